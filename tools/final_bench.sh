#!/bin/bash
# every workload's bench line (+ the reference arm) into gpurun_out/bench_<tag>_<workload>.json
tag=${1:-r1b}
python bench.py --verify > gpurun_out/bench_${tag}_cfg2.json 2> gpurun_out/bench_${tag}_cfg2.err
for w in cfg3 cfg4a cfg4b cfg5 cfg6ch; do
  python bench.py --workload $w --steps 50 --verify > gpurun_out/bench_${tag}_$w.json 2> gpurun_out/bench_${tag}_$w.err
done
python bench.py --impl reference --steps 10 > gpurun_out/bench_${tag}_reference_arm.json 2> gpurun_out/bench_${tag}_reference_arm.err
for f in gpurun_out/bench_${tag}_*.json; do python - "$f" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d.get("roofline") or {}
e = d.get("e2e") or {}
print(sys.argv[1].split("/")[-1], round(d["value"]), round(d["ms_per_step"], 4), r.get("frac") and round(r["frac"], 4), r.get("kernel"),
      "e2e", e.get("value") and round(e["value"]), (e.get("parity_spot_check") or "")[:40], "cpu", d.get("cpu_baseline") and round(d["cpu_baseline"]["value"]))
PY
done
