#!/bin/bash
# Round 2, last GPU call: the downmix kernels after the batch metering / 32-bit quotient change (cfg4b, cfg4c lines and
# full captures), the whole GPU suite and the default line as the driver runs it, on ONE B200.
O=gpurun_out/r2_final3
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/gpu.txt
timeout 150 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
for w in cfg4b cfg4c; do
  timeout 60 python bench.py --workload $w --steps 100 > $O/bench_$w.json 2> $O/bench_$w.err
done
timeout 90 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_cfg5_as_driver.json 2> $O/bench_cfg5_as_driver.err
cap() {
  timeout 90 ncu --set full --clock-control none --import-source on -k regex:$2 -s 4 -c 1 -f -o $O/$3 \
      python bench.py --workload $1 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > $O/ncu_$3.log 2>&1
  python tools/ncu_summary.py $O/$3.ncu-rep $O/$3_ncu_full.txt $1 > /dev/null 2>&1
  rm -f $O/$3.ncu-rep
}
cap cfg4b mix8to2 mix8to2_cfg4b
cap cfg4c mix8to2 mix8to2_cfg4c
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
tail -2 $O/pytest_gpu.log; cat $O/smoke.log | tail -2
python - <<'P'
import json
for w in ("cfg4b", "cfg4c", "cfg5_as_driver"):
    try:
        d = json.loads(open(f"gpurun_out/r2_final3/bench_{w}.json").read().strip().splitlines()[-1])
        print(w, d["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("sustained", {}).get("frac"), d["e2e"]["value"], d["e2e"].get("parity_spot_check", "")[:40])
    except Exception as e:
        print(w, "no line:", e)
P
