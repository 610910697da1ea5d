#!/bin/bash
# A/B on one box: the fully metered 8 -> 2 kernel as CTAs of 8 warps at 128 registers (2 per SM, lib/exp/mix_256x2.so)
# against CTAs of 4 warps at 96 registers (5 per SM, the build in lib/), then the downmix and shim tests on the new build.
O=gpurun_out/r2_mix_ab
mkdir -p $O
for rep in 1 2; do
  for v in old new; do
    [ $v = old ] && export CMGPU_LIB=$PWD/libcoolmic-dsp_b200/lib/exp/mix_256x2.so || unset CMGPU_LIB
    timeout 40 python bench.py --workload cfg4b --steps 100 --no-e2e --no-cpu-baseline --no-extras > $O/bench_cfg4b_${v}_$rep.json 2> $O/bench_cfg4b_${v}_$rep.err
  done
done
unset CMGPU_LIB
timeout 60 python -m pytest tests/test_gpu_mix.py -q -x > $O/pytest_mix.log 2>&1; echo "pytest rc=$?" >> $O/pytest_mix.log
timeout 40 python bench.py --workload cfg4b --steps 100 --no-cpu-baseline > $O/bench_cfg4b.json 2> $O/bench_cfg4b.err
tail -2 $O/pytest_mix.log
python - <<'P'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_mix_ab/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], d["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("sustained", {}).get("frac"), (d.get("e2e") or {}).get("parity_spot_check", "")[:30])
    except Exception as e:
        print(f, "no line:", e)
P
