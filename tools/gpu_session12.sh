#!/bin/bash
# Round-2 GPU session 12: pass-through (meter-only, read-only) tick on the device; single tick with the early tick read.
O=gpurun_out/s12
mkdir -p $O
timeout 300 python bench.py --workload cfg5pt --steps 50 --no-e2e --no-cpu-baseline > $O/bench_cfg5pt.json 2> $O/bench_cfg5pt.err
timeout 300 python bench.py --workload cfg3 --steps 50 --no-e2e --no-cpu-baseline > $O/bench_cfg3.json 2> $O/bench_cfg3.err
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "small_buffer or full_blocks or random_all or cycle_equals or appendix or sine" > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_tick -s 4 -c 1 -f -o $O/fused_tick_cfg5pt \
    python bench.py --workload cfg5pt --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > $O/ncu.log 2>&1
python tools/ncu_summary.py $O/fused_tick_cfg5pt.ncu-rep $O/fused_tick_cfg5pt_ncu_full.txt > /dev/null 2>&1
rm -f $O/fused_tick_cfg5pt.ncu-rep
ls -la $O
