#!/bin/bash
# Round-2 GPU session 7: the cp.async-ring any_tick -- parity, bounds-checking build, bench A/B by item size, ncu summary.
O=gpurun_out/s7
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_post.py -q \
   -k "random_all_channel or exhaustive or planar or cycle_equals or transform_only or separate_out or full_blocks or gather or results_batch or passthrough" > $O/pytest_any.log 2>&1; echo "rc=$?" >> $O/pytest_any.log
CMGPU_LIB=$PWD/libcoolmic-dsp_b200/lib/libcoolmic_b200_dbg.so timeout 900 python -m pytest tests/test_gpu_parity.py -q \
   -k "random_all_channel or planar or full_blocks" > $O/pytest_any_boundscheck.log 2>&1; echo "rc=$?" >> $O/pytest_any_boundscheck.log
timeout 300 python bench.py --workload cfg6ch --steps 100 --no-e2e --no-cpu-baseline > $O/bench_cfg6ch.json 2> $O/bench_cfg6ch.err
for v in 2048 4096; do
  CMGPU_ITEM_VECS=$v timeout 300 python bench.py --workload cfg6ch --steps 100 --no-e2e --no-cpu-baseline --no-extras > $O/bench_cfg6ch_$v.json 2>/dev/null
done
timeout 300 python bench.py --workload cfg6ch --steps 100 --mode copy --no-e2e --no-cpu-baseline --no-extras > $O/bench_cfg6ch_copy.json 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:any_tick -s 4 -c 1 -f -o $O/any_tick_cfg6ch \
    python bench.py --workload cfg6ch --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > $O/ncu_any.log 2>&1
python tools/ncu_summary.py $O/any_tick_cfg6ch.ncu-rep $O/any_tick_cfg6ch_ncu_full.txt cfg6ch > /dev/null 2>&1
rm -f $O/any_tick_cfg6ch.ncu-rep
ls -la $O
