#!/bin/bash
# Round-2 GPU session 4: cp.async-staged span kernel, dynamic claims in the any-channel and downmix kernels,
# gather test, bounds-checking build over the parity + downmix suites, then the whole GPU suite.
O=gpurun_out/s4
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_post.py -q -k "nccl" > $O/pytest_gather.log 2>&1; echo "rc=$?" >> $O/pytest_gather.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -q -k "stream_major or cycle_equals or small_buffer or config3" > $O/pytest_span.log 2>&1; echo "rc=$?" >> $O/pytest_span.log
for w in cfg3 cfg6ch cfg4b cfg2p; do
  timeout 300 python bench.py --workload $w --steps 50 --no-e2e --no-cpu-baseline > $O/bench_$w.json 2> $O/bench_$w.err
done
CMGPU_SPAN_BY_TICK=1 timeout 300 python bench.py --workload cfg3 --steps 50 --no-e2e --no-cpu-baseline > $O/bench_cfg3_bytick.json 2>/dev/null
for v in 4096 7000; do
  CMGPU_ITEM_VECS=$v timeout 300 python bench.py --workload cfg6ch --steps 50 --no-e2e --no-cpu-baseline --no-extras > $O/bench_cfg6ch_$v.json 2>/dev/null
done
CMGPU_STATIC_ITEMS=1 timeout 300 python bench.py --workload cfg6ch --steps 50 --no-e2e --no-cpu-baseline --no-extras > $O/bench_cfg6ch_static.json 2>/dev/null
CMGPU_STATIC_ITEMS=1 timeout 300 python bench.py --workload cfg4b --steps 50 --no-e2e --no-cpu-baseline --no-extras > $O/bench_cfg4b_static.json 2>/dev/null
timeout 300 python bench.py --workload cfg4c --steps 50 --no-e2e --no-cpu-baseline --no-extras > $O/bench_cfg4c.json 2>/dev/null
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_all.log 2>&1; echo "pytest rc=$?" >> $O/pytest_all.log
CMGPU_LIB=$PWD/libcoolmic-dsp_b200/lib/libcoolmic_b200_dbg.so timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mix.py -q \
    > $O/pytest_boundscheck.log 2>&1; echo "rc=$?" >> $O/pytest_boundscheck.log
CMGPU_LIB=$PWD/libcoolmic-dsp_b200/lib/libcoolmic_b200_dbg.so python -c "
from __graft_entry__ import load_package
cm = load_package(); print('violations counted by the bounds-checking build:', cm.lib().cmgpu_debug_violations())" >> $O/pytest_boundscheck.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:span_tick -s 3 -c 1 -f -o $O/span_tick_cfg3 \
    python bench.py --workload cfg3 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > $O/ncu_cfg3.log 2>&1
ls -la $O
