#!/bin/bash
# Round-2 GPU session 3: dynamic work claims under overlapping launches, the stream-major span kernel, the NCCL gather
# test with its failure detail, the bounds-checking build over a test subset.
O=gpurun_out/s3
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_post.py -q -k "nccl" > $O/pytest_gather.log 2>&1; echo "rc=$?" >> $O/pytest_gather.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "stream_major or cycle_equals or small_buffer or overlapping or tick_numbering" > $O/pytest_span.log 2>&1; echo "rc=$?" >> $O/pytest_span.log
timeout 600 python bench.py --steps 50 > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "rc=$?" >> $O/bench_cfg5.err
CMGPU_STATIC_ITEMS=1 timeout 300 python bench.py --steps 50 --no-e2e --no-cpu-baseline > $O/bench_cfg5_static.json 2>/dev/null
for w in cfg2 cfg3 cfg4a; do
  timeout 300 python bench.py --workload $w --steps 50 --no-e2e --no-cpu-baseline > $O/bench_$w.json 2> $O/bench_$w.err
done
CMGPU_SPAN_BY_TICK=1 timeout 300 python bench.py --workload cfg3 --steps 50 --no-e2e --no-cpu-baseline > $O/bench_cfg3_bytick.json 2>/dev/null
CMGPU_STATIC_ITEMS=1 timeout 300 python bench.py --workload cfg2 --steps 50 --no-e2e --no-cpu-baseline > $O/bench_cfg2_static.json 2>/dev/null
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -ge 2 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
      bench.py --gpus 2 --steps 50 --no-cpu-baseline > $O/bench_cfg5_n2.json 2> $O/bench_cfg5_n2.err
fi
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_all.log 2>&1; echo "pytest rc=$?" >> $O/pytest_all.log
CMGPU_LIB=$PWD/libcoolmic-dsp_b200/lib/libcoolmic_b200_dbg.so timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mix.py -q \
   -k "not tma" > $O/pytest_boundscheck.log 2>&1; echo "rc=$?" >> $O/pytest_boundscheck.log
CMGPU_LIB=$PWD/libcoolmic-dsp_b200/lib/libcoolmic_b200_dbg.so python -c "
from __graft_entry__ import load_package
cm = load_package(); print('violations:', cm.lib().cmgpu_debug_violations())" >> $O/pytest_boundscheck.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:span_tick -s 3 -c 1 -f -o $O/span_tick_cfg3 \
    python bench.py --workload cfg3 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > $O/ncu_cfg3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/launches_cfg3.csv \
    python bench.py --workload cfg3 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > /dev/null 2>&1
ls -la $O
