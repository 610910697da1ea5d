#!/bin/bash
# Round-2 GPU session 9: the single tick of the small-buffer regime -- kernel time in isolation, A/B through the span kernel.
O=gpurun_out/s9
mkdir -p $O
for i in 1 2; do
timeout 300 python bench.py --workload cfg3 --steps 50 --no-e2e --no-cpu-baseline > $O/bench_cfg3_$i.json 2>/dev/null
CMGPU_SPAN_SINGLE=1 timeout 300 python bench.py --workload cfg3 --steps 50 --no-e2e --no-cpu-baseline > $O/bench_cfg3_spansingle_$i.json 2>/dev/null
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_cfg3.csv \
    python bench.py --workload cfg3 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > /dev/null 2>&1
CMGPU_SPAN_SINGLE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_cfg3_spansingle.csv \
    python bench.py --workload cfg3 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > /dev/null 2>&1
CMGPU_SPAN_SINGLE=1 CMGPU_SPAN_BY_STREAM=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "small_buffer or full_blocks or random_all or stream_major or overlapping" > $O/pytest_spansingle.log 2>&1; echo "rc=$?" >> $O/pytest_spansingle.log
ls -la $O
