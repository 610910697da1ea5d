#!/bin/bash
# Round-2 GPU session 2: whole GPU suite (incl. full-size + exhaustive-on-device + NCCL gather tests), copy probe at two
# footprints, A/B of dynamic work claims, bench lines of every workload, compute-sanitizer on a subset.
O=gpurun_out/s2
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
./tools/copy_probe 2 > $O/copy_probe_2g.txt 2>&1
./tools/copy_probe 11.7 > $O/copy_probe_11g.txt 2>&1
timeout 600 python bench.py --steps 50 > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "rc=$?" >> $O/bench_cfg5.err
CMGPU_STATIC_ITEMS=1 timeout 300 python bench.py --steps 50 --no-e2e --no-cpu-baseline > $O/bench_cfg5_static.json 2>/dev/null
for w in cfg2 cfg3 cfg4a cfg6ch cfg4b cfg2p cfg5x; do
  timeout 300 python bench.py --workload $w --steps 30 --no-e2e --no-cpu-baseline > $O/bench_$w.json 2> $O/bench_$w.err
done
CMGPU_STATIC_ITEMS=1 timeout 300 python bench.py --workload cfg2 --steps 30 --no-e2e --no-cpu-baseline > $O/bench_cfg2_static.json 2>/dev/null
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -ge 2 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus 2 --steps 50 --no-cpu-baseline > $O/bench_cfg5_n2.json 2> $O/bench_cfg5_n2.err
fi
# sanitizers on a reduced subset (kernels run 10-100x slower under them)
SUB='test_random_all_channel_counts or test_small_buffer_regime or test_overlapping_ticks or test_cycle_equals_individual or test_tma_and_ldg or test_full_blocks'
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "$SUB" > $O/sanitizer_memcheck.log 2>&1; echo "rc=$?" >> $O/sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "test_small_buffer_regime or test_overlapping_ticks or test_tma_and_ldg" > $O/sanitizer_racecheck.log 2>&1; echo "rc=$?" >> $O/sanitizer_racecheck.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_post.py tests/test_gpu_mix.py -q -x -k "not nccl" > $O/sanitizer_memcheck_post.log 2>&1; echo "rc=$?" >> $O/sanitizer_memcheck_post.log
ls -la $O
