#!/bin/bash
# Round-2 GPU session 13: whole GPU suite with the store-free pass-through kernels, the pass-through line again.
O=gpurun_out/s13
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 300 python bench.py --workload cfg5pt --steps 50 --no-e2e --no-cpu-baseline > $O/bench_cfg5pt.json 2> $O/bench_cfg5pt.err
timeout 300 python bench.py --steps 50 --no-cpu-baseline > $O/bench_cfg5.json 2> $O/bench_cfg5.err
CMGPU_LIB=$PWD/libcoolmic-dsp_b200/lib/libcoolmic_b200_dbg.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mix.py -q > $O/pytest_boundscheck.log 2>&1; echo "rc=$?" >> $O/pytest_boundscheck.log
ls -la $O
