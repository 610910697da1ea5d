#!/bin/bash
# Round-2 GPU session 20: the final build -- everything tools/final_bench_r2.sh measures, the bounds-checking suite,
# the latency probe.
bash tools/final_bench_r2.sh r2_final2
O=gpurun_out/r2_final2
CMGPU_LIB=$PWD/libcoolmic-dsp_b200/lib/libcoolmic_b200_dbg.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mix.py tests/test_gpu_post.py -q > $O/pytest_boundscheck.log 2>&1; echo "rc=$?" >> $O/pytest_boundscheck.log
./tools/latency_probe > $O/latency_probe.txt 2>&1
tail -3 $O/pytest_gpu.log $O/pytest_boundscheck.log
