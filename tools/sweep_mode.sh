#!/bin/bash
# library build x bench mode sweep: tools/sweep_mode.sh "<libs>" "<modes>" <workload> [steps]
libs=${1:-"cur"}; modes=${2:-"fused"}; w=${3:-cfg2}; steps=${4:-40}
for lib in $libs; do
  if [ $lib = cur ]; then unset CMGPU_LIB; else export CMGPU_LIB=$PWD/libcoolmic-dsp_b200/lib/exp/$lib.so; fi
  for m in $modes; do timeout 120 python bench.py --workload $w --mode $m --steps $steps --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', '$w', '$m', round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['roofline']['kernel'])"; done
done
