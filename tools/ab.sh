#!/bin/bash
# A/B of library builds on the GPU box: tools/ab.sh "<lib names under lib/exp, or 'cur'>" "<workloads>" [reps] [steps]
libs=${1:-"old cur"}; wls=${2:-"cfg2 cfg5"}; reps=${3:-2}; steps=${4:-40}
for rep in $(seq $reps); do for lib in $libs; do
  if [ $lib = cur ]; then unset CMGPU_LIB; else export CMGPU_LIB=$PWD/libcoolmic-dsp_b200/lib/exp/$lib.so; fi
  for w in $wls; do python bench.py --workload $w --steps $steps --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', '$w', round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])"; done
done; done
unset CMGPU_LIB
