#!/bin/bash
# per-item size sweep (CMGPU_ITEM_VECS tuning hook): tools/sweep_item.sh "<libs>" "<sizes>" <workload> [steps]
libs=${1:-"cur"}; sizes=${2:-"2048"}; w=${3:-cfg2}; steps=${4:-40}
for lib in $libs; do
  if [ $lib = cur ]; then unset CMGPU_LIB; else export CMGPU_LIB=$PWD/libcoolmic-dsp_b200/lib/exp/$lib.so; fi
  for v in $sizes; do CMGPU_ITEM_VECS=$v python bench.py --workload $w --steps $steps --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', '$w', 'item_vecs=$v', round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])"; done
done
unset CMGPU_LIB
