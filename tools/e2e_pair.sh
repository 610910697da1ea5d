#!/bin/bash
# two e2e_probe.py processes started at the same instant, one per GPU: tools/e2e_pair.sh <visible: each|all> [probe flags]
mode=$1; shift
T=$(python -c "import time; print(time.time()+30)")
if [ "$mode" = each ]; then
  (CUDA_VISIBLE_DEVICES=0 python tools/e2e_probe.py --start-at $T "$@" > gpurun_out/pair0.txt &)
  CUDA_VISIBLE_DEVICES=1 python tools/e2e_probe.py --start-at $T "$@"
else
  (python tools/e2e_probe.py --device 0 --start-at $T "$@" > gpurun_out/pair0.txt &)
  python tools/e2e_probe.py --device 1 --start-at $T "$@"
fi
sleep 2; cat gpurun_out/pair0.txt
