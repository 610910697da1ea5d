#!/bin/bash
# Round-2 GPU session 11: the channel counts no BASELINE config names (1, 4, 16 with long blocks; 3, 7, 12 through any_tick).
O=gpurun_out/s11
mkdir -p $O
for w in cfg1ch cfg4ch cfg16ch; do
  timeout 300 python bench.py --workload $w --steps 50 --no-e2e --no-cpu-baseline > $O/bench_$w.json 2> $O/bench_$w.err
done
ls -la $O
