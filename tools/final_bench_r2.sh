#!/bin/bash
# Round 2, final build on ONE B200: the default bench line as the driver runs it, every other workload's line, the
# reference arm, the copy probe, ncu launch lists and full captures of the kernels the lines are about.
tag=${1:-r2_final}
O=gpurun_out/$tag
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/gpu.txt; nproc >> $O/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_cfg5_as_driver.json 2> $O/bench_cfg5_as_driver.err
python bench.py > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "rc=$?" >> $O/bench_cfg5.err
for w in cfg2 cfg3 cfg4a cfg4b cfg4c cfg6ch cfg2p; do
  timeout 600 python bench.py --workload $w --steps 100 > $O/bench_$w.json 2> $O/bench_$w.err
done
timeout 600 python bench.py --impl reference --steps 10 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
./tools/copy_probe 2 > $O/copy_probe_2g.txt 2>&1
./tools/copy_probe 11.7 > $O/copy_probe_11g.txt 2>&1
# launch lists (ncu serialises launches: per-launch times are cold and isolated)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_cfg5.csv \
    python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_cfg3.csv \
    python bench.py --workload cfg3 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > /dev/null 2>&1
# full captures
cap() {  # workload kernel-regex name [keep]: full capture, summarised HERE (the reports are 30 MB each, gpurun_out/ takes 64)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s 4 -c 1 -f -o $O/$3 \
      python bench.py --workload $1 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > $O/ncu_$3.log 2>&1
  python tools/ncu_summary.py $O/$3.ncu-rep $O/$3_ncu_full.txt $1 > /dev/null 2>&1
  [ -z "$4" ] && rm -f $O/$3.ncu-rep
}
cap cfg5 fused_tick fused_tick_cfg5 keep
cap cfg2 fused_tick fused_tick_cfg2
cap cfg4a fused_tick fused_tick_cfg4a
cap cfg6ch any_tick any_tick_cfg6ch
cap cfg4b mix8to2 mix8to2_cfg4b
cap cfg4c mix8to2 mix8to2_cfg4c
cap cfg3 span_tick span_tick_cfg3
cap cfg2p fused_tick fused_tick_cfg2p
cp $O/traffic.json $O/traffic_from_this_run.json 2>/dev/null
ls -la $O; du -sh $O
