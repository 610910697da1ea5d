#!/bin/bash
# Round-2 GPU session 1: full GPU test suite, first bench lines of the new default (cfg5) at N=1 and N=2
# through the C NCCL path, item-size sweep on cfg5, link probe A/B, ncu launch list + one full capture.
O=gpurun_out/s1
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
nproc >> $O/gpus.txt; free -g >> $O/gpus.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 600 python bench.py --steps 50 > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "rc=$?" >> $O/bench_cfg5.err
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -ge 2 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps 50 > $O/bench_cfg5_n2.json 2> $O/bench_cfg5_n2.err; echo "rc=$?" >> $O/bench_cfg5_n2.err
fi
timeout 300 python bench.py --workload cfg2 --steps 100 --no-cpu-baseline > $O/bench_cfg2.json 2> $O/bench_cfg2.err
for v in 768 1024 1536 2048 3072; do
  CMGPU_ITEM_VECS=$v timeout 200 python bench.py --steps 30 --no-e2e --no-cpu-baseline --no-extras > $O/sweep_cfg5_$v.json 2>/dev/null
done
timeout 200 python bench.py --workload cfg5x --steps 30 --no-e2e --no-cpu-baseline --no-extras > $O/bench_cfg5x.json 2>/dev/null
for m in copy transform meter; do
  timeout 200 python bench.py --steps 30 --mode $m --no-e2e --no-cpu-baseline --no-extras > $O/mode_cfg5_$m.json 2>/dev/null
done
python - > $O/link_probe.txt 2>&1 <<'P'
from __graft_entry__ import load_package
cm = load_package()
for wc in (False, True):
    for nb in (16 << 20, 256 << 20):
        print("wc", wc, nb, cm.link_probe(0, nb, reps=max(4, (2 << 30) // nb), wc=wc))
P
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches_cfg5.csv \
    python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > $O/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_tick -s 4 -c 1 -f -o $O/fused_tick_cfg5 \
    python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > $O/ncu_full.log 2>&1
ls -la $O
