#!/bin/bash
# Round-2 GPU session 10: the whole GPU suite on a 2-GPU box (gather over 2 ranks, new stress / adopt tests), the 2-GPU line.
O=gpurun_out/s10
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29530 \
    bench.py --gpus 2 --steps 100 > $O/bench_cfg5_n2.json 2> $O/bench_cfg5_n2.err; echo "rc=$?" >> $O/bench_cfg5_n2.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
ls -la $O
