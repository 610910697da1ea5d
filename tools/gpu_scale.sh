#!/bin/bash
# N-GPU run as the driver launches it (torchrun spawns the ranks; the script itself uses no torch): the default
# strong-scaling workload on every GPU of the box, plus the NCCL gather test over all of them.
N=${1:-8}
O=gpurun_out/scale
mkdir -p $O
nvidia-smi -L > $O/gpus_n$N.txt; nproc >> $O/gpus_n$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 \
    bench.py --gpus $N --steps 100 > $O/bench_cfg5_n$N.json 2> $O/bench_cfg5_n$N.err; echo "rc=$?" >> $O/bench_cfg5_n$N.err
timeout 600 python -m pytest tests/test_gpu_post.py -q -k nccl > $O/pytest_gather_n$N.log 2>&1; echo "rc=$?" >> $O/pytest_gather_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus $N --steps 10 --impl reference > $O/bench_ref_n$N.json 2> $O/bench_ref_n$N.err
tail -3 $O/bench_cfg5_n$N.err
