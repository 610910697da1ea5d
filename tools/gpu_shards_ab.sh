#!/bin/bash
# A/B of the end-to-end leg's shards on N GPUs: in proportion to each rank's host-link share (default) vs equal.
N=${1:-2}
O=gpurun_out/shards
mkdir -p $O
for v in ${VARIANTS:-prop equal prop equal}; do
  extra=""; [ $v = equal ] && extra="--e2e-equal-shards"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29530 \
      bench.py --gpus $N --steps 20 --no-cpu-baseline --no-extras $extra > $O/bench_n${N}_$v.json 2> $O/bench_n${N}_$v.err; echo "rc=$?" >> $O/bench_n${N}_$v.err
  python - <<P
import json
d=json.load(open("$O/bench_n${N}_$v.json")); e=d["e2e"]
print("$v", d["ms_per_step"], "e2e", round(e["value"]), "GB/s each way", round(e["gbs_each_way_all_ranks"],1), e.get("shards",{}).get("streams_per_rank"), e.get("shards",{}).get("link_gbs_per_rank"), e["parity_spot_check"][:12])
P
done
if [ -z "$SKIP_N1" ]; then
  python bench.py --steps 20 --no-cpu-baseline --no-extras > $O/bench_n1.json 2> $O/bench_n1.err
  python -c "
import json; d=json.load(open('$O/bench_n1.json')); print('n1', d['ms_per_step'], d['e2e']['value'], d['e2e']['parity_spot_check'][:12])"
fi
