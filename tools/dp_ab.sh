#!/bin/bash
# N-GPU A/B of the gradient all-reduce paths on the headline step: tools/dp_ab.sh <ngpus> <steps> <rounds>
# alternates NCCL (NF_B200_DP_P2P_MAX_BYTES=0) and the peer-memory kernel so both see the same box state
N=$1; STEPS=${2:-100}; ROUNDS=${3:-2}
for r in $(seq 1 $ROUNDS); do
  for v in 0 16777216; do
    NF_B200_DP_P2P_MAX_BYTES=$v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps $STEPS --warmup 10 --configs "" --tarflow 0 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('p2p_max_bytes=$v', 'N=%d' % d['n_gpus'], 'C2a train %.4f ms (median %.4f) e2e %.4f ms' % (d['ms_per_step'], d['train_ms_median_max'][0], d['e2e']['ms_per_step']))
"
  done
done
