#!/bin/bash
# N-GPU timing of data-parallel variants: tools/dp_variants.sh <ngpus> <configs> "VAR=.. VAR=.." ...
N=$1; CFG=$2; shift 2
for v in "$@"; do
  env $v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 30 --warmup 5 --configs "$CFG" 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][-1])
pc = ' '.join('%s train %.4f e2e %.4f' % (k, e['train']['ms_per_step'], e['e2e']['train']['ms_per_step']) for k, e in d['per_config'].items() if 'train' in e)
print('$v', 'N=%d' % d['n_gpus'], 'C2a train %.4f ms e2e %.4f ms' % (d['ms_per_step'], d['e2e']['ms_per_step']), pc)
"
done
