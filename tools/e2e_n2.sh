#!/bin/bash
# end-to-end loop on 2 GPUs under different settings
for cfg in "$@"; do
  env $cfg timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 100 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']
print('%-24s N=2 e2e xRT %8.0f ms/step %.3f  f32 wire ms/step %.3f' % ('$cfg' or 'default', e['value'], e['ms_per_step'], e['f32_wire']['ms_per_step']))"
done
