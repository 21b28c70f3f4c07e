#!/bin/bash
# A/B of the time-tiled MAC variants on the GPU box: parity first, then ms/block per T and S.
python -m pytest tests/test_engine_gpu.py -m gpu -x -q 2>&1 | tail -2
run() { python bench.py --steps 50 --no-cpu-baseline --skip-e2e --blocks-per-step $1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']; T=d['config']['blocks_per_step']; print('$2 T',T,'S','${FCV_TT_S:-def}','xRT', round(d['value']), 'ms/block', round(d['ms_per_step']/T,4), {a: round(b/T,4) for a,b in k.items()})"; }
for t in 4 8; do
  FCV_MAC_V1=1 run $t v1
  for s in 2 4; do FCV_TT_S=$s run $t v2; done
done
