#!/bin/bash
# parity tests, then ms/block per kernel at the given blocks-per-step values (default 4 8)
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for t in ${TS:-4 8}; do python bench.py --steps 50 --no-cpu-baseline --skip-e2e --blocks-per-step $t | python -c "import json,sys; d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']; T=d['config']['blocks_per_step']; print('T',T,'xRT', round(d['value']), 'ms/block', round(d['ms_per_step']/T,4), {a: round(b/T,4) for a,b in k.items()})"; done
