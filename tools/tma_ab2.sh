#!/bin/bash
run() { timeout 120 python bench.py --steps 50 --no-cpu-baseline --skip-e2e --blocks-per-step $1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']; T=d['config']['blocks_per_step']; print('$2 T',T,'xRT', round(d['value']), 'ms/block', round(d['ms_per_step']/T,4), {a: round(b/T,4) for a,b in k.items()})"; }
FCV_MAC_TMA=3 timeout 200 python -m pytest tests/test_engine_gpu.py -m gpu -x -q -k "tiled" 2>&1 | tail -2
run 8 reg; FCV_MAC_TMA=1 run 8 tma_s2; FCV_MAC_TMA=3 run 8 tma_s1
