#!/bin/bash
python -m pytest tests/test_engine_gpu.py -m gpu -x -q 2>&1 | tail -2
for t in ${TS:-4 8}; do for a in ${AH:-0 148 296 592 1184}; do FCV_FFT_AHEAD=$a python bench.py --steps 50 --no-cpu-baseline --skip-e2e --blocks-per-step $t | python -c "import json,sys; d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']; T=d['config']['blocks_per_step']; print('T',T,'ahead',$a,'xRT', round(d['value']), 'ms/block', round(d['ms_per_step']/T,4), {a: round(b/T,4) for a,b in k.items()})"; done; done
