#!/bin/bash
# TMA-staged MAC against the register-pipelined one: parity of the tiled paths, then timing
export FCV_MAC_TMA=1
timeout 200 python -m pytest tests/test_engine_gpu.py tests/test_scale_properties_gpu.py -m gpu -x -q 2>&1 | tail -3
unset FCV_MAC_TMA
run() { timeout 120 python bench.py --steps 50 --no-cpu-baseline --skip-e2e --blocks-per-step $1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']; T=d['config']['blocks_per_step']; print('$2 T',T,'xRT', round(d['value']), 'ms/block', round(d['ms_per_step']/T,4), {a: round(b/T,4) for a,b in k.items()})"; }
run 8 reg; FCV_MAC_TMA=1 run 8 tma
run 4 reg; FCV_MAC_TMA=1 run 4 tma_s4; FCV_MAC_TMA=2 run 4 tma_s2
