#!/bin/bash
# which kernels slow each other down? runs the device loop with subsets of the three kernels
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.sw_power_cap --format=csv,noheader -lms 100 > gpurun_out/clk.csv &
SMI=$!
for o in 7 1 4 2 5 3 6; do FCV_ONLY=$o python bench.py --steps 1500 --no-cpu-baseline --skip-e2e --blocks-per-step 4 | python -c "import json,sys; d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']; T=d['config']['blocks_per_step']; print('only',$o, 'ms/block', round(d['ms_per_step']/T,4), {a: round(b/T,4) for a,b in k.items()}, d['clocks'])"; done
kill $SMI
sort gpurun_out/clk.csv | uniq -c | sort -rn | head -12
