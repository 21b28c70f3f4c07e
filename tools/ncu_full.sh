#!/bin/bash
# --set full captures (tag = $1) of the kernels of the default step (warm-up launches skipped)
tag=${1:-r02}
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --skip-e2e"
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:mac_tma -s 3 -c 1 -f -o gpurun_out/${tag}_mac $B > gpurun_out/${tag}_mac.log 2>&1
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k "regex:fwd13_stream|inv13|dcny" -s 9 -c 3 -f -o gpurun_out/${tag}_fft $B > gpurun_out/${tag}_fft.log 2>&1
# the MIMO case: dense 6x6 (X rows shared by six outputs: L2 hit rate of the output-fastest item order)
timeout -s KILL 200 ncu --set full --clock-control none -k regex:mac_tma -s 3 -c 1 -f -o gpurun_out/${tag}_mac_dense $B --workload surround51_dense --streams 256 > gpurun_out/${tag}_mac_dense.log 2>&1
ls -la gpurun_out | grep ${tag}
