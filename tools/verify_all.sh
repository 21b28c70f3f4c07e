#!/bin/bash
# round-end style verification on one GPU: parity tests, default bench, launch list and
# --set full captures (tag = $1) of the three kernels of the default step
tag=${1:-r01h}
(timeout -s KILL 200 python -m pytest tests -m gpu -x -q) > gpurun_out/gputests.log 2>&1; tail -2 gpurun_out/gputests.log
timeout -s KILL 300 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cut -c1-300 gpurun_out/bench_default.json
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --skip-e2e > gpurun_out/${tag}_launches.log 2>&1
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:mac_tma -s 10 -c 1 -f -o gpurun_out/${tag}_mac python bench.py --steps 3 --warmup 3 --no-cpu-baseline --skip-e2e > gpurun_out/${tag}_mac.log 2>&1
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k "regex:fwd13_stream|inv13|dcny" -s 30 -c 3 -f -o gpurun_out/${tag}_fft python bench.py --steps 3 --warmup 3 --no-cpu-baseline --skip-e2e > gpurun_out/${tag}_fft.log 2>&1
ls gpurun_out | grep ${tag}
