#!/bin/bash
# final verification of the round on one GPU: all GPU tests, smoke, default bench (timed), launch list, full captures
tag=${1:-r02c}
(time timeout -s KILL 600 python -m pytest tests -m gpu -x -q --timeout 300) > gpurun_out/${tag}_gputests.log 2>&1; tail -4 gpurun_out/${tag}_gputests.log
(time timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/${tag}_smoke.log 2>&1; tail -5 gpurun_out/${tag}_smoke.log
(time timeout -s KILL 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err); cut -c1-400 gpurun_out/${tag}_bench.json
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --skip-e2e --no-configs --no-library > gpurun_out/${tag}_launches.log 2>&1
bash tools/ncu_full.sh ${tag}
