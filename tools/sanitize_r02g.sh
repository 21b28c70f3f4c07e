#!/bin/bash
# synccheck over the time-tiled batch tests with the tensor-memory inverse kernel off (FCV_INV_TMEM=0): the control for
# the "Missing init" report that the tool attaches to tcgen05.alloc's result word (tools/sanitize_r02f.sh)
OUT=gpurun_out; CS="compute-sanitizer --error-exitcode 7 --print-limit 10"
FCV_INV_TMEM=0 timeout -s KILL 240 $CS --tool synccheck python -m pytest tests/test_engine_gpu.py -x -q --timeout 200 -k "time_tiled" > $OUT/san5_synccheck_notmem.log 2>&1
echo "synccheck FCV_INV_TMEM=0 rc=$?" | tee -a $OUT/san5_summary.txt
grep -E "passed|failed|ERROR SUMMARY" $OUT/san5_synccheck_notmem.log | tail -3 | tee -a $OUT/san5_summary.txt
