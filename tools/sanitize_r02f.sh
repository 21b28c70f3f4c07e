#!/bin/bash
# sanitizer passes over the kernels changed at the end of round 2: mac_tma_kernel (straight-line consumer rows, producer
# passes of 4) and inv13_stream_kernel<.., TM = true> (tensor-memory scratch), on the T = 8 batch tests
OUT=gpurun_out; CS="compute-sanitizer --error-exitcode 7 --print-limit 10"
for tool in memcheck synccheck; do
  timeout -s KILL 240 $CS --tool $tool python -m pytest tests/test_engine_gpu.py -x -q --timeout 200 -k "tiled or time_tiled or blocks_per_step or T8 or batch" > $OUT/san4_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a $OUT/san4_summary.txt
  grep -E "passed|failed|ERROR SUMMARY" $OUT/san4_$tool.log | tail -3 | tee -a $OUT/san4_summary.txt
done
