#!/bin/bash
# sanitizer passes over what changed after r02b: stride addressing of the batch kernels, the stereo-pair inverse kernel
OUT=gpurun_out; CS="compute-sanitizer --error-exitcode 7 --print-limit 10"
SEL='stereo_pair_inverse or time_tiled_batch_equals_block_by_block or batch_matches_single'
for tool in memcheck synccheck racecheck; do
  timeout 900 $CS --tool $tool python -m pytest tests/test_engine_gpu.py -x -q --timeout 800 -k "$SEL" > $OUT/san2_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a $OUT/san2_summary.txt
  grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY" $OUT/san2_$tool.log | tail -3 | tee -a $OUT/san2_summary.txt
  grep -E "Race reported|and (Read|Write) access" $OUT/san2_$tool.log | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | sort -rn | head -6 | tee -a $OUT/san2_summary.txt
done
