#!/bin/bash
# sanitizer passes over the per-file path's new kernels: mac_group_kernel, fwd13_pair_kernel (cluster + distributed
# shared memory), inv13_pair_kernel on the per-file path, host_copy_out with one system fence
OUT=gpurun_out; CS="compute-sanitizer --error-exitcode 7 --print-limit 10"
for tool in memcheck synccheck racecheck; do
  timeout 1200 $CS --tool $tool python -m pytest tests/test_coalesce_gpu.py -x -q --timeout 1100 -k "one_launch_group or integer_wire or concurrent_threads" > $OUT/san3_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a $OUT/san3_summary.txt
  grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY" $OUT/san3_$tool.log | tail -3 | tee -a $OUT/san3_summary.txt
  grep -E "Race reported|and (Read|Write) access" $OUT/san3_$tool.log | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | sort -rn | head -6 | tee -a $OUT/san3_summary.txt
done
