#!/bin/bash
# Race evidence for the round (run on a GPU box; summaries go to gpurun_out/, copied into profiles/):
#   1. compute-sanitizer racecheck + synccheck on the kernels with shared-memory hand-offs: the
#      TMA / mbarrier ring of mac_tma_kernel (T = 4, 8), mac_tt (T = 2), the T = 1 kernel, the FFT
#      passes of every block size, the single-stream path with the host copy-out;
#   2. ThreadSanitizer over the host-side concurrency (tools/tsan_host.cc).
set -u
OUT=gpurun_out
mkdir -p $OUT
CS="compute-sanitizer --error-exitcode 7 --print-limit 20"
SEL='time_tiled_batch_equals_block_by_block or time_tiled_other or all_partition_sizes_mono or mimo_dense or stereo_long_reverb'
for tool in racecheck synccheck; do
  timeout 1500 $CS --tool $tool python -m pytest tests/test_engine_gpu.py -x -q --timeout 1400 -k "$SEL" > $OUT/san_$tool.log 2>&1
  echo "$tool engine rc=$?" | tee -a $OUT/san_summary.txt
  grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|hazard" $OUT/san_$tool.log | tail -5 | tee -a $OUT/san_summary.txt
done
timeout 900 $CS --tool racecheck python -m pytest tests/test_coalesce_gpu.py -x -q --timeout 800 -k "concurrent_threads or integer_wire or one_launch_group" > $OUT/san_racecheck_single.log 2>&1
echo "racecheck single-stream path rc=$?" | tee -a $OUT/san_summary.txt
grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|hazard" $OUT/san_racecheck_single.log | tail -5 | tee -a $OUT/san_summary.txt
# ---- ThreadSanitizer
H=folve_b200/host; S=folve_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC,-fsanitize=thread \
     -c -o /tmp/fcv_engine_tsan.o $S/fcv_engine.cu > $OUT/tsan_build.log 2>&1 &&
g++ -fsanitize=thread -O1 -g -std=c++17 -I$H -I$H/sndfile_shim -Iinclude -o /tmp/tsan_host tools/tsan_host.cc \
    $H/sound-processor.cc $H/filter-config.cc $H/processor-pool.cc $H/batch-convolver.cc $H/sndfile_shim/sndfile_shim.cc \
    /tmp/fcv_engine_tsan.o $S/fcv_k_fft.o $S/fcv_k_fft13.o $S/fcv_k_mac.o $S/fcv_k_mac_tma.o $S/fcv_k_fused13.o \
    -L/usr/local/cuda/lib64 -lcudart -lpthread >> $OUT/tsan_build.log 2>&1
if [ -x /tmp/tsan_host ]; then
  TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 history_size=4" timeout 600 /tmp/tsan_host > $OUT/tsan_run.log 2>&1
  echo "tsan_host rc=$?" | tee -a $OUT/san_summary.txt
  echo "ThreadSanitizer warnings: $(grep -c 'WARNING: ThreadSanitizer' $OUT/tsan_run.log)" | tee -a $OUT/san_summary.txt
  grep -A12 "WARNING: ThreadSanitizer" $OUT/tsan_run.log | grep -E "WARNING|#0|#1|#2" | head -40 >> $OUT/san_summary.txt
else
  echo "tsan build failed" | tee -a $OUT/san_summary.txt; tail -5 $OUT/tsan_build.log
fi
