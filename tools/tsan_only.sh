#!/bin/bash
# ThreadSanitizer over the host side (dispatcher, BatchConvolver, MultiDeviceConvolver instances); see tools/sanitize.sh
OUT=gpurun_out; H=folve_b200/host; S=folve_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC,-fsanitize=thread \
     -c -o /tmp/fcv_engine_tsan.o $S/fcv_engine.cu > $OUT/tsan_build.log 2>&1 &&
nvcc -gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC,-fsanitize=thread \
     -c -o /tmp/fcv_nonuniform_tsan.o $S/fcv_nonuniform.cu >> $OUT/tsan_build.log 2>&1 &&
g++ -fsanitize=thread -O1 -g -std=c++17 -I$H -I$H/sndfile_shim -Iinclude -o /tmp/tsan_host tools/tsan_host.cc \
    $H/sound-processor.cc $H/filter-config.cc $H/processor-pool.cc $H/batch-convolver.cc $H/sndfile_shim/sndfile_shim.cc \
    /tmp/fcv_engine_tsan.o /tmp/fcv_nonuniform_tsan.o $S/fcv_k_fft.o $S/fcv_k_fft13.o $S/fcv_k_mac.o $S/fcv_k_mac_tma.o $S/fcv_k_fused13.o \
    -L/usr/local/cuda/lib64 -lcudart -lpthread >> $OUT/tsan_build.log 2>&1
if [ -x /tmp/tsan_host ]; then
  for d in 1 2; do
    FCV_DISPATCHERS=$d TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 history_size=4" timeout 600 /tmp/tsan_host > $OUT/tsan_run_$d.log 2>&1
    echo "tsan_host (FCV_DISPATCHERS=$d) rc=$? warnings: $(grep -c 'WARNING: ThreadSanitizer' $OUT/tsan_run_$d.log)" | tee -a $OUT/tsan2_summary.txt
    grep -A12 "WARNING: ThreadSanitizer" $OUT/tsan_run_$d.log | grep -E "WARNING|#0|#1|#2" | head -20 | tee -a $OUT/tsan2_summary.txt
  done
else
  echo "tsan build failed" | tee -a $OUT/tsan2_summary.txt; tail -8 $OUT/tsan_build.log
fi
