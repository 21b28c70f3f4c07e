CS="compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 10"
timeout 900 $CS python -m pytest tests/test_coalesce_gpu.py -x -q --timeout 800 > gpurun_out/san_memcheck_single.log 2>&1
echo "memcheck single-stream path rc=$?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/san_memcheck_single.log | tail -3
timeout 1200 $CS python -m pytest tests/test_engine_gpu.py tests/test_soundprocessor_gpu.py tests/test_dropin_stack.py -m gpu -x -q --timeout 1100 -k "time_tiled_other or mimo_dense or block_edges or batch_frames_valid or async_submit or golden or gapless or truncated or replaced_impulse or reference_callers or demo_filters or batch_convolver_equals" > gpurun_out/san_memcheck_engine.log 2>&1
echo "memcheck engine + host layer rc=$?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/san_memcheck_engine.log | tail -3
