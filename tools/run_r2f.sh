CS="compute-sanitizer --error-exitcode 7 --print-limit 5 --tool racecheck"
timeout 600 $CS --num-cuda-barriers 32 python -m pytest tests/test_engine_gpu.py -x -q --timeout 500 -k "time_tiled_batch_equals_block_by_block" > gpurun_out/san_racecheck_nb32.log 2>&1
echo "racecheck --num-cuda-barriers 32: rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY|Warning" gpurun_out/san_racecheck_nb32.log | tail -4
FCV_MAC_TMA=0 timeout 600 $CS python -m pytest tests/test_engine_gpu.py -x -q --timeout 500 -k "time_tiled_batch_equals_block_by_block" > gpurun_out/san_racecheck_notma.log 2>&1
echo "racecheck FCV_MAC_TMA=0 (register-pipelined MAC): rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY|Warning" gpurun_out/san_racecheck_notma.log | tail -4
FOLVE_B200_TRACE=1 timeout 300 python bench.py --steps 10 --no-configs --no-cpu-baseline 2>&1 >/dev/null | grep "BatchConvolver::Run" | tail -3
