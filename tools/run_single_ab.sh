#!/bin/bash
# the per-file path after a change: parity tests, kernel durations of a lone stream, throughput for 1 / 16 / 32 host threads
timeout 600 python -m pytest tests/test_coalesce_gpu.py tests/test_soundprocessor_gpu.py tests/test_engine_gpu.py tests/test_nonuniform_gpu.py -m gpu -x -q --timeout 300 2>&1 | tail -2
bash tools/single_launches.sh
for i in 1 2; do python tools/sp_sync.py 1 16 32 2>&1 | tail -3; done
