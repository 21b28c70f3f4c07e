#!/bin/bash
# A/B of the number of dispatcher threads on the per-file path
for rep in 1 2; do for n in 1 2 3; do echo "FCV_DISPATCHERS=$n"; FCV_DISPATCHERS=$n timeout 200 python tools/sp_sync.py 8 16 32 2>&1 | tail -3; done; done
FCV_DISPATCHERS=2 timeout 300 python -m pytest tests/test_coalesce_gpu.py -m gpu -x -q --timeout 250 2>&1 | tail -1
FCV_DISPATCHERS=2 SOAK_SECONDS=20 timeout 200 python tools/soak_sp.py 2>&1 | tail -1
