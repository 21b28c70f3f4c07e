set -x
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -25
timeout 120 python tools/sp_sync.py 1 4 16 32 2>&1 | tail -5
FCV_COMBINE_DEPTH=1 timeout 120 python tools/sp_sync.py 16 32 2>&1 | tail -3
FCV_COMBINE_DEPTH=3 timeout 120 python tools/sp_sync.py 16 32 2>&1 | tail -3
FCV_COMBINE_DEPTH=4 timeout 120 python tools/sp_sync.py 16 32 2>&1 | tail -3
FCV_STREAM_ZEROCOPY=0 timeout 120 python tools/sp_sync.py 1 16 2>&1 | tail -3
