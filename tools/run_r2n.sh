for rep in 1 2; do
timeout 200 python tools/sp_sync.py 8 16 32 2>&1 | tail -3
FCV_GROUP_MAC_S=4 timeout 200 python tools/sp_sync.py 8 16 32 2>&1 | tail -3 | sed "s/^/S=auto /"
done
FCV_COMBINE_TRACE=1 timeout 200 python tools/sp_sync.py 16 2>&1 | tail -2
timeout 300 python -m pytest tests/test_coalesce_gpu.py -x -q --timeout 120 2>&1 | tail -2
