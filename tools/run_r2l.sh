timeout 300 python -m pytest tests/test_coalesce_gpu.py tests/test_engine_gpu.py -x -q --timeout 120 2>&1 | tail -4
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for rep in 1 2; do timeout 200 python tools/sp_sync.py 1 2 4 8 16 32 64 2>&1 | tail -7; done
FCV_COMBINE_TRACE=1 timeout 200 python tools/sp_sync.py 16 2>&1 | tail -2
FCV_COMBINE_TRACE=1 timeout 200 python tools/sp_sync.py 32 2>&1 | tail -2
