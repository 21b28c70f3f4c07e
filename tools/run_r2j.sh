timeout 300 python -m pytest tests/test_coalesce_gpu.py -x -q --timeout 120 2>&1 | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
export FCV_COMBINE_TRACE=1
for t in 1 8 16 32; do timeout 120 python tools/sp_sync.py $t 2>&1 | grep -v "^$" | tail -2; done
for t in 1 16; do FCV_FUSED=0 timeout 120 python tools/sp_sync.py $t 2>&1 | grep -v "^$" | tail -2; done
unset FCV_COMBINE_TRACE
for d in 4 6 8 12; do FCV_COMBINE_DEPTH=$d timeout 120 python tools/sp_sync.py 16 32 2>&1 | tail -2; done
FCV_FUSED=0 timeout 120 python tools/sp_sync.py 16 2>&1 | tail -1
