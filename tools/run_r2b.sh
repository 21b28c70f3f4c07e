export FCV_COMBINE_TRACE=1
for d in 2 4 6; do for sp in 1 0; do FCV_COMBINE_SPIN=$sp FCV_COMBINE_DEPTH=$d timeout 120 python tools/sp_sync.py 16 2>&1 | grep -v "^$" | tail -2; done; done
FCV_COMBINE_DEPTH=4 timeout 120 python tools/sp_sync.py 32 2>&1 | grep -v "^$" | tail -2
FCV_COMBINE_DEPTH=4 timeout 120 python tools/sp_sync.py 8 2>&1 | grep -v "^$" | tail -2
