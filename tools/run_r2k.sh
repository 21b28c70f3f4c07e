export FCV_COMBINE_TRACE=1
for rep in 1 2 3 4; do
timeout 200 python tools/sp_sync.py 16 2>&1 | tail -2
done
