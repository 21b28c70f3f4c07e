#!/bin/bash
# config-5 album library on one GPU with 1 / 2 / 3 BatchConvolver instances per GPU
for rep in 1 2; do for k in 1 2 3; do
FOLVE_B200_LIBRARY_INSTANCES=$k timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); a=d['e2e']['album_library']
print('instances $k: library %.0f x realtime, wall %.2f s' % (a['value'], a['wall_s']))"
done; done
