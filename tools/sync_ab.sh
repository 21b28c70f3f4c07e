#!/bin/bash
# per-file drop-in path with and without per-slot CUDA graphs: block latency and the
# 16-thread SoundProcessor loop (bench.py's e2e.soundprocessor_sync)
for g in 1 0; do
FCV_STREAM_GRAPHS=$g python bench.py --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('graphs=$g latency us', d['block_latency_us']['median'], d['block_latency_us']['p99'], 'soundprocessor_sync xRT', d['e2e']['soundprocessor_sync']['value'], 'threads', d['e2e']['soundprocessor_sync']['threads'])"
done
