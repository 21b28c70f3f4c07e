#!/bin/bash
# per-file drop-in path: block latency and the 16-thread SoundProcessor loop (bench.py's e2e.soundprocessor_sync)
for cfg in "$@"; do
env $cfg timeout -s KILL 200 python bench.py --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$cfg latency us', round(d['block_latency_us']['median'],1), round(d['block_latency_us']['p99'],1), 'soundprocessor_sync xRT', round(d['e2e']['soundprocessor_sync']['value']), 'threads', d['e2e']['soundprocessor_sync']['threads'], 'cpu ref', round(d['cpu_baseline']['value']))"
done
