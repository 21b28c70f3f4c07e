timeout 600 python -m pytest tests/test_nonuniform_gpu.py tests/test_coalesce_gpu.py tests/test_engine_gpu.py -x -q --timeout 300 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-configs --no-library 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(json.dumps(d['block_latency_us']['nonuniform_q1024'], indent=1))"
