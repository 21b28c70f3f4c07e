#!/bin/bash
# quick GPU check used during development: parity tests + short bench
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps ${1:-50} --warmup 5 --no-cpu-baseline ${2:-} | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('xRT', round(d['value']), 'ms/step', round(d['ms_per_step'],4), d['kernel_ms_per_step'], 'e2e', d['e2e']['value'], 'roofline', round(d['roofline']['frac'],3))"
