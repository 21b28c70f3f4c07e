set -x
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -25
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; tail -c 1500 gpurun_out/bench_c.err
