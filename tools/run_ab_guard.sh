# guarded A/B of one environment switch: quick parity run first (120 s limit), stop if it fails or hangs
# tools/run_ab_guard.sh VAR v1 v2 ...   (the first value is tested first)
var=$1; shift
for v in "$@"; do
  echo "== engine parity tests with $var=$v"
  env $var=$v timeout -s KILL 150 python -m pytest tests/test_engine_gpu.py tests/test_scale_properties_gpu.py -m gpu -x -q --timeout 100 2>&1 | tail -3
  [ ${PIPESTATUS[0]} -ne 0 ] && { echo "parity run failed for $var=$v: stopping"; exit 1; }
done
for rep in 1 2; do for v in "$@"; do
env $var=$v timeout -s KILL 200 python bench.py --steps 100 --no-cpu-baseline --skip-e2e --no-configs --no-library 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']
print('$var=%-3s xRT %8.0f ms/step %.4f  fwd %.4f mac %.4f inv %.4f frac %.3f' % ('$v', d['value'], d['ms_per_step'], k['fwd_fft'], k['mac'], k['inv_fft'], d['roofline']['frac']))"
done; done
