# A/B of one environment switch on one box: tools/run_ab_env.sh FCV_FWD_BOTH 0 1
var=$1; shift
for v in "$@"; do echo "== engine parity tests with $var=$v"; env $var=$v timeout 400 python -m pytest tests/test_engine_gpu.py tests/test_scale_properties_gpu.py -m gpu -x -q --timeout 200 2>&1 | tail -1; done
for rep in 1 2; do for v in "$@"; do
env $var=$v timeout 300 python bench.py --steps 100 --no-cpu-baseline --skip-e2e --no-configs --no-library 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']
print('$var=%-3s xRT %8.0f ms/step %.4f  fwd %.4f mac %.4f inv %.4f frac %.3f' % ('$v', d['value'], d['ms_per_step'], k['fwd_fft'], k['mac'], k['inv_fft'], d['roofline']['frac']))"
done; done
