# A/B of whole environment settings on one box: tools/run_ab_envs.sh "A=1" "A=2 B=3" ... ("-" = none)
for rep in 1 2; do for v in "$@"; do
e="$v"; [ "$v" = "-" ] && e="FCV_NONE=1"
env $e timeout 300 python bench.py --steps 100 --no-cpu-baseline --skip-e2e --no-configs --no-library 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']
print('%-28s xRT %8.0f ms/step %.4f  fwd %.4f mac %.4f inv %.4f frac %.3f' % ('$v', d['value'], d['ms_per_step'], k['fwd_fft'], k['mac'], k['inv_fft'], d['roofline']['frac']))"
done; done
