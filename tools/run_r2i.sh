timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -3
for w in "surround51 256" "surround51_dense 256"; do set -- $w
timeout 300 python bench.py --workload $1 --streams $2 --steps 40 --no-cpu-baseline --skip-e2e --wire s24 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']; r=d['roofline']
print('%-18s xRT %8.0f ms/step %.4f  fwd %.4f mac %.4f inv %.4f  fwd_frac %.3f' % ('$1', d['value'], d['ms_per_step'], k['fwd_fft'], k['mac'], k['inv_fft'], r['per_kernel']['fwd_fft']['frac']))"
done
