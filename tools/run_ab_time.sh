# timing-only A/B of library builds on one box (no tests): tools/run_ab_time.sh base abl1 abl2 ...
cp folve_b200/libfolve_b200.so /tmp/base.so
for v in "$@"; do [ $v = base ] || cp folve_b200/libfolve_b200_$v.so /tmp/$v.so; done
for rep in 1 2; do for v in "$@"; do cp /tmp/$v.so folve_b200/libfolve_b200.so
timeout 300 python bench.py --steps 100 --no-cpu-baseline --skip-e2e --no-configs --no-library 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']
print('%-10s xRT %8.0f ms/step %.4f  fwd %.4f mac %.4f inv %.4f' % ('$v', d['value'], d['ms_per_step'], k['fwd_fft'], k['mac'], k['inv_fft']))"
done; done
cp /tmp/base.so folve_b200/libfolve_b200.so
