# A/B of two builds of the library on one box: folve_b200/libfolve_b200.so (new) against libfolve_b200_ab.so (old)
cp folve_b200/libfolve_b200.so /tmp/new.so; cp folve_b200/libfolve_b200_ab.so /tmp/old.so
for rep in 1 2; do for v in new old; do cp /tmp/$v.so folve_b200/libfolve_b200.so
timeout 300 python bench.py --steps 100 --no-cpu-baseline --skip-e2e ${BENCH_ARGS:-} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']
print('$v  xRT %8.0f ms/step %.4f  fwd %.4f mac %.4f inv %.4f' % (d['value'], d['ms_per_step'], k['fwd_fft'], k['mac'], k['inv_fft']))"
done; done
cp /tmp/new.so folve_b200/libfolve_b200.so
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --no-configs --no-library 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); p=d['cpu_baseline']['parity']; print({k:v for k,v in p.items() if k!='what'})"
