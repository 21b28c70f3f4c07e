# A/B of several builds of the library on one box: tools/run_ab_lib.sh base split splitpf ...
# (folve_b200/libfolve_b200.so = "base", folve_b200/libfolve_b200_<name>.so the variants)
cp folve_b200/libfolve_b200.so /tmp/base.so
for v in "$@"; do [ $v = base ] || cp folve_b200/libfolve_b200_$v.so /tmp/$v.so; done
for rep in 1 2; do for v in "$@"; do cp /tmp/$v.so folve_b200/libfolve_b200.so
timeout 300 python bench.py --steps 100 --no-cpu-baseline --skip-e2e ${BENCH_ARGS:-} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']
print('%-10s xRT %8.0f ms/step %.4f  fwd %.4f mac %.4f inv %.4f' % ('$v', d['value'], d['ms_per_step'], k['fwd_fft'], k['mac'], k['inv_fft']))"
done; done
for v in "$@"; do cp /tmp/$v.so folve_b200/libfolve_b200.so; echo "== tests with $v"; timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -x -q --timeout 300 2>&1 | tail -1; done
cp /tmp/base.so folve_b200/libfolve_b200.so
