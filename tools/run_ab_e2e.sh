# A/B of library builds on the END-TO-END loop: tools/run_ab_e2e.sh base nq8 ...
cp folve_b200/libfolve_b200.so /tmp/base.so
for v in "$@"; do [ $v = base ] || cp folve_b200/libfolve_b200_$v.so /tmp/$v.so; done
for rep in 1 2 3; do for v in "$@"; do cp /tmp/$v.so folve_b200/libfolve_b200.so
timeout 300 python bench.py --steps 40 --no-cpu-baseline --no-configs --no-library ${BENCH_ARGS:-} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']
print('%-8s e2e %8.0f (%.3f ms) ceiling %8.0f frac %.3f' % ('$v', e['value'], e['ms_per_step'], e['link_ceiling']['value'], e['link_ceiling']['e2e_frac_of_ceiling']))"
done; done
cp /tmp/base.so folve_b200/libfolve_b200.so
