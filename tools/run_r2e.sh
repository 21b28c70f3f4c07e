timeout 300 python -m pytest tests/test_soundprocessor_gpu.py -m gpu -q --timeout 200 -k "one_process or placement" 2>&1 | tail -5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 300 gpurun_out/bench_n2.err
FCV_HUGEPAGES=1 timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-library > gpurun_out/bench_n2_huge.json 2> gpurun_out/bench_n2_huge.err; tail -c 300 gpurun_out/bench_n2_huge.err
grep -i huge /proc/meminfo | head -5; cat /sys/kernel/mm/transparent_hugepage/enabled
timeout 300 python bench.py --steps 20 --warmup 5 --no-configs --no-cpu-baseline > gpurun_out/bench_n1_on2.json 2> gpurun_out/bench_n1_on2.err; tail -c 300 gpurun_out/bench_n1_on2.err
for f in bench_n2 bench_n2_huge bench_n1_on2; do python - $f <<'PY'
import json,sys
d=json.load(open('gpurun_out/%s.json'%sys.argv[1])); e=d['e2e']
print(sys.argv[1], 'value %.0f e2e %.0f ceiling %.0f frac %.3f' % (d['value'], e['value'], e['link_ceiling']['value'], e['link_ceiling']['e2e_frac_of_ceiling']), 'library', {k:round(v['value']) for k,v in e.items() if isinstance(v,dict) and k.startswith('album')})
PY
done
