# final N = 8 run (driver's launch line): bench line incl. library leg, one-process library, reference arm
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533"
nproc
timeout 600 $TR bench.py --gpus 4 --steps 20 --warmup 5 --no-configs > gpurun_out/r02e_bench_n4.json 2> gpurun_out/r02e_bench_n4.err; tail -c 300 gpurun_out/r02e_bench_n4.err
timeout 200 python bench.py --impl reference --gpus 4 --steps 3 --warmup 3 > gpurun_out/r02e_bench_ref_on4.json 2>/dev/null
for f in r02e_bench_n4; do python - $f <<'PY'
import json,sys
d=json.load(open('gpurun_out/%s.json'%sys.argv[1])); e=d['e2e']
print(sys.argv[1], 'value %.0f e2e %.0f ceiling %.0f frac %.3f' % (d['value'], e['value'], e['link_ceiling']['value'], e['link_ceiling']['e2e_frac_of_ceiling']), 'library', {k:round(v['value']) for k,v in e.items() if isinstance(v,dict) and k.startswith('album')})
PY
done
python -c "import json; d=json.load(open('gpurun_out/r02e_bench_ref_on4.json')); print('reference arm', round(d['value']), d['cpu_baseline']['cores'])"
