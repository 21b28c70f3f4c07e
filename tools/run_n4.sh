TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544"
timeout 600 $TR bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; tail -c 200 gpurun_out/bench_n4.err
timeout 300 $TR bench.py --impl reference --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_ref_n4.json 2>/dev/null
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n4.json')); e=d['e2e']
print('N=4 value %.0f e2e %.0f ceiling %.0f frac %.3f' % (d['value'], e['value'], e['link_ceiling']['value'], e['link_ceiling']['e2e_frac_of_ceiling']), 'library', {k:round(v['value']) for k,v in e.items() if isinstance(v,dict) and k.startswith('album')})
r=json.load(open('gpurun_out/bench_ref_n4.json')); print('reference arm', round(r['value']), r['cpu_baseline']['cores'])
PY
timeout 300 python -m pytest tests/test_soundprocessor_gpu.py -m gpu -q --timeout 200 -k "one_process" 2>&1 | tail -2
