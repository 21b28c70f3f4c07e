#!/bin/bash
# device-resident throughput of every BASELINE config (parity-test cases; bench.py's line is santalucia)
for w in lowpass santalucia roomcorr96 roomcorr192 crossfeed surround51 surround51_dense; do
  timeout -s KILL 120 python bench.py --workload $w --streams ${STREAMS:-256} --steps 50 --no-cpu-baseline --skip-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']; r=d['roofline']
print('%-18s xRT %9.0f ms/step %.4f fwd %.4f mac %.4f inv %.4f  %s' % ('$w', d['value'], d['ms_per_step'], k['fwd_fft'], k['mac'], k['inv_fft'], d['config']['workload'][:90]))"
done
