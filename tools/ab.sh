#!/bin/bash
# A/B of environment toggles on the device-resident loop: tools/ab.sh "VAR=val VAR2=val" "..." ...
# (each argument is one configuration; "" = defaults).  Prints ms per step and per kernel.
for cfg in "$@"; do
  env $cfg python bench.py --steps ${STEPS:-300} --no-cpu-baseline --skip-e2e ${BENCH_ARGS:-} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']
print('%-40s xRT %8.0f ms/step %.4f  fwd %.4f mac %.4f inv %.4f  clk %s %s' % ('$cfg' or 'default', d['value'], d['ms_per_step'], k['fwd_fft'], k['mac'], k['inv_fft'], d['clocks']['sm_mhz'], d['clocks']['reasons']))"
done
