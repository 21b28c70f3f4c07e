#!/bin/bash
# end-to-end loop (fcv_batch_submit / wait, pinned host buffers) under different settings
for cfg in "$@"; do
  env $cfg python bench.py --steps ${STEPS:-150} --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']
print('%-32s e2e xRT %8.0f ms/step %.3f  f32 wire ms/step %.3f  (device ms/step %.3f)' % ('$cfg' or 'default', e['value'], e['ms_per_step'], e['f32_wire']['ms_per_step'], d['ms_per_step']))"
done
