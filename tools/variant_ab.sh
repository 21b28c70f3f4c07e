#!/bin/bash
# A/B of library variants built into folve_b200/variants/ (scratch copy on the GPU box only)
cp folve_b200/libfolve_b200.so /tmp/base.so
for v in base "$@"; do
  if [ $v = base ]; then cp /tmp/base.so folve_b200/libfolve_b200.so; else cp folve_b200/variants/libfolve_b200_$v.so folve_b200/libfolve_b200.so; fi
  echo "== variant $v"; python -m pytest tests/test_engine_gpu.py -m gpu -x -q 2>&1 | tail -1
  TS="${TS:-4}" bash tools/quick.sh 2>&1 | grep "^T "
done
cp /tmp/base.so folve_b200/libfolve_b200.so
