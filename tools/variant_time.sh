#!/bin/bash
# timing-only comparison of experiment variants (results of EXP variants are wrong on purpose)
cp folve_b200/libfolve_b200.so /tmp/base.so
for v in base "$@"; do
  if [ $v = base ]; then cp /tmp/base.so folve_b200/libfolve_b200.so; else cp folve_b200/variants/libfolve_b200_$v.so folve_b200/libfolve_b200.so; fi
  echo "== variant $v"; TS="${TS:-4}" bash tools/quick.sh 2>&1 | grep "^T "
done
cp /tmp/base.so folve_b200/libfolve_b200.so
