# A/B of the MAC producer's rows-per-pass (FCV_MAC_G) on one box: tools/run_ab_macg.sh 8 4 2 1 0
for rep in 1 2; do for v in "$@"; do
FCV_MAC_G=$v timeout 300 python bench.py --steps 100 --no-cpu-baseline --skip-e2e --no-configs --no-library 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']
print('FCV_MAC_G=%-3s xRT %8.0f ms/step %.4f  fwd %.4f mac %.4f inv %.4f frac %.3f' % ('$v', d['value'], d['ms_per_step'], k['fwd_fft'], k['mac'], k['inv_fft'], d['roofline']['frac']))"
done; done
for v in "$@"; do echo "== engine parity tests with FCV_MAC_G=$v"; FCV_MAC_G=$v timeout 400 python -m pytest tests/test_engine_gpu.py tests/test_scale_properties_gpu.py -m gpu -x -q --timeout 200 2>&1 | tail -1; done
