for T in 1 2 4 8; do
timeout 300 python bench.py --blocks-per-step $T --steps 40 --no-cpu-baseline --skip-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']; r=d['roofline']; T=$T
print('T=%d xRT %8.0f ms/step %.4f per block: fwd %.4f mac %.4f inv %.4f | mac bytes/block %.3f GB (compulsory) frac %.3f | block-sync bytes/block %.3f GB' % (T, d['value'], d['ms_per_step'], k['fwd_fft']/T, k['mac']/T, k['inv_fft']/T, r['algorithmic_bytes_per_launch']/T/1e9, r['frac'], r['block_sync_bytes_per_launch']/T/1e9))"
done
FCV_HUGEPAGES=1 timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-configs --no-library 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']; print('hugepages N=1: e2e %.0f ceiling %.0f' % (e['value'], e['link_ceiling']['value']))"
timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-configs --no-library 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']; print('default   N=1: e2e %.0f ceiling %.0f' % (e['value'], e['link_ceiling']['value']))"
