# tools/run_ab_wl.sh <workload> <streams> "ENV=.." ... : device-resident kernel times of one workload under several settings
wl=$1; st=$2; shift 2
for v in "$@"; do
e="$v"; [ "$v" = "-" ] && e="FCV_NONE=1"
env $e timeout -s KILL 120 python bench.py --workload $wl --streams $st --steps 60 --no-cpu-baseline --skip-e2e --no-configs --no-library 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']
print('%-10s %-18s xRT %8.0f ms/step %.4f  fwd %.4f mac %.4f inv %.4f' % ('$wl', '$v', d['value'], d['ms_per_step'], k['fwd_fft'], k['mac'], k['inv_fft']))"
done
