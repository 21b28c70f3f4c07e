for i in 1 2 3; do for h in 1 0; do
FCV_HUGEPAGES=$h timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-configs --no-library 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']; print('hugepages=$h: e2e %.0f ceiling %.0f f32 %.0f' % (e['value'], e['link_ceiling']['value'], e['f32_wire']['value']))"
done; done
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -3
