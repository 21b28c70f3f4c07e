"""BatchConvolver throughput (16-bit files, int16 wire) against the number of host threads."""
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from folve_b200 import workloads  # noqa: E402

wl = workloads.WORKLOADS["santalucia"]()
with tempfile.TemporaryDirectory() as tmp:
    d = workloads.write_filter_dir(wl, os.path.join(tmp, wl.name))
    for th in [int(a) for a in sys.argv[1:]] or [4, 8, 12, 16]:
        a, w = bench.library_run(wl, d, 1024, 2, 120.0, 8, th, True)
        print(f"threads {th:3d}: {a / w:9.0f} x realtime", flush=True)
