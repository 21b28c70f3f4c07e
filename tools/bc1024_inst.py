#!/usr/bin/env python
"""bench.py's batch_convolver leg (1024 gapless chains x 2 files) with 1 / 2 / 3 BatchConvolver instances on the GPU."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from folve_b200 import workloads
wl = workloads.WORKLOADS["santalucia"]()
with tempfile.TemporaryDirectory() as tmp:
    d = workloads.write_filter_dir(wl, os.path.join(tmp, wl.name))
    for k in (1, 2, 3, 1, 2, 3):
        os.environ["FOLVE_B200_LIBRARY_INSTANCES"] = str(k)
        a, w = bench.library_run(wl, d, 1024, 2, 120.0, 8, len(os.sched_getaffinity(0)), True)
        print(f"instances {k}: {a / w:.0f} x realtime ({w:.2f} s)", flush=True)
