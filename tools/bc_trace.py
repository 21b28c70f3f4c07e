#!/usr/bin/env python
"""BatchConvolver on the benchmark's 1024 chains with its host trace (FOLVE_B200_TRACE=1):
tools/bc_trace.py [threads ...]"""
import os, sys, tempfile
os.environ["FOLVE_B200_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from folve_b200 import workloads
wl = workloads.WORKLOADS["santalucia"]()
cores = len(os.sched_getaffinity(0))
with tempfile.TemporaryDirectory() as tmp:
    d = workloads.write_filter_dir(wl, os.path.join(tmp, wl.name))
    for t in [int(a) for a in sys.argv[1:]] or [cores]:
        a, w = bench.library_run(wl, d, 1024, 2, 120.0, 8, t, True)
        print(f"threads {t}: {a / w:.0f} x realtime ({w:.2f} s)", flush=True)
