#!/usr/bin/env python
"""Soak of the per-file path: many rounds of the threaded SoundProcessor block loop (pool churn: every
round creates and destroys its processors), thread counts from 1 to 4 x cores.  Any hang trips the timeout."""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from folve_b200 import workloads
wl = workloads.WORKLOADS["santalucia"]()
cores = len(os.sched_getaffinity(0))
t0 = time.time()
with tempfile.TemporaryDirectory() as tmp:
    d = workloads.write_filter_dir(wl, os.path.join(tmp, wl.name))
    rounds = 0
    while time.time() - t0 < float(os.environ.get("SOAK_SECONDS", "40")):
        for t in (1, 3, cores, 2 * cores + 1, 4 * cores):
            a, w = bench.harness_run(bench.HOST_SO, wl, d, t, 150)
            rounds += 1
print(f"soak: {rounds} rounds in {time.time() - t0:.1f} s, no hang")
