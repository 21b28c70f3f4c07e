#!/usr/bin/env python
"""The drop-in API itself on the GPU: SoundProcessor::FillBuffer / WriteProcessed, one file per host
thread (harness.cc fh_bench_threads), for several thread counts.  A/B knobs: FCV_COMBINE_MAX=1 (one launch
group per block), FCV_FUSED=1 (one cooperative launch per group), FCV_COMBINE_TRACE=1.
usage: tools/sp_sync.py [threads ...]"""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from folve_b200 import workloads

wl = workloads.WORKLOADS[os.environ.get("WORKLOAD", "santalucia")]()
threads = [int(a) for a in sys.argv[1:]] or [1, 4, 16, 32]
with tempfile.TemporaryDirectory() as tmp:
    d = workloads.write_filter_dir(wl, os.path.join(tmp, wl.name))
    bench.harness_run(bench.HOST_SO, wl, d, 4, 20)
    for t in threads:
        a, w = bench.harness_run(bench.HOST_SO, wl, d, t, 400)
        print(f"threads {t:3d}: {a / w:9.0f} x realtime, {1e6 * w / 400:7.1f} us per block and thread "
              f"[COMBINE_MAX={os.environ.get('FCV_COMBINE_MAX', '')} FUSED={os.environ.get('FCV_FUSED', '')}]", flush=True)
