#!/usr/bin/env python
"""One stream as a batch of one (device-resident PCM, no host copy-out): kernel durations to compare with the
single-stream path's (tools/single_launches.sh) -- the difference is what the kernel-side host access costs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from folve_b200 import capi, workloads
wl = workloads.WORKLOADS["santalucia"]()
f = wl.load(capi.Filter(wl.ninp, wl.nout, wl.size, wl.fragm)).commit(0)
b = capi.Batch(f, 1)
b.host_in[:] = np.random.default_rng(0).uniform(-0.03, 0.03, b.host_in.shape).astype(np.float32)
b.process()
for _ in range(80):
    b.process_device()
b.sync()
b.close(); f.close()
