"""One stream, 40 blocks through fcv_stream_process (SantaLucia): run under
`ncu --metrics gpu__time_duration.sum` to see what each kernel of a block costs on its own."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from folve_b200 import capi, workloads  # noqa: E402

wl = workloads.WORKLOADS["santalucia"]()
flt = wl.load(capi.Filter(wl.ninp, wl.nout, wl.size, wl.fragm)).commit(0)
st = capi.Stream(flt)
N = wl.fragm
st.buffer[: N * wl.ninp] = np.random.default_rng(0).uniform(-0.03, 0.03, N * wl.ninp).astype(np.float32)
m = C.c_float(0)
L = capi.lib()
ts = []
for k in range(int(sys.argv[1]) if len(sys.argv) > 1 else 40):
    t0 = time.perf_counter()
    L.fcv_stream_process(st._h, N, C.byref(m))
    ts.append(time.perf_counter() - t0)
print("median call us", 1e6 * float(np.median(ts[5:])))
st.close()
flt.close()
