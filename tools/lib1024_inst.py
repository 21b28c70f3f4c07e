#!/usr/bin/env python
"""1024 gapless chains on one GPU through 1 / 2 BatchConvolver instances (FOLVE_B200_LIBRARY_INSTANCES)."""
import ctypes as C, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from folve_b200 import workloads
wl = workloads.WORKLOADS["santalucia"]()
L = C.CDLL(bench.HOST_SO)
L.fh_bench_albums.restype = C.c_double
L.fh_bench_albums.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                              C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]
with tempfile.TemporaryDirectory() as tmp:
    d = workloads.write_filter_dir(wl, os.path.join(tmp, wl.name))
    cfg = os.path.join(d, f"filter-{wl.fs}.conf")
    audio, chains = C.c_double(0), C.c_int(0)
    wall = L.fh_bench_albums(cfg.encode(), wl.fs, wl.ninp, 1024, 1, 0, 1, 1, 8, len(os.sched_getaffinity(0)), 1, C.byref(audio), C.byref(chains))
    print(f"instances {os.environ.get('FOLVE_B200_LIBRARY_INSTANCES', '1')}: {chains.value} chains, {audio.value / wall:.0f} x realtime, wall {wall:.2f} s")
