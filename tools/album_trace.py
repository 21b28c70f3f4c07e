#!/usr/bin/env python
"""One rank's share of the config-5 album library at a given world size, with BatchConvolver's host trace:
tools/album_trace.py <world> [threads]   (FOLVE_B200_TRACE=1 is set here)"""
import ctypes as C, os, sys, tempfile
os.environ["FOLVE_B200_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from folve_b200 import workloads
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cores = len(os.sched_getaffinity(0))
threads = int(sys.argv[2]) if len(sys.argv) > 2 else max(2, cores // world)
wl = workloads.WORKLOADS["santalucia"]()
L = C.CDLL(bench.HOST_SO)
L.fh_bench_albums.restype = C.c_double
L.fh_bench_albums.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                              C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]
with tempfile.TemporaryDirectory() as tmp:
    d = workloads.write_filter_dir(wl, os.path.join(tmp, wl.name))
    cfg = os.path.join(d, f"filter-{wl.fs}.conf")
    audio, chains = C.c_double(0), C.c_int(0)
    wall = L.fh_bench_albums(cfg.encode(), wl.fs, wl.ninp, 128, 8, 0, world, 0, 8, threads, 1, C.byref(audio), C.byref(chains))
    print(f"world {world}: {chains.value} chains, {threads} threads: {audio.value / wall:.0f} x realtime ({wall:.3f} s)")
