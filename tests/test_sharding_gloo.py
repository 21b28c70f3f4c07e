"""CPU, world_size 2 over gloo: the N>1 host logic -- album-level sharding with no
data-path collective, and the barrier + max-over-ranks timing reduction bench.py uses."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from folve_b200 import sharding
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    paths = [f"/music/album{a:03d}/track{t:02d}.flac" for a in range(37) for t in range(8)]
    mine = sharding.shard(paths, rank, world)
    # gapless neighbours are co-located: an album is never split
    albums = {sharding.album_of(p) for p in mine}
    assert all(sharding.device_for_path(p, world) == rank for p in mine)
    assert len(mine) == 8 * len(albums)
    # the union over ranks covers every file exactly once (checked with a gather of counts and ids)
    ids = torch.zeros(len(paths), dtype=torch.int32)
    for p in mine:
        ids[paths.index(p)] = 1
    dist.all_reduce(ids, op=dist.ReduceOp.SUM)
    assert bool((ids == 1).all())
    # timing reduction: barrier, then the maximum over ranks is what every rank reports
    dist.barrier()
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert float(t) == float(world)
    bal = sharding.balanced_albums(128, rank, world)
    assert len(bal) == 128 // world and all(a % world == rank for a in bal)
    q.put((rank, len(mine)))
    dist.destroy_process_group()


def test_album_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = dict(q.get() for _ in range(2))
    assert got[0] + got[1] == 37 * 8
    assert min(got.values()) > 0


def test_cxx_placement_equals_the_python_rule():
    """SoundProcessor::DeviceForKey (the in-process placement of the C++ host layer) and sharding.device_for_path
    (what the ranks of bench.py agree on) are the same function of the album directory: 3000 random keys, non-ASCII
    names included, 1 to 8 devices.  CRC32 only -- no GPU involved."""
    import ctypes as C

    import numpy as np
    sys.path.insert(0, ROOT)
    from folve_b200 import sharding
    so = os.path.join(ROOT, "folve_b200", "libfolve_host.so")
    if not os.path.exists(so):
        pytest.skip("needs folve_b200/libfolve_host.so")
    f = C.CDLL(so).fh_device_for_key
    f.restype, f.argtypes = C.c_int, [C.c_char_p, C.c_int]
    r = np.random.default_rng(5)
    words = ["music", "Ünïcode", "a b", "日本語", "x" * 200, "1999 - Live", "cd.2", "flac"]
    used = set()
    for _ in range(3000):
        album = "/" + "/".join(words[int(k)] for k in r.integers(0, len(words), int(r.integers(1, 6))))
        n = int(r.integers(1, 9))
        want = sharding.device_for_path(album + "/01 - track.flac", n)
        assert f(album.encode(), n) == want, (album, n)
        used.add((n, want))
    assert len(used) == sum(range(1, 9))            # every device of every box size is hit
