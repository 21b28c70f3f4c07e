"""Sharding of independent streams over the GPUs of one box.

Files of one directory (album) are chained gaplessly through ONE processor
(convolve-file-handler.cc:390-415), so the unit of placement is the directory:
all files of an album go to the same GPU and no state ever crosses GPUs.
No collective is needed; ranks only agree on the assignment, which is a pure
function of the path."""
from __future__ import annotations

import os
import zlib


def album_of(path: str) -> str:
    return os.path.dirname(path.rstrip("/")) or "/"


def device_for_path(path: str, ndevices: int) -> int:
    """GPU that owns `path`'s album; stable across processes and runs (CRC32, not hash())."""
    return zlib.crc32(album_of(path).encode()) % max(1, ndevices)


def shard(paths, rank: int, world: int):
    """The files rank `rank` of `world` processes (one per GPU) is responsible for."""
    return [p for p in paths if device_for_path(p, world) == rank]


def balanced_albums(nalbums: int, rank: int, world: int):
    """Benchmark placement: album a of a synthetic library goes to rank a % world."""
    return list(range(rank, nalbums, world))
