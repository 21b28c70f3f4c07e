"""ctypes binding of oracle/dropin_driver.cc: the reference's OWN filesystem layer
(FolveFilesystem, ConvolveFileHandler, ConversionBuffer, BufferThread, FileHandlerCache,
ProcessorPool -- compiled unmodified from /root/reference by `make -C oracle dropin refstack`):

dropin()   -> oracle/_ref/libfolve_dropin.so   : those callers on THIS repository's SoundProcessor
              and CUDA engine (needs a GPU)
refstack() -> oracle/_ref/libfolve_refstack.so : the same callers on the reference's SoundProcessor
              over the restated zita-convolver (CPU).  TEST INFRASTRUCTURE.

A file is read like a media player reads it from the mounted filesystem: sequential read() calls.
WAV input comes back as 24-bit "FLAC" (convolve-file-handler.cc:245-248) -- with the sndfile
shim an uncompressed stand-in: 42 header bytes, then interleaved little-endian int24 frames.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN_SO = os.path.join(ROOT, "oracle", "_ref", "libfolve_dropin.so")
REFSTACK_SO = os.path.join(ROOT, "oracle", "_ref", "libfolve_refstack.so")
HEADER_BYTES = 42


class Mount:
    def __init__(self, so, music_dir, config_base_dir, filter_name, gapless=True, pre_buffer_bytes=0):
        L = C.CDLL(so)
        L.dd_variant.restype = C.c_char_p
        L.dd_open.restype = C.c_void_p
        L.dd_open.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int]
        L.dd_read_file.restype = C.c_long
        L.dd_read_file.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_long, C.c_int, C.POINTER(C.c_int),
                                   C.POINTER(C.c_float)]
        self.L = L
        self.variant = L.dd_variant().decode()
        self.h = L.dd_open(str(music_dir).encode(), str(config_base_dir).encode(), filter_name.encode(),
                           1 if gapless else 0, pre_buffer_bytes)
        if not self.h:
            raise RuntimeError("dd_open failed")

    def read(self, fs_path, channels, read_size=65536, cap=1 << 28):
        """-> (int32 [frames, channels] of 24-bit samples, flags, max_output_value, header bytes)"""
        buf = np.zeros(cap, np.uint8)
        flags, mx = C.c_int(0), C.c_float(0)
        n = self.L.dd_read_file(self.h, fs_path.encode(), buf.ctypes.data, cap, read_size, C.byref(flags), C.byref(mx))
        if n < 0:
            raise RuntimeError(f"cannot read {fs_path}")
        raw = buf[:n]
        body = raw[HEADER_BYTES:]
        body = body[: len(body) // (3 * channels) * 3 * channels].reshape(-1, 3).astype(np.int32)
        v = body[:, 0] | (body[:, 1] << 8) | (body[:, 2] << 16)
        v = np.where(v & 0x800000, v - (1 << 24), v)
        return v.reshape(-1, channels), flags.value, mx.value, bytes(raw[:HEADER_BYTES])


def have_dropin():
    return os.path.exists(DROPIN_SO)


def have_refstack():
    return os.path.exists(REFSTACK_SO)
