"""ctypes binding of the C++ caller-protocol harness (folve_b200/host/harness.cc).

product()   -> folve_b200/libfolve_host.so : this repo's SoundProcessor on the B200 engine
reference() -> oracle/_ref/libfolve_ref.so : the reference's own sound-processor.cc /
               zita-config.cc / processor-pool.cc compiled unmodified against the
               restated Convproc (CPU).  TEST INFRASTRUCTURE.
"""
from __future__ import annotations

import ctypes as C
import os
import struct

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRODUCT_SO = os.path.join(ROOT, "folve_b200", "libfolve_host.so")
REFERENCE_SO = os.path.join(ROOT, "oracle", "_ref", "libfolve_ref.so")

SF_FORMAT_PCM_16, SF_FORMAT_PCM_24, SF_FORMAT_PCM_32, SF_FORMAT_FLOAT = 2, 3, 4, 6
_DTYPE = {SF_FORMAT_PCM_16: np.int16, SF_FORMAT_PCM_24: np.int32, SF_FORMAT_PCM_32: np.int32,
          SF_FORMAT_FLOAT: np.float32}

_libs = {}


def _load(path):
    if path in _libs:
        return _libs[path]
    L = C.CDLL(path)
    vp, i = C.c_void_p, C.c_int
    L.fh_version.restype = C.c_char_p
    L.fh_set_reset_is_fresh.argtypes = [i]
    L.fh_run_chain.restype = i
    L.fh_run_chain.argtypes = [C.c_char_p, i, i, i, i, i, C.POINTER(vp), C.POINTER(C.c_long), i, i,
                               C.POINTER(vp), C.POINTER(C.c_long), C.POINTER(C.c_float), C.POINTER(i),
                               C.c_char_p, i]
    L.fh_config_open.restype = vp
    L.fh_config_open.argtypes = [C.c_char_p, i, i]
    L.fh_config_info.argtypes = [vp, C.POINTER(i)]
    L.fh_config_impulse.restype = i
    L.fh_config_impulse.argtypes = [vp, i, i, C.POINTER(C.c_float), i]
    L.fh_config_close.argtypes = [vp]
    _libs[path] = L
    return L


def have_reference():
    return os.path.exists(REFERENCE_SO)


def have_product():
    return os.path.exists(PRODUCT_SO)


class Harness:
    def __init__(self, path):
        self.L = _load(path)
        self.kind = self.L.fh_version().decode()

    def run_chain(self, filter_dir, samplerate, channels, bits, files, gapless=True,
                  in_format=SF_FORMAT_FLOAT, out_format=SF_FORMAT_FLOAT):
        """files: list of [frames, channels] arrays in the dtype of in_format.
        Returns (outputs list of [frames, nout], max_values, gapless_flags)."""
        n = len(files)
        arrs = [np.ascontiguousarray(f, dtype=_DTYPE[in_format]) for f in files]
        frames = (C.c_long * n)(*[a.shape[0] for a in arrs])
        pin = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        nout_cap = 64
        outs = [np.zeros((a.shape[0], nout_cap), dtype=_DTYPE[out_format]) for a in arrs]
        pout = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
        oframes = (C.c_long * n)()
        mx = (C.c_float * n)()
        flags = (C.c_int * n)()
        err = C.create_string_buffer(512)
        rc = self.L.fh_run_chain(str(filter_dir).encode(), samplerate, channels, bits, 1 if gapless else 0, n,
                                 pin, frames, in_format, out_format, pout, oframes, mx, flags, err, 512)
        if rc < 0:
            raise RuntimeError(f"fh_run_chain failed: {err.value.decode()}")
        nout = rc
        res = []
        for k in range(n):
            flat = outs[k].reshape(-1)[: oframes[k] * nout]
            res.append(flat.reshape(oframes[k], nout).copy())
        return res, list(mx), list(flags)

    def run_library(self, config_file, samplerate, channels, chains, gapless=True, slots=4, threads=2,
                    blocks_per_step=1, pcm16=False):
        """chains: list of lists of [frames, channels] float32 arrays (pcm16: int16 arrays = 16-bit
        files in and out, int16 on the wire).  Product only.
        Returns (outputs per chain per file, max values, flags, steps)."""
        if not hasattr(self.L, "fh_run_library"):
            raise RuntimeError("fh_run_library is only in the product harness")
        fn = self.L.fh_run_library_pcm16 if pcm16 else self.L.fh_run_library_tiled
        fn.restype = C.c_int
        dt, ct = (np.int16, C.c_short) if pcm16 else (np.float32, C.c_float)
        files = [np.ascontiguousarray(f, dt) for c in chains for f in c]
        cof = [ci for ci, c in enumerate(chains) for _ in c]
        n = len(files)
        fp = C.POINTER(ct)
        pin = (fp * n)(*[f.ctypes.data_as(fp) for f in files])
        frames = (C.c_long * n)(*[f.shape[0] for f in files])
        outs = [np.zeros((f.shape[0], 64), dt) for f in files]
        pout = (fp * n)(*[o.ctypes.data_as(fp) for o in outs])
        oframes = (C.c_long * n)()
        mx = (C.c_float * n)()
        flags = (C.c_int * n)()
        steps = C.c_long(0)
        fn.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int),
                       C.POINTER(fp), C.POINTER(C.c_long), C.POINTER(fp), C.POINTER(C.c_long),
                       C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_long), C.c_int]
        rc = fn(str(config_file).encode(), samplerate, channels, 1 if gapless else 0, slots, threads, n,
                (C.c_int * n)(*cof), pin, frames, pout, oframes, mx, flags, C.byref(steps), blocks_per_step)
        if rc < 0:
            raise RuntimeError(f"fh_run_library failed ({rc})")
        nout = rc
        res, k = [], 0
        for c in chains:
            row = []
            for _ in c:
                row.append(outs[k].reshape(-1)[: oframes[k] * nout].reshape(oframes[k], nout).copy())
                k += 1
            res.append(row)
        return res, list(mx), list(flags), steps.value

    def load_config(self, config_file, samplerate, channels):
        """-> dict(rc, created, ninp, nout, size, fragm, npar, pairs={(i,o): (state, impulse)})"""
        h = self.L.fh_config_open(str(config_file).encode(), samplerate, channels)
        info = (C.c_int * 7)()
        self.L.fh_config_info(h, info)
        d = dict(zip(["rc", "created", "ninp", "nout", "size", "fragm", "npar"], list(info)))
        d["pairs"] = {}
        if d["created"]:
            cap = d["npar"] * d["fragm"]
            for i in range(d["ninp"]):
                for o in range(d["nout"]):
                    buf = np.zeros(max(cap, 1), np.float32)
                    st = self.L.fh_config_impulse(h, i, o, buf.ctypes.data_as(C.POINTER(C.c_float)), cap)
                    if st > 0:
                        d["pairs"][(i, o)] = (st, buf[:cap].copy())
        self.L.fh_config_close(h)
        return d

    def set_reset_is_fresh(self, on):
        self.L.fh_set_reset_is_fresh(1 if on else 0)

    def drop_pool(self):
        self.L.fh_drop_pool()


def product():
    return Harness(PRODUCT_SO)


def reference():
    return Harness(REFERENCE_SO)


def write_wav(path, data, rate, fmt="pcm16"):
    """data: [frames, channels] float in [-1, 1) (quantised here) or already-int array."""
    data = np.atleast_2d(np.asarray(data))
    if data.shape[0] < data.shape[1] and data.ndim == 2 and data.shape[0] <= 8 and data.shape[1] > 8:
        data = data.T
    frames, ch = data.shape
    if fmt == "pcm16":
        q = np.clip(np.rint(data * 32768.0), -32768, 32767).astype("<i2") if data.dtype.kind == "f" else data.astype("<i2")
        raw, bits, tag = q.tobytes(), 16, 1
    elif fmt == "pcm24":
        q = np.clip(np.rint(data * 8388608.0), -8388608, 8388607).astype(np.int32) if data.dtype.kind == "f" else data.astype(np.int32)
        b = q.astype("<i4").view(np.uint8).reshape(-1, 4)[:, :3]
        raw, bits, tag = b.tobytes(), 24, 1
    elif fmt == "pcm32":
        q = np.clip(np.rint(data * 2147483648.0), -2**31, 2**31 - 1).astype("<i4") if data.dtype.kind == "f" else data.astype("<i4")
        raw, bits, tag = q.tobytes(), 32, 1
    elif fmt == "float":
        raw, bits, tag = data.astype("<f4").tobytes(), 32, 3
    else:
        raise ValueError(fmt)
    block = ch * bits // 8
    hdr = b"RIFF" + struct.pack("<I", 36 + len(raw)) + b"WAVE" + b"fmt " + struct.pack(
        "<IHHIIHH", 16, tag, ch, rate, rate * block, block, bits) + b"data" + struct.pack("<I", len(raw))
    with open(path, "wb") as f:
        f.write(hdr + raw + (b"\0" if len(raw) & 1 else b""))
