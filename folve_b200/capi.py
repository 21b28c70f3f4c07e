"""ctypes binding of the C ABI in include/folve_b200.h (libfolve_b200.so).

This is plumbing for tests/ and bench.py -- the product is the shared library
and the C++ host layer in folve_b200/host/.  There is no fallback: if the
library is missing or no sm_100 device is usable, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfolve_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "folve_b200.h")

PCM_F32, PCM_S16, PCM_S24 = 0, 1, 2
_PCM_DTYPE = {PCM_F32: np.float32, PCM_S16: np.int16, PCM_S24: np.int32}

_lib = None


class FcvError(RuntimeError):
    pass


def declared_symbols() -> list[str]:
    """Every function name include/folve_b200.h declares."""
    src = open(HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fcv_[a-z0-9_]+)\s*\(", src)))


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FcvError(
            f"{LIB_PATH} is missing: run `make` (or __graft_entry__.build()); "
            "there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, i, u, f = C.c_void_p, C.c_int, C.c_uint, C.c_float
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
    sig = {
        "fcv_abi_version": (i, []),
        "fcv_last_error": (C.c_char_p, []),
        "fcv_device_count": (i, []),
        "fcv_filter_begin": (vp, [i, i, u, u]),
        "fcv_filter_add": (i, [vp, i, i, i, fp, i, i]),
        "fcv_filter_link": (i, [vp, i, i, i, i]),
        "fcv_filter_commit": (i, [vp, i]),
        "fcv_filter_ref": (None, [vp]),
        "fcv_filter_unref": (None, [vp]),
        "fcv_filter_ninp": (i, [vp]),
        "fcv_filter_nout": (i, [vp]),
        "fcv_filter_fragm": (i, [vp]),
        "fcv_filter_partitions": (i, [vp]),
        "fcv_filter_ring_depth": (i, [vp]),
        "fcv_filter_active_rows": (i, [vp]),
        "fcv_filter_active_pairs": (i, [vp]),
        "fcv_filter_device": (i, [vp]),
        "fcv_stream_create": (vp, [vp]),
        "fcv_stream_destroy": (None, [vp]),
        "fcv_stream_reset": (i, [vp]),
        "fcv_stream_buffer": (fp, [vp]),
        "fcv_stream_process": (i, [vp, i, fp]),
        "fcv_stream_create_fmt": (vp, [vp, i, i]),
        "fcv_stream_buffer_bytes": (C.c_size_t, [vp]),
        "fcv_stream_submit": (i, [vp, i]),
        "fcv_stream_await": (i, [vp, fp]),
        "fcv_batch_set_copy_only": (i, [vp, i]),
        "fcv_stream_filter": (vp, [vp]),
        "fcv_nufilter_begin": (vp, [i, i, u, u, u]),
        "fcv_nufilter_add": (i, [vp, i, i, i, fp, i, i]),
        "fcv_nufilter_link": (i, [vp, i, i, i, i]),
        "fcv_nufilter_commit": (i, [vp, i]),
        "fcv_nufilter_ref": (None, [vp]),
        "fcv_nufilter_unref": (None, [vp]),
        "fcv_nufilter_quantum": (i, [vp]),
        "fcv_nufilter_head_partitions": (i, [vp]),
        "fcv_nufilter_tail_partitions": (i, [vp]),
        "fcv_nustream_create": (vp, [vp]),
        "fcv_nustream_destroy": (None, [vp]),
        "fcv_nustream_reset": (i, [vp]),
        "fcv_nustream_buffer": (fp, [vp]),
        "fcv_nustream_process": (i, [vp, i, fp]),
        "fcv_batch_create": (vp, [vp, i, i, i]),
        "fcv_batch_create_tiled": (vp, [vp, i, i, i, i]),
        "fcv_batch_blocks_per_step": (i, [vp]),
        "fcv_batch_destroy": (None, [vp]),
        "fcv_batch_nstreams": (i, [vp]),
        "fcv_batch_host_in": (vp, [vp]),
        "fcv_batch_host_out": (vp, [vp]),
        "fcv_batch_host_in_bytes": (C.c_size_t, [vp]),
        "fcv_batch_host_out_bytes": (C.c_size_t, [vp]),
        "fcv_batch_device_in": (vp, [vp]),
        "fcv_batch_device_out": (vp, [vp]),
        "fcv_batch_process": (i, [vp, ip]),
        "fcv_batch_host_in_slot": (vp, [vp, i]),
        "fcv_batch_host_out_slot": (vp, [vp, i]),
        "fcv_batch_submit": (i, [vp, i, ip]),
        "fcv_batch_wait": (i, [vp, i]),
        "fcv_batch_process_device": (i, [vp, ip]),
        "fcv_batch_sync": (i, [vp]),
        "fcv_batch_reset_slot": (i, [vp, i]),
        "fcv_batch_get_max": (i, [vp, fp]),
        "fcv_batch_get_block_max": (i, [vp, fp]),
        "fcv_batch_reset_slot_async": (i, [vp, i]),
        "fcv_batch_host_block_max_slot": (vp, [vp, i]),
        "fcv_batch_cuda_stream": (vp, [vp]),
        "fcv_batch_event_record": (i, [vp, i]),
        "fcv_batch_event_elapsed_ms": (i, [vp, i, i, fp]),
        "fcv_batch_set_profiling": (i, [vp, i]),
        "fcv_batch_profile": (i, [vp, fp, ip]),
        "fcv_kernel_launches": (C.c_ulonglong, []),
        "fcv_filter_get_spectrum": (i, [vp, i, i, i, fp]),
        "fcv_stream_get_input_spectrum": (i, [vp, i, i, fp]),
        "fcv_filter_get_impulse": (i, [vp, i, i, fp, i]),
        "fcv_debug_set_fused": (None, [i]),
        "fcv_debug_fused_launches": (C.c_ulonglong, []),
        "fcv_debug_set_inv_pair": (None, [i]),
        "fcv_debug_set_tmem": (None, [i]),
        "fcv_debug_get_tmem": (i, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _check(rc: int) -> int:
    if rc < 0:
        raise FcvError(f"fcv error {rc}: {lib().fcv_last_error().decode()}")
    return rc


def _fp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class Filter:
    """Mirror of the configure / impdata_* part of a Convproc."""

    def __init__(self, ninp: int, nout: int, size: int, fragm: int):
        self._h = lib().fcv_filter_begin(ninp, nout, size, fragm)
        if not self._h:
            raise FcvError(lib().fcv_last_error().decode())
        self.ninp, self.nout, self.size, self.fragm = ninp, nout, size, fragm

    def add(self, inp: int, out: int, data, ind0: int, step: int = 1, ind1: int | None = None):
        data = np.ascontiguousarray(data, dtype=np.float32)
        if ind1 is None:
            ind1 = ind0 + (len(data) + step - 1) // step
        _check(lib().fcv_filter_add(self._h, inp, out, step, _fp(data), ind0, ind1))

    def link(self, inp1, out1, inp2, out2):
        _check(lib().fcv_filter_link(self._h, inp1, out1, inp2, out2))

    def commit(self, device: int = 0):
        _check(lib().fcv_filter_commit(self._h, device))
        return self

    @property
    def partitions(self): return lib().fcv_filter_partitions(self._h)
    @property
    def ring_depth(self): return lib().fcv_filter_ring_depth(self._h)
    @property
    def active_rows(self): return lib().fcv_filter_active_rows(self._h)
    @property
    def active_pairs(self): return lib().fcv_filter_active_pairs(self._h)

    def spectrum(self, inp, out, j):
        dst = np.zeros(2 * (self.fragm + 1), np.float32)
        rc = _check(lib().fcv_filter_get_spectrum(self._h, inp, out, j, _fp(dst)))
        return (dst[0::2] + 1j * dst[1::2]) if rc == 1 else None

    def close(self):
        if self._h:
            lib().fcv_filter_unref(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Stream:
    """Mirror of one running Convproc as SoundProcessor drives it."""

    def __init__(self, flt: Filter, in_format=PCM_F32, out_format=PCM_F32):
        self.flt = flt
        self.in_format, self.out_format = in_format, out_format
        self._h = lib().fcv_stream_create_fmt(flt._h, in_format, out_format)
        if not self._h:
            raise FcvError(lib().fcv_last_error().decode())
        nbytes = lib().fcv_stream_buffer_bytes(self._h)
        raw = (C.c_char * nbytes).from_address(C.cast(lib().fcv_stream_buffer(self._h), C.c_void_p).value)
        self.raw = np.frombuffer(raw, dtype=np.uint8)
        self.buffer = self.raw.view(np.float32)          # the float view SoundProcessor uses
        self.in_view = self.raw.view(_PCM_DTYPE[in_format])
        self.out_view = self.raw.view(_PCM_DTYPE[out_format])
        self.max_value = 0.0

    def process(self, block: np.ndarray) -> np.ndarray:
        """block: [frames<=fragm, ninp] float32 -> [frames, nout]; rest of the block is zero."""
        frames = block.shape[0]
        f = self.flt
        self.buffer[: frames * f.ninp] = np.ascontiguousarray(block, np.float32).reshape(-1)
        m = C.c_float(self.max_value)
        _check(lib().fcv_stream_process(self._h, frames, C.byref(m)))
        self.max_value = m.value
        return self.buffer[: frames * f.nout].reshape(frames, f.nout).copy()

    def submit(self, block: np.ndarray):
        """queue one block ([frames<=fragm, ninp] in the stream's input wire format) without waiting"""
        frames = block.shape[0]
        self.in_view[: frames * self.flt.ninp] = np.ascontiguousarray(block, self.in_view.dtype).reshape(-1)
        self._frames = frames
        _check(lib().fcv_stream_submit(self._h, frames))

    def wait(self) -> np.ndarray:
        m = C.c_float(self.max_value)
        _check(lib().fcv_stream_await(self._h, C.byref(m)))
        self.max_value = m.value
        f = self.flt
        return self.out_view[: self._frames * f.nout].reshape(self._frames, f.nout).copy()

    def reset(self):
        _check(lib().fcv_stream_reset(self._h))
        self.max_value = 0.0

    def input_spectrum(self, inp, age=0):
        dst = np.zeros(2 * (self.flt.fragm + 1), np.float32)
        _check(lib().fcv_stream_get_input_spectrum(self._h, inp, age, _fp(dst)))
        return dst[0::2] + 1j * dst[1::2]

    def close(self):
        if self._h:
            lib().fcv_stream_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class NuFilter:
    """Non-uniform partitioning (Convproc::configure with quantum = minpart < maxpart)."""

    def __init__(self, ninp: int, nout: int, size: int, quantum: int, maxpart: int):
        self._h = lib().fcv_nufilter_begin(ninp, nout, size, quantum, maxpart)
        if not self._h:
            raise FcvError(lib().fcv_last_error().decode())
        self.ninp, self.nout, self.size, self.quantum, self.maxpart = ninp, nout, size, quantum, maxpart
        self.fragm = quantum     # block size of the streams (what run_blocks steps by)

    def add(self, inp, out, data, ind0, step=1, ind1=None):
        data = np.ascontiguousarray(data, dtype=np.float32)
        if ind1 is None:
            ind1 = ind0 + (len(data) + step - 1) // step
        _check(lib().fcv_nufilter_add(self._h, inp, out, step, _fp(data), ind0, ind1))

    def link(self, inp1, out1, inp2, out2):
        _check(lib().fcv_nufilter_link(self._h, inp1, out1, inp2, out2))

    def commit(self, device: int = 0):
        _check(lib().fcv_nufilter_commit(self._h, device))
        return self

    @property
    def head_partitions(self): return lib().fcv_nufilter_head_partitions(self._h)
    @property
    def tail_partitions(self): return lib().fcv_nufilter_tail_partitions(self._h)

    def close(self):
        if self._h:
            lib().fcv_nufilter_unref(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class NuStream:
    def __init__(self, flt: NuFilter):
        self.flt = flt
        self._h = lib().fcv_nustream_create(flt._h)
        if not self._h:
            raise FcvError(lib().fcv_last_error().decode())
        n = flt.quantum * max(flt.ninp, flt.nout)
        self.buffer = np.ctypeslib.as_array(lib().fcv_nustream_buffer(self._h), shape=(n,))
        self.max_value = 0.0

    def process(self, block: np.ndarray) -> np.ndarray:
        frames = block.shape[0]
        f = self.flt
        self.buffer[: frames * f.ninp] = np.ascontiguousarray(block, np.float32).reshape(-1)
        m = C.c_float(self.max_value)
        _check(lib().fcv_nustream_process(self._h, frames, C.byref(m)))
        self.max_value = m.value
        return self.buffer[: frames * f.nout].reshape(frames, f.nout).copy()

    def reset(self):
        _check(lib().fcv_nustream_reset(self._h))
        self.max_value = 0.0

    def close(self):
        if self._h:
            lib().fcv_nustream_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Batch:
    def __init__(self, flt: Filter, nstreams: int, in_format=PCM_F32, out_format=PCM_F32, blocks_per_step=1):
        self.flt = flt
        self.n = nstreams
        self.blocks_per_step = blocks_per_step
        self._h = lib().fcv_batch_create_tiled(flt._h, nstreams, in_format, out_format, blocks_per_step)
        if not self._h:
            raise FcvError(lib().fcv_last_error().decode())
        L = lib()
        N = flt.fragm * blocks_per_step
        din, dout = _PCM_DTYPE[in_format], _PCM_DTYPE[out_format]
        self.in_bytes = L.fcv_batch_host_in_bytes(self._h)
        self.out_bytes = L.fcv_batch_host_out_bytes(self._h)
        ibuf = (C.c_char * self.in_bytes).from_address(L.fcv_batch_host_in(self._h))
        obuf = (C.c_char * self.out_bytes).from_address(L.fcv_batch_host_out(self._h))
        self.host_in = np.frombuffer(ibuf, dtype=din).reshape(nstreams, N, flt.ninp)
        self.host_out = np.frombuffer(obuf, dtype=dout).reshape(nstreams, N, flt.nout)

    def _fv(self, frames_valid):
        if frames_valid is None:
            return None, None
        a = np.ascontiguousarray(frames_valid, dtype=np.int32)
        assert a.shape == (self.n,)
        return a, a.ctypes.data_as(C.POINTER(C.c_int))

    def process(self, frames_valid=None):
        keep, p = self._fv(frames_valid)
        _check(lib().fcv_batch_process(self._h, p))

    def slot_views(self, slot):
        """numpy views of the pinned in/out staging of `slot` (0 or 1)."""
        L = lib()
        ibuf = (C.c_char * self.in_bytes).from_address(L.fcv_batch_host_in_slot(self._h, slot))
        obuf = (C.c_char * self.out_bytes).from_address(L.fcv_batch_host_out_slot(self._h, slot))
        return (np.frombuffer(ibuf, dtype=self.host_in.dtype).reshape(self.host_in.shape),
                np.frombuffer(obuf, dtype=self.host_out.dtype).reshape(self.host_out.shape))

    def submit(self, slot, frames_valid=None):
        keep, p = self._fv(frames_valid)
        _check(lib().fcv_batch_submit(self._h, slot, p))

    def wait(self, slot):
        _check(lib().fcv_batch_wait(self._h, slot))

    def process_device(self, frames_valid=None):
        keep, p = self._fv(frames_valid)
        _check(lib().fcv_batch_process_device(self._h, p))

    def sync(self):
        _check(lib().fcv_batch_sync(self._h))

    def reset_slot(self, slot):
        _check(lib().fcv_batch_reset_slot(self._h, slot))

    def get_max(self):
        m = np.zeros(self.n, np.float32)
        _check(lib().fcv_batch_get_max(self._h, _fp(m)))
        return m

    def get_block_max(self):
        """[nstreams, blocks_per_step]: signed maximum of every block of the last step"""
        m = np.zeros((self.n, self.blocks_per_step), np.float32)
        _check(lib().fcv_batch_get_block_max(self._h, _fp(m)))
        return m

    @property
    def cuda_stream(self): return lib().fcv_batch_cuda_stream(self._h)
    @property
    def device_in(self): return lib().fcv_batch_device_in(self._h)
    @property
    def device_out(self): return lib().fcv_batch_device_out(self._h)

    def event_record(self, slot):
        _check(lib().fcv_batch_event_record(self._h, slot))

    def event_elapsed_ms(self, slot0, slot1):
        ms = C.c_float(0)
        _check(lib().fcv_batch_event_elapsed_ms(self._h, slot0, slot1, C.byref(ms)))
        return ms.value

    def set_copy_only(self, on):
        _check(lib().fcv_batch_set_copy_only(self._h, 1 if on else 0))

    def set_profiling(self, on):
        _check(lib().fcv_batch_set_profiling(self._h, 1 if on else 0))

    def profile(self):
        ms = (C.c_float * 3)()
        n = C.c_int(0)
        _check(lib().fcv_batch_profile(self._h, ms, C.byref(n)))
        return [ms[0], ms[1], ms[2]], n.value

    def close(self):
        if self._h:
            lib().fcv_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
