"""ctypes binding of oracle/libzita_oracle.so plus the float64 ground truth.

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU
baseline legs import this.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "oracle", "libzita_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_SO)
        vp, u, i, f = C.c_void_p, C.c_uint32, C.c_int32, C.c_float
        fp = C.POINTER(C.c_float)
        L.zo_new.restype = vp
        L.zo_delete.argtypes = [vp]
        L.zo_set_reset_is_fresh.argtypes = [vp, C.c_int]
        L.zo_configure.argtypes = [vp, u, u, u, u, u, u, f]
        L.zo_impdata_create.argtypes = [vp, u, u, i, fp, i, i]
        L.zo_impdata_link.argtypes = [vp, u, u, u, u]
        L.zo_reset.argtypes = [vp]
        L.zo_start_process.argtypes = [vp, C.c_int, C.c_int]
        L.zo_process.argtypes = [vp]
        L.zo_stop_process.argtypes = [vp]
        L.zo_cleanup.argtypes = [vp]
        L.zo_state.argtypes = [vp]
        L.zo_inpdata.argtypes = [vp, u]
        L.zo_inpdata.restype = fp
        L.zo_outdata.argtypes = [vp, u]
        L.zo_outdata.restype = fp
        L.zo_parsize.argtypes = [vp]
        L.zo_parsize.restype = u
        L.zo_npar.argtypes = [vp]
        L.zo_npar.restype = u
        L.zo_ptind.argtypes = [vp]
        L.zo_ptind.restype = u
        L.zo_fftb.argtypes = [vp, u, u, u]
        L.zo_fftb.restype = fp
        L.zo_ffta.argtypes = [vp, u, u]
        L.zo_ffta.restype = fp
        _lib = L
    return _lib


def fragm_for(size: int) -> int:
    """zita-fconfig.cc:74-77: MAXQUANT halved while > MINPART and >= 2*size."""
    fragm = 8192
    while fragm > 64 and fragm >= 2 * size:
        fragm //= 2
    return fragm


class OracleConvproc:
    """The Convproc facade driven exactly as sound-processor.cc drives it."""

    def __init__(self, ninp, nout, size, fragm=None, reset_is_fresh=False):
        L = lib()
        self.ninp, self.nout, self.size = ninp, nout, size
        self.fragm = fragm if fragm is not None else fragm_for(size)
        self._h = L.zo_new()
        L.zo_set_reset_is_fresh(self._h, 1 if reset_is_fresh else 0)
        rc = L.zo_configure(self._h, ninp, nout, size, self.fragm, self.fragm, self.fragm, 0.0)
        if rc:
            raise RuntimeError(f"zo_configure -> {rc}")
        self.started = False

    def add(self, inp, out, data, ind0, step=1, ind1=None):
        data = np.ascontiguousarray(data, dtype=np.float32)
        if ind1 is None:
            ind1 = ind0 + (len(data) + step - 1) // step
        rc = lib().zo_impdata_create(self._h, inp, out, step, data.ctypes.data_as(C.POINTER(C.c_float)), ind0, ind1)
        if rc:
            raise RuntimeError(f"zo_impdata_create -> {rc}")

    def link(self, inp1, out1, inp2, out2):
        rc = lib().zo_impdata_link(self._h, inp1, out1, inp2, out2)
        if rc:
            raise RuntimeError(f"zo_impdata_link -> {rc}")

    def reset(self):
        """SoundProcessor::Reset (sound-processor.cc:139-145)."""
        L = lib()
        L.zo_reset(self._h)
        L.zo_start_process(self._h, 0, 0)
        self.started = True

    def process(self, block: np.ndarray) -> np.ndarray:
        """SoundProcessor::Process (sound-processor.cc:98-127) for one block of
        `frames` <= fragm frames: copy the valid frames per channel, process,
        read back the same number of frames."""
        if not self.started:
            self.reset()
        L = lib()
        frames = block.shape[0]
        block = np.asarray(block, np.float32)
        for ch in range(self.ninp):
            dst = np.ctypeslib.as_array(L.zo_inpdata(self._h, ch), shape=(self.fragm,))
            dst[:frames] = block[:, ch]
        L.zo_process(self._h)
        out = np.zeros((frames, self.nout), np.float32)
        for ch in range(self.nout):
            src = np.ctypeslib.as_array(L.zo_outdata(self._h, ch), shape=(self.fragm,))
            out[:, ch] = src[:frames]
        return out

    def fftb(self, inp, out, j):
        p = lib().zo_fftb(self._h, inp, out, j)
        if not p:
            return None
        a = np.ctypeslib.as_array(p, shape=(2 * (self.fragm + 1),))
        return a[0::2] + 1j * a[1::2]

    def ffta(self, inp, slot):
        p = lib().zo_ffta(self._h, inp, slot)
        if not p:
            return None
        a = np.ctypeslib.as_array(p, shape=(2 * (self.fragm + 1),))
        return a[0::2] + 1j * a[1::2]

    @property
    def npar(self):
        return lib().zo_npar(self._h)

    @property
    def ptind(self):
        return lib().zo_ptind(self._h)

    def close(self):
        if self._h:
            lib().zo_delete(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run_blocks(proc, x: np.ndarray, fragm: int) -> np.ndarray:
    """Feed x [frames, ninp] block by block (last block short), as
    ConvolveFileHandler::AddMoreSoundData does for one file
    (convolve-file-handler.cc:370-424)."""
    outs = []
    for s in range(0, x.shape[0], fragm):
        outs.append(proc.process(x[s:s + fragm]))
    return np.concatenate(outs, axis=0) if outs else np.zeros((0, 0), np.float32)


def truth_f64(x: np.ndarray, h: dict, nout: int) -> np.ndarray:
    """float64 ground truth: y[o] = sum_i h[i,o] * x[i], truncated to len(x).
    h maps (inp, out) -> 1-D float array (the summed impulse of that pair)."""
    from scipy.signal import fftconvolve
    n = x.shape[0]
    y = np.zeros((n, nout), np.float64)
    for (i, o), taps in h.items():
        taps = np.asarray(taps, np.float64)
        if not taps.size or not np.any(taps):
            continue
        last = np.flatnonzero(taps)[-1] + 1
        y[:, o] += fftconvolve(x[:, i].astype(np.float64), taps[:last])[:n]
    return y


class FilterSpec:
    """A filter described once and loaded identically into the oracle and the engine."""

    def __init__(self, ninp, nout, size):
        self.ninp, self.nout, self.size = ninp, nout, size
        self.fragm = fragm_for(size)
        self.ops = []  # ("add", inp, out, data, ind0) | ("link", i1, o1, i2, o2)

    def add(self, inp, out, data, ind0=0):
        self.ops.append(("add", inp, out, np.asarray(data, np.float32), ind0))
        return self

    def link(self, i1, o1, i2, o2):
        self.ops.append(("link", i1, o1, i2, o2))
        return self

    def load(self, target):
        for op in self.ops:
            if op[0] == "add":
                target.add(op[1], op[2], op[3], op[4])
            else:
                target.link(*op[1:])
        return target

    def impulses(self):
        """Effective time-domain impulse per pair (float64), following zita's
        add / link semantics."""
        npar = (self.size + self.fragm - 1) // self.fragm
        total = npar * self.fragm
        own, link, exists = {}, {}, set()
        for op in self.ops:
            if op[0] == "add":
                _, i, o, d, i0 = op
                if i0 >= total or i0 + len(d) <= 0:
                    continue
                exists.add((i, o))
                if (i, o) in link:
                    continue
                a = own.setdefault((i, o), np.zeros(total, np.float64))
                n = min(len(d), total - i0)
                a[i0:i0 + n] += d[:n].astype(np.float64)
            else:
                _, i1, o1, i2, o2 = op
                if (i1, o1) not in exists:
                    continue
                exists.add((i2, o2))
                own.pop((i2, o2), None)
                link[(i2, o2)] = (i1, o1)
        res = dict(own)
        for dst, src in link.items():
            if src in own and src not in link:
                res[dst] = own[src]
        return res
