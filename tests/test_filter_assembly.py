"""CPU: the host half of the C ABI's filter object -- fcv_filter_begin / fcv_filter_add / fcv_filter_link, which
replace Convproc::configure / impdata_create / impdata_link (zita-config.cc:163,203,252,274) -- against the
add / link rules the oracle is tested with (tests/oracle_py.py FilterSpec.impulses; tests/test_oracle.py runs the
same random structures through the oracle against float64 convolution).  Nothing here touches a GPU: the filter
is assembled on the host and read back with fcv_filter_get_impulse before any commit."""
import ctypes as C

import numpy as np
import pytest

from folve_b200 import capi
from oracle_py import FilterSpec


def _random_spec(r):
    ninp, nout = int(r.integers(1, 5)), int(r.integers(1, 5))
    size = int(r.choice([40, 64, 100, 129, 300, 700, 1500, 3000, 6000, 9000]))
    spec = FilterSpec(ninp, nout, size)
    pairs = [(i, o) for i in range(ninp) for o in range(nout)]
    for _ in range(int(r.integers(1, 9))):
        i, o = pairs[int(r.integers(len(pairs)))]
        if r.random() < 0.25:
            i2, o2 = pairs[int(r.integers(len(pairs)))]
            if (i2, o2) != (i, o):
                spec.link(i, o, i2, o2)          # before or after the source has data, chains, re-linked targets
            continue
        taps = int(r.integers(1, size + 1))
        spec.add(i, o, r.standard_normal(taps) * 0.3 / np.sqrt(taps), int(r.integers(0, size)))
    return spec, pairs


def test_random_add_and_link_sequences_assemble_the_same_impulses():
    L = capi.lib()
    linked = silent = 0
    for seed in range(300):
        spec, pairs = _random_spec(np.random.default_rng(1000 + seed))
        f = spec.load(capi.Filter(spec.ninp, spec.nout, spec.size, spec.fragm))
        want = spec.impulses()
        cap = (spec.size + spec.fragm - 1) // spec.fragm * spec.fragm
        assert f.partitions == cap // spec.fragm
        for (i, o) in pairs:
            buf = np.zeros(cap, np.float32)
            st = L.fcv_filter_get_impulse(f._h, i, o, buf.ctypes.data_as(C.POINTER(C.c_float)), cap)
            got = buf.astype(np.float64) * 2 * spec.fragm          # stored with zita's 1 / (2 fragm)
            if (i, o) in want:
                assert np.abs(got - want[(i, o)]).max() < 1e-6, (seed, i, o, st)
            else:
                assert not np.any(got), (seed, i, o, st)
                silent += 1
            linked += st == 2
        f.close()
    assert linked > 50 and silent > 500


def test_arguments_out_of_range_are_refused():
    """impdata_create / impdata_link with a channel out of range or an inverted interval are errors (Converror::
    BAD_PARAM in zita-convolver); nothing is stored."""
    L = capi.lib()
    f = capi.Filter(2, 1, 100, 64)
    d = np.ones(4, np.float32)
    p = d.ctypes.data_as(C.POINTER(C.c_float))
    for (i, o) in ((2, 0), (0, 1), (-1, 0)):
        assert L.fcv_filter_add(f._h, i, o, 1, p, 0, 4) != 0
    assert L.fcv_filter_link(f._h, 0, 0, 0, 0) != 0                 # a pair cannot be its own source
    assert L.fcv_filter_link(f._h, 0, 0, 2, 0) != 0
    buf = np.zeros(128, np.float32)
    assert L.fcv_filter_get_impulse(f._h, 0, 0, buf.ctypes.data_as(C.POINTER(C.c_float)), 128) == 0 and not buf.any()
    with pytest.raises(capi.FcvError):
        capi.Filter(0, 1, 100, 64)
    with pytest.raises(capi.FcvError):
        capi.Filter(1, 1, 100, 100)                                  # block size must be a power of two
    f.close()
