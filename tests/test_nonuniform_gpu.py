"""GPU: non-uniform partitioning (north star) -- a head of small
partitions evaluated every `quantum` frames plus a tail of `maxpart`-sized partitions, i.e.
Convproc::configure(ninp, nout, size, quantum, quantum, maxpart).  The output is that of the uniform
engine (fragm = maxpart, what folve configures: /root/reference/zita-fconfig.cc:74-93) and of the
float64 truth, delivered in blocks of `quantum` frames."""
import numpy as np
import pytest

from folve_b200 import capi
from oracle_py import FilterSpec, OracleConvproc, run_blocks, truth_f64

pytestmark = pytest.mark.gpu


def _spec(seed, nin=2, nout=2, size=60000):
    r = np.random.default_rng(seed)
    spec = FilterSpec(nin, nout, size)
    for i in range(nin):
        for o in range(nout):
            if i == o or (i + o) % 3 == 0:
                n = size - 700 - 1000 * i
                spec.add(i, o, r.standard_normal(n) * np.exp(-4.0 * np.arange(n) / n) * 0.01, 300 + 17 * o)
        spec.add(i, i % nout, [0.4], 0)
    return spec


@pytest.mark.parametrize("quantum", [256, 1024, 4096, 8192])
def test_nonuniform_equals_uniform_and_truth(quantum):
    spec = _spec(quantum)
    x = np.random.default_rng(1).uniform(-0.25, 0.25, (3 * 8192 + 5 * quantum + 77, 2)).astype(np.float32)
    nu = spec.load(capi.NuFilter(2, 2, spec.size, quantum, 8192)).commit(0)
    if quantum < 8192:
        assert nu.head_partitions == 8192 // quantum and nu.tail_partitions == (spec.size - 8192 + 8191) // 8192
    else:   # quantum == maxpart is the uniform engine: one level
        assert nu.head_partitions == (spec.size + 8191) // 8192 and nu.tail_partitions == 0
    s = capi.NuStream(nu)
    y = run_blocks(s, x, quantum)
    uf = spec.load(capi.Filter(2, 2, spec.size, 8192)).commit(0)
    us = capi.Stream(uf)
    yu = run_blocks(us, x, 8192)
    t = truth_f64(x, spec.impulses(), 2)
    fs = max(1.0, float(np.abs(t).max()))
    assert y.shape == yu.shape == t.shape
    assert np.abs(y - t).max() / fs < 1e-5
    assert np.abs(y - yu).max() / fs < 1e-5
    assert np.abs(np.rint(y * 32767.0) - np.rint(yu * 32767.0)).max() <= 1
    assert s.max_value == pytest.approx(max(0.0, float(y.max())), abs=0)
    # reset == fresh (whole blocks: the rounding of a block's output depends on everything in the block)
    s.reset()
    y2 = run_blocks(s, x[: 2 * 8192 + quantum], quantum)
    assert np.array_equal(y2, y[: 2 * 8192 + quantum])
    for h in (s, us, nu, uf):
        h.close()


def test_nonuniform_short_filter_links_and_oracle():
    """a filter that fits the head level (no tail), a 3 x 2 matrix with /impulse/copy links whose source gets
    data after the link was made, against the CPU oracle"""
    r = np.random.default_rng(5)
    for size, quantum, maxpart in ((3000, 256, 4096), (20000, 512, 4096)):
        spec = FilterSpec(3, 2, size)
        spec.add(0, 0, r.standard_normal(size - 100) * 0.01, 50)
        spec.link(0, 0, 1, 1)                       # (1,1) uses the spectra of (0,0) ...
        spec.add(0, 0, r.standard_normal(500) * 0.01, size - 600)   # ... also what is added later, in head or tail
        spec.add(2, 1, [0.5], 3)
        x = r.uniform(-0.5, 0.5, (2 * maxpart + 3 * quantum + 11, 3)).astype(np.float32)
        nu = spec.load(capi.NuFilter(3, 2, size, quantum, maxpart)).commit(0)
        s = capi.NuStream(nu)
        y = run_blocks(s, x, quantum)
        o = OracleConvproc(3, 2, size, reset_is_fresh=True)
        # the oracle follows folve's fragm rule (zita-fconfig.cc:74-77); only compare where it agrees with maxpart
        yo = run_blocks(spec.load(o), x, o.fragm)
        t = truth_f64(x, spec.impulses(), 2)
        fs = max(1.0, float(np.abs(t).max()))
        assert np.abs(y - t).max() / fs < 1e-5
        if o.fragm == maxpart:
            assert np.abs(y - yo).max() / fs < 1e-5
        s.close()
        nu.close()
