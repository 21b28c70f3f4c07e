"""GPU: BASELINE.json's full-size shapes through size-independent properties
(the oracle is too slow to run 1024 streams): linearity, stream independence,
agreement of the batched / time-tiled paths with the single-stream path that the
oracle tests pin, impulse -> filter, and a long gapless-style run."""
import numpy as np
import pytest

from folve_b200 import capi, workloads

pytestmark = pytest.mark.gpu


def _load(name):
    wl = workloads.WORKLOADS[name]()
    return wl, wl.load(capi.Filter(wl.ninp, wl.nout, wl.size, wl.fragm)).commit(0)


@pytest.mark.parametrize("T", [4, 8])
def test_santalucia_1024_streams_linearity_and_independence(T):
    """T = 8 is the bench configuration: the time-tiled MAC runs as a persistent grid there (32768 work items)."""
    wl, f = _load("santalucia")
    N, B = wl.fragm, 1024
    assert f.partitions == 25 and f.ring_depth == 22 and f.active_rows == 44
    r = np.random.default_rng(1)
    nsteps = 32 // T                             # 32 blocks > 22 partitions: the whole ring is exercised
    xa = r.uniform(-0.02, 0.02, (nsteps, 16, T * N, 2)).astype(np.float32)
    xb = r.uniform(-0.02, 0.02, (nsteps, 16, T * N, 2)).astype(np.float32)
    bt = capi.Batch(f, B, blocks_per_step=T)
    ya, yb, yab, ydup = [], [], [], []
    for k in range(nsteps):
        bt.host_in[:] = 0
        bt.host_in[0:16] = xa[k]                 # a
        bt.host_in[16:32] = xb[k]                # b
        bt.host_in[32:48] = xa[k] + xb[k]        # a + b
        bt.host_in[1000:1016] = xa[k]            # a again, far away in the batch (another chunk / CUDA stream)
        bt.process()
        o = bt.host_out
        ya.append(o[0:16].copy()); yb.append(o[16:32].copy()); yab.append(o[32:48].copy()); ydup.append(o[1000:1016].copy())
        assert not np.any(o[48:1000])            # silent streams stay silent: no cross-talk between streams
    ya, yb, yab, ydup = (np.concatenate(v, axis=1) for v in (ya, yb, yab, ydup))
    assert np.array_equal(ya, ydup)              # same input, same output, wherever the stream sits
    fs = max(1.0, np.abs(yab).max())
    assert np.abs(yab - (ya + yb)).max() / fs < 1e-5     # linearity
    # the batched time-tiled path equals the synchronous single-stream path (pinned by the oracle tests)
    s = capi.Stream(f)
    x0 = np.concatenate([xa[k][3] for k in range(nsteps)], axis=0)
    y0 = np.concatenate([s.process(x0[i:i + N]) for i in range(0, x0.shape[0], N)], axis=0)
    assert np.abs(y0 - ya[3]).max() < 2e-6
    s.close(); bt.close(); f.close()


@pytest.mark.parametrize("name", ["lowpass", "roomcorr96", "roomcorr192", "crossfeed", "surround51", "surround51_dense"])
def test_every_baseline_config_impulse_returns_the_filter(name):
    """delta in on every input at once -> out[o] = sum_i h[i][o], sample-exact positions."""
    wl, f = _load(name)
    N = wl.fragm
    npart = f.partitions
    blocks = npart + 1
    x = np.zeros((blocks * N, wl.ninp), np.float32)
    x[0, :] = 1.0
    s = capi.Stream(f)
    y = np.concatenate([s.process(x[i:i + N]) for i in range(0, x.shape[0], N)], axis=0)
    want = np.zeros((blocks * N, wl.nout), np.float64)
    pair_h = {}
    for (i, o, d, i0) in wl.adds:
        h = pair_h.setdefault((i, o), np.zeros(blocks * N))
        h[i0:i0 + len(d)] += d
    for (i1, o1, i2, o2) in wl.links:
        pair_h[(i2, o2)] = pair_h[(i1, o1)]
    for (i, o), h in pair_h.items():
        want[:, o] += h
    scale = max(1e-3, np.abs(want).max())
    assert np.abs(y - want).max() / scale < 2e-5
    s.close(); f.close()


def test_long_run_state_does_not_drift():
    """300 s of audio (1616 blocks) through one stream in 8-block steps: the last blocks
    still match a fresh stream fed only the last ring-depth+1 blocks of history."""
    wl, f = _load("santalucia")
    N, T = wl.fragm, 8
    r = np.random.default_rng(2)
    nblocks = 1616
    bt = capi.Batch(f, 2, blocks_per_step=T)
    x = r.uniform(-0.03, 0.03, (nblocks * N, 2)).astype(np.float32)
    last = None
    for k in range(nblocks // T):
        bt.host_in[0] = x[k * T * N:(k + 1) * T * N]
        bt.host_in[1] = 0
        bt.process()
        last = bt.host_out[0].copy()
    # history needed by the last T blocks: 25 partitions back
    hist = 25 + T
    s = capi.Stream(f)
    xs = x[(nblocks - hist) * N:]
    ys = np.concatenate([s.process(xs[i:i + N]) for i in range(0, xs.shape[0], N)], axis=0)
    assert np.abs(ys[-T * N:] - last).max() < 2e-6
    assert bt.get_max()[1] == 0.0
    s.close(); bt.close(); f.close()
