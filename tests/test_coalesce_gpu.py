"""GPU: the coalescer behind fcv_stream_process / fcv_stream_submit / fcv_stream_await.

folve convolves every open file on its own host thread, one block per call
(SoundProcessor::Process, /root/reference/sound-processor.cc:98-127).  Calls that arrive together
are sent through the GPU as ONE launch sequence; every stream must still get exactly what a call
of its own would have produced: the comparisons here are bit for bit against the same stream run
alone, and against the CPU oracle within the north-star tolerance.
"""
import threading

import numpy as np
import pytest

from folve_b200 import capi
from oracle_py import FilterSpec, OracleConvproc, run_blocks

pytestmark = pytest.mark.gpu


def _spec(seed, size=30000, nch=2):
    r = np.random.default_rng(seed)
    spec = FilterSpec(nch, nch, size)
    for ch in range(nch):
        spec.add(ch, ch, r.standard_normal(size - 600) * 0.004, 500).add(ch, ch, [0.4], 0)
    return spec


def _engine(spec):
    return spec.load(capi.Filter(spec.ninp, spec.nout, spec.size, spec.fragm)).commit(0)


def _inputs(n, frames, nch, seed):
    r = np.random.default_rng(seed)
    return [r.uniform(-0.03, 0.03, (frames + 37 * k, nch)).astype(np.float32) for k in range(n)]


@pytest.mark.parametrize("size", [30000, 3000])
def test_concurrent_threads_equal_the_lone_stream(size):
    """16 host threads, one stream each, blocks racing into shared launch groups."""
    spec = _spec(1, size)
    f = _engine(spec)
    N = spec.fragm
    xs = _inputs(16, 6 * N + 11, 2, 5)
    alone = []
    for x in xs:
        s = capi.Stream(f)
        alone.append(run_blocks(s, x, N))
        s.close()
    streams = [capi.Stream(f) for _ in xs]
    got = [None] * len(xs)
    start = threading.Barrier(len(xs))

    def work(i):
        start.wait()
        got[i] = run_blocks(streams[i], xs[i], N)

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(xs))]
    [t.start() for t in th]
    [t.join() for t in th]
    for i in range(len(xs)):
        assert np.array_equal(got[i], alone[i]), i
    # and the lone stream is the oracle's stream
    yo = run_blocks(spec.load(OracleConvproc(2, 2, spec.size, reset_is_fresh=True)), xs[3], N)
    assert np.abs(alone[3] - yo).max() < 1e-5
    # maxima travel with their own stream
    for i, s in enumerate(streams):
        assert s.max_value == pytest.approx(max(0.0, float(alone[i].max())), abs=0)
        s.close()
    f.close()


def test_submit_await_many_files_from_one_thread():
    """one thread keeps 40 files going: submit all, then await all -- groups of up to 32 streams"""
    spec = _spec(2)
    f = _engine(spec)
    N = spec.fragm
    xs = _inputs(40, 3 * N, 2, 9)
    alone = []
    for x in xs[:5]:
        s = capi.Stream(f)
        alone.append(run_blocks(s, x, N))
        s.close()
    streams = [capi.Stream(f) for _ in xs]
    outs = [[] for _ in xs]
    nblocks = max((len(x) + N - 1) // N for x in xs)
    for k in range(nblocks):
        live = [i for i, x in enumerate(xs) if k * N < len(x)]
        for i in live:
            streams[i].submit(xs[i][k * N:(k + 1) * N])
        for i in reversed(live):   # awaiting in another order than submitting is allowed
            outs[i].append(streams[i].wait())
    for i in range(5):
        assert np.array_equal(np.concatenate(outs[i]), alone[i]), i
    # a second submit without an await in between is refused, the stream stays usable
    streams[0].submit(xs[0][:N])
    with pytest.raises(capi.FcvError):
        streams[0].submit(xs[0][:N])
    streams[0].wait()
    with pytest.raises(capi.FcvError):
        streams[0].wait()
    [s.close() for s in streams]
    f.close()


@pytest.mark.parametrize("fmt,scale_in,scale_out", [(capi.PCM_S16, 32768.0, 32767.0), (capi.PCM_S24, 8388608.0, 8388607.0)])
def test_integer_wire_formats_on_the_single_stream_path(fmt, scale_in, scale_out):
    """int16 / int24 blocks in the stream's buffer: the device applies libsndfile's conversions
    (x / 2^(b-1) in, lrintf(y * (2^(b-1) - 1)) out, both in float32) -- the very samples the float
    path gives when the host converts around it"""
    spec = _spec(3)
    f = _engine(spec)
    N = spec.fragm
    r = np.random.default_rng(4)
    xi = np.rint(r.uniform(-0.03, 0.03, (4 * N + 100, 2)) * scale_in).astype(np.int32)
    sf = capi.Stream(f)
    yf = run_blocks(sf, (xi / scale_in).astype(np.float32), N)
    want = np.rint(yf * np.float32(scale_out)).astype(np.int64)   # float32 product, round half even
    si = capi.Stream(f, fmt, fmt)
    outs = []
    for k in range(0, len(xi), N):
        si.submit(xi[k:k + N])
        outs.append(si.wait())
    yi = np.concatenate(outs).astype(np.int64)
    assert np.array_equal(yi, want)
    assert si.max_value == pytest.approx(sf.max_value, abs=0)
    sf.close()
    si.close()
    f.close()


def test_short_block_leaves_the_rest_of_the_buffer_zero():
    """sound-processor.cc:99-103,116-125: behind the frames that were read the block is zeroed,
    and only that many output frames are written back"""
    spec = _spec(5)
    f = _engine(spec)
    N = spec.fragm
    s = capi.Stream(f)
    r = np.random.default_rng(6)
    s.process(r.uniform(-0.5, 0.5, (N, 2)).astype(np.float32))
    s.buffer[:] = 7.0   # stale content
    y = s.process(r.uniform(-0.5, 0.5, (100, 2)).astype(np.float32))
    assert y.shape == (100, 2) and np.abs(y).max() > 0
    assert not np.any(s.buffer[200:2 * N])
    s.close()
    f.close()


@pytest.mark.parametrize("fmt", [capi.PCM_F32, capi.PCM_S16, capi.PCM_S24])
def test_one_launch_group_equals_three_launches(fmt):
    """a group of single-stream blocks as ONE cooperative launch (fcv_k_fused13.cu) gives bit for bit
    what the forward / MAC / inverse launches give: lone streams, short last blocks, silence,
    many streams per group, every wire format -- and so does the inverse transform as a cluster pair that
    writes the frames straight into the callers' pinned blocks (fcv_debug_set_inv_pair(2))"""
    spec = _spec(7)
    f = _engine(spec)
    N = spec.fragm
    scale = {capi.PCM_F32: 1.0, capi.PCM_S16: 32768.0, capi.PCM_S24: 8388608.0}[fmt]
    r = np.random.default_rng(8)
    # more streams than launch slots: the later ones queue up and travel in groups of several
    lens = [3 * N + 17, 2 * N, N // 3, 4 * N + 1, 2 * N + N // 2, N] + [2 * N + 100 * k for k in range(18)]
    xs = []
    for n in lens:
        x = r.uniform(-0.03, 0.03, (n, 2))
        x[N // 2:N // 2 + 300] = 0.0
        xs.append(x.astype(np.float32) if fmt == capi.PCM_F32 else np.rint(x * scale).astype(np.int32))
    L = capi.lib()

    def run(fused):
        L.fcv_debug_set_fused(1 if fused else 0)
        try:
            streams = [capi.Stream(f, fmt, fmt) for _ in xs]
            outs = [[] for _ in xs]
            for k in range(max((len(x) + N - 1) // N for x in xs) + 1):
                live = [i for i, x in enumerate(xs) if k * N < len(x)]
                if k == 2:   # a block of pure silence for stream 1 in the middle of the run
                    streams[1].submit(np.zeros((0, 2), xs[1].dtype))
                    assert streams[1].wait().shape == (0, 2)
                for i in live:
                    streams[i].submit(xs[i][k * N:(k + 1) * N])
                for i in live:
                    outs[i].append(streams[i].wait())
            mx = [s.max_value for s in streams]
            [s.close() for s in streams]
            return [np.concatenate(o) for o in outs], mx
        finally:
            L.fcv_debug_set_fused(0)

    def run_pair():
        # the inverse transform as a cluster pair writing the frames straight into the callers' blocks
        L.fcv_debug_set_inv_pair(2)
        try:
            return run(False)
        finally:
            L.fcv_debug_set_inv_pair(0)

    yp, mp = run_pair()
    c0 = L.fcv_debug_fused_launches()
    y3, m3 = run(False)
    c1 = L.fcv_debug_fused_launches()
    y1, m1 = run(True)
    c2 = L.fcv_debug_fused_launches()
    assert c1 == c0 and c2 - c1 >= 5           # the cooperative kernel really ran in the second pass only
    for a, b, c in zip(y3, y1, yp):
        assert a.shape == b.shape and np.array_equal(a, b)
        assert a.shape == c.shape and np.array_equal(a, c)
    assert m3 == m1 and m3 == mp
    f.close()
