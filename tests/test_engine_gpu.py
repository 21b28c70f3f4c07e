"""GPU parity: the CUDA engine (through the C ABI) against the CPU oracle and the
float64 ground truth on the same seeded inputs.

Tolerances (BASELINE.json north_star): float32 output within 1e-5 of full scale
of the oracle; <= 1 LSB after 16-bit quantisation.
"""
import numpy as np
import pytest

from folve_b200 import capi
from oracle_py import FilterSpec, OracleConvproc, run_blocks, truth_f64

pytestmark = pytest.mark.gpu

TOL_FS = 1e-5


def _rng(seed):
    return np.random.default_rng(seed)


def _engine(spec):
    return spec.load(capi.Filter(spec.ninp, spec.nout, spec.size, spec.fragm)).commit(0)


def _oracle(spec):
    return spec.load(OracleConvproc(spec.ninp, spec.nout, spec.size, reset_is_fresh=True))


def _three_way(spec, x):
    f = _engine(spec)
    s = capi.Stream(f)
    y = run_blocks(s, x, spec.fragm)
    yo = run_blocks(_oracle(spec), x, spec.fragm)
    t = truth_f64(x, spec.impulses(), spec.nout)
    fs = max(1.0, np.abs(t).max())
    e_eo = np.abs(y - yo).max() / fs
    e_et = np.abs(y - t).max() / fs
    e_ot = np.abs(yo - t).max() / fs
    assert y.shape == yo.shape == t.shape
    assert e_eo < TOL_FS, (e_eo, e_et, e_ot)
    assert e_et < TOL_FS, (e_eo, e_et, e_ot)
    s.close()
    f.close()
    return y, yo, t


@pytest.mark.parametrize("size", [20, 100, 200, 400, 900, 1800, 3000, 4096, 5000])
def test_all_partition_sizes_mono(size):
    """fragm = 64 ... 8192 (every FFT plan)."""
    r = _rng(size)
    spec = FilterSpec(1, 1, size).add(0, 0, r.standard_normal(size) / np.sqrt(size))
    x = r.uniform(-0.5, 0.5, (3 * spec.fragm + 5, 1)).astype(np.float32)
    _three_way(spec, x)


def test_filter_spectra_match_oracle():
    r = _rng(11)
    spec = FilterSpec(2, 2, 30000)
    spec.add(0, 0, r.standard_normal(20000) * 0.01, 500).add(1, 1, r.standard_normal(9000) * 0.01, 0)
    spec.add(0, 0, [0.4], 0)
    f, o = _engine(spec), _oracle(spec)
    assert f.partitions == o.npar == 4
    for (i, out) in ((0, 0), (1, 1), (0, 1)):
        for j in range(4):
            a, b = f.spectrum(i, out, j), o.fftb(i, out, j)
            if b is None or not np.any(b):
                assert a is None
                continue
            assert a is not None
            assert np.abs(a - b).max() <= 2e-6 * max(np.abs(b).max(), 1e-3), (i, out, j)
    f.close()


def test_input_spectrum_matches_oracle():
    r = _rng(12)
    spec = FilterSpec(2, 2, 20000).add(0, 0, r.standard_normal(100)).add(1, 1, r.standard_normal(100))
    f, o = _engine(spec), _oracle(spec)
    s = capi.Stream(f)
    x = r.uniform(-1, 1, (8192, 2)).astype(np.float32)
    s.process(x)
    o.process(x)
    for ch in range(2):
        a, b = s.input_spectrum(ch, 0), o.ffta(ch, 0)
        assert np.abs(a - b).max() <= 3e-6 * np.abs(b).max()
    s.close()
    f.close()


def test_stereo_long_reverb_shape():
    """SantaLucia-shaped: 178193-tap IR at delay 500 plus a dirac, stereo diagonal."""
    r = _rng(13)
    spec = FilterSpec(2, 2, 204800)
    env = np.exp(-np.arange(178193) / 40000.0)
    for ch in range(2):
        spec.add(ch, ch, r.standard_normal(178193) * env * 0.004, 500)
        spec.add(ch, ch, [0.4], 0)
    x = r.uniform(-0.03, 0.03, (30 * 8192 + 1234, 2)).astype(np.float32)
    f = _engine(spec)
    assert f.partitions == 25 and f.ring_depth == 22 and f.active_rows == 44
    f.close()
    y, yo, t = _three_way(spec, x)
    # <= 1 LSB after 16-bit quantisation (lrintf(y * 32767), libsndfile convention)
    q, qo = np.rint(y * 32767.0), np.rint(yo * 32767.0)
    assert np.abs(q - qo).max() <= 1


def test_mimo_crossfeed_with_link():
    r = _rng(14)
    spec = FilterSpec(2, 2, 10000)
    spec.add(0, 0, r.standard_normal(4096) * 0.02).add(1, 1, r.standard_normal(4096) * 0.02)
    spec.link(0, 0, 1, 0)
    spec.add(0, 0, [0.25], 9000)
    spec.add(0, 1, [0.3], 700)
    x = r.uniform(-0.2, 0.2, (4 * 8192 + 5, 2)).astype(np.float32)
    _three_way(spec, x)


@pytest.mark.parametrize("nin,nout", [(1, 2), (3, 1), (6, 6), (5, 9)])
def test_mimo_dense(nin, nout):
    r = _rng(100 * nin + nout)
    spec = FilterSpec(nin, nout, 9000)
    for i in range(nin):
        for o in range(nout):
            if (i + o) % 3 != 2:  # leave some pairs empty
                spec.add(i, o, r.standard_normal(2000 + 500 * o) * 0.01, 100 * i)
    x = r.uniform(-0.3, 0.3, (3 * 8192 + 11, nin)).astype(np.float32)
    _three_way(spec, x)


def test_block_edges_clicks_and_short_files():
    r = _rng(15)
    spec = FilterSpec(1, 1, 20000).add(0, 0, r.standard_normal(20000) * 0.01)
    N = spec.fragm
    for frames in (1, N - 1, N, N + 1, 3 * N + 7):
        x = np.zeros((frames, 1), np.float32)
        for pos in (0, N - 1, N, 2 * N - 1):
            if pos < frames:
                x[pos, 0] = 1.0
        _three_way(spec, x)


def test_impulse_returns_filter_exact_positions():
    r = _rng(16)
    h = (r.standard_normal(9000) * 0.01).astype(np.float32)
    spec = FilterSpec(1, 1, 9000).add(0, 0, h)
    x = np.zeros((3 * 8192, 1), np.float32)
    x[0, 0] = 1.0
    y, _, _ = _three_way(spec, x)
    assert np.abs(y[:9000, 0] - h).max() < 2e-6
    assert np.abs(y[9000:, 0]).max() < 2e-6


def test_reset_equals_fresh_and_signed_max():
    r = _rng(17)
    spec = FilterSpec(1, 1, 300).add(0, 0, [-1.0])   # inverter: negative peaks become positive
    f = _engine(spec)
    s = capi.Stream(f)
    x = r.uniform(-0.5, 0.25, (3 * spec.fragm, 1)).astype(np.float32)
    y1 = run_blocks(s, x, spec.fragm)
    # sound-processor.cc:120-123: signed comparison, no fabs
    assert s.max_value == pytest.approx(max(0.0, float(y1.max())), abs=1e-6)
    s.process(x[:spec.fragm])          # odd number of extra blocks, then reset
    s.reset()
    assert s.max_value == 0.0
    y2 = run_blocks(s, x, spec.fragm)
    assert np.array_equal(y1, y2)
    s.close()
    f.close()


def test_batch_matches_single_streams_and_formats():
    r = _rng(18)
    spec = FilterSpec(2, 2, 20000)
    for ch in range(2):
        spec.add(ch, ch, r.standard_normal(20000) * 0.005, 0)
    f = _engine(spec)
    N, B, nblk = spec.fragm, 7, 4
    x = r.uniform(-0.4, 0.4, (B, nblk * N, 2)).astype(np.float32)
    ref = [run_blocks(capi.Stream(f), x[b], N) for b in range(B)]
    bt = capi.Batch(f, B)
    outs = []
    for k in range(nblk):
        bt.host_in[:] = x[:, k * N:(k + 1) * N]
        bt.process()
        outs.append(bt.host_out.copy())
    y = np.concatenate(outs, axis=1)
    for b in range(B):
        assert np.array_equal(y[b], ref[b])
    mx = bt.get_max()
    assert np.allclose(mx, np.maximum(0.0, y.max(axis=(1, 2))), atol=1e-7)
    bt.close()

    # int16 wire format, fused conversion: x/32768 in, lrintf(y*32767) out
    xi = np.rint(x * 20000).astype(np.int16)
    bs = capi.Batch(f, B, capi.PCM_S16, capi.PCM_S16)
    oi = []
    for k in range(nblk):
        bs.host_in[:] = xi[:, k * N:(k + 1) * N]
        bs.process()
        oi.append(bs.host_out.copy())
    yi = np.concatenate(oi, axis=1)
    o = _oracle(spec)
    yo = run_blocks(o, xi[0].astype(np.float32) / 32768.0, N)
    assert np.abs(yi[0].astype(np.int64) - np.rint(yo * 32767.0).astype(np.int64)).max() <= 1
    bs.close()
    f.close()


def test_batch_frames_valid_and_slot_reset():
    r = _rng(19)
    spec = FilterSpec(1, 1, 5000).add(0, 0, r.standard_normal(5000) * 0.01)
    f = _engine(spec)
    N, B = spec.fragm, 5
    bt = capi.Batch(f, B)
    x = r.uniform(-0.4, 0.4, (B, N, 1)).astype(np.float32)
    fv = np.array([N, 1, N // 2, 0, N - 1], np.int32)
    bt.host_in[:] = x
    bt.process(fv)
    o = _oracle(spec)
    for b in range(B):
        o.reset()
        if fv[b]:
            yo = o.process(x[b, :fv[b]])
            assert np.abs(bt.host_out[b, :fv[b]] - yo).max() < TOL_FS
    # reset slot 0 only: slot 0 behaves fresh, slot 4 keeps its history
    bt.reset_slot(0)
    bt.host_in[:] = 0
    bt.process()
    assert np.all(bt.host_out[0] == 0.0)
    assert np.abs(bt.host_out[4]).max() > 0.0
    bt.close()
    f.close()


@pytest.mark.parametrize("T", [2, 4, 8])
def test_time_tiled_batch_equals_block_by_block(T):
    """T blocks per step (time-tiled MAC, X rows read once for all outputs that need
    them) must reproduce the one-block-per-step results, including short steps,
    frames_valid inside a block, int16 wire format and a MIMO filter with a link."""
    r = _rng(30 + T)
    spec = FilterSpec(2, 2, 60000)
    spec.add(0, 0, r.standard_normal(50000) * 0.003, 500).add(1, 1, r.standard_normal(60000) * 0.003, 0)
    spec.add(0, 1, r.standard_normal(9000) * 0.003, 20000)   # partitions 2 and 3 only
    spec.link(0, 0, 1, 0)
    f = _engine(spec)
    N, B, nsteps = spec.fragm, 5, 3
    total = nsteps * T * N
    x = r.uniform(-0.3, 0.3, (B, total, 2)).astype(np.float32)
    one = capi.Batch(f, B)
    want = np.zeros((B, total, 2), np.float32)
    for k in range(nsteps * T):
        one.host_in[:] = x[:, k * N:(k + 1) * N]
        one.process()
        want[:, k * N:(k + 1) * N] = one.host_out
    wmax = one.get_max()
    one.close()
    tt = capi.Batch(f, B, blocks_per_step=T)
    assert tt.host_in.shape == (B, T * N, 2)
    got = np.zeros_like(want)
    for k in range(nsteps):
        tt.host_in[:] = x[:, k * T * N:(k + 1) * T * N]
        tt.process()
        got[:, k * T * N:(k + 1) * T * N] = tt.host_out
    assert np.abs(got - want).max() < 2e-6
    assert np.allclose(tt.get_max(), wmax, atol=2e-6)
    # a short last step: stream b has (b+1) * 1000 valid frames, then everything is reset
    fv = np.array([(b + 1) * 1000 for b in range(B)], np.int32)
    fv[-1] = T * N
    tt.host_in[:] = x[:, :T * N]
    tt.process(fv)
    o = _oracle(spec)
    for b in range(B):
        o.reset()
        xs = x[b, :T * N]
        # oracle: continue the history of `want`'s stream: rebuild it block by block
        hist = run_blocks(o, x[b], N)
        tail = run_blocks(o, xs[:fv[b]], N)
        assert np.abs(tt.host_out[b, :fv[b]] - tail).max() < TOL_FS, b
    tt.close()
    # int16 wire
    xi = np.rint(x * 20000).astype(np.int16)
    a = capi.Batch(f, B, capi.PCM_S16, capi.PCM_S16)
    c = capi.Batch(f, B, capi.PCM_S16, capi.PCM_S16, blocks_per_step=T)
    c.host_in[:] = xi[:, :T * N]
    c.process()
    for k in range(T):
        a.host_in[:] = xi[:, k * N:(k + 1) * N]
        a.process()
        assert np.abs(a.host_out.astype(np.int32) - c.host_out[:, k * N:(k + 1) * N].astype(np.int32)).max() <= 1
    a.close(); c.close()
    f.close()


@pytest.mark.parametrize("fmt", ["f32", "s16", "s24"])
def test_stereo_pair_inverse_kernel_is_bit_identical(fmt):
    """The inverse transform of a stereo batch with the two channels' CTAs as a cluster that writes whole
    interleaved frames (fcv_debug_set_inv_pair; an experiment that stays off) gives the same bits as the
    default kernel: block contents, running maximum, per-block maxima; T = 1 and T = 8, short steps."""
    r = _rng(77)
    spec = FilterSpec(2, 2, 40000)
    spec.add(0, 0, r.standard_normal(40000) * 0.004, 100).add(1, 1, r.standard_normal(30000) * 0.004, 0)
    spec.add(1, 0, r.standard_normal(5000) * 0.004, 9000)
    f = _engine(spec)
    N, B = spec.fragm, 7
    pcm = {"f32": capi.PCM_F32, "s16": capi.PCM_S16, "s24": capi.PCM_S24}[fmt]
    L = capi.lib()
    for T in (1, 8):
        x = r.uniform(-0.3, 0.3, (3, B, T * N, 2))
        xin = (x.astype(np.float32) if fmt == "f32" else
               np.rint(x * (20000 if fmt == "s16" else 5000000)).astype(np.int16 if fmt == "s16" else np.int32))
        fv = np.array([T * N, 1, N - 1, N, T * N - 3, 0, 5000][:B], np.int32)
        outs = []
        for pair in (0, 1):
            L.fcv_debug_set_inv_pair(pair)
            try:
                bt = capi.Batch(f, B, pcm, pcm, blocks_per_step=T)
                got = []
                for k in range(3):
                    bt.host_in[:] = xin[k]
                    bt.process(fv if k == 2 else None)
                    got.append(bt.host_out.copy())
                got.append(bt.get_max().copy())
                got.append(bt.get_block_max().copy())
                bt.close()
                outs.append(got)
            finally:
                L.fcv_debug_set_inv_pair(0)
        for a, c in zip(*outs):
            assert np.array_equal(a, c)
    f.close()


def test_persistent_mac_grid_equals_small_batches():
    """Stereo filters at 8 blocks per step run the time-tiled MAC as a persistent grid (3 CTAs per SM walking the
    work items) once a batch has more items than CTA slots (here: 24 streams = 768 items); a batch of two streams
    (64 items) gets one CTA per item.  Streams are independent, so the big batch must give the bits of the small
    ones: two steps, int16 wire, a filter whose window does not end on a chunk boundary."""
    r = _rng(79)
    spec = FilterSpec(2, 2, 9 * 8192 - 100)
    spec.add(0, 0, r.standard_normal(9 * 8192 - 100) * 0.003, 0).add(1, 1, r.standard_normal(50000) * 0.003, 300)
    spec.add(1, 0, r.standard_normal(4000) * 0.003, 8192)
    f = _engine(spec)
    N, B, T = spec.fragm, 24, 8
    x = np.rint(r.uniform(-0.3, 0.3, (2, B, T * N, 2)) * 20000).astype(np.int16)
    big = capi.Batch(f, B, capi.PCM_S16, capi.PCM_S16, blocks_per_step=T)
    want = []
    for k in range(2):
        big.host_in[:] = x[k]
        big.process()
        want.append(big.host_out.copy())
    wmax = big.get_max().copy()
    big.close()
    for s0 in range(0, B, 2):
        small = capi.Batch(f, 2, capi.PCM_S16, capi.PCM_S16, blocks_per_step=T)
        for k in range(2):
            small.host_in[:] = x[k][s0:s0 + 2]
            small.process()
            assert np.array_equal(small.host_out, want[k][s0:s0 + 2])
        assert np.array_equal(small.get_max(), wmax[s0:s0 + 2])
        small.close()
    f.close()


@pytest.mark.parametrize("fmt", ["f32", "s16", "s24"])
def test_tensor_memory_transform_kernels_are_bit_identical(fmt):
    """Batches with several blocks per step: the inverse transform that keeps each thread's twiddles and the
    overlap tail in tensor memory (default) and the forward transform that does so with its twiddles (an
    experiment that stays off) give the same bits as the kernels that re-read them from global memory
    (fcv_debug_set_tmem): block contents, running maximum, per-block maxima; stereo and 3 -> 2 channels,
    T = 2 and 8, short steps, silence, several steps (the tail is carried from step to step)."""
    r = _rng(78)
    L = capi.lib()
    pcm = {"f32": capi.PCM_F32, "s16": capi.PCM_S16, "s24": capi.PCM_S24}[fmt]
    before = L.fcv_debug_get_tmem()
    for nin, nout in ((2, 2), (3, 2)):
        spec = FilterSpec(nin, nout, 40000)
        spec.add(0, 0, r.standard_normal(40000) * 0.004, 100).add(1, 1, r.standard_normal(30000) * 0.004, 0)
        spec.add(nin - 1, 0, r.standard_normal(5000) * 0.004, 9000)
        f = _engine(spec)
        N, B = spec.fragm, 7
        for T in (2, 8):
            x = r.uniform(-0.3, 0.3, (3, B, T * N, nin))
            xin = (x.astype(np.float32) if fmt == "f32" else
                   np.rint(x * (20000 if fmt == "s16" else 5000000)).astype(np.int16 if fmt == "s16" else np.int32))
            fv = np.array([T * N, 1, N - 1, N, T * N - 3, 0, 5000][:B], np.int32)
            outs = []
            for mask in (0, 1, 3):
                L.fcv_debug_set_tmem(mask)
                try:
                    bt = capi.Batch(f, B, pcm, pcm, blocks_per_step=T)
                    got = []
                    for k in range(3):
                        bt.host_in[:] = xin[k]
                        bt.process(fv if k == 1 else None)
                        got.append(bt.host_out.copy())
                    got.append(bt.get_max().copy())
                    got.append(bt.get_block_max().copy())
                    bt.close()
                    outs.append(got)
                finally:
                    L.fcv_debug_set_tmem(before)
            for other in outs[1:]:
                for a, c in zip(outs[0], other):
                    assert np.array_equal(a, c)
        f.close()
    assert L.fcv_debug_get_tmem() == before


@pytest.mark.parametrize("nin,nout,T,size", [(1, 1, 4, 30000), (3, 2, 4, 30000), (1, 2, 8, 30000), (6, 6, 2, 30000),
                                             (3, 2, 8, 30000), (6, 6, 8, 30000), (2, 2, 8, 110000), (2, 3, 4, 110000)])
def test_time_tiled_other_channel_counts_and_s24(nin, nout, T, size):
    """fragm = 8192 with mono, odd and 5.1 channel counts: the forward kernel's mono and
    scalar-load paths, several blocks per step, 24-bit wire format; tiled == block by block
    and both == oracle.  The T = 8 cases with several inputs per output walk more than one
    (input, output) pair per work item of the TMA-staged MAC (rows of a pair restart on
    pipeline stage 0), the long filters more than one ring revolution per pair; 3 streams
    leave the last stream group half empty."""
    r = _rng(90 + 10 * nin + nout + T + size // 10000)
    spec = FilterSpec(nin, nout, size)
    for i in range(nin):
        for o in range(nout):
            if (i + o) % 2 == 0 or nin == 1 or size > 30000:
                taps = 8000 + 4000 * ((i + o) % 3) if size == 30000 else size - 20000 - 7000 * ((i + o) % 3)
                spec.add(i, o, r.standard_normal(taps) * (0.004 if size == 30000 else 0.001), 3000 * ((i * nout + o) % 5))
    f = _engine(spec)
    N, B = spec.fragm, 3
    assert N == 8192
    total = 2 * T * N
    xi = r.integers(-(1 << 21), 1 << 21, (B, total, nin)).astype(np.int32)       # 24-bit samples
    one = capi.Batch(f, B, capi.PCM_S24, capi.PCM_F32)
    want = np.zeros((B, total, nout), np.float32)
    for k in range(2 * T):
        one.host_in[:] = xi[:, k * N:(k + 1) * N]
        one.process()
        want[:, k * N:(k + 1) * N] = one.host_out
    one.close()
    tt = capi.Batch(f, B, capi.PCM_S24, capi.PCM_F32, blocks_per_step=T)
    got = np.zeros_like(want)
    for k in range(2):
        tt.host_in[:] = xi[:, k * T * N:(k + 1) * T * N]
        tt.process()
        got[:, k * T * N:(k + 1) * T * N] = tt.host_out
    tt.close()
    assert np.abs(got - want).max() < 2e-6
    o = _oracle(spec)
    x0 = (xi[0].astype(np.float64) / 8388608.0).astype(np.float32)
    ref = run_blocks(o, x0, N)
    assert np.abs(got[0] - ref).max() < TOL_FS
    f.close()


def test_async_submit_wait_two_slots():
    r = _rng(20)
    spec = FilterSpec(2, 2, 20000)
    for ch in range(2):
        spec.add(ch, ch, r.standard_normal(20000) * 0.005, 0)
    f = _engine(spec)
    N, B, nblk = spec.fragm, 300, 5          # > 256 streams: several chunks over several CUDA streams
    x = r.uniform(-0.4, 0.4, (B, nblk * N, 2)).astype(np.float32)
    ref = capi.Batch(f, B)
    want = []
    for k in range(nblk):
        ref.host_in[:] = x[:, k * N:(k + 1) * N]
        ref.process()
        want.append(ref.host_out.copy())
    ref.close()
    bt = capi.Batch(f, B)
    views = [bt.slot_views(0), bt.slot_views(1)]
    got = []
    views[0][0][:] = x[:, :N]
    bt.submit(0)
    for k in range(1, nblk):
        views[k & 1][0][:] = x[:, k * N:(k + 1) * N]
        bt.submit(k & 1)
        bt.wait((k - 1) & 1)
        got.append(views[(k - 1) & 1][1].copy())
    bt.wait((nblk - 1) & 1)
    got.append(views[(nblk - 1) & 1][1].copy())
    for k in range(nblk):
        assert np.array_equal(got[k], want[k]), k
    bt.close()
    f.close()


def test_errors_are_reported_not_fatal():
    L = capi.lib()
    assert not L.fcv_filter_begin(0, 2, 100, 64)
    assert b"out of range" in L.fcv_last_error()
    assert not L.fcv_filter_begin(2, 2, 100, 100)       # fragm not a power of two
    f = capi.Filter(1, 1, 100, 64)
    with pytest.raises(capi.FcvError):
        f.add(1, 0, [1.0], 0)
    with pytest.raises(capi.FcvError):
        capi.Stream(f)                                   # not committed
    f.commit(0)
    with pytest.raises(capi.FcvError):
        f.add(0, 0, [1.0], 0)                            # immutable after commit
    f.close()


@pytest.mark.parametrize("signal", ["sine_on_bin", "sine_off_bin", "square_full_scale", "dc_step"])
def test_deterministic_signals_against_truth(signal):
    """SURVEY 8(c) golden set (ii): tones that sit exactly on an FFT bin and between two bins, a
    full-scale square wave (every sample at +-1: the worst case for accumulation error) and a DC
    step, through a 20000-tap stereo filter with a cross path -- engine vs oracle vs float64 truth."""
    r = _rng(400)
    spec = FilterSpec(2, 2, 20000)
    env = np.exp(-np.arange(20000) / 3000.0)
    for (i, o, g) in ((0, 0, 1.0), (1, 1, 1.0), (0, 1, 0.3)):
        h = r.standard_normal(20000) * env
        spec.add(i, o, (g * 0.9 / np.abs(h).sum()) * h, 0)          # sum |h| < 1.2 per output: no overflow at full scale
    N = spec.fragm
    n = np.arange(6 * N + 321)
    if signal == "sine_on_bin":
        k = 37                                                      # exactly bin 37 of the 2N-point transform
        x = np.stack([np.sin(2 * np.pi * k * n / (2 * N)), np.cos(2 * np.pi * (k + 5) * n / (2 * N))], 1)
    elif signal == "sine_off_bin":
        x = np.stack([np.sin(2 * np.pi * 37.5 * n / (2 * N)), np.sin(2 * np.pi * 0.123456 * n)], 1)
    elif signal == "square_full_scale":
        x = np.stack([np.where((n // 50) % 2 == 0, 1.0, -1.0), np.where((n // 3) % 2 == 0, -1.0, 1.0)], 1)
    else:
        x = np.stack([(n >= N - 1).astype(np.float64), -(n >= 2 * N).astype(np.float64)], 1)
    _three_way(spec, x.astype(np.float32))


def test_maximum_filter_size_128_partitions():
    """MAXSIZE = 0x100000 taps (zita-config.h:61): 128 partitions of 8192; one stream block by block
    and a 4-stream batch stepping 8 blocks at once (the window of 135 ring slots walks the TMA ring 17
    times per pair) against the oracle."""
    r = _rng(500)
    spec = FilterSpec(1, 1, 0x100000)
    taps = 0x100000 - 1000
    spec.add(0, 0, r.standard_normal(taps) * np.exp(-np.arange(taps) / 250000.0) * 2e-4, 1000)
    N = spec.fragm
    f = _engine(spec)
    assert f.partitions == 128 and f.ring_depth == 128
    x = r.uniform(-0.5, 0.5, (16 * N, 1)).astype(np.float32)
    yo = run_blocks(_oracle(spec), x, N)
    s = capi.Stream(f)
    y = run_blocks(s, x, N)
    s.close()
    assert np.abs(y - yo).max() < TOL_FS
    B, T = 4, 8
    bt = capi.Batch(f, B, capi.PCM_F32, capi.PCM_F32, blocks_per_step=T)
    got = np.zeros((16 * N, 1), np.float32)
    for k in range(2):
        bt.host_in[:] = 0
        bt.host_in[2] = x[k * T * N:(k + 1) * T * N]
        bt.process()
        got[k * T * N:(k + 1) * T * N] = bt.host_out[2]
    bt.close()
    f.close()
    assert np.abs(got - yo).max() < TOL_FS


def test_maximum_channel_counts_64x64():
    """Convproc::MAXINP x MAXOUT = 64 x 64 (zita-fconfig.cc:49,55), every fourth pair populated."""
    r = _rng(501)
    spec = FilterSpec(64, 64, 9000)
    for i in range(64):
        for o in range(64):
            if (i * 7 + o) % 4 == 0:
                spec.add(i, o, r.standard_normal(3000 + 50 * ((i + o) % 7)) * 0.002, 100 * ((i + 3 * o) % 9))
    N = spec.fragm
    x = r.uniform(-0.2, 0.2, (3 * N + 11, 64)).astype(np.float32)
    f = _engine(spec)
    s = capi.Stream(f)
    y = run_blocks(s, x, N)
    s.close()
    f.close()
    yo = run_blocks(_oracle(spec), x, N)
    assert y.shape == yo.shape == (3 * N + 11, 64)
    assert np.abs(y - yo).max() < TOL_FS
