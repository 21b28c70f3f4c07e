"""CPU: the zita-convolver restatement (oracle/) against the float64 ground truth.

The reference ships no golden vectors, so the oracle is pinned
against the mathematical definition of the path: y[o] = sum_i h[i,o] * x[i],
output length == input length, zero latency.
"""
import numpy as np
import pytest

from oracle_py import FilterSpec, OracleConvproc, fragm_for, run_blocks, truth_f64


def _rng(seed):
    return np.random.default_rng(seed)


def _check(spec, x, tol=2e-6):
    o = spec.load(OracleConvproc(spec.ninp, spec.nout, spec.size))
    y = run_blocks(o, x, spec.fragm)
    t = truth_f64(x, spec.impulses(), spec.nout)
    scale = max(1.0, np.abs(t).max())
    err = np.abs(y - t).max() / scale
    assert y.shape == t.shape
    assert err < tol, err
    return y, t


def test_fragm_rule():
    # zita-fconfig.cc:74-77
    assert fragm_for(65536) == 8192
    assert fragm_for(204800) == 8192
    assert fragm_for(4097) == 8192
    assert fragm_for(4096) == 4096
    assert fragm_for(2048) == 2048
    assert fragm_for(100) == 128
    assert fragm_for(64) == 64
    assert fragm_for(1) == 64


@pytest.mark.parametrize("size,taps", [(64, 40), (300, 300), (1000, 777), (5000, 5000), (20000, 20000)])
def test_mono_random_filter(size, taps):
    r = _rng(size)
    spec = FilterSpec(1, 1, size).add(0, 0, r.standard_normal(taps) / np.sqrt(taps))
    x = r.uniform(-0.5, 0.5, (3 * spec.fragm + 17, 1)).astype(np.float32)
    _check(spec, x)


def test_impulse_returns_filter():
    r = _rng(1)
    h = (r.standard_normal(9000) * 0.01).astype(np.float32)
    spec = FilterSpec(1, 1, 9000).add(0, 0, h)
    x = np.zeros((3 * 8192, 1), np.float32)
    x[0, 0] = 1.0
    y, _ = _check(spec, x)
    assert np.abs(y[:9000, 0] - h).max() < 1e-6
    assert np.abs(y[9000:, 0]).max() < 1e-6


def test_stereo_diag_accumulate_and_dirac():
    # the SantaLucia pattern: /impulse/read and /impulse/dirac on the same pair add
    r = _rng(2)
    spec = FilterSpec(2, 2, 30000)
    for ch in range(2):
        spec.add(ch, ch, r.standard_normal(20000) * 0.004, 500)
        spec.add(ch, ch, [0.4], 0)
    x = r.uniform(-0.03, 0.03, (5 * 8192 + 123, 2)).astype(np.float32)
    _check(spec, x)


def test_mimo_with_link():
    r = _rng(3)
    spec = FilterSpec(2, 2, 10000)
    spec.add(0, 0, r.standard_normal(4096) * 0.02)
    spec.add(1, 1, r.standard_normal(4096) * 0.02)
    spec.link(0, 0, 1, 0)           # cross path uses the spectra of (0,0)
    spec.add(0, 0, [0.25], 9000)    # later addition to the source is seen by the link
    spec.add(0, 1, [0.3], 700)
    x = r.uniform(-0.2, 0.2, (4 * 8192 + 5, 2)).astype(np.float32)
    _check(spec, x)


def test_block_edges_and_short_files():
    r = _rng(4)
    spec = FilterSpec(1, 1, 20000).add(0, 0, r.standard_normal(20000) * 0.01)
    N = spec.fragm
    for frames in (1, N - 1, N, N + 1, 3 * N + 7):
        x = np.zeros((frames, 1), np.float32)
        for pos in (0, N - 1, N, 2 * N - 1):
            if pos < frames:
                x[pos, 0] = 1.0
        _check(spec, x)


def test_reset_quirk_documented():
    """with the recalled library behaviour a Reset()
    after an odd number of blocks leaves Convproc::_inpoffs in the second half,
    so the re-used processor lags by one block; with reset_is_fresh it does not."""
    r = _rng(5)
    h = r.standard_normal(100) * 0.1
    x = r.uniform(-0.5, 0.5, (2 * 128, 1)).astype(np.float32)
    t = truth_f64(x, {(0, 0): h}, 1)
    for fresh in (False, True):
        o = OracleConvproc(1, 1, 100, reset_is_fresh=fresh)
        o.add(0, 0, h, 0)
        o.reset()
        o.process(x[:128])          # one (odd) block
        o.reset()                   # ProcessorPool::Return -> Reset
        y = run_blocks(o, x, 128)
        if fresh:
            assert np.abs(y - t).max() < 1e-5
        else:
            assert np.abs(y[:128]).max() == 0.0          # one block of silence
            assert np.abs(y[128:] - t[:128]).max() < 1e-5  # then everything one block late


def test_unused_output_is_silent():
    spec = FilterSpec(2, 2, 100).add(0, 0, [1.0])
    x = _rng(6).uniform(-1, 1, (300, 2)).astype(np.float32)
    y, _ = _check(spec, x)
    assert np.all(y[:, 1] == 0.0)
    assert np.abs(y[:, 0] - x[:, 0]).max() < 1e-6


@pytest.mark.parametrize("signal", ["sine_on_bin", "sine_off_bin", "square_full_scale"])
def test_deterministic_signals(signal):
    """Tones on and between FFT bins and a full-scale square wave (SURVEY 8(c) golden set ii):
    the oracle against the float64 truth; the GPU suite runs the same inputs through the engine."""
    r = _rng(400)
    spec = FilterSpec(2, 2, 20000)
    env = np.exp(-np.arange(20000) / 3000.0)
    for (i, o, g) in ((0, 0, 1.0), (1, 1, 1.0), (0, 1, 0.3)):
        h = r.standard_normal(20000) * env
        spec.add(i, o, (g * 0.9 / np.abs(h).sum()) * h, 0)
    N = spec.fragm
    n = np.arange(4 * N + 321)
    if signal == "sine_on_bin":
        x = np.stack([np.sin(2 * np.pi * 37 * n / (2 * N)), np.cos(2 * np.pi * 42 * n / (2 * N))], 1)
    elif signal == "sine_off_bin":
        x = np.stack([np.sin(2 * np.pi * 37.5 * n / (2 * N)), np.sin(2 * np.pi * 0.123456 * n)], 1)
    else:
        x = np.stack([np.where((n // 50) % 2 == 0, 1.0, -1.0), np.where((n // 3) % 2 == 0, -1.0, 1.0)], 1)
    _check(spec, x.astype(np.float32))


@pytest.mark.parametrize("seed", range(24))
def test_random_filter_structures(seed):
    """Random channel counts, sizes (every block size of the fragm rule), overlapping additions at random
    offsets, links made before and after their source gets data, unused pairs, ragged last block: the oracle
    against direct float64 convolution."""
    r = _rng(1000 + seed)
    ninp, nout = int(r.integers(1, 5)), int(r.integers(1, 5))
    size = int(r.choice([40, 64, 100, 129, 300, 700, 1500, 3000, 6000, 9000]))
    spec = FilterSpec(ninp, nout, size)
    pairs = [(i, o) for i in range(ninp) for o in range(nout)]
    for _ in range(int(r.integers(1, 9))):
        i, o = pairs[int(r.integers(len(pairs)))]
        if r.random() < 0.25:
            i2, o2 = pairs[int(r.integers(len(pairs)))]
            if (i2, o2) != (i, o):
                spec.link(i, o, i2, o2)
            continue
        taps = int(r.integers(1, size + 1))
        i0 = int(r.integers(0, size))
        spec.add(i, o, r.standard_normal(taps) * 0.3 / np.sqrt(taps), i0)
    frames = int(r.integers(1, 4)) * spec.fragm + int(r.integers(0, spec.fragm))
    x = r.uniform(-0.5, 0.5, (frames, ninp)).astype(np.float32)
    _check(spec, x)


@pytest.mark.parametrize("nreal", [8, 16, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 65536])
def test_fft_stand_in_against_float64(nreal):
    """oracle/fft_oracle.c stands in for FFTW's r2c / c2r (unnormalised, forward sign -1): against numpy's float64
    transforms at float32 accuracy, round trip == nreal * x, DC / Nyquist imaginary parts ignored by c2r, inputs left
    untouched.  Sizes: 2 * fragm for every block size of the fragm rule, and the ends of the plan's range."""
    import ctypes as C
    from oracle_py import lib
    L = lib()
    fp = C.POINTER(C.c_float)
    L.offt_plan_create.restype, L.offt_plan_create.argtypes = C.c_void_p, [C.c_int]
    L.offt_plan_destroy.restype, L.offt_plan_destroy.argtypes = None, [C.c_void_p]
    for f in (L.offt_r2c, L.offt_c2r):
        f.restype, f.argtypes = None, [C.c_void_p, fp, fp]
    p = L.offt_plan_create(nreal)
    assert p
    r = _rng(nreal)
    x = r.uniform(-1, 1, nreal).astype(np.float32)
    x0 = x.copy()
    X = np.zeros(nreal + 2, np.float32)
    L.offt_r2c(p, x.ctypes.data_as(fp), X.ctypes.data_as(fp))
    assert np.array_equal(x, x0)
    want = np.fft.rfft(x.astype(np.float64))
    got = X[0::2].astype(np.float64) + 1j * X[1::2].astype(np.float64)
    rms = np.sqrt(np.mean(np.abs(want) ** 2))
    assert np.abs(got - want).max() < 4e-6 * rms * np.log2(nreal)
    assert X[1] == 0 and X[nreal + 1] == 0
    Xd = X.copy()
    Xd[1], Xd[nreal + 1] = 123.0, -7.0                   # must be ignored
    Xd0 = Xd.copy()
    y = np.zeros(nreal, np.float32)
    L.offt_c2r(p, Xd.ctypes.data_as(fp), y.ctypes.data_as(fp))
    assert np.array_equal(Xd, Xd0)
    assert np.abs(y / nreal - x).max() < 2e-6 * np.log2(nreal)
    # a spectrum of its own: c2r against irfft
    Z = (r.standard_normal(nreal // 2 + 1) + 1j * r.standard_normal(nreal // 2 + 1))
    Z[0], Z[-1] = Z[0].real, Z[-1].real
    Zi = np.zeros(nreal + 2, np.float32)
    Zi[0::2], Zi[1::2] = Z.real, Z.imag
    L.offt_c2r(p, Zi.ctypes.data_as(fp), y.ctypes.data_as(fp))
    zt = np.fft.irfft(Zi[0::2].astype(np.float64) + 1j * Zi[1::2].astype(np.float64), nreal) * nreal
    assert np.abs(y - zt).max() < 4e-6 * np.sqrt(np.mean(zt ** 2)) * np.log2(nreal)
    L.offt_plan_destroy(p)
