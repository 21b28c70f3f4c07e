"""CPU: the reference's own SoundProcessor (sound-processor.cc compiled unmodified,
driven through the restated caller protocol) against the float64 ground truth,
including the gapless hand-off and the quirks listed in SURVEY.md section 8(a)."""
import os

import numpy as np
import pytest

import harness_py as H
from configs import make_filter_dirs
from oracle_py import OracleConvproc, run_blocks, truth_f64

pytestmark = pytest.mark.skipif(not H.have_reference(), reason="needs oracle/_ref/libfolve_ref.so")


def _noise(frames, ch, peak, seed):
    r = np.random.default_rng(seed)
    return (np.rint(r.uniform(-peak, peak, (frames, ch)) * 32768) / 32768).astype(np.float32)


@pytest.fixture(autouse=True)
def _fresh_pool():
    """Parity contract: processors are fresh, and Reset()
    restores the fresh state.  The pooled-reuse lag is covered by its own test."""
    R = H.reference()
    R.drop_pool()
    R.set_reset_is_fresh(True)
    yield
    R.drop_pool()


@pytest.fixture(scope="module")
def dirs(tmp_path_factory):
    return make_filter_dirs(tmp_path_factory.mktemp("filters"))


def _impulses(name, dirs):
    d, rate, ch, bits = dirs[name]
    conf = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".conf")][0]
    c = H.reference().load_config(conf, rate, ch)
    scale = 2.0 * c["fragm"]
    return c, {k: v[1].astype(np.float64) * scale for k, v in c["pairs"].items()}


def test_single_file_matches_truth(dirs):
    R = H.reference()
    for name in ("crossfeed", "tiny", "hilbert", "surround51"):
        d, rate, ch, bits = dirs[name]
        c, h = _impulses(name, dirs)
        N = c["fragm"]
        x = _noise(2 * N + N // 3, ch, 0.25, 7)
        (y,), mx, flags = R.run_chain(d, rate, ch, bits, [x])
        t = truth_f64(x, h, c["nout"])
        assert y.shape == t.shape                      # output length == input length (quirk 7)
        assert np.abs(y - t).max() < 2e-6 * max(1.0, np.abs(t).max())
        assert mx[0] == pytest.approx(max(0.0, float(y.max())), abs=1e-7)   # signed maximum (quirk 1)
        assert flags == [0]


def test_gapless_chain_equals_one_long_file(dirs):
    R = H.reference()
    d, rate, ch, bits = dirs["crossfeed"]
    c, h = _impulses("crossfeed", dirs)
    N = c["fragm"]
    for r in (1, N // 2, N - 1):
        lens = [N + r, 2 * N + 5, N + N // 3]
        x = _noise(sum(lens), ch, 0.25, 100 + r)
        files = np.split(x, np.cumsum(lens)[:-1])
        outs, mx, flags = R.run_chain(d, rate, ch, bits, files, gapless=True)
        t = truth_f64(x, h, c["nout"])
        y = np.concatenate(outs)
        assert [o.shape[0] for o in outs] == lens
        assert np.abs(y - t).max() < 2e-6
        assert flags == [2, 3, 1]                      # out | in+out | in
        # without -g every file starts from silence
        outs2, _, flags2 = R.run_chain(d, rate, ch, bits, files, gapless=False)
        assert flags2 == [0, 0, 0]
        t1 = truth_f64(files[1], h, c["nout"])
        assert np.abs(outs2[1] - t1).max() < 2e-6
        assert np.abs(outs2[1] - outs[1]).max() > 1e-4


def test_quirk_no_handoff_when_length_is_multiple_of_fragm(dirs):
    R = H.reference()
    d, rate, ch, bits = dirs["crossfeed"]
    c, h = _impulses("crossfeed", dirs)
    N = c["fragm"]
    files = [_noise(2 * N, ch, 0.25, 1), _noise(N + 7, ch, 0.25, 2)]
    outs, _, flags = R.run_chain(d, rate, ch, bits, files, gapless=True)
    assert flags == [0, 0]                             # quirk 3: the tail of file 0 is lost
    assert np.abs(outs[1] - truth_f64(files[1], h, c["nout"])).max() < 2e-6


def test_quirk_successor_consumed_by_topup(dirs):
    R = H.reference()
    d, rate, ch, bits = dirs["crossfeed"]
    c, h = _impulses("crossfeed", dirs)
    N = c["fragm"]
    files = [_noise(N + 100, ch, 0.25, 3), _noise(50, ch, 0.25, 4)]   # 50 <= N - 100
    outs, _, flags = R.run_chain(d, rate, ch, bits, files, gapless=True)
    assert flags == [2, 1]
    assert outs[0].shape[0] == N + 100
    assert outs[1].shape[0] == 0                       # quirk 4: B never writes anything
    t = truth_f64(files[0], h, c["nout"])
    assert np.abs(outs[0] - t).max() < 2e-6


def test_reference_soundprocessor_equals_python_driven_oracle(dirs):
    """The facade under the reference's SoundProcessor is the same C restatement the
    Python tests drive directly: both routes must agree bit for bit."""
    R = H.reference()
    d, rate, ch, bits = dirs["tiny"]
    c, h = _impulses("tiny", dirs)
    x = _noise(5 * c["fragm"] + 17, ch, 0.5, 9)
    (y,), _, _ = R.run_chain(d, rate, ch, bits, [x])
    o = OracleConvproc(1, 1, c["size"])
    o.add(0, 0, (h[(0, 0)] ).astype(np.float32), 0)
    yo = run_blocks(o, x, c["fragm"])
    assert np.abs(y - yo).max() < 1e-6


def test_int_wire_formats_follow_libsndfile_scaling(dirs):
    R = H.reference()
    d, rate, ch, bits = dirs["tiny"]
    c, h = _impulses("tiny", dirs)
    xi = (np.random.default_rng(5).integers(-8000, 8000, (1000, 1))).astype(np.int16)
    (yi,), _, _ = R.run_chain(d, rate, ch, bits, [xi], in_format=H.SF_FORMAT_PCM_16, out_format=H.SF_FORMAT_PCM_16)
    t = truth_f64(xi.astype(np.float64) / 32768.0, h, 1)
    assert np.abs(yi.astype(np.int64) - np.rint(t * 32767.0).astype(np.int64)).max() <= 1
    x24 = (np.random.default_rng(6).integers(-2**21, 2**21, (1000, 1))).astype(np.int32)
    (y24,), _, _ = R.run_chain(d, rate, ch, bits, [x24], in_format=H.SF_FORMAT_PCM_24, out_format=H.SF_FORMAT_PCM_24)
    t = truth_f64(x24.astype(np.float64) / 8388608.0, h, 1)
    assert np.abs(y24.astype(np.int64) - np.rint(t * 8388607.0).astype(np.int64)).max() <= 4


def test_pooled_processor_reuse_quirk(dirs):
    """Quirk 5: a processor returned to the pool after an odd number of blocks and
    reused lags one block with the recalled zita behaviour; not with reset_is_fresh."""
    R = H.reference()
    d, rate, ch, bits = dirs["tiny"]
    c, h = _impulses("tiny", dirs)
    N = c["fragm"]
    a, b = _noise(N, ch, 0.5, 11), _noise(3 * N, ch, 0.5, 12)
    t = truth_f64(b, h, 1)
    try:
        R.drop_pool()
        R.set_reset_is_fresh(False)
        R.run_chain(d, rate, ch, bits, [a], gapless=False)          # one block, then pooled
        (y,), _, _ = R.run_chain(d, rate, ch, bits, [b], gapless=False)
        assert np.abs(y[:N]).max() == 0.0 and np.abs(y[N:] - t[:2 * N]).max() < 2e-6
        R.drop_pool()
        R.set_reset_is_fresh(True)
        R.run_chain(d, rate, ch, bits, [a], gapless=False)
        (y,), _, _ = R.run_chain(d, rate, ch, bits, [b], gapless=False)
        assert np.abs(y - t).max() < 2e-6
    finally:
        R.drop_pool()
        R.set_reset_is_fresh(False)


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "chains.npz")


def test_golden_fixture_is_what_the_reference_build_gives(dirs):
    """tests/golden/chains.npz against the reference build of this checkout, through the committed
    generator's own cases and inputs: every stored head, tail, maximum and flag bit for bit.  (The GPU
    suite checks the CUDA engine against the same file; this pins the file to its generator.)"""
    from golden.make_golden import CASES, case_inputs, summarize
    g = np.load(GOLDEN)
    R = H.reference()
    for name, fdir, lens, gapless, seed in CASES:
        d, rate, ch, bits = dirs[fdir]
        conf = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".conf")][0]
        fragm = R.load_config(conf, rate, ch)["fragm"]
        files = case_inputs(fragm, ch, lens, seed)
        R.drop_pool()
        outs, mx, flags = R.run_chain(d, rate, ch, bits, files, gapless=gapless)
        assert list(g[f"{name}/flags"]) == flags, name
        assert np.array_equal(g[f"{name}/max"], np.array(mx, np.float32)), name
        for k, y in enumerate(outs):
            assert y.shape[0] == int(g[f"{name}/{k}/frames"][0]), (name, k)
            head, tail, s, sa = summarize(y)
            assert np.array_equal(head, g[f"{name}/{k}/head"]), (name, k)
            assert np.array_equal(tail, g[f"{name}/{k}/tail"]), (name, k)
            assert np.array_equal(s, g[f"{name}/{k}/sum"]) and np.array_equal(sa, g[f"{name}/{k}/abssum"]), (name, k)


def test_golden_fixture_agrees_with_float64_truth(dirs):
    """The stored outputs against direct float64 convolution of the regenerated inputs: a gapless chain is
    one long signal as far as the stored hand-off flags say, a chain without -g starts every file from silence."""
    from golden.make_golden import CASES, case_inputs
    g = np.load(GOLDEN)
    for name, fdir, lens, gapless, seed in CASES:
        c, h = _impulses(fdir, dirs)
        ch = dirs[fdir][2]
        files = case_inputs(c["fragm"], ch, lens, seed)
        # a file whose flags carry "in" (bit 0) continues its predecessor's signal, any other starts from silence
        # (tiny_chain: file 1 disappears in file 0's top-up -- quirk 4 -- and file 2 then opens fresh)
        flags = [int(v) for v in g[f"{name}/flags"]]
        assert gapless or not any(flags)
        truths, k = [], 0
        while k < len(files):
            e = k + 1
            while e < len(files) and flags[e] & 1:
                e += 1
            t = truth_f64(np.concatenate(files[k:e]), h, c["nout"])
            truths += np.split(t, np.cumsum([f.shape[0] for f in files[k:e]])[:-1])
            k = e
        for k, t in enumerate(truths):
            n = int(g[f"{name}/{k}/frames"][0])
            if n == 0:
                continue                                           # quirk 4: consumed by the predecessor's top-up
            assert n == t.shape[0], (name, k)
            fs = max(1.0, np.abs(t).max())
            assert np.abs(g[f"{name}/{k}/head"] - t[:1024]).max() < 2e-6 * fs, (name, k)
            assert np.abs(g[f"{name}/{k}/tail"] - t[-1024:]).max() < 2e-6 * fs, (name, k)
            assert np.allclose(g[f"{name}/{k}/sum"], t.sum(axis=0), atol=1e-5 * n ** 0.5 + 1e-4), (name, k)
