"""The reference's own callers of SoundProcessor -- FolveFilesystem::GetOrCreateHandler,
FileHandler::Read, ConversionBuffer::FillUntil, ConvolveFileHandler::AddMoreSoundData /
PassoverProcessor (/root/reference/convolve-file-handler.cc:328-351,370-424), BufferThread,
FileHandlerCache and the reference's ProcessorPool (/root/reference/processor-pool.cc:48-118) --
compiled UNMODIFIED and run

  * on the reference's SoundProcessor (CPU, not gpu): pins oracle/dropin_driver.cc and the caller
    protocol restated in folve_b200/host/harness.cc against the real thing;
  * on THIS repository's SoundProcessor and CUDA engine (gpu): "drops in unchanged", through
    gapless albums, the quirk cases and the prebuffer thread.
"""
import os

import numpy as np
import pytest

import dropin_py as D
import harness_py as H
from configs import make_filter_dirs
from harness_py import write_wav


def _noise(frames, ch, peak, seed):
    r = np.random.default_rng(seed)
    return (np.rint(r.uniform(-peak, peak, (frames, ch)) * 32768) / 32768).astype(np.float32)


@pytest.fixture(scope="module")
def library(tmp_path_factory):
    """music/<album>/NN.wav (16-bit stereo 44.1 kHz) + the filter directories of tests/configs.py"""
    root = tmp_path_factory.mktemp("mount")
    filters = root / "filters"
    dirs = make_filter_dirs(filters)
    N = 8192   # block size of the crossfeed filter (size 8192 -> fragm 8192)
    music = root / "music"
    albums = {
        # ends inside a block -> hand-off; a multiple of the block size -> none (quirk 3);
        # a successor swallowed whole by the top-up (quirk 4); the rest
        "a": [3 * N + 100, 2 * N, N + 5000, 200, 2 * N + 77],
        "b": [N // 2, N // 2 + 1, 5 * N + 1],
    }
    files = {}
    for name, lengths in albums.items():
        os.makedirs(music / name)
        for k, n in enumerate(lengths):
            x = _noise(n, 2, 0.25, 100 * len(files) + k)
            write_wav(str(music / name / f"{k + 1:02d}.wav"), x, 44100, "pcm16")
            files[f"/{name}/{k + 1:02d}.wav"] = x
    return dict(music=str(music), filters=str(filters), dirs=dirs, files=files, albums=albums)


def _read_album(mount, lib, album):
    out = []
    for path in sorted(p for p in lib["files"] if p.startswith(f"/{album}/")):
        out.append((path,) + mount.read(path, 2)[:3])
    return out


@pytest.mark.skipif(not (D.have_refstack() and H.have_reference()), reason="needs oracle/_ref (make -C oracle ref refstack)")
@pytest.mark.parametrize("gapless", [True, False])
def test_reference_filesystem_equals_the_restated_caller(library, gapless):
    """the reference's real AddMoreSoundData / PassoverProcessor and the restatement in harness.cc
    deliver the same 24-bit samples, frame counts and gapless flags (both on the reference's
    SoundProcessor, CPU)"""
    R = H.reference()
    R.drop_pool()
    R.set_reset_is_fresh(True)
    m = D.Mount(D.REFSTACK_SO, library["music"], library["filters"], "crossfeed", gapless=gapless)
    assert m.variant == "refstack"
    d, rate, ch, bits = library["dirs"]["crossfeed"]
    for album in library["albums"]:
        got = _read_album(m, library, album)
        xs = [library["files"][p] for (p, _, _, _) in got]
        ys, mx, flags = R.run_chain(d, rate, ch, bits, xs, gapless=gapless, out_format=H.SF_FORMAT_PCM_24)
        for (p, y, fl, m_out), yr, fr in zip(got, ys, flags):
            assert fl & 4, p                       # a convolving handler, not pass-through
            assert y.shape == yr.shape, p
            assert np.array_equal(y, yr), p
            assert (fl & 3) == fr, (p, fl, fr)


@pytest.mark.gpu
@pytest.mark.skipif(not (D.have_dropin() and D.have_refstack()), reason="needs oracle/_ref (make -C oracle dropin refstack)")
@pytest.mark.parametrize("gapless,pre_buffer", [(True, 0), (False, 0), (True, 128 << 10)])
def test_reference_callers_on_the_b200_engine(library, gapless, pre_buffer):
    """folve's own filesystem layer, unmodified, on this repository's SoundProcessor: same frame
    counts and gapless flags as on the reference's SoundProcessor, samples within the 24-bit
    bound of tests/test_soundprocessor_gpu.py (float32 noise floor of a 16384-point transform)"""
    ref = D.Mount(D.REFSTACK_SO, library["music"], library["filters"], "crossfeed", gapless=gapless)
    gpu = D.Mount(D.DROPIN_SO, library["music"], library["filters"], "crossfeed", gapless=gapless,
                  pre_buffer_bytes=pre_buffer)
    assert gpu.variant == "dropin"
    worst, n_off, n_all = 0, 0, 0
    for album in library["albums"]:
        a, b = _read_album(gpu, library, album), _read_album(ref, library, album)
        for (p, y, fl, mx), (_, yr, flr, mxr) in zip(a, b):
            assert fl & 4, p
            assert y.shape == yr.shape, p
            assert (fl & 3) == (flr & 3), (p, fl, flr)
            d = np.abs(y.astype(np.int64) - yr.astype(np.int64))
            worst = max(worst, int(d.max()) if d.size else 0)
            n_off += int((d > 1).sum())
            n_all += d.size
            assert mx == pytest.approx(mxr, abs=2e-6), p
    assert worst <= 4, worst
    assert n_off <= 0.03 * n_all


@pytest.mark.skipif(not (D.have_refstack() and H.have_reference()), reason="needs oracle/_ref (make -C oracle ref refstack)")
def test_random_albums_reference_filesystem_equals_the_restated_caller(tmp_path):
    """The same comparison over 12 random albums of a mono 22.05 kHz library with the block size of 256 frames:
    track lengths drawn from the classes the hand-off logic distinguishes (whole blocks, a few frames more,
    one frame short of a block, shorter than a block, a single frame), gapless."""
    filters = tmp_path / "filters"
    dirs = make_filter_dirs(filters)
    d, rate, ch, bits = dirs["tiny"]
    N = 256
    r = np.random.default_rng(31)
    music = tmp_path / "music"
    albums = {}
    for a in range(12):
        lengths = []
        for _ in range(int(r.integers(2, 7))):
            k = int(r.integers(0, 4))
            lengths.append(max(1, [k * N, k * N + int(r.integers(1, 40)), (k + 1) * N - 1, int(r.integers(2, N)), 1]
                                  [int(r.integers(0, 5))]))
        name = f"album{a:02d}"
        os.makedirs(music / name)
        albums[name] = []
        for k, n in enumerate(lengths):
            x = _noise(n, ch, 0.25, 1000 * a + k)
            write_wav(str(music / name / f"{k + 1:02d}.wav"), x, rate, "pcm16")
            albums[name].append((f"/{name}/{k + 1:02d}.wav", x))
    R = H.reference()
    R.drop_pool()
    R.set_reset_is_fresh(True)
    m = D.Mount(D.REFSTACK_SO, str(music), str(filters), "tiny", gapless=True)
    handed = 0
    for name, tracks in albums.items():
        got = [(p,) + m.read(p, ch)[:3] for (p, _) in tracks]
        ys, mx, flags = R.run_chain(d, rate, ch, bits, [x for (_, x) in tracks], gapless=True,
                                    out_format=H.SF_FORMAT_PCM_24)
        for (p, y, fl, m_out), yr, fr in zip(got, ys, flags):
            assert fl & 4, p
            assert y.shape == yr.shape, (p, [x.shape[0] for (_, x) in tracks])
            assert np.array_equal(y, yr), p
            assert (fl & 3) == fr, (p, fl, fr, [x.shape[0] for (_, x) in tracks])
            handed += fr & 1
    assert handed >= 10
