"""Synthetic filter directories in the reference's own .conf syntax (SURVEY.md section 8(c)
golden-vector list, item v): MIMO with /impulse/copy, per-output impulse sets,
/impulse/hilbert, quoted file names + /cd, short partitions, and broken files."""
from __future__ import annotations

import os

import numpy as np

from harness_py import write_wav


def _ir(n, seed, scale=1.0):
    r = np.random.default_rng(seed)
    env = np.exp(-5.0 * np.arange(n) / n)
    h = r.standard_normal(n) * env
    return (h / np.abs(h).sum() * scale).astype(np.float64)


def make_filter_dirs(root):
    """Creates the filter directories under `root`; returns {name: (dir, rate, channels, bits)}."""
    root = str(root)
    out = {}

    def mk(name, rate, ch, bits, conf, wavs=(), conf_name=None):
        d = os.path.join(root, name)
        os.makedirs(d, exist_ok=True)
        for (fn, data, fmt, wrate) in wavs:
            os.makedirs(os.path.dirname(os.path.join(d, fn)), exist_ok=True)
            write_wav(os.path.join(d, fn), data, wrate, fmt)
        with open(os.path.join(d, conf_name or f"filter-{rate}.conf"), "w") as f:
            f.write(conf)
        out[name] = (d, rate, ch, bits)

    # room correction, 96 kHz, 24-bit stereo IR file, both channels of one WAV
    ir = np.stack([_ir(65536, 10), _ir(65536, 11)], axis=1)
    mk("roomcorr96", 96000, 2, 24,
       "# synthetic room correction\n/convolver/new 2 2 1024 65536\n"
       "/impulse/read 1 1 1.0 0 0 0 1 ir96.wav\n/impulse/read 2 2 1.0 0 0 0 2 ir96.wav\n",
       [("ir96.wav", ir, "pcm24", 96000)])

    # 2x2 crossfeed: direct paths from a float WAV, cross paths linked + delayed dirac
    d = _ir(4096, 20, 0.6)
    mk("crossfeed", 44100, 2, 16,
       "/convolver/new 2 2 256 8192 1.0\n"
       "/impulse/read 1 1 1 0 0 0 1 direct.wav\n"
       "/impulse/read 2 2 1 0 0 0 1 direct.wav\n"
       "/impulse/copy 1 2 1 1\n/impulse/copy 2 1 2 2\n"
       "# later additions to the source are seen by the copies\n"
       "/impulse/dirac 1 1 0.1 13\n",
       [("direct.wav", d[:, None], "float", 44100)])

    # 5.1: six diagonal IRs + LFE (out 4) fed from the five mains; offsets, lengths, gains, delays
    wavs, lines = [], ["/convolver/new 6 6 1024 20000 0.3"]
    six = np.stack([_ir(12000, 30 + c) for c in range(6)], axis=1)
    wavs.append(("set/six.wav", six, "pcm32", 48000))
    lines.append("/cd set")
    for c in range(6):
        lines.append(f"/impulse/read {c + 1} {c + 1} 0.9 {10 * c} {5 * c} 0 {c + 1} six.wav")
    lfe = _ir(3000, 40)
    wavs.append(("set/lfe mix.wav", lfe[:, None], "pcm16", 48000))
    for c in (1, 2, 3, 5, 6):
        lines.append(f"/impulse/read {c} 4 0.2 100 0 2500 1 \"lfe mix.wav\"")
    lines.append("/input/name 1 Left")
    lines.append("/output/name 4 LFE")
    mk("surround51", 48000, 6, 24, "\n".join(lines) + "\n", wavs, conf_name="filter-48000-6-24.conf")

    # hilbert pair + dirac (complex matrix example of README.CONFIG.txt)
    mk("hilbert", 44100, 2, 16,
       "/convolver/new 2 2 64 4096\n/impulse/hilbert 1 1 0.5 2000 4000\n/impulse/dirac 2 2 0.5 2000\n"
       "/impulse/hilbert 1 2 -0.25 1000 1999\n")

    # small filter -> fragm 256 (size 200), truncation of an over-long read, mono
    mk("tiny", 22050, 1, 16,
       "/convolver/new 1 1 64 200\n/impulse/read 1 1 2.0 50 10 0 1 long.wav\n",
       [("long.wav", _ir(1000, 50)[:, None], "pcm16", 22050)])

    # escapes and single quotes in file names, tab separators, absolute /cd
    absd = os.path.join(root, "abs dir")
    os.makedirs(absd, exist_ok=True)
    write_wav(os.path.join(absd, "it's.wav"), _ir(300, 60)[:, None], 44100, "pcm16")
    mk("quoting", 44100, 1, 16,
       f"/convolver/new 1 1 64 512\n/cd \"{absd}\"\n/impulse/read\t1 1 1 0 0 0 1 it\\'s.wav\n"
       f"/impulse/read 1 1 0.5 100 0 0 1 \"it\\'s.wav\"\n")
    # a single quote inside double quotes is an error in zita-sstring.cc:80-95
    mk("quoting_bad", 44100, 1, 16,
       f"/convolver/new 1 1 64 512\n/cd \"{absd}\"\n/impulse/read 1 1 0.5 100 0 0 1 \"it's.wav\"\n")

    # --- broken or odd files -------------------------------------------------------
    mk("missing_wav", 44100, 2, 16,   # ERR_OTHER is swallowed AND stops parsing: later lines are ignored
       "/convolver/new 2 2 64 1000\n/impulse/dirac 1 1 0.5 0\n/impulse/read 2 2 1 0 0 0 1 nothere.wav\n"
       "/impulse/dirac 2 2 0.25 3\n")
    mk("no_convolver", 44100, 2, 16, "# nothing but a comment\n\n")
    mk("impulse_before_new", 44100, 2, 16, "/impulse/dirac 1 1 1.0 0\n/convolver/new 2 2 64 1000\n")
    mk("bad_ionum", 44100, 2, 16, "/convolver/new 2 2 64 1000\n/impulse/dirac 3 1 1.0 0\n")
    mk("syntax", 44100, 2, 16, "/convolver/new 2 2 64 1000\nthis is not a comment\n")
    mk("unknown_cmd", 44100, 2, 16, "/convolver/new 2 2 64 1000\n/impulse/bogus 1 1\n")
    mk("too_many_inputs", 44100, 2, 16, "/convolver/new 65 2 64 1000\n/impulse/dirac 1 1 1 0\n")
    mk("copy_self", 44100, 2, 16, "/convolver/new 2 2 64 1000\n/impulse/dirac 1 1 1 0\n/impulse/copy 1 1 1 1\n")
    mk("copy_no_source", 44100, 2, 16, "/convolver/new 2 2 64 1000\n/impulse/copy 2 2 1 1\n/impulse/dirac 1 1 1 0\n")
    mk("dirac_beyond_size", 44100, 1, 16, "/convolver/new 1 1 64 1000\n/impulse/dirac 1 1 1 1000\n/impulse/dirac 1 1 0.5 999\n")
    mk("size_zero", 44100, 1, 16, "/convolver/new 1 1 64 0\n")
    mk("indented_command", 44100, 1, 16, "/convolver/new 1 1 64 100\n  /impulse/dirac 1 1 1 0\n")
    return out
