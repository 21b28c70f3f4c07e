#!/usr/bin/env python
"""Generates tests/golden/chains.npz from the REFERENCE build (oracle/_ref/libfolve_ref.so:
the reference's own sound-processor.cc / zita-config.cc / processor-pool.cc on the
restated Convproc).  Run in the build container after `make -C oracle ref`:

    python tests/golden/make_golden.py

Inputs are regenerated from seeds by the tests; only the reference's outputs are stored
(float32, first and last 1024 frames of every file plus a float64 checksum per channel).
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import harness_py as H  # noqa: E402
from configs import make_filter_dirs  # noqa: E402

CASES = [
    # name, filter dir, file lengths as (blocks, extra frames), gapless, seed
    ("crossfeed_gapless", "crossfeed", [(1, 1), (2, 5), (1, 2730)], True, 101),
    ("tiny_chain", "tiny", [(3, 17), (0, 100), (2, 255)], True, 102),
    ("surround_single", "surround51", [(2, 1234)], False, 103),
    ("hilbert_nogap", "hilbert", [(1, 7), (1, 9)], False, 104),
    ("roomcorr_gapless", "roomcorr96", [(2, 4096), (1, 1)], True, 105),
]


def case_inputs(fragm, channels, lens, seed):
    r = np.random.default_rng(seed)
    files = []
    for (blocks, extra) in lens:
        n = blocks * fragm + extra
        files.append((np.rint(r.uniform(-0.25, 0.25, (n, channels)) * 32768) / 32768).astype(np.float32))
    return files


def summarize(y):
    return y[:1024].copy(), y[-1024:].copy(), y.astype(np.float64).sum(axis=0), np.abs(y.astype(np.float64)).sum(axis=0)


def main():
    R = H.reference()
    R.drop_pool()
    R.set_reset_is_fresh(True)
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        dirs = make_filter_dirs(tmp)
        for name, fdir, lens, gapless, seed in CASES:
            d, rate, ch, bits = dirs[fdir]
            conf = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".conf")][0]
            fragm = R.load_config(conf, rate, ch)["fragm"]
            files = case_inputs(fragm, ch, lens, seed)
            R.drop_pool()
            outs, mx, flags = R.run_chain(d, rate, ch, bits, files, gapless=gapless)
            out[f"{name}/max"] = np.array(mx, np.float32)
            out[f"{name}/flags"] = np.array(flags, np.int32)
            for k, y in enumerate(outs):
                head, tail, s, sa = summarize(y)
                out[f"{name}/{k}/frames"] = np.array([y.shape[0]], np.int64)
                out[f"{name}/{k}/head"] = head
                out[f"{name}/{k}/tail"] = tail
                out[f"{name}/{k}/sum"] = s
                out[f"{name}/{k}/abssum"] = sa
    np.savez_compressed(os.path.join(HERE, "chains.npz"), **out)
    print("wrote", os.path.join(HERE, "chains.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
