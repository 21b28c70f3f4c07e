"""Synthetic workloads of the BASELINE.json configs (shapes, seeds: SURVEY.md section 8(d)).

/root/reference (and with it demo-filters/) does not exist on the GPU box, so
the filters used for measurement are synthetic stand-ins with exactly the
structure of the reference's config files:

  santalucia  demo-filters/SantaLucia/filter-44100.conf:44-53 --
              /convolver/new 2 2 256 204800 0.5; per channel one 178193-tap
              impulse (santalucia.wav frames 1400..) at delay 500, gain 4e-3,
              plus a 0.4 dirac at 0.  => fragm 8192, 25 partitions allocated,
              22 non-zero, 2 pairs.
  lowpass     demo-filters/lowpass/filter-44100.conf:6-8 -- size 65536, only
              the first 123 taps non-zero => 1 active partition of 8.
  roomcorr    65536-tap dense FIR per channel (96/192 kHz room correction).
  crossfeed   2x2 with /impulse/copy links on the cross paths.
  surround51  6x6: six diagonal 65536-tap IRs + LFE fed from the five mains.
"""
from __future__ import annotations

import numpy as np


class Workload:
    def __init__(self, name, fs, ninp, nout, size, adds, links=()):
        self.name, self.fs, self.ninp, self.nout, self.size = name, fs, ninp, nout, size
        self.adds = adds    # list of (inp, out, data float32, ind0)
        self.links = links  # list of (inp1, out1, inp2, out2)
        fragm = 8192
        while fragm > 64 and fragm >= 2 * size:
            fragm //= 2
        self.fragm = fragm

    def load(self, target):
        """target: anything with add(inp, out, data, ind0) / link(i1, o1, i2, o2)."""
        for (i, o, d, i0) in self.adds:
            target.add(i, o, d, i0)
        for l in self.links:
            target.link(*l)
        return target


def _decay_noise(n, seed, t60_frac=1.0):
    r = np.random.default_rng(seed)
    env = np.exp(-6.9 * np.arange(n) / (n * t60_frac))
    return (r.standard_normal(n) * env).astype(np.float32)


def santalucia(fs=44100):
    adds = []
    for ch in range(2):
        ir = _decay_noise(178193, 3 + ch)
        ir *= np.float32(25.0 / np.abs(ir).sum())      # sum|h| ~ 25 like the real IR after gain
        adds.append((ch, ch, ir, 500))
        adds.append((ch, ch, np.array([0.4], np.float32), 0))
    return Workload("santalucia", fs, 2, 2, 204800, adds)


def lowpass(fs=44100):
    n = np.arange(123) - 61
    h = (np.sinc(n / 8.0) / 8.0 * np.hamming(123)).astype(np.float32) * np.float32(0.75)
    return Workload("lowpass", fs, 2, 2, 65536, [(ch, ch, h, 0) for ch in range(2)])


def roomcorr(fs=96000):
    adds = []
    for ch in range(2):
        h = _decay_noise(65536, 10 + ch)
        h /= np.float32(np.abs(h).sum())
        adds.append((ch, ch, h, 0))
    return Workload("roomcorr", fs, 2, 2, 65536, adds)


def crossfeed(fs=44100):
    direct = _decay_noise(4096, 20)
    direct /= np.float32(np.abs(direct).sum() * 1.5)
    adds = [(0, 0, direct, 0), (1, 1, direct.copy(), 0)]
    links = [(0, 0, 0, 1), (1, 1, 1, 0)]
    return Workload("crossfeed", fs, 2, 2, 8192, adds, links)


def surround51(fs=48000, dense=False):
    adds = []
    for i in range(6):
        for o in range(6):
            if dense or i == o or (o == 3 and i != 3):
                h = _decay_noise(65536, 30 + 6 * i + o)
                h /= np.float32(np.abs(h).sum() * 6.0)
                adds.append((i, o, h, 0))
    return Workload("surround51_dense" if dense else "surround51", fs, 6, 6, 65536, adds)


def write_float_wav(path, data, rate):
    """Mono/multi-channel IEEE float32 RIFF/WAVE, [frames, channels]."""
    import struct
    data = np.asarray(data, "<f4")
    if data.ndim == 1:
        data = data[:, None]
    raw = data.tobytes()
    ch = data.shape[1]
    hdr = b"RIFF" + struct.pack("<I", 36 + len(raw)) + b"WAVE" + b"fmt " + struct.pack(
        "<IHHIIHH", 16, 3, ch, rate, rate * ch * 4, ch * 4, 32) + b"data" + struct.pack("<I", len(raw))
    with open(path, "wb") as f:
        f.write(hdr + raw)


def write_filter_dir(wl, directory):
    """The workload as a folve filter directory in the reference's .conf syntax:
    filter-<rate>.conf plus one float32 WAV per multi-tap impulse."""
    import os
    os.makedirs(directory, exist_ok=True)
    lines = [f"# synthetic stand-in for the {wl.name} workload (folve_b200/workloads.py)",
             f"/convolver/new {wl.ninp} {wl.nout} 256 {wl.size}"]
    for k, (i, o, d, i0) in enumerate(wl.adds):
        if len(d) == 1:
            lines.append(f"/impulse/dirac {i + 1} {o + 1} {float(d[0])!r} {i0}")
        else:
            name = f"ir_{k}.wav"
            write_float_wav(os.path.join(directory, name), d, wl.fs)
            lines.append(f"/impulse/read {i + 1} {o + 1} 1.0 {i0} 0 0 1 {name}")
    for (i1, o1, i2, o2) in wl.links:   # fcv/zita order: source first; config order: destination first
        lines.append(f"/impulse/copy {i2 + 1} {o2 + 1} {i1 + 1} {o1 + 1}")
    with open(os.path.join(directory, f"filter-{wl.fs}.conf"), "w") as f:
        f.write("\n".join(lines) + "\n")
    return directory


WORKLOADS = {
    "santalucia": santalucia,
    "lowpass": lowpass,
    "roomcorr96": lambda: roomcorr(96000),
    "roomcorr192": lambda: roomcorr(192000),
    "crossfeed": crossfeed,
    "surround51": surround51,
    "surround51_dense": lambda: surround51(dense=True),
}


def synthetic_pcm(nstreams, frames, nchan, peak, seed):
    """Uniform white noise, 16-bit-quantised like the configs' synthetic FLAC stand-in."""
    r = np.random.default_rng(seed)
    x = r.uniform(-peak, peak, (nstreams, frames, nchan))
    return (np.rint(x * 32768.0) / 32768.0).astype(np.float32)
