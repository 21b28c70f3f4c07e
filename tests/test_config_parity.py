"""CPU: this repo's filter-config loader against the reference's OWN parser
(zita-config.cc / zita-fconfig.cc / zita-audiofile.cc / zita-sstring.cc compiled
unmodified from /root/reference into oracle/_ref/libfolve_ref.so).

Both sides report, per (input, output) pair, the accumulated time-domain impulse
exactly as it is handed to the convolver; the comparison is bit-exact."""
import os

import numpy as np
import pytest

import harness_py as H
from configs import make_filter_dirs

pytestmark = pytest.mark.skipif(not (H.have_reference() and H.have_product()),
                                reason="needs oracle/_ref/libfolve_ref.so and folve_b200/libfolve_host.so")

DEMO = "/root/reference/demo-filters"


def _conf_of(d, rate, ch, bits):
    for n in (f"filter-{rate}-{ch}-{bits}.conf", f"filter-{rate}-{ch}.conf", f"filter-{rate}.conf"):
        p = os.path.join(d, n)
        if os.path.exists(p):
            return p
    raise FileNotFoundError(d)


def _compare(conf, rate, ch):
    a = H.reference().load_config(conf, rate, ch)
    b = H.product().load_config(conf, rate, ch)
    assert a["created"] == b["created"], (a["rc"], b["rc"])
    assert (a["rc"] == 0) == (b["rc"] == 0)
    if a["rc"] not in (0,):
        assert a["rc"] == b["rc"]          # same error code (zita-config.h:51 numbering)
    if not a["created"]:
        return a, b
    for k in ("ninp", "nout", "size", "fragm", "npar"):
        assert a[k] == b[k], k
    assert set(a["pairs"]) == set(b["pairs"])
    for k in a["pairs"]:
        assert a["pairs"][k][0] == b["pairs"][k][0], k
        assert np.array_equal(a["pairs"][k][1], b["pairs"][k][1]), k
    return a, b


@pytest.mark.skipif(not os.path.isdir(DEMO), reason="reference demo-filters not present (GPU box)")
@pytest.mark.parametrize("name,rate", [("lowpass", 44100), ("highpass", 44100), ("echo", 44100),
                                       ("echo", 192000), ("SantaLucia", 44100)])
def test_every_demo_filter_loads_identically(name, rate):
    a, b = _compare(f"{DEMO}/{name}/filter-{rate}.conf", rate, 2)
    assert a["created"] == 1 and a["fragm"] == 8192
    if name == "SantaLucia":
        # 178193 frames of santalucia.wav from offset 1400 at delay 500 + dirac 0.4 at 0
        h = b["pairs"][(0, 0)][1] * (2 * 8192)
        assert a["npar"] == 25
        assert np.flatnonzero(h)[-1] < 500 + 178193
        assert h[0] == pytest.approx(0.4)
        assert 25.0 < np.abs(h).sum() < 26.0
    if name == "lowpass":
        assert np.flatnonzero(b["pairs"][(1, 1)][1])[-1] < 123


def test_synthetic_filter_directories(tmp_path):
    dirs = make_filter_dirs(tmp_path)
    created = {}
    for name, (d, rate, ch, bits) in dirs.items():
        a, b = _compare(_conf_of(d, rate, ch, bits), rate, ch)
        created[name] = a["created"]
    # which broken files still yield a processor
    assert created["missing_wav"] == 1          # quirk 6: ERR_OTHER swallowed
    assert created["too_many_inputs"] == 0
    assert created["no_convolver"] == 0
    assert created["impulse_before_new"] == 0   # ERR_NOCONV
    assert created["bad_ionum"] == 0
    assert created["syntax"] == 0
    assert created["unknown_cmd"] == 0
    assert created["copy_self"] == 0            # ERR_PARAM
    assert created["copy_no_source"] == 1
    assert created["indented_command"] == 0     # a command must start in column 0
    assert created["size_zero"] == 1
    assert created["quoting_bad"] == 0
    for ok in ("roomcorr96", "crossfeed", "surround51", "hilbert", "tiny", "quoting", "dirac_beyond_size"):
        assert created[ok] == 1, ok


def test_loader_semantics_spot_checks(tmp_path):
    dirs = make_filter_dirs(tmp_path)
    P = H.product()
    d, rate, ch, bits = dirs["missing_wav"]
    c = P.load_config(_conf_of(d, rate, ch, bits), rate, ch)
    # parsing stopped at the missing file: the dirac on (2,2) after it is NOT loaded
    assert set(c["pairs"]) == {(0, 0)}
    d, rate, ch, bits = dirs["crossfeed"]
    c = P.load_config(_conf_of(d, rate, ch, bits), rate, ch)
    assert c["pairs"][(0, 1)][0] == 2 and c["pairs"][(1, 0)][0] == 2       # links
    assert np.array_equal(c["pairs"][(0, 1)][1], c["pairs"][(0, 0)][1])   # incl. the later dirac
    assert c["pairs"][(0, 0)][1][13] != c["pairs"][(1, 1)][1][13]
    d, rate, ch, bits = dirs["tiny"]
    c = P.load_config(_conf_of(d, rate, ch, bits), rate, ch)
    assert c["fragm"] == 256 and c["npar"] == 1
    h = c["pairs"][(0, 0)][1]
    assert np.all(h[:50] == 0) and h[50] != 0 and len(h) == 256
    assert np.all(h[200:] == 0)                 # truncated to size - delay = 150 frames
    d, rate, ch, bits = dirs["hilbert"]
    c = P.load_config(_conf_of(d, rate, ch, bits), rate, ch)
    h = c["pairs"][(0, 0)][1] * (2 * c["fragm"])
    assert h[2000] == 0 and h[2001] == pytest.approx(-0.5 * 2 / np.pi * (0.43 + 0.57 * np.cos(np.pi / 2000)), rel=1e-6)
    assert h[1999] == -h[2001]
    assert (0, 1) in c["pairs"]                 # delay 1000 >= length / 2 = 999: kept
