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


def _random_config(r, wavs):
    """One random jconvolver-style config: mostly well-formed commands with random parameters, sometimes a
    token damaged, dropped or added, sometimes a command out of order."""
    ninp, nout = int(r.integers(1, 5)), int(r.integers(1, 5))
    size = int(r.choice([0, 1, 100, 256, 1000, 5000, 5000, 8192, 20000, 20000]))
    part = int(r.choice([0, 64, 256, 1000, 1024, 8192]))
    lines = []
    if r.random() < 0.9:
        dens = ["", " 0.5", " 1", " x"][int(r.choice(4, p=[0.6, 0.2, 0.15, 0.05]))]
        lines.append(f"/convolver/new {ninp} {nout} {part} {size}{dens}")

    def io(n):
        return int(r.integers(1, n + 1)) if r.random() < 0.975 else int(r.choice([0, n + 1, -1, 65]))

    def num(kind):
        if r.random() < 0.012:
            return str(r.choice(["abc", "", "1e", "0x10", "-", "3.5.1", "1,5"]))
        if kind == "gain":
            return str(r.choice(["1", "0.5", "-0.25", "1e-1", "2.", ".75", "+1", "-1.0e0"]))
        return str(int(r.choice([0, 0, 0, 1, 7, 50, 63, 64, 255, 256, 300, 999, 4000, 30000, -1])))

    for _ in range(int(r.integers(0, 12))):
        c = r.random()
        if c < 0.35:
            w = wavs[int(r.integers(len(wavs)))] if r.random() < 0.9 else ("nonexistent.wav", 1, 0)
            name = w[0] if r.random() < 0.7 else '"' + w[0] + '"'
            chan = int(r.integers(1, w[1] + 1)) if r.random() < 0.9 else w[1] + 1
            # offset + length stay inside the file: past its end the reference's read loop (zita-config.cc:146-171)
            # gets 0 frames for ever and Create never returns
            off = int(r.choice([0, 0, 1, 7, w[2] // 2, w[2], w[2] + 1]))
            length = 0 if r.random() < 0.5 or off >= w[2] else int(r.integers(1, w[2] - off + 1))
            ln = f"/impulse/read {io(ninp)} {io(nout)} {num('gain')} {num('d')} {off} {length} {chan} {name}"
        elif c < 0.55:
            ln = f"/impulse/dirac {io(ninp)} {io(nout)} {num('gain')} {num('d')}"
        elif c < 0.70:
            ln = f"/impulse/hilbert {io(ninp)} {io(nout)} {num('gain')} {num('d')} {int(r.choice([0, 63, 64, 100, 999, 4096, 4097]))}"
        elif c < 0.82:
            ln = f"/impulse/copy {io(ninp)} {io(nout)} {io(ninp)} {io(nout)}"
        elif c < 0.88:
            ln = str(r.choice(["/cd sub", "/cd .", "/cd", '/cd "sub"', "/cd nowhere"]))
        elif c < 0.93:
            ln = str(r.choice(["/input/name 1 left", "/output/name 2 right", "/input/name", "/output/name 9 x y z"]))
        elif c < 0.97:
            ln = str(r.choice(["# a comment", "", "   ", "\t", "#/impulse/dirac 1 1 1 0", "  # indented comment"]))
        else:
            ln = str(r.choice(["/unknown/cmd 1 2", "garbage", " /impulse/dirac 1 1 1 0", "/impulse/dirac", "/convolver/new 2 2 64 100"]))
        m = r.random()
        toks = ln.split(" ")
        if m < 0.02 and len(toks) > 1:
            toks.pop(int(r.integers(1, len(toks))))
        elif m < 0.04:
            toks.append(str(r.choice(["extra", "7", "#c"])))
        elif m < 0.08:
            ln = ln.replace(" ", str(r.choice(["  ", "\t", " \t "])))
            toks = None
        lines.append(" ".join(toks) if toks is not None else ln)
    if r.random() < 0.08:
        lines.insert(0, lines.pop()) if lines else None        # a command in front of /convolver/new
    end = "\n" if r.random() < 0.9 else ""
    return ("\r\n" if r.random() < 0.05 else "\n").join(lines) + end


def _fuzz_main(root, count=600):
    from harness_py import write_wav
    r = np.random.default_rng(2024)
    d = os.path.join(root, "fuzz")
    os.makedirs(os.path.join(d, "sub"))
    wavs = [("mono16.wav", 1, 3000), ("stereo24.wav", 2, 700), ("quad_float.wav", 4, 9000), ("sub/short.wav", 1, 10),
            ("with space.wav", 2, 257)]
    for k, (fn, ch, n) in enumerate(wavs):
        data = np.random.default_rng(50 + k).uniform(-0.5, 0.5, (n, ch)) * np.exp(-np.arange(n) / (n / 4))[:, None]
        write_wav(os.path.join(d, fn), data, [44100, 48000, 44100, 44100, 96000][k],
                  ["pcm16", "pcm24", "float", "pcm16", "pcm32"][k])
    seen = {"created": 0, "failed": 0, "pairs": 0, "second_new": 0}
    for k in range(count):
        text = _random_config(r, wavs)
        conf = os.path.join(d, f"filter-{k}.conf")
        with open(conf, "w") as f:
            f.write(text)
        rate, ch = (44100, 96000)[k % 2], (2, 1, 6)[k % 3]       # what Create() is called with
        print(f"config {k} ({rate} Hz, {ch} channels): {text!r}", flush=True)
        if sum(ln.startswith("/convolver/new") for ln in text.splitlines()) > 1:
            # a second /convolver/new that is reached: the reference keeps the first line's engine under the second
            # line's ninp / nout / size / fragm (Create then gives NULL or a mismatched processor); here such a
            # file never yields a processor (filter-config.cc).  Same verdict from the parser.
            a, b = H.reference().load_config(conf, rate, ch), H.product().load_config(conf, rate, ch)
            assert a["rc"] == b["rc"] and (b["created"] == 0 or a["created"] == 1)
            if a["created"] != b["created"]:
                seen["second_new"] += 1
                continue
        a, b = _compare(conf, rate, ch)
        seen["created" if a["created"] else "failed"] += 1
        seen["pairs"] += len(a["pairs"])
    print("SEEN", seen["created"], seen["failed"], seen["pairs"], flush=True)


def test_random_configs_load_identically(tmp_path):
    """Differential test of the loader against the reference's parser: 600 seeded random configs (well-formed and
    damaged) must give the same verdict (created or not, error code) and bit-identical impulses for every pair.
    Runs in a child process with a time limit: a parser that never returns is a failure, not a hung suite."""
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    code = f"import sys; sys.path.insert(0, {here!r}); import test_config_parity as T; T._fuzz_main({str(tmp_path)!r})"
    try:
        p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=240, cwd=here)
    except subprocess.TimeoutExpired as e:
        out = e.stdout.decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
        raise AssertionError("loader did not return on:\n" + "\n".join(out.splitlines()[-1:]))
    tail = "\n".join(p.stdout.splitlines()[-2:])
    assert p.returncode == 0, tail + "\n" + p.stderr[-2000:]
    created, failed, pairs = (int(v) for v in p.stdout.splitlines()[-1].split()[1:])
    assert created > 200 and failed > 200 and pairs > 150, (created, failed, pairs)


def test_quoted_string_scanner_equals_the_reference():
    """ScanString against the reference's sstring (zita-sstring.cc, compiled unmodified into oracle/_ref) on
    20 000 random byte strings over the characters that matter to it (quotes, backslashes, blanks, control
    characters) and on several destination sizes: same return value, same destination bytes."""
    import ctypes as C
    ref = C.CDLL(H.REFERENCE_SO)._Z7sstringPKcPci
    own = C.CDLL(H.PRODUCT_SO)._ZN10folve_b20010ScanStringEPKcPci
    for f in (ref, own):
        f.restype = C.c_int
        f.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    r = np.random.default_rng(77)
    alphabet = np.frombuffer(b"\"\"\"'''\\\\\\   \t\nab/.x#", np.uint8)
    checked = nonzero = 0
    for k in range(20000):
        n = int(r.integers(0, 24))
        src = bytes(r.choice(alphabet, n)) if r.random() < 0.95 else bytes(r.integers(1, 256, n, dtype=np.uint8))
        size = int(r.choice([1, 2, 3, 5, 8, 64, 1024]))
        da, db = C.create_string_buffer(b"\xee" * 1100, 1100), C.create_string_buffer(b"\xee" * 1100, 1100)
        ra, rb = ref(src, da, size), own(src, db, size)
        assert ra == rb, (src, size, ra, rb)
        if ra:
            assert da.raw[:size] == db.raw[:size], (src, size, da.raw[:size], db.raw[:size])
            nonzero += 1
        assert da.raw[size:] == db.raw[size:] == b"\xee" * (1100 - size), (src, size)   # nothing written past `size`
        checked += 1
    assert nonzero > 5000


def test_block_size_and_partition_count_for_every_size_class(tmp_path):
    """/convolver/new over the sizes around every power of two up to MAXSIZE and beyond it: same verdict, block size
    (zita-fconfig.cc:74-77) and partition count on both sides; the dirac at the last valid position lands in the last
    partition on both.  (From 2^31 on the reference reads the size into a signed int and its range check no longer
    sees it; this loader refuses everything above MAXSIZE.)"""
    sizes = [0, 1, 2, 3]
    for k in range(5, 21):
        sizes += [2 ** k - 1, 2 ** k, 2 ** k + 1]
    sizes += [0x100000 + 2, 2 ** 30, 2 ** 31 - 1]
    seen = set()
    for size in sizes:
        conf = tmp_path / f"filter-{size}.conf"
        conf.write_text(f"/convolver/new 1 1 0 {size}\n/impulse/dirac 1 1 0.5 {max(size - 1, 0)}\n")
        a, b = _compare(str(conf), 44100, 1)
        if a["created"]:
            seen.add(a["fragm"])
            assert a["fragm"] == b["fragm"] and a["npar"] == b["npar"]
    assert seen == {64, 128, 256, 512, 1024, 2048, 4096, 8192}


def test_impulse_files_are_stamped_for_the_staleness_check(tmp_path):
    """SURVEY 8(f)4 / the reference's TODO at sound-processor.cc:129-133, host side without a GPU: the loader records
    every file an /impulse/read line reached -- the missing one included, not the one behind the line parsing stopped
    at -- and the record goes stale when one of them is rewritten with the same size within the same second, changes
    size, disappears, or when the missing one appears."""
    import ctypes as C
    import time
    from harness_py import write_wav
    d = tmp_path / "stale"
    (d / "sub").mkdir(parents=True)
    ir = np.random.default_rng(1).uniform(-0.5, 0.5, (300, 1))
    write_wav(str(d / "a.wav"), ir, 44100, "pcm16")
    write_wav(str(d / "sub" / "b.wav"), ir[::-1], 44100, "pcm16")
    conf = d / "filter-44100.conf"
    conf.write_text("/convolver/new 2 2 64 1000\n/impulse/read 1 1 1 0 0 0 1 a.wav\n/cd sub\n/impulse/read 2 2 1 0 0 0 1 b.wav\n"
                    "/impulse/dirac 1 2 0.1 5\n/impulse/read 2 1 1 0 0 0 1 later.wav\n/impulse/read 1 2 1 0 0 0 1 never-reached.wav\n")
    L = H.product().L
    L.fh_config_impulse_stamps.restype = C.c_int
    L.fh_config_impulse_stamps.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    L.fh_config_open.restype = C.c_void_p

    def load():
        return C.c_void_p(L.fh_config_open(str(conf).encode(), 44100, 2))

    def state(h):
        cur = C.c_int(-1)
        return L.fh_config_impulse_stamps(h, C.byref(cur)), cur.value

    h = load()
    assert state(h) == (3, 1)                       # a.wav, sub/b.wav, sub/later.wav (missing: parsing stops there)
    time.sleep(0.01)
    write_wav(str(d / "a.wav"), ir * 0.5, 44100, "pcm16")          # same size, same second, new content
    assert state(h) == (3, 0)
    L.fh_config_close(h)
    h = load()
    assert state(h) == (3, 1)
    write_wav(str(d / "sub" / "later.wav"), ir, 44100, "pcm16")     # the missing file appears
    assert state(h) == (3, 0)
    L.fh_config_close(h)
    h = load()
    assert state(h) == (4, 1)                       # now the line after it is reached as well
    write_wav(str(d / "sub" / "b.wav"), ir[:100], 44100, "pcm16")   # size change
    assert state(h) == (4, 0)
    L.fh_config_close(h)
    h = load()
    os.remove(d / "a.wav")
    assert state(h) == (4, 0)
    L.fh_config_close(h)
    R = H.reference().L
    R.fh_config_impulse_stamps.restype = C.c_int
    R.fh_config_impulse_stamps.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    R.fh_config_open.restype = C.c_void_p
    hr = C.c_void_p(R.fh_config_open(str(conf).encode(), 44100, 2))
    cur = C.c_int(0)
    assert R.fh_config_impulse_stamps(hr, C.byref(cur)) == -1       # the reference keeps no such record
    R.fh_config_close(hr)
