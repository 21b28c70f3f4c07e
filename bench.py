#!/usr/bin/env python
"""bench.py -- the convolution hot path on N B200s of one node.

Metric (BASELINE.json): convolved audio-seconds per second (x realtime).
A step = one block of `fragm` frames for every stream of the batch (forward FFT
of each input channel, complex MAC over the partition history, inverse FFT +
overlap-add) -- SoundProcessor::Process() (sound-processor.cc:98-127) for
`--streams` independent SoundProcessors per GPU at once.

  value : device-resident (PCM already in HBM), CUDA events, max over ranks
  e2e   : through the C-ABI call fcv_batch_process with pinned HOST buffers,
          host->device and device->host copies inside the timed region
  roofline : the complex-MAC kernel against the measured HBM copy bandwidth
  cpu_baseline : the zita-convolver restatement on all host cores (rank 0, N=1)

`--impl reference` times the CPU path (one Convproc per stream, one stream per
thread, all cores) on the same workload; see DESIGN.md section "Measurement".
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from folve_b200 import workloads  # noqa: E402


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU every 100 ms (NVML)."""

    BITS = {
        0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
        0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting",
    }

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._th = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.BITS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.nv:
            self._th = threading.Thread(target=self._run, daemon=True)
            self._th.start()

    def stop(self):
        self._stop.set()
        if self._th:
            self._th.join()
        return {
            "sm_mhz": statistics.median(self.samples) if self.samples else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
        }


# --------------------------------------------------------------------------- CPU arm
def _oracle_lib():
    so = os.path.join(ROOT, "oracle", "libzita_oracle.so")
    if not os.path.exists(so):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    L = C.CDLL(so)

    class Imp(C.Structure):
        _fields_ = [("inp", C.c_int), ("out", C.c_int), ("ind0", C.c_int), ("len", C.c_int),
                    ("data", C.POINTER(C.c_float))]

    L.zb_run.restype = C.c_double
    L.zb_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_uint, C.c_int,
                         C.POINTER(Imp), C.c_float, C.POINTER(C.c_double)]
    return L, Imp


def cpu_run(wl, nthreads, nblocks, amplitude=0.03):
    """One Convproc per stream, one stream per thread; returns (audio_s, wall_s)."""
    L, Imp = _oracle_lib()
    keep = [np.ascontiguousarray(d, np.float32) for (_, _, d, _) in wl.adds]
    imps = (Imp * len(wl.adds))()
    for k, (i, o, d, i0) in enumerate(wl.adds):
        imps[k].inp, imps[k].out, imps[k].ind0, imps[k].len = i, o, i0, len(d)
        imps[k].data = keep[k].ctypes.data_as(C.POINTER(C.c_float))
    cs = C.c_double(0)
    wall = L.zb_run(nthreads, nblocks, wl.ninp, wl.nout, wl.size, wl.fragm, len(wl.adds), imps,
                    amplitude, C.byref(cs))
    if wall <= 0:
        raise RuntimeError("cpu baseline failed")
    return nthreads * nblocks * wl.fragm / wl.fs, wall


REF_SO = os.path.join(ROOT, "oracle", "_ref", "libfolve_ref.so")
HOST_SO = os.path.join(ROOT, "folve_b200", "libfolve_host.so")


def harness_run(so, wl, filter_dir, nthreads, nblocks):
    """The block loop of ConvolveFileHandler::AddMoreSoundData through SoundProcessor
    (FillBuffer / WriteProcessed), one file per thread; returns (audio_s, wall_s)."""
    L = C.CDLL(so)
    L.fh_bench_threads.restype = C.c_double
    L.fh_bench_threads.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    fragm = C.c_int(0)
    wall = L.fh_bench_threads(filter_dir.encode(), wl.fs, wl.ninp, 16, nthreads, nblocks, C.byref(fragm))
    if wall <= 0:
        raise RuntimeError(f"fh_bench_threads failed in {so}")
    return nthreads * nblocks * fragm.value / wl.fs, wall


def library_run(wl, filter_dir, nchains, files_per_chain, seconds_per_file, blocks_per_step, threads, pcm16):
    """BatchConvolver (folve_b200/host/batch-convolver.cc): gapless album chains of in-memory files
    through the batched submit layer; returns (audio_s, wall_s)."""
    L = C.CDLL(HOST_SO)
    L.fh_bench_library.restype = C.c_double
    L.fh_bench_library.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_long, C.c_int, C.c_int,
                                   C.c_int, C.POINTER(C.c_double)]
    cfg = os.path.join(filter_dir, f"filter-{wl.fs}-{wl.ninp}.conf")
    if not os.path.exists(cfg):
        cfg = os.path.join(filter_dir, f"filter-{wl.fs}.conf")
    audio = C.c_double(0)
    wall = L.fh_bench_library(cfg.encode(), wl.fs, wl.ninp, 1, nchains, files_per_chain,
                              int(seconds_per_file * wl.fs), blocks_per_step, threads, 1 if pcm16 else 0, C.byref(audio))
    if wall <= 0:
        raise RuntimeError("fh_bench_library failed")
    return audio.value, wall


def cpu_arm(wl, filter_dir):
    """-> (kind, runner(nthreads, nblocks) -> (audio_s, wall_s), description)"""
    if os.path.exists(REF_SO):
        return ("reference", lambda nt, nb: harness_run(REF_SO, wl, filter_dir, nt, nb),
                "the reference's own sound-processor.cc / zita-config.cc / processor-pool.cc (compiled unmodified "
                "into oracle/_ref) on the restated zita-convolver (oracle/zita_oracle.c, own float32 FFT, no FFTW)")
    return ("port", lambda nt, nb: cpu_run(wl, nt, nb), "restated Convproc (oracle/zita_oracle.c)")


def fft_calibration(n=16384, reps=300):
    """The CPU arm's FFT is the oracle's own (FFTW is not installed): time it next to
    pocketfft (scipy, float32) on one core so that the GPU/CPU ratio cannot be silently
    flattered by a slow CPU transform."""
    L = C.CDLL(os.path.join(ROOT, "oracle", "libzita_oracle.so"))
    L.offt_plan_create.restype = C.c_void_p
    L.offt_plan_create.argtypes = [C.c_int]
    L.offt_r2c.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.offt_plan_destroy.argtypes = [C.c_void_p]
    x = np.random.default_rng(0).standard_normal(n).astype(np.float32)
    out = np.zeros(n + 2, np.float32)
    plan = L.offt_plan_create(n)
    for _ in range(20):
        L.offt_r2c(plan, x.ctypes.data, out.ctypes.data)
    t0 = time.perf_counter()
    for _ in range(reps):
        L.offt_r2c(plan, x.ctypes.data, out.ctypes.data)
    t_oracle = (time.perf_counter() - t0) / reps
    L.offt_plan_destroy(plan)
    res = {"n": n, "oracle_r2c_us": round(1e6 * t_oracle, 1)}
    try:
        import scipy.fft as sfft
        for _ in range(20):
            sfft.rfft(x)
        t0 = time.perf_counter()
        for _ in range(reps):
            sfft.rfft(x)
        res["pocketfft_rfft_us"] = round(1e6 * (time.perf_counter() - t0) / reps, 1)
    except Exception:
        pass
    return res


def real_libraries():
    """the real libzita-convolver / libfftw3f are looked for at run time.  They are in neither
    this image nor the GPU box's; the record says what was found, so that the arm cannot be mistaken for them."""
    import ctypes.util
    return {n: ctypes.util.find_library(n) for n in ("zita-convolver", "fftw3f", "sndfile")}


def cpu_baseline(wl, filter_dir, target_s=12.0):
    cores = len(os.sched_getaffinity(0))
    kind, run, what = cpu_arm(wl, filter_dir)
    audio, wall = run(cores, 8)                     # calibration (also warms the cores)
    nb = max(8, min(100000, int(8 * target_s / wall)))
    audio, wall = run(cores, nb)
    cal = fft_calibration()
    value = audio / wall
    # What the arm would reach with a pocketfft-class transform: per block and thread (ninp + nout)
    # transforms of 2 * fragm points; the measured one-core difference is taken out of the block time.
    calibrated = None
    if "pocketfft_rfft_us" in cal and cal["n"] == 2 * wl.fragm:
        t_blk = cores * (wl.fragm / wl.fs) / value
        t_fft_gain = (wl.ninp + wl.nout) * 1e-6 * max(0.0, cal["oracle_r2c_us"] - cal["pocketfft_rfft_us"])
        if t_blk > t_fft_gain:
            calibrated = value * t_blk / (t_blk - t_fft_gain)
    return {
        "value": value, "unit": "x realtime (audio-s per wall-s)", "cores": cores, "kind": kind,
        "value_with_pocketfft_class_fft": calibrated,
        "sample": f"{cores} files x {nb} blocks of {wl.fragm} frames ({wl.name}), one SoundProcessor/Convproc per "
                  f"file, one file per thread, {wall:.1f} s wall; {what}",
        "fft_calibration_one_core": cal,
        "real_libraries_found": real_libraries(),
    }


def run_reference(args, rank, world):
    if rank != 0:
        return
    import tempfile
    wl = workloads.WORKLOADS[args.workload]()
    cores = len(os.sched_getaffinity(0))
    with tempfile.TemporaryDirectory() as tmp:
        filter_dir = workloads.write_filter_dir(wl, os.path.join(tmp, wl.name))
        kind, run, what = cpu_arm(wl, filter_dir)
        audio, wall = run(cores, 8)
        # bounded sample: ~4 s of CPU work per step, and at most ~90 s for the whole run whatever K is
        per_step_s = min(4.0, 90.0 / max(1, args.steps))
        nb = max(8, min(20000, int(8 * per_step_s / wall)))
        for _ in range(args.warmup):
            run(cores, max(4, nb // 8))
        t_audio = t_wall = 0.0
        for _ in range(args.steps):
            a, w = run(cores, nb)
            t_audio += a
            t_wall += w
    v = t_audio / t_wall
    unit = "x realtime (audio-s per wall-s)"
    line = {
        "impl": "reference", "metric": "convolved audio-seconds per second", "value": v, "unit": unit,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_wall / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl.name, "fs": wl.fs, "channels": wl.ninp, "fragm": wl.fragm,
                   "files": cores, "blocks_per_step": nb},
        "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": kind,
                         "sample": f"per step: {cores} files x {nb} blocks of {wl.fragm} frames, one file per "
                                   f"thread; {what}"},
        "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# --------------------------------------------------------------------------- GPU arm
def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def emit(line):
    """The ONE JSON line of the contract goes to the real stdout; everything libraries print on
    file descriptor 1 meanwhile (NCCL's version banner at N > 1) was redirected to stderr."""
    _STDOUT.write(json.dumps(line) + "\n")
    _STDOUT.flush()


_STDOUT = sys.stdout

WIRE = {"f32": (0, 4, np.float32, 1.0), "s16": (1, 2, np.int16, 32768.0), "s24": (2, 4, np.int32, 8388608.0)}


def byte_model(flt, wl, B, T, wire_in, wire_out):
    """HBM bytes of one step of B streams x T blocks, N = fragm, 8N bytes per spectrum row.

    MAC, compulsory (time-tiled): every stream's window of P+T-1 ring slots of each input once, the
    filter rows once, the T output rows of every output once.
    MAC, traffic model: what the kernel requests -- one window per (output, input) PAIR with impulse
    data (equal to the compulsory figure for diagonal filters; ncu cross-check in profiles/).
    MAC, block-synchronous: every block re-reads its whole partition history.
    FFT kernels: PCM in / spectra out, spectra in / PCM out + the overlap tails once per step."""
    N, P, rows, I, O, K = wl.fragm, flt.ring_depth, flt.active_rows, wl.ninp, wl.nout, flt.active_pairs
    row = 8 * N
    D = P + T - 1
    return {
        "mac_compulsory": row * (B * D * I + rows + B * T * O),
        "mac_traffic_model": row * (B * D * K + rows + B * T * O),
        "mac_block_sync": 8 * (N + 1) * (B * T * P * I + rows + B * T * O),
        "fwd": B * T * I * (N * wire_in + row),
        "inv": B * T * O * (row + N * wire_out) + B * O * 2 * 4 * N,
    }


class Measure:
    """One workload on this rank's GPU: device-resident loop, end-to-end loop, copy-only loop."""

    def __init__(self, ctx, wl_name, streams, T, wire):
        from folve_b200 import capi
        self.ctx, self.capi = ctx, capi
        self.wl = workloads.WORKLOADS[wl_name]()
        self.flt = self.wl.load(capi.Filter(self.wl.ninp, self.wl.nout, self.wl.size, self.wl.fragm)).commit(ctx.local_rank)
        self.B, self.T, self.wire = streams, T, wire
        fmt, self.wire_bytes, dt, scale = WIRE[wire]
        self.batch = capi.Batch(self.flt, streams, fmt, fmt, blocks_per_step=T)
        peak_in = 0.03 if wl_name == "santalucia" else 0.25
        x = workloads.synthetic_pcm(streams, T * self.wl.fragm, self.wl.ninp, peak_in, 1000 + ctx.rank)
        self.x = x
        self.batch.host_in[:] = x if wire == "f32" else np.rint(x * scale).astype(dt)
        self.audio_per_step = ctx.world * streams * T * self.wl.fragm / self.wl.fs

    def e2e(self, steps, warm, copy_only=False):
        """-> (seconds for `steps` steps through fcv_batch_submit / fcv_batch_wait, max over ranks; launches)"""
        L, bt = self.capi.lib(), self.batch
        in1, _ = bt.slot_views(1)
        in1[:] = bt.host_in
        bt.set_copy_only(copy_only)
        for _ in range(warm):
            bt.process()
        self.ctx.barrier()
        n0 = L.fcv_kernel_launches()
        t0 = time.perf_counter()
        bt.submit(0)
        for k in range(1, steps):
            bt.submit(k & 1)
            bt.wait((k - 1) & 1)      # step k-1 is complete in its host_out slot
        bt.wait((steps - 1) & 1)
        self.ctx.sync()
        sec = self.ctx.max_over_ranks(time.perf_counter() - t0)
        nl = L.fcv_kernel_launches() - n0
        bt.set_copy_only(False)
        self.ctx.barrier()
        return sec, nl

    def device(self, steps, warm):
        """-> (ms for `steps` device-resident steps (CUDA events, max over ranks), launches, per-kernel ms per step)"""
        L, bt = self.capi.lib(), self.batch
        bt.set_profiling(not os.environ.get("FCV_DEVICE_CHUNKS"))
        for _ in range(warm):
            bt.process_device()
        bt.profile()              # drop warm-up timings
        self.ctx.barrier()
        n0 = L.fcv_kernel_launches()
        bt.event_record(0)
        for _ in range(steps):
            bt.process_device()
        bt.event_record(1)
        bt.sync()
        self.ctx.sync()
        ms = self.ctx.max_over_ranks(bt.event_elapsed_ms(0, 1))
        launches = L.fcv_kernel_launches() - n0
        kms, ksteps = bt.profile()
        bt.set_profiling(False)
        self.ctx.barrier()
        k = max(1, ksteps)
        return ms, launches, {"fwd_fft": kms[0] / k, "mac": kms[1] / k, "inv_fft": kms[2] / k}

    def roofline(self, kernel_ms, step_ms, peak, peak_src):
        bm = byte_model(self.flt, self.wl, self.B, self.T, self.wire_bytes, self.wire_bytes)
        mac_s = kernel_ms["mac"] * 1e-3
        gbs = lambda b, ms: b / (ms * 1e-3) / 1e9 if ms > 0 else float("nan")
        achieved = bm["mac_compulsory"] / mac_s / 1e9
        T = self.T
        step_bytes = bm["fwd"] + bm["mac_compulsory"] + bm["inv"]
        return {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": bm["mac_traffic_model"],
            "kernel": "mac_kernel" if T == 1 else (f"mac_tma_kernel<T={T}>" if T >= 4 and not
                                                    os.environ.get("FCV_MAC_TMA") == "0" else f"mac_tt_kernel<T={T}>"),
            "algorithmic_bytes_per_launch": bm["mac_compulsory"], "peak_source": peak_src,
            "what": "frac = compulsory HBM bytes of the time-tiled MAC launch (window of P+T-1 ring slots per stream and "
                    "input once, filter rows once, T output rows per output once) / its CUDA-event time / peak; "
                    "traffic = bytes the kernel requests (one window per (output, input) pair; equals the compulsory "
                    "bytes for diagonal filters; ncu dram__bytes cross-check: profiles/)",
            "traffic_frac_of_peak": bm["mac_traffic_model"] / mac_s / 1e9 / peak,
            "block_sync_bytes_per_launch": bm["mac_block_sync"],
            "block_sync_frac": bm["mac_block_sync"] / mac_s / 1e9 / peak,
            "per_kernel": {
                "fwd_fft": {"bytes": bm["fwd"], "gbs": gbs(bm["fwd"], kernel_ms["fwd_fft"]),
                            "frac": gbs(bm["fwd"], kernel_ms["fwd_fft"]) / peak},
                "mac": {"bytes": bm["mac_compulsory"], "gbs": achieved, "frac": achieved / peak},
                "inv_fft": {"bytes": bm["inv"], "gbs": gbs(bm["inv"], kernel_ms["inv_fft"]),
                            "frac": gbs(bm["inv"], kernel_ms["inv_fft"]) / peak},
            },
            "step": {"bytes": step_bytes, "gbs": gbs(step_bytes, step_ms), "frac": gbs(step_bytes, step_ms) / peak,
                     "what": "all kernels of a step: compulsory bytes / device time per step / peak"},
        }

    def close(self):
        self.batch.close()
        self.flt.close()


class Ctx:
    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))

    def sync(self):
        self.torch.cuda.synchronize()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def config_block(ctx, peak, peak_src, T):
    """Short device-resident + end-to-end runs of the other BASELINE configs (1, 3, 4) so that every
    config has numbers from the same run as the headline: (workload, streams per GPU, wire)."""
    plan = [("lowpass", 1024, "s16"), ("roomcorr96", 1024, "s24"), ("roomcorr192", 1024, "s24"),
            ("crossfeed", 1024, "s16"), ("surround51", 256, "s24"), ("surround51_dense", 256, "s24")]
    out = {}
    for name, streams, wire in plan:
        m = Measure(ctx, name, streams, T, wire)
        ms, _, kms = m.device(12, 3)
        sec, _ = m.e2e(6, 3)
        sec_c, _ = m.e2e(6, 2, copy_only=True)
        step_ms = ms / 12
        rf = m.roofline(kms, step_ms, peak, peak_src)
        wl, flt = m.wl, m.flt
        out[name] = {
            "workload": f"{wl.ninp}x{wl.nout} fs={wl.fs} size={wl.size} fragm={wl.fragm} ring={flt.ring_depth} "
                        f"rows={flt.active_rows} pairs={flt.active_pairs}",
            "streams_per_gpu": streams, "blocks_per_step": T, "wire_format": wire,
            "value": m.audio_per_step * 12 / (ms * 1e-3), "ms_per_step": step_ms, "kernel_ms_per_step": kms,
            "e2e": {"value": m.audio_per_step * 6 / sec, "ms_per_step": 1e3 * sec / 6,
                    "link_ceiling": m.audio_per_step * 6 / sec_c,
                    "h2d_bytes_per_step": streams * T * wl.fragm * wl.ninp * m.wire_bytes,
                    "d2h_bytes_per_step": streams * T * wl.fragm * wl.nout * m.wire_bytes},
            "mac_frac_of_hbm": rf["frac"], "mac_traffic_frac_of_hbm": rf["traffic_frac_of_peak"],
            "step_frac_of_hbm": rf["step"]["frac"],
            "per_kernel_frac": {k: v["frac"] for k, v in rf["per_kernel"].items()},
        }
        m.close()
    return out


def album_library(ctx, wl, filter_dir, T, pcm16, in_process_gpus=0):
    """BASELINE config 5 as specified: 128 albums x 8 tracks of U[120 s, 360 s], every album ONE gapless
    chain (convolve-file-handler.cc:390-415), all chains of a shard in flight at once, sharded by
    album (folve_b200/sharding.py balanced_albums) over the ranks -- or, with in_process_gpus > 1, over
    that many GPUs from ONE process (MultiDeviceConvolver).  -> dict"""
    from folve_b200 import sharding
    L = C.CDLL(HOST_SO)
    L.fh_bench_albums.restype = C.c_double
    L.fh_bench_albums.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    cfg = os.path.join(filter_dir, f"filter-{wl.fs}.conf")
    cores = len(os.sched_getaffinity(0))
    nalbums, tracks = 128, 8
    rank, world = (0, 1) if in_process_gpus > 1 else (ctx.rank, ctx.world)
    threads = max(2, cores // max(1, ctx.world if in_process_gpus <= 1 else in_process_gpus))
    audio, chains = C.c_double(0), C.c_int(0)
    ctx.barrier()
    # two BatchConvolvers per GPU (MultiDeviceConvolver's instances_per_device): four steps in flight instead of two;
    # the small per-GPU batches of this config are bound by the latency of a step, not by the link
    instances = int(os.environ.setdefault("FOLVE_B200_LIBRARY_INSTANCES", "2"))
    wall = L.fh_bench_albums(cfg.encode(), wl.fs, wl.ninp, nalbums, tracks, rank, world, in_process_gpus, T, threads,
                             1 if pcm16 else 0, C.byref(audio), C.byref(chains))
    if wall <= 0:
        raise RuntimeError("fh_bench_albums failed")
    assert chains.value == len(sharding.balanced_albums(nalbums, rank, world))
    if in_process_gpus > 1:
        total_audio, max_wall = audio.value, wall
    else:
        total_audio, max_wall = ctx.sum_over_ranks(audio.value), ctx.max_over_ranks(wall)
    return {"value": total_audio / max_wall, "unit": "x realtime (audio-s per wall-s)", "audio_seconds": total_audio,
            "wall_s": max_wall, "albums": nalbums, "tracks_per_album": tracks, "chains_per_gpu": chains.value
            if in_process_gpus <= 1 else nalbums // in_process_gpus,
            "host_threads_per_gpu": threads, "batch_convolvers_per_gpu": instances,
            "wire_format": "s16" if pcm16 else "f32", "blocks_per_step": T,
            "sharding": (f"one process, {in_process_gpus} GPUs (MultiDeviceConvolver)" if in_process_gpus > 1 else
                         f"{world} ranks, albums a = rank (mod {world}) (sharding.balanced_albums)"),
            "what": "BatchConvolver::Run over SNDFILE in / SNDFILE out: 128 gapless album chains x 8 tracks of "
                    "U[120 s, 360 s] (seed 100 + album), all chains prebuffered at t = 0, two steps in flight per BatchConvolver"}


def parity_gate(wl, flt):
    """Correctness gate reported with the numbers: one stream of the benchmark's own
    filter through the single-stream C ABI against the CPU oracle (the checker) and the float64 truth."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from folve_b200 import capi
    from oracle_py import OracleConvproc, run_blocks, truth_f64
    N = wl.fragm
    x = workloads.synthetic_pcm(1, 5 * N + 321, wl.ninp, 0.03 if wl.name == "santalucia" else 0.25, 4242)[0]
    s = capi.Stream(flt)
    y = run_blocks(s, x, N)
    s.close()
    yo = run_blocks(wl.load(OracleConvproc(wl.ninp, wl.nout, wl.size, reset_is_fresh=True)), x, N)
    h = {}
    for (i, o, d, i0) in wl.adds:
        t = h.setdefault((i, o), np.zeros(wl.size + 1, np.float64))
        t[i0:i0 + len(d)] += np.asarray(d, np.float64)[: len(t) - i0]
    for (i1, o1, i2, o2) in wl.links:
        h[(i2, o2)] = h[(i1, o1)]
    t = truth_f64(x, h, wl.nout)
    fs = max(1.0, float(np.abs(t).max()))

    def lsb_hist(a, b, scale):
        d = np.abs(np.rint(a.astype(np.float64) * scale) - np.rint(b.astype(np.float64) * scale)).astype(np.int64)
        return {str(k): int((d == k).sum()) for k in range(0, 4)} | {">=4": int((d >= 4).sum())}

    snr = lambda a: float(10 * np.log10((t ** 2).sum() / max(1e-300, ((a - t) ** 2).sum())))
    return {
        "frames": int(x.shape[0]), "max_err_engine_vs_oracle_fs": float(np.abs(y - yo).max() / fs),
        "max_err_engine_vs_truth_fs": float(np.abs(y - t).max() / fs),
        "max_err_oracle_vs_truth_fs": float(np.abs(yo - t).max() / fs),
        "snr_db_engine_vs_truth": snr(y), "snr_db_oracle_vs_truth": snr(yo),
        "lsb16_engine_vs_oracle": lsb_hist(y, yo, 32767.0), "lsb24_engine_vs_oracle": lsb_hist(y, yo, 8388607.0),
        "lsb24_engine_vs_truth": lsb_hist(y, t, 8388607.0), "lsb24_oracle_vs_truth": lsb_hist(yo, t, 8388607.0),
        "what": "one stream, 5 blocks + 321 frames of the benchmark's filter through fcv_stream_process; "
                "histograms count samples by |difference| in LSBs after 16 / 24-bit quantisation",
    }


def main():
    global _STDOUT
    sys.stdout.flush()
    _STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="santalucia", choices=sorted(workloads.WORKLOADS))
    ap.add_argument("--streams", type=int, default=1024, help="concurrent streams PER GPU")
    ap.add_argument("--wire", default="s16", choices=["f32", "s16", "s24"],
                    help="PCM format of the host and device staging buffers (the workload is 16-bit audio; "
                         "the conversions are fused into the FFT kernels)")
    ap.add_argument("--blocks-per-step", type=int, default=8, choices=[1, 2, 4, 8],
                    help="consecutive blocks of every stream per step (>1: time-tiled MAC)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling aid: only the device-resident loop")
    ap.add_argument("--no-configs", action="store_true", help="skip the short runs of the other BASELINE configs")
    ap.add_argument("--no-library", action="store_true", help="skip the config-5 album library leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.warmup < 3:
        args.warmup = 3

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    # the C++ host layer (SoundProcessor, BatchConvolver) of this rank works on this rank's GPU
    os.environ.setdefault("FOLVE_B200_DEVICE", str(local_rank))
    ctx = Ctx()
    from folve_b200 import capi
    L = capi.lib()
    K, W, T, B = args.steps, args.warmup, args.blocks_per_step, args.streams
    peak, peak_src = measured_peak()
    small = B < 256   # contract test: keep the auxiliary legs out

    m = Measure(ctx, args.workload, B, T, args.wire)
    wl, flt, N = m.wl, m.flt, m.wl.fragm
    clocks = ClockSampler(local_rank)
    clocks.start()

    # ---- end to end through the C ABI: pinned host in -> pinned host out, every step.
    # Two host staging slots: step k+1 is submitted before step k is awaited, the way a prebuffering
    # server keeps the copy engines busy; every step still moves its own input host->device and its
    # own output device->host.  link_ceiling: the same loop with the kernels switched off.
    e2e_s = launches_e2e = ceil_s = None
    if not args.skip_e2e:
        e2e_s, launches_e2e = m.e2e(K, W)
        ceil_s, _ = m.e2e(max(5, K // 2), 2, copy_only=True)

    # ---- single-stream block latency through the synchronous drop-in call
    lat = None
    if not args.skip_e2e and rank == 0:
        st = capi.Stream(flt)
        st.buffer[: N * wl.ninp] = m.x[0, :N].reshape(-1)
        mx = C.c_float(0)
        ts = []
        for k in range(1100):
            t1 = time.perf_counter()
            L.fcv_stream_process(st._h, N, C.byref(mx))
            ts.append(time.perf_counter() - t1)
        ts = np.array(ts[100:]) * 1e6
        lat = {"median": float(np.median(ts)), "p99": float(np.percentile(ts, 99)), "blocks": int(ts.size),
               "what": "fcv_stream_process, one stream: pinned block read by the forward kernel over the link, "
                       "3 kernels, output written to the pinned block by the inverse kernel"}
        st.close()
        # the same filter with non-uniform partitions (fcv_nustream: head partitions of 1024 frames, tail
        # partitions of fragm): a block of 1024 frames = 23 ms of audio instead of 186 ms
        if N == 8192:
            q = 1024
            nuf = wl.load(capi.NuFilter(wl.ninp, wl.nout, wl.size, q, N)).commit(local_rank)
            nus = capi.NuStream(nuf)
            nus.buffer[: q * wl.ninp] = m.x[0, :q].reshape(-1)
            ts = []
            for k in range(8 * 300):
                t1 = time.perf_counter()
                L.fcv_nustream_process(nus._h, q, C.byref(mx))
                ts.append(time.perf_counter() - t1)
            ts = np.array(ts[8 * 50:]) * 1e6
            per_big = ts.reshape(-1, 8).sum(axis=1)
            lat["nonuniform_q1024"] = {
                "median": float(np.median(ts)), "p99": float(np.percentile(ts, 99)),
                "median_of_tail_calls": float(np.median(ts.reshape(-1, 8)[:, 7])),
                "us_per_8192_frames": float(np.median(per_big)), "calls": int(ts.size),
                "what": "fcv_nustream_process, one stream, blocks of 1024 frames: 8 head partitions of 1024 + tail "
                        "partitions of 8192; every 8th call also evaluates the tail level for the next large block"}
            nus.close()
            nuf.close()
    ctx.barrier()

    # ---- device resident: PCM already in HBM (left there by the steps above)
    dev_ms, launches, kms = m.device(K, W)
    clk = clocks.stop()

    value = m.audio_per_step * K / (dev_ms * 1e-3)
    # the same end-to-end loop with float32 on the wire (what SoundProcessor's float buffer
    # would ship unconverted): twice the PCIe bytes
    e2e_f32 = None
    if not args.skip_e2e and args.wire != "f32" and not small:
        K2 = max(10, K // 4)
        m32 = Measure(ctx, args.workload, B, T, "f32")
        sec, _ = m32.e2e(K2, W)
        e2e_f32 = {"value": m32.audio_per_step * K2 / sec, "ms_per_step": 1e3 * sec / K2, "steps": K2,
                   "h2d_bytes_per_step": B * T * N * wl.ninp * 4, "d2h_bytes_per_step": B * T * N * wl.nout * 4}
        m32.close()
    e2e_value = None if args.skip_e2e else m.audio_per_step * K / e2e_s
    link_ceiling = None
    if ceil_s:
        kc = max(5, K // 2)
        cv = m.audio_per_step * kc / ceil_s
        link_ceiling = {"value": cv, "ms_per_step": 1e3 * ceil_s / kc, "steps": kc, "e2e_frac_of_ceiling": e2e_value / cv,
                        "what": "the same submit/wait loop with the kernels switched off (fcv_batch_set_copy_only): "
                                "only the host->device and device->host copies of every step"}

    rf = m.roofline(kms, dev_ms / K, peak, peak_src)
    P, rows = flt.ring_depth, flt.active_rows
    configs = None
    if world == 1 and not args.no_configs and not args.skip_e2e and not small:
        configs = config_block(ctx, peak, peak_src, T)

    library = library_1p = None
    if not args.no_library and not args.skip_e2e and not small and os.path.exists(HOST_SO):
        import tempfile
        with tempfile.TemporaryDirectory() as tmp:
            filter_dir = workloads.write_filter_dir(wl, os.path.join(tmp, wl.name))
            library = album_library(ctx, wl, filter_dir, T, args.wire == "s16")
            ngpu_box = L.fcv_device_count()
            if world == 1 and ngpu_box > 1:
                # every GPU of the box from this ONE process, the way folve (a single process) would
                library_1p = album_library(ctx, wl, filter_dir, T, args.wire == "s16", in_process_gpus=ngpu_box)

    if rank == 0:
        wb = m.wire_bytes
        line = {
            "metric": "convolved audio-seconds per second", "value": value,
            "unit": "x realtime (audio-s per wall-s)", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": f"{wl.name}: {wl.ninp}x{wl.nout} fs={wl.fs} size={wl.size} fragm={N} "
                            f"partitions={flt.partitions} (non-zero ring depth {P}, {rows} filter rows)",
                "streams_per_gpu": B, "blocks_per_step": T, "frames_per_block": N,
                "audio_seconds_per_step": m.audio_per_step, "wire_format": args.wire,
                "l2": f"per-step working set {rf['step']['bytes'] / 1e9:.2f} GB >> 126 MB L2 (inputs larger than L2)",
                "parallelism": f"{world} x independent stream shards, no collective",
            },
            "e2e": {"value": e2e_value, "unit": "x realtime (audio-s per wall-s)",
                    "h2d_bytes_per_step": B * T * N * wl.ninp * wb, "d2h_bytes_per_step": B * T * N * wl.nout * wb,
                    "ms_per_step": None if args.skip_e2e else 1e3 * e2e_s / K, "wire_format": args.wire,
                    "link_ceiling": link_ceiling, "f32_wire": e2e_f32},
            "gpu_launches": int(launches), "gpu_launches_e2e": None if launches_e2e is None else int(launches_e2e),
            "kernel_ms_per_step": kms,
            "roofline": rf,
            "clocks": clk,
            "block_latency_us": lat,
        }
        if configs is not None:
            line["configs"] = configs
        if library is not None:
            line["e2e"]["album_library"] = library
        if library_1p is not None:
            line["e2e"]["album_library_one_process"] = library_1p
        if world == 1 and not args.no_cpu_baseline:
            import tempfile
            with tempfile.TemporaryDirectory() as tmp:
                filter_dir = workloads.write_filter_dir(wl, os.path.join(tmp, wl.name))
                line["cpu_baseline"] = cpu_baseline(wl, filter_dir)
                line["cpu_baseline"]["parity"] = parity_gate(wl, flt)
                if os.path.exists(HOST_SO) and not args.skip_e2e:
                    # the drop-in API itself: one synchronous SoundProcessor per host thread on the GPU
                    cores = len(os.sched_getaffinity(0))
                    harness_run(HOST_SO, wl, filter_dir, cores, 20)
                    a, w = harness_run(HOST_SO, wl, filter_dir, cores, 400)
                    a2x, w2x = harness_run(HOST_SO, wl, filter_dir, 2 * cores, 400)
                    a2, w2 = library_run(wl, filter_dir, B, 2, 120.0, T, cores, args.wire == "s16")
                    a3, w3 = library_run(wl, filter_dir, B, 2, 120.0, T, cores, False)
                    line["e2e"]["batch_convolver"] = {
                        "value": a2 / w2, "unit": "x realtime (audio-s per wall-s)", "threads": cores,
                        "wire_format": args.wire, "f32_wire_value": a3 / w3,
                        "what": f"BatchConvolver::Run: {B} gapless chains x 2 in-memory files of ~120 s in flight at once, "
                                f"{T} blocks per chain and step, SNDFILE in / SNDFILE out on {cores} host threads, two "
                                f"steps in flight"}
                    line["e2e"]["soundprocessor_sync"] = {
                        "value": a / w, "unit": "x realtime (audio-s per wall-s)", "threads": cores,
                        "value_2x_threads": a2x / w2x, "threads_2x": 2 * cores,
                        "what": "SoundProcessor::FillBuffer/WriteProcessed block loop, one file per host thread (as many "
                                "threads as cores, and twice as many: the callers mostly wait); the library's dispatcher "
                                "gathers the concurrent synchronous calls into shared launch groups"}
        emit(line)

    m.close()
    ctx.close()


if __name__ == "__main__":
    main()
