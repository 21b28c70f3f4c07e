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


def cpu_baseline(wl, filter_dir, target_s=12.0):
    cores = len(os.sched_getaffinity(0))
    kind, run, what = cpu_arm(wl, filter_dir)
    audio, wall = run(cores, 8)                     # calibration (also warms the cores)
    nb = max(8, min(100000, int(8 * target_s / wall)))
    audio, wall = run(cores, nb)
    return {
        "value": audio / wall, "unit": "x realtime (audio-s per wall-s)", "cores": cores, "kind": kind,
        "sample": f"{cores} files x {nb} blocks of {wl.fragm} frames ({wl.name}), one SoundProcessor/Convproc per "
                  f"file, one file per thread, {wall:.1f} s wall; {what}",
        "fft_calibration_one_core": fft_calibration(),
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


def ncu_traffic(wl_name, streams, T=1):
    """dram bytes per MAC launch from the committed ncu --set full summary, if one matches."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "mac_traffic.json")))
        e = d.get(f"{wl_name}:{streams}" if T == 1 else f"{wl_name}:{streams}:T{T}")
        return float(e["dram_bytes_per_launch"]) if e else None
    except Exception:
        return None


def emit(line):
    """The ONE JSON line of the contract goes to the real stdout; everything libraries print on
    file descriptor 1 meanwhile (NCCL's version banner at N > 1) was redirected to stderr."""
    _STDOUT.write(json.dumps(line) + "\n")
    _STDOUT.flush()


_STDOUT = sys.stdout


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
    ap.add_argument("--wire", default="s16", choices=["f32", "s16"],
                    help="PCM format of the host and device staging buffers (the workload is 16-bit audio; "
                         "the conversions are fused into the FFT kernels)")
    ap.add_argument("--blocks-per-step", type=int, default=8, choices=[1, 2, 4, 8],
                    help="consecutive blocks of every stream per step (>1: time-tiled MAC)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling aid: only the device-resident loop")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.warmup < 3:
        args.warmup = 3

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from folve_b200 import capi

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    wl = workloads.WORKLOADS[args.workload]()
    flt = wl.load(capi.Filter(wl.ninp, wl.nout, wl.size, wl.fragm)).commit(local_rank)
    B, N, K, W, T = args.streams, wl.fragm, args.steps, args.warmup, args.blocks_per_step
    fmt = capi.PCM_F32 if args.wire == "f32" else capi.PCM_S16
    batch = capi.Batch(flt, B, fmt, fmt, blocks_per_step=T)
    x = workloads.synthetic_pcm(B, T * N, wl.ninp, 0.03, 1000 + rank)
    batch.host_in[:] = x if fmt == capi.PCM_F32 else np.rint(x * 32768.0).astype(np.int16)
    L = capi.lib()

    clocks = ClockSampler(local_rank)

    # ---- end to end through the C ABI: pinned host in -> pinned host out, every step.
    # Two host staging slots: block k+1 is submitted before block k is awaited, the
    # way a prebuffering server keeps the copy engines busy; every step still moves
    # its own input host->device and its own output device->host.
    def measure_e2e(bt, steps):
        """-> (seconds for `steps` steps, max over ranks; kernel launches)"""
        in1, _ = bt.slot_views(1)
        in1[:] = bt.host_in
        for _ in range(1 if args.skip_e2e else W):
            bt.process()
        barrier()
        n_0 = L.fcv_kernel_launches()
        t_0 = time.perf_counter()
        if not args.skip_e2e:
            bt.submit(0)
            for k in range(1, steps):
                bt.submit(k & 1)
                bt.wait((k - 1) & 1)      # block k-1 is complete in its host_out slot
            bt.wait((steps - 1) & 1)
        torch.cuda.synchronize()
        sec = max_over_ranks(time.perf_counter() - t_0)
        nl = L.fcv_kernel_launches() - n_0
        barrier()
        return sec, nl

    clocks.start()
    e2e_s, launches_e2e = measure_e2e(batch, K)

    # ---- single-stream block latency through the synchronous drop-in call
    lat = None
    if not args.skip_e2e and rank == 0:
        st = capi.Stream(flt)
        st.buffer[: N * wl.ninp] = x[0, :N].reshape(-1)
        m = C.c_float(0)
        ts = []
        for k in range(1100):
            t1 = time.perf_counter()
            L.fcv_stream_process(st._h, N, C.byref(m))
            ts.append(time.perf_counter() - t1)
        ts = np.array(ts[100:]) * 1e6
        lat = {"median": float(np.median(ts)), "p99": float(np.percentile(ts, 99)), "blocks": int(ts.size),
               "what": "fcv_stream_process, one stream: pinned block read by the forward kernel over the link, "
                       "3 kernels, D2H -> pinned block"}
        st.close()
    barrier()

    # ---- device resident: PCM already in HBM (left there by the steps above)
    batch.set_profiling(not os.environ.get("FCV_DEVICE_CHUNKS"))
    for _ in range(W):
        batch.process_device()
    batch.profile()              # drop warm-up timings
    barrier()
    n0 = L.fcv_kernel_launches()
    batch.event_record(0)
    for _ in range(K):
        batch.process_device()
    batch.event_record(1)
    batch.sync()
    torch.cuda.synchronize()
    dev_ms = max_over_ranks(batch.event_elapsed_ms(0, 1))
    launches = L.fcv_kernel_launches() - n0
    kms, ksteps = batch.profile()
    clk = clocks.stop()
    barrier()

    audio_per_step = world * B * T * N / wl.fs
    value = audio_per_step * K / (dev_ms * 1e-3)
    # the same end-to-end loop with float32 on the wire (what SoundProcessor's float buffer
    # would ship unconverted): twice the PCIe bytes
    e2e_f32 = None
    if not args.skip_e2e and fmt != capi.PCM_F32:
        K2 = max(10, K // 4)
        b32 = capi.Batch(flt, B, capi.PCM_F32, capi.PCM_F32, blocks_per_step=T)
        b32.host_in[:] = x
        sec, _ = measure_e2e(b32, K2)
        e2e_f32 = {"value": audio_per_step * K2 / sec, "ms_per_step": 1e3 * sec / K2, "steps": K2,
                   "h2d_bytes_per_step": B * T * N * wl.ninp * 4, "d2h_bytes_per_step": B * T * N * wl.nout * 4}
        b32.close()
    e2e_value = None if args.skip_e2e else audio_per_step * K / e2e_s

    # roofline of the complex-MAC kernel: SURVEY section 8(d) algorithmic bytes
    P, rows, I, O = flt.ring_depth, flt.active_rows, wl.ninp, wl.nout
    # block-synchronous streaming model: every one of the B*T stream-blocks of a launch
    # reads its full partition history (a time-tiled launch moves fewer bytes: see traffic)
    bytes_mac = 8 * (N + 1) * (B * T * P * I + rows + B * T * O)
    mac_ms = kms[1] / max(1, ksteps) if ksteps else float("nan")
    peak, peak_src = measured_peak()
    achieved = bytes_mac / (mac_ms * 1e-3) / 1e9
    traffic = ncu_traffic(wl.name, B, T)

    if rank == 0:
        wire_bytes = 4 if fmt == capi.PCM_F32 else 2
        line = {
            "metric": "convolved audio-seconds per second", "value": value,
            "unit": "x realtime (audio-s per wall-s)", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": f"{wl.name}: {wl.ninp}x{wl.nout} fs={wl.fs} size={wl.size} fragm={N} "
                            f"partitions={flt.partitions} (non-zero ring depth {P}, {rows} filter rows)",
                "streams_per_gpu": B, "blocks_per_step": T, "frames_per_block": N,
                "audio_seconds_per_step": audio_per_step, "wire_format": args.wire,
                "l2": f"per-step working set {(bytes_mac + 0) / 1e9:.2f} GB >> 126 MB L2 (inputs larger than L2)",
                "parallelism": f"{world} x independent stream shards, no collective",
            },
            "e2e": {"value": e2e_value, "unit": "x realtime (audio-s per wall-s)",
                    "h2d_bytes_per_step": B * T * N * I * wire_bytes, "d2h_bytes_per_step": B * T * N * O * wire_bytes,
                    "ms_per_step": 1e3 * e2e_s / K, "wire_format": args.wire, "f32_wire": e2e_f32},
            "gpu_launches": int(launches), "gpu_launches_e2e": int(launches_e2e),
            "kernel_ms_per_step": {"fwd_fft": kms[0] / max(1, ksteps), "mac": mac_ms,
                                   "inv_fft": kms[2] / max(1, ksteps)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "kernel": "mac_kernel" if T == 1 else (f"mac_tma_kernel<T={T}>" if T >= 4 and not
                                                                os.environ.get("FCV_MAC_TMA") == "0" else f"mac_tt_kernel<T={T}>"),
                         "algorithmic_bytes_per_launch": bytes_mac, "peak_source": peak_src,
                         # what the kernel really moved (ncu dram bytes) over the same measured time:
                         # for a time-tiled launch this, not `frac`, is the fraction of the HBM peak in use
                         "traffic_gbs": (traffic / (mac_ms * 1e-3) / 1e9) if traffic else None,
                         "traffic_frac_of_peak": (traffic / (mac_ms * 1e-3) / 1e9 / peak) if traffic else None},
            "clocks": clk,
            "block_latency_us": lat,
        }
        if world == 1 and not args.no_cpu_baseline:
            import tempfile
            with tempfile.TemporaryDirectory() as tmp:
                filter_dir = workloads.write_filter_dir(wl, os.path.join(tmp, wl.name))
                line["cpu_baseline"] = cpu_baseline(wl, filter_dir)
                if os.path.exists(HOST_SO) and not args.skip_e2e:
                    # the drop-in API itself: one synchronous SoundProcessor per host thread on the GPU
                    cores = len(os.sched_getaffinity(0))
                    harness_run(HOST_SO, wl, filter_dir, cores, 20)
                    a, w = harness_run(HOST_SO, wl, filter_dir, cores, 400)
                    a2, w2 = library_run(wl, filter_dir, B, 2, 120.0, T, cores, fmt == capi.PCM_S16)
                    a3, w3 = library_run(wl, filter_dir, B, 2, 120.0, T, cores, False)
                    line["e2e"]["batch_convolver"] = {
                        "value": a2 / w2, "unit": "x realtime (audio-s per wall-s)", "threads": cores,
                        "wire_format": args.wire, "f32_wire_value": a3 / w3,
                        "what": f"BatchConvolver::Run: {B} gapless chains x 2 in-memory files of ~120 s in flight at once, "
                                f"{T} blocks per chain and step, SNDFILE in / SNDFILE out on {cores} host threads, two "
                                f"steps in flight"}
                    line["e2e"]["soundprocessor_sync"] = {
                        "value": a / w, "unit": "x realtime (audio-s per wall-s)", "threads": cores,
                        "what": "SoundProcessor::FillBuffer/WriteProcessed block loop, one file per host thread, "
                                "one synchronous fcv_stream_process per block"}
        emit(line)

    batch.close()
    flt.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
