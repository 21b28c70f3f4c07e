// fcv_internal.h -- what the translation units of libfolve_b200.so share: the handle
// structs behind include/folve_b200.h, error helpers, and the launch interface between the
// host side (fcv_engine.cu) and the kernel files (fcv_k_*.cu).  The library is split so that the
// kernels compile in parallel and a change to one kernel family rebuilds one file.
#pragma once
#include "../../include/folve_b200.h"

#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <vector>

#include "fcv_stream_dev.cuh"
#include "fcv_types.h"

namespace fcv {

// ---- errors -----------------------------------------------------------------------------
int fail(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));
extern std::atomic<unsigned long long> g_launches;

#define CU_TRY(expr)                                                                              \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess)                                                                   \
            return fcv::fail(FCV_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                             __FILE__, __LINE__);                                                 \
    } while (0)

}  // namespace fcv

// ---- handles --------------------------------------------------------------------------------
struct FcvPair {
    bool exists = false;      // a MAC node was created for this pair
    int link = -1;            // index of the pair whose spectra are used instead
    std::vector<float> h;     // time domain, npar * fragm, already scaled by 0.5/fragm
    std::vector<int> row;     // per partition: filter row or -1 (after commit)
};

struct FcvCombiner;

struct fcv_filter {
    std::atomic<int> refs{1};
    int ninp = 0, nout = 0;
    unsigned size = 0;
    int fragm = 0, log2n = 0;
    int npar = 0;  // partitions zita allocates room for
    bool committed = false;
    std::vector<FcvPair> pairs;  // [inp * nout + out]
    // after commit
    int device = -1;
    int ring = 1;       // depth of the input-spectra ring
    int nrows = 0;      // non-zero (pair, partition) spectra
    int active_pairs = 0;
    int group_no = 1;   // outputs per MAC group
    int ngroups = 1;
    int nsteps = 0;
    float2 *dH = nullptr;
    fcv::MacStep *dsteps = nullptr;
    int *dgroup_off = nullptr;
    // per-output pair lists (time-tiled MAC, DC/Nyquist products)
    fcv::TTPair *dpairs = nullptr;
    int *dpair_off = nullptr;
    int *dtt_rows = nullptr;
    fcv::FftTables tb{};
    fcv::f13::Tables tb13{};   // fragm = 8192 only
    bool k13 = false;          // fragm = 8192 transforms of fcv_fft13.cuh in use
    std::vector<fcv::MacStep> hsteps;
    std::vector<int> hgroup_off;
    FcvCombiner *combiner = nullptr;   // coalesces concurrent single-stream calls (created at commit)
};

struct fcv_batch {
    fcv_filter *f = nullptr;
    int B = 0;
    int T = 1;   // blocks per stream per step
    int R = 1;   // ring depth = filter ring + T - 1
    int in_fmt = FCV_PCM_F32, out_fmt = FCV_PCM_F32;
    size_t in_block = 0, out_block = 0;  // bytes per stream per STEP (T blocks)
    size_t out_pad = 0;                  // extra bytes after device_out (single-stream max mirror)
    unsigned long long step = 0;         // blocks processed so far (ring slot = step % ring)
    bool per_block_max = false;          // single-stream mode: maxv is the maximum of the last block only
    int num_sms = 148;                   // SMs of the device (persistent grids)
    bool in_zero_copy = false;           // single-stream mode: kernels read the PCM block from pinned host memory
    const void *hin_dev = nullptr;       // device address of hin in that case
    size_t host_block = 0;               // single-stream mode: bytes of the shared in/out host block (max mirror behind it)
    bool copy_only = false;              // diagnostic: submit moves the PCM but launches no kernel
    unsigned seq_src = 0;                // single-stream staged mode: source of the completion word's host-to-host copy
    int last_path = 0;                   // which entry point enqueued last (device loop / submit): see enter_path
    // device
    unsigned char *dmem = nullptr;       // one slab
    float2 *xring = nullptr;
    float *tail = nullptr;
    unsigned char *din = nullptr, *dout = nullptr;
    float2 *Y = nullptr;
    float *maxv = nullptr;
    float *bmax = nullptr;               // [B][T] per-block maxima of the last step
    float2 *zc0 = nullptr;               // [B][nout][T] entry 0 of the sequences to inverse-transform (dcny_kernel)
    fcv::StreamDev *dst = nullptr;
    int *dfv = nullptr;
    size_t state_bytes_per_stream = 0;
    // host
    unsigned char *hin = nullptr, *hout = nullptr;
    int *hfv = nullptr;
    // second host staging slot + per-slot completion events for the asynchronous submit/wait pair
    unsigned char *hin1 = nullptr, *hout1 = nullptr;
    int *dfv1 = nullptr, *hfv1 = nullptr;
    float *hbmax[2] = {nullptr, nullptr};   // [B][T] block maxima of the step submitted from each host slot
    cudaEvent_t slot_done[2][8] = {};
    bool slot_busy[2] = {false, false};
    // streams
#ifndef FCV_BATCH_NQ
#define FCV_BATCH_NQ 4
#endif
    static const int NQ = FCV_BATCH_NQ;   // CUDA streams the chunks of a submit rotate over (<= 8)
    cudaStream_t q[NQ] = {};
    cudaEvent_t fj[9] = {};  // fork/join events of the chunked device path
    // stopwatch
    cudaEvent_t sw[16] = {};
    // profiling
    bool profiling = false;
    std::vector<cudaEvent_t> ev;  // 4 events per step: t0 | fwd | mac | inv
    size_t ev_used = 0;
    int prof_steps = 0;
};

namespace fcv {

// ---- launch interface: one step of `cnt` streams on CUDA stream q ---------------------------
struct StepArgs {
    const fcv_filter *f = nullptr;
    int T = 1, R = 1;
    int in_fmt = PCM_F32, out_fmt = PCM_F32;
    bool pdl = false;            // programmatic dependent launch between the kernels (single-stream path)
    bool per_block_max = false;  // the forward kernel zeroes the stream's running maximum
    int num_sms = 148;
    int cnt = 0;
    // batch addressing ...
    BatchSel bsel{};
    // ... or a coalesced group of single streams (cnt <= GROUP_MAX), T == 1
    const GroupSel *grp = nullptr;
    // batch only, T > 1: the step's Y rows and entry-0 slots of stream 0 of the launch
    float2 *Y = nullptr;
    float2 *zc0 = nullptr;
};

// fcv_k_fft.cu (any block size) / fcv_k_fft13.cu (fragm = 8192)
int fft_tables(int device, int log2n, FftTables *out);
int fft13_tables(int device, f13::Tables *out);
int fft_entry_of_bin(int log2n, int k);   // entry of bin k in the generic kernels' spectrum layout
void launch_fwd(const StepArgs &a, cudaStream_t q);
void launch_inv(const StepArgs &a, cudaStream_t q);
void launch_fwd13(const StepArgs &a, cudaStream_t q);
void launch_inv13(const StepArgs &a, cudaStream_t q);
// filter preparation: src[row][N] floats -> dst[row][N] spectra on the current device (synchronous to stream 0)
int launch_filter_fft(const fcv_filter *f, const float *dsrc, float2 *dst, int nrows);
void launch_filter_fft13(const fcv_filter *f, const float *dsrc, float2 *dst, int nrows);
// fcv_k_mac.cu / fcv_k_mac_tma.cu
void launch_mac_t1(const StepArgs &a, cudaStream_t q);                  // T == 1 (DC / Nyquist inside)
void launch_mac_tt(const StepArgs &a, int newest, cudaStream_t q);      // T = 2, 4, 8
bool launch_mac_tma(const StepArgs &a, int newest, cudaStream_t q);     // T = 4, 8; false: shape not covered
void launch_dcny(const StepArgs &a, cudaStream_t q);                    // T > 1

// fcv_engine.cu, for fcv_nonuniform.cu
int stream_set_mix(fcv_stream *s, const float *device_frames);
const float *stream_device_out(const fcv_stream *s);

// fcv_k_fused13.cu: the three kernels of a single-stream group as one cooperative launch
bool fused13_available(const fcv_filter *f, int in_fmt, int out_fmt);
bool launch_fused13(const StepArgs &a, cudaStream_t q);   // false: not launched, take the three-launch path

// <<<>>> with the programmatic-stream-serialization attribute when `pdl` (single-stream path:
// the next kernel's launch latency hides behind the tail of the previous one)
template <typename... KArgs, typename... Args>
static inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t q, bool pdl,
                            Args... args) {
    if (!pdl) {
        kernel<<<grid, block, smem, q>>>(KArgs(args)...);
        return;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = q;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

}  // namespace fcv
