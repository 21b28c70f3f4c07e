// fcv_engine.cu -- C ABI of the B200 convolution engine (include/folve_b200.h):
// filter spectra resident in HBM, per-stream device state, the three kernel
// launches per block, pinned staging and the batched entry points.
//
// Replaces the Convproc object behind folve's SoundProcessor
// (/root/reference/sound-processor.cc:34-145) and the impulse-loading calls of
// the zita-config loader (/root/reference/zita-config.cc:163,203,252,274,
// /root/reference/zita-fconfig.cc:78-93).  No CPU fallback: without a usable
// sm_100 device every entry point fails with FCV_E_CUDA.
#include "../../include/folve_b200.h"

#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "fcv_fft.cuh"
#include "fcv_fft13.cuh"
#include "fcv_mac.cuh"
#include "fcv_mac_tma.cuh"

using namespace fcv;

#ifndef FFT_MIN_CTAS
#define FFT_MIN_CTAS 2
#endif
#ifndef FWD_MIN_CTAS
#define FWD_MIN_CTAS 4
#endif

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<unsigned long long> g_launches{0};

static int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU_TRY(expr)                                                                         \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess)                                                              \
            return fail(FCV_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                                 \
    } while (0)

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------

// Signed maximum (>= 0) of one block over all output channels: warp reduction, then one
// atomic per warp (positive floats order like their bit patterns).
__device__ __forceinline__ void block_max_update(float *dst, float m) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(reinterpret_cast<int *>(dst), __float_as_int(m));
}

// Forward transform of the current block of every (stream, input channel):
// fused int/float conversion + de-interleave + zero padding + real FFT, written
// into ring slot `pt` of the stream's input-spectra ring.
template <int LOG2N>
__global__ void __launch_bounds__(fft_threads(LOG2N, 1), FWD_MIN_CTAS)
fwd_stream_kernel(const StreamDev *__restrict__ st, const int *__restrict__ fv, int fv_all, FftTables tb,
                  int ninp, int R, int T, int pt, int in_fmt, int reset_max) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *sm = reinterpret_cast<float2 *>(smem_raw);
    constexpr int N = 1 << LOG2N, NT = fft_threads(LOG2N, 1);
    // grid: x = 2 * input channel + half, y = stream, z = block of the step
    const int i = blockIdx.x >> 1, h = blockIdx.x & 1, b = blockIdx.y, bt = blockIdx.z;
    const StreamDev s = st[b];
    int frames = (fv ? fv[b] : fv_all) - bt * N;
    frames = frames < 0 ? 0 : (frames > N ? N : frames);
    int slot = pt + bt;
    if (slot >= R) slot -= R;
    float2 *row = s.xring + (size_t)(i * R + slot) * N;
    // per-block maximum mode: the inverse kernel of this block starts from zero
    if (reset_max && blockIdx.x == 0 && bt == 0 && threadIdx.x == 0) *s.maxv = 0.0f;
    if (blockIdx.x == 0 && threadIdx.x == 0) s.bmax[bt] = 0.0f;  // this block's maximum starts from zero
    if (frames == 0) {  // silence: its spectrum is zero
        for (int e = threadIdx.x; e < N / 2; e += NT) row[h * (N / 2) + e] = make_float2(0.f, 0.f);
        return;
    }
    const size_t boff = (size_t)bt * N * ninp;  // samples before this block in the staging area
    if (in_fmt == PCM_F32) fwd_body<LOG2N, PCM_F32, 1>(sm, tb, (const float *)s.din + boff, ninp, i, frames, row, h);
    else if (in_fmt == PCM_S16) fwd_body<LOG2N, PCM_S16, 1>(sm, tb, (const short *)s.din + boff, ninp, i, frames, row, h);
    else fwd_body<LOG2N, PCM_S24, 1>(sm, tb, (const int *)s.din + boff, ninp, i, frames, row, h);
}

// Forward transform of raw float partitions (filter preparation, K6):
// src[row][N] -> dst[row][M].
template <int LOG2N>
__global__ void __launch_bounds__(fft_threads(LOG2N), FFT_MIN_CTAS)
fwd_raw_kernel(const float *__restrict__ src, float2 *__restrict__ dst, FftTables tb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *sm = reinterpret_cast<float2 *>(smem_raw);
    constexpr int N = 1 << LOG2N;
    const size_t r = blockIdx.x;
    fwd_body<LOG2N, PCM_F32, 2>(sm, tb, src + r * N, 1, 0, N, dst + r * N);
}

// Overlap-add, tail save, re-interleave, float/int conversion and signed maximum
// of one output channel; returns this thread's maximum over the valid frames.
template <int LOG2N, int FMT>
__device__ __forceinline__ float inv_epilogue(const float2 *sm, const FftTables &tb, float2 *__restrict__ tail,
                                              void *dout, int nout, int o, int frames) {
    constexpr int N = 1 << LOG2N, Q = N / 2;
    constexpr int NT = fft_threads(LOG2N);
    constexpr int CH = (Q / NT) < 8 ? (Q / NT) : 8;
    const int tid = threadIdx.x;
    float lmax = 0.0f;
#pragma unroll 1
    for (int c = 0; c < Q / NT; c += CH) {
        float2 w[CH], tl[CH];
#pragma unroll
        for (int i = 0; i < CH; i++) {
            const int n = tid + (c + i) * NT;
            w[i] = __ldg(&tb.twA[n]);
            tl[i] = tail[n];
        }
#pragma unroll
        for (int i = 0; i < CH; i++) {
            const int n = tid + (c + i) * NT;
            const float2 a = sm[smem_pad(n)];
            const float2 t = cmulconj(sm[smem_pad(Q + n)], w[i]);
            const float y0 = a.x + t.x + tl[i].x;
            const float y1 = a.y + t.y + tl[i].y;
            tail[n] = make_float2(a.x - t.x, a.y - t.y);
            const int f0 = 2 * n;
            pcm_store<FMT>(dout, (size_t)f0 * nout + o, y0);
            pcm_store<FMT>(dout, (size_t)(f0 + 1) * nout + o, y1);
            if (f0 < frames) lmax = fmaxf(lmax, y0);
            if (f0 + 1 < frames) lmax = fmaxf(lmax, y1);
        }
    }
    return lmax;
}

// DC / Nyquist products (dcny_warp, fcv_mac.cuh) of a multi-block step: one warp per (stream,
// output, block of the step).  Runs between the MAC and the inverse transform, off their
// critical paths; the block-by-block MAC kernel does the same inside its own launch.
__global__ void __launch_bounds__(256)
dcny_kernel(const StreamDev *__restrict__ st, const TTPair *__restrict__ pairs, const int *__restrict__ pair_off,
            const int *__restrict__ tt_rows, const float2 *__restrict__ H, float2 *__restrict__ zc0, int nwarps,
            int nout, int P, int R, int T, int pt, int M) {
    const int w = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= nwarps) return;
    const int bt = w % T, o = (w / T) % nout, b = w / (T * nout);
    int newest = pt + bt;
    if (newest >= R) newest -= R;
    const float2 z = dcny_warp(st[b].xring, pairs, pair_off, tt_rows, H, o, P, R, newest, M, lane);
    if (lane == 0) zc0[w] = z;
}

// Inverse transform of every (stream, output channel) with fused DC/Nyquist
// products, overlap-add, tail save, re-interleave, float/int conversion and
// running signed maximum.  The T blocks of a step are done one after the other
// by the same CTA (block t+1 overlap-adds the tail block t just saved).
template <int LOG2N>
__global__ void __launch_bounds__(fft_threads(LOG2N), FFT_MIN_CTAS)
inv_stream_kernel(const StreamDev *__restrict__ st, const int *__restrict__ fv, int fv_all, FftTables tb,
                  const float2 *__restrict__ Y, const float2 *__restrict__ zc0, int nout, int T, int out_fmt) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *sm = reinterpret_cast<float2 *>(smem_raw);
    __shared__ float red[32];
    constexpr int N = 1 << LOG2N, M = N;
    constexpr int NT = fft_threads(LOG2N);
    const int tid = threadIdx.x;
    const int o = blockIdx.x, b = blockIdx.y;
    const StreamDev s = st[b];
    const int fvb = fv ? fv[b] : fv_all;
    float2 *tail = reinterpret_cast<float2 *>(s.tail + (size_t)o * N);
    float lmax = 0.0f;

    for (int bt = 0; bt < T; bt++) {
        int frames = fvb - bt * N;
        frames = frames < 0 ? 0 : (frames > N ? N : frames);
        inv_load<LOG2N>(sm, tb, Y + (((size_t)b * nout + o) * T + bt) * M);
        if (tid == 0) sm[0] = zc0[((size_t)b * nout + o) * T + bt];  // Zc[0] from the two real bins (dcny_kernel)
        __syncthreads();

        inv_body<LOG2N>(sm, tb);

        const size_t boff = (size_t)bt * N * nout;
        float m;
        if (out_fmt == PCM_F32) m = inv_epilogue<LOG2N, PCM_F32>(sm, tb, tail, (float *)s.dout + boff, nout, o, frames);
        else if (out_fmt == PCM_S16) m = inv_epilogue<LOG2N, PCM_S16>(sm, tb, tail, (short *)s.dout + boff, nout, o, frames);
        else m = inv_epilogue<LOG2N, PCM_S24>(sm, tb, tail, (int *)s.dout + boff, nout, o, frames);
        lmax = fmaxf(lmax, m);
        block_max_update(s.bmax + bt, m);
        if (bt + 1 < T) __syncthreads();  // shared memory and the tail are reused by the next block
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, d));
    if ((tid & 31) == 0) red[tid >> 5] = lmax;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < NT / 32; w++) lmax = fmaxf(lmax, red[w]);
        // running maximum is >= 0, positive floats order like their bit patterns
        if (lmax > 0.0f) atomicMax(reinterpret_cast<int *>(s.maxv), __float_as_int(lmax));
    }
}


// ---------------------------------------------------------------------------
// fragm = 8192: the wavefront-lean transforms of fcv_fft13.cuh
// ---------------------------------------------------------------------------
#ifndef F13_INV_NT
#define F13_INV_NT 256   // threads of the inverse kernel: 256 (2 CTAs/SM, <= 128 registers) or 128 (3 CTAs/SM)
#endif
constexpr int f13_min_ctas(int nt) { return nt >= 256 ? 2 : 3; }

// Forward transform of the current block: one CTA = one half (blockIdx.x & 1) of the
// spectra of C consecutive input channels (blockIdx.x >> 1 = channel group) of one
// (stream, block): PCM and twiddles are fetched once for C transforms.
template <int FMT, int NCH, int C>
__global__ void __launch_bounds__(128 * C, f13_min_ctas(128 * C))
fwd13_stream_kernel(const StreamDev *__restrict__ st, const int *__restrict__ fv, int fv_all, f13::Tables tb,
                    int ninp, int R, int T, int pt, int reset_max) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c2 *sm = reinterpret_cast<c2 *>(smem_raw);
    constexpr int N = f13::N;
    pdl_trigger();
    pdl_wait();
    const int h = blockIdx.x & 1, ch0 = (blockIdx.x >> 1) * C, b = blockIdx.y, bt = blockIdx.z;
    const StreamDev s = st[b];
    int frames = (fv ? fv[b] : fv_all) - bt * N;
    frames = frames < 0 ? 0 : (frames > N ? N : frames);
    int slot = pt + bt;
    if (slot >= R) slot -= R;
    float2 *rows[C];
#pragma unroll
    for (int c = 0; c < C; c++) rows[c] = s.xring + (size_t)((ch0 + c) * R + slot) * N;
    // per-block maximum mode: the inverse kernel of this block starts from zero
    if (reset_max && blockIdx.x == 0 && bt == 0 && threadIdx.x == 0) *s.maxv = 0.0f;
    if (blockIdx.x == 0 && threadIdx.x == 0) s.bmax[bt] = 0.0f;  // this block's maximum starts from zero
    if (frames == 0) {  // silence: its spectrum is zero
#pragma unroll
        for (int c = 0; c < C; c++)
            for (int e = threadIdx.x; e < f13::Q; e += 128 * C) rows[c][h * f13::Q + e] = make_float2(0.f, 0.f);
        return;
    }
    const size_t wire = FMT == PCM_S16 ? 2 : 4;
    const void *in = reinterpret_cast<const char *>(s.din) + (size_t)bt * N * ninp * wire;
    if (h == 0) f13::fwd_half<0, FMT, NCH, C, 128 * C>(sm, tb, in, ninp, ch0, frames, rows);
    else f13::fwd_half<1, FMT, NCH, C, 128 * C>(sm, tb, in, ninp, ch0, frames, rows);
}

// Filter preparation (K6): src[row][N] floats -> dst[row][N] spectra, one half per CTA.
__global__ void __launch_bounds__(128, f13_min_ctas(128))
fwd13_raw_kernel(const float *__restrict__ src, float2 *__restrict__ dst, f13::Tables tb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c2 *sm = reinterpret_cast<c2 *>(smem_raw);
    const size_t r = blockIdx.y;
    float2 *rows[1] = {dst + r * f13::N};
    if (blockIdx.x == 0) f13::fwd_half<0, PCM_F32, 1, 1, 128>(sm, tb, src + r * f13::N, 1, 0, f13::N, rows);
    else f13::fwd_half<1, PCM_F32, 1, 1, 128>(sm, tb, src + r * f13::N, 1, 0, f13::N, rows);
}

// Inverse transform of every (stream, output channel), T blocks one after the other.
template <int FMT, bool PF>
__global__ void __launch_bounds__(F13_INV_NT, f13_min_ctas(F13_INV_NT))
inv13_stream_kernel(const StreamDev *__restrict__ st, const int *__restrict__ fv, int fv_all, f13::Tables tb,
                    const float2 *__restrict__ Y, const float2 *__restrict__ zc0, int nout, int T) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c2 *sm = reinterpret_cast<c2 *>(smem_raw);
    constexpr int NT = F13_INV_NT;
    __shared__ float red[NT / 32];
    constexpr int N = f13::N, M = N;
    pdl_trigger();
    pdl_wait();
    const int tid = threadIdx.x;
    const int o = blockIdx.x, b = blockIdx.y;
    const StreamDev s = st[b];
    const int fvb = fv ? fv[b] : fv_all;
    float2 *tail = reinterpret_cast<float2 *>(s.tail + (size_t)o * N);
    const size_t wire = FMT == PCM_S16 ? 2 : 4;
    float lmax = 0.0f;

    for (int bt = 0; bt < T; bt++) {
        int frames = fvb - bt * N;
        frames = frames < 0 ? 0 : (frames > N ? N : frames);
        // entry 0 of the sequence to transform comes from the two real bins (dcny_kernel)
        const float2 z0 = tid == 0 ? zc0[((size_t)b * nout + o) * T + bt] : make_float2(0.f, 0.f);
        const float2 *yrow = Y + (((size_t)b * nout + o) * T + bt) * M;
        if (PF && bt + 1 < T) {  // the next block's spectrum row (64 KB) is requested into L2 now
#pragma unroll
            for (int i = 0; i < (M * 8 / 128) / NT; i++) prefetch_l2(yrow + M + (size_t)(tid + i * NT) * 16);
        }
#pragma unroll 1
        for (int j = tid; j < 256; j += NT) {
            if (j < 128) f13::inv_pass_c<0>(sm, tb, yrow, c2_pack(z0.x, z0.y), j);
            else f13::inv_pass_c<1>(sm + f13::HALF_ELEMS, tb, yrow, 0ull, j - 128);
        }
        __syncthreads();
        f13::pass_b<+1, 2, NT>(sm, tb);
        __syncthreads();
        void *dout = reinterpret_cast<char *>(s.dout) + (size_t)bt * N * nout * wire;
        const float m = f13::inv_pass_a<FMT, NT>(sm, tb, tail, dout, nout, o, frames);
        lmax = fmaxf(lmax, m);
        block_max_update(s.bmax + bt, m);
        if (bt + 1 < T) __syncthreads();  // shared memory and the tail are reused by the next block
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, d));
    if ((tid & 31) == 0) red[tid >> 5] = lmax;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < NT / 32; w++) lmax = fmaxf(lmax, red[w]);
        // running maximum is >= 0, positive floats order like their bit patterns
        if (lmax > 0.0f) atomicMax(reinterpret_cast<int *>(s.maxv), __float_as_int(lmax));
    }
}

// ---------------------------------------------------------------------------
// per-device context: twiddle tables per partition size
// ---------------------------------------------------------------------------
struct DeviceCtx {
    std::mutex mu;
    bool checked = false;
    bool ok = false;
    std::string why;
    bool have[16] = {};
    FftTables tab[16];
    bool attr_set[16] = {};
    bool have13 = false;
    f13::Tables tab13{};
};
static std::mutex g_ctx_mu;
static std::map<int, DeviceCtx *> g_ctx;

static DeviceCtx *get_ctx(int device) {
    std::lock_guard<std::mutex> l(g_ctx_mu);
    auto it = g_ctx.find(device);
    if (it != g_ctx.end()) return it->second;
    DeviceCtx *c = new DeviceCtx();
    g_ctx[device] = c;
    return c;
}

static int check_device(int device) {
    DeviceCtx *c = get_ctx(device);
    std::lock_guard<std::mutex> l(c->mu);
    if (!c->checked) {
        c->checked = true;
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess) {
            c->why = std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e);
        } else if (device < 0 || device >= n) {
            c->why = "no such CUDA device";
        } else {
            cudaDeviceProp p;
            e = cudaGetDeviceProperties(&p, device);
            if (e != cudaSuccess) c->why = cudaGetErrorString(e);
            else if (p.major != 10) {
                char b[128];
                snprintf(b, sizeof(b), "device %d is sm_%d%d; this library is built for sm_100a only", device,
                         p.major, p.minor);
                c->why = b;
            } else c->ok = true;
        }
    }
    if (!c->ok) return fail(FCV_E_CUDA, "device %d unusable: %s", device, c->why.c_str());
    return 0;
}

template <int LOG2N>
static int set_attrs() {
    const int bytes = (int)fft_smem_bytes(LOG2N);
    CU_TRY(cudaFuncSetAttribute(fwd_stream_kernel<LOG2N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)fft_smem_bytes(LOG2N, 1)));
    CU_TRY(cudaFuncSetAttribute(fwd_raw_kernel<LOG2N>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CU_TRY(cudaFuncSetAttribute(inv_stream_kernel<LOG2N>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    return 0;
}

#define DISPATCH_LOG2N(l2, CALL)                    \
    switch (l2) {                                   \
        case 6: { constexpr int L = 6; CALL; } break;   \
        case 7: { constexpr int L = 7; CALL; } break;   \
        case 8: { constexpr int L = 8; CALL; } break;   \
        case 9: { constexpr int L = 9; CALL; } break;   \
        case 10: { constexpr int L = 10; CALL; } break; \
        case 11: { constexpr int L = 11; CALL; } break; \
        case 12: { constexpr int L = 12; CALL; } break; \
        case 13: { constexpr int L = 13; CALL; } break; \
        default: return fail(FCV_E_PARAM, "unsupported partition size 2^%d", l2); \
    }

// Builds (once per device and size) the twiddle tables in double precision.
static int get_tables(int device, int log2n, FftTables *out) {
    DeviceCtx *c = get_ctx(device);
    std::lock_guard<std::mutex> l(c->mu);
    if (!c->have[log2n]) {
        const int q = log2n - 1, Q = 1 << q, M = 2 * Q;
        const double PI = 3.14159265358979323846264338327950288;
        std::vector<float2> h;
        size_t offA = 0, offU, offP[4] = {0, 0, 0, 0};
        h.resize(Q);
        for (int n = 0; n < Q; n++) {
            const double a = -2.0 * PI * n / M;
            h[n] = make_float2((float)cos(a), (float)sin(a));
        }
        offU = h.size();
        h.resize(offU + M);
        for (int e = 0; e < M; e++) {
            const int half = e >> q, k = 2 * plan_revinv(q, e & (Q - 1)) + half;
            const double a = -PI * k / M;
            h[offU + e] = make_float2((float)cos(a), (float)sin(a));
        }
        const int np = plan_npass(q);
        for (int t = 0; t < np; t++) {
            const int R = 1 << plan_lr(q, t), S = 1 << plan_ls(q, t), Qt = 1 << plan_lqt(q, t);
            offP[t] = h.size();
            if (S > 1) {
                h.resize(offP[t] + (size_t)(R - 1) * S);
                for (int k1 = 1; k1 < R; k1++)
                    for (int u = 0; u < S; u++) {
                        const double a = -2.0 * PI * (double)u * (double)k1 / (double)Qt;
                        h[offP[t] + (size_t)(k1 - 1) * S + u] = make_float2((float)cos(a), (float)sin(a));
                    }
            }
        }
        // conjugate-partner entry of every entry: bin k <-> bin M - k (same half)
        std::vector<unsigned short> part((size_t)M);
        for (int e = 0; e < M; e++) {
            const int half = e >> q, kp = plan_revinv(q, e & (Q - 1));
            const int kpp = half ? (Q - 1 - kp) : ((Q - kp) & (Q - 1));
            part[e] = (unsigned short)((half << q) + plan_rev(q, kpp));
        }
        float2 *d = nullptr;
        unsigned short *dpart = nullptr;
        CU_TRY(cudaMalloc(&d, h.size() * sizeof(float2)));
        CU_TRY(cudaMemcpy(d, h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice));
        CU_TRY(cudaMalloc(&dpart, part.size() * sizeof(unsigned short)));
        CU_TRY(cudaMemcpy(dpart, part.data(), part.size() * sizeof(unsigned short), cudaMemcpyHostToDevice));
        FftTables tb;
        tb.part = dpart;
        tb.twA = d + offA;
        tb.twU = d + offU;
        for (int t = 0; t < 4; t++) tb.twP[t] = d + offP[t];
        c->tab[log2n] = tb;
        int rc = 0;
        DISPATCH_LOG2N(log2n, rc = set_attrs<L>());
        if (rc) return rc;
        c->have[log2n] = true;
    }
    *out = c->tab[log2n];
    return 0;
}


// fragm = 8192 runs the transforms of fcv_fft13.cuh (their spectrum layout differs from the
// generic kernels', so the choice is process-wide); FCV_GENERIC_FFT=1 keeps the generic ones.
static bool use_f13(int log2n) {
    static const bool generic = [] {
        const char *v = getenv("FCV_GENERIC_FFT");
        return v && *v && *v != '0';
    }();
    return log2n == f13::LOG2N && !generic;
}

template <int FMT>
static int set_attrs13() {
    const int one = (int)f13::HALF_BYTES, two = 2 * (int)f13::HALF_BYTES;
    CU_TRY(cudaFuncSetAttribute(fwd13_stream_kernel<FMT, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, two));
    CU_TRY(cudaFuncSetAttribute(fwd13_stream_kernel<FMT, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, one));
    CU_TRY(cudaFuncSetAttribute(fwd13_stream_kernel<FMT, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, one));
    CU_TRY(cudaFuncSetAttribute(inv13_stream_kernel<FMT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, two));
    CU_TRY(cudaFuncSetAttribute(inv13_stream_kernel<FMT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, two));
    return 0;
}

// Twiddle tables of the fragm = 8192 transforms (fcv_fft13.cuh), double precision on the host.
static int get_tables13(int device, f13::Tables *out) {
    DeviceCtx *c = get_ctx(device);
    std::lock_guard<std::mutex> l(c->mu);
    if (!c->have13) {
        const int M = f13::N, Q = f13::Q;
        const double PI = 3.14159265358979323846264338327950288;
        std::vector<float2> h((size_t)15 * 256 + 16 * 256 + 256 + 2 * Q);
        auto unit = [&](double turns) {  // exp(-2 pi i turns)
            const double a = -2.0 * PI * turns;
            return make_float2((float)cos(a), (float)sin(a));
        };
        size_t o0 = 0, o1 = o0 + 15 * 256, oB = o1 + 16 * 256, oU = oB + 256;
        for (int k0 = 1; k0 < 16; k0++)
            for (int u = 0; u < 256; u++) h[o0 + (size_t)(k0 - 1) * 256 + u] = unit((double)(2 * u * k0 % M) / M);
        for (int k0 = 0; k0 < 16; k0++)
            for (int u = 0; u < 256; u++) h[o1 + (size_t)k0 * 256 + u] = unit((double)(u * (2 * k0 + 1) % M) / M);
        for (int n0 = 0; n0 < 16; n0++)
            for (int k1 = 0; k1 < 16; k1++) h[oB + (size_t)n0 * 16 + k1] = unit((double)(n0 * k1) / 256.0);
        for (int e = 0; e < 2 * Q; e++) {
            const int k = 2 * (e & (Q - 1)) + (e >> (f13::LOG2N - 1));
            h[oU + e] = unit((double)k / (2.0 * M));
        }
        float2 *d = nullptr;
        CU_TRY(cudaMalloc(&d, h.size() * sizeof(float2)));
        CU_TRY(cudaMemcpy(d, h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice));
        c->tab13.twA0 = d + o0;
        c->tab13.twA1 = d + o1;
        c->tab13.twB = d + oB;
        c->tab13.twU = d + oU;
        int rc = set_attrs13<PCM_F32>();
        if (!rc) rc = set_attrs13<PCM_S16>();
        if (!rc) rc = set_attrs13<PCM_S24>();
        if (rc) return rc;
        CU_TRY(cudaFuncSetAttribute(fwd13_raw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f13::HALF_BYTES));
        c->have13 = true;
    }
    *out = c->tab13;
    return 0;
}

// ---------------------------------------------------------------------------
// filter
// ---------------------------------------------------------------------------
struct Pair {
    bool exists = false;      // a MAC node was created for this pair
    int link = -1;            // index of the pair whose spectra are used instead
    std::vector<float> h;     // time domain, npar * fragm, already scaled by 0.5/fragm
    std::vector<int> row;     // per partition: filter row or -1 (after commit)
};

struct fcv_filter {
    std::atomic<int> refs{1};
    int ninp = 0, nout = 0;
    unsigned size = 0;
    int fragm = 0, log2n = 0;
    int npar = 0;  // partitions zita allocates room for
    bool committed = false;
    std::vector<Pair> pairs;  // [inp * nout + out]
    // after commit
    int device = -1;
    int ring = 1;       // depth of the input-spectra ring
    int nrows = 0;      // non-zero (pair, partition) spectra
    int active_pairs = 0;
    int group_no = 1;   // outputs per MAC group
    int ngroups = 1;
    int nsteps = 0;
    float2 *dH = nullptr;
    MacStep *dsteps = nullptr;
    int *dgroup_off = nullptr;
    // per-output pair lists (time-tiled MAC, DC/Nyquist products)
    TTPair *dpairs = nullptr;
    int *dpair_off = nullptr;
    int *dtt_rows = nullptr;
    FftTables tb{};
    f13::Tables tb13{};   // fragm = 8192 only
    std::vector<MacStep> hsteps;
    std::vector<int> hgroup_off;
};

extern "C" int fcv_abi_version(void) { return FCV_ABI_VERSION; }
extern "C" const char *fcv_last_error(void) { return g_err.c_str(); }
extern "C" unsigned long long fcv_kernel_launches(void) { return g_launches.load(); }

extern "C" int fcv_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail(FCV_E_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    int ok = 0;
    for (int d = 0; d < n; d++) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, d) == cudaSuccess && p.major == 10) ok++;
    }
    return ok;
}

extern "C" fcv_filter *fcv_filter_begin(int ninp, int nout, unsigned size, unsigned fragm) {
    if (ninp < 1 || ninp > FCV_MAXINP || nout < 1 || nout > FCV_MAXOUT) {
        fail(FCV_E_PARAM, "inputs/outputs out of range (%d, %d)", ninp, nout);
        return nullptr;
    }
    if (fragm < FCV_MINPART || fragm > FCV_MAXQUANT || (fragm & (fragm - 1))) {
        fail(FCV_E_PARAM, "fragm %u is not a power of two in [%d, %d]", fragm, FCV_MINPART, FCV_MAXQUANT);
        return nullptr;
    }
    if (size > FCV_MAXSIZE) {
        fail(FCV_E_PARAM, "size %u exceeds %u", size, (unsigned)FCV_MAXSIZE);
        return nullptr;
    }
    fcv_filter *f = new (std::nothrow) fcv_filter();
    if (!f) { fail(FCV_E_ALLOC, "out of memory"); return nullptr; }
    f->ninp = ninp;
    f->nout = nout;
    f->size = size;
    f->fragm = (int)fragm;
    f->log2n = 0;
    while ((1u << f->log2n) < fragm) f->log2n++;
    f->npar = (int)((size + fragm - 1) / fragm);
    f->pairs.resize((size_t)ninp * nout);
    return f;
}

extern "C" int fcv_filter_add(fcv_filter *f, int inp, int out, int step, const float *data, int ind0, int ind1) {
    if (!f) return fail(FCV_E_PARAM, "null filter");
    if (f->committed) return fail(FCV_E_STATE, "filter already committed");
    if (inp < 0 || inp >= f->ninp || out < 0 || out >= f->nout) return fail(FCV_E_PARAM, "bad input/output index");
    // Convlevel::impdata_write range test: nothing to do if the data lies outside
    // [0, npar * fragm) -- no MAC node is created in that case either.
    const long n = (long)ind1 - (long)ind0;
    const long total = (long)f->npar * f->fragm;
    const long i0 = -(long)ind0;
    if (i0 >= n || i0 + total <= 0) return 0;
    Pair &p = f->pairs[(size_t)inp * f->nout + out];
    p.exists = true;
    if (p.link >= 0) return 0;  // linked pairs ignore new data
    if (!data) return 0;
    try {
        if (p.h.empty()) p.h.assign((size_t)total, 0.0f);
    } catch (...) {
        return fail(FCV_E_ALLOC, "out of memory");
    }
    const float norm = 0.5f / (float)f->fragm;
    const long j0 = i0 < 0 ? 0 : i0;
    const long j1 = (i0 + total > n) ? n : i0 + total;
    for (long j = j0; j < j1; j++) p.h[(size_t)(j - i0)] += norm * data[j * step];
    return 0;
}

extern "C" int fcv_filter_link(fcv_filter *f, int inp1, int out1, int inp2, int out2) {
    if (!f) return fail(FCV_E_PARAM, "null filter");
    if (inp1 < 0 || inp1 >= f->ninp || out1 < 0 || out1 >= f->nout || inp2 < 0 || inp2 >= f->ninp ||
        out2 < 0 || out2 >= f->nout || (inp1 == inp2 && out1 == out2))
        return fail(FCV_E_PARAM, "bad link");
    if (f->committed) return fail(FCV_E_STATE, "filter already committed");
    const int src = inp1 * f->nout + out1, dst = inp2 * f->nout + out2;
    if (!f->pairs[src].exists) return 0;  // Convlevel::impdata_link: no source node, no-op
    Pair &d = f->pairs[dst];
    d.exists = true;
    d.h.clear();
    d.h.shrink_to_fit();
    d.link = src;
    return 0;
}

static void filter_free_device(fcv_filter *f) {
    if (f->device >= 0) cudaSetDevice(f->device);
    if (f->dH) cudaFree(f->dH);
    if (f->dsteps) cudaFree(f->dsteps);
    if (f->dgroup_off) cudaFree(f->dgroup_off);
    if (f->dpairs) cudaFree(f->dpairs);
    if (f->dpair_off) cudaFree(f->dpair_off);
    if (f->dtt_rows) cudaFree(f->dtt_rows);
    f->dpairs = nullptr;
    f->dpair_off = nullptr;
    f->dtt_rows = nullptr;
    f->dH = nullptr;
    f->dsteps = nullptr;
    f->dgroup_off = nullptr;
}

extern "C" int fcv_filter_commit(fcv_filter *f, int device) {
    if (!f) return fail(FCV_E_PARAM, "null filter");
    if (f->committed) return fail(FCV_E_STATE, "filter already committed");
    int rc = check_device(device);
    if (rc) return rc;
    CU_TRY(cudaSetDevice(device));
    rc = get_tables(device, f->log2n, &f->tb);
    if (rc) return rc;
    if (use_f13(f->log2n)) {
        rc = get_tables13(device, &f->tb13);
        if (rc) return rc;
    }
    f->device = device;
    const int N = f->fragm;

    // Rows: every non-zero partition of every pair that owns data.
    std::vector<float> rows;
    int nrows = 0, last_part = -1;
    for (auto &p : f->pairs) {
        p.row.assign((size_t)(f->npar > 0 ? f->npar : 1), -1);  // at least one slot: ring depth is >= 1
        if (!p.exists || p.link >= 0 || p.h.empty()) continue;
        for (int j = 0; j < f->npar; j++) {
            const float *src = p.h.data() + (size_t)j * N;
            bool nz = false;
            for (int k = 0; k < N; k++)
                if (src[k] != 0.0f) { nz = true; break; }
            if (!nz) continue;
            p.row[j] = nrows++;
            rows.insert(rows.end(), src, src + N);
        }
    }
    // Resolve links (one level, as zita does: a link to a link sees no data).
    for (auto &p : f->pairs) {
        if (p.exists && p.link >= 0) {
            const Pair &s = f->pairs[(size_t)p.link];
            if (s.link < 0 && !s.row.empty()) p.row = s.row;
        }
    }
    f->active_pairs = 0;
    for (auto &p : f->pairs) {
        bool any = false;
        for (int j = 0; j < f->npar; j++)
            if (p.row[j] >= 0) { any = true; if (j > last_part) last_part = j; }
        if (any) f->active_pairs++;
    }
    f->nrows = nrows;
    f->ring = last_part + 1 > 0 ? last_part + 1 : 1;

    // MAC step table: outputs in groups of group_no.
    f->group_no = f->nout >= 8 ? 8 : (f->nout > 4 ? 8 : (f->nout > 2 ? 4 : f->nout));
    f->ngroups = (f->nout + f->group_no - 1) / f->group_no;
    f->hsteps.clear();
    f->hgroup_off.assign(1, 0);
    for (int g = 0; g < f->ngroups; g++) {
        for (int i = 0; i < f->ninp; i++)
            for (int j = 0; j < f->ring; j++) {
                MacStep sp;
                sp.inp = i;
                sp.part = j;
                bool any = false;
                for (int o = 0; o < MAC_NO_MAX; o++) {
                    const int oo = g * f->group_no + o;
                    sp.row[o] = (o < f->group_no && oo < f->nout) ? f->pairs[(size_t)i * f->nout + oo].row[j] : -1;
                    any |= sp.row[o] >= 0;
                }
                if (any) f->hsteps.push_back(sp);
            }
        f->hgroup_off.push_back((int)f->hsteps.size());
    }
    f->nsteps = (int)f->hsteps.size();

    // Per-output pair lists: for every output the inputs that feed it and, per
    // partition j < ring, the filter row (or -1).
    std::vector<TTPair> hpairs;
    std::vector<int> hpair_off(1, 0), htt_rows;
    for (int o = 0; o < f->nout; o++) {
        for (int i = 0; i < f->ninp; i++) {
            const Pair &p = f->pairs[(size_t)i * f->nout + o];
            bool any = false;
            for (int j = 0; j < f->ring && j < (int)p.row.size(); j++) any |= p.row[j] >= 0;
            if (!any) continue;
            TTPair tp;
            tp.inp = i;
            tp.rowbase = (int)htt_rows.size();
            for (int j = 0; j < f->ring; j++) htt_rows.push_back(j < (int)p.row.size() ? p.row[j] : -1);
            hpairs.push_back(tp);
        }
        hpair_off.push_back((int)hpairs.size());
    }
    CU_TRY(cudaMalloc(&f->dpairs, (hpairs.size() + 1) * sizeof(TTPair)));
    CU_TRY(cudaMalloc(&f->dpair_off, hpair_off.size() * sizeof(int)));
    CU_TRY(cudaMalloc(&f->dtt_rows, (htt_rows.size() + 1) * sizeof(int)));
    if (!hpairs.empty())
        CU_TRY(cudaMemcpy(f->dpairs, hpairs.data(), hpairs.size() * sizeof(TTPair), cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(f->dpair_off, hpair_off.data(), hpair_off.size() * sizeof(int), cudaMemcpyHostToDevice));
    if (!htt_rows.empty())
        CU_TRY(cudaMemcpy(f->dtt_rows, htt_rows.data(), htt_rows.size() * sizeof(int), cudaMemcpyHostToDevice));

    const size_t M = (size_t)N;
    // one extra, all-zero row after the filter rows: what the TMA-staged MAC loads for a
    // partition in which a pair has no data
    CU_TRY(cudaMalloc(&f->dH, (size_t)(nrows + 1) * M * sizeof(float2)));
    CU_TRY(cudaMemset(f->dH + (size_t)nrows * M, 0, M * sizeof(float2)));
    CU_TRY(cudaMalloc(&f->dsteps, (size_t)(f->nsteps > 0 ? f->nsteps : 1) * sizeof(MacStep)));
    CU_TRY(cudaMalloc(&f->dgroup_off, f->hgroup_off.size() * sizeof(int)));
    if (f->nsteps)
        CU_TRY(cudaMemcpy(f->dsteps, f->hsteps.data(), (size_t)f->nsteps * sizeof(MacStep), cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(f->dgroup_off, f->hgroup_off.data(), f->hgroup_off.size() * sizeof(int), cudaMemcpyHostToDevice));
    if (nrows) {
        float *dsrc = nullptr;
        CU_TRY(cudaMalloc(&dsrc, rows.size() * sizeof(float)));
        CU_TRY(cudaMemcpy(dsrc, rows.data(), rows.size() * sizeof(float), cudaMemcpyHostToDevice));
        const FftTables tb = f->tb;
        if (use_f13(f->log2n)) {
            fwd13_raw_kernel<<<dim3(2, nrows), 128, f13::HALF_BYTES>>>(dsrc, f->dH, f->tb13);
        } else {
            DISPATCH_LOG2N(f->log2n, (fwd_raw_kernel<L><<<nrows, fft_threads(L), fft_smem_bytes(L)>>>(dsrc, f->dH, tb)));
        }
        g_launches++;
        cudaError_t e = cudaDeviceSynchronize();
        cudaFree(dsrc);
        if (e != cudaSuccess) return fail(FCV_E_CUDA, "filter transform failed: %s", cudaGetErrorString(e));
    }
    // time-domain copies are no longer needed
    for (auto &p : f->pairs) { p.h.clear(); p.h.shrink_to_fit(); }
    f->committed = true;
    return 0;
}

extern "C" void fcv_filter_ref(fcv_filter *f) { if (f) f->refs++; }
extern "C" void fcv_filter_unref(fcv_filter *f) {
    if (!f) return;
    if (--f->refs == 0) {
        filter_free_device(f);
        delete f;
    }
}
extern "C" int fcv_filter_ninp(const fcv_filter *f) { return f ? f->ninp : 0; }
extern "C" int fcv_filter_nout(const fcv_filter *f) { return f ? f->nout : 0; }
extern "C" int fcv_filter_fragm(const fcv_filter *f) { return f ? f->fragm : 0; }
extern "C" int fcv_filter_partitions(const fcv_filter *f) { return f ? f->npar : 0; }
extern "C" int fcv_filter_ring_depth(const fcv_filter *f) { return f ? f->ring : 0; }
extern "C" int fcv_filter_active_rows(const fcv_filter *f) { return f ? f->nrows : 0; }
extern "C" int fcv_filter_active_pairs(const fcv_filter *f) { return f ? f->active_pairs : 0; }
extern "C" int fcv_filter_device(const fcv_filter *f) { return f ? f->device : -1; }

// packed-permuted device row -> natural order (N+1 interleaved complex)
static void unpermute_row(int log2n, const float2 *row, float *dst) {
    const int q = log2n - 1, Q = 1 << q, M = 2 * Q;
    dst[0] = row[0].x; dst[1] = 0.f;
    dst[2 * M] = row[0].y; dst[2 * M + 1] = 0.f;
    for (int k = 1; k < M; k++) {
        // fragm = 8192: split-parity natural layout (fcv_fft13.cuh); else digit-reversed halves
        const int e = ((k & 1) << q) + (use_f13(log2n) ? (k >> 1) : plan_rev(q, k >> 1));
        dst[2 * k] = row[e].x;
        dst[2 * k + 1] = row[e].y;
    }
}

extern "C" int fcv_filter_get_spectrum(fcv_filter *f, int inp, int out, int j, float *dst) {
    if (!f || !f->committed) return fail(FCV_E_STATE, "filter not committed");
    if (inp < 0 || inp >= f->ninp || out < 0 || out >= f->nout || j < 0 || j >= f->npar)
        return fail(FCV_E_PARAM, "bad index");
    const int row = f->pairs[(size_t)inp * f->nout + out].row[j];
    if (row < 0) return 0;
    CU_TRY(cudaSetDevice(f->device));
    std::vector<float2> h((size_t)f->fragm);
    CU_TRY(cudaMemcpy(h.data(), f->dH + (size_t)row * f->fragm, h.size() * sizeof(float2), cudaMemcpyDeviceToHost));
    unpermute_row(f->log2n, h.data(), dst);
    return 1;
}

extern "C" int fcv_filter_get_impulse(fcv_filter *f, int inp, int out, float *dst, int capacity) {
    if (!f || !dst || capacity < 0) return fail(FCV_E_PARAM, "null argument");
    if (f->committed) return fail(FCV_E_STATE, "impulses are dropped at commit");
    if (inp < 0 || inp >= f->ninp || out < 0 || out >= f->nout) return fail(FCV_E_PARAM, "bad index");
    const Pair &p = f->pairs[(size_t)inp * f->nout + out];
    if (!p.exists) return 0;
    const Pair *src = &p;
    if (p.link >= 0) src = &f->pairs[(size_t)p.link];
    const size_t total = (size_t)f->npar * f->fragm;
    const size_t n = total < (size_t)capacity ? total : (size_t)capacity;
    memset(dst, 0, (size_t)capacity * sizeof(float));
    if (src->link < 0 && !src->h.empty()) memcpy(dst, src->h.data(), n * sizeof(float));
    return p.link >= 0 ? 2 : 1;
}

// ---------------------------------------------------------------------------
// batch
// ---------------------------------------------------------------------------
static size_t pcm_bytes(int fmt) { return fmt == FCV_PCM_S16 ? 2 : 4; }

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
    // device
    unsigned char *dmem = nullptr;       // one slab
    float2 *xring = nullptr;
    float *tail = nullptr;
    unsigned char *din = nullptr, *dout = nullptr;
    float2 *Y = nullptr;
    float *maxv = nullptr;
    float *bmax = nullptr;               // [B][T] per-block maxima of the last step
    float2 *zc0 = nullptr;               // [B][nout][T] entry 0 of the sequences to inverse-transform (dcny_kernel)
    StreamDev *dst = nullptr;
    int *dfv = nullptr;
    size_t state_bytes_per_stream = 0;
    // host
    unsigned char *hin = nullptr, *hout = nullptr;
    int *hfv = nullptr;
    // second host staging slot + per-slot completion events for the asynchronous submit/wait pair
    unsigned char *hin1 = nullptr, *hout1 = nullptr;
    int *dfv1 = nullptr, *hfv1 = nullptr;
    float *hbmax[2] = {nullptr, nullptr};   // [B][T] block maxima of the step submitted from each host slot
    cudaEvent_t slot_done[2][4] = {};
    bool slot_busy[2] = {false, false};
    // streams
    static const int NQ = 4;
    cudaStream_t q[NQ] = {};
    cudaEvent_t fj[5] = {};  // fork/join events of the chunked device path
    // stopwatch
    cudaEvent_t sw[16] = {};
    // profiling
    bool profiling = false;
    std::vector<cudaEvent_t> ev;  // 4 events per step: t0 | fwd | mac | inv
    size_t ev_used = 0;
    int prof_steps = 0;
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static void batch_free(fcv_batch *b) {
    if (!b) return;
    if (b->f && b->f->device >= 0) cudaSetDevice(b->f->device);
    for (auto e : b->ev) cudaEventDestroy(e);
    for (int i = 0; i < 16; i++) if (b->sw[i]) cudaEventDestroy(b->sw[i]);
    for (int i = 0; i < 5; i++) if (b->fj[i]) cudaEventDestroy(b->fj[i]);
    for (int i = 0; i < fcv_batch::NQ; i++)
        if (b->q[i]) { cudaStreamSynchronize(b->q[i]); cudaStreamDestroy(b->q[i]); }
    if (b->dmem) cudaFree(b->dmem);
    if (b->hin) cudaFreeHost(b->hin);
    if (b->hout) cudaFreeHost(b->hout);
    if (b->hfv) cudaFreeHost(b->hfv);
    if (b->hin1) cudaFreeHost(b->hin1);
    if (b->hout1) cudaFreeHost(b->hout1);
    if (b->hfv1) cudaFreeHost(b->hfv1);
    if (b->dfv1) cudaFree(b->dfv1);
    for (int k = 0; k < 2; k++) if (b->hbmax[k]) cudaFreeHost(b->hbmax[k]);
    for (int k = 0; k < 2; k++)
        for (int i = 0; i < 4; i++)
            if (b->slot_done[k][i]) cudaEventDestroy(b->slot_done[k][i]);
    if (b->f) fcv_filter_unref(b->f);
    delete b;
}

static fcv_batch *batch_create(fcv_filter *f, int nstreams, int in_fmt, int out_fmt, bool shared_host_buffer, int T) {
    if (!f || !f->committed) { fail(FCV_E_STATE, "filter not committed"); return nullptr; }
    if (nstreams < 1) { fail(FCV_E_PARAM, "nstreams < 1"); return nullptr; }
    if (T != 1 && T != 2 && T != 4 && T != 8) { fail(FCV_E_PARAM, "blocks per step must be 1, 2, 4 or 8"); return nullptr; }
    if (in_fmt < 0 || in_fmt > FCV_PCM_S24 || out_fmt < 0 || out_fmt > FCV_PCM_S24) {
        fail(FCV_E_PARAM, "bad PCM format");
        return nullptr;
    }
    if (cudaSetDevice(f->device) != cudaSuccess) { fail(FCV_E_CUDA, "cudaSetDevice failed"); return nullptr; }
    fcv_batch *b = new (std::nothrow) fcv_batch();
    if (!b) { fail(FCV_E_ALLOC, "out of memory"); return nullptr; }
    b->f = f;
    fcv_filter_ref(f);
    b->B = nstreams;
    b->in_fmt = in_fmt;
    b->out_fmt = out_fmt;
    const size_t N = (size_t)f->fragm, B = (size_t)nstreams;
    b->T = T;
    b->R = f->ring + T - 1;
    b->in_block = (size_t)T * N * f->ninp * pcm_bytes(in_fmt);
    b->out_block = (size_t)T * N * f->nout * pcm_bytes(out_fmt);
    b->out_pad = 256;
    b->per_block_max = shared_host_buffer;
    if (cudaDeviceGetAttribute(&b->num_sms, cudaDevAttrMultiProcessorCount, f->device) != cudaSuccess || b->num_sms < 1)
        b->num_sms = 148;

    const size_t xring_b = align_up(B * f->ninp * b->R * N * sizeof(float2), 256);
    const size_t tail_b = align_up(B * f->nout * N * sizeof(float), 256);
    const size_t din_b = align_up(B * b->in_block, 256);
    const size_t dout_b = align_up(B * b->out_block + b->out_pad, 256);
    const size_t y_b = align_up(B * f->nout * T * N * sizeof(float2), 256);
    const size_t max_b = align_up(B * sizeof(float), 256);
    const size_t bmax_b = align_up(B * (size_t)T * sizeof(float), 256);
    const size_t st_b = align_up(B * sizeof(StreamDev), 256);
    const size_t fv_b = align_up(B * sizeof(int), 256);
    const size_t zc_b = align_up(B * f->nout * (size_t)T * sizeof(float2), 256);
    const size_t total = xring_b + tail_b + din_b + dout_b + y_b + max_b + bmax_b + st_b + fv_b + zc_b;
    cudaError_t e = cudaMalloc(&b->dmem, total);
    if (e != cudaSuccess) {
        fail(FCV_E_ALLOC, "cudaMalloc(%zu bytes) failed: %s", total, cudaGetErrorString(e));
        batch_free(b);
        return nullptr;
    }
    unsigned char *p = b->dmem;
    b->xring = (float2 *)p; p += xring_b;
    b->tail = (float *)p; p += tail_b;
    b->din = p; p += din_b;
    b->dout = p; p += dout_b;
    b->Y = (float2 *)p; p += y_b;
    // single-stream mode keeps the running maximum right behind the output block
    // so that one device->host copy brings back both
    b->maxv = shared_host_buffer ? (float *)(b->dout + B * b->out_block) : (float *)p;
    p += max_b;
    b->bmax = (float *)p; p += bmax_b;
    b->dst = (StreamDev *)p; p += st_b;
    b->dfv = (int *)p; p += fv_b;
    b->zc0 = (float2 *)p; p += zc_b;
    b->state_bytes_per_stream = (size_t)f->ninp * b->R * N * sizeof(float2);

    bool ok = cudaMemset(b->dmem, 0, total) == cudaSuccess;
    if (shared_host_buffer) {
        // SoundProcessor::buffer_: fragm * max(ninp, nout) floats (+ the max mirror)
        const size_t bytes = N * (size_t)(f->ninp > f->nout ? f->ninp : f->nout) * sizeof(float) + b->out_pad;
        ok = ok && cudaHostAlloc((void **)&b->hin, bytes, cudaHostAllocMapped) == cudaSuccess;
        if (ok) memset(b->hin, 0, bytes);
        // The forward kernel of a single stream reads its block straight from this pinned host
        // buffer (64 KB of coalesced reads over the link): one copy operation and one driver call
        // less per block than staging it in device memory first.  FCV_STREAM_ZEROCOPY=0: stage.
        static const bool zc = !(getenv("FCV_STREAM_ZEROCOPY") && atoi(getenv("FCV_STREAM_ZEROCOPY")) == 0);
        void *dp = nullptr;
        b->in_zero_copy = ok && zc && cudaHostGetDevicePointer(&dp, b->hin, 0) == cudaSuccess && dp;
        if (b->in_zero_copy) b->hin_dev = dp;
    } else {
        ok = ok && cudaHostAlloc((void **)&b->hin, B * b->in_block, cudaHostAllocDefault) == cudaSuccess;
        ok = ok && cudaHostAlloc((void **)&b->hout, B * b->out_block, cudaHostAllocDefault) == cudaSuccess;
        if (ok) { memset(b->hin, 0, B * b->in_block); memset(b->hout, 0, B * b->out_block); }
    }
    std::vector<StreamDev> hs(B);
    for (size_t s = 0; s < B; s++) {
        hs[s].xring = b->xring + s * f->ninp * b->R * N;
        hs[s].tail = b->tail + s * f->nout * N;
        hs[s].din = b->in_zero_copy ? b->hin_dev : (const void *)(b->din + s * b->in_block);
        hs[s].dout = b->dout + s * b->out_block;
        hs[s].maxv = b->maxv + s;
        hs[s].bmax = b->bmax + s * (size_t)T;
    }
    ok = ok && cudaMemcpy(b->dst, hs.data(), B * sizeof(StreamDev), cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaHostAlloc((void **)&b->hfv, B * sizeof(int), cudaHostAllocDefault) == cudaSuccess;
    const int nq = shared_host_buffer ? 1 : fcv_batch::NQ;
    for (int i = 0; ok && i < nq; i++) ok = cudaStreamCreateWithFlags(&b->q[i], cudaStreamNonBlocking) == cudaSuccess;
    if (!ok) {
        fail(FCV_E_ALLOC, "batch allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        batch_free(b);
        return nullptr;
    }
    return b;
}

// <<<>>> with the programmatic-stream-serialization attribute when `pdl` (single-stream path:
// the next kernel's launch latency hides behind the tail of the previous one)
template <typename... KArgs, typename... Args>
static void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t q, bool pdl, Args... args) {
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
static bool use_pdl(const fcv_batch *b) {
    static const bool on = !(getenv("FCV_PDL") && atoi(getenv("FCV_PDL")) == 0);
    return on && b->per_block_max;   // single-stream handles only
}

template <int NO, int S>
static void launch_mac(const fcv_batch *b, int off, int cnt, int pt, cudaStream_t q) {
    const fcv_filter *f = b->f;
    const int M4 = f->fragm / 2;
    const int TPB = M4 >= 128 ? 128 : M4;  // M4 is a power of two >= 32
    dim3 grid(M4 / TPB + 1, (cnt + S - 1) / S, f->ngroups);   // + 1: the DC / Nyquist column
    const float4 *H = reinterpret_cast<const float4 *>(f->dH);
    float4 *Y = reinterpret_cast<float4 *>(b->Y + (size_t)off * f->nout * f->fragm);
#define FCV_MAC1_ARGS b->dst + off, cnt, f->dsteps, f->dgroup_off, H, Y, M4, b->R, pt, f->nout, f->dpairs, f->dpair_off, \
                      f->dtt_rows, b->zc0 + (size_t)off * f->nout, f->ring
    if (TPB == 128)
        launch_k(mac_kernel<NO, S, 128>, grid, dim3(128), 0, q, use_pdl(b), FCV_MAC1_ARGS);
    else if (TPB == 64)
        launch_k(mac_kernel<NO, S, 64>, grid, dim3(64), 0, q, use_pdl(b), FCV_MAC1_ARGS);
    else
        launch_k(mac_kernel<NO, S, 32>, grid, dim3(32), 0, q, use_pdl(b), FCV_MAC1_ARGS);
#undef FCV_MAC1_ARGS
}

// Time-tiled MAC: T blocks per stream per launch, one output channel per grid.z.
template <int T, int S>
static void launch_mac_tt_s(const fcv_batch *b, int off, int cnt, int newest, cudaStream_t q) {
    const fcv_filter *f = b->f;
    const int M4 = f->fragm / 2;
    const int TPB = M4 >= 128 ? 128 : M4;
    dim3 grid(M4 / TPB, (cnt + S - 1) / S, f->nout);
    const float4 *H = reinterpret_cast<const float4 *>(f->dH);
    float4 *Y = reinterpret_cast<float4 *>(b->Y + (size_t)off * f->nout * T * f->fragm);
#define FCV_TT_ARGS b->dst + off, cnt, f->dpairs, f->dpair_off, f->dtt_rows, H, Y, M4, f->ring, b->R, newest, f->nout
    if (TPB == 128) mac_tt_kernel<T, S, 128><<<grid, 128, 0, q>>>(FCV_TT_ARGS);
    else if (TPB == 64) mac_tt_kernel<T, S, 64><<<grid, 64, 0, q>>>(FCV_TT_ARGS);
    else mac_tt_kernel<T, S, 32><<<grid, 32, 0, q>>>(FCV_TT_ARGS);
#undef FCV_TT_ARGS
}

// TMA-staged variant (fcv_mac_tma.cuh).  Returns false when the shape is not covered.
template <int T, int S, int NS, int MC = tma::min_ctas(T, S)>
static bool launch_mac_tma(const fcv_batch *b, int off, int cnt, int newest, cudaStream_t q) {
    const fcv_filter *f = b->f;
    const int M4 = f->fragm / 2;
    if (M4 % tma::TPB != 0 || cnt < S) return false;
    const size_t smem = tma::smem_bytes(S, NS);
    // dynamic + static shared memory exceeds the 48 KB default; the attribute is per device
    if (cudaFuncSetAttribute(tma::mac_tma_kernel<T, S, NS, MC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
        return false;
    // one CTA per work item; FCV_MAC_PERSIST=n runs a persistent grid of n CTAs per SM instead
    // (measured 5 % slower on SantaLucia x 1024 streams: 1.05 vs 1.00 ms, profiles/r01_experiments.md)
    static const int persist = getenv("FCV_MAC_PERSIST") ? atoi(getenv("FCV_MAC_PERSIST")) : 0;
    const int ntiles = M4 / tma::TPB, ngroups = (cnt + S - 1) / S, nitems = ntiles * ngroups * f->nout;
    int grid = nitems;
    if (persist > 0 && b->num_sms * persist < nitems) grid = b->num_sms * persist;
    const float4 *H = reinterpret_cast<const float4 *>(f->dH);
    float4 *Y = reinterpret_cast<float4 *>(b->Y + (size_t)off * f->nout * T * f->fragm);
    tma::mac_tma_kernel<T, S, NS, MC><<<grid, tma::THREADS, smem, q>>>(b->dst + off, cnt, f->dpairs, f->dpair_off,
                                                                     f->dtt_rows, H, Y, M4, f->ring, b->R, newest,
                                                                     f->nout, f->nrows, ntiles, ngroups, nitems);
    return true;
}

template <int T>
static void launch_mac_tt(const fcv_batch *b, int off, int cnt, int newest, cudaStream_t q) {
    // T = 4 and T = 8 stream their rows through TMA-staged tiles (fcv_mac_tma.cuh) whenever the
    // shape allows (spectrum tiles of 2 KB, at least two streams); measured equal to the
    // register-pipelined kernel below at T = 8 and 2 % faster at T = 4.  FCV_MAC_TMA=0 turns
    // it off, 2 / 3 select other stream tilings (experiments).
    static const int use_tma = getenv("FCV_MAC_TMA") ? atoi(getenv("FCV_MAC_TMA")) : 1;
    if (use_tma) {
        if (T == 8 && use_tma == 3 && launch_mac_tma<8, 1, 12>(b, off, cnt, newest, q)) return;
        if (T == 8 && launch_mac_tma<8, 2, 8>(b, off, cnt, newest, q)) return;
        if (T == 4 && use_tma == 2 && launch_mac_tma<4, 4, 6>(b, off, cnt, newest, q)) return;
        if (T == 4 && launch_mac_tma<4, 2, 8>(b, off, cnt, newest, q)) return;
    }
    // Streams per thread, sharing each filter value from registers.  Measured on B200
    // (SantaLucia, 1024 streams): T=4: S=4 0.150 ms/block (HBM floor 0.148), S=2 0.166;
    // T=8: S=2 0.134, S=4 0.143 (228 registers).  FCV_TT_S overrides for experiments.
    static const int env_s = getenv("FCV_TT_S") ? atoi(getenv("FCV_TT_S")) : 0;
    int S = env_s ? env_s : (T >= 8 ? 2 : 4);
    if (cnt < S) S = cnt >= 2 ? 2 : 1;
    if (S >= 4) launch_mac_tt_s<T, 4>(b, off, cnt, newest, q);
    else if (S >= 2) launch_mac_tt_s<T, 2>(b, off, cnt, newest, q);
    else launch_mac_tt_s<T, 1>(b, off, cnt, newest, q);
}

// fragm = 8192 forward launch: stereo and mono blocks take the vector-load kernels (stereo:
// both channels per CTA), any other channel count one channel per CTA with scalar loads.
template <int FMT>
static void launch_fwd13_fmt(const fcv_batch *b, int off, int cnt, const int *fv, int fv_all, int pt, cudaStream_t q) {
    const fcv_filter *f = b->f;
    const int rm = b->per_block_max ? 1 : 0;
    if (f->ninp == 2)
        launch_k(fwd13_stream_kernel<FMT, 2, 2>, dim3(2, cnt, b->T), dim3(256), 2 * f13::HALF_BYTES, q, use_pdl(b),
                 b->dst + off, fv, fv_all, f->tb13, f->ninp, b->R, b->T, pt, rm);
    else if (f->ninp == 1)
        launch_k(fwd13_stream_kernel<FMT, 1, 1>, dim3(2, cnt, b->T), dim3(128), f13::HALF_BYTES, q, use_pdl(b),
                 b->dst + off, fv, fv_all, f->tb13, f->ninp, b->R, b->T, pt, rm);
    else
        launch_k(fwd13_stream_kernel<FMT, 0, 1>, dim3(2 * f->ninp, cnt, b->T), dim3(128), f13::HALF_BYTES, q, use_pdl(b),
                 b->dst + off, fv, fv_all, f->tb13, f->ninp, b->R, b->T, pt, rm);
}
static void launch_fwd13(const fcv_batch *b, int off, int cnt, const int *fv, int fv_all, int pt, cudaStream_t q) {
    if (b->in_fmt == PCM_F32) launch_fwd13_fmt<PCM_F32>(b, off, cnt, fv, fv_all, pt, q);
    else if (b->in_fmt == PCM_S16) launch_fwd13_fmt<PCM_S16>(b, off, cnt, fv, fv_all, pt, q);
    else launch_fwd13_fmt<PCM_S24>(b, off, cnt, fv, fv_all, pt, q);
}

// The three launches for streams [off, off+cnt) of the batch on CUDA stream q.
// fv_base: per-stream valid-frame counts on the device, or nullptr = `fv_all` frames for every stream
// (-1: the whole step)
static int run_kernels(fcv_batch *b, int off, int cnt, const int *fv_base, cudaStream_t q, cudaEvent_t *ev,
                       int fv_all = -1) {
    fcv_filter *f = b->f;
    const int T = b->T, R = b->R;
    // the step's first block goes to ring slot (step * T) mod R
    const int pt = (int)((b->step * (unsigned long long)T) % (unsigned long long)R);
    const int *fv = fv_base ? fv_base + off : nullptr;
    if (fv_all < 0) fv_all = b->T * f->fragm;
    const FftTables tb = f->tb;
    // diagnostic: FCV_ONLY=1|2|4 (bit mask fwd|mac|inv) launches only those kernels
    static const int only = getenv("FCV_ONLY") ? atoi(getenv("FCV_ONLY")) : 7;
    if (ev) cudaEventRecord(ev[0], q);
    const bool k13 = use_f13(f->log2n);
    if (!(only & 1)) {
    } else if (k13) {
        launch_fwd13(b, off, cnt, fv, fv_all, pt, q);
    } else
    DISPATCH_LOG2N(f->log2n, (fwd_stream_kernel<L><<<dim3(2 * f->ninp, cnt, T), fft_threads(L, 1), fft_smem_bytes(L, 1), q>>>(
                                  b->dst + off, fv, fv_all, tb, f->ninp, R, T, pt, b->in_fmt, b->per_block_max ? 1 : 0)));
    if (ev) cudaEventRecord(ev[1], q);
    if (!(only & 2)) {
    } else if (T == 1) {
        const int S = cnt >= 4 ? 4 : (cnt >= 2 ? 2 : 1);
        switch (f->group_no) {
            case 1: if (S == 4) launch_mac<1, 4>(b, off, cnt, pt, q); else if (S == 2) launch_mac<1, 2>(b, off, cnt, pt, q); else launch_mac<1, 1>(b, off, cnt, pt, q); break;
            case 2: if (S == 4) launch_mac<2, 4>(b, off, cnt, pt, q); else if (S == 2) launch_mac<2, 2>(b, off, cnt, pt, q); else launch_mac<2, 1>(b, off, cnt, pt, q); break;
            case 4: if (S >= 2) launch_mac<4, 2>(b, off, cnt, pt, q); else launch_mac<4, 1>(b, off, cnt, pt, q); break;
            default: if (S >= 2) launch_mac<8, 2>(b, off, cnt, pt, q); else launch_mac<8, 1>(b, off, cnt, pt, q); break;
        }
    } else {
        const int newest = (pt + T - 1) % R;
        if (T == 2) launch_mac_tt<2>(b, off, cnt, newest, q);
        else if (T == 4) launch_mac_tt<4>(b, off, cnt, newest, q);
        else launch_mac_tt<8>(b, off, cnt, newest, q);
    }
    if (ev) cudaEventRecord(ev[2], q);
    const float2 *Y = b->Y + (size_t)off * f->nout * T * f->fragm;
    float2 *zc0 = b->zc0 + (size_t)off * f->nout * T;
    if ((only & 4) && T > 1) {   // T == 1: done inside mac_kernel
        const int nwarps = cnt * f->nout * T;
        dcny_kernel<<<(nwarps + 7) / 8, 256, 0, q>>>(b->dst + off, f->dpairs, f->dpair_off, f->dtt_rows, f->dH, zc0,
                                                     nwarps, f->nout, f->ring, R, T, pt, f->fragm);
    }
    if (!(only & 4)) {
    } else if (k13) {
#define FCV_INV13_ARGS b->dst + off, fv, fv_all, f->tb13, Y, zc0, f->nout, T
        const dim3 grid(f->nout, cnt);
        const size_t smem = 2 * f13::HALF_BYTES;
        static const bool pf = !(getenv("FCV_INV_PF") && atoi(getenv("FCV_INV_PF")) == 0);
#define FCV_INV13_LAUNCH(F) \
        do { if (pf && T > 1) inv13_stream_kernel<F, true><<<grid, F13_INV_NT, smem, q>>>(FCV_INV13_ARGS); \
             else launch_k(inv13_stream_kernel<F, false>, grid, dim3(F13_INV_NT), smem, q, use_pdl(b), FCV_INV13_ARGS); } while (0)
        if (b->out_fmt == PCM_F32) FCV_INV13_LAUNCH(PCM_F32);
        else if (b->out_fmt == PCM_S16) FCV_INV13_LAUNCH(PCM_S16);
        else FCV_INV13_LAUNCH(PCM_S24);
#undef FCV_INV13_LAUNCH
#undef FCV_INV13_ARGS
    } else
    DISPATCH_LOG2N(f->log2n, (inv_stream_kernel<L><<<dim3(f->nout, cnt), fft_threads(L), fft_smem_bytes(L), q>>>(
                                  b->dst + off, fv, fv_all, tb, Y, zc0, f->nout, T, b->out_fmt)));
    if (ev) cudaEventRecord(ev[3], q);
    g_launches += T > 1 ? 4 : 3;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(FCV_E_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    return 0;
}

static cudaEvent_t *prof_events(fcv_batch *b) {
    if (!b->profiling) return nullptr;
    if (b->ev_used + 4 > b->ev.size()) {
        for (int i = 0; i < 4; i++) {
            cudaEvent_t e;
            if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
            b->ev.push_back(e);
        }
    }
    cudaEvent_t *p = &b->ev[b->ev_used];
    b->ev_used += 4;
    b->prof_steps++;
    return p;
}

extern "C" fcv_batch *fcv_batch_create(fcv_filter *f, int nstreams, int in_format, int out_format) {
    return batch_create(f, nstreams, in_format, out_format, false, 1);
}
extern "C" fcv_batch *fcv_batch_create_tiled(fcv_filter *f, int nstreams, int in_format, int out_format,
                                             int blocks_per_step) {
    return batch_create(f, nstreams, in_format, out_format, false, blocks_per_step);
}
extern "C" int fcv_batch_blocks_per_step(const fcv_batch *b) { return b ? b->T : 0; }
extern "C" void fcv_batch_destroy(fcv_batch *b) { batch_free(b); }
extern "C" int fcv_batch_nstreams(const fcv_batch *b) { return b ? b->B : 0; }
extern "C" void *fcv_batch_host_in(fcv_batch *b) { return b ? b->hin : nullptr; }
extern "C" void *fcv_batch_host_out(fcv_batch *b) { return b ? b->hout : nullptr; }
extern "C" size_t fcv_batch_host_in_bytes(const fcv_batch *b) { return b ? (size_t)b->B * b->in_block : 0; }
extern "C" size_t fcv_batch_host_out_bytes(const fcv_batch *b) { return b ? (size_t)b->B * b->out_block : 0; }
extern "C" void *fcv_batch_device_in(fcv_batch *b) { return b ? b->din : nullptr; }
extern "C" void *fcv_batch_device_out(fcv_batch *b) { return b ? b->dout : nullptr; }
extern "C" void *fcv_batch_cuda_stream(fcv_batch *b) { return b ? (void *)b->q[0] : nullptr; }

static int stage_fv(fcv_batch *b, const int *frames_valid, cudaStream_t q) {
    for (int s = 0; s < b->B; s++) {
        const int v = frames_valid[s];
        if (v < 0 || v > b->T * b->f->fragm) return fail(FCV_E_PARAM, "frames_valid[%d] = %d out of range", s, v);
        b->hfv[s] = v;
    }
    CU_TRY(cudaMemcpyAsync(b->dfv, b->hfv, (size_t)b->B * sizeof(int), cudaMemcpyHostToDevice, q));
    return 0;
}

extern "C" int fcv_batch_process_device(fcv_batch *b, const int *frames_valid) {
    if (!b) return fail(FCV_E_PARAM, "null batch");
    CU_TRY(cudaSetDevice(b->f->device));
    if (frames_valid) {
        // the staging array is reused: the previous block must be done with it
        CU_TRY(cudaStreamSynchronize(b->q[0]));
        int rc = stage_fv(b, frames_valid, b->q[0]);
        if (rc) return rc;
    }
    // Optional fork/join over the batch's CUDA streams: the FFT kernels (issue and
    // shared-memory bound) of one part of the batch can then overlap the MAC kernel
    // (HBM bound) of another.  q[0] stays the stream everything is ordered on.
    static const int env_chunks = getenv("FCV_DEVICE_CHUNKS") ? atoi(getenv("FCV_DEVICE_CHUNKS")) : 0;
    int nchunk = (b->profiling || env_chunks < 2) ? 1 : env_chunks;
    if (nchunk > b->B / 8) nchunk = b->B / 8 > 0 ? b->B / 8 : 1;
    if (nchunk == 1) {
        int rc = run_kernels(b, 0, b->B, frames_valid ? b->dfv : nullptr, b->q[0], prof_events(b));
        if (rc) return rc;
        b->step++;
        return 0;
    }
    for (int i = 0; i < fcv_batch::NQ + 1; i++)
        if (!b->fj[i]) CU_TRY(cudaEventCreateWithFlags(&b->fj[i], cudaEventDisableTiming));
    CU_TRY(cudaEventRecord(b->fj[fcv_batch::NQ], b->q[0]));
    for (int i = 1; i < fcv_batch::NQ; i++) CU_TRY(cudaStreamWaitEvent(b->q[i], b->fj[fcv_batch::NQ], 0));
    const int per = ((b->B + nchunk - 1) / nchunk + 1) & ~1;
    for (int c = 0, off = 0; off < b->B; c++, off += per) {
        const int cnt = (b->B - off) < per ? (b->B - off) : per;
        int rc = run_kernels(b, off, cnt, frames_valid ? b->dfv : nullptr, b->q[c % fcv_batch::NQ], nullptr);
        if (rc) return rc;
    }
    for (int i = 1; i < fcv_batch::NQ; i++) {
        CU_TRY(cudaEventRecord(b->fj[i], b->q[i]));
        CU_TRY(cudaStreamWaitEvent(b->q[0], b->fj[i], 0));
    }
    b->step++;
    return 0;
}

extern "C" int fcv_batch_sync(fcv_batch *b) {
    if (!b) return fail(FCV_E_PARAM, "null batch");
    CU_TRY(cudaSetDevice(b->f->device));
    for (int i = 0; i < fcv_batch::NQ; i++)
        if (b->q[i]) CU_TRY(cudaStreamSynchronize(b->q[i]));
    return 0;
}

static int ensure_slot1(fcv_batch *b) {
    if (b->hin1) return 0;
    const size_t B = (size_t)b->B;
    CU_TRY(cudaHostAlloc((void **)&b->hin1, B * b->in_block, cudaHostAllocDefault));
    CU_TRY(cudaHostAlloc((void **)&b->hout1, B * b->out_block, cudaHostAllocDefault));
    CU_TRY(cudaHostAlloc((void **)&b->hfv1, B * sizeof(int), cudaHostAllocDefault));
    CU_TRY(cudaMalloc((void **)&b->dfv1, B * sizeof(int)));
    memset(b->hin1, 0, B * b->in_block);
    memset(b->hout1, 0, B * b->out_block);
    return 0;
}

extern "C" void *fcv_batch_host_in_slot(fcv_batch *b, int slot) {
    if (!b || !b->hout || slot < 0 || slot > 1) return nullptr;
    if (slot == 1 && (cudaSetDevice(b->f->device) != cudaSuccess || ensure_slot1(b))) return nullptr;
    return slot ? b->hin1 : b->hin;
}
extern "C" void *fcv_batch_host_out_slot(fcv_batch *b, int slot) {
    if (!b || !b->hout || slot < 0 || slot > 1) return nullptr;
    if (slot == 1 && (cudaSetDevice(b->f->device) != cudaSuccess || ensure_slot1(b))) return nullptr;
    return slot ? b->hout1 : b->hout;
}

extern "C" int fcv_batch_wait(fcv_batch *b, int slot) {
    if (!b || slot < 0 || slot > 1) return fail(FCV_E_PARAM, "bad slot");
    if (!b->slot_busy[slot]) return 0;
    CU_TRY(cudaSetDevice(b->f->device));
    for (int i = 0; i < fcv_batch::NQ; i++)
        if (b->slot_done[slot][i]) CU_TRY(cudaEventSynchronize(b->slot_done[slot][i]));
    b->slot_busy[slot] = false;
    return 0;
}

// Chunks of streams a submit is split into (each chunk: copy in, kernels, copy out on one of NQ
// CUDA streams, round robin -- a given stream of the batch always lands on the same CUDA stream).
static int submit_chunks(const fcv_batch *b) {
    if (b->profiling) return 1;
    static const int env_chunks = getenv("FCV_CHUNKS") ? atoi(getenv("FCV_CHUNKS")) : 0;  // tuning knob
    int nchunk = env_chunks > 0 ? env_chunks : b->B / 128;
    if (nchunk > 16) nchunk = 16;
    if (nchunk > b->B) nchunk = b->B;
    if (nchunk < 1) nchunk = 1;
    return nchunk;
}
static cudaStream_t submit_stream_of(const fcv_batch *b, int stream_index) {
    const int nchunk = submit_chunks(b);
    const int per = (b->B + nchunk - 1) / nchunk;
    return b->q[nchunk == 1 ? 0 : (stream_index / per) % fcv_batch::NQ];
}

// Enqueue one block for every stream from host staging slot `slot`:
// per chunk of streams host->device copy, the three kernels, device->host copy,
// round-robin over NQ CUDA streams.  Chunks of consecutive submits run in order
// on their stream, so the device state needs no double buffering; only the host
// staging has two slots.
extern "C" int fcv_batch_submit(fcv_batch *b, int slot, const int *frames_valid) {
    if (!b || slot < 0 || slot > 1) return fail(FCV_E_PARAM, "bad slot");
    if (!b->hout) return fail(FCV_E_STATE, "batch has no host staging");
    CU_TRY(cudaSetDevice(b->f->device));
    if (slot == 1) { int rc = ensure_slot1(b); if (rc) return rc; }
    int rc = fcv_batch_wait(b, slot);  // the slot's previous submit must have drained
    if (rc) return rc;
    unsigned char *hin = slot ? b->hin1 : b->hin, *hout = slot ? b->hout1 : b->hout;
    int *hfv = slot ? b->hfv1 : b->hfv, *dfv = slot ? b->dfv1 : b->dfv;
    if (frames_valid) {
        for (int s = 0; s < b->B; s++) {
            const int v = frames_valid[s];
            if (v < 0 || v > b->T * b->f->fragm) return fail(FCV_E_PARAM, "frames_valid[%d] = %d out of range", s, v);
            hfv[s] = v;
        }
    }
    const int nchunk = submit_chunks(b);
    if (!b->hbmax[slot]) {
        CU_TRY(cudaHostAlloc((void **)&b->hbmax[slot], (size_t)b->B * b->T * sizeof(float), cudaHostAllocDefault));
        memset(b->hbmax[slot], 0, (size_t)b->B * b->T * sizeof(float));
    }
    static const bool env_nokernels = getenv("FCV_COPY_ONLY") != nullptr;  // diagnostic: copies without kernels
    const int per = (b->B + nchunk - 1) / nchunk;
    for (int c = 0, off = 0; off < b->B; c++, off += per) {
        const int cnt = (b->B - off) < per ? (b->B - off) : per;
        cudaStream_t q = b->q[nchunk == 1 ? 0 : c % fcv_batch::NQ];
        if (frames_valid)
            CU_TRY(cudaMemcpyAsync(dfv + off, hfv + off, (size_t)cnt * sizeof(int), cudaMemcpyHostToDevice, q));
        CU_TRY(cudaMemcpyAsync(b->din + (size_t)off * b->in_block, hin + (size_t)off * b->in_block,
                               (size_t)cnt * b->in_block, cudaMemcpyHostToDevice, q));
        if (!env_nokernels) rc = run_kernels(b, off, cnt, frames_valid ? dfv : nullptr, q, nchunk == 1 ? prof_events(b) : nullptr);
        if (rc) return rc;
        CU_TRY(cudaMemcpyAsync(hout + (size_t)off * b->out_block, b->dout + (size_t)off * b->out_block,
                               (size_t)cnt * b->out_block, cudaMemcpyDeviceToHost, q));
        CU_TRY(cudaMemcpyAsync(b->hbmax[slot] + (size_t)off * b->T, b->bmax + (size_t)off * b->T,
                               (size_t)cnt * b->T * sizeof(float), cudaMemcpyDeviceToHost, q));
    }
    for (int i = 0; i < fcv_batch::NQ; i++) {
        if (!b->slot_done[slot][i]) CU_TRY(cudaEventCreateWithFlags(&b->slot_done[slot][i], cudaEventDisableTiming));
        CU_TRY(cudaEventRecord(b->slot_done[slot][i], b->q[i]));
    }
    b->slot_busy[slot] = true;
    b->step++;
    return 0;
}

extern "C" int fcv_batch_process(fcv_batch *b, const int *frames_valid) {
    int rc = fcv_batch_submit(b, 0, frames_valid);
    if (rc) return rc;
    return fcv_batch_wait(b, 0);
}

extern "C" int fcv_batch_reset_slot(fcv_batch *b, int slot) {
    if (!b || slot < 0 || slot >= b->B) return fail(FCV_E_PARAM, "bad slot");
    CU_TRY(cudaSetDevice(b->f->device));
    const fcv_filter *f = b->f;
    const size_t N = (size_t)f->fragm;
    // chunks of a batch run on several CUDA streams: order the reset after everything
    // already enqueued and before anything enqueued later
    int rc = fcv_batch_sync(b);
    if (rc) return rc;
    CU_TRY(cudaMemsetAsync(b->xring + (size_t)slot * f->ninp * b->R * N, 0, b->state_bytes_per_stream, b->q[0]));
    CU_TRY(cudaMemsetAsync(b->tail + (size_t)slot * f->nout * N, 0, (size_t)f->nout * N * sizeof(float), b->q[0]));
    CU_TRY(cudaMemsetAsync(b->maxv + slot, 0, sizeof(float), b->q[0]));
    CU_TRY(cudaStreamSynchronize(b->q[0]));
    return 0;
}

extern "C" int fcv_batch_reset_slot_async(fcv_batch *b, int slot) {
    if (!b || slot < 0 || slot >= b->B) return fail(FCV_E_PARAM, "bad slot");
    CU_TRY(cudaSetDevice(b->f->device));
    const fcv_filter *f = b->f;
    const size_t N = (size_t)f->fragm;
    // the CUDA stream every submit processes this stream of the batch on: the reset lands behind
    // the steps already submitted and ahead of the next one
    cudaStream_t q = submit_stream_of(b, slot);
    CU_TRY(cudaMemsetAsync(b->xring + (size_t)slot * f->ninp * b->R * N, 0, b->state_bytes_per_stream, q));
    CU_TRY(cudaMemsetAsync(b->tail + (size_t)slot * f->nout * N, 0, (size_t)f->nout * N * sizeof(float), q));
    CU_TRY(cudaMemsetAsync(b->maxv + slot, 0, sizeof(float), q));
    return 0;
}

extern "C" const float *fcv_batch_host_block_max_slot(fcv_batch *b, int slot) {
    if (!b || slot < 0 || slot > 1) return nullptr;
    return b->hbmax[slot];
}

extern "C" int fcv_batch_get_max(fcv_batch *b, float *max_out) {
    if (!b || !max_out) return fail(FCV_E_PARAM, "null argument");
    CU_TRY(cudaSetDevice(b->f->device));
    int rc = fcv_batch_sync(b);
    if (rc) return rc;
    CU_TRY(cudaMemcpy(max_out, b->maxv, (size_t)b->B * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int fcv_batch_get_block_max(fcv_batch *b, float *max_out) {
    if (!b || !max_out) return fail(FCV_E_PARAM, "null argument");
    CU_TRY(cudaSetDevice(b->f->device));
    int rc = fcv_batch_sync(b);
    if (rc) return rc;
    CU_TRY(cudaMemcpy(max_out, b->bmax, (size_t)b->B * b->T * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int fcv_batch_event_record(fcv_batch *b, int slot) {
    if (!b || slot < 0 || slot >= 16) return fail(FCV_E_PARAM, "bad event slot");
    CU_TRY(cudaSetDevice(b->f->device));
    if (!b->sw[slot]) CU_TRY(cudaEventCreate(&b->sw[slot]));
    CU_TRY(cudaEventRecord(b->sw[slot], b->q[0]));
    return 0;
}

extern "C" int fcv_batch_event_elapsed_ms(fcv_batch *b, int slot0, int slot1, float *ms) {
    if (!b || !ms || slot0 < 0 || slot0 >= 16 || slot1 < 0 || slot1 >= 16 || !b->sw[slot0] || !b->sw[slot1])
        return fail(FCV_E_PARAM, "bad event slot");
    CU_TRY(cudaSetDevice(b->f->device));
    CU_TRY(cudaEventSynchronize(b->sw[slot1]));
    CU_TRY(cudaEventElapsedTime(ms, b->sw[slot0], b->sw[slot1]));
    return 0;
}

extern "C" int fcv_batch_set_profiling(fcv_batch *b, int on) {
    if (!b) return fail(FCV_E_PARAM, "null batch");
    b->profiling = on != 0;
    b->ev_used = 0;
    b->prof_steps = 0;
    return 0;
}

extern "C" int fcv_batch_profile(fcv_batch *b, float ms[3], int *steps) {
    if (!b || !ms) return fail(FCV_E_PARAM, "null argument");
    int rc = fcv_batch_sync(b);
    if (rc) return rc;
    ms[0] = ms[1] = ms[2] = 0.f;
    for (size_t i = 0; i + 3 < b->ev_used; i += 4) {
        for (int k = 0; k < 3; k++) {
            float t = 0.f;
            CU_TRY(cudaEventElapsedTime(&t, b->ev[i + k], b->ev[i + k + 1]));
            ms[k] += t;
        }
    }
    if (steps) *steps = b->prof_steps;
    b->ev_used = 0;
    b->prof_steps = 0;
    return 0;
}

// ---------------------------------------------------------------------------
// single stream == batch of one with a shared in/out host block
// ---------------------------------------------------------------------------
struct fcv_stream {
    fcv_batch *b = nullptr;
};

extern "C" fcv_stream *fcv_stream_create(fcv_filter *f) {
    fcv_batch *b = batch_create(f, 1, FCV_PCM_F32, FCV_PCM_F32, true, 1);
    if (!b) return nullptr;
    fcv_stream *s = new (std::nothrow) fcv_stream();
    if (!s) { batch_free(b); fail(FCV_E_ALLOC, "out of memory"); return nullptr; }
    s->b = b;
    return s;
}

extern "C" void fcv_stream_destroy(fcv_stream *s) {
    if (!s) return;
    batch_free(s->b);
    delete s;
}

extern "C" int fcv_stream_reset(fcv_stream *s) {
    if (!s) return fail(FCV_E_PARAM, "null stream");
    return fcv_batch_reset_slot(s->b, 0);
}

extern "C" float *fcv_stream_buffer(fcv_stream *s) { return s ? (float *)s->b->hin : nullptr; }
extern "C" fcv_filter *fcv_stream_filter(fcv_stream *s) { return s ? s->b->f : nullptr; }

extern "C" int fcv_stream_process(fcv_stream *s, int frames_valid, float *max_inout) {
    if (!s) return fail(FCV_E_PARAM, "null stream");
    fcv_batch *b = s->b;
    const fcv_filter *f = b->f;
    if (frames_valid < 0 || frames_valid > f->fragm) return fail(FCV_E_PARAM, "frames_valid out of range");
    CU_TRY(cudaSetDevice(f->device));
    cudaStream_t q = b->q[0];
    if (frames_valid > 0 && !b->in_zero_copy)
        CU_TRY(cudaMemcpyAsync(b->din, b->hin, (size_t)frames_valid * f->ninp * sizeof(float), cudaMemcpyHostToDevice, q));
    // the valid-frame count of the one stream travels as a kernel argument
    int rc = run_kernels(b, 0, 1, nullptr, q, nullptr, frames_valid);
    if (rc) return rc;
    // one copy brings back the whole output block and the running maximum behind it
    CU_TRY(cudaMemcpyAsync(b->hin, b->dout, b->out_block + sizeof(float), cudaMemcpyDeviceToHost, q));
    b->step++;
    CU_TRY(cudaStreamSynchronize(q));
    if (max_inout) {
        float m;
        memcpy(&m, b->hin + b->out_block, sizeof(float));
        if (m > *max_inout) *max_inout = m;
    }
    return 0;
}

extern "C" int fcv_stream_get_input_spectrum(fcv_stream *s, int inp, int age, float *dst) {
    if (!s || !dst) return fail(FCV_E_PARAM, "null argument");
    fcv_batch *b = s->b;
    const fcv_filter *f = b->f;
    if (inp < 0 || inp >= f->ninp || age < 0 || age >= f->ring || (unsigned long long)age >= b->step)
        return fail(FCV_E_PARAM, "bad index");
    CU_TRY(cudaSetDevice(f->device));
    const int slot = (int)((b->step - 1 - age) % (unsigned long long)b->R);  // single streams have T == 1
    std::vector<float2> h((size_t)f->fragm);
    CU_TRY(cudaMemcpy(h.data(), b->xring + (size_t)(inp * b->R + slot) * f->fragm, h.size() * sizeof(float2),
                      cudaMemcpyDeviceToHost));
    unpermute_row(f->log2n, h.data(), dst);
    return 0;
}
