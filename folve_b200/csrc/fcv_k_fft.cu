// fcv_k_fft.cu -- kernels and launches of the any-size transforms (fcv_fft.cuh), block sizes
// 64 .. 8192 (fragm rule: /root/reference/zita-fconfig.cc:74-77; 8192 normally runs the kernels of
// fcv_k_fft13.cu instead).  Role in the reference: FFTW's r2c / c2r inside zita-convolver's
// Convlevel::process(), reached from SoundProcessor::Process()
// (/root/reference/sound-processor.cc:98-127).
#include <cmath>
#include <map>

#include "fcv_internal.h"
#include "fcv_fft.cuh"

using namespace fcv;

#ifndef FFT_MIN_CTAS
#define FFT_MIN_CTAS 2
#endif
#ifndef FWD_MIN_CTAS
#define FWD_MIN_CTAS 4
#endif

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
template <class SEL, int LOG2N>
__global__ void __launch_bounds__(fft_threads(LOG2N, 1), FWD_MIN_CTAS)
fwd_stream_kernel(const __grid_constant__ SEL sel, FftTables tb, int ninp, int R, int T, int in_fmt, int reset_max) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *sm = reinterpret_cast<float2 *>(smem_raw);
    constexpr int N = 1 << LOG2N, NT = fft_threads(LOG2N, 1);
    pdl_trigger();
    pdl_wait();
    // grid: x = 2 * input channel + half, y = stream, z = block of the step
    const int i = blockIdx.x >> 1, h = blockIdx.x & 1, b = blockIdx.y, bt = blockIdx.z;
    const StreamDev s = sel.stream(b);
    int frames = sel.frames(b) - bt * N;
    frames = frames < 0 ? 0 : (frames > N ? N : frames);
    int slot = sel.slot(b) + bt;
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
// MIX: add mix[frame * nout + o] (the tail level of a non-uniformly partitioned filter) to every sample.
template <int LOG2N, int FMT, bool MIX>
__device__ __forceinline__ float inv_epilogue(const float2 *sm, const FftTables &tb, float2 *__restrict__ tail,
                                              void *dout, int nout, int o, int frames, const float *__restrict__ mix) {
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
            float y0 = a.x + t.x + tl[i].x;
            float y1 = a.y + t.y + tl[i].y;
            tail[n] = make_float2(a.x - t.x, a.y - t.y);
            const int f0 = 2 * n;
            if (MIX && mix) {
                y0 += mix[(size_t)f0 * nout + o];
                y1 += mix[(size_t)(f0 + 1) * nout + o];
            }
            pcm_store<FMT>(dout, (size_t)f0 * nout + o, y0);
            pcm_store<FMT>(dout, (size_t)(f0 + 1) * nout + o, y1);
            if (f0 < frames) lmax = fmaxf(lmax, y0);
            if (f0 + 1 < frames) lmax = fmaxf(lmax, y1);
        }
    }
    return lmax;
}

// Inverse transform of every (stream, output channel) with overlap-add, tail save,
// re-interleave, float/int conversion and running signed maximum.  The T blocks of a step are
// done one after the other by the same CTA (block t+1 overlap-adds the tail block t just saved).
template <class SEL, int LOG2N>
__global__ void __launch_bounds__(fft_threads(LOG2N), FFT_MIN_CTAS)
inv_stream_kernel(const __grid_constant__ SEL sel, FftTables tb, int nout, int T, int out_fmt) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *sm = reinterpret_cast<float2 *>(smem_raw);
    __shared__ float red[32];
    constexpr int N = 1 << LOG2N, M = N;
    constexpr int NT = fft_threads(LOG2N);
    pdl_trigger();
    pdl_wait();
    const int tid = threadIdx.x;
    const int o = blockIdx.x, b = blockIdx.y;
    const StreamDev s = sel.stream(b);
    const int fvb = sel.frames(b);
    float2 *tail = reinterpret_cast<float2 *>(s.tail + (size_t)o * N);
    float lmax = 0.0f;

    for (int bt = 0; bt < T; bt++) {
        int frames = fvb - bt * N;
        frames = frames < 0 ? 0 : (frames > N ? N : frames);
        inv_load<LOG2N>(sm, tb, s.Y + ((size_t)o * T + bt) * M);
        if (tid == 0) sm[0] = s.zc0[(size_t)o * T + bt];  // Zc[0] from the two real bins (DC / Nyquist products)
        __syncthreads();

        inv_body<LOG2N>(sm, tb);

        const size_t boff = (size_t)bt * N * nout;
        float m;
        const float *mix = sel.mix(b);
        if (out_fmt == PCM_F32) m = inv_epilogue<LOG2N, PCM_F32, SEL::kSingle>(sm, tb, tail, (float *)s.dout + boff, nout, o, frames, mix);
        else if (out_fmt == PCM_S16) m = inv_epilogue<LOG2N, PCM_S16, SEL::kSingle>(sm, tb, tail, (short *)s.dout + boff, nout, o, frames, mix);
        else m = inv_epilogue<LOG2N, PCM_S24, SEL::kSingle>(sm, tb, tail, (int *)s.dout + boff, nout, o, frames, mix);
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
    if (SEL::kSingle && s.hout) {
        const int fr = fvb < 0 ? 0 : (fvb > N ? N : fvb);
        host_copy_out(s, nout, (size_t)fr * nout * (out_fmt == PCM_S16 ? 2 : 4), sel.seq(b), tid, NT);
    }
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
        default: break; \
    }

template <int LOG2N>
static int set_attrs() {
    const int bytes = (int)fft_smem_bytes(LOG2N), one = (int)fft_smem_bytes(LOG2N, 1);
    CU_TRY(cudaFuncSetAttribute(fwd_stream_kernel<BatchSel, LOG2N>, cudaFuncAttributeMaxDynamicSharedMemorySize, one));
    CU_TRY(cudaFuncSetAttribute(fwd_stream_kernel<GroupSel, LOG2N>, cudaFuncAttributeMaxDynamicSharedMemorySize, one));
    CU_TRY(cudaFuncSetAttribute(fwd_raw_kernel<LOG2N>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CU_TRY(cudaFuncSetAttribute(inv_stream_kernel<BatchSel, LOG2N>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CU_TRY(cudaFuncSetAttribute(inv_stream_kernel<GroupSel, LOG2N>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    return 0;
}

namespace {
struct Ctx {
    bool have[16] = {};
    FftTables tab[16];
};
std::mutex g_mu;
std::map<int, Ctx> g_ctx;
}  // namespace

// Builds (once per device and size) the twiddle tables in double precision.
int fcv::fft_tables(int device, int log2n, FftTables *out) {
    if (log2n < 6 || log2n > 13) return fail(FCV_E_PARAM, "unsupported partition size 2^%d", log2n);
    std::lock_guard<std::mutex> l(g_mu);
    Ctx &c = g_ctx[device];
    if (!c.have[log2n]) {
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
        c.tab[log2n] = tb;
        int rc = 0;
        DISPATCH_LOG2N(log2n, rc = set_attrs<L>());
        if (rc) return rc;
        c.have[log2n] = true;
    }
    *out = c.tab[log2n];
    return 0;
}

// Entry-permuted spectrum layout of the generic kernels (test hooks un-permute rows on the host).
int fcv::fft_entry_of_bin(int log2n, int k) {
    const int q = log2n - 1;
    return ((k & 1) << q) + plan_rev(q, k >> 1);
}

int fcv::launch_filter_fft(const fcv_filter *f, const float *dsrc, float2 *dst, int nrows) {
    if (f->k13) {
        launch_filter_fft13(f, dsrc, dst, nrows);
    } else {
        const FftTables tb = f->tb;
        DISPATCH_LOG2N(f->log2n, (fwd_raw_kernel<L><<<nrows, fft_threads(L), fft_smem_bytes(L)>>>(dsrc, dst, tb)));
    }
    g_launches++;
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return fail(FCV_E_CUDA, "filter transform failed: %s", cudaGetErrorString(e));
    return 0;
}

template <class SEL>
static void launch_fwd_sel(const StepArgs &a, const SEL &sel, cudaStream_t q) {
    const fcv_filter *f = a.f;
    DISPATCH_LOG2N(f->log2n, (launch_k(fwd_stream_kernel<SEL, L>, dim3(2 * f->ninp, a.cnt, a.T), dim3(fft_threads(L, 1)),
                                       fft_smem_bytes(L, 1), q, a.pdl, sel, f->tb, f->ninp, a.R, a.T, a.in_fmt,
                                       a.per_block_max ? 1 : 0)));
}
void fcv::launch_fwd(const StepArgs &a, cudaStream_t q) {
    if (a.grp) launch_fwd_sel<GroupSel>(a, *a.grp, q);
    else launch_fwd_sel<BatchSel>(a, a.bsel, q);
}

template <class SEL>
static void launch_inv_sel(const StepArgs &a, const SEL &sel, cudaStream_t q) {
    const fcv_filter *f = a.f;
    DISPATCH_LOG2N(f->log2n, (launch_k(inv_stream_kernel<SEL, L>, dim3(f->nout, a.cnt), dim3(fft_threads(L)),
                                       fft_smem_bytes(L), q, a.pdl, sel, f->tb, f->nout, a.T, a.out_fmt)));
}
void fcv::launch_inv(const StepArgs &a, cudaStream_t q) {
    if (a.grp) launch_inv_sel<GroupSel>(a, *a.grp, q);
    else launch_inv_sel<BatchSel>(a, a.bsel, q);
}
