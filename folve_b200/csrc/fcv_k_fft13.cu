// fcv_k_fft13.cu -- kernels and launches of the fragm = 8192 transforms (fcv_fft13.cuh): every
// filter longer than 4096 taps (/root/reference/zita-fconfig.cc:74-77).  Role in the reference:
// FFTW's r2c / c2r inside zita-convolver's Convlevel::process(), reached from
// SoundProcessor::Process() (/root/reference/sound-processor.cc:98-127), with the
// (de)interleave, zero padding, overlap-add, int/float conversion and maximum fused in.
#include <cmath>
#include <map>

#include "fcv_internal.h"
#include "fcv_fft13.cuh"

using namespace fcv;

#ifndef F13_INV_SPLIT
#define F13_INV_SPLIT 0   // experiment: per-half barriers between passes C^-1 and B of the inverse kernel
#endif
#ifndef F13_INV_PFTW
#define F13_INV_PFTW 0    // experiment (with F13_INV_SPLIT): L1 prefetch of the pass-A twiddles
#endif
#ifndef F13_INV_NT
#define F13_INV_NT 256   // threads of the inverse kernel: 256 (2 CTAs/SM, <= 128 registers) or 128 (3 CTAs/SM)
#endif
constexpr int f13_min_ctas(int nt) { return nt >= 256 ? 2 : 3; }

// Forward transform of the current block: one CTA = one half (blockIdx.x & 1) of the
// spectra of C consecutive input channels (blockIdx.x >> 1 = channel group) of one
// (stream, block): PCM and twiddles are fetched once for C transforms.
// (One CTA computing both halves one after the other -- half as many CTAs, the second half's PCM
// loads served by L1 / L2 -- was measured 20 % slower, profiles/r02_experiments.md.)
template <class SEL, int FMT, int NCH, int C>
__global__ void __launch_bounds__(128 * C, f13_min_ctas(128 * C))
fwd13_stream_kernel(const __grid_constant__ SEL sel, f13::Tables tb, int ninp, int R, int T, int reset_max) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c2 *sm = reinterpret_cast<c2 *>(smem_raw);
    constexpr int N = f13::N;
    pdl_trigger();
    pdl_wait();
    const int h = blockIdx.x & 1, ch0 = (blockIdx.x >> 1) * C, b = blockIdx.y, bt = blockIdx.z;
    int frames = sel.frames(b) - bt * N;
    frames = frames < 0 ? 0 : (frames > N ? N : frames);
    int slot = sel.slot(b) + bt;
    if (slot >= R) slot -= R;
    float2 *const xring = sel.xring(b);
    float2 *rows[C];
#pragma unroll
    for (int c = 0; c < C; c++) rows[c] = xring + (size_t)((ch0 + c) * R + slot) * N;
    if (blockIdx.x == 0 && threadIdx.x == 0) {   // the only thread that needs the stream's descriptor
        const StreamDev s = sel.stream(b);
        // per-block maximum mode: the inverse kernel of this block starts from zero
        if ((reset_max & 1) && bt == 0) *s.maxv = 0.0f;
        s.bmax[bt] = 0.0f;  // this block's maximum starts from zero
    }
    if (frames == 0) {  // silence: its spectrum is zero
#pragma unroll
        for (int c = 0; c < C; c++)
            for (int e = threadIdx.x; e < f13::Q; e += 128 * C) rows[c][h * f13::Q + e] = make_float2(0.f, 0.f);
        return;
    }
    const size_t wire = FMT == PCM_S16 ? 2 : 4;
    const void *in = reinterpret_cast<const char *>(sel.din(b)) + (size_t)bt * N * ninp * wire;
    const bool fp = NCH == 2 && C == 2 && (reset_max & 2);   // stereo: FCV_FWD_FULL (launch_fwd13_fmt)
    if (h == 0) f13::fwd_half<0, FMT, NCH, C, 128 * C>(sm, tb, in, ninp, ch0, frames, rows, fp);
    else f13::fwd_half<1, FMT, NCH, C, 128 * C>(sm, tb, in, ninp, ch0, frames, rows, fp);
}

// Stereo blocks of a batch with several blocks per step: one CTA = one half of both channels' spectra of one
// stream for all T blocks, the thread's 47 twiddles in tensor memory (f13::fwd_half_tm).  Bit-identical and measured 6 % SLOWER than
// one CTA per block (0.417 against 0.394 ms, profiles/r02_experiments.md): an experiment, FCV_FWD_TMEM=1 only.
template <class SEL, int FMT>
__global__ void __launch_bounds__(256, 2)
fwd13_tm_kernel(const __grid_constant__ SEL sel, f13::Tables tb, int R, int T, int reset_max) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c2 *sm = reinterpret_cast<c2 *>(smem_raw);
    __shared__ uint32_t tm_slot;
    constexpr int N = f13::N;
    pdl_trigger();
    pdl_wait();
    const int tid = threadIdx.x;
    const int h = blockIdx.x, b = blockIdx.y;   // gridDim.x == 2
    if (tid < 32) f13::tm::alloc(&tm_slot, f13::tm::COLS);
    f13::tm::fence_before_sync();
    __syncthreads();
    f13::tm::fence_after_sync();
    const uint32_t tmem = f13::tm::thread_base(tm_slot);
    if (h == 0) f13::fwd_tm_fill<0>(tmem, tb);
    else f13::fwd_tm_fill<1>(tmem, tb);
    const int fvb = sel.frames(b), slot0 = sel.slot(b);
    float2 *const xring = sel.xring(b);
    const size_t wire = FMT == PCM_S16 ? 2 : 4;
    if (h == 0 && tid == 0) {   // the only thread that needs the stream's descriptor
        const StreamDev s = sel.stream(b);
        if (reset_max) *s.maxv = 0.0f;   // per-block maximum mode: the inverse kernel of this step starts from zero
        for (int bt = 0; bt < T; bt++) s.bmax[bt] = 0.0f;   // every block's maximum starts from zero
    }
    for (int bt = 0; bt < T; bt++) {
        int frames = fvb - bt * N;
        frames = frames < 0 ? 0 : (frames > N ? N : frames);
        int slot = slot0 + bt;
        if (slot >= R) slot -= R;
        float2 *rows[2] = {xring + (size_t)(0 * R + slot) * N, xring + (size_t)(1 * R + slot) * N};
        if (frames == 0) {  // silence: its spectrum is zero
#pragma unroll
            for (int c = 0; c < 2; c++)
                for (int e = tid; e < f13::Q; e += 256) rows[c][h * f13::Q + e] = make_float2(0.f, 0.f);
            continue;
        }
        const void *in = reinterpret_cast<const char *>(sel.din(b)) + (size_t)bt * N * 2 * wire;
        if (h == 0) f13::fwd_half_tm<0, FMT>(sm, tmem, in, frames, rows);
        else f13::fwd_half_tm<1, FMT>(sm, tmem, in, frames, rows);
        __syncthreads();   // shared memory is written again by the next block's pass A
    }
    f13::tm::fence_before_sync();
    __syncthreads();
    if (tid < 32) {
        f13::tm::fence_after_sync();
        f13::tm::dealloc(tm_slot, f13::tm::COLS);
    }
}

// The same for the stereo blocks of the per-file path as a cluster of the two halves' CTAs that reads the
// caller's block ONCE (f13::stage_half / fwd_pass_a_staged): half the bytes over the link.
template <class SEL, int FMT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
fwd13_pair_kernel(const __grid_constant__ SEL sel, f13::Tables tb, int R, int T, int reset_max) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c2 *sm = reinterpret_cast<c2 *>(smem_raw);
    unsigned char *stage = smem_raw + 2 * f13::HALF_BYTES;
    constexpr int N = f13::N;
    pdl_trigger();
    pdl_wait();
    const int h = blockIdx.x, b = blockIdx.y, bt = blockIdx.z;   // gridDim.x == 2: h is the rank in the cluster
    int frames = sel.frames(b) - bt * N;
    frames = frames < 0 ? 0 : (frames > N ? N : frames);
    int slot = sel.slot(b) + bt;
    if (slot >= R) slot -= R;
    float2 *const xring = sel.xring(b);
    float2 *rows[2] = {xring + (size_t)(0 * R + slot) * N, xring + (size_t)(1 * R + slot) * N};
    if (h == 0 && threadIdx.x == 0) {
        const StreamDev s = sel.stream(b);
        if (reset_max && bt == 0) *s.maxv = 0.0f;
        s.bmax[bt] = 0.0f;
    }
    if (frames == 0) {  // silence (both CTAs of the pair): its spectrum is zero
#pragma unroll
        for (int c = 0; c < 2; c++)
            for (int e = threadIdx.x; e < f13::Q; e += 256) rows[c][h * f13::Q + e] = make_float2(0.f, 0.f);
        return;
    }
    const size_t wire = FMT == PCM_S16 ? 2 : 4;
    const void *in = reinterpret_cast<const char *>(sel.din(b)) + (size_t)bt * N * 2 * wire;
    f13::stage_half<FMT>(stage, in, h, frames);
    f13::cluster_arrive();
    f13::cluster_wait();   // both halves of the block are staged
    const uint32_t st0 = f13::peer_smem(stage, 0u), st1 = f13::peer_smem(stage, 1u);
    if (h == 0) f13::fwd_half_staged<0, FMT>(sm, tb, st0, st1, frames, rows);
    else f13::fwd_half_staged<1, FMT>(sm, tb, st0, st1, frames, rows);
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

// Signed maximum (>= 0) of one block over all output channels: warp reduction, then one
// atomic per warp (positive floats order like their bit patterns).
__device__ __forceinline__ void block_max_update13(float *dst, float m) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(reinterpret_cast<int *>(dst), __float_as_int(m));
}

// Inverse transform of every (stream, output channel), T blocks one after the other
// (block t+1 overlap-adds the tail block t just saved), with overlap-add, tail save,
// re-interleave, float -> PCM and the signed maximum of the valid frames fused in.
// TM (batches, T > 1, 256 threads): the thread's twiddles and the overlap tail live in tensor memory for
// the T blocks (f13::tm, fcv_fft13.cuh); FCV_INV_TMEM=0 selects the kernel without it.
template <class SEL, int FMT, bool PF, bool TM = false>
__global__ void __launch_bounds__(F13_INV_NT, f13_min_ctas(F13_INV_NT))
inv13_stream_kernel(const __grid_constant__ SEL sel, f13::Tables tb, int nout_flags, int T) {
    const int nout = nout_flags & 0xffff;          // bit 16: whole blocks skip the per-sample test of the maximum (TM)
    const bool full_path = (nout_flags >> 16) != 0;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c2 *sm = reinterpret_cast<c2 *>(smem_raw);
    constexpr int NT = F13_INV_NT;
    static_assert(!TM || NT == 256, "tensor-memory variant: one column per thread");
    __shared__ float red[NT / 32];
    __shared__ uint32_t tm_slot;
    constexpr int N = f13::N, M = N;
    pdl_trigger();
    pdl_wait();
    const int tid = threadIdx.x;
    const int o = blockIdx.x, b = blockIdx.y;
    const StreamDev s = sel.stream(b);
    const int fvb = sel.frames(b);
    float2 *tail = reinterpret_cast<float2 *>(s.tail + (size_t)o * N);
    const size_t wire = FMT == PCM_S16 ? 2 : 4;
    float lmax = 0.0f;
    uint32_t tmem = 0;
    if (TM) {
        if (tid < 32) f13::tm::alloc(&tm_slot, f13::tm::COLS);
        f13::tm::fence_before_sync();
        __syncthreads();
        f13::tm::fence_after_sync();
        tmem = f13::tm::thread_base(tm_slot);
        f13::inv_tm_fill(tmem, tb, tail);
    }

    for (int bt = 0; bt < T; bt++) {
        int frames = fvb - bt * N;
        frames = frames < 0 ? 0 : (frames > N ? N : frames);
        // entry 0 of the sequence to transform comes from the two real bins (DC / Nyquist products)
        const float2 z0 = tid == 0 ? s.zc0[(size_t)o * T + bt] : make_float2(0.f, 0.f);
        const float2 *yrow = s.Y + ((size_t)o * T + bt) * M;
        if (PF && bt + 1 < T) {  // the next block's spectrum row (64 KB) is requested into L2 now
#pragma unroll
            for (int i = 0; i < (M * 8 / 128) / NT; i++)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(yrow + M + (size_t)(tid + i * NT) * 16));
        }
#if F13_INV_SPLIT
        static_assert(NT == 256, "two halves of 128 threads");
        {   // each half: C^-1 -> own 128-thread barrier -> B; only pass A^-1 needs both halves
            const int half = tid >> 7, ht = tid & 127;
            if (half == 0) f13::inv_pass_c<0>(sm, tb, yrow, c2_pack(z0.x, z0.y), ht);
            else f13::inv_pass_c<1>(sm + f13::HALF_ELEMS, tb, yrow, 0ull, ht);
#if F13_INV_PFTW
            // the 31 pass-A twiddles of this thread's column: into L1 while passes C and B run
            if ((tid & 15) == 0) {
#pragma unroll
                for (int k0 = 1; k0 < 16; k0++) asm volatile("prefetch.global.L1 [%0];" ::"l"(tb.twA0 + (k0 - 1) * 256 + tid));
#pragma unroll
                for (int k0 = 0; k0 < 16; k0++) asm volatile("prefetch.global.L1 [%0];" ::"l"(tb.twA1 + k0 * 256 + tid));
            }
#endif
            if (half == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
            else asm volatile("bar.sync 2, 128;" ::: "memory");
            f13::pass_b_one<+1, 128>(sm + half * f13::HALF_ELEMS, tb, ht);
        }
        __syncthreads();
#else
        if (TM) {
            uint32_t wr[32];
            f13::tm::ld_issue32(tmem + f13::tm::TWC, wr);   // waited for inside, behind the spectrum loads
            if (tid < 128) f13::inv_pass_c_w<0>(sm, yrow, c2_pack(z0.x, z0.y), tid, wr);
            else f13::inv_pass_c_w<1>(sm + f13::HALF_ELEMS, yrow, 0ull, tid - 128, wr);
        } else {
#pragma unroll 1
            for (int j = tid; j < 256; j += NT) {
                if (j < 128) f13::inv_pass_c<0>(sm, tb, yrow, c2_pack(z0.x, z0.y), j);
                else f13::inv_pass_c<1>(sm + f13::HALF_ELEMS, tb, yrow, 0ull, j - 128);
            }
        }
        __syncthreads();
        f13::pass_b<+1, 2, NT>(sm, tb);
        __syncthreads();
#endif
        void *dout = reinterpret_cast<char *>(s.dout) + (size_t)bt * N * nout * wire;
        const float m = TM ? f13::inv_pass_a_tm<FMT>(sm, tmem, dout, nout, o, frames, full_path)
                           : f13::inv_pass_a<FMT, NT>(sm, tb, tail, dout, nout, o, frames);
        lmax = fmaxf(lmax, m);
        block_max_update13(s.bmax + bt, m);
        if (bt + 1 < T) __syncthreads();  // shared memory and the tail are reused by the next block
    }
    if (TM) f13::inv_tm_save_tail(tmem, tail);   // the last block's tail: the next step starts from it
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, d));
    if ((tid & 31) == 0) red[tid >> 5] = lmax;
    if (TM) f13::tm::fence_before_sync();
    __syncthreads();
    if (TM && tid < 32) {   // every warp's tensor-memory accesses are complete (their loads wait inside)
        f13::tm::fence_after_sync();
        f13::tm::dealloc(tm_slot, f13::tm::COLS);
    }
    if (tid == 0) {
        for (int w = 1; w < NT / 32; w++) lmax = fmaxf(lmax, red[w]);
        // running maximum is >= 0, positive floats order like their bit patterns
        if (lmax > 0.0f) atomicMax(reinterpret_cast<int *>(s.maxv), __float_as_int(lmax));
    }
    if (SEL::kSingle && s.hout) {
        const int fr = fvb < 0 ? 0 : (fvb > N ? N : fvb);
        host_copy_out(s, nout, (size_t)fr * nout * wire, sel.seq(b), tid, NT);
    }
}

// ---- tables ---------------------------------------------------------------------------------
// The same for blocks of exactly two output channels: the two channels' CTAs form a cluster and
// write the interleaved PCM together (f13::inv_pass_a_pair).  256 threads.
template <class SEL, int FMT, bool PF>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 2)
inv13_pair_kernel(const __grid_constant__ SEL sel, f13::Tables tb, int T) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c2 *sm = reinterpret_cast<c2 *>(smem_raw);
    constexpr int NT = 256;
    __shared__ float red[NT / 32];
    constexpr int N = f13::N, M = N;
    pdl_trigger();
    pdl_wait();
    const int tid = threadIdx.x;
    const int o = blockIdx.x, b = blockIdx.y;   // gridDim.x == 2: o is the rank in the cluster
    const StreamDev s = sel.stream(b);
    const int fvb = sel.frames(b);
    float2 *tail = reinterpret_cast<float2 *>(s.tail + (size_t)o * N);
    const size_t wire = FMT == PCM_S16 ? 2 : 4;
    float lmax = 0.0f;

    for (int bt = 0; bt < T; bt++) {
        int frames = fvb - bt * N;
        frames = frames < 0 ? 0 : (frames > N ? N : frames);
        const float2 z0 = tid == 0 ? s.zc0[(size_t)o * T + bt] : make_float2(0.f, 0.f);
        const float2 *yrow = s.Y + ((size_t)o * T + bt) * M;
        if (PF && bt + 1 < T) {  // the next block's spectrum row (64 KB) is requested into L2 now
#pragma unroll
            for (int i = 0; i < (M * 8 / 128) / NT; i++)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(yrow + M + (size_t)(tid + i * NT) * 16));
        }
        if (bt == 0) {
            if (tid < 128) f13::inv_pass_c<0>(sm, tb, yrow, c2_pack(z0.x, z0.y), tid);
            else f13::inv_pass_c<1>(sm + f13::HALF_ELEMS, tb, yrow, 0ull, tid - 128);
        } else {   // waits for the pair's second barrier of the previous block before writing shared memory
            if (tid < 128) f13::inv_pass_c<0, true>(sm, tb, yrow, c2_pack(z0.x, z0.y), tid);
            else f13::inv_pass_c<1, true>(sm + f13::HALF_ELEMS, tb, yrow, 0ull, tid - 128);
        }
        __syncthreads();
        f13::pass_b<+1, 2, NT>(sm, tb);
        __syncthreads();
        void *dout = reinterpret_cast<char *>(s.dout) + (size_t)bt * N * 2 * wire;
        float m;
        void *hout = SEL::kSingle ? s.hout : nullptr;   // per-file path: the vectors go to the caller's block as well
        f13::inv_pass_a_pair<FMT>(sm, tb, tail, dout, hout, o, frames, m, [&]() { block_max_update13(s.bmax + bt, m); });
        lmax = fmaxf(lmax, m);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, d));
    if ((tid & 31) == 0) red[tid >> 5] = lmax;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < NT / 32; w++) lmax = fmaxf(lmax, red[w]);
        if (lmax > 0.0f) atomicMax(reinterpret_cast<int *>(s.maxv), __float_as_int(lmax));
    }
    // The pair's last barrier: the other CTA is done with this CTA's shared memory -- and, on the per-file
    // path, both CTAs' stores to the caller's block and both maxima are ordered before what follows.
    f13::cluster_wait_divergent();
    if (SEL::kSingle && s.hout) {
        // one more rendezvous behind the maximum's atomics (the wait above belongs to the arrive inside the epilogue)
        f13::cluster_arrive();
        f13::cluster_wait();
        if (o == 0 && tid == 0) {
            *s.hmax = __ldcg(s.maxv);
            __threadfence_system();   // cumulative: every store of both CTAs before the word (see host_copy_out)
            *reinterpret_cast<volatile unsigned *>(s.hdone) = sel.seq(b);
        }
    }
}

template <class SEL, int FMT>
static int set_attrs13() {
    const int one = (int)f13::HALF_BYTES, two = 2 * (int)f13::HALF_BYTES;
    CU_TRY(cudaFuncSetAttribute(fwd13_stream_kernel<SEL, FMT, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, two));
    if constexpr (!SEL::kSingle)
        CU_TRY(cudaFuncSetAttribute(fwd13_tm_kernel<SEL, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, two));
    if constexpr (SEL::kSingle)
        CU_TRY(cudaFuncSetAttribute(fwd13_pair_kernel<SEL, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    two + (int)f13::STAGE_BYTES_MAX));
    CU_TRY(cudaFuncSetAttribute(fwd13_stream_kernel<SEL, FMT, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, one));
    CU_TRY(cudaFuncSetAttribute(fwd13_stream_kernel<SEL, FMT, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, one));
    CU_TRY(cudaFuncSetAttribute(fwd13_stream_kernel<SEL, FMT, 0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, two));
    CU_TRY(cudaFuncSetAttribute(inv13_stream_kernel<SEL, FMT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, two));
    CU_TRY(cudaFuncSetAttribute(inv13_stream_kernel<SEL, FMT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, two));
    if constexpr (!SEL::kSingle && F13_INV_NT == 256) {
        CU_TRY(cudaFuncSetAttribute(inv13_stream_kernel<SEL, FMT, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, two));
        CU_TRY(cudaFuncSetAttribute(inv13_stream_kernel<SEL, FMT, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, two));
    }
    {
        CU_TRY(cudaFuncSetAttribute(inv13_pair_kernel<SEL, FMT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, two));
        CU_TRY(cudaFuncSetAttribute(inv13_pair_kernel<SEL, FMT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, two));
    }
    return 0;
}

namespace {
struct Ctx13 {
    bool have = false;
    f13::Tables tab{};
};
std::mutex g_mu13;
std::map<int, Ctx13> g_ctx13;
}  // namespace

// Twiddle tables of the fragm = 8192 transforms, double precision on the host; once per device.
int fcv::fft13_tables(int device, f13::Tables *out) {
    std::lock_guard<std::mutex> l(g_mu13);
    Ctx13 &c = g_ctx13[device];
    if (!c.have) {
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
#if F13_TWU_FACTORED
        // base values of the unpack / repack twiddles: exp(-i pi (2c + h) / M), h = 0, 1, c = 0..255
        for (int hh = 0; hh < 2; hh++)
            for (int cc = 0; cc < 256; cc++) h[oU + (size_t)hh * 256 + cc] = unit((double)(2 * cc + hh) / (2.0 * M));
#else
        for (int e = 0; e < 2 * Q; e++) {
            const int k = 2 * (e & (Q - 1)) + (e >> (f13::LOG2N - 1));
            h[oU + e] = unit((double)k / (2.0 * M));
        }
#endif
        float2 *d = nullptr;
        CU_TRY(cudaMalloc(&d, h.size() * sizeof(float2)));
        CU_TRY(cudaMemcpy(d, h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice));
        c.tab.twA0 = d + o0;
        c.tab.twA1 = d + o1;
        c.tab.twB = d + oB;
        c.tab.twU = d + oU;
        int rc = set_attrs13<BatchSel, PCM_F32>();
        if (!rc) rc = set_attrs13<BatchSel, PCM_S16>();
        if (!rc) rc = set_attrs13<BatchSel, PCM_S24>();
        if (!rc) rc = set_attrs13<GroupSel, PCM_F32>();
        if (!rc) rc = set_attrs13<GroupSel, PCM_S16>();
        if (!rc) rc = set_attrs13<GroupSel, PCM_S24>();
        if (rc) return rc;
        CU_TRY(cudaFuncSetAttribute(fwd13_raw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f13::HALF_BYTES));
        c.have = true;
    }
    *out = c.tab;
    return 0;
}

void fcv::launch_filter_fft13(const fcv_filter *f, const float *dsrc, float2 *dst, int nrows) {
    fwd13_raw_kernel<<<dim3(2, nrows), 128, f13::HALF_BYTES>>>(dsrc, dst, f->tb13);
}

// Tensor-memory variants (f13::tm): the inverse kernel of a batch uses it by default (FCV_INV_TMEM=0: off), the
// forward kernel's stays an experiment (FCV_FWD_TMEM=1).  fcv_debug_set_tmem: bit 0 inverse, bit 1 forward (tests).
static std::atomic<bool> g_inv_tmem{!(getenv("FCV_INV_TMEM") && atoi(getenv("FCV_INV_TMEM")) == 0)};
static std::atomic<bool> g_fwd_tmem{getenv("FCV_FWD_TMEM") && atoi(getenv("FCV_FWD_TMEM")) != 0};
extern "C" void fcv_debug_set_tmem(int mask) {
    g_inv_tmem.store((mask & 1) != 0);
    g_fwd_tmem.store((mask & 2) != 0);
}
extern "C" int fcv_debug_get_tmem(void) { return (g_inv_tmem.load() ? 1 : 0) | (g_fwd_tmem.load() ? 2 : 0); }

// ---- launches -------------------------------------------------------------------------------
// Stereo and mono blocks take the vector-load kernels (stereo: both channels per CTA), any other
// channel count scalar loads: two channels per CTA when the count is even, else one.
template <class SEL, int FMT>
static void launch_fwd13_fmt(const StepArgs &a, const SEL &sel, cudaStream_t q) {
    const fcv_filter *f = a.f;
    const int rm = a.per_block_max ? 1 : 0;
    // per-file path, stereo: the two halves' CTAs as a cluster that reads the caller's block once (FCV_FWD_PAIR=0: off)
    static const bool fwd_pair = !(getenv("FCV_FWD_PAIR") && atoi(getenv("FCV_FWD_PAIR")) == 0);
    if constexpr (SEL::kSingle) {
        if (f->ninp == 2 && fwd_pair) {
            launch_k(fwd13_pair_kernel<SEL, FMT>, dim3(2, a.cnt, a.T), dim3(256), 2 * f13::HALF_BYTES + f13::STAGE_BYTES_MAX,
                     q, a.pdl, sel, f->tb13, a.R, a.T, rm);
            return;
        }
    }
    // batches of stereo blocks with several blocks per step: one CTA per (half, stream), twiddles in tensor memory
    // (experiment, off: slower)
    const bool fwd_tm = g_fwd_tmem.load(std::memory_order_relaxed);
    if constexpr (!SEL::kSingle) {
        if (f->ninp == 2 && a.T > 1 && fwd_tm) {
            fwd13_tm_kernel<SEL, FMT><<<dim3(2, a.cnt), 256, 2 * f13::HALF_BYTES, q>>>(sel, f->tb13, a.R, a.T, rm);
            return;
        }
    }
    // bit 1 of the last argument: whole blocks take pass A without the per-sample zeroing (FCV_FWD_FULL=0: off)
    static const int fwd_full = getenv("FCV_FWD_FULL") ? atoi(getenv("FCV_FWD_FULL")) : 1;
    if (f->ninp == 2)
        launch_k(fwd13_stream_kernel<SEL, FMT, 2, 2>, dim3(2, a.cnt, a.T), dim3(256), 2 * f13::HALF_BYTES, q, a.pdl, sel,
                 f->tb13, f->ninp, a.R, a.T, rm | (fwd_full ? 2 : 0));
    else if (f->ninp == 1)
        launch_k(fwd13_stream_kernel<SEL, FMT, 1, 1>, dim3(2, a.cnt, a.T), dim3(128), f13::HALF_BYTES, q, a.pdl, sel,
                 f->tb13, f->ninp, a.R, a.T, rm);
    else if (f->ninp % 2 == 0)   // 4, 6, 8 ... channels: pairs of channels share a CTA's twiddle loads like a stereo block
        launch_k(fwd13_stream_kernel<SEL, FMT, 0, 2>, dim3(f->ninp, a.cnt, a.T), dim3(256), 2 * f13::HALF_BYTES, q, a.pdl,
                 sel, f->tb13, f->ninp, a.R, a.T, rm);
    else
        launch_k(fwd13_stream_kernel<SEL, FMT, 0, 1>, dim3(2 * f->ninp, a.cnt, a.T), dim3(128), f13::HALF_BYTES, q, a.pdl,
                 sel, f->tb13, f->ninp, a.R, a.T, rm);
}
template <class SEL>
static void launch_fwd13_sel(const StepArgs &a, const SEL &sel, cudaStream_t q) {
    if (a.in_fmt == PCM_F32) launch_fwd13_fmt<SEL, PCM_F32>(a, sel, q);
    else if (a.in_fmt == PCM_S16) launch_fwd13_fmt<SEL, PCM_S16>(a, sel, q);
    else launch_fwd13_fmt<SEL, PCM_S24>(a, sel, q);
}
void fcv::launch_fwd13(const StepArgs &a, cudaStream_t q) {
    if (a.grp) launch_fwd13_sel<GroupSel>(a, *a.grp, q);
    else launch_fwd13_sel<BatchSel>(a, a.bsel, q);
}

static std::atomic<bool> g_inv_pair{getenv("FCV_INV_PAIR") && atoi(getenv("FCV_INV_PAIR")) != 0};
static std::atomic<bool> g_inv_pair_single{getenv("FCV_INV_PAIR_SINGLE") && atoi(getenv("FCV_INV_PAIR_SINGLE")) != 0};
extern "C" void fcv_debug_set_inv_pair(int mask) {   // bit 0: batches, bit 1: the per-file path
    g_inv_pair.store((mask & 1) != 0);
    g_inv_pair_single.store((mask & 2) != 0);
}

template <class SEL, int FMT>
static void launch_inv13_fmt(const StepArgs &a, const SEL &sel, cudaStream_t q) {
    const dim3 grid(a.f->nout, a.cnt);
    const size_t smem = 2 * f13::HALF_BYTES;
    static const bool pf = !(getenv("FCV_INV_PF") && atoi(getenv("FCV_INV_PF")) == 0);
    // Stereo blocks: the two channels' CTAs as a cluster that writes whole interleaved frames.
    // An experiment that stays OFF on both paths (profiles/r02_experiments.md):
    //   batches (FCV_INV_PAIR=1 / fcv_debug_set_inv_pair(1)): 4 % slower;
    //   per-file path (FCV_INV_PAIR_SINGLE=1 / fcv_debug_set_inv_pair(2)): the frames go straight to the caller's
    //     pinned block from both CTAs, no second pass by the last CTA: 14.6 -> 14.0 us for a lone block, but
    //     16 / 32 concurrent callers lose 5-10 % (clusters of concurrent groups are harder to place).
    const bool pair = (SEL::kSingle ? g_inv_pair_single : g_inv_pair).load(std::memory_order_relaxed);
    if (pair && a.f->nout == 2 && F13_INV_NT == 256) {
        if (pf && a.T > 1) inv13_pair_kernel<SEL, FMT, true><<<grid, 256, smem, q>>>(sel, a.f->tb13, a.T);
        else launch_k(inv13_pair_kernel<SEL, FMT, false>, grid, dim3(256), smem, q, a.pdl, sel, a.f->tb13, a.T);
        return;
    }
    // batches with several blocks per step: twiddles and overlap tail in tensor memory (FCV_INV_TMEM=0: off)
    const bool use_tm = g_inv_tmem.load(std::memory_order_relaxed);
    if constexpr (!SEL::kSingle && F13_INV_NT == 256) {
        if (use_tm && a.T > 1) {
            // bit 16 of the channel-count argument: whole blocks take the epilogue without the per-sample test of the
            // maximum (FCV_INV_FULL=0: off)
            static const int inv_full = getenv("FCV_INV_FULL") ? atoi(getenv("FCV_INV_FULL")) : 1;
            const int nout_arg = a.f->nout | (inv_full ? 1 << 16 : 0);
            if (pf) inv13_stream_kernel<SEL, FMT, true, true><<<grid, 256, smem, q>>>(sel, a.f->tb13, nout_arg, a.T);
            else inv13_stream_kernel<SEL, FMT, false, true><<<grid, 256, smem, q>>>(sel, a.f->tb13, nout_arg, a.T);
            return;
        }
    }
    if (pf && a.T > 1) inv13_stream_kernel<SEL, FMT, true><<<grid, F13_INV_NT, smem, q>>>(sel, a.f->tb13, a.f->nout, a.T);
    else launch_k(inv13_stream_kernel<SEL, FMT, false>, grid, dim3(F13_INV_NT), smem, q, a.pdl, sel, a.f->tb13, a.f->nout, a.T);
}
template <class SEL>
static void launch_inv13_sel(const StepArgs &a, const SEL &sel, cudaStream_t q) {
    if (a.out_fmt == PCM_F32) launch_inv13_fmt<SEL, PCM_F32>(a, sel, q);
    else if (a.out_fmt == PCM_S16) launch_inv13_fmt<SEL, PCM_S16>(a, sel, q);
    else launch_inv13_fmt<SEL, PCM_S24>(a, sel, q);
}
void fcv::launch_inv13(const StepArgs &a, cudaStream_t q) {
    if (a.grp) launch_inv13_sel<GroupSel>(a, *a.grp, q);
    else launch_inv13_sel<BatchSel>(a, a.bsel, q);
}
