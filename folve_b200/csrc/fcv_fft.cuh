// fcv_fft.cuh -- shared-memory real FFT / inverse real FFT of one zero-padded
// partition, hand written for sm_100a.
//
// Role in the reference path: these two kernels are what zita-convolver's
// Convlevel::process() hands to FFTW (fftwf_execute_dft_r2c on the zero padded
// input block, fftwf_execute_dft_c2r on the accumulated spectrum), called from
// SoundProcessor::Process() at /root/reference/sound-processor.cc:113, with the
// de-interleave (sound-processor.cc:106-111), overlap-add, re-interleave and
// running maximum (sound-processor.cc:115-125) fused in.
//
// Method (ours, not FFTW's):
//   * A real transform of length 2N is done as a complex transform of length
//     M = N on z[n] = x[2n] + i x[2n+1].
//   * The block is zero padded (x[N..2N) == 0), so z[n] == 0 for n >= M/2 and
//     the first radix-2 DIF stage degenerates: even bins = FFT_Q(z), odd bins
//     = FFT_Q(z * w_M^n), Q = M/2.  The two length-Q transforms live side by
//     side in shared memory ("half 0" and "half 1").
//   * Each length-Q transform is an in-place decimation-in-frequency FFT with
//     radix-16/8/4 register butterflies and NO reordering pass: the spectrum
//     stays in digit-reversed order ("packed-permuted layout").  The complex
//     multiply-accumulate is element-wise, so the order is irrelevant to it,
//     and the inverse runs the same passes backwards (decimation in time).
//   * DC and Nyquist are both real; they share entry 0 (re = DC, im = Nyquist)
//     so that one spectrum is exactly M complex values = 8*N bytes.
//   * Shared memory is padded by one float2 every 16 so that the stride-R
//     accesses of the last pass are bank-conflict free.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fcv_c2.cuh"
#include "fcv_types.h"

namespace fcv {

// ---- compile-time plan ----------------------------------------------------
// Q = 2^q is the length of each half transform.  Radix exponents per pass,
// first pass in the low nibble.
#ifndef FCV_PLAN12
#define FCV_PLAN12 0x444  // 4096 = 16*16*16 (256 threads); 0x3333 = 8*8*8*8 (512 threads)
#endif
__host__ __device__ constexpr int plan_code(int q) {
    return q == 12 ? FCV_PLAN12 : q == 11 ? 0x344 : q == 10 ? 0x334 : q == 9 ? 0x333
         : q == 8 ? 0x44 : q == 7 ? 0x34 : q == 6 ? 0x33 : q == 5 ? 0x23 : 0;
}
__host__ __device__ constexpr int plan_npass(int q) {
    return (plan_code(q) >> 12) ? 4 : ((plan_code(q) >> 8) ? 3 : 2);
}
__host__ __device__ constexpr int plan_lr(int q, int t) { return (plan_code(q) >> (4 * t)) & 0xF; }
// log2 of the block length pass t works on, and of its butterfly stride
__host__ __device__ constexpr int plan_lqt(int q, int t) {
    int l = q;
    for (int s = 0; s < t; s++) l -= plan_lr(q, s);
    return l;
}
__host__ __device__ constexpr int plan_ls(int q, int t) { return plan_lqt(q, t) - plan_lr(q, t); }
// bin k' -> position in the half array (mixed-radix digit reversal) and back
__host__ __device__ constexpr int plan_rev(int q, int k) {
    int pos = 0, sh = 0;
    for (int t = 0; t < plan_npass(q); t++) {
        const int lr = plan_lr(q, t);
        pos |= ((k >> sh) & ((1 << lr) - 1)) << plan_ls(q, t);
        sh += lr;
    }
    return pos;
}
__host__ __device__ constexpr int plan_revinv(int q, int pos) {
    int k = 0, sh = 0;
    for (int t = 0; t < plan_npass(q); t++) {
        const int lr = plan_lr(q, t);
        k |= ((pos >> plan_ls(q, t)) & ((1 << lr) - 1)) << sh;
        sh += lr;
    }
    return k;
}
// NH = number of half transforms one CTA works on: 2 = the whole spectrum, 1 = one half
// (the halves are independent from the zero-padded load up to and including the unpack
// of the forward transform, so the forward kernels run one half per CTA: half the
// threads and shared memory per CTA, twice as many CTAs per SM, and their load, compute
// and store phases overlap instead of alternating).
__host__ __device__ constexpr int fft_threads(int log2n, int nh = 2) {
    const int n = (1 << log2n) * nh / 2;
    if (log2n == 13 && plan_npass(12) == 4) return 256 * nh;  // radix-8 plan: one more pass, twice the warps
    return (n / 32) > 256 ? 256 : ((n / 32) < 32 ? 32 : (n / 32));
}
__host__ __device__ constexpr int smem_pad(int e) { return e + (e >> 4); }
__host__ __device__ constexpr size_t fft_smem_bytes(int log2n, int nh = 2) {
    return (size_t)smem_pad((1 << log2n) * nh / 2) * sizeof(float2) + 64;
}

// ---- complex helpers --------------------------------------------------------
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmulconj(float2 a, float2 b) {  // a * conj(b)
    return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
// multiply by DIR*i  (DIR = -1: forward transform, +1: inverse)
template <int DIR>
__device__ __forceinline__ float2 mul_i(float2 a) {
    return DIR < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
}
// multiply by (c + DIR*i*s)
template <int DIR>
__device__ __forceinline__ float2 mul_w(float2 a, float c, float s) {
    const float ss = DIR < 0 ? -s : s;
    return make_float2(a.x * c - a.y * ss, a.x * ss + a.y * c);
}

// ---- register butterflies ---------------------------------------------------
// run<DIR>(v): v <- DFT_R(v) with kernel exp(DIR * 2 pi i j k / R).
// The result for bin k is left in register out(r) == k, i.e. register r holds
// bin out(r).
// All complex values are packed pairs (fcv_c2.cuh): a complex add/sub is ONE FADD2,
// multiplications by +-i fold into the operand modifiers of the add that consumes
// them, a twiddle multiplication is FMUL2 + FFMA2.
template <int R> struct Bfly;

// a * (DIR * i)
template <int DIR>
__device__ __forceinline__ c2 c2_mul_i(c2 a) { return DIR < 0 ? c2_mul_ni(a) : c2_mul_pi(a); }
// a * (c + DIR*i*s), c and s compile-time constants
template <int DIR>
__device__ __forceinline__ c2 c2_mul_w(c2 a, float c, float s) { return c2_cmul(a, c2_pack(c, DIR < 0 ? -s : s)); }

template <> struct Bfly<2> {
    __host__ __device__ static constexpr int out(int r) { return r; }
    template <int DIR> __device__ __forceinline__ static void run(c2 (&v)[2]) {
        const c2 a = v[0], b = v[1];
        v[0] = c2_add(a, b);
        v[1] = c2_sub(a, b);
    }
};

template <int DIR>
__device__ __forceinline__ void dft4(c2 &a, c2 &b, c2 &c, c2 &d) {
    const c2 s0 = c2_add(a, c), d0 = c2_sub(a, c), s1 = c2_add(b, d), t = c2_sub(b, d);
    a = c2_add(s0, s1);
    c = c2_sub(s0, s1);
    b = c2_add(d0, c2_mul_i<DIR>(t));
    d = c2_sub(d0, c2_mul_i<DIR>(t));
}

template <> struct Bfly<4> {
    __host__ __device__ static constexpr int out(int r) { return r; }
    template <int DIR> __device__ __forceinline__ static void run(c2 (&v)[4]) {
        dft4<DIR>(v[0], v[1], v[2], v[3]);
    }
};

// j = j0 + 4 j1 (j1 in {0,1}), k = k1 + 2 k0: register k0 + 4 k1 holds bin k1 + 2 k0
template <> struct Bfly<8> {
    __host__ __device__ static constexpr int out(int r) { return (r >> 2) + 2 * (r & 3); }
    template <int DIR> __device__ __forceinline__ static void run(c2 (&v)[8]) {
        constexpr float C2 = 0.70710678118654752440f;
#pragma unroll
        for (int j0 = 0; j0 < 4; j0++) {
            const c2 a = v[j0], b = v[j0 + 4];
            v[j0] = c2_add(a, b);
            v[j0 + 4] = c2_sub(a, b);
        }
        // twiddle w_8^(j0*k1), k1 = 1 row only
        v[5] = c2_mul_w<DIR>(v[5], C2, C2);
        v[6] = c2_mul_i<DIR>(v[6]);
        v[7] = c2_mul_w<DIR>(v[7], -C2, C2);
        dft4<DIR>(v[0], v[1], v[2], v[3]);
        dft4<DIR>(v[4], v[5], v[6], v[7]);
    }
};

// j = j0 + 4 j1, k = k1 + 4 k0: register k0 + 4 k1 holds bin k1 + 4 k0
template <> struct Bfly<16> {
    __host__ __device__ static constexpr int out(int r) { return (r >> 2) + 4 * (r & 3); }
    template <int DIR> __device__ __forceinline__ static void run(c2 (&v)[16]) {
        constexpr float C1 = 0.92387953251128675613f;  // cos(pi/8)
        constexpr float S1 = 0.38268343236508977173f;  // sin(pi/8)
        constexpr float C2 = 0.70710678118654752440f;
#pragma unroll
        for (int j0 = 0; j0 < 4; j0++) dft4<DIR>(v[j0], v[j0 + 4], v[j0 + 8], v[j0 + 12]);
        // v[j0 + 4 k1] *= w_16^(j0 k1)
        v[5] = c2_mul_w<DIR>(v[5], C1, S1);     // 1
        v[6] = c2_mul_w<DIR>(v[6], C2, C2);     // 2
        v[7] = c2_mul_w<DIR>(v[7], S1, C1);     // 3
        v[9] = c2_mul_w<DIR>(v[9], C2, C2);     // 2
        v[10] = c2_mul_i<DIR>(v[10]);           // 4
        v[11] = c2_mul_w<DIR>(v[11], -C2, C2);  // 6
        v[13] = c2_mul_w<DIR>(v[13], S1, C1);   // 3
        v[14] = c2_mul_w<DIR>(v[14], -C2, C2);  // 6
        v[15] = c2_mul_w<DIR>(v[15], -C1, -S1); // 9
#pragma unroll
        for (int k1 = 0; k1 < 4; k1++) dft4<DIR>(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
    }
};

// ---- passes over both half transforms ----------------------------------------
// When the butterfly stride S divides the thread count, every butterfly a thread
// handles in a pass has the same u = b mod S, i.e. the same R-1 twiddles: they are
// loaded once per pass (TW_INVARIANT), otherwise once per butterfly.
//
// Shared-memory addressing: element base + j*S of a butterfly sits at
// smem_pad(base) + smem_pad(j*S) whenever the pass's block length is >= 16 (the low
// four bits of base are then u < S and never carry), i.e. one register plus
// compile-time offsets.
template <int LS, int LQT>
__device__ __forceinline__ int bfly_pos(int pbase, int base, int j) {
    if (LQT >= 4) return pbase + smem_pad(j << LS);
    return smem_pad(base + (j << LS));
}

template <int Q_LOG2, int T, int NT, int NH>
__device__ __forceinline__ void fwd_pass(float2 *smf, const float2 *__restrict__ twf, int tid) {
    constexpr int LR = plan_lr(Q_LOG2, T), R = 1 << LR;
    constexpr int LS = plan_ls(Q_LOG2, T), S = 1 << LS;
    constexpr int LQT = plan_lqt(Q_LOG2, T);
    constexpr int NB = (NH << Q_LOG2) >> LR;
    constexpr bool TW_INVARIANT = S > 1 && (NT % S) == 0;
    constexpr int NI = (NB + NT - 1) / NT;       // butterflies per thread
    constexpr int UN = (NI % 2 == 0 && R <= 16) ? 2 : 1;  // two independent butterflies in flight
    c2 *sm = reinterpret_cast<c2 *>(smf);
    const c2 *tw = reinterpret_cast<const c2 *>(twf);
    c2 w[R];
    if (TW_INVARIANT) {
#pragma unroll
        for (int k1 = 1; k1 < R; k1++) w[k1] = __ldg(&tw[(k1 - 1) * S + (tid & (S - 1))]);
    }
#pragma unroll 1
    for (int b0 = tid; b0 < NB; b0 += NT * UN) {
        c2 v[UN][R];
        int pb[UN], bs[UN];
#pragma unroll
        for (int q = 0; q < UN; q++) {
            const int b = b0 + q * NT;
            const int u = b & (S - 1);
            bs[q] = ((b >> LS) << LQT) + u;
            pb[q] = smem_pad(bs[q]);
#pragma unroll
            for (int j = 0; j < R; j++) v[q][j] = sm[bfly_pos<LS, LQT>(pb[q], bs[q], j)];
        }
#pragma unroll
        for (int q = 0; q < UN; q++) {
            const int b = b0 + q * NT;
            if (b < NB || UN == 1) {
                if (S > 1 && !TW_INVARIANT) {
                    const int u = b & (S - 1);
#pragma unroll
                    for (int k1 = 1; k1 < R; k1++) w[k1] = __ldg(&tw[(k1 - 1) * S + u]);
                }
                Bfly<R>::template run<-1>(v[q]);
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const int k1 = Bfly<R>::out(r);
                    c2 x = v[q][r];
                    if (S > 1 && k1 > 0) x = c2_cmul(x, w[k1]);
                    sm[bfly_pos<LS, LQT>(pb[q], bs[q], k1)] = x;
                }
            }
        }
    }
    __syncthreads();
}

template <int Q_LOG2, int T, int NT, int NH>
__device__ __forceinline__ void inv_pass(float2 *smf, const float2 *__restrict__ twf, int tid) {
    constexpr int LR = plan_lr(Q_LOG2, T), R = 1 << LR;
    constexpr int LS = plan_ls(Q_LOG2, T), S = 1 << LS;
    constexpr int LQT = plan_lqt(Q_LOG2, T);
    constexpr int NB = (NH << Q_LOG2) >> LR;
    constexpr bool TW_INVARIANT = S > 1 && (NT % S) == 0;
    constexpr int NI = (NB + NT - 1) / NT;
    constexpr int UN = (NI % 2 == 0 && R <= 16) ? 2 : 1;
    c2 *sm = reinterpret_cast<c2 *>(smf);
    const c2 *tw = reinterpret_cast<const c2 *>(twf);
    c2 w[R];
    if (TW_INVARIANT) {
#pragma unroll
        for (int k1 = 1; k1 < R; k1++) w[k1] = __ldg(&tw[(k1 - 1) * S + (tid & (S - 1))]);
    }
#pragma unroll 1
    for (int b0 = tid; b0 < NB; b0 += NT * UN) {
        c2 v[UN][R];
        int pb[UN], bs[UN];
#pragma unroll
        for (int q = 0; q < UN; q++) {
            const int b = b0 + q * NT;
            const int u = b & (S - 1);
            bs[q] = ((b >> LS) << LQT) + u;
            pb[q] = smem_pad(bs[q]);
#pragma unroll
            for (int k1 = 0; k1 < R; k1++) v[q][k1] = sm[bfly_pos<LS, LQT>(pb[q], bs[q], k1)];
        }
#pragma unroll
        for (int q = 0; q < UN; q++) {
            const int b = b0 + q * NT;
            if (S > 1 && !TW_INVARIANT) {
                const int u = b & (S - 1);
#pragma unroll
                for (int k1 = 1; k1 < R; k1++) w[k1] = __ldg(&tw[(k1 - 1) * S + u]);
            }
            if (S > 1) {
#pragma unroll
                for (int k1 = 1; k1 < R; k1++) v[q][k1] = c2_cmulconj(v[q][k1], w[k1]);
            }
            Bfly<R>::template run<+1>(v[q]);
#pragma unroll
            for (int r = 0; r < R; r++) sm[bfly_pos<LS, LQT>(pb[q], bs[q], Bfly<R>::out(r))] = v[q][r];
        }
    }
    __syncthreads();
}

template <int Q_LOG2, int NT, int NH>
__device__ __forceinline__ void fwd_passes(float2 *sm, const FftTables &tb, int tid) {
    fwd_pass<Q_LOG2, 0, NT, NH>(sm, tb.twP[0], tid);
    fwd_pass<Q_LOG2, 1, NT, NH>(sm, tb.twP[1], tid);
    if constexpr (plan_npass(Q_LOG2) >= 3) fwd_pass<Q_LOG2, 2, NT, NH>(sm, tb.twP[2], tid);
    if constexpr (plan_npass(Q_LOG2) >= 4) fwd_pass<Q_LOG2, 3, NT, NH>(sm, tb.twP[3], tid);
}
template <int Q_LOG2, int NT, int NH>
__device__ __forceinline__ void inv_passes(float2 *sm, const FftTables &tb, int tid) {
    if constexpr (plan_npass(Q_LOG2) >= 4) inv_pass<Q_LOG2, 3, NT, NH>(sm, tb.twP[3], tid);
    if constexpr (plan_npass(Q_LOG2) >= 3) inv_pass<Q_LOG2, 2, NT, NH>(sm, tb.twP[2], tid);
    inv_pass<Q_LOG2, 1, NT, NH>(sm, tb.twP[1], tid);
    inv_pass<Q_LOG2, 0, NT, NH>(sm, tb.twP[0], tid);
}

// entry e of the packed-permuted layout <-> its conjugate-partner entry
template <int Q_LOG2>
__device__ __forceinline__ int partner_entry(int e, int &kp_out, int &kpp_out) {
    constexpr int Q = 1 << Q_LOG2;
    const int half = e >> Q_LOG2;
    const int kp = plan_revinv(Q_LOG2, e & (Q - 1));
    const int kpp = half ? (Q - 1 - kp) : ((Q - kp) & (Q - 1));
    kp_out = kp;
    kpp_out = kpp;
    return (half << Q_LOG2) + plan_rev(Q_LOG2, kpp);
}

// Entry holding the conjugate partner (bin M - k) of the bin stored at entry e,
// computed in position space: half 1 is a mirror (k'' = Q-1-k' complements every
// digit); half 0 negates k' digit by digit with the carry running from the
// lowest bin digit (the highest position digit) upwards.
template <int Q_LOG2>
__device__ __forceinline__ int partner_of(int e) {
    constexpr int Q = 1 << Q_LOG2;
    const int pos = e & (Q - 1);
    if (e >> Q_LOG2) return Q + (Q - 1 - pos);
    int carry = 1, pos2 = 0;
#pragma unroll
    for (int t = 0; t < plan_npass(Q_LOG2); t++) {
        const int lr = plan_lr(Q_LOG2, t), ls = plan_ls(Q_LOG2, t), mask = (1 << lr) - 1;
        int nd = (mask - ((pos >> ls) & mask)) + carry;
        carry = nd >> lr;
        pos2 |= (nd & mask) << ls;
    }
    return pos2;
}

// ---- PCM wire formats ---------------------------------------------------------

// Frames 2n and 2n+1 of channel `chan` of an interleaved block as (x[2n], x[2n+1]).
// NCH = 2 / 1: stereo / mono blocks, one vector load per pair of frames;
// NCH = 0: any channel count, two scalar loads.
template <int FMT, int NCH>
__device__ __forceinline__ float2 pcm_load2(const void *in, int nchan, int chan, int n) {
    if (FMT == PCM_F32) {
        if (NCH == 2) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(in) + n);
            return chan ? make_float2(v.y, v.w) : make_float2(v.x, v.z);
        }
        if (NCH == 1) return __ldg(reinterpret_cast<const float2 *>(in) + n);
        const float *p = reinterpret_cast<const float *>(in) + (size_t)(2 * n) * nchan + chan;
        return make_float2(__ldg(p), __ldg(p + nchan));
    } else if (FMT == PCM_S16) {
        constexpr float K = 1.0f / 32768.0f;
        if (NCH == 2) {
            const short4 v = __ldg(reinterpret_cast<const short4 *>(in) + n);
            return chan ? make_float2(v.y * K, v.w * K) : make_float2(v.x * K, v.z * K);
        }
        if (NCH == 1) {
            const short2 v = __ldg(reinterpret_cast<const short2 *>(in) + n);
            return make_float2(v.x * K, v.y * K);
        }
        const short *p = reinterpret_cast<const short *>(in) + (size_t)(2 * n) * nchan + chan;
        return make_float2(__ldg(p) * K, __ldg(p + nchan) * K);
    } else {
        constexpr float K = 1.0f / 8388608.0f;
        if (NCH == 2) {
            const int4 v = __ldg(reinterpret_cast<const int4 *>(in) + n);
            return chan ? make_float2(v.y * K, v.w * K) : make_float2(v.x * K, v.z * K);
        }
        if (NCH == 1) {
            const int2 v = __ldg(reinterpret_cast<const int2 *>(in) + n);
            return make_float2(v.x * K, v.y * K);
        }
        const int *p = reinterpret_cast<const int *>(in) + (size_t)(2 * n) * nchan + chan;
        return make_float2(__ldg(p) * K, __ldg(p + nchan) * K);
    }
}

template <int FMT>
__device__ __forceinline__ void pcm_store(void *p, size_t idx, float v) {
    if (FMT == PCM_F32) {
        reinterpret_cast<float *>(p)[idx] = v;
    } else if (FMT == PCM_S16) {
        // libsndfile f2s_array without clipping: lrintf(x * 0x7FFF) stored to a short (wraps)
        reinterpret_cast<short *>(p)[idx] = (short)__float2int_rn(v * 32767.0f);
    } else {
        reinterpret_cast<int *>(p)[idx] = __float2int_rn(v * 8388607.0f);
    }
}

// ---- forward: one zero-padded partition -> packed-permuted spectrum -----------
// in: interleaved PCM, `nchan` channels, channel `chan`; frames >= frames_valid read as 0.
// All global loads of a stage are issued before their first use (the stages are
// latency bound otherwise: one CTA only has 8 warps).
template <int LOG2N, int FMT, int NCH, int NH>
__device__ __forceinline__ void fwd_load(float2 *sm, const FftTables &tb, const void *in, int nchan,
                                         int chan, int frames_valid, int h0) {
    constexpr int QL = LOG2N - 1, Q = 1 << QL;
    constexpr int NT = fft_threads(LOG2N, NH);
    constexpr int IT = Q / NT;
    constexpr int CH = IT < 8 ? IT : 8;
    const int tid = threadIdx.x;
#pragma unroll 1
    for (int c = 0; c < IT; c += CH) {
        float2 z[CH], w[CH];
#pragma unroll
        for (int i = 0; i < CH; i++) z[i] = pcm_load2<FMT, NCH>(in, nchan, chan, tid + (c + i) * NT);
        if (NH == 2 || h0) {
#pragma unroll
            for (int i = 0; i < CH; i++) w[i] = __ldg(&tb.twA[tid + (c + i) * NT]);
        }
#pragma unroll
        for (int i = 0; i < CH; i++) {
            const int n = tid + (c + i) * NT;
            float2 v = z[i];
            if (2 * n >= frames_valid) v.x = 0.0f;
            if (2 * n + 1 >= frames_valid) v.y = 0.0f;
            if (NH == 2) {
                sm[smem_pad(n)] = v;
                sm[smem_pad(Q + n)] = cmul(v, w[i]);
            } else {
                sm[smem_pad(n)] = h0 ? cmul(v, w[i]) : v;
            }
        }
    }
}

// NH == 2: the whole spectrum; NH == 1: half h0 of it (entries [h0*Q, (h0+1)*Q) of out_row).
template <int LOG2N, int FMT, int NH>
__device__ __forceinline__ void fwd_body(float2 *sm, const FftTables &tb, const void *in, int nchan,
                                         int chan, int frames_valid, float2 *__restrict__ out_row, int h0 = 0) {
    constexpr int QL = LOG2N - 1, Q = 1 << QL, ME = NH * Q;
    constexpr int NT = fft_threads(LOG2N, NH);
    const int tid = threadIdx.x;
    if (nchan == 2) fwd_load<LOG2N, FMT, 2, NH>(sm, tb, in, nchan, chan, frames_valid, h0);
    else if (nchan == 1) fwd_load<LOG2N, FMT, 1, NH>(sm, tb, in, nchan, chan, frames_valid, h0);
    else fwd_load<LOG2N, FMT, 0, NH>(sm, tb, in, nchan, chan, frames_valid, h0);
    __syncthreads();
    fwd_passes<QL, NT, NH>(sm, tb, tid);
    // unpack: X[k] = E - i w D from the bin and its conjugate partner (same half)
    const int e0 = NH == 2 ? 0 : h0 * Q;  // first global entry this CTA holds
    constexpr int CH = (ME / NT) < 8 ? (ME / NT) : 8;
#pragma unroll 1
    for (int c = 0; c < ME / NT; c += CH) {
        float2 w[CH], x[CH];
        int e2[CH];
#pragma unroll
        for (int i = 0; i < CH; i++) {
            w[i] = __ldg(&tb.twU[e0 + tid + (c + i) * NT]);
            e2[i] = (int)__ldg(&tb.part[e0 + tid + (c + i) * NT]) - e0;
        }
#pragma unroll
        for (int i = 0; i < CH; i++) {
            const int e = tid + (c + i) * NT;
            const float2 zk = sm[smem_pad(e)];
            const float2 zp = sm[smem_pad(e2[i])];
            const float2 ev = make_float2(0.5f * (zk.x + zp.x), 0.5f * (zk.y - zp.y));
            const float2 dv = make_float2(0.5f * (zk.x - zp.x), 0.5f * (zk.y + zp.y));
            const float2 t = cmul(w[i], dv);
            x[i] = make_float2(ev.x + t.y, ev.y - t.x);              // E - i w D
            if (e0 + e == 0) x[i] = make_float2(zk.x + zk.y, zk.x - zk.y);  // DC, Nyquist
        }
#pragma unroll
        for (int i = 0; i < CH; i++) out_row[e0 + tid + (c + i) * NT] = x[i];
    }
}

// ---- inverse: packed-permuted spectrum -> 2N real samples ------------------------
// inv_load reads the accumulated spectrum Y and its conjugate partners straight
// from global memory (the partner reads hit L2: the same CTA has just fetched
// the row) and stores Zc[k] = (Y[k] + conj Y[M-k]) + i conj(w) (Y[k] - conj Y[M-k])
// to shared memory; entry 0 (DC, Nyquist) is supplied by the caller.
// inv_body then turns it in place into a'[n] (half 0) and b'[n] (half 1); the
// caller combines them:
//   z[n] = a' + conj(twA[n]) b' -> samples 2n, 2n+1;  z[n+Q] = a' - conj(twA[n]) b' -> samples N+2n, N+2n+1.
template <int LOG2N>
__device__ __forceinline__ void inv_load(float2 *sm, const FftTables &tb, const float2 *__restrict__ yrow) {
    constexpr int M = 1 << LOG2N;
    constexpr int NT = fft_threads(LOG2N);
    constexpr int CH = (M / NT) < 8 ? (M / NT) : 8;
    const int tid = threadIdx.x;
#pragma unroll 1
    for (int c = 0; c < M / NT; c += CH) {
        float2 y[CH], yp[CH], w[CH];
#pragma unroll
        for (int i = 0; i < CH; i++) {
            const int e = tid + (c + i) * NT;
            y[i] = __ldg(&yrow[e]);
            yp[i] = __ldg(&yrow[partner_of<LOG2N - 1>(e)]);
            w[i] = __ldg(&tb.twU[e]);
        }
#pragma unroll
        for (int i = 0; i < CH; i++) {
            const int e = tid + (c + i) * NT;
            const float2 ev = make_float2(y[i].x + yp[i].x, y[i].y - yp[i].y);
            const float2 dv = make_float2(y[i].x - yp[i].x, y[i].y + yp[i].y);
            const float2 t = cmulconj(dv, w[i]);
            if (e != 0) sm[smem_pad(e)] = make_float2(ev.x - t.y, ev.y + t.x);  // E + i conj(w) D
        }
    }
}

template <int LOG2N>
__device__ __forceinline__ void inv_body(float2 *sm, const FftTables &tb) {
    constexpr int QL = LOG2N - 1;
    constexpr int NT = fft_threads(LOG2N);
    inv_passes<QL, NT, 2>(sm, tb, threadIdx.x);
}

}  // namespace fcv
