// fcv_fft13.cuh -- the real FFT / inverse real FFT of one zero-padded partition for
// fragm = 8192 (every filter longer than 4096 taps, /root/reference/zita-fconfig.cc:74-77),
// built around what the profile says bounds these kernels on B200: the L1/LSU data pipe
// (one 128-byte wavefront per clock per SM, shared and global accesses together), not
// instruction issue, not DRAM latency, not occupancy (profiles/r01_fft_findings.md).
//
// Role in the reference path: same as fcv_fft.cuh -- FFTW's r2c / c2r inside
// zita-convolver's Convlevel::process(), reached from SoundProcessor::Process()
// (/root/reference/sound-processor.cc:98-127), with the (de)interleave, zero padding,
// overlap-add, int/float conversion and running maximum fused in.
//
// Method.  z[n] = x[2n] + i x[2n+1]; the block is zero padded, so the M = 8192 point
// transform of z splits into two independent Q = 4096 point transforms ("halves"):
// half h holds the bins k = 2k' + h.  Each half is 16 x 16 x 16:
//   pass A  over n2 (n = n0 + 16 n1 + 256 n2), straight from global memory into registers
//           (half 1: inputs times w_32^n2, constants; w_M^u folded into the output twiddles);
//           out  A[k0][u] * w,  u = n0 + 16 n1                         -> shared, row k0
//   pass B  over n1, in place in shared memory:  B[k0][n0 + 16 k1]
//   pass C  over n0, shared -> registers; one thread takes the run (k0,k1) AND the run of
//           its conjugate partners, so the real-spectrum unpack X = E - i w D happens in
//           registers and both results go straight to global memory.
// Shared memory is touched 4 times per element instead of 9, every table is read in
// lane order, and one CTA transforms both channels of a stereo block so that PCM and
// twiddles are fetched once for two transforms.
// The inverse runs the same passes backwards (C^-1 from global with the Hermitian
// repack in registers, B^-1 in place, A^-1 + overlap-add + PCM store from registers).
//
// Spectrum layout ("split-parity natural"): entry e = h*Q + k' holds bin k = 2k' + h;
// DC and Nyquist (both real) share entry 0.  The conjugate partner of bin k (bin M - k)
// is entry (Q - k') mod Q of the same half for h = 0 and Q - 1 - k' for h = 1.  The MAC
// kernels are element-wise and never look at the order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "fcv_c2.cuh"
#include "fcv_fft.cuh"

namespace fcv {
namespace f13 {

constexpr int LOG2N = 13;
constexpr int N = 1 << LOG2N;      // frames per block = complex entries per spectrum
constexpr int Q = N / 2;           // entries per half
constexpr int ROW = 257;           // shared-memory row stride (elements): odd -> stride-ROW accesses are conflict free
constexpr int HALF_ELEMS = 16 * ROW;
constexpr size_t HALF_BYTES = (size_t)HALF_ELEMS * sizeof(float2);

__host__ __device__ constexpr int out16(int r) { return (r >> 2) + 4 * (r & 3); }   // Bfly<16>::out
__host__ __device__ constexpr int reg16(int k) { return ((k & 3) << 2) | (k >> 2); }  // its inverse

// w_32^j = exp(-2 pi i j / 32), j = 0..15 (compile-time j after unrolling)
__device__ __forceinline__ float2 w32(int j) {
    constexpr float C[16] = {1.0f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                             0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f,
                             0.19509032201612826785f, 0.0f, -0.19509032201612826785f, -0.38268343236508977173f,
                             -0.55557023301960222474f, -0.70710678118654752440f, -0.83146961230254523708f,
                             -0.92387953251128675613f, -0.98078528040323044913f};
    constexpr float S[16] = {0.0f, 0.19509032201612826785f, 0.38268343236508977173f, 0.55557023301960222474f,
                             0.70710678118654752440f, 0.83146961230254523708f, 0.92387953251128675613f,
                             0.98078528040323044913f, 1.0f, 0.98078528040323044913f, 0.92387953251128675613f,
                             0.83146961230254523708f, 0.70710678118654752440f, 0.55557023301960222474f,
                             0.38268343236508977173f, 0.19509032201612826785f};
    return make_float2(C[j], -S[j]);
}

#ifndef F13_TWU_FACTORED
#define F13_TWU_FACTORED 0
#endif
// Unpack / repack twiddle exp(-i pi k / M) of entry 256 k2 + c of half H (bin k = 2 (256 k2 + c) + H):
//   exp(-i pi (2c + H) / M) * exp(-i pi k2 / 16)  =  base[H][c] * w_32^k2.
// -DF13_TWU_FACTORED=1 (experiment, OFF): only the 512 base values are read from memory (4 KB
// instead of the 64 KB table) and the 15 other twiddles of a run are one multiplication by a
// compile-time constant each, so that the twiddle tables of an SM shrink from 128 KB -- which cycle
// through ~100 KB of L1 and miss it three times out of four (ncu, profiles/r02b_kernels.md) -- to
// 68 KB.  Measured on one box, alternating builds (profiles/r02_experiments.md): forward 0.4084 ->
// 0.4357 ms, inverse 0.5672 -> 0.5710 ms: the 30 extra packed multiplies per run cost more than
// the L2 round trips they save (the kernels are as sensitive to their math / issue slots as to
// load latency); parity unchanged (SNR vs float64 132.4 against 132.8 dB).
__device__ __forceinline__ c2 ldg_c2(const float2 *p);
#if F13_TWU_FACTORED
#define F13_W(twu, k2, c, base) twu_of(base, k2)
#else
#define F13_W(twu, k2, c, base) ldg_c2((twu) + 256 * (k2) + (c))
#endif
__device__ __forceinline__ c2 twu_of(c2 base, int k2) { return k2 == 0 ? base : c2_cmul(base, c2_pack(w32(k2))); }

__device__ __forceinline__ c2 ldg_c2(const float2 *p) {
    const float2 v = __ldg(p);
    return c2_pack(v.x, v.y);
}
// Read-once data (spectrum rows written by the previous kernel): kept out of L1, where the
// ~100 KB of twiddle tables every CTA of the SM re-reads have to survive next to 2 x 66 KB
// of shared memory.
__device__ __forceinline__ c2 ldg_stream_c2(const float2 *p) {
    c2 v;
    asm volatile("ld.global.nc.L1::no_allocate.b64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

// Two stereo frames in wire format (frames 2n, 2n+1: 16 bytes as float32 / int24-in-int32, 8 bytes as int16, in
// r.x, r.y) -> the pairs (x[2n], x[2n+1]) of the left and of the right channel, libsndfile's scaling.
template <int FMT>
__device__ __forceinline__ void raw_to_pairs(uint4 r, float2 &left, float2 &right) {
    if (FMT == PCM_F32) {
        left = make_float2(__uint_as_float(r.x), __uint_as_float(r.z));
        right = make_float2(__uint_as_float(r.y), __uint_as_float(r.w));
    } else if (FMT == PCM_S16) {
        constexpr float K = 1.0f / 32768.0f;
        left = make_float2((short)(r.x & 0xffffu) * K, (short)(r.y & 0xffffu) * K);
        right = make_float2((short)(r.x >> 16) * K, (short)(r.y >> 16) * K);
    } else {
        constexpr float K = 1.0f / 8388608.0f;
        left = make_float2((int)r.x * K, (int)r.z * K);
        right = make_float2((int)r.y * K, (int)r.w * K);
    }
}

// z[n] = (x[2n], x[2n+1]) of C consecutive channels starting at ch0, frames >= fv read as 0.
// NCH = 2: stereo block and both channels wanted (one vector load); NCH = 1: mono block;
// NCH = 0: any layout, scalar loads.
// FULL: every frame of the block is valid (fv == N): no zeroing.
template <int FMT, int NCH, int C, bool FULL = false>
__device__ __forceinline__ void load_z(const void *in, int nchan, int ch0, int n, int fv, c2 (&z)[C]) {
    float2 v[C];
    if (NCH == 2 && C == 2) {
        if (FMT == PCM_F32) {
            const float4 t = __ldcs(reinterpret_cast<const float4 *>(in) + n);
            raw_to_pairs<FMT>(make_uint4(__float_as_uint(t.x), __float_as_uint(t.y), __float_as_uint(t.z), __float_as_uint(t.w)), v[0], v[1]);
        } else if (FMT == PCM_S16) {
            const uint2 t = __ldcs(reinterpret_cast<const uint2 *>(in) + n);
            raw_to_pairs<FMT>(make_uint4(t.x, t.y, 0u, 0u), v[0], v[1]);
        } else {
            const int4 t = __ldcs(reinterpret_cast<const int4 *>(in) + n);
            raw_to_pairs<FMT>(make_uint4((uint32_t)t.x, (uint32_t)t.y, (uint32_t)t.z, (uint32_t)t.w), v[0], v[1]);
        }
    } else {
#pragma unroll
        for (int c = 0; c < C; c++) v[c] = pcm_load2<FMT, (NCH == 1 ? 1 : 0)>(in, nchan, ch0 + c, n);
    }
#pragma unroll
    for (int c = 0; c < C; c++) {
        if (!FULL) {
            if (2 * n >= fv) v[c].x = 0.0f;
            if (2 * n + 1 >= fv) v[c].y = 0.0f;
        }
        z[c] = c2_pack(v[c].x, v[c].y);
    }
}

// ---- forward ---------------------------------------------------------------------------
// Pass A of half H for C channels: global PCM -> registers -> shared row k0.
template <int H, int FMT, int NCH, int C, int NT, bool FULL = false>
__device__ __forceinline__ void fwd_pass_a(c2 *sm, const Tables &tb, const void *in, int nchan, int ch0, int fv) {
#pragma unroll 1
    for (int u = threadIdx.x; u < 256; u += NT) {
        c2 v[C][16];
#pragma unroll
        for (int j = 0; j < 16; j++) {
            c2 z[C];
            load_z<FMT, NCH, C, FULL>(in, nchan, ch0, u + 256 * j, fv, z);
#pragma unroll
            for (int c = 0; c < C; c++) v[c][j] = z[c];
        }
        c2 w[16];
        if (H == 0) {
#pragma unroll
            for (int k0 = 1; k0 < 16; k0++) w[k0] = ldg_c2(tb.twA0 + (k0 - 1) * 256 + u);
        } else {
#pragma unroll
            for (int k0 = 0; k0 < 16; k0++) w[k0] = ldg_c2(tb.twA1 + k0 * 256 + u);
        }
#pragma unroll
        for (int c = 0; c < C; c++) {
            if (H == 1) {
#pragma unroll
                for (int j = 1; j < 16; j++) v[c][j] = c2_cmul(v[c][j], c2_pack(w32(j)));
            }
            Bfly<16>::template run<-1>(v[c]);
#pragma unroll
            for (int r = 0; r < 16; r++) {
                const int k0 = out16(r);
                c2 x = v[c][r];
                if (H == 1 || k0 > 0) x = c2_cmul(x, w[k0]);
                sm[c * HALF_ELEMS + k0 * ROW + u] = x;
            }
        }
    }
}

// Pass B, in place: thread (k0, n0) transforms over n1 / k1 at row k0, columns n0 + 16 j.
// NA arrays of HALF_ELEMS (channels in the forward kernel, halves in the inverse one).
template <int DIR, int NA, int NT>
__device__ __forceinline__ void pass_b(c2 *sm, const Tables &tb) {
    const int k0 = threadIdx.x & 15;
#pragma unroll 1
    for (int n0 = threadIdx.x >> 4; n0 < 16; n0 += NT / 16) {
        c2 w[16];
#pragma unroll
        for (int k1 = 1; k1 < 16; k1++) w[k1] = ldg_c2(tb.twB + n0 * 16 + k1);
#pragma unroll
        for (int a = 0; a < NA; a++) {
            c2 *p = sm + a * HALF_ELEMS + k0 * ROW + n0;
            c2 v[16];
#pragma unroll
            for (int j = 0; j < 16; j++) v[j] = p[16 * j];
            if (DIR > 0) {
#pragma unroll
                for (int k1 = 1; k1 < 16; k1++) v[k1] = c2_cmulconj(v[k1], w[k1]);
            }
            Bfly<16>::template run<DIR>(v);
#pragma unroll
            for (int r = 0; r < 16; r++) {
                const int k = out16(r);
                c2 x = v[r];
                if (DIR < 0 && k > 0) x = c2_cmul(x, w[k]);
                p[16 * k] = x;
            }
        }
    }
}

// Pass B of ONE half by NTH threads (ht = 0 .. NTH-1): the two halves of an inverse transform are
// independent up to pass A^-1, so each half's 128 threads can run C^-1 -> B behind a barrier of
// their own (inverse kernel, F13_INV_SPLIT).
template <int DIR, int NTH>
__device__ __forceinline__ void pass_b_one(c2 *sm, const Tables &tb, int ht) {
    const int k0 = ht & 15;
#pragma unroll 1
    for (int n0 = ht >> 4; n0 < 16; n0 += NTH / 16) {
        c2 w[16];
#pragma unroll
        for (int k1 = 1; k1 < 16; k1++) w[k1] = ldg_c2(tb.twB + n0 * 16 + k1);
        c2 *p = sm + k0 * ROW + n0;
        c2 v[16];
#pragma unroll
        for (int j = 0; j < 16; j++) v[j] = p[16 * j];
        if (DIR > 0) {
#pragma unroll
            for (int k1 = 1; k1 < 16; k1++) v[k1] = c2_cmulconj(v[k1], w[k1]);
        }
        Bfly<16>::template run<DIR>(v);
#pragma unroll
        for (int r = 0; r < 16; r++) {
            const int k = out16(r);
            c2 x = v[r];
            if (DIR < 0 && k > 0) x = c2_cmul(x, w[k]);
            p[16 * k] = x;
        }
    }
}

// X[k] = E - i w D and X[M-k] = conj(E + i w D) from Z[k] (zk) and Z[M-k] (zp)
__device__ __forceinline__ void unpack_pair(c2 zk, c2 zp, c2 w, c2 &xk, c2 &xp) {
    const c2 e = c2_scale(c2_add(zk, c2_conj(zp)), 0.5f);
    const c2 d = c2_scale(c2_sub(zk, c2_conj(zp)), 0.5f);
    const c2 t = c2_cmul(d, w);
    xk = c2_add(e, c2_mul_ni(t));           // E - i t
    xp = c2_conj(c2_add(e, c2_mul_pi(t)));  // conj(E + i t)
}

// Pass C of half H + unpack, one channel: shared -> registers -> global row (entries [H*Q, (H+1)*Q)).
template <int H>
__device__ __forceinline__ void fwd_pass_c(const c2 *sm, const Tables &tb, float2 *__restrict__ row, int t) {
    c2 *out = reinterpret_cast<c2 *>(row) + H * Q;
#if F13_TWU_FACTORED
    const float2 *twu = tb.twU + H * 256;   // base values exp(-i pi (2c + H) / M), c = 0..255
#else
    const float2 *twu = tb.twU + H * Q;
#endif
    if (H == 0 && t == 0) {
        // runs c = 0 and c = 128 hold their own partners: k2 <-> 16 - k2 and k2 <-> 15 - k2
#if F13_TWU_FACTORED
        const c2 b0 = ldg_c2(twu), b128 = ldg_c2(twu + 128);
#endif
        c2 v1[16], v2[16];
#pragma unroll
        for (int j = 0; j < 16; j++) {
            v1[j] = sm[j];                  // (k0, k1) = (0, 0): row 0, columns n0
            v2[j] = sm[16 * 8 + j];         // (k0, k1) = (0, 8): row 0, columns 128 + n0
        }
        Bfly<16>::template run<-1>(v1);
        Bfly<16>::template run<-1>(v2);
        {
            const float2 z0 = c2_unpack(v1[reg16(0)]);
            out[0] = c2_pack(z0.x + z0.y, z0.x - z0.y);  // DC, Nyquist
        }
#pragma unroll
        for (int k2 = 1; k2 <= 8; k2++) {
            c2 xk, xp;
            unpack_pair(v1[reg16(k2)], v1[reg16(16 - k2)], F13_W(twu, k2, 0, b0), xk, xp);
            out[256 * k2] = xk;
            if (k2 != 8) out[256 * (16 - k2)] = xp;
        }
#pragma unroll
        for (int k2 = 0; k2 < 8; k2++) {
            c2 xk, xp;
            unpack_pair(v2[reg16(k2)], v2[reg16(15 - k2)], F13_W(twu, k2, 128, b128), xk, xp);
            out[256 * k2 + 128] = xk;
            out[256 * (15 - k2) + 128] = xp;
        }
        return;
    }
    const int c = t, cc = H == 0 ? 256 - t : 255 - t;
#if F13_TWU_FACTORED
    const c2 bc = ldg_c2(twu + c);
#endif
    c2 v1[16], v2[16];
    {
        const c2 *p1 = sm + (c & 15) * ROW + (c >> 4) * 16;
        const c2 *p2 = sm + (cc & 15) * ROW + (cc >> 4) * 16;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            v1[j] = p1[j];
            v2[j] = p2[j];
        }
    }
    Bfly<16>::template run<-1>(v1);
    Bfly<16>::template run<-1>(v2);
#pragma unroll
    for (int k2 = 0; k2 < 16; k2++) {
        c2 xk, xp;
        unpack_pair(v1[reg16(k2)], v2[reg16(15 - k2)], F13_W(twu, k2, c, bc), xk, xp);
        out[256 * k2 + c] = xk;
        out[256 * (15 - k2) + cc] = xp;
    }
}

// One half (H) of the transforms of C channels of one block.  rows[c] = spectrum row of channel ch0 + c.
// NT = 128 * C threads: pass A one u per thread (C = 2) or two (C = 1), pass C one job per thread.
template <int H, int FMT, int NCH, int C, int NT>
__device__ __forceinline__ void fwd_half(c2 *sm, const Tables &tb, const void *in, int nchan, int ch0, int fv,
                                         float2 *const (&rows)[C], bool full_path = false) {
    // a whole block (the usual case) takes pass A without the per-sample zeroing (32 ISETP + 64 FSEL per thread)
    if (full_path && fv >= N) fwd_pass_a<H, FMT, NCH, C, NT, true>(sm, tb, in, nchan, ch0, fv);
    else fwd_pass_a<H, FMT, NCH, C, NT>(sm, tb, in, nchan, ch0, fv);
    __syncthreads();
    pass_b<-1, C, NT>(sm, tb);
    __syncthreads();
#pragma unroll 1
    for (int j = threadIdx.x; j < 128 * C; j += NT) {
        const int c = j >> 7;
        fwd_pass_c<H>(sm + c * HALF_ELEMS, tb, C == 1 ? rows[0] : (c ? rows[C - 1] : rows[0]), j & 127);
    }
}

// ---- inverse ---------------------------------------------------------------------------
// Zc[k] = (Y[k] + conj Y[M-k]) + i conj(w) (Y[k] - conj Y[M-k]),  Zc[M-k] = conj(E - i conj(w) D)
__device__ __forceinline__ void repack_pair(c2 yk, c2 yp, c2 w, c2 &zk, c2 &zp) {
    const c2 e = c2_add(yk, c2_conj(yp));
    const c2 d = c2_sub(yk, c2_conj(yp));
    const c2 t = c2_cmulconj(d, w);
    zk = c2_add(e, c2_mul_pi(t));           // E + i t
    zp = c2_conj(c2_add(e, c2_mul_ni(t)));  // conj(E - i t)
}

// Pass C^-1 of half H: global spectrum row -> registers (Hermitian repack) -> shared.
// zc0 = the value of entry 0 (from the two real bins), used by thread 0 of half 0 only.
// CWAIT (stereo-pair kernel): the second cluster barrier of the previous block is waited for here, just
// before the first write to shared memory, instead of at the end of that block.
__device__ __forceinline__ void cluster_wait_divergent() { asm volatile("barrier.cluster.wait.acquire;" ::: "memory"); }
template <int H, bool CWAIT = false>
__device__ __forceinline__ void inv_pass_c(c2 *sm, const Tables &tb, const float2 *__restrict__ yrow, c2 zc0, int t) {
    const float2 *y = yrow + H * Q;
#if F13_TWU_FACTORED
    const float2 *twu = tb.twU + H * 256;
#else
    const float2 *twu = tb.twU + H * Q;
#endif
    if (H == 0 && t == 0) {
#if F13_TWU_FACTORED
        const c2 b0 = ldg_c2(twu), b128 = ldg_c2(twu + 128);
#endif
        c2 y1[16], y2[16], v1[16], v2[16];
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) {
            y1[k2] = ldg_stream_c2(y + 256 * k2);
            y2[k2] = ldg_stream_c2(y + 256 * k2 + 128);
        }
        v1[0] = zc0;
#pragma unroll
        for (int k2 = 1; k2 <= 8; k2++) {
            c2 zk, zp;
            repack_pair(y1[k2], y1[16 - k2], F13_W(twu, k2, 0, b0), zk, zp);
            v1[k2] = zk;
            if (k2 != 8) v1[16 - k2] = zp;
        }
#pragma unroll
        for (int k2 = 0; k2 < 8; k2++) {
            c2 zk, zp;
            repack_pair(y2[k2], y2[15 - k2], F13_W(twu, k2, 128, b128), zk, zp);
            v2[k2] = zk;
            v2[15 - k2] = zp;
        }
        Bfly<16>::template run<+1>(v1);
        Bfly<16>::template run<+1>(v2);
        if (CWAIT) cluster_wait_divergent();
#pragma unroll
        for (int r = 0; r < 16; r++) {
            sm[out16(r)] = v1[r];
            sm[16 * 8 + out16(r)] = v2[r];
        }
        return;
    }
    const int c = t, cc = H == 0 ? 256 - t : 255 - t;
#if F13_TWU_FACTORED
    const c2 bc = ldg_c2(twu + c);
#endif
    c2 v1[16], v2[16];
    {
        c2 y1[16], y2[16], w[16];
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) {
            y1[k2] = ldg_stream_c2(y + 256 * k2 + c);
            y2[k2] = ldg_stream_c2(y + 256 * k2 + cc);
            w[k2] = F13_W(twu, k2, c, bc);
        }
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) repack_pair(y1[k2], y2[15 - k2], w[k2], v1[k2], v2[15 - k2]);
    }
    Bfly<16>::template run<+1>(v1);
    Bfly<16>::template run<+1>(v2);
    c2 *p1 = sm + (c & 15) * ROW + (c >> 4) * 16;
    c2 *p2 = sm + (cc & 15) * ROW + (cc >> 4) * 16;
    if (CWAIT) cluster_wait_divergent();
#pragma unroll
    for (int r = 0; r < 16; r++) {
        p1[out16(r)] = v1[r];
        p2[out16(r)] = v2[r];
    }
}

// pcm_store at a byte address.  The epilogues below address a thread's samples as
// base + (512 n2) * frame_bytes (+ frame_bytes) with compile-time n2: one IMAD.WIDE per pair instead of the
// 64-bit (2n * nout + o) * width chain per sample (14 -> 3 address instructions per pair of samples).
template <int FMT>
__device__ __forceinline__ void pcm_store_b(char *p, float v) {
    if (FMT == PCM_F32) *reinterpret_cast<float *>(p) = v;
    else if (FMT == PCM_S16) *reinterpret_cast<short *>(p) = (short)__float2int_rn(v * 32767.0f);
    else *reinterpret_cast<int *>(p) = __float2int_rn(v * 8388607.0f);
}

// Pass A^-1 of both halves + overlap-add, tail save, re-interleave, float -> PCM and the
// signed maximum of the valid frames (sound-processor.cc:115-125), all from registers.
// sm: [2][HALF_ELEMS] (half 0, half 1).  Returns this thread's maximum.
template <int FMT, int NT>
__device__ __forceinline__ float inv_pass_a(const c2 *sm, const Tables &tb, float2 *__restrict__ tail, void *dout,
                                            int nout, int o, int frames) {
    float lmax = 0.0f;
    const uint32_t fbytes = (uint32_t)nout * (FMT == PCM_S16 ? 2u : 4u);   // bytes per frame
#pragma unroll 1
    for (int u = threadIdx.x; u < 256; u += NT) {
        char *const pbase = reinterpret_cast<char *>(dout) + (size_t)(2 * u) * fbytes + (size_t)o * (FMT == PCM_S16 ? 2 : 4);
        c2 va[16], vb[16];
#pragma unroll
        for (int k0 = 0; k0 < 16; k0++) {
            va[k0] = sm[k0 * ROW + u];
            vb[k0] = sm[HALF_ELEMS + k0 * ROW + u];
        }
#pragma unroll
        for (int k0 = 1; k0 < 16; k0++) va[k0] = c2_cmulconj(va[k0], ldg_c2(tb.twA0 + (k0 - 1) * 256 + u));
#pragma unroll
        for (int k0 = 0; k0 < 16; k0++) vb[k0] = c2_cmulconj(vb[k0], ldg_c2(tb.twA1 + k0 * 256 + u));
        Bfly<16>::template run<+1>(va);
        Bfly<16>::template run<+1>(vb);
        float2 tl[16];
#pragma unroll
        for (int r = 0; r < 16; r++) tl[r] = __ldcg(&tail[u + 256 * out16(r)]);  // L2 only: read once per block
#pragma unroll
        for (int r = 0; r < 16; r++) {
            const int n2 = out16(r), n = u + 256 * n2;
            const c2 b = n2 == 0 ? vb[r] : c2_cmulconj(vb[r], c2_pack(w32(n2)));
            const float2 s = c2_unpack(c2_add(va[r], b));
            const float2 d = c2_unpack(c2_sub(va[r], b));
            const float y0 = s.x + tl[r].x, y1 = s.y + tl[r].y;
            tail[n] = d;
            const int f0 = 2 * n;
            char *p = pbase + (size_t)((uint32_t)(512 * n2) * fbytes);
            pcm_store_b<FMT>(p, y0);
            pcm_store_b<FMT>(p + fbytes, y1);
            if (f0 < frames) lmax = fmaxf(lmax, y0);
            if (f0 + 1 < frames) lmax = fmaxf(lmax, y1);
        }
    }
    return lmax;
}


// ---- tensor memory as a per-thread scratch: what a thread needs again in every block ----------
// The inverse kernel of a batch walks T blocks of one (stream, output).  In every block each thread
// re-reads the SAME 47 twiddles (31 of pass A^-1, 16 of the repack in pass C^-1: 128 KB of tables per
// SM cycling through ~100 KB of L1, hit rate 16 %, profiles/r02b_kernels.md) and reads back the 16
// overlap values it stored itself one block earlier -- 79 of its ~145 global-memory instructions per
// block.  B200's 256 KB of tensor memory per SM are idle in this kernel (no MMA): each CTA allocates
// 256 columns (2 CTAs per SM = all 512), every thread owns 128 words of them -- its lane of the warp's
// lane quarter, columns [128 (warp / 4), +128) -- and keeps there
//   words   0 ..  31   the overlap tail (16 x float2), carried from block to block
//   words  32 ..  95   pass A^-1 twiddles: slot k0 (1..15) = twA0[k0-1][u], slot 16 + k0 = twA1[k0][u]
//   words  96 .. 127   repack twiddles of pass C^-1: slot k2 = twU[256 k2 + c]
// tcgen05.ld / tcgen05.st (shape 32x32b: thread i <-> lane i, register j <-> column j) move 16 or 32
// registers per instruction.  Same values, same arithmetic: bit-identical to the kernel without it.
namespace tm {
constexpr uint32_t COLS = 256;            // per CTA
constexpr uint32_t TAIL = 0, TWA = 32, TWC = 96;
__device__ __forceinline__ void alloc(uint32_t *slot, uint32_t ncols) {   // one whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void dealloc(uint32_t addr, uint32_t ncols) {   // one whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// this thread's 128 words: lane quarter of its warp, column half of its warp's group of four
__device__ __forceinline__ uint32_t thread_base(uint32_t cta_base) {
    const uint32_t warp = threadIdx.x >> 5;
    return cta_base + (((warp & 3u) * 32u) << 16) + (warp >> 2) * 128u;
}
// NC2 packed complex values = 2 NC2 columns; the load waits for its data inside the same statement
template <int NC2> __device__ __forceinline__ void tm_ld(uint32_t taddr, c2 *v);
template <int NC2> __device__ __forceinline__ void tm_st(uint32_t taddr, const c2 *v);
template <>
__device__ __forceinline__ void tm_ld<8>(uint32_t taddr, c2 *v) {
    asm volatile("{ .reg .b32 t<16>; tcgen05.ld.sync.aligned.32x32b.x16.b32 {t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15}, [%8]; tcgen05.wait::ld.sync.aligned; mov.b64 %0, {t0, t1}; mov.b64 %1, {t2, t3}; mov.b64 %2, {t4, t5}; mov.b64 %3, {t6, t7}; mov.b64 %4, {t8, t9}; mov.b64 %5, {t10, t11}; mov.b64 %6, {t12, t13}; mov.b64 %7, {t14, t15}; }"
                 : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]), "=l"(v[4]), "=l"(v[5]), "=l"(v[6]), "=l"(v[7])
                 : "r"(taddr)
                 : "memory");
}
template <>
__device__ __forceinline__ void tm_ld<16>(uint32_t taddr, c2 *v) {
    asm volatile("{ .reg .b32 t<32>; tcgen05.ld.sync.aligned.32x32b.x32.b32 {t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, t20, t21, t22, t23, t24, t25, t26, t27, t28, t29, t30, t31}, [%16]; tcgen05.wait::ld.sync.aligned; mov.b64 %0, {t0, t1}; mov.b64 %1, {t2, t3}; mov.b64 %2, {t4, t5}; mov.b64 %3, {t6, t7}; mov.b64 %4, {t8, t9}; mov.b64 %5, {t10, t11}; mov.b64 %6, {t12, t13}; mov.b64 %7, {t14, t15}; mov.b64 %8, {t16, t17}; mov.b64 %9, {t18, t19}; mov.b64 %10, {t20, t21}; mov.b64 %11, {t22, t23}; mov.b64 %12, {t24, t25}; mov.b64 %13, {t26, t27}; mov.b64 %14, {t28, t29}; mov.b64 %15, {t30, t31}; }"
                 : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]), "=l"(v[4]), "=l"(v[5]), "=l"(v[6]), "=l"(v[7]), "=l"(v[8]), "=l"(v[9]), "=l"(v[10]), "=l"(v[11]), "=l"(v[12]), "=l"(v[13]), "=l"(v[14]), "=l"(v[15])
                 : "r"(taddr)
                 : "memory");
}
template <>
__device__ __forceinline__ void tm_st<16>(uint32_t taddr, const c2 *v) {
    asm volatile("{ .reg .b32 t<32>; mov.b64 {t0, t1}, %1; mov.b64 {t2, t3}, %2; mov.b64 {t4, t5}, %3; mov.b64 {t6, t7}, %4; mov.b64 {t8, t9}, %5; mov.b64 {t10, t11}, %6; mov.b64 {t12, t13}, %7; mov.b64 {t14, t15}, %8; mov.b64 {t16, t17}, %9; mov.b64 {t18, t19}, %10; mov.b64 {t20, t21}, %11; mov.b64 {t22, t23}, %12; mov.b64 {t24, t25}, %13; mov.b64 {t26, t27}, %14; mov.b64 {t28, t29}, %15; mov.b64 {t30, t31}, %16; tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {t0, t1, t2, t3, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18, t19, t20, t21, t22, t23, t24, t25, t26, t27, t28, t29, t30, t31}; }"
                 :
                 : "r"(taddr), "l"(v[0]), "l"(v[1]), "l"(v[2]), "l"(v[3]), "l"(v[4]), "l"(v[5]), "l"(v[6]), "l"(v[7]), "l"(v[8]), "l"(v[9]), "l"(v[10]), "l"(v[11]), "l"(v[12]), "l"(v[13]), "l"(v[14]), "l"(v[15])
                 : "memory");
}
// Split form: the load is issued here and its 32 registers may only be touched behind ld_wait32, which names
// them as read-write operands so that neither nvcc nor ptxas moves a use in front of the wait.
__device__ __forceinline__ void ld_issue32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void ld_wait32(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}
__device__ __forceinline__ c2 pair_of(const uint32_t (&r)[32], int i) {
    c2 v;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "r"(r[2 * i]), "r"(r[2 * i + 1]));
    return v;
}
}  // namespace tm

// Pass C^-1 as inv_pass_c with the repack twiddles coming from tensor memory: wr = the 32 registers of a load
// issued by the caller (w[k2] = twU[256 k2 + c]; the thread of entry 0, t = 0 of half 0, finds those of its second
// run in the slots its first run leaves free: w[0] = twU[128], w[8 + k2] = twU[256 k2 + 128], k2 = 1..7 --
// fill_twc).  Every thread issues the same 32 spectrum loads (entry 0's thread with its own two columns) in front
// of the wait, which all lanes of a warp must reach together; only then the entry-0 thread takes its own path.
template <int H>
__device__ __forceinline__ void inv_pass_c_w(c2 *sm, const float2 *__restrict__ yrow, c2 zc0, int t, uint32_t (&wr)[32]) {
    const float2 *y = yrow + H * Q;
    const bool entry0 = H == 0 && t == 0;
    const int c = t, cc = entry0 ? 128 : (H == 0 ? 256 - t : 255 - t);
    c2 y1[16], y2[16];
#pragma unroll
    for (int k2 = 0; k2 < 16; k2++) {
        y1[k2] = ldg_stream_c2(y + 256 * k2 + c);
        y2[k2] = ldg_stream_c2(y + 256 * k2 + cc);
    }
    tm::ld_wait32(wr);
    c2 w[16];
#pragma unroll
    for (int k2 = 0; k2 < 16; k2++) w[k2] = tm::pair_of(wr, k2);
    if (entry0) {
        c2 v1[16], v2[16];
        v1[0] = zc0;
#pragma unroll
        for (int k2 = 1; k2 <= 8; k2++) {
            c2 zk, zp;
            repack_pair(y1[k2], y1[16 - k2], w[k2], zk, zp);
            v1[k2] = zk;
            if (k2 != 8) v1[16 - k2] = zp;
        }
#pragma unroll
        for (int k2 = 0; k2 < 8; k2++) {
            c2 zk, zp;
            repack_pair(y2[k2], y2[15 - k2], k2 == 0 ? w[0] : w[8 + k2], zk, zp);
            v2[k2] = zk;
            v2[15 - k2] = zp;
        }
        Bfly<16>::template run<+1>(v1);
        Bfly<16>::template run<+1>(v2);
#pragma unroll
        for (int r = 0; r < 16; r++) {
            sm[out16(r)] = v1[r];
            sm[16 * 8 + out16(r)] = v2[r];
        }
        return;
    }
    c2 v1[16], v2[16];
#pragma unroll
    for (int k2 = 0; k2 < 16; k2++) repack_pair(y1[k2], y2[15 - k2], w[k2], v1[k2], v2[15 - k2]);
    Bfly<16>::template run<+1>(v1);
    Bfly<16>::template run<+1>(v2);
    c2 *p1 = sm + (c & 15) * ROW + (c >> 4) * 16;
    c2 *p2 = sm + (cc & 15) * ROW + (cc >> 4) * 16;
#pragma unroll
    for (int r = 0; r < 16; r++) {
        p1[out16(r)] = v1[r];
        p2[out16(r)] = v2[r];
    }
}
// the w[] of inv_pass_c_w / fwd_pass_c_w for run t (0..127) of half H, from the table
__device__ __forceinline__ void fill_twc(const Tables &tb, int H, int t, c2 (&w)[16]) {
    const float2 *twu = tb.twU + H * Q;
#pragma unroll
    for (int k2 = 0; k2 < 16; k2++) w[k2] = ldg_c2(twu + 256 * k2 + t);
    if (H == 0 && t == 0) {
        w[0] = ldg_c2(twu + 128);
#pragma unroll
        for (int k2 = 1; k2 < 8; k2++) w[8 + k2] = ldg_c2(twu + 256 * k2 + 128);
    }
}
__device__ __forceinline__ void inv_fill_twc(const Tables &tb, int j, c2 (&w)[16]) { fill_twc(tb, j >> 7, j & 127, w); }

// Pass A^-1 as inv_pass_a (256 threads, one column each) with the twiddles and the overlap tail in the thread's
// tensor-memory words: no table load, no tail load, no tail store.  The tensor-memory loads are issued one
// phase ahead of their use (the first twiddles while shared memory is read, the tail in front of the
// butterflies); the tail store stays in flight until the next block's tail load (wait_st in front of it).
template <int FMT>
__device__ __forceinline__ float inv_pass_a_tm(const c2 *sm, uint32_t tmem, void *dout, int nout, int o, int frames,
                                               bool full_path = false) {
    float lmax = 0.0f;
    const int u = threadIdx.x;
    const uint32_t fbytes = (uint32_t)nout * (FMT == PCM_S16 ? 2u : 4u);   // bytes per frame
    char *const pbase = reinterpret_cast<char *>(dout) + (size_t)(2 * u) * fbytes + (size_t)o * (FMT == PCM_S16 ? 2 : 4);
    uint32_t wr[32];
    tm::ld_issue32(tmem + tm::TWA, wr);
    c2 va[16], vb[16];
#pragma unroll
    for (int k0 = 0; k0 < 16; k0++) {
        va[k0] = sm[k0 * ROW + u];
        vb[k0] = sm[HALF_ELEMS + k0 * ROW + u];
    }
    tm::ld_wait32(wr);
#pragma unroll
    for (int k0 = 1; k0 < 16; k0++) va[k0] = c2_cmulconj(va[k0], tm::pair_of(wr, k0));
    tm::ld_issue32(tmem + tm::TWA + 32, wr);
    tm::ld_wait32(wr);
#pragma unroll
    for (int k0 = 0; k0 < 16; k0++) vb[k0] = c2_cmulconj(vb[k0], tm::pair_of(wr, k0));
    tm::wait_st();   // the previous block's tail store (or the fill)
    tm::ld_issue32(tmem + tm::TAIL, wr);
    Bfly<16>::template run<+1>(va);
    Bfly<16>::template run<+1>(vb);
    tm::ld_wait32(wr);
    c2 tl[16];
    // FULL: every frame of the block counts for the maximum (a whole block, the usual case): no per-sample test
    auto epilogue = [&](auto full) {
        constexpr bool FULL = decltype(full)::value;
#pragma unroll
        for (int r = 0; r < 16; r++) {
            const int n2 = out16(r), n = u + 256 * n2;
            const c2 b = n2 == 0 ? vb[r] : c2_cmulconj(vb[r], c2_pack(w32(n2)));
            const float2 s = c2_unpack(c2_add(va[r], b));
            const float2 t = c2_unpack(tm::pair_of(wr, r));
            const float y0 = s.x + t.x, y1 = s.y + t.y;
            tl[r] = c2_sub(va[r], b);
            const int f0 = 2 * n;
            char *p = pbase + (size_t)((uint32_t)(512 * n2) * fbytes);
            pcm_store_b<FMT>(p, y0);
            pcm_store_b<FMT>(p + fbytes, y1);
            if (FULL || f0 < frames) lmax = fmaxf(lmax, y0);
            if (FULL || f0 + 1 < frames) lmax = fmaxf(lmax, y1);
        }
    };
    if (full_path && frames >= N) epilogue(std::true_type{});
    else epilogue(std::false_type{});
    tm::tm_st<16>(tmem + tm::TAIL, tl);
    return lmax;
}
// the thread's words: tail from / to global memory (first / behind the last block), twiddles from the tables
__device__ __forceinline__ void inv_tm_fill(uint32_t tmem, const Tables &tb, const float2 *__restrict__ tail) {
    const int u = threadIdx.x;
    c2 v[16];
#pragma unroll
    for (int r = 0; r < 16; r++) {
        const float2 t = __ldcg(&tail[u + 256 * out16(r)]);
        v[r] = c2_pack(t.x, t.y);
    }
    tm::tm_st<16>(tmem + tm::TAIL, v);
    v[0] = 0ull;
#pragma unroll
    for (int k0 = 1; k0 < 16; k0++) v[k0] = ldg_c2(tb.twA0 + (k0 - 1) * 256 + u);
    tm::tm_st<16>(tmem + tm::TWA, v);
#pragma unroll
    for (int k0 = 0; k0 < 16; k0++) v[k0] = ldg_c2(tb.twA1 + k0 * 256 + u);
    tm::tm_st<16>(tmem + tm::TWA + 32, v);
    inv_fill_twc(tb, u, v);
    tm::tm_st<16>(tmem + tm::TWC, v);
    tm::wait_st();
}
__device__ __forceinline__ void inv_tm_save_tail(uint32_t tmem, float2 *__restrict__ tail) {
    const int u = threadIdx.x;
    c2 v[16];
    tm::wait_st();
    tm::tm_ld<16>(tmem + tm::TAIL, v);
#pragma unroll
    for (int r = 0; r < 16; r++) tail[u + 256 * out16(r)] = c2_unpack(v[r]);
}


// ---- forward transform with the thread's twiddles in tensor memory -----------------------------
// Stereo blocks of a batch: one CTA = one half of both channels' spectra of ONE stream for the T blocks of
// the step (the kernel without it: one CTA per block).  47 of a thread's 63 global loads per block are
// twiddles that do not change from block to block: pass A 16 (slot k0), pass B 15 (slot k1), unpack 16
// (slot k2, as in the inverse kernel) -- 96 words of the thread's 128.
namespace tm {
constexpr uint32_t F_TWA = 0, F_TWB = 32, F_TWC = 64;
}
template <int H>
__device__ __forceinline__ void fwd_tm_fill(uint32_t tmem, const Tables &tb) {
    const int u = threadIdx.x;
    c2 v[16];
    v[0] = 0ull;
    if (H == 0) {
#pragma unroll
        for (int k0 = 1; k0 < 16; k0++) v[k0] = ldg_c2(tb.twA0 + (k0 - 1) * 256 + u);
    } else {
#pragma unroll
        for (int k0 = 0; k0 < 16; k0++) v[k0] = ldg_c2(tb.twA1 + k0 * 256 + u);
    }
    tm::tm_st<16>(tmem + tm::F_TWA, v);
    v[0] = 0ull;
#pragma unroll
    for (int k1 = 1; k1 < 16; k1++) v[k1] = ldg_c2(tb.twB + (u >> 4) * 16 + k1);
    tm::tm_st<16>(tmem + tm::F_TWB, v);
    fill_twc(tb, H, u & 127, v);
    tm::tm_st<16>(tmem + tm::F_TWC, v);
    tm::wait_st();
}
// fwd_half<H, FMT, 2, 2, 256> with every table value from the thread's tensor-memory words; each load is issued
// in front of the phase's global / shared loads and waited for behind them.
template <int H, int FMT>
__device__ __forceinline__ void fwd_half_tm(c2 *sm, uint32_t tmem, const void *in, int fv, float2 *const (&rows)[2]) {
    const int u = threadIdx.x;
    uint32_t wr[32];
    {   // pass A
        tm::ld_issue32(tmem + tm::F_TWA, wr);
        c2 v[2][16];
#pragma unroll
        for (int j = 0; j < 16; j++) {
            c2 z[2];
            load_z<FMT, 2, 2>(in, 2, 0, u + 256 * j, fv, z);
            v[0][j] = z[0];
            v[1][j] = z[1];
        }
        tm::ld_wait32(wr);
#pragma unroll
        for (int c = 0; c < 2; c++) {
            if (H == 1) {
#pragma unroll
                for (int j = 1; j < 16; j++) v[c][j] = c2_cmul(v[c][j], c2_pack(w32(j)));
            }
            Bfly<16>::template run<-1>(v[c]);
#pragma unroll
            for (int r = 0; r < 16; r++) {
                const int k0 = out16(r);
                c2 x = v[c][r];
                if (H == 1 || k0 > 0) x = c2_cmul(x, tm::pair_of(wr, k0));
                sm[c * HALF_ELEMS + k0 * ROW + u] = x;
            }
        }
    }
    tm::ld_issue32(tmem + tm::F_TWB, wr);
    __syncthreads();
    {   // pass B
        const int k0 = u & 15, n0 = u >> 4;
        c2 *p = sm + k0 * ROW + n0;
        c2 v[16];
#pragma unroll
        for (int j = 0; j < 16; j++) v[j] = p[16 * j];
        tm::ld_wait32(wr);
#pragma unroll
        for (int a = 0; a < 2; a++) {
            if (a) {
                p += HALF_ELEMS;
#pragma unroll
                for (int j = 0; j < 16; j++) v[j] = p[16 * j];
            }
            Bfly<16>::template run<-1>(v);
#pragma unroll
            for (int r = 0; r < 16; r++) {
                const int k = out16(r);
                c2 x = v[r];
                if (k > 0) x = c2_cmul(x, tm::pair_of(wr, k));
                p[16 * k] = x;
            }
        }
    }
    tm::ld_issue32(tmem + tm::F_TWC, wr);
    __syncthreads();
    {   // pass C + unpack: thread u = run t = u % 128 of channel u / 128; the entry-0 threads (t = 0 of half 0)
        // read their own two runs with the same code and leave the common path behind the wait
        const c2 *smc = sm + (u >> 7) * HALF_ELEMS;
        c2 *out = reinterpret_cast<c2 *>(rows[u >> 7]) + H * Q;
        const int t = u & 127;
        const bool entry0 = H == 0 && t == 0;
        const int c = t, cc = entry0 ? 128 : (H == 0 ? 256 - t : 255 - t);
        c2 v1[16], v2[16];
        {
            const c2 *p1 = smc + (c & 15) * ROW + (c >> 4) * 16;
            const c2 *p2 = smc + (cc & 15) * ROW + (cc >> 4) * 16;
#pragma unroll
            for (int j = 0; j < 16; j++) {
                v1[j] = p1[j];
                v2[j] = p2[j];
            }
        }
        Bfly<16>::template run<-1>(v1);
        Bfly<16>::template run<-1>(v2);
        tm::ld_wait32(wr);
        if (entry0) {
            {
                const float2 z0 = c2_unpack(v1[reg16(0)]);
                out[0] = c2_pack(z0.x + z0.y, z0.x - z0.y);  // DC, Nyquist
            }
#pragma unroll
            for (int k2 = 1; k2 <= 8; k2++) {
                c2 xk, xp;
                unpack_pair(v1[reg16(k2)], v1[reg16(16 - k2)], tm::pair_of(wr, k2), xk, xp);
                out[256 * k2] = xk;
                if (k2 != 8) out[256 * (16 - k2)] = xp;
            }
#pragma unroll
            for (int k2 = 0; k2 < 8; k2++) {
                c2 xk, xp;
                unpack_pair(v2[reg16(k2)], v2[reg16(15 - k2)], tm::pair_of(wr, k2 == 0 ? 0 : 8 + k2), xk, xp);
                out[256 * k2 + 128] = xk;
                out[256 * (15 - k2) + 128] = xp;
            }
        } else {
#pragma unroll
            for (int k2 = 0; k2 < 16; k2++) {
                c2 xk, xp;
                unpack_pair(v1[reg16(k2)], v2[reg16(15 - k2)], tm::pair_of(wr, k2), xk, xp);
                out[256 * k2 + c] = xk;
                out[256 * (15 - k2) + cc] = xp;
            }
        }
    }
}


// ---- stereo pair: the two output channels of a stream as a cluster of two CTAs -------------
// A CTA owns ONE channel of an interleaved block, so on its own it can only store 2- or 4-byte
// scalars, every 32-byte sector of the block being written four times (timing-only ablation,
// profiles/r02_experiments.md: the inverse kernel without its PCM stores runs 20 % faster).
// Here the two CTAs (cluster rank = channel) put their converted samples into their own shared
// memory -- the transform buffers are free once pass A^-1 has them in registers -- and after a
// cluster barrier each writes HALF of the block's frames with both channels, as 16-byte vectors,
// reading the other channel through distributed shared memory.
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t peer_smem(const void *p, uint32_t rank) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p), r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
__device__ __forceinline__ uint2 ld_cluster_u2(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared::cluster.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
    return v;
}
template <int FMT>
__device__ __forceinline__ uint32_t pcm_word(float v) {   // pcm_store's conversions, as a register value
    if (FMT == PCM_F32) return __float_as_uint(v);
    if (FMT == PCM_S16) return (uint32_t)__float2int_rn(v * 32767.0f) & 0xffffu;
    return (uint32_t)__float2int_rn(v * 8388607.0f);
}

// Pass A^-1 + overlap-add + tail save + maximum as inv_pass_a, PCM through the pair exchange.
// o = this CTA's channel = its rank in the cluster; the block has exactly two channels.
// sm is read (transform result) and then reused as the exchange buffer; the caller must keep both
// CTAs from writing sm again until the second cluster barrier (inside) has been passed.
// hout (or null): the caller's pinned block (per-file path); it receives the same vectors as dout, but only the
// first `frames` frames (the reference writes back only the frames it read, sound-processor.cc:116-125).
template <int FMT, class BETWEEN>
__device__ __forceinline__ float inv_pass_a_pair(c2 *sm, const Tables &tb, float2 *__restrict__ tail, void *dout, void *hout,
                                                 int o, int frames, float &lmax_out, BETWEEN between) {
    float &lmax = lmax_out;
    lmax = 0.0f;
    const int u = threadIdx.x;   // 256 threads: one column each
    c2 va[16], vb[16];
#pragma unroll
    for (int k0 = 0; k0 < 16; k0++) {
        va[k0] = sm[k0 * ROW + u];
        vb[k0] = sm[HALF_ELEMS + k0 * ROW + u];
    }
    __syncthreads();   // the transform buffers are in registers everywhere: they become the exchange buffer
#pragma unroll
    for (int k0 = 1; k0 < 16; k0++) va[k0] = c2_cmulconj(va[k0], ldg_c2(tb.twA0 + (k0 - 1) * 256 + u));
#pragma unroll
    for (int k0 = 0; k0 < 16; k0++) vb[k0] = c2_cmulconj(vb[k0], ldg_c2(tb.twA1 + k0 * 256 + u));
    Bfly<16>::template run<+1>(va);
    Bfly<16>::template run<+1>(vb);
    float2 tl[16];
#pragma unroll
    for (int r = 0; r < 16; r++) tl[r] = __ldcg(&tail[u + 256 * out16(r)]);  // L2 only: read once per block
    uint32_t *x32 = reinterpret_cast<uint32_t *>(sm);   // 16-bit: one word per n (frames 2n, 2n+1 of this channel)
    uint2 *x64 = reinterpret_cast<uint2 *>(sm);         // 32-bit formats: two words per n
#pragma unroll
    for (int r = 0; r < 16; r++) {
        const int n2 = out16(r), n = u + 256 * n2;
        const c2 b = n2 == 0 ? vb[r] : c2_cmulconj(vb[r], c2_pack(w32(n2)));
        const float2 s = c2_unpack(c2_add(va[r], b));
        const float2 d = c2_unpack(c2_sub(va[r], b));
        const float y0 = s.x + tl[r].x, y1 = s.y + tl[r].y;
        tail[n] = d;
        const int f0 = 2 * n;
        if (FMT == PCM_S16) x32[n] = pcm_word<FMT>(y0) | (pcm_word<FMT>(y1) << 16);
        else x64[n] = make_uint2(pcm_word<FMT>(y0), pcm_word<FMT>(y1));
        if (f0 < frames) lmax = fmaxf(lmax, y0);
        if (f0 + 1 < frames) lmax = fmaxf(lmax, y1);
    }
    cluster_arrive();
    between();        // work that needs neither buffer (the block maximum)
    cluster_wait();   // both channels' samples are in place
    const uint32_t peer = peer_smem(sm, (uint32_t)(o ^ 1));
    uint4 *out = reinterpret_cast<uint4 *>(dout);
    if (FMT == PCM_S16) {
        // 16 bytes = frames 2n .. 2n+3 = columns n, n+1 of both channels
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int n = o * (Q / 2) + i * 512 + 2 * u;
            const uint2 own = *reinterpret_cast<const uint2 *>(x32 + n);
            const uint2 oth = ld_cluster_u2(peer + n * 4);
            const uint2 c0 = o == 0 ? own : oth, c1 = o == 0 ? oth : own;
            uint4 v;
            v.x = __byte_perm(c0.x, c1.x, 0x5410);
            v.y = __byte_perm(c0.x, c1.x, 0x7632);
            v.z = __byte_perm(c0.y, c1.y, 0x5410);
            v.w = __byte_perm(c0.y, c1.y, 0x7632);
            out[n / 2] = v;
            if (hout) {   // frames 2n .. 2n+3
                const int f0 = 2 * n;
                if (f0 + 4 <= frames) reinterpret_cast<uint4 *>(hout)[n / 2] = v;
                else {
                    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if (f0 + k < frames) reinterpret_cast<uint32_t *>(hout)[f0 + k] = w[k];
                }
            }
        }
    } else {
        // 16 bytes = frames 2n, 2n+1 = column n of both channels
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int n = o * (Q / 2) + i * 256 + u;
            const uint2 own = x64[n];
            const uint2 oth = ld_cluster_u2(peer + n * 8);
            const uint2 c0 = o == 0 ? own : oth, c1 = o == 0 ? oth : own;
            const uint4 v = make_uint4(c0.x, c1.x, c0.y, c1.y);
            out[n] = v;
            if (hout) {   // frames 2n, 2n+1
                const int f0 = 2 * n;
                if (f0 + 2 <= frames) reinterpret_cast<uint4 *>(hout)[n] = v;
                else if (f0 < frames) reinterpret_cast<uint2 *>(hout)[f0] = make_uint2(v.x, v.y);
            }
        }
    }
    cluster_arrive();   // this CTA has read the other's exchange buffer; the matching wait comes before the
    return lmax;        // next write to shared memory (inv_pass_c<.., true>) or before the CTA exits
}


// ---- forward pair: the two halves' CTAs of a stereo block share ONE read of the PCM -----------
// Each of the two CTAs (half 0 / half 1 of the spectrum) needs every frame of the block.  On the per-file
// path the block lies in the caller's pinned host memory: read by both CTAs it crosses the link twice.
// As a cluster, CTA r copies only frames [r N/2, (r+1) N/2) into its shared memory (wire format, as they
// are), and pass A takes the first half of its inputs from CTA 0's copy and the second from CTA 1's
// (ld.shared::cluster).  The arithmetic is that of fwd_pass_a: bit-identical spectra.
template <int FMT>
struct StageVec {   // two stereo frames in wire format
    static constexpr int BYTES = FMT == PCM_S16 ? 8 : 16;
};
constexpr size_t STAGE_BYTES_MAX = (size_t)(N / 2) * 16;   // the whole block as float32 / int32 frames

__device__ __forceinline__ uint4 ld_cluster_u4(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared::cluster.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}

// CTA `rank` copies its half of the block (frames < fv only) from `in` to its staging buffer.
template <int FMT>
__device__ __forceinline__ void stage_half(unsigned char *stage, const void *in, int rank, int fv) {
    constexpr int VB = StageVec<FMT>::BYTES;
    const int n0 = rank * (Q / 2);
#pragma unroll
    for (int i = 0; i < (Q / 2) / 256; i++) {
        const int n = n0 + i * 256 + threadIdx.x;
        if (VB == 16) {
            uint4 t = make_uint4(0u, 0u, 0u, 0u);
            if (2 * n < fv) t = __ldcs(reinterpret_cast<const uint4 *>(in) + n);
            reinterpret_cast<uint4 *>(stage)[n] = t;
        } else {
            uint2 t = make_uint2(0u, 0u);
            if (2 * n < fv) t = __ldcs(reinterpret_cast<const uint2 *>(in) + n);
            reinterpret_cast<uint2 *>(stage)[n] = t;
        }
    }
}

// fwd_pass_a for a stereo block (both channels, 256 threads) with the inputs taken from the pair's staging buffers.
template <int H, int FMT>
__device__ __forceinline__ void fwd_pass_a_staged(c2 *sm, const Tables &tb, uint32_t stage0, uint32_t stage1, int fv) {
    constexpr int VB = StageVec<FMT>::BYTES;
    const int u = threadIdx.x;
    c2 v[2][16];
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const int n = u + 256 * j;
        const uint32_t a = (j < 8 ? stage0 : stage1) + (uint32_t)n * VB;
        uint4 r;
        if (VB == 16) r = ld_cluster_u4(a);
        else {
            const uint2 t = ld_cluster_u2(a);
            r = make_uint4(t.x, t.y, 0u, 0u);
        }
        float2 p[2];
        raw_to_pairs<FMT>(r, p[0], p[1]);
#pragma unroll
        for (int c = 0; c < 2; c++) {
            if (2 * n >= fv) p[c].x = 0.0f;
            if (2 * n + 1 >= fv) p[c].y = 0.0f;
            v[c][j] = c2_pack(p[c].x, p[c].y);
        }
    }
    c2 w[16];
    if (H == 0) {
#pragma unroll
        for (int k0 = 1; k0 < 16; k0++) w[k0] = ldg_c2(tb.twA0 + (k0 - 1) * 256 + u);
    } else {
#pragma unroll
        for (int k0 = 0; k0 < 16; k0++) w[k0] = ldg_c2(tb.twA1 + k0 * 256 + u);
    }
#pragma unroll
    for (int c = 0; c < 2; c++) {
        if (H == 1) {
#pragma unroll
            for (int j = 1; j < 16; j++) v[c][j] = c2_cmul(v[c][j], c2_pack(w32(j)));
        }
        Bfly<16>::template run<-1>(v[c]);
#pragma unroll
        for (int r = 0; r < 16; r++) {
            const int k0 = out16(r);
            c2 x = v[c][r];
            if (H == 1 || k0 > 0) x = c2_cmul(x, w[k0]);
            sm[c * HALF_ELEMS + k0 * ROW + u] = x;
        }
    }
}

// One half of both channels' transforms, inputs from the staging buffers (the rest is fwd_half).
template <int H, int FMT>
__device__ __forceinline__ void fwd_half_staged(c2 *sm, const Tables &tb, uint32_t stage0, uint32_t stage1, int fv,
                                                float2 *const (&rows)[2]) {
    fwd_pass_a_staged<H, FMT>(sm, tb, stage0, stage1, fv);
    cluster_arrive();   // this CTA has read the other's staging buffer (waited for before the CTA exits)
    __syncthreads();
    pass_b<-1, 2, 256>(sm, tb);
    __syncthreads();
    const int j = threadIdx.x, c = j >> 7;
    fwd_pass_c<H>(sm + c * HALF_ELEMS, tb, c ? rows[1] : rows[0], j & 127);
    cluster_wait();
}

}  // namespace f13
}  // namespace fcv
