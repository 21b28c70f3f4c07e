// fcv_mac.cuh -- frequency-domain complex multiply-accumulate over the partition
// history and the input x output matrix: the roofline kernel.
//
// Role in the reference path: the inner loops of zita-convolver's
// Convlevel::process() (for each output, for each MAC node, for j < npar:
// freq_data[k] += ffta[ptind - j][k] * fftb[j][k]), reached from
// SoundProcessor::Process() at /root/reference/sound-processor.cc:113.
//
//   Y[s][o][e] = sum over (i, j) with H[i][o][j] present of
//                X[s][i][(pt - j) mod P][e] * H[i][o][j][e]
//
// e runs over the M = fragm entries of the packed-permuted spectrum layout
// (fcv_fft.cuh).  Entry 0 holds two REAL bins (DC, Nyquist); this kernel
// treats it as complex like every other entry and the inverse-FFT kernel
// overwrites it with the correct real products, so there is no special case
// in the streaming loop.
//
// Pure streaming: every X row is read from HBM exactly once per launch; the
// filter rows H are shared by all streams and stay L2 resident; S streams per
// thread reuse each H value from registers.  0.5-1 flop per byte -> HBM bound,
// no tensor cores.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fcv_c2.cuh"
#include "fcv_stream_dev.cuh"
#include "fcv_types.h"

namespace fcv {

// Two complex multiply-accumulates on a 16-byte vector: 4 FFMA2 (fcv_c2.cuh).
__device__ __forceinline__ void cmac2(float4 &acc, const float4 x, const float4 h) {
    const c2 a = c2_cmac(c2_pack(acc.x, acc.y), c2_pack(x.x, x.y), c2_pack(h.x, h.y));
    const c2 b = c2_cmac(c2_pack(acc.z, acc.w), c2_pack(x.z, x.w), c2_pack(h.z, h.w));
    const float2 af = c2_unpack(a), bf = c2_unpack(b);
    acc = make_float4(af.x, af.y, bf.x, bf.y);
}

// Streaming (read-once) 16-byte load: keep it out of L1 so that the shared
// filter rows (read through ld_keep) stay cached.
__device__ __forceinline__ float4 ld_stream(const float4 *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ld_keep(const float4 *p) {
    return __ldg(p);
}

// DC and Nyquist are real bins sharing entry 0 of a spectrum: their products are two real
// multiply-accumulates over the partition history (the MAC kernels treat entry 0 as one
// complex value, which the inverse transform ignores).  One warp per (stream, output, block):
// lanes split the partitions; returns, on every lane, entry 0 of the complex N-point
// sequence the inverse transform starts from: (dc + ny, dc - ny).  One code path for every
// kernel that needs it, so that all block sizes / tilings round identically.
__device__ __forceinline__ float2 dcny_warp(const float2 *__restrict__ xring, const TTPair *__restrict__ pairs,
                                            const int *__restrict__ pair_off, const int *__restrict__ tt_rows,
                                            const float2 *__restrict__ H, int o, int P, int R, int newest, int M,
                                            int lane) {
    float dc = 0.f, ny = 0.f;
    for (int p = pair_off[o]; p < pair_off[o + 1]; p++) {
        const int inp = pairs[p].inp;
        const int *rows = tt_rows + pairs[p].rowbase;
        for (int j = lane; j < P; j += 32) {
            const int row = rows[j];
            if (row >= 0) {
                int slot = newest - j;
                if (slot < 0) slot += R;
                const float2 x = xring[(size_t)(inp * R + slot) * M];
                const float2 h = H[(size_t)row * M];
                dc = fmaf(x.x, h.x, dc);
                ny = fmaf(x.y, h.y, ny);
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        dc += __shfl_xor_sync(0xffffffffu, dc, d);
        ny += __shfl_xor_sync(0xffffffffu, ny, d);
    }
    return make_float2(dc + ny, dc - ny);
}

// The partition walk of ONE stream for one 16-byte column (the per-file path; S = 1 in mac_kernel
// and the MAC phase of the fused single-stream kernel): a lone stream is latency bound -- a chain
// of dependent round trips, step table -> X row / filter rows -> FMA -- so all loads of U steps
// are issued before the first of them is used.  acc[o][0] += X[inp][pt - part] * H[row[o]].
template <int NO, int S>
__device__ __forceinline__ void mac_single_steps(float4 (&acc)[NO][S], const float4 *xb, int pt,
                                                 const MacStep *__restrict__ steps, int t0, int t1,
                                                 const float4 *__restrict__ H, int M4, int P, int e4) {
    constexpr int U = NO <= 2 ? 8 : 4;
#pragma unroll 1
    for (int t = t0; t < t1; t += U) {
        float4 x[U], h[U][NO];
        int row[U][NO];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const bool live = t + u < t1;
            const MacStep *sp = steps + (live ? t + u : t);
            const int inp = __ldg(&sp->inp), part = __ldg(&sp->part);
            int slot = pt - part;
            if (slot < 0) slot += P;
            x[u] = ld_stream(xb + (size_t)(inp * P + slot) * (size_t)M4);
#pragma unroll
            for (int o = 0; o < NO; o++) row[u][o] = live ? __ldg(&sp->row[o]) : -1;
        }
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int o = 0; o < NO; o++)
                h[u][o] = row[u][o] >= 0 ? ld_keep(H + (size_t)row[u][o] * (size_t)M4 + e4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int o = 0; o < NO; o++)
                if (row[u][o] >= 0) cmac2(acc[o][0], x[u], h[u][o]);
    }
}

// grid: x = M/2/TPB tiles of 2*TPB entries + 1 DC/Nyquist column, y = ceil(nstreams/S), z = output groups.
// SEL: how the streams of the launch are addressed (fcv_stream_dev.cuh); every stream carries its
// own ring position, Y rows and entry-0 slot.
template <class SEL, int NO, int S, int TPB>
__global__ void __launch_bounds__(TPB)
mac_kernel(const __grid_constant__ SEL sel, int nstreams, const MacStep *__restrict__ steps,
           const int *__restrict__ group_off, const float4 *__restrict__ H,
           int M4, int P, int nout, const TTPair *__restrict__ pairs, const int *__restrict__ pair_off,
           const int *__restrict__ tt_rows, int Pfilt) {
    pdl_trigger();
    pdl_wait();
    const int e4 = blockIdx.x * TPB + threadIdx.x;
    const int b0 = blockIdx.y * S;
    const int g = blockIdx.z;

    // one extra column of CTAs (blockIdx.x == gridDim.x - 1) does the DC / Nyquist products of its
    // streams and outputs beside the spectrum tiles (block-by-block path: no separate launch)
    if (blockIdx.x == gridDim.x - 1) {
        if (threadIdx.x < 32) {
            for (int o = 0; o < NO; o++) {
                const int oo = g * NO + o;
                if (oo >= nout) break;
                for (int s = 0; s < S && b0 + s < nstreams; s++) {
                    const StreamDev sd = sel.stream(b0 + s);
                    const float2 z = dcny_warp(sd.xring, pairs, pair_off, tt_rows, reinterpret_cast<const float2 *>(H),
                                               oo, Pfilt, P, sel.slot(b0 + s), 2 * M4, threadIdx.x);
                    if (threadIdx.x == 0) sd.zc0[oo] = z;
                }
            }
        }
        return;
    }

    const float4 *xb[S];
    float4 *yb[S];
    int pts[S];
#pragma unroll
    for (int s = 0; s < S; s++) {
        const int b = min(b0 + s, nstreams - 1);
        const StreamDev sd = sel.stream(b);
        xb[s] = reinterpret_cast<const float4 *>(sd.xring) + e4;
        yb[s] = reinterpret_cast<float4 *>(sd.Y) + e4;
        pts[s] = sel.slot(b);
    }
    float4 acc[NO][S];
#pragma unroll
    for (int o = 0; o < NO; o++)
#pragma unroll
        for (int s = 0; s < S; s++) acc[o][s] = make_float4(0.f, 0.f, 0.f, 0.f);

    const int t0 = group_off[g], t1 = group_off[g + 1];
    if (S == 1) {
        mac_single_steps<NO, S>(acc, xb[0], pts[0], steps, t0, t1, H, M4, P, e4);
    } else {
#pragma unroll 2
    for (int t = t0; t < t1; t++) {
        const MacStep *sp = steps + t;
        const int inp = __ldg(&sp->inp), part = __ldg(&sp->part);
        float4 x[S];
#pragma unroll
        for (int s = 0; s < S; s++) {
            int slot = pts[s] - part;
            if (slot < 0) slot += P;
            x[s] = ld_stream(xb[s] + (size_t)(inp * P + slot) * (size_t)M4);
        }
#pragma unroll
        for (int o = 0; o < NO; o++) {
            const int row = __ldg(&sp->row[o]);
            if (row >= 0) {
                const float4 h = ld_keep(H + (size_t)row * (size_t)M4 + e4);
#pragma unroll
                for (int s = 0; s < S; s++) cmac2(acc[o][s], x[s], h);
            }
        }
    }
    }
#pragma unroll
    for (int o = 0; o < NO; o++) {
        const int oo = g * NO + o;
        if (oo < nout) {
#pragma unroll
            for (int s = 0; s < S; s++) {
                if (b0 + s < nstreams) __stcs(yb[s] + (size_t)oo * (size_t)M4, acc[o][s]);
            }
        }
    }
}

// ---- the launch groups of the per-file path ------------------------------------------------
// One CTA per (spectrum tile, OUTPUT, stream): a thread walks only the partitions of the inputs that
// feed its output -- X row, filter row, 4 FFMA2 per step, addresses by increments -- instead of the
// step table shared by the outputs of a group.  A lone stream is bound by the length of one warp's
// instruction stream (one warp per scheduler: ~10 cycles per instruction; mac_kernel<.., 2, 1, 128>
// executes 3000 instructions per warp for SantaLucia, 15 us): here it is ~600, on twice as many SMs.
// Per output the additions happen in the order of mac_kernel's step table (inputs ascending,
// partitions ascending, absent rows skipped), so the results are bit-identical.
// grid: x = M/2/TPB tiles + 1 DC/Nyquist column, y = outputs, z = streams of the launch.
template <class SEL, int TPB>
__global__ void __launch_bounds__(TPB)
mac_group_kernel(const __grid_constant__ SEL sel, const float4 *__restrict__ H, int M4, int R, const TTPair *__restrict__ pairs,
                 const int *__restrict__ pair_off, const int *__restrict__ tt_rows, int Pfilt) {
    pdl_trigger();
    pdl_wait();
    const int o = blockIdx.y, b = blockIdx.z;
    const StreamDev sd = sel.stream(b);
    const int pt = sel.slot(b);
    if (blockIdx.x == gridDim.x - 1) {   // DC / Nyquist products of this (stream, output)
        if (threadIdx.x < 32) {
            const float2 z = dcny_warp(sd.xring, pairs, pair_off, tt_rows, reinterpret_cast<const float2 *>(H), o, Pfilt, R,
                                       pt, 2 * M4, threadIdx.x);
            if (threadIdx.x == 0) sd.zc0[o] = z;
        }
        return;
    }
    const int e4 = blockIdx.x * TPB + threadIdx.x;
    const float4 *hb = H + e4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    constexpr int U = 12;   // rows in flight together: SantaLucia's 22 partitions in two round trips
    const int p1 = pair_off[o + 1];
    for (int p = pair_off[o]; p < p1; p++) {
        const int *rows = tt_rows + __ldg(&pairs[p].rowbase);
        const float4 *xin = reinterpret_cast<const float4 *>(sd.xring) + (size_t)__ldg(&pairs[p].inp) * R * (size_t)M4 + e4;
#pragma unroll 1
        for (int j0 = 0; j0 < Pfilt; j0 += U) {
            int row[U];
            float4 x[U], h[U];
#pragma unroll
            for (int u = 0; u < U; u++) row[u] = j0 + u < Pfilt ? __ldg(rows + j0 + u) : -1;
#pragma unroll
            for (int u = 0; u < U; u++) {
                int slot = pt - (j0 + u);
                if (slot < 0) slot += R;
                if (row[u] >= 0) {
                    x[u] = ld_stream(xin + (size_t)slot * (size_t)M4);
                    h[u] = ld_keep(hb + (size_t)row[u] * (size_t)M4);
                }
            }
#pragma unroll
            for (int u = 0; u < U; u++)
                if (row[u] >= 0) cmac2(acc, x[u], h[u]);
        }
    }
    __stcs(reinterpret_cast<float4 *>(sd.Y) + (size_t)o * (size_t)M4 + e4, acc);
}

// ---- time-tiled variant ------------------------------------------------------------
// When T consecutive blocks of a stream are available at once (prebuffered files),
// output block t0+t needs ring slots t0+t-j, j < P: the T outputs share all but
// T-1 of their P input rows.  This kernel walks the window of P+T-1 slots ONCE,
// newest first, and feeds every loaded X row to all T accumulators; the filter
// rows slide through a circular register window (one new H row per X row).
// HBM traffic per block drops from P rows to (P+T-1)/T rows per input.
//
//   step d = 0 .. P+T-2 :  X row of block u = t0+T-1-d ;  output t uses H[j = t-(T-1)+d]
//   H[j] for output t at step d lives in window register (t + d) mod T.

__device__ __forceinline__ void prefetch_l2(const void *p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

#ifndef FCV_TT_PREFETCH
#define FCV_TT_PREFETCH 6  // X rows requested into L2 this many steps ahead of their use
#endif

template <int T, int S, int TPB>
__global__ void __launch_bounds__(TPB)
mac_tt_kernel(const StreamDev *__restrict__ st, int nstreams, const TTPair *__restrict__ pairs,
              const int *__restrict__ pair_off, const int *__restrict__ tt_rows,
              const float4 *__restrict__ H, float4 *__restrict__ Y, int M4, int P, int R, int newest_slot,
              int nout) {
    const int e4 = blockIdx.x * TPB + threadIdx.x;
    const int b0 = blockIdx.y * S;
    const int o = blockIdx.z;

    const float4 *xb[S];
#pragma unroll
    for (int s = 0; s < S; s++) {
        const int b = min(b0 + s, nstreams - 1);
        xb[s] = reinterpret_cast<const float4 *>(st[b].xring) + e4;
    }
    float4 acc[T][S];
#pragma unroll
    for (int t = 0; t < T; t++)
#pragma unroll
        for (int s = 0; s < S; s++) acc[t][s] = make_float4(0.f, 0.f, 0.f, 0.f);

    const int D = P + T - 1;
    for (int p = pair_off[o]; p < pair_off[o + 1]; p++) {
        const int inp = __ldg(&pairs[p].inp);
        const int *rows = tt_rows + __ldg(&pairs[p].rowbase);
        const size_t xin = (size_t)inp * R;
        float4 hw[T];
#pragma unroll
        for (int t = 0; t < T; t++) hw[t] = make_float4(0.f, 0.f, 0.f, 0.f);
        // software pipeline, one step deep: the X rows and the filter row of step
        // d+1 are requested before the multiply-accumulates of step d
        int slot = newest_slot;
        // L2 prefetch runs FCV_TT_PREFETCH rows ahead of the register pipeline, so that
        // the loads below mostly hit L2 and few bytes need to be in flight per thread
        int pslot = newest_slot;
#pragma unroll
        for (int k = 0; k < FCV_TT_PREFETCH; k++) {
            if (k > 0 && k < D) {
#pragma unroll
                for (int s = 0; s < S; s++) prefetch_l2(xb[s] + (xin + pslot) * (size_t)M4);
            }
            pslot = pslot == 0 ? R - 1 : pslot - 1;
        }
        float4 xn[S], hn;
#pragma unroll
        for (int s = 0; s < S; s++) xn[s] = ld_stream(xb[s] + (xin + slot) * (size_t)M4);
        {
            const int row = __ldg(&rows[0]);
            hn = row >= 0 ? ld_keep(H + (size_t)row * (size_t)M4 + e4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int d0 = 0; d0 < D; d0 += T) {
#pragma unroll
            for (int r = 0; r < T; r++) {
                const int d = d0 + r;
                if (d < D) {
                    float4 x[S];
#pragma unroll
                    for (int s = 0; s < S; s++) x[s] = xn[s];
                    hw[(T - 1 + r) % T] = hn;  // the newest output (t = T-1) starts on partition j = d
                    slot = slot == 0 ? R - 1 : slot - 1;
                    if (d + FCV_TT_PREFETCH < D) {
#pragma unroll
                        for (int s = 0; s < S; s++) prefetch_l2(xb[s] + (xin + pslot) * (size_t)M4);
                    }
                    pslot = pslot == 0 ? R - 1 : pslot - 1;
                    if (d + 1 < D) {
#pragma unroll
                        for (int s = 0; s < S; s++) xn[s] = ld_stream(xb[s] + (xin + slot) * (size_t)M4);
                        const int row = d + 1 < P ? __ldg(&rows[d + 1]) : -1;
                        hn = row >= 0 ? ld_keep(H + (size_t)row * (size_t)M4 + e4) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int t = 0; t < T; t++)
#pragma unroll
                        for (int s = 0; s < S; s++) cmac2(acc[t][s], x[s], hw[(t + r) % T]);
                }
            }
        }
    }
#pragma unroll
    for (int t = 0; t < T; t++)
#pragma unroll
        for (int s = 0; s < S; s++)
            if (b0 + s < nstreams)
                __stcs(Y + (((size_t)(b0 + s) * nout + o) * T + t) * (size_t)M4 + e4, acc[t][s]);
}


}  // namespace fcv
