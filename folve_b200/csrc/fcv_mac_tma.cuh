// fcv_mac_tma.cuh -- time-tiled complex multiply-accumulate with the spectra streamed from
// HBM through bulk-async-copy (TMA, cp.async.bulk) staged tiles and an mbarrier pipeline.
//
// Same arithmetic and the same walk as mac_tt_kernel (fcv_mac.cuh): for T consecutive
// blocks of a stream the window of P+T-1 ring slots is read once, newest first, every X
// row feeds all T accumulators, the filter rows slide through a register window.  What
// changes is who moves the data: one producer warp issues 2 KB bulk copies (one per
// stream row tile and one for the filter row tile) into a ring of NS shared-memory stages,
// completion is signalled on an mbarrier per stage, and the four consumer warps only
// execute  try_wait / LDS.128 / FFMA2 / arrive.  Loads in flight no longer occupy
// registers or issue slots (no LDG, no L2 prefetch, no address arithmetic in the math
// warps), so the pipeline can run NS rows ahead of the arithmetic.
//
//   work item = (spectrum tile of 128 float4 = 2 KB, group of S streams, output), tile fastest
//   grid : persistent, 1-D: CTA c takes items c, c + gridDim.x, ...; the producer runs ahead of
//          the consumers across item boundaries, so the pipeline is filled once per CTA and the
//          next item's rows arrive while the consumers store the current item's Y
//   block: 160 threads = 4 consumer warps (one float4 column each) + 1 producer warp
//   smem : NS stages x (S + 1) x 2 KB
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "fcv_c2.cuh"
#include "fcv_mac.cuh"

namespace fcv {
namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared bulk copy, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ c2x2 lds_c2x2(const unsigned char *p) {
    c2x2 v;
    asm volatile("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(v.a), "=l"(v.b) : "r"(smem_u32(p)));
    return v;
}

constexpr int CONSUMER_WARPS = 4;
constexpr int TPB = 32 * CONSUMER_WARPS;   // float4 columns per tile
constexpr int THREADS = TPB + 32;          // + producer warp
constexpr int TILE_BYTES = TPB * 16;

__host__ __device__ constexpr size_t smem_bytes(int S, int NS) { return (size_t)NS * (S + 1) * TILE_BYTES; }

// zero_row: index of an all-zero filter row (pairs without data in a partition)
// CTAs per SM the register allocation is held to (T * S accumulators + T window entries)
__host__ __device__ constexpr int min_ctas(int T, int S) { return T * S >= 16 ? 3 : 4; }

template <int T, int S, int NS, int MC = min_ctas(T, S)>
__global__ void __launch_bounds__(THREADS, MC)
mac_tma_kernel(const StreamDev *__restrict__ st, int nstreams, const TTPair *__restrict__ pairs,
               const int *__restrict__ pair_off, const int *__restrict__ tt_rows,
               const float4 *__restrict__ H, float4 *__restrict__ Y, int M4, int P, int R, int newest_slot,
               int nout, int zero_row, int ntiles, int ngroups, int nitems) {
    constexpr int STAGE_BYTES = (S + 1) * TILE_BYTES;
    extern __shared__ __align__(128) unsigned char stages[];
    __shared__ uint64_t full[NS], empty[NS];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NS; i++) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], CONSUMER_WARPS);
        }
        fence_barrier_init();
    }
    __syncthreads();

    const int D = P + T - 1;
    const size_t rowb = (size_t)M4 * 16;

    if (warp == CONSUMER_WARPS) {
        // ---- producer: one lane walks the same (item, pair, step) sequence as the consumers
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const int tile = item % ntiles, b0 = (item / ntiles) % ngroups * S, o = item / (ntiles * ngroups);
            const int p0 = pair_off[o], p1 = pair_off[o + 1];
            const unsigned char *xbase[S];
#pragma unroll
            for (int s = 0; s < S; s++) {
                const int b = min(b0 + s, nstreams - 1);
                xbase[s] = reinterpret_cast<const unsigned char *>(st[b].xring) + (size_t)tile * TILE_BYTES;
            }
            const unsigned char *hbase = reinterpret_cast<const unsigned char *>(H) + (size_t)tile * TILE_BYTES;
            for (int p = p0; p < p1; p++) {
                const int inp = pairs[p].inp;
                const int *rows = tt_rows + pairs[p].rowbase;
                int slot = newest_slot;
                int rown = rows[0];  // filter row of the next step, fetched one step early
                for (int d = 0; d < D; d++) {
                    const int rowd = rown;
                    if (d + 1 < P) rown = rows[d + 1];
                    mbar_wait(&empty[stage], phase ^ 1);
                    unsigned char *dst = stages + (size_t)stage * STAGE_BYTES;
                    mbar_expect_tx(&full[stage], (uint32_t)((d < P ? S + 1 : S) * TILE_BYTES));
                    const size_t xoff = ((size_t)inp * R + slot) * rowb;
#pragma unroll
                    for (int s = 0; s < S; s++) bulk_g2s(dst + s * TILE_BYTES, xbase[s] + xoff, TILE_BYTES, &full[stage]);
                    if (d < P) {
                        const int row = rowd < 0 ? zero_row : rowd;
                        bulk_g2s(dst + S * TILE_BYTES, hbase + (size_t)row * rowb, TILE_BYTES, &full[stage]);
                    }
                    slot = slot == 0 ? R - 1 : slot - 1;
                    if (++stage == NS) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
            }
        }
        return;
    }

    // ---- consumers
    int stage = 0;
    uint32_t phase = 0;
    const unsigned char *mine = stages + (size_t)threadIdx.x * 16;
#pragma unroll 1
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int tile = item % ntiles, b0 = (item / ntiles) % ngroups * S, o = item / (ntiles * ngroups);
    const int p0 = pair_off[o], p1 = pair_off[o + 1];
    c2 acc[T][S][2];
#pragma unroll
    for (int t = 0; t < T; t++)
#pragma unroll
        for (int s = 0; s < S; s++) acc[t][s][0] = acc[t][s][1] = 0ull;

    for (int p = p0; p < p1; p++) {
        c2x2 hw[T];
#pragma unroll
        for (int t = 0; t < T; t++) hw[t].a = hw[t].b = 0ull;
        // T steps d0 .. d0+T-1; GUARD: the chunk contains steps where some outputs have no
        // partition (j < 0 at the head of the window, j >= P at its tail)
        auto chunk = [&](int d0, auto guard) {
            constexpr bool GUARD = decltype(guard)::value;
#pragma unroll
            for (int r = 0; r < T; r++) {
                const int d = d0 + r;
                if (!GUARD || d < D) {
                    mbar_wait(&full[stage], phase);
                    const unsigned char *sp = mine + (size_t)stage * STAGE_BYTES;
                    c2x2 x[S];
#pragma unroll
                    for (int s = 0; s < S; s++) x[s] = lds_c2x2(sp + s * TILE_BYTES);
                    c2x2 h;
                    h.a = h.b = 0ull;
                    if (!GUARD || d < P) h = lds_c2x2(sp + S * TILE_BYTES);
                    hw[(T - 1 + r) % T] = h;  // H[d]: the newest output (t = T-1) starts on partition j = d
#pragma unroll
                    for (int t = 0; t < T; t++) {
                        const int j = d - (T - 1) + t;
                        if (!GUARD || (j >= 0 && j < P)) {
                            const c2x2 hh = hw[(t + r) % T];
#pragma unroll
                            for (int s = 0; s < S; s++) {
                                acc[t][s][0] = c2_cmac(acc[t][s][0], x[s].a, hh.a);
                                acc[t][s][1] = c2_cmac(acc[t][s][1], x[s].b, hh.b);
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[stage]);
                    if (++stage == NS) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        };
        int d0 = 0;
        chunk(d0, std::true_type{});
#pragma unroll 1
        for (d0 = T; d0 + T <= P; d0 += T) chunk(d0, std::false_type{});
#pragma unroll 1
        for (; d0 < D; d0 += T) chunk(d0, std::true_type{});
    }
    const int e4 = tile * TPB + threadIdx.x;
#pragma unroll
    for (int t = 0; t < T; t++)
#pragma unroll
        for (int s = 0; s < S; s++)
            if (b0 + s < nstreams) {
                c2x2 v;
                v.a = acc[t][s][0];
                v.b = acc[t][s][1];
                __stcs(Y + (((size_t)(b0 + s) * nout + o) * T + t) * (size_t)M4 + e4, c2x2_to(v));
            }
    }
}

}  // namespace tma
}  // namespace fcv
