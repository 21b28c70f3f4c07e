// fcv_mac_tma.cuh -- time-tiled complex multiply-accumulate with the spectra streamed from
// HBM through bulk-async-copy (TMA, cp.async.bulk) staged tiles and an mbarrier pipeline.
//
// Same arithmetic and the same walk as mac_tt_kernel (fcv_mac.cuh): for T consecutive
// blocks of a stream the window of P+T-1 ring slots is read once, newest first, every X
// row feeds all T accumulators, the filter rows slide through a register window.  What
// changes is who moves the data: one producer warp issues 2 KB bulk copies (one per
// stream row tile and one for the filter row tile) into a ring of NS shared-memory stages
// (G lanes per pass, a row each: template parameter G below), completion is signalled on an
// mbarrier per stage, and the four consumer warps only execute
// try_wait / LDS.128 / FFMA2 / elected arrive -- straight-line code without per-output
// predicates, 81 instructions per warp and row of which 64 are FFMA2 (see the consumer
// section).  Loads in flight no longer occupy registers or issue slots (no LDG, no L2
// prefetch, no address arithmetic in the math warps), so the pipeline runs NS - G rows
// ahead of the arithmetic.
//
//   work item = (spectrum tile of 128 float4 = 2 KB, group of S streams, output), OUTPUT fastest:
//          the CTAs that run side by side are the outputs of one (tile, stream group).  Where
//          several outputs are fed from the same input (crossfeed, 5.1, dense matrices) they read
//          the same X row tiles at the same time, so all but the first read hit L2 instead of
//          HBM; for diagonal filters the order makes no difference.
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
// The same operations on shared-window addresses held in registers.  The addresses are made
// opaque to the compiler once per kernel (hold_u32): at the 128-register cap ptxas otherwise
// rematerialises them in every row (S2R SR_CgaCtaId / SR_TID + LEA chains in front of each
// try_wait and arrive).
__device__ __forceinline__ uint32_t hold_u32(uint32_t v) {
    uint32_t r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
// one elected lane of the (converged) warp arrives; elect.sync is also the warp-level
// rendezvous that orders every lane's shared-memory reads of the stage before the release.
// Predicated, so the row loop carries no divergent branch.
__device__ __forceinline__ void mbar_arrive_elect_a(uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "@p mbarrier.arrive.shared::cta.b64 _, [%0];\n\t}" ::"r"(bar)
        : "memory");
}
__device__ __forceinline__ c2x2 lds_c2x2_a(uint32_t p) {
    c2x2 v;
    asm volatile("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(v.a), "=l"(v.b) : "r"(p));
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

// G: how the producer warp hands out the rows of the window.
//   G > 0: passes of G rows, lane l of a pass takes row g0 + l (stage (g0 + l) % NS).  The lanes of
//          a pass reconverge behind their waits (ptxas brackets the try_wait spin with BSSY /
//          BSYNC), so a pass issues its copies when the LAST of its G stages has been released:
//          with G = NS the ring is refilled only once it has drained completely (no copy of a CTA
//          in flight while its consumers work); with G < NS the copies run NS - G rows ahead of
//          the arithmetic at G-fold amortised address arithmetic.
//   G = 0: the addresses of NS rows are computed lane-parallel, then the rows are issued in order,
//          one lane at a time, each as soon as its own stage is released (NS - 1 rows ahead).
template <int T, int S, int NS, int MC = min_ctas(T, S), int G = NS>
__global__ void __launch_bounds__(THREADS, MC)
mac_tma_kernel(const float2 *__restrict__ xring0, size_t xring_stride, int nstreams, const TTPair *__restrict__ pairs,
               const int *__restrict__ pair_off, const int *__restrict__ tt_rows,
               const float4 *__restrict__ H, float4 *__restrict__ Y, int M4, int P, int R, int newest_slot,
               int nout, int zero_row, int ntiles, int ngroups, int nitems) {
    constexpr int STAGE_BYTES = (S + 1) * TILE_BYTES;
    extern __shared__ __align__(128) unsigned char stages[];
    __shared__ uint64_t bars[2 * NS];
    uint64_t *const full = bars, *const empty = bars + NS;

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
        // ---- producer: GL lanes of the warp take the GL rows of one pass, one row each -- every
        // lane waits for its own stage to be released and issues that row's copies (the lanes of a
        // pass leave their waits together, see G above).  The address arithmetic and the barrier
        // hand-shake are executed once per GL rows instead of once per row: the producer shares its
        // scheduler with consumer warp 0, and as a single lane walking the rows (~90 instructions
        // each) it, not HBM, paced the kernel.  Row d of every (input, output) pair lives in stage
        // d % NS and is issued by lane d % GL, so a stage is always served by the same lane, which
        // keeps that stage's phase bit; the consumers index the stages of a T-row chunk with
        // compile-time constants when NS divides T.
        static_assert(NS <= 32, "one lane per stage");
        static_assert(G >= 0 && G <= NS && (G == 0 || NS % G == 0), "lanes per pass: a stage is always served by the same lane");
        constexpr int GL = G > 0 ? G : NS;   // rows whose addresses one pass computes
        uint32_t ph = 0;   // bit s: parity of the last fill of stage s (a lane only ever looks at the stages it serves)
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const int o = item % nout, tile = (item / nout) % ntiles, b0 = item / (nout * ntiles) * S;
            const int p0 = pair_off[o], p1 = pair_off[o + 1];
            const unsigned char *xbase[S];
#pragma unroll
            for (int s = 0; s < S; s++) {
                const int b = min(b0 + s, nstreams - 1);
                // the streams of a batch lie in one slab: no descriptor load in front of the first copies
                xbase[s] = reinterpret_cast<const unsigned char *>(xring0 + (size_t)b * xring_stride) + (size_t)tile * TILE_BYTES;
            }
            const unsigned char *hbase = reinterpret_cast<const unsigned char *>(H) + (size_t)tile * TILE_BYTES;
            for (int p = p0; p < p1; p++) {
                const int inp = pairs[p].inp;
                const int *rows = tt_rows + pairs[p].rowbase;
                for (int g0 = 0; g0 < D; g0 += GL) {
                    const int d = g0 + lane;
                    const bool has_row = lane < GL && d < D;
                    const int stage = d % NS;
                    int slot = newest_slot - d;   // d < D = R: at most one wrap
                    if (slot < 0) slot += R;
                    int row = (has_row && d < P) ? rows[d] : 0;
                    if (row < 0) row = zero_row;
                    unsigned char *dst = stages + (size_t)stage * STAGE_BYTES;
                    const size_t xoff = ((size_t)inp * R + slot) * rowb;
                    auto issue = [&]() {
                        mbar_wait(&empty[stage], ((ph >> stage) & 1) ^ 1);
                        ph ^= 1u << stage;
                        mbar_expect_tx(&full[stage], (uint32_t)((d < P ? S + 1 : S) * TILE_BYTES));
#pragma unroll
                        for (int s = 0; s < S; s++) bulk_g2s(dst + s * TILE_BYTES, xbase[s] + xoff, TILE_BYTES, &full[stage]);
                        if (d < P) bulk_g2s(dst + S * TILE_BYTES, hbase + (size_t)row * rowb, TILE_BYTES, &full[stage]);
                    };
                    if (G > 0) {
                        if (has_row) issue();
                    } else {
#pragma unroll
                        for (int i = 0; i < NS; i++)
                            if (lane == i && has_row) issue();
                    }
                    __syncwarp();
                }
            }
        }
        return;
    }

    // ---- consumers
    uint32_t phases = 0;   // bit s: parity of the next fill of stage s
    constexpr bool FIXED = T % NS == 0;   // the stage of row r of a chunk is the constant r % NS
    const uint32_t mine = hold_u32(smem_u32(stages) + threadIdx.x * 16);
    const uint32_t bar0 = hold_u32(smem_u32(bars));
#pragma unroll 1
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int o = item % nout, tile = (item / nout) % ntiles, b0 = item / (nout * ntiles) * S;
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
        // Rows d0 .. d0+NR-1 of the window (NR <= T, a compile-time count: straight-line code, one
        // entry, one exit -- an early exit per row makes ptxas re-home every accumulator in front
        // of each exit branch, ~90 MOVs per row).
        // HEAD (d0 = 0): output t starts on partition 0 at row T-1-t, so row r feeds the outputs
        // t >= T-1-r -- a compile-time staircase.  Otherwise every row feeds every output: the
        // outputs whose partition index j = d-(T-1)+t has run past P-1 are NOT skipped, their window
        // entries are the zeros that rows d >= P put there (h = 0 below), so they add x * 0 to an
        // accumulator that is never -0 -- bit-identical to skipping them, and much cheaper than the
        // per-output ISETP / BRA pairs that skipping costs (21 of SantaLucia's 29 rows were such
        // guarded rows: 150 against 89 instructions of an unguarded row).
        auto chunk = [&](int d0, auto head, auto nrows) {
            constexpr bool HEAD = decltype(head)::value;
            constexpr int NR = decltype(nrows)::value;
#pragma unroll
            for (int r = 0; r < NR; r++) {
                const int d = d0 + r;
                const int stage = FIXED ? r % NS : (d0 + r) % NS;
                mbar_wait_a(bar0 + 8 * stage, (phases >> stage) & 1);
                const uint32_t sp = mine + stage * STAGE_BYTES;
                c2x2 x[S];
#pragma unroll
                for (int s = 0; s < S; s++) x[s] = lds_c2x2_a(sp + s * TILE_BYTES);
                c2x2 h;
                h.a = h.b = 0ull;
                if (d < P) h = lds_c2x2_a(sp + S * TILE_BYTES);
                hw[(T - 1 + r) % T] = h;  // H[d]: the newest output (t = T-1) starts on partition j = d
#pragma unroll
                for (int t = (HEAD ? T - 1 - r : 0); t < T; t++) {
                    const c2x2 hh = hw[(t + r) % T];
#pragma unroll
                    for (int s = 0; s < S; s++) {
                        acc[t][s][0] = c2_cmac(acc[t][s][0], x[s].a, hh.a);
                        acc[t][s][1] = c2_cmac(acc[t][s][1], x[s].b, hh.b);
                    }
                }
                mbar_arrive_elect_a(bar0 + 8 * (NS + stage));
                phases ^= 1u << stage;
            }
        };
        using IT = std::integral_constant<int, T>;
        int d0 = 0;
        chunk(d0, std::true_type{}, IT{});   // D = P+T-1 >= T: all T rows of the head exist
#pragma unroll 1
        for (d0 = T; d0 + T <= D; d0 += T) chunk(d0, std::false_type{}, IT{});
        // the last D % T rows, as one of T-1 straight-line bodies (one per launch: P is the filter's)
        switch (D - d0) {
#define FCV_TAIL(n) case n: if constexpr (n < T) chunk(d0, std::false_type{}, std::integral_constant<int, n>{}); break;
            FCV_TAIL(1) FCV_TAIL(2) FCV_TAIL(3) FCV_TAIL(4) FCV_TAIL(5) FCV_TAIL(6) FCV_TAIL(7)
#undef FCV_TAIL
            default: break;
        }
        static_assert(T <= 8, "tail bodies");
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
