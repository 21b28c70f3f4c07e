// fcv_k_mac.cu -- launches of the complex multiply-accumulate kernels of fcv_mac.cuh: the
// block-by-block kernel (T = 1, DC / Nyquist products inside), the register-pipelined
// time-tiled kernel (T = 2, and the fallback for shapes the TMA-staged kernel does not cover)
// and the DC / Nyquist kernel of multi-block steps.  Role in the reference: the inner loops of
// zita-convolver's Convlevel::process(), reached from /root/reference/sound-processor.cc:113.
#include "fcv_internal.h"
#include "fcv_mac.cuh"

using namespace fcv;

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

void fcv::launch_dcny(const StepArgs &a, cudaStream_t q) {
    const fcv_filter *f = a.f;
    const int nwarps = a.cnt * f->nout * a.T;
    dcny_kernel<<<(nwarps + 7) / 8, 256, 0, q>>>(a.bsel.st, f->dpairs, f->dpair_off, f->dtt_rows, f->dH, a.zc0, nwarps,
                                                 f->nout, f->ring, a.R, a.T, a.bsel.pt, f->fragm);
}

template <class SEL, int NO, int S>
static void launch_mac1(const StepArgs &a, const SEL &sel, cudaStream_t q) {
    const fcv_filter *f = a.f;
    const int M4 = f->fragm / 2;
    const int TPB = M4 >= 128 ? 128 : M4;  // M4 is a power of two >= 32
    dim3 grid(M4 / TPB + 1, (a.cnt + S - 1) / S, f->ngroups);   // + 1: the DC / Nyquist column
    const float4 *H = reinterpret_cast<const float4 *>(f->dH);
#define FCV_MAC1_ARGS sel, a.cnt, f->dsteps, f->dgroup_off, H, M4, a.R, f->nout, f->dpairs, f->dpair_off, f->dtt_rows, f->ring
    if (TPB == 128) launch_k(mac_kernel<SEL, NO, S, 128>, grid, dim3(128), 0, q, a.pdl, FCV_MAC1_ARGS);
    else if (TPB == 64) launch_k(mac_kernel<SEL, NO, S, 64>, grid, dim3(64), 0, q, a.pdl, FCV_MAC1_ARGS);
    else launch_k(mac_kernel<SEL, NO, S, 32>, grid, dim3(32), 0, q, a.pdl, FCV_MAC1_ARGS);
#undef FCV_MAC1_ARGS
}

// The launch groups of the per-file path: one CTA per (tile, output, stream) -- mac_group_kernel.
template <class SEL>
static void launch_mac_group(const StepArgs &a, const SEL &sel, cudaStream_t q) {
    const fcv_filter *f = a.f;
    const int M4 = f->fragm / 2;
    const int TPB = M4 >= 128 ? 128 : M4;
    dim3 grid(M4 / TPB + 1, f->nout, a.cnt);
    const float4 *H = reinterpret_cast<const float4 *>(f->dH);
#define FCV_MACG_ARGS sel, H, M4, a.R, f->dpairs, f->dpair_off, f->dtt_rows, f->ring
    if (TPB == 128) launch_k(mac_group_kernel<SEL, 128>, grid, dim3(128), 0, q, a.pdl, FCV_MACG_ARGS);
    else if (TPB == 64) launch_k(mac_group_kernel<SEL, 64>, grid, dim3(64), 0, q, a.pdl, FCV_MACG_ARGS);
    else launch_k(mac_group_kernel<SEL, 32>, grid, dim3(32), 0, q, a.pdl, FCV_MACG_ARGS);
#undef FCV_MACG_ARGS
}

template <class SEL>
static void launch_mac_t1_sel(const StepArgs &a, const SEL &sel, cudaStream_t q) {
    // FCV_MAC_GROUP=0: the step-table kernel for launch groups too (A/B)
    static const bool group_kernel = !(getenv("FCV_MAC_GROUP") && atoi(getenv("FCV_MAC_GROUP")) == 0);
    if (SEL::kSingle && group_kernel && a.cnt <= 65535) {
        launch_mac_group<SEL>(a, sel, q);
        return;
    }
    const int S = a.cnt >= 4 ? 4 : (a.cnt >= 2 ? 2 : 1);
    switch (a.f->group_no) {
        case 1: if (S == 4) launch_mac1<SEL, 1, 4>(a, sel, q); else if (S == 2) launch_mac1<SEL, 1, 2>(a, sel, q); else launch_mac1<SEL, 1, 1>(a, sel, q); break;
        case 2: if (S == 4) launch_mac1<SEL, 2, 4>(a, sel, q); else if (S == 2) launch_mac1<SEL, 2, 2>(a, sel, q); else launch_mac1<SEL, 2, 1>(a, sel, q); break;
        case 4: if (S >= 2) launch_mac1<SEL, 4, 2>(a, sel, q); else launch_mac1<SEL, 4, 1>(a, sel, q); break;
        default: if (S >= 2) launch_mac1<SEL, 8, 2>(a, sel, q); else launch_mac1<SEL, 8, 1>(a, sel, q); break;
    }
}
void fcv::launch_mac_t1(const StepArgs &a, cudaStream_t q) {
    if (a.grp) launch_mac_t1_sel<GroupSel>(a, *a.grp, q);
    else launch_mac_t1_sel<BatchSel>(a, a.bsel, q);
}

// Time-tiled MAC: T blocks per stream per launch, one output channel per grid.z.
template <int T, int S>
static void launch_mac_tt_s(const StepArgs &a, int newest, cudaStream_t q) {
    const fcv_filter *f = a.f;
    const int M4 = f->fragm / 2;
    const int TPB = M4 >= 128 ? 128 : M4;
    dim3 grid(M4 / TPB, (a.cnt + S - 1) / S, f->nout);
    const float4 *H = reinterpret_cast<const float4 *>(f->dH);
    float4 *Y = reinterpret_cast<float4 *>(a.Y);
#define FCV_TT_ARGS a.bsel.st, a.cnt, f->dpairs, f->dpair_off, f->dtt_rows, H, Y, M4, f->ring, a.R, newest, f->nout
    if (TPB == 128) mac_tt_kernel<T, S, 128><<<grid, 128, 0, q>>>(FCV_TT_ARGS);
    else if (TPB == 64) mac_tt_kernel<T, S, 64><<<grid, 64, 0, q>>>(FCV_TT_ARGS);
    else mac_tt_kernel<T, S, 32><<<grid, 32, 0, q>>>(FCV_TT_ARGS);
#undef FCV_TT_ARGS
}

template <int T>
static void launch_mac_tt_t(const StepArgs &a, int newest, cudaStream_t q) {
    // Streams per thread, sharing each filter value from registers.  Measured on B200
    // (SantaLucia, 1024 streams): T=4: S=4 0.150 ms/block (HBM floor 0.148), S=2 0.166;
    // T=8: S=2 0.134, S=4 0.143 (228 registers).  FCV_TT_S overrides for experiments.
    static const int env_s = getenv("FCV_TT_S") ? atoi(getenv("FCV_TT_S")) : 0;
    int S = env_s ? env_s : (T >= 8 ? 2 : 4);
    if (a.cnt < S) S = a.cnt >= 2 ? 2 : 1;
    if (S >= 4) launch_mac_tt_s<T, 4>(a, newest, q);
    else if (S >= 2) launch_mac_tt_s<T, 2>(a, newest, q);
    else launch_mac_tt_s<T, 1>(a, newest, q);
}

void fcv::launch_mac_tt(const StepArgs &a, int newest, cudaStream_t q) {
    if (a.T == 2) launch_mac_tt_t<2>(a, newest, q);
    else if (a.T == 4) launch_mac_tt_t<4>(a, newest, q);
    else launch_mac_tt_t<8>(a, newest, q);
}
