// fcv_k_mac_tma.cu -- launch of the TMA-staged time-tiled complex multiply-accumulate
// (fcv_mac_tma.cuh): T = 4 and T = 8 blocks per step, the batch path's roofline kernel.
#include "fcv_internal.h"
#include "fcv_mac_tma.cuh"

using namespace fcv;

// Returns false when the shape is not covered.
template <int T, int S, int NS, int G = NS, int MC = tma::min_ctas(T, S)>
static bool launch_tma(const StepArgs &a, int newest, cudaStream_t q) {
    const fcv_filter *f = a.f;
    const int M4 = f->fragm / 2;
    if (M4 % tma::TPB != 0 || a.cnt < S) return false;
    const size_t smem = tma::smem_bytes(S, NS);
    // dynamic + static shared memory exceeds the 48 KB default; the attribute is per device
    if (cudaFuncSetAttribute(tma::mac_tma_kernel<T, S, NS, MC, G>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
        return false;
    // Grid: persistent, MC CTAs per SM walking the items, for filters with one or two outputs; one CTA per work
    // item otherwise.  Measured with the final kernels (profiles/r02_experiments.md): SantaLucia 0.822 -> 0.811 ms,
    // crossfeed (ring depth 1: a work item is only 8 rows, the pipeline's cold start per CTA is what the persistent
    // grid saves) 0.476 -> 0.414 ms, dense 6 x 6 1.40 -> 1.53 ms (the CTAs drift apart and the outputs of a tile no
    // longer read their shared X rows together).  FCV_MAC_PERSIST=n forces n CTAs per SM (0: never persistent).
    static const int persist_env = getenv("FCV_MAC_PERSIST") ? atoi(getenv("FCV_MAC_PERSIST")) : -1;
    const int persist = persist_env >= 0 ? persist_env : (T == 8 && f->nout <= 2 ? MC : 0);
    const int ntiles = M4 / tma::TPB, ngroups = (a.cnt + S - 1) / S, nitems = ntiles * ngroups * f->nout;
    int grid = nitems;
    if (persist > 0 && a.num_sms * persist < nitems) grid = a.num_sms * persist;
    const float4 *H = reinterpret_cast<const float4 *>(f->dH);
    float4 *Y = reinterpret_cast<float4 *>(a.Y);
    tma::mac_tma_kernel<T, S, NS, MC, G><<<grid, tma::THREADS, smem, q>>>(a.bsel.xring0, a.bsel.xring_stride, a.cnt, f->dpairs, f->dpair_off,
                                                                     f->dtt_rows, H, Y, M4, f->ring, a.R, newest,
                                                                     f->nout, f->nrows, ntiles, ngroups, nitems);
    return true;
}

bool fcv::launch_mac_tma(const StepArgs &a, int newest, cudaStream_t q) {
    // T = 4 and T = 8 stream their rows through TMA-staged tiles whenever the shape allows
    // (spectrum tiles of 2 KB, at least two streams); measured equal to the register-pipelined
    // kernel at T = 8 and 2 % faster at T = 4.  FCV_MAC_TMA=0 turns it off, 2 / 3 select other
    // stream tilings (experiments).
    static const int use_tma = getenv("FCV_MAC_TMA") ? atoi(getenv("FCV_MAC_TMA")) : 1;
    if (!use_tma) return false;
    if (a.T == 8 && use_tma == 3 && launch_tma<8, 1, 12>(a, newest, q)) return true;
    // FCV_MAC_G: rows per producer pass (fcv_mac_tma.cuh; 0 = in order, one lane at a time)
    static const int g = getenv("FCV_MAC_G") ? atoi(getenv("FCV_MAC_G")) : 4;
    if (a.T == 8 && g == 0 && launch_tma<8, 2, 8, 0>(a, newest, q)) return true;
    if (a.T == 8 && g == 1 && launch_tma<8, 2, 8, 1>(a, newest, q)) return true;
    if (a.T == 8 && g == 2 && launch_tma<8, 2, 8, 2>(a, newest, q)) return true;
    if (a.T == 8 && g == 4 && launch_tma<8, 2, 8, 4>(a, newest, q)) return true;
    if (a.T == 8 && g == 8 && launch_tma<8, 2, 8, 8>(a, newest, q)) return true;
    if (a.T == 8 && launch_tma<8, 2, 8, 4>(a, newest, q)) return true;
    if (a.T == 4 && use_tma == 2 && launch_tma<4, 4, 6>(a, newest, q)) return true;
    if (a.T == 4 && launch_tma<4, 2, 8>(a, newest, q)) return true;
    return false;
}
