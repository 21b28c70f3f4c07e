// fcv_k_fused13.cu -- ONE launch per group of single-stream blocks (the per-file path of folve:
// SoundProcessor::Process(), /root/reference/sound-processor.cc:98-127, for up to 32 files at once).
//
// The three kernels of a T = 1 step -- forward transform, complex multiply-accumulate over the
// partition history, inverse transform with overlap-add and the copy-out to the caller's pinned
// block -- as the three phases of one cooperative kernel separated by grid-wide barriers.  Nothing
// is computed differently: the phases call the very device functions of fcv_fft13.cuh and
// fcv_mac.cuh that the separate kernels call, so results are bit-identical to the three-launch
// path (tests/test_coalesce_gpu.py).  What changes is the cost of getting a group onto the GPU:
// one launch instead of three and no launch gaps between the phases.
//
// MEASURED (profiles/r02_experiments.md): host time per group 27 -> 13 us when many threads launch
// (6 us from the single dispatcher thread), but a lone block takes 58 instead of 51 us (two
// grid-wide barriers, 15 of 17 CTAs idle in the first and last phase), and under load concurrent
// cooperative grids overlap worse than chained ordinary launches: 16 callers reach 20-22 k x realtime
// against 24 k.  It is therefore OFF by default (FCV_FUSED=1 turns it on); kept because the
// equivalence is tested and the measurement answers the question what one launch per block buys.
//
// Covered shape: fragm = 8192, stereo in and out (what folve feeds it); everything else keeps
// the three-launch path.
//
//   grid : cooperative, min(work items of the widest phase, what is co-resident) CTAs x 256 threads
//   phase 1: item = (stream, half of the spectrum): fwd_half of both channels        2 n items
//   phase 2: item = (stream, 4 KB tile of the spectrum row): both outputs' MAC       16 n items
//            + one warp per (stream, output): DC / Nyquist products                  ceil(2 n / 8) items
//   phase 3: item = (stream, output): inverse transform, overlap-add, PCM out,
//            last CTA of a stream copies block + maximum to the pinned host block    2 n items
#include <cooperative_groups.h>

#include <map>

#include "fcv_internal.h"
#include "fcv_fft13.cuh"
#include "fcv_mac.cuh"

using namespace fcv;
namespace cg = cooperative_groups;

struct FusedFilter {
    const MacStep *steps;
    const int *group_off;
    const float4 *H;
    const TTPair *pairs;
    const int *pair_off;
    const int *tt_rows;
    int R;       // ring depth of the streams (T = 1: the filter's)
    int Pfilt;   // partitions with data
};

constexpr int FUSED_NT = 256;
constexpr int FUSED_TILE = FUSED_NT;                   // float4 columns per MAC item
constexpr int FUSED_TILES = (f13::N / 2) / FUSED_TILE; // 16 tiles per spectrum row

template <int FIN, int FOUT>
__global__ void __launch_bounds__(FUSED_NT, 2)
fused13_stereo_kernel(const __grid_constant__ GroupSel sel, int n, f13::Tables tb, FusedFilter ff) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c2 *sm = reinterpret_cast<c2 *>(smem_raw);
    __shared__ float red[FUSED_NT / 32];
    cg::grid_group grid = cg::this_grid();
    constexpr int N = f13::N, M = N, M4 = N / 2;
    const int tid = threadIdx.x;

    // ---- phase 1: forward transforms into the ring slot of this block
    for (int item = blockIdx.x; item < 2 * n; item += gridDim.x) {
        const int b = item >> 1, h = item & 1;
        const StreamDev s = sel.stream(b);
        int frames = sel.frames(b);
        frames = frames < 0 ? 0 : (frames > N ? N : frames);
        const int slot = sel.slot(b);
        float2 *rows[2] = {s.xring + (size_t)(0 * ff.R + slot) * N, s.xring + (size_t)(1 * ff.R + slot) * N};
        if (h == 0 && tid == 0) {   // the block's maximum starts from zero
            *s.maxv = 0.0f;
            s.bmax[0] = 0.0f;
        }
        if (frames == 0) {          // silence: its spectrum is zero
            for (int c = 0; c < 2; c++)
                for (int e = tid; e < f13::Q; e += FUSED_NT) rows[c][h * f13::Q + e] = make_float2(0.f, 0.f);
        } else if (h == 0) {
            f13::fwd_half<0, FIN, 2, 2, FUSED_NT>(sm, tb, s.din, 2, 0, frames, rows);
        } else {
            f13::fwd_half<1, FIN, 2, 2, FUSED_NT>(sm, tb, s.din, 2, 0, frames, rows);
        }
        __syncthreads();   // shared memory is reused by the next item
    }
    grid.sync();

    // ---- phase 2: Y[o] = sum over (input, partition) of X * H, and the two real bins
    {
        const int nmac = FUSED_TILES * n, ndc = (2 * n + 7) / 8;
        for (int item = blockIdx.x; item < nmac + ndc; item += gridDim.x) {
            if (item < nmac) {
                const int b = item / FUSED_TILES, e4 = (item % FUSED_TILES) * FUSED_TILE + tid;
                const StreamDev s = sel.stream(b);
                float4 acc[2][1];
                acc[0][0] = acc[1][0] = make_float4(0.f, 0.f, 0.f, 0.f);
                mac_single_steps<2, 1>(acc, reinterpret_cast<const float4 *>(s.xring) + e4, sel.slot(b), ff.steps,
                                       ff.group_off[0], ff.group_off[1], ff.H, M4, ff.R, e4);
                float4 *y = reinterpret_cast<float4 *>(s.Y) + e4;
                __stcs(y, acc[0][0]);
                __stcs(y + M4, acc[1][0]);
            } else {
                const int w = (item - nmac) * 8 + (tid >> 5);
                if (w < 2 * n) {
                    const int b = w >> 1, o = w & 1;
                    const StreamDev s = sel.stream(b);
                    const float2 z = dcny_warp(s.xring, ff.pairs, ff.pair_off, ff.tt_rows,
                                               reinterpret_cast<const float2 *>(ff.H), o, ff.Pfilt, ff.R, sel.slot(b), M,
                                               tid & 31);
                    if ((tid & 31) == 0) s.zc0[o] = z;
                }
            }
        }
    }
    grid.sync();

    // ---- phase 3: inverse transform, overlap-add, PCM out, maximum, copy-out to the host block
    for (int item = blockIdx.x; item < 2 * n; item += gridDim.x) {
        const int b = item >> 1, o = item & 1;
        const StreamDev s = sel.stream(b);
        int frames = sel.frames(b);
        frames = frames < 0 ? 0 : (frames > N ? N : frames);
        float2 *tail = reinterpret_cast<float2 *>(s.tail + (size_t)o * N);
        const float2 z0 = tid == 0 ? s.zc0[o] : make_float2(0.f, 0.f);
        const float2 *yrow = s.Y + (size_t)o * M;
        if (tid < 128) f13::inv_pass_c<0>(sm, tb, yrow, c2_pack(z0.x, z0.y), tid);
        else f13::inv_pass_c<1>(sm + f13::HALF_ELEMS, tb, yrow, 0ull, tid - 128);
        __syncthreads();
        f13::pass_b<+1, 2, FUSED_NT>(sm, tb);
        __syncthreads();
        float lmax = f13::inv_pass_a<FOUT, FUSED_NT>(sm, tb, tail, s.dout, 2, o, frames);
        {   // this block's maximum (bmax) and the stream's (maxv): both start from zero in phase 1
            float m = lmax;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
            if ((tid & 31) == 0) red[tid >> 5] = m;
            __syncthreads();
            if (tid == 0) {
                for (int w = 1; w < FUSED_NT / 32; w++) m = fmaxf(m, red[w]);
                if (m > 0.0f) {
                    atomicMax(reinterpret_cast<int *>(s.bmax), __float_as_int(m));
                    atomicMax(reinterpret_cast<int *>(s.maxv), __float_as_int(m));
                }
            }
        }
        if (s.hout) host_copy_out(s, 2, (size_t)frames * 2 * (FOUT == PCM_S16 ? 2 : 4), sel.seq(b), tid, FUSED_NT);
        __syncthreads();   // shared memory (and `red`) are reused by the next item
    }
}

namespace {
struct FusedCtx {
    bool checked = false;
    bool ok = false;
    int max_ctas = 0;
};
std::mutex g_mu;
std::map<int, FusedCtx> g_ctx;
}  // namespace

template <int FIN, int FOUT>
static bool prepare(int device, FusedCtx &c) {
    const size_t smem = 2 * f13::HALF_BYTES;
    if (cudaFuncSetAttribute(fused13_stereo_kernel<FIN, FOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess)
        return false;
    int per_sm = 0, sms = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fused13_stereo_kernel<FIN, FOUT>, FUSED_NT, smem) !=
            cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || per_sm < 1)
        return false;
    const int m = per_sm * sms;
    if (c.max_ctas == 0 || m < c.max_ctas) c.max_ctas = m;
    return true;
}

// Whether a group of this filter can take the one-launch path on `device` (decided once per device).
bool fcv::fused13_available(const fcv_filter *f, int in_fmt, int out_fmt) {
    if (!f->k13 || f->ninp != 2 || f->nout != 2 || f->group_no != 2 || f->ngroups != 1) return false;
    if (in_fmt != out_fmt) return false;   // instantiated for equal wire formats only
    std::lock_guard<std::mutex> l(g_mu);
    FusedCtx &c = g_ctx[f->device];
    if (!c.checked) {
        c.checked = true;
        int coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, f->device);
        c.ok = coop != 0 && prepare<PCM_F32, PCM_F32>(f->device, c) && prepare<PCM_S16, PCM_S16>(f->device, c) &&
               prepare<PCM_S24, PCM_S24>(f->device, c);
        if (!c.ok) cudaGetLastError();
    }
    return c.ok;
}

template <int FMT>
static cudaError_t launch_fmt(const StepArgs &a, int grid, cudaStream_t q) {
    const fcv_filter *f = a.f;
    GroupSel sel = *a.grp;
    int n = a.cnt;
    f13::Tables tb = f->tb13;
    FusedFilter ff;
    ff.steps = f->dsteps;
    ff.group_off = f->dgroup_off;
    ff.H = reinterpret_cast<const float4 *>(f->dH);
    ff.pairs = f->dpairs;
    ff.pair_off = f->dpair_off;
    ff.tt_rows = f->dtt_rows;
    ff.R = a.R;
    ff.Pfilt = f->ring;
    void *args[] = {&sel, &n, &tb, &ff};
    return cudaLaunchCooperativeKernel((const void *)fused13_stereo_kernel<FMT, FMT>, dim3(grid), dim3(FUSED_NT), args,
                                       2 * f13::HALF_BYTES, q);
}

// One cooperative launch for the whole group; false if it could not be launched (the caller then
// takes the three-launch path).
bool fcv::launch_fused13(const StepArgs &a, cudaStream_t q) {
    int max_ctas;
    {
        std::lock_guard<std::mutex> l(g_mu);
        max_ctas = g_ctx[a.f->device].max_ctas;
    }
    int items = FUSED_TILES * a.cnt + (2 * a.cnt + 7) / 8;
    int grid = items < max_ctas ? items : max_ctas;
    if (grid < 1) return false;
    cudaError_t e;
    if (a.in_fmt == PCM_F32) e = launch_fmt<PCM_F32>(a, grid, q);
    else if (a.in_fmt == PCM_S16) e = launch_fmt<PCM_S16>(a, grid, q);
    else e = launch_fmt<PCM_S24>(a, grid, q);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    g_launches += 1;
    return true;
}
