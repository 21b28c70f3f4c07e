// fcv_types.h -- plain structs shared by the host side and the kernels.
#pragma once
#include <cuda_runtime.h>

namespace fcv {

// PCM wire formats (== FCV_PCM_* of include/folve_b200.h)
enum { PCM_F32 = 0, PCM_S16 = 1, PCM_S24 = 2 };

// ---- tables (defined next to the kernels that read them) -----------------------------------
struct FftTables {
    const float2 *twA;     // [Q]   w_M^n = exp(-2 pi i n / M)
    const float2 *twU;     // [M]   exp(-i pi k / M) for the bin stored at entry e
    const float2 *twP[4];  // per pass t with stride S_t > 1: [(k1-1)*S_t + u] = exp(-2 pi i u k1 / Q_t)
    const unsigned short *part;  // [M] entry holding the conjugate-partner bin (M - k) of entry e
};
namespace f13 {
struct Tables {
    const float2 *twA0;  // [15][256]  w_M^(2 u k0),      k0 = 1..15   (half 0, pass A)
    const float2 *twA1;  // [16][256]  w_M^(u (2 k0 + 1)), k0 = 0..15   (half 1, pass A, premultiply folded in)
    const float2 *twB;   // [16][16]   w_256^(n0 k1)
    const float2 *twU;   // [2][256]   exp(-i pi (2c + h) / M): base of the unpack / repack twiddles (fcv_fft13.cuh)
};
}  // namespace f13

constexpr int MAC_NO_MAX = 8;
// One (input, partition) pair that feeds at least one output of the group.
struct MacStep {
    int inp;
    int part;
    int row[MAC_NO_MAX];  // filter row per output of the group, -1 = absent
};
struct TTPair {
    int inp;      // input channel feeding this output
    int rowbase;  // index into tt_rows: P consecutive filter-row numbers (absent -> the zero row)
};

}  // namespace fcv
