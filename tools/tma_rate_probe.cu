// tma_rate_probe.cu -- is the time-tiled MAC bound by the number of bulk copies (cp.async.bulk)
// rather than by their bytes?  One producer lane streams a contiguous region through an
// NS-stage shared-memory ring in pieces of SZ bytes; 4 consumer warps read one float4 per
// thread and piece and release the stage.  3 CTAs per SM, 174 KB per CTA like mac_tma_kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I folve_b200/csrc -o tma_rate_probe tma_rate_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include "../folve_b200/csrc/fcv_mac_tma.cuh"
using namespace fcv::tma;

template <int SZ, int OPS>   // a stage = OPS copies of SZ bytes
__global__ void __launch_bounds__(160, 3) k(const unsigned char *__restrict__ x, size_t per_cta, float *sink) {
    constexpr int STAGE = SZ * OPS, NS = 48 * 1024 / STAGE;
    extern __shared__ __align__(128) unsigned char stages[];
    __shared__ uint64_t full[NS], empty[NS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < NS; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 4); }
        fence_barrier_init();
    }
    __syncthreads();
    const unsigned char *src = x + (size_t)blockIdx.x * per_cta;
    const int n = (int)(per_cta / STAGE);
    int stage = 0; uint32_t phase = 0;
    if (warp == 4) {
        if (lane == 0)
            for (int i = 0; i < n; i++) {
                mbar_wait(&empty[stage], phase ^ 1);
                mbar_expect_tx(&full[stage], STAGE);
                for (int o = 0; o < OPS; o++)
                    bulk_g2s(stages + stage * STAGE + o * SZ, src + (size_t)i * STAGE + o * SZ, SZ, &full[stage]);
                if (++stage == NS) { stage = 0; phase ^= 1; }
            }
        return;
    }
    float a = 0.f;
    for (int i = 0; i < n; i++) {
        mbar_wait(&full[stage], phase);
        for (int o = threadIdx.x * 16; o < STAGE; o += 128 * 16) {
            const fcv::c2x2 v = lds_c2x2(stages + stage * STAGE + o);
            a += __uint_as_float((unsigned)v.a) + __uint_as_float((unsigned)v.b);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == NS) { stage = 0; phase ^= 1; }
    }
    if (a == 12345.678f) *sink = a;
}

template <int SZ, int OPS>
void run(const unsigned char *x, float *sink, const char *what) {
    const int ctas = 32768 / 4;
    const size_t per_cta = 4 * 174 * 1024 / (SZ * OPS) * (SZ * OPS);
    cudaFuncSetAttribute(k<SZ, OPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        k<SZ, OPS><<<ctas, 160, 48 * 1024>>>(x, per_cta, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    printf("%-28s %.3f ms  %.0f GB/s  %.1f M copies/s per SM\n", what, ms, ctas * (double)per_cta / ms * 1e-6,
           ctas * (double)per_cta / SZ / ms * 1e-3 / 148);
}

int main() {
    const size_t bytes = (size_t)8192 * 4 * 174 * 1024;
    unsigned char *x; float *sink;
    cudaMalloc(&x, bytes + (1 << 20)); cudaMalloc(&sink, 4); cudaMemset(x, 0, bytes);
    run<1024, 1>(x, sink, "1 KB copies, 1 per stage");
    run<2048, 1>(x, sink, "2 KB copies, 1 per stage");
    run<2048, 3>(x, sink, "2 KB copies, 3 per stage");
    run<4096, 1>(x, sink, "4 KB copies, 1 per stage");
    run<6144, 1>(x, sink, "6 KB copies, 1 per stage");
    run<8192, 1>(x, sink, "8 KB copies, 1 per stage");
    run<16384, 1>(x, sink, "16 KB copies, 1 per stage");
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
