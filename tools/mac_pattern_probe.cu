// mac_pattern_probe.cu -- does the layout of the input-spectra ring matter to the time-tiled MAC?
// Reproduces only the memory traffic of mac_tma_kernel<8,2,*> (SantaLucia, 1024 streams): CTA
// (tile, stream pair, output) reads D = 29 pieces of 2 KB per stream and writes 8 pieces per stream.
//   layout 0 (current): ring[stream][input][slot][tile]  -> pieces 64 KB apart, neighbours adjacent
//   layout 1 (tile major): ring[stream][input][tile][slot] -> one contiguous 58 KB run per CTA and stream
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mac_pattern_probe mac_pattern_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
constexpr int D = 29, T = 8, NT = 32, TILE16 = 128;   // 2 KB tiles of 16-byte words

template <int LAYOUT>
__global__ void __launch_bounds__(128, 4)
k(const float4 *__restrict__ x, float4 *__restrict__ y, int nstreams, float *sink) {
    const int tile = blockIdx.x, b0 = blockIdx.y * 2, o = blockIdx.z;
    float4 a = make_float4(0, 0, 0, 0);
    for (int s = 0; s < 2; s++) {
        const size_t ring = ((size_t)(b0 + s) * 2 + o) * D * NT * TILE16;
        const float4 *p = x + ring + (LAYOUT == 0 ? (size_t)tile * TILE16 : (size_t)tile * D * TILE16) + threadIdx.x;
        const size_t step = LAYOUT == 0 ? (size_t)NT * TILE16 : TILE16;
        float4 v[D];
#pragma unroll
        for (int d = 0; d < D; d++) v[d] = __ldcs(p + d * step);
#pragma unroll
        for (int d = 0; d < D; d++) { a.x += v[d].x; a.y += v[d].y; a.z += v[d].z; a.w += v[d].w; }
    }
    for (int s = 0; s < 2; s++)
#pragma unroll
        for (int t = 0; t < T; t++)
            __stcs(y + ((((size_t)(b0 + s) * 2 + o) * T + t) * NT + tile) * TILE16 + threadIdx.x, a);
    if (a.x == 12345.678f) *sink = a.y;
}

int main() {
    const int nstreams = 1024;
    const size_t xb = (size_t)nstreams * 2 * D * NT * TILE16 * 16, yb = (size_t)nstreams * 2 * T * NT * TILE16 * 16;
    float4 *x, *y; float *sink;
    cudaMalloc(&x, xb); cudaMalloc(&y, yb); cudaMalloc(&sink, 4);
    cudaMemset(x, 0, xb);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const dim3 grid(NT, nstreams / 2, 2);
    for (int layout = 0; layout < 2; layout++)
        for (int r = 0; r < 4; r++) {
            cudaEventRecord(e0);
            if (layout == 0) k<0><<<grid, 128>>>(x, y, nstreams, sink); else k<1><<<grid, 128>>>(x, y, nstreams, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (r == 3) printf("layout %d: %.3f ms, %.0f GB/s (%.2f GB read, %.2f GB written)\n", layout, ms, (xb + yb) / ms * 1e-6, xb * 1e-9, yb * 1e-9);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
