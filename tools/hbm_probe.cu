// hbm_probe.cu -- what the FFT kernels' memory pattern costs on B200, without the FFT:
//   A. streaming read / write / copy bandwidth (grid-stride, 16-byte accesses)
//   B. "row jobs": each CTA reads RIN KB, spins for SPIN cycles of register work, then
//      writes ROUT KB to a row that lies ~3.7 MB away from its neighbour's (one ring
//      slot per stream), CTAS_PER_SM CTAs resident.  Shows how much of a CTA's store
//      burst is exposed when only 2-3 CTAs share an SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hbm_probe hbm_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__global__ void k_read(const float4 *p, size_t n, float *sink) {
    float4 a = make_float4(0, 0, 0, 0);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = __ldcs(p + i);
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    if (a.x + a.y + a.z + a.w == 12345.678f) *sink = a.x;
}
__global__ void k_write(float4 *p, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}
__global__ void k_copy(const float4 *s, float4 *d, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        d[i] = __ldcs(s + i);
}

// one job per CTA; dynamic shared memory only to pin the CTAs-per-SM count
__global__ void __launch_bounds__(256)
k_rows(const float4 *in, float4 *out, int rin16, int rout16, int spin, size_t in_stride16, size_t out_stride16,
       int nstreams, float *sink) {
    extern __shared__ float4 sm[];
    const int job = blockIdx.x;
    const int stream = job % nstreams, blk = job / nstreams;
    const float4 *src = in + (size_t)job * in_stride16;
    float4 *dst = out + (size_t)stream * out_stride16 + (size_t)blk * rout16;
    float4 a = make_float4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < rin16; i += 256) {
        const float4 v = __ldg(src + i);
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    sm[threadIdx.x] = a;
    __syncthreads();
    float x = a.x + sm[(threadIdx.x + 1) & 255].y;
    for (int i = 0; i < spin; i++) x = fmaf(x, 1.0000001f, 1e-7f);
    for (int i = threadIdx.x; i < rout16; i += 256) dst[i] = make_float4(x, a.y, a.z, a.w);
    if (x == 12345.678f) *sink = x;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

int main(int argc, char **argv) {
    const size_t bytes = 4ull << 30, n = bytes / 16;
    float4 *a, *b; float *sink;
    cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&sink, 4);
    cudaMemset(a, 1, bytes); cudaMemset(b, 2, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 8;
    for (int mode = 0; mode < 3; mode++) {
        float best = 1e9f;
        for (int r = 0; r < 4; r++) {
            cudaEventRecord(e0);
            if (mode == 0) k_read<<<grid, 512>>>(a, n, sink);
            else if (mode == 1) k_write<<<grid, 512>>>(b, n);
            else k_copy<<<grid, 512>>>(a, b, n);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            const float ms = time_ms(e0, e1); if (ms < best) best = ms;
        }
        const double moved = mode == 2 ? 2.0 * bytes : bytes;
        printf("%-6s %8.3f ms  %7.1f GB/s\n", mode == 0 ? "read" : mode == 1 ? "write" : "copy", best, moved / best * 1e-6);
    }
    // row jobs: 8192 jobs, 1024 streams, in 32 KB (shared PCM block: 64 KB per two CTAs), out 64 KB
    const int jobs = 8192, nstreams = 1024;
    const int rin16 = 32 * 1024 / 16, rout16 = 64 * 1024 / 16;
    const size_t out_stride16 = (size_t)(3800 * 1024) / 16;   // ~3.7 MB between the rows of neighbouring streams
    for (int ctas = 2; ctas <= 6; ctas++) {
        const int smem = (227 * 1024 / ctas - 1024) & ~1023;
        cudaFuncSetAttribute(k_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        for (int spin = 0; spin <= 6000; spin += 1500) {
            float best = 1e9f;
            for (int r = 0; r < 3; r++) {
                cudaEventRecord(e0);
                k_rows<<<jobs, 256, smem>>>(a, b, rin16, rout16, spin, rin16, out_stride16, nstreams, sink);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                const float ms = time_ms(e0, e1); if (ms < best) best = ms;
            }
            printf("rows: %d CTAs/SM spin %5d  %7.3f ms  %6.1f ns/job  (%.0f GB/s)\n", ctas, spin, best,
                   best * 1e6 / jobs, (double)jobs * (rin16 + rout16) * 16 / best * 1e-6);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
