// hostlink_probe.cu -- how fast can G CTAs of 256 threads write / read one 64 KB block in mapped pinned host
// memory (what host_copy_out and the forward kernel's zero-copy read do), as a function of G?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/hostlink_probe tools/hostlink_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void wr(const int4 *__restrict__ src, int4 *dst, int n16, volatile unsigned *flag, unsigned *cnt, unsigned seq) {
    const int per = (n16 + gridDim.x - 1) / gridDim.x, lo = blockIdx.x * per, hi = min(n16, lo + per);
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) dst[i] = __ldcg(src + i);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(cnt, 1u) == gridDim.x - 1) { *cnt = 0; __threadfence_system(); *flag = seq; }
    }
}
// one system fence only: by the thread that publishes the flag, after the CTA barrier (fence cumulativity)
__global__ void wr1(const int4 *__restrict__ src, int4 *dst, int n16, volatile unsigned *flag, unsigned *cnt, unsigned seq) {
    const int per = (n16 + gridDim.x - 1) / gridDim.x, lo = blockIdx.x * per, hi = min(n16, lo + per);
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) dst[i] = __ldcg(src + i);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(cnt, 1u) == gridDim.x - 1) { *cnt = 0; __threadfence_system(); *flag = seq; }
    }
}
// no fence at all (not a valid protocol: shows what the fences cost)
__global__ void wr0(const int4 *__restrict__ src, int4 *dst, int n16, volatile unsigned *flag, unsigned *cnt, unsigned seq) {
    const int per = (n16 + gridDim.x - 1) / gridDim.x, lo = blockIdx.x * per, hi = min(n16, lo + per);
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) dst[i] = __ldcg(src + i);
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(cnt, 1u) == gridDim.x - 1) { *cnt = 0; *flag = seq; }
}
__global__ void rd(const int4 *__restrict__ src, int4 *dst, int n16) {
    const int per = (n16 + gridDim.x - 1) / gridDim.x, lo = blockIdx.x * per, hi = min(n16, lo + per);
    int4 v[8];
    for (int i0 = lo + threadIdx.x; i0 < hi; i0 += blockDim.x * 8) {
#pragma unroll
        for (int k = 0; k < 8; k++) if (i0 + k * blockDim.x < hi) v[k] = __ldcs(src + i0 + k * blockDim.x);
#pragma unroll
        for (int k = 0; k < 8; k++) if (i0 + k * blockDim.x < hi) dst[i0 + k * blockDim.x] = v[k];
    }
}
int main() {
    const int bytes = 65536, n16 = bytes / 16;
    int4 *h, *hd, *d; unsigned *cnt;
    cudaHostAlloc(&h, bytes + 256, cudaHostAllocMapped); cudaHostGetDevicePointer(&hd, h, 0);
    cudaMalloc(&d, bytes); cudaMalloc(&cnt, 4); cudaMemset(cnt, 0, 4); cudaMemset(d, 1, bytes);
    volatile unsigned *flag = (volatile unsigned *)((char *)hd + bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int G : {1, 2, 8}) {
        for (int mode = 0; mode < 6; mode++) {
            float best = 1e9, sum = 0; const int reps = 200;
            for (int r = 0; r < reps + 20; r++) {
                cudaEventRecord(e0);
                if (mode == 0) wr<<<G, 256>>>(d, hd, n16, flag, cnt, r + 1);
                else if (mode == 1) rd<<<G, 256>>>(hd, d, n16);
                else if (mode == 2) wr1<<<G, 256>>>(d, hd, n16, flag, cnt, r + 1);
                else if (mode == 3) wr0<<<G, 256>>>(d, hd, n16, flag, cnt, r + 1);
                else if (mode == 4) wr1<<<G, 256>>>(d, hd, n16 / 2, flag, cnt, r + 1);
                else rd<<<G, 256>>>(hd, d, n16 / 2);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (r >= 20) { sum += ms; best = ms < best ? ms : best; }
            }
            const char *names[6] = {"write 64 KB, fence by all + fence", "read  64 KB", "write 64 KB, one fence", "write 64 KB, no fence", "write 32 KB, one fence", "read  32 KB"};
            printf("%-36s %2d CTAs: mean %.2f us, best %.2f us\n", names[mode], G, 1e3 * sum / reps, 1e3 * best);
        }
    }
    // empty-kernel floor
    { float sum = 0; for (int r = 0; r < 220; r++) { cudaEventRecord(e0); rd<<<1, 256>>>(hd, d, 0); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r >= 20) sum += ms; }
      printf("empty kernel: mean %.2f us\n", 1e3 * sum / 200); }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
