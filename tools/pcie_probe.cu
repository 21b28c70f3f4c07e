// tools/pcie_probe.cu -- measures the host<->device copy ceilings that bound bench.py's
// e2e number: H2D alone, D2H alone, both directions at once (pinned memory, 64 MiB).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
int main() {
    const size_t n = 64u << 20;
    void *h0, *h1, *d0, *d1;
    cudaHostAlloc(&h0, n, cudaHostAllocDefault);
    cudaHostAlloc(&h1, n, cudaHostAllocDefault);
    memset(h0, 1, n); memset(h1, 2, n);
    cudaMalloc(&d0, n); cudaMalloc(&d1, n);
    cudaStream_t s0, s1;
    cudaStreamCreate(&s0); cudaStreamCreate(&s1);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float ms;
    for (int mode = 0; mode < 3; mode++) {
        for (int rep = 0; rep < 3; rep++) {
            cudaDeviceSynchronize();
            cudaEventRecord(a, s0);
            const int iters = 10;
            for (int i = 0; i < iters; i++) {
                if (mode == 0 || mode == 2) cudaMemcpyAsync(d0, h0, n, cudaMemcpyHostToDevice, s0);
                if (mode == 1 || mode == 2) cudaMemcpyAsync(h1, d1, n, cudaMemcpyDeviceToHost, mode == 2 ? s1 : s0);
            }
            cudaStreamSynchronize(s1);
            cudaEventRecord(b, s0);
            cudaEventSynchronize(b);
            cudaEventElapsedTime(&ms, a, b);
            if (rep == 2)
                printf("%s: %.2f ms per 64 MiB%s -> %.1f GB/s per direction\n",
                       mode == 0 ? "H2D" : mode == 1 ? "D2H" : "H2D+D2H concurrent", ms / iters,
                       mode == 2 ? " each way" : "", n / (ms / iters * 1e-3) / 1e9);
        }
    }
    // chunked: 8 x 8 MiB like fcv_batch_process
    for (int rep = 0; rep < 3; rep++) {
        cudaDeviceSynchronize();
        cudaEventRecord(a, s0);
        for (int c = 0; c < 8; c++) {
            cudaMemcpyAsync((char *)d0 + c * (n / 8), (char *)h0 + c * (n / 8), n / 8, cudaMemcpyHostToDevice, s0);
            cudaMemcpyAsync((char *)h1 + c * (n / 8), (char *)d1 + c * (n / 8), n / 8, cudaMemcpyDeviceToHost, s1);
        }
        cudaStreamSynchronize(s1);
        cudaEventRecord(b, s0);
        cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
        if (rep == 2) printf("8 x 8 MiB both ways: %.2f ms\n", ms);
    }
    return 0;
}
