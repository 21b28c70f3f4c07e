// tools/pcie_probe.cu -- measures the host<->device copy ceilings that bound bench.py's
// e2e number: H2D alone, D2H alone, both directions at once (pinned memory, 64 MiB).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <time.h>
// usage: pcie_probe [iters [sync]] -- with "sync" the measurement starts on the next multiple of 2 s of the
// wall clock, so that probes started together on several GPUs really overlap
int main(int argc, char **argv) {
    const int iters_arg = argc > 1 ? atoi(argv[1]) : 10;
    const bool wall_sync = argc > 2;
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
            if (wall_sync && rep == 2) {
                // all probes started within the same 8 s window agree on `base`; mode m starts at base + 2 m
                static time_t base = 0;
                timespec ts;
                clock_gettime(CLOCK_REALTIME, &ts);
                if (!base) base = ((ts.tv_sec + 1) / 8 + 1) * 8;
                do clock_gettime(CLOCK_REALTIME, &ts); while (ts.tv_sec < base + 2 * mode);
            }
            cudaEventRecord(a, s0);
            const int iters = rep == 2 ? iters_arg : 3;
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
