// fp32x2_probe.cu -- issue-rate probe for sm_100a packed fp32 (FFMA2/FADD2) against
// scalar FFMA/FADD: warp-instructions per clock per SM with 8 independent chains.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32x2_probe fp32x2_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0,{%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float2 up(u64 v) { float2 r; asm("mov.b64 {%0,%1},%2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }

template <int MODE>
__global__ void probe(float2 *out, const float2 *ab, int iters) {
    float2 a = ab[threadIdx.x & 1], b = ab[2 + (threadIdx.x & 1)];  // per-thread registers, not uniform
    float2 c[8];
#pragma unroll
    for (int i = 0; i < 8; i++) c[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    if (MODE == 0) {  // scalar FFMA: 16 per iteration
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 8; i++) { c[i].x = fmaf(c[i].x, a.x, b.x); c[i].y = fmaf(c[i].y, a.y, b.y); }
        }
    } else if (MODE == 1) {  // FFMA2: 8 per iteration (same flops as MODE 0)
        u64 C[8], A = pk(a.x, a.y), B = pk(b.x, b.y);
#pragma unroll
        for (int i = 0; i < 8; i++) C[i] = pk(c[i].x, c[i].y);
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32x2 %0,%0,%1,%2;" : "+l"(C[i]) : "l"(A), "l"(B));
        }
#pragma unroll
        for (int i = 0; i < 8; i++) c[i] = up(C[i]);
    } else if (MODE == 2) {  // complex MAC via 2 FFMA2 (broadcast + swizzle operands), 8 cMACs per iteration
        u64 C[8];
#pragma unroll
        for (int i = 0; i < 8; i++) C[i] = pk(c[i].x, c[i].y);
        u64 X = pk(a.x, a.y), Xs = pk(-a.y, a.x), Hr = pk(b.x, b.x), Hi = pk(b.y, b.y);
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                asm volatile("fma.rn.f32x2 %0,%1,%2,%0;" : "+l"(C[i]) : "l"(X), "l"(Hr));
                asm volatile("fma.rn.f32x2 %0,%1,%2,%0;" : "+l"(C[i]) : "l"(Xs), "l"(Hi));
            }
        }
#pragma unroll
        for (int i = 0; i < 8; i++) c[i] = up(C[i]);
    } else if (MODE == 3) {  // scalar FADD: 16 per iteration
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 8; i++) { c[i].x += a.x; c[i].y += a.y; }
        }
    } else if (MODE == 4) {  // FADD2: 8 per iteration
        u64 C[8], A = pk(a.x, a.y);
#pragma unroll
        for (int i = 0; i < 8; i++) C[i] = pk(c[i].x, c[i].y);
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("add.rn.f32x2 %0,%0,%1;" : "+l"(C[i]) : "l"(A));
        }
#pragma unroll
        for (int i = 0; i < 8; i++) c[i] = up(C[i]);
    } else if (MODE == 5) {  // scalar complex MAC: 4 FFMA per cMAC, 8 cMACs per iteration, register operands
        float2 x = a, h = b;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                c[i].x = fmaf(x.x, h.x, c[i].x); c[i].x = fmaf(-x.y, h.y, c[i].x);
                c[i].y = fmaf(x.x, h.y, c[i].y); c[i].y = fmaf(x.y, h.x, c[i].y);
            }
            x.x += 1e-9f;  // keep operands in registers (not uniform/constant)
        }
    }
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; i++) { s.x += c[i].x; s.y += c[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, int per_iter, int flops_per_instr) {
    int dev; cudaGetDevice(&dev);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev);
    const int threads = 512, blocks = p.multiProcessorCount * 2, iters = 1 << 15;
    float2 *out; cudaMalloc(&out, sizeof(float2) * threads * blocks);
    float2 hab[4] = {{0.999f, 1.001f}, {0.999f, 1.001f}, {1e-3f, -1e-3f}, {1e-3f, -1e-3f}};
    float2 *ab; cudaMalloc(&ab, sizeof(hab)); cudaMemcpy(ab, hab, sizeof(hab), cudaMemcpyHostToDevice);
    probe<MODE><<<blocks, threads>>>(out, ab, 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0); probe<MODE><<<blocks, threads>>>(out, ab, iters); cudaEventRecord(e1);
        cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double winstr = (double)blocks * (threads / 32) * (double)iters * per_iter;
    double per_s = winstr / (best * 1e-3);
    printf("%-28s %8.3f ms  %7.1f G warp-instr/s  = %.3f /clk/SM at %d MHz (max clock)  %.1f TFLOP/s\n", name, best,
           per_s * 1e-9, per_s / p.multiProcessorCount / (clk_khz * 1e3), clk_khz / 1000,
           per_s * 32 * flops_per_instr * 1e-12);
    cudaFree(out);
}

int main() {
    run<0>("FFMA (scalar, reg ops)", 16, 2);
    run<1>("FFMA2 (packed)", 8, 4);
    run<2>("cMAC = 2 x FFMA2 swz/bcast", 16, 4);
    run<5>("cMAC = 4 x FFMA", 32, 2);
    run<3>("FADD (scalar)", 16, 1);
    run<4>("FADD2 (packed)", 8, 2);
    return 0;
}
