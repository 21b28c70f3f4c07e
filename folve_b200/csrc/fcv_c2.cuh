// fcv_c2.cuh -- packed complex arithmetic on sm_100a.
//
// Blackwell executes two fp32 operations per lane in ONE issue slot with the packed
// instructions FFMA2 / FADD2 / FMUL2 (PTX fma.rn.f32x2, add.rn.f32x2, mul.rn.f32x2 on a
// 64-bit register pair).  A complex number (re, im) is exactly one such pair, and
// ptxas folds the operand shapes complex arithmetic needs into operand modifiers:
//   {a, a}     -> scalar broadcast operand      (R.F32)
//   {-b.y,b.x} -> swapped, half-negated operand (-R.F32x2.LO_HI.NP)
// so a complex multiply-accumulate is two FFMA2 and nothing else, and a complex
// add/sub with a multiplication by +-i folded in is one FADD2.  Measured on B200
// (tools/fp32x2_probe.cu): FFMA2 issues at half the rate of FFMA, i.e. the same flops
// per clock but HALF the issue slots -- the other half is free for the loads, address
// arithmetic and shared-memory traffic that the MAC and FFT kernels are otherwise
// issue-bound on.
//
// Every packed operation rounds exactly like the two scalar operations it stands for
// (same operands, same fused multiply-adds).
#pragma once
#include <cuda_runtime.h>

namespace fcv {

typedef unsigned long long c2;  // packed complex: low word = re, high word = im

__device__ __forceinline__ c2 c2_pack(float re, float im) {
    c2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(re), "f"(im));
    return r;
}
__device__ __forceinline__ c2 c2_pack(float2 v) { return c2_pack(v.x, v.y); }
__device__ __forceinline__ float2 c2_unpack(c2 v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ float c2_re(c2 v) { return c2_unpack(v).x; }
__device__ __forceinline__ float c2_im(c2 v) { return c2_unpack(v).y; }

__device__ __forceinline__ c2 c2_fma(c2 a, c2 b, c2 c) {  // (a.re*b.re + c.re, a.im*b.im + c.im)
    c2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ c2 c2_mul2(c2 a, c2 b) {  // element-wise product
    c2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ c2 c2_add(c2 a, c2 b) {
    c2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ c2 c2_neg(c2 a) {
    const float2 f = c2_unpack(a);
    return c2_pack(-f.x, -f.y);
}
__device__ __forceinline__ c2 c2_sub(c2 a, c2 b) { return c2_add(a, c2_neg(b)); }
__device__ __forceinline__ c2 c2_conj(c2 a) {
    const float2 f = c2_unpack(a);
    return c2_pack(f.x, -f.y);
}
// a * (+i) = (-im, re);  a * (-i) = (im, -re)
__device__ __forceinline__ c2 c2_mul_pi(c2 a) {
    const float2 f = c2_unpack(a);
    return c2_pack(-f.y, f.x);
}
__device__ __forceinline__ c2 c2_mul_ni(c2 a) {
    const float2 f = c2_unpack(a);
    return c2_pack(f.y, -f.x);
}
__device__ __forceinline__ c2 c2_scale(c2 a, float s) { return c2_mul2(a, c2_pack(s, s)); }

// acc += x * h (complex):
//   acc.re = fma(x.re, h.re, acc.re); acc.re = fma(-x.im, h.im, acc.re);
//   acc.im = fma(x.im, h.re, acc.im); acc.im = fma( x.re, h.im, acc.im);
// x is the operand that gets swapped/negated and h the one that is broadcast: with h
// shared by several x (streams) and x by several h (outputs) in the MAC kernels this
// is the form ptxas folds completely into operand modifiers -- 2 FFMA2, no moves.
__device__ __forceinline__ c2 c2_cmac(c2 acc, c2 x, c2 h) {
    const float2 xf = c2_unpack(x), hf = c2_unpack(h);
    acc = c2_fma(x, c2_pack(hf.x, hf.x), acc);
    acc = c2_fma(c2_pack(-xf.y, xf.x), c2_pack(hf.y, hf.y), acc);
    return acc;
}
// a * b (complex): (a.re*b.re - a.im*b.im, a.re*b.im + a.im*b.re) = FMUL2 + FFMA2.
// Operand order matters to ptxas: the swapped/half-negated pair must be the FIRST
// multiplicand and the broadcast scalar the second for both to fold into modifiers.
__device__ __forceinline__ c2 c2_cmul(c2 a, c2 b) {
    const float2 af = c2_unpack(a), bf = c2_unpack(b);
    const c2 t = c2_mul2(b, c2_pack(af.x, af.x));                        // (a.re*b.re, a.re*b.im)
    return c2_fma(c2_pack(-bf.y, bf.x), c2_pack(af.y, af.y), t);
}
// a * conj(b): (a.re*b.re + a.im*b.im, a.im*b.re - a.re*b.im) = b.re*(a.re, a.im) + b.im*(a.im, -a.re)
__device__ __forceinline__ c2 c2_cmulconj(c2 a, c2 b) {
    const float2 af = c2_unpack(a), bf = c2_unpack(b);
    const c2 t = c2_mul2(a, c2_pack(bf.x, bf.x));
    return c2_fma(c2_pack(af.y, -af.x), c2_pack(bf.y, bf.y), t);
}

struct c2x2 {  // two packed complex values = one 16-byte vector
    c2 a, b;
};
__device__ __forceinline__ c2x2 c2x2_from(float4 v) {
    c2x2 r;
    r.a = c2_pack(v.x, v.y);
    r.b = c2_pack(v.z, v.w);
    return r;
}
__device__ __forceinline__ float4 c2x2_to(c2x2 v) {
    const float2 a = c2_unpack(v.a), b = c2_unpack(v.b);
    return make_float4(a.x, a.y, b.x, b.y);
}

}  // namespace fcv
