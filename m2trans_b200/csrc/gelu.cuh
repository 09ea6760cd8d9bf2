// GELU for the tail epilogues (ref M2Trans_network.py:44,:50: nn.GELU(), the exact erf form).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace m2t {

__device__ __forceinline__ float gelu_fast(float x) {
    const float z = fabsf(x) * 0.70710678118654752f;
    float t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float erf_abs = fmaf(-p * t, e, 1.f);          // erf(|x|/sqrt 2)
    const float hx = 0.5f * x;
    return fmaf(fabsf(hx), erf_abs, hx);                 // 0.5 x (1 + sign(x) erf_abs) = hx + |hx| erf_abs
}

// Two GELUs at a time on the packed fp32x2 pipe (FFMA2/FMUL2/FADD2 on sm_100).  The tail epilogues are bound by the
// FMA pipe (a packed instruction occupies it for two cycles, so packing saves issue slots, not pipe time) and then by
// the MUFU unit, so the form below minimises FMA-pipe operations and uses ONE MUFU per element:
//     gelu(x) = relu(x) - |x| * 2^(q(|x|) - 1),      2^q(a) ~= erfc(a / sqrt 2),  q(a) = a (c1 + c2 a + ... + c5 a^4)
// (q(0) = 0 exactly; the coefficients are a weighted minimax fit of the GELU error itself, max |error| 7.1e-7 over all of
// fp32 in fp32 arithmetic including 2 ulp of ex2.approx -- three orders below the fp16 rounding of the stored activation;
// q -> -inf for large |x|, so the correction underflows to zero and gelu = relu exactly there).  relu runs on the ALU pipe;
// 7 FMA-pipe operations per pair including the bias add.  Round 1 used erfc ~= (1 + p5(|x|))^-16 with a reciprocal and four
// squarings: 12 FMA-pipe operations per pair and 2.1e-6 of error.
__device__ __forceinline__ uint64_t f2_pack(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& a, float& b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t f2_splat(float a) { return f2_pack(a, a); }

// (gelu(x.lo), gelu(x.hi)) of a packed pair, fp32 results
__device__ __forceinline__ uint64_t gelu_pair(uint64_t x) {
    float x0, x1;
    f2_unpack(x, x0, x1);
    const uint64_t a = f2_pack(fabsf(x0), fabsf(x1));
    uint64_t p = f2_fma(f2_splat(-0.00048811757005751133f), a, f2_splat(0.007198806386440992f));
    p = f2_fma(p, a, f2_splat(-0.052146803587675095f));
    p = f2_fma(p, a, f2_splat(-0.4595957100391388f));
    p = f2_fma(p, a, f2_splat(-1.1510006189346313f));
    const uint64_t q = f2_fma(p, a, f2_splat(-1.f));          // log2(erfc(|x|/sqrt 2) / 2)
    float q0, q1, e0, e1;
    f2_unpack(q, q0, q1);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(q0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(q1));
    const uint64_t na = f2_pack(-fabsf(x0), -fabsf(x1));
    return f2_fma(na, f2_pack(e0, e1), f2_pack(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));
}

// returns half2 bits of (gelu(acc.x + bias.x), gelu(acc.y + bias.y))
__device__ __forceinline__ uint32_t gelu_pair_h2(uint64_t acc, uint64_t bias) {
    float g0, g1;
    f2_unpack(gelu_pair(f2_add(acc, bias)), g0, g1);
    const __half2 hv = __floats2half2_rn(g0, g1);
    return *reinterpret_cast<const uint32_t*>(&hv);
}

}  // namespace m2t
