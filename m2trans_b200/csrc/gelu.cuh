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
//     gelu(x) = relu(x) - 0.5 |x| erfc(|x| / sqrt 2),      erfc(|x| / sqrt 2) ~= (1 + c1|x| + ... + c5|x|^5)^-16
// (the shape of Abramowitz-Stegun 7.1.28 with the 1/sqrt 2 folded in and the coefficients re-fitted for degree 5:
// max |gelu error| 2.1e-6 over [-14, 14] in fp32 arithmetic, two orders below the fp16 rounding of the stored
// activation).  relu runs on the ALU pipe; 12 FMA-pipe operations per pair including the bias add.
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

// returns half2 bits of (gelu(acc.x + bias.x), gelu(acc.y + bias.y))
__device__ __forceinline__ uint32_t gelu_pair_h2(uint64_t acc, uint64_t bias) {
    const uint64_t x = f2_add(acc, bias);
    float x0, x1;
    f2_unpack(x, x0, x1);
    const uint64_t a = f2_pack(fabsf(x0), fabsf(x1));
    uint64_t p = f2_fma(f2_splat(9.250150469597429e-05f), a, f2_splat(-9.215229511028156e-05f));
    p = f2_fma(p, a, f2_splat(0.00345434108749032f));
    p = f2_fma(p, a, f2_splat(0.02103373408317566f));
    p = f2_fma(p, a, f2_splat(0.04988996684551239f));
    p = f2_fma(p, a, f2_splat(1.f));
    float p0, p1, r0, r1;
    f2_unpack(p, p0, p1);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(p0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(p1));
    uint64_t r = f2_pack(r0, r1);
    r = f2_mul(r, r);
    r = f2_mul(r, r);
    r = f2_mul(r, r);
    r = f2_mul(r, r);                                          // (1 + ...)^-16 = erfc(|x|/sqrt 2)
    const uint64_t t = f2_mul(a, r);
    const uint64_t g = f2_fma(t, f2_splat(-0.5f), f2_pack(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));
    float g0, g1;
    f2_unpack(g, g0, g1);
    const __half2 hv = __floats2half2_rn(g0, g1);
    return *reinterpret_cast<const uint32_t*>(&hv);
}

}  // namespace m2t
