// Shared epilogue of the ff-conv kernels (CUDA-core and tcgen05 variants):
//   Xout = acc + bias + Xin      (ref M2Trans_network.py:164: feed_forward(xc) + x)
// written as coalesced float4 rows of the fp32 NHWC residual stream, while accumulating the
// per-(image, channel) sum and sum of squares of Xout for the next CFTM's InstanceNorm (ref :135).
// Partial sums: <= 16 values per thread in fp32, 8 threads combined in fp32, then one fp64 atomicAdd per
// channel per CTA, so the 138k-element reductions of the largest frames keep ~1e-7 relative accuracy.
#pragma once
#include "common.cuh"

namespace m2t {

constexpr int EPI_LD = NF + 1;   // fp32 words per staged pixel row (odd: conflict-free column writes)

// barrier over the 128 epilogue threads: the whole CTA (BAR == 0) or named barrier BAR
template <int BAR>
__device__ __forceinline__ void epi_sync() {
    if constexpr (BAR == 0) __syncthreads();
    else asm volatile("bar.sync %0, 128;" ::"n"(BAR) : "memory");
}

// Called by threads 0..127.  Os holds a 128-pixel x 64-channel fp32 tile, pixel p at (y0 + p / TW, x0 + p % TW).
// A barrier must separate the writes to Os from this call, and another one this call from the next writes.
template <int TW, int BAR>
__device__ __forceinline__ void epilogue_residual_stats(const float* Os, const float* __restrict__ bias,
                                                        const float* Xin, float* Xout,  /* may alias (in-place) */
                                                        double* __restrict__ stats, int b, int y0, int x0, int Hp,
                                                        int Wp, const float* __restrict__ res = nullptr,
                                                        __half* __restrict__ xr = nullptr) {
    __shared__ float red[4][2][NF];
    const int t = threadIdx.x, c4 = t & 15, lane = t & 31, wid = t >> 5;
    const float4 bv = *reinterpret_cast<const float4*>(bias + 4 * c4);
    float s[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    // all 16 residual loads first: 32 KB in flight per CTA keeps the memory system busy
    float4 xi[16];
#pragma unroll
    for (int it = 0; it < 16; ++it) {
        const int p = (t >> 4) + 8 * it;
        const long pix = ((long)b * Hp + (y0 + p / TW)) * Wp + (x0 + p % TW);
        xi[it] = *reinterpret_cast<const float4*>(Xin + pix * NF + 4 * c4);
    }
#pragma unroll
    for (int it = 0; it < 16; ++it) {
        const int p = (t >> 4) + 8 * it;
        const long pix = ((long)b * Hp + (y0 + p / TW)) * Wp + (x0 + p % TW);
        const float* o = Os + p * EPI_LD + 4 * c4;
        float v[4];
        v[0] = o[0] + bv.x + xi[it].x; v[1] = o[1] + bv.y + xi[it].y;
        v[2] = o[2] + bv.z + xi[it].z; v[3] = o[3] + bv.w + xi[it].w;
        *reinterpret_cast<float4*>(Xout + pix * NF + 4 * c4) = make_float4(v[0], v[1], v[2], v[3]);
        if (xr != nullptr) {   // last CFTM: also emit fp16(res + x), the tail's first GEMM operand (ref :70)
            const float4 rv = *reinterpret_cast<const float4*>(res + pix * NF + 4 * c4);
            const __half2 h0 = __floats2half2_rn(v[0] + rv.x, v[1] + rv.y);
            const __half2 h1 = __floats2half2_rn(v[2] + rv.z, v[3] + rv.w);
            uint2 u;
            u.x = *reinterpret_cast<const uint32_t*>(&h0);
            u.y = *reinterpret_cast<const uint32_t*>(&h1);
            *reinterpret_cast<uint2*>(xr + pix * NF + 4 * c4) = u;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) { s[e] += v[e]; s2[e] = fmaf(v[e], v[e], s2[e]); }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        s[e] += __shfl_xor_sync(0xffffffffu, s[e], 16);
        s2[e] += __shfl_xor_sync(0xffffffffu, s2[e], 16);
    }
    if (lane < 16) {
#pragma unroll
        for (int e = 0; e < 4; ++e) { red[wid][0][4 * c4 + e] = s[e]; red[wid][1][4 * c4 + e] = s2[e]; }
    }
    epi_sync<BAR>();
    {
        const int c = t >> 1, k = t & 1;
        const double tot = (double)red[0][k][c] + (double)red[1][k][c] + (double)red[2][k][c] + (double)red[3][k][c];
        atomicAdd(&stats[((long)b * NF + c) * 2 + k], tot);
    }
}

}  // namespace m2t
