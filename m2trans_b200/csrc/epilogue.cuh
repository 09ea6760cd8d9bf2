// Shared epilogue of the ff-conv kernels (CUDA-core and tcgen05 variants):
//   Xout = acc + bias + Xin      (ref M2Trans_network.py:164: feed_forward(xc) + x)
// written as coalesced float4 rows of the fp32 NHWC residual stream, while accumulating the
// per-(image, channel) sum and sum of squares of Xout for the next CFTM's InstanceNorm (ref :135).
// 128 threads (t = 0..127) work on one 128-pixel tile; thread t owns channels 4*(t&15).. of pixels (t>>4)+8*it.
// Statistics: every thread keeps fp32 partial sums over the tiles it sees of ONE image (<= a few hundred
// values), and the 128 threads flush them with one fp64 atomicAdd per channel when the image changes or the
// kernel ends (contended fp64 atomics per tile were the bottleneck of the first version of these kernels).
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

struct EpiStats {
    float s[4], s2[4];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int e = 0; e < 4; ++e) { s[e] = 0.f; s2[e] = 0.f; }
    }
};

// Residual rows of the tile: 16 independent 16-byte loads per thread.  Issued before the accumulator is
// ready, they keep 32 KB per epilogue warpgroup in flight while the MMAs run.
template <int TW>
__device__ __forceinline__ void epilogue_load_residual(int t, const float* Xin, int b, int y0, int x0, int Hp, int Wp,
                                                       float4 xi[16]) {
    const int c4 = t & 15;
#pragma unroll
    for (int it = 0; it < 16; ++it) {
        const int p = (t >> 4) + 8 * it;
        const long pix = ((long)b * Hp + (y0 + p / TW)) * Wp + (x0 + p % TW);
        xi[it] = *reinterpret_cast<const float4*>(Xin + pix * NF + 4 * c4);
    }
}

// Os holds the 128-pixel x 64-channel fp32 tile, pixel p at (y0 + p / TW, x0 + p % TW).  A barrier must separate
// the writes to Os from this call, and another one this call from the next writes to Os.
template <int TW>
__device__ __forceinline__ void epilogue_apply(const float* Os, int t, const float4 xi[16], const float* __restrict__ bias,
                                               float* Xout, EpiStats& st, int b, int y0, int x0, int Hp, int Wp,
                                               const float* __restrict__ res, __half* __restrict__ xr) {
    const int c4 = t & 15;
    const float4 bv = *reinterpret_cast<const float4*>(bias + 4 * c4);
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
        for (int e = 0; e < 4; ++e) { st.s[e] += v[e]; st.s2[e] = fmaf(v[e], v[e], st.s2[e]); }
    }
}

// Adds the 128 threads' partial sums to stats[b] and clears them.  `red` is a [4][2][64] float scratch private
// to these threads; the call contains two barriers, so all 128 threads must reach it together.
template <int BAR>
__device__ __forceinline__ void epilogue_flush_stats(float (*red)[2][NF], int t, EpiStats& st, double* __restrict__ stats,
                                                     int b) {
    const int c4 = t & 15, lane = t & 31, wid = t >> 5;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        st.s[e] += __shfl_xor_sync(0xffffffffu, st.s[e], 16);
        st.s2[e] += __shfl_xor_sync(0xffffffffu, st.s2[e], 16);
    }
    if (lane < 16) {
#pragma unroll
        for (int e = 0; e < 4; ++e) { red[wid][0][4 * c4 + e] = st.s[e]; red[wid][1][4 * c4 + e] = st.s2[e]; }
    }
    epi_sync<BAR>();
    {
        const int c = t >> 1, k = t & 1;
        const double tot = (double)red[0][k][c] + (double)red[1][k][c] + (double)red[2][k][c] + (double)red[3][k][c];
        atomicAdd(&stats[((long)b * NF + c) * 2 + k], tot);
    }
    epi_sync<BAR>();
    st.clear();
}

}  // namespace m2t
