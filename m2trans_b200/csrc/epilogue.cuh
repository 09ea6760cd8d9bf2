// Shared epilogue of the ff-conv kernels (CUDA-core and tcgen05 variants):
//   Xout = acc + bias + Xin      (ref M2Trans_network.py:164: feed_forward(xc) + x)
// written as rows of the fp32 NHWC residual stream, while accumulating the per-(image, channel) sum and sum of
// squares of Xout for the next CFTM's InstanceNorm (ref :135).
// 128 threads (t = 0..127) work on one 128-pixel tile; thread t owns channels 8*(t&7).. of pixels (t>>3)+16*it:
// 32 contiguous bytes per thread and access (256-bit loads/stores, one sector each), 8 threads per 256-byte pixel row.
// Statistics: every thread keeps fp32 partial sums over the tiles it sees of ONE image (<= a few hundred
// values), and the 128 threads flush them with one fp64 atomicAdd per channel when the image changes or the
// kernel ends (contended fp64 atomics per tile were the bottleneck of the first version of these kernels).
#pragma once
#include "common.cuh"

namespace m2t {

// fp32 words per staged pixel row: 68 = 4 (mod 32), so both the per-pixel float4 writes (thread = pixel, 8 lanes per
// wavefront hit 8 different bank quads) and the per-row float4 reads (8 threads = one row) are conflict-free
constexpr int EPI_LD = NF + 4;

// barrier over the 128 epilogue threads: the whole CTA (BAR == 0) or named barrier BAR
template <int BAR>
__device__ __forceinline__ void epi_sync() {
    if constexpr (BAR == 0) __syncthreads();
    else asm volatile("bar.sync %0, 128;" ::"n"(BAR) : "memory");
}

struct EpiStats {
    float s[8], s2[8];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int e = 0; e < 8; ++e) { s[e] = 0.f; s2[e] = 0.f; }
    }
};

// Residual rows of the tile: 8 independent 32-byte loads per thread.  Issued before the accumulator is
// ready, they keep 32 KB per epilogue warpgroup in flight while the MMAs run.
// `res` (last CFTM only, else nullptr): the head output rows for xr = fp16(res + x) are requested at the same time;
// loading them inside the store loop exposed one DRAM round trip per pixel row (measured: 10 K cycles per tile).
template <int TW>
__device__ __forceinline__ void epilogue_load_residual(int t, const float* Xin, int b, int y0, int x0, int Hp, int Wp,
                                                       uint4 xi[16], const float* __restrict__ res, uint4 rv[16]) {
    const int c8 = t & 7;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int p = (t >> 3) + 16 * it;
        const long pix = ((long)b * Hp + (y0 + p / TW)) * Wp + (x0 + p % TW);
        ldg256(Xin + pix * NF + 8 * c8, xi[2 * it], xi[2 * it + 1]);
    }
    if (res != nullptr) {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int p = (t >> 3) + 16 * it;
            const long pix = ((long)b * Hp + (y0 + p / TW)) * Wp + (x0 + p % TW);
            ldg256(res + pix * NF + 8 * c8, rv[2 * it], rv[2 * it + 1]);
        }
    }
}

// Os holds the 128-pixel x 64-channel fp32 tile (row pitch EPI_LD), pixel p at (y0 + p / TW, x0 + p % TW).  A barrier
// must separate the writes to Os from this call, and another one this call from the next writes to Os.
template <int TW>
__device__ __forceinline__ void epilogue_apply(const float* Os, int t, const uint4 xi[16], const float* __restrict__ bias,
                                               float* Xout, EpiStats& st, int b, int y0, int x0, int Hp, int Wp,
                                               const uint4 rv[16], __half* __restrict__ xr) {
    const int c8 = t & 7;
    const float4 b0 = *reinterpret_cast<const float4*>(bias + 8 * c8);
    const float4 b1 = *reinterpret_cast<const float4*>(bias + 8 * c8 + 4);
    const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int p = (t >> 3) + 16 * it;
        const long pix = ((long)b * Hp + (y0 + p / TW)) * Wp + (x0 + p % TW);
        const float4 o0 = *reinterpret_cast<const float4*>(Os + p * EPI_LD + 8 * c8);
        const float4 o1 = *reinterpret_cast<const float4*>(Os + p * EPI_LD + 8 * c8 + 4);
        const float ov[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
        const uint32_t xw[8] = {xi[2 * it].x, xi[2 * it].y, xi[2 * it].z, xi[2 * it].w,
                                xi[2 * it + 1].x, xi[2 * it + 1].y, xi[2 * it + 1].z, xi[2 * it + 1].w};
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = ov[e] + bv[e] + __uint_as_float(xw[e]);
        stg256(Xout + pix * NF + 8 * c8,
               make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3])),
               make_uint4(__float_as_uint(v[4]), __float_as_uint(v[5]), __float_as_uint(v[6]), __float_as_uint(v[7])));
        if (xr != nullptr) {   // last CFTM: also emit fp16(res + x), the tail's first GEMM operand (ref :70)
            const uint4 r0 = rv[2 * it], r1 = rv[2 * it + 1];
            const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
            uint4 u;
            uint32_t* pu = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const __half2 h = __floats2half2_rn(v[2 * e] + __uint_as_float(rw[2 * e]), v[2 * e + 1] + __uint_as_float(rw[2 * e + 1]));
                pu[e] = *reinterpret_cast<const uint32_t*>(&h);
            }
            *reinterpret_cast<uint4*>(xr + pix * NF + 8 * c8) = u;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) { st.s[e] += v[e]; st.s2[e] = fmaf(v[e], v[e], st.s2[e]); }
    }
}

// Adds the 128 threads' partial sums to stats[b] and clears them.  `red` is a [4][2][64] float scratch private
// to these threads; the call contains two barriers, so all 128 threads must reach it together.
template <int BAR>
__device__ __forceinline__ void epilogue_flush_stats(float (*red)[2][NF], int t, EpiStats& st, double* __restrict__ stats,
                                                     int b) {
    const int c8 = t & 7, lane = t & 31, wid = t >> 5;
#pragma unroll
    for (int e = 0; e < 8; ++e) {      // lanes with equal (lane & 7) hold the same channels
        st.s[e] += __shfl_xor_sync(0xffffffffu, st.s[e], 8);
        st.s2[e] += __shfl_xor_sync(0xffffffffu, st.s2[e], 8);
        st.s[e] += __shfl_xor_sync(0xffffffffu, st.s[e], 16);
        st.s2[e] += __shfl_xor_sync(0xffffffffu, st.s2[e], 16);
    }
    if (lane < 8) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { red[wid][0][8 * c8 + e] = st.s[e]; red[wid][1][8 * c8 + e] = st.s2[e]; }
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
