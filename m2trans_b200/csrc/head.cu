// Head stage: frame reflect-pad (ref M2Trans_network.py:78-86) + 3->64 3x3 reflect conv + bias
// (ref :34,:63), written as the fp32 NHWC residual stream, plus the InstanceNorm partial sums of
// the first CFTM (ref :127,:135).  K = 27 is too thin for the tensor cores: this stage is
// HBM-bound on paper (12 B in, 256 B out per pixel) and runs on the CUDA cores.
// One CTA = a 32 x 8 pixel tile; its 3 x 10 x 34 input patch is staged in shared memory with both reflections
// resolved.  A thread owns FOUR output channels and keeps their 27 x 4 weights in registers; a half-warp covers
// the 64 channels of one pixel, so the patch reads are shared-memory broadcasts (27 per pixel) and the arithmetic
// is 54 packed FFMA2 per pixel: FMA-bound, not LSU-bound like the first version that re-read the weights
// from shared memory for every pixel.  Each half-warp stores the 256 contiguous bytes of its pixel.
#include "common.cuh"

namespace m2t {

// reflect without repeating the edge, for a coordinate in [-1, n]
__device__ __forceinline__ int reflect1(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }
// frame padding: bottom/right only (ref :85)
__device__ __forceinline__ int frame_src(int i, int n) { return i < n ? i : 2 * (n - 1) - i; }

constexpr int HEAD_TW = 32, HEAD_TH = 4;          // tile; Hp and Wp are multiples of 32
constexpr int HEAD_THREADS = 32 * HEAD_TH;        // warp w computes row w of the tile

__device__ __forceinline__ uint64_t hd_pack(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ uint64_t hd_fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

__global__ void __launch_bounds__(HEAD_THREADS)
head_conv_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                 float* __restrict__ res, double* __restrict__ stats, int B, int H, int W, int Hp, int Wp, int resident) {
    __shared__ float sx[3][HEAD_TH + 2][HEAD_TW + 2];
    __shared__ float red[HEAD_TH][2][NF];
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int cq = lane & 15, half = lane >> 4;
    // weights of this thread's 4 channels, as packed pairs: constants, loaded before the dependency wait
    uint64_t w01[27], w23[27];
#pragma unroll
    for (int k = 0; k < 27; ++k) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w + k * NF + 4 * cq));
        w01[k] = hd_pack(wv.x, wv.y);
        w23[k] = hd_pack(wv.z, wv.w);
    }
    const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + 4 * cq));
    pdl_trigger_last_wave(resident); // the CTAs of the last wave admit the next kernel's launch: it finds no free slot before they leave
    pdl_wait();

    const int tiles_x = Wp / HEAD_TW, tiles_y = Hp / HEAD_TH;
    const int b = blockIdx.x / (tiles_x * tiles_y), r = blockIdx.x - b * tiles_x * tiles_y;
    const int y0 = (r / tiles_x) * HEAD_TH, x0 = (r % tiles_x) * HEAD_TW;
    const float* xb = x + (long)b * 3 * H * W;
    for (int i = t; i < 3 * (HEAD_TH + 2) * (HEAD_TW + 2); i += HEAD_THREADS) {
        const int c = i / ((HEAD_TH + 2) * (HEAD_TW + 2)), rr = i - c * (HEAD_TH + 2) * (HEAD_TW + 2);
        const int py = rr / (HEAD_TW + 2), px = rr - py * (HEAD_TW + 2);
        const int sy = frame_src(reflect1(y0 + py - 1, Hp), H);
        const int sxx = frame_src(reflect1(x0 + px - 1, Wp), W);
        sx[c][py][px] = __ldg(xb + ((long)c * H + sy) * W + sxx);
    }
    __syncthreads();

    float s[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    float* orow = res + (((long)b * Hp + y0 + warp) * Wp + x0) * NF + 4 * cq;
#pragma unroll 2
    for (int i = 0; i < HEAD_TW / 2; ++i) {
        const int px = 2 * i + half;
        uint64_t a01 = hd_pack(bv.x, bv.y), a23 = hd_pack(bv.z, bv.w);
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float v = sx[c][warp + ky][px + kx];
                    const uint64_t vv = hd_pack(v, v);
                    a01 = hd_fma2(vv, w01[c * 9 + ky * 3 + kx], a01);
                    a23 = hd_fma2(vv, w23[c * 9 + ky * 3 + kx], a23);
                }
        float o[4];
        asm("mov.b64 {%0, %1}, %2;" : "=f"(o[0]), "=f"(o[1]) : "l"(a01));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(o[2]), "=f"(o[3]) : "l"(a23));
        *reinterpret_cast<float4*>(orow + (long)px * NF) = make_float4(o[0], o[1], o[2], o[3]);
#pragma unroll
        for (int e = 0; e < 4; ++e) { s[e] += o[e]; s2[e] = fmaf(o[e], o[e], s2[e]); }
    }

    // lanes l and l ^ 16 hold the same channels; one slot per (warp, channel), fixed summation order below
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        s[e] += __shfl_xor_sync(0xffffffffu, s[e], 16);
        s2[e] += __shfl_xor_sync(0xffffffffu, s2[e], 16);
    }
    if (half == 0) {
#pragma unroll
        for (int e = 0; e < 4; ++e) { red[warp][0][4 * cq + e] = s[e]; red[warp][1][4 * cq + e] = s2[e]; }
    }
    __syncthreads();
    if (t < 2 * NF) {
        const int c = t >> 1, k = t & 1;
        double tot = 0.0;
#pragma unroll
        for (int wv = 0; wv < HEAD_TH; ++wv) tot += (double)red[wv][k][c];
        atomicAdd(&stats[((long)b * NF + c) * 2 + k], tot);
    }
}

int launch_head(const float* x, const float* w, const float* b, float* res, double* stats, const Geom& g,
                cudaStream_t s) {
    const int tiles = g.B * (g.Hp / HEAD_TH) * (g.Wp / HEAD_TW);
    M2T_CUDA(launch_pdl(head_conv_kernel, dim3((unsigned)tiles), dim3(HEAD_THREADS), 0, s, x, w, b, res, stats, g.B, g.H,
                        g.W, g.Hp, g.Wp, resident_ctas(head_conv_kernel, HEAD_THREADS, 0)));
    return M2T_OK;
}

// (sum, sumsq) -> (mean, 1/sqrt(var+eps)), biased variance (ref :127 nn.InstanceNorm2d defaults)
__global__ void stats_finalize_kernel(const double* __restrict__ stats, float2* __restrict__ munorm, int n,
                                      double inv_npix) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double m = stats[2 * i] * inv_npix;
    double var = stats[2 * i + 1] * inv_npix - m * m;
    if (var < 0.0) var = 0.0;
    munorm[i] = make_float2((float)m, (float)(1.0 / sqrt(var + (double)IN_EPS)));
}

int launch_stats_finalize(const double* stats, float2* munorm, int B, int npix, cudaStream_t s) {
    const int n = B * NF;
    M2T_CUDA(launch_pdl(stats_finalize_kernel, dim3(cdiv(n, 128)), dim3(128), 0, s, stats, munorm, n, 1.0 / (double)npix));
    return M2T_OK;
}

}  // namespace m2t
