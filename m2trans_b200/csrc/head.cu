// Head stage: frame reflect-pad (ref M2Trans_network.py:78-86) + 3->64 3x3 reflect conv + bias
// (ref :34,:63), written as the fp32 NHWC residual stream, plus the InstanceNorm partial sums of
// the first CFTM (ref :127,:135).  K = 27 is too thin for the tensor cores: this stage is
// HBM-bound (12 B in, 256 B out per pixel) and runs on the CUDA cores.
// One CTA = 32 consecutive pixels of one padded row: the 3 x 3 x 34 input patch is staged in shared memory
// once (both reflections resolved there), 4 threads per pixel each produce 16 channels and store 4 float4
// (64 B contiguous per pixel and instruction).
#include "common.cuh"

namespace m2t {

// reflect without repeating the edge, for a coordinate in [-1, n]
__device__ __forceinline__ int reflect1(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }
// frame padding: bottom/right only (ref :85)
__device__ __forceinline__ int frame_src(int i, int n) { return i < n ? i : 2 * (n - 1) - i; }

constexpr int HEAD_PX = 32;  // pixels per CTA (4 threads per pixel); Wp is a multiple of 32

constexpr int HEAD_CHUNKS = 8;    // chunks of HEAD_PX pixels per CTA: 256 pixels, always inside one image

__global__ void __launch_bounds__(HEAD_PX * 4)
head_conv_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                 float* __restrict__ res, double* __restrict__ stats, int B, int H, int W, int Hp, int Wp) {
    __shared__ float sw[27 * NF];
    __shared__ float sb[NF];
    __shared__ float sx[2][3][3][HEAD_PX + 2];
    __shared__ float red[HEAD_PX * 4 / 32][2][NF];
    const int t = threadIdx.x;
    for (int i = t; i < 27 * NF; i += HEAD_PX * 4) sw[i] = w[i];
    if (t < NF) sb[t] = bias[t];
    pdl_wait();          // weights above are constants; everything below touches activations / statistics

    const int npix = Hp * Wp;
    const long cta_px0 = (long)blockIdx.x * HEAD_PX * HEAD_CHUNKS;
    const int b = (int)(cta_px0 / npix);
    const float* xb = x + (long)b * 3 * H * W;
    const int q = t & 3, p = t >> 2;
    // per-thread InstanceNorm partial sums of its 16 channels over all chunks: one set of atomics per CTA
    float ssum[4][4], ssq[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) { ssum[j][e] = 0.f; ssq[j][e] = 0.f; }

    auto stage = [&](int chunk, int buf) {
        const int rem = (int)(cta_px0 + (long)chunk * HEAD_PX - (long)b * npix);
        const int y = rem / Wp, x0 = rem - y * Wp;
        for (int i = t; i < 9 * (HEAD_PX + 2); i += HEAD_PX * 4) {
            const int c = i / (3 * (HEAD_PX + 2)), r = i - c * 3 * (HEAD_PX + 2);
            const int ky = r / (HEAD_PX + 2), col = r - ky * (HEAD_PX + 2);
            const int sy = frame_src(reflect1(y + ky - 1, Hp), H);
            const int sxx = frame_src(reflect1(x0 + col - 1, Wp), W);
            sx[buf][c][ky][col] = __ldg(xb + ((long)c * H + sy) * W + sxx);
        }
    };
    stage(0, 0);
    __syncthreads();
    for (int chunk = 0; chunk < HEAD_CHUNKS; ++chunk) {
        const int buf = chunk & 1;
        if (chunk + 1 < HEAD_CHUNKS) stage(chunk + 1, buf ^ 1);
        float acc[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 bv = *reinterpret_cast<const float4*>(&sb[4 * (q + 4 * j)]);
            acc[j][0] = bv.x; acc[j][1] = bv.y; acc[j][2] = bv.z; acc[j][3] = bv.w;
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float v = sx[buf][c][ky][p + kx];
                    const float* wr = &sw[(c * 9 + ky * 3 + kx) * NF];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 wv = *reinterpret_cast<const float4*>(&wr[4 * (q + 4 * j)]);
                        acc[j][0] = fmaf(v, wv.x, acc[j][0]);
                        acc[j][1] = fmaf(v, wv.y, acc[j][1]);
                        acc[j][2] = fmaf(v, wv.z, acc[j][2]);
                        acc[j][3] = fmaf(v, wv.w, acc[j][3]);
                    }
                }
            }
        }
        float* o = res + (cta_px0 + (long)chunk * HEAD_PX + p) * NF;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            *reinterpret_cast<float4*>(o + 4 * (q + 4 * j)) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
#pragma unroll
            for (int e = 0; e < 4; ++e) { ssum[j][e] += acc[j][e]; ssq[j][e] = fmaf(acc[j][e], acc[j][e], ssq[j][e]); }
        }
        __syncthreads();          // next chunk's patch is staged; this chunk's patch may be overwritten
    }

    // lanes with equal (lane & 3) hold the same 16 channels
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float s = ssum[j][e], s2 = ssq[j][e];
#pragma unroll
            for (int m = 4; m < 32; m <<= 1) {
                s += __shfl_xor_sync(0xffffffffu, s, m);
                s2 += __shfl_xor_sync(0xffffffffu, s2, m);
            }
            if ((t & 31) < 4) {       // one slot per (warp, channel): fixed summation order below
                red[t >> 5][0][4 * (q + 4 * j) + e] = s;
                red[t >> 5][1][4 * (q + 4 * j) + e] = s2;
            }
        }
    }
    __syncthreads();
    if (t < 2 * NF) {
        const int c = t >> 1, k = t & 1;
        double tot = 0.0;
#pragma unroll
        for (int wv = 0; wv < HEAD_PX * 4 / 32; ++wv) tot += (double)red[wv][k][c];
        atomicAdd(&stats[((long)b * NF + c) * 2 + k], tot);
    }
    pdl_trigger();       // multi-wave grid: admit the next kernel only as this one drains
}

int launch_head(const float* x, const float* w, const float* b, float* res, double* stats, const Geom& g,
                cudaStream_t s) {
    const long total = (long)g.B * g.Hp * g.Wp;            // multiple of 1024: Hp and Wp are multiples of 32
    M2T_CUDA(launch_pdl(head_conv_kernel, dim3((unsigned)(total / (HEAD_PX * HEAD_CHUNKS))), dim3(HEAD_PX * 4), 0, s, x, w, b,
                        res, stats, g.B, g.H, g.W, g.Hp, g.Wp));
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
