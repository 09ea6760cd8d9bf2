// GMSD of the reference's test loop on the device (ref test.py:98-99: piq.gmsd(hr, sr, data_range=1.,
// reduction='none'); `piq` is a third-party package that is absent offline, so this restates its published algorithm --
// Xue, Zhang, Mou, Bovik, "Gradient Magnitude Similarity Deviation", IEEE TIP 2014, as piq implements it):
//   luma  = 0.299 R + 0.587 G + 0.114 B of x / data_range            (the Y row of piq's rgb2yiq; 1-channel input as is)
//   pad bottom and right by p = max(H % 2, W % 2) zeros, 2 x 2 average pooling (stride 2)
//   gradient magnitude with the Prewitt pair [[1,0,-1]]*3 / 3 and its transpose, zero padding 1:  g = sqrt(gx^2 + gy^2 + 1e-12)
//   GMS   = (2 g_x g_y + c) / (g_x^2 + g_y^2 + c),  c = 170 / 255^2
//   GMSD  = sqrt(mean((GMS - mean(GMS))^2))  per image
// Two passes: pooled luma of both images into the workspace, then one thread per pooled pixel; sums in fp64, fixed order.
#include "common.cuh"

namespace m2t {
namespace {

constexpr int GM_THREADS = 256;

__device__ __forceinline__ float gm_luma(const float* p, long plane, int colors, float inv_range) {
    if (colors == 1) return p[0] * inv_range;
    return (0.299f * p[0] + 0.587f * p[plane] + 0.114f * p[2 * plane]) * inv_range;
}

__global__ void __launch_bounds__(GM_THREADS)
gmsd_pool_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ la, float* __restrict__ lb,
                 int B, int colors, int H, int W, int H2, int W2, float inv_range) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)B * H2 * W2) return;
    const int x2 = (int)(i % W2), y2 = (int)((i / W2) % H2), img = (int)(i / ((long)W2 * H2));
    const long plane = (long)H * W;
    const float* pa = a + (long)img * colors * plane;
    const float* pb = b + (long)img * colors * plane;
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            const int y = 2 * y2 + dy, x = 2 * x2 + dx;
            if (y < H && x < W) {                            // the padded row / column is zero
                sa += gm_luma(pa + (long)y * W + x, plane, colors, inv_range);
                sb += gm_luma(pb + (long)y * W + x, plane, colors, inv_range);
            }
        }
    la[i] = 0.25f * sa;
    lb[i] = 0.25f * sb;
}

__device__ __forceinline__ float gm_grad(const float* __restrict__ l, int y, int x, int H2, int W2) {
    float v[3][3];
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            const int yy = y + dy, xx = x + dx;
            v[dy + 1][dx + 1] = (yy >= 0 && yy < H2 && xx >= 0 && xx < W2) ? l[(long)yy * W2 + xx] : 0.f;
        }
    const float gx = (v[0][0] + v[1][0] + v[2][0] - v[0][2] - v[1][2] - v[2][2]) * (1.f / 3.f);
    const float gy = (v[0][0] + v[0][1] + v[0][2] - v[2][0] - v[2][1] - v[2][2]) * (1.f / 3.f);
    return sqrtf(gx * gx + gy * gy + 1e-12f);
}

// one block per (image, slab of pooled pixels): partial {sum, sum of squares} of the GMS map in fp64
__global__ void __launch_bounds__(GM_THREADS)
gmsd_map_kernel(const float* __restrict__ la, const float* __restrict__ lb, double2* __restrict__ partials, int H2, int W2,
                int blocks_per_img) {
    __shared__ double rs[GM_THREADS / 32], rq[GM_THREADS / 32];
    const int img = blockIdx.x / blocks_per_img, blk = blockIdx.x - img * blocks_per_img;
    const long n = (long)H2 * W2;
    const float* pa = la + (long)img * n;
    const float* pb = lb + (long)img * n;
    const float c = 170.f / (255.f * 255.f);
    double s = 0.0, q = 0.0;
    for (long i = (long)blk * GM_THREADS + threadIdx.x; i < n; i += (long)blocks_per_img * GM_THREADS) {
        const int y = (int)(i / W2), x = (int)(i - (long)y * W2);
        const float ga = gm_grad(pa, y, x, H2, W2), gb = gm_grad(pb, y, x, H2, W2);
        const float gms = (2.f * ga * gb + c) / (ga * ga + gb * gb + c);
        s += (double)gms;
        q += (double)gms * (double)gms;
    }
#pragma unroll
    for (int m = 16; m; m >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, m); q += __shfl_xor_sync(0xffffffffu, q, m); }
    if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = s; rq[threadIdx.x >> 5] = q; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ts = 0.0, tq = 0.0;
        for (int w = 0; w < GM_THREADS / 32; ++w) { ts += rs[w]; tq += rq[w]; }
        partials[blockIdx.x] = make_double2(ts, tq);
    }
}

__global__ void gmsd_final_kernel(const double2* __restrict__ partials, float* __restrict__ out, int blocks_per_img, double n) {
    const int img = blockIdx.x * blockDim.x + threadIdx.x;
    if (img >= (int)gridDim.x * (int)blockDim.x) return;
    double s = 0.0, q = 0.0;
    for (int k = 0; k < blocks_per_img; ++k) { const double2 p = partials[(long)img * blocks_per_img + k]; s += p.x; q += p.y; }
    const double mean = s / n;
    double var = q / n - mean * mean;
    if (var < 0.0) var = 0.0;
    out[img] = (float)sqrt(var);
}

inline void gm_geom(int H, int W, int* H2, int* W2, int* blocks) {
    const int p = (H % 2) > (W % 2) ? (H % 2) : (W % 2);
    *H2 = (H + p) / 2;
    *W2 = (W + p) / 2;
    const long n = (long)*H2 * *W2;
    long b = (n + 4 * GM_THREADS - 1) / (4 * GM_THREADS);
    *blocks = (int)(b < 1 ? 1 : (b > 256 ? 256 : b));
}

}  // namespace
}  // namespace m2t

using namespace m2t;

extern "C" {

size_t m2t_gmsd_workspace_bytes(int B, int H, int W) {
    if (B < 1 || H < 2 || W < 2) return 0;
    int H2, W2, blocks;
    gm_geom(H, W, &H2, &W2, &blocks);
    return align_up((size_t)2 * B * H2 * W2 * sizeof(float), 256) + (size_t)B * blocks * sizeof(double2);
}

int m2t_eval_gmsd(const float* d_x, const float* d_y, int B, int colors, int H, int W, float data_range, float* d_out,
                  void* d_workspace, void* stream) {
    if (!d_x || !d_y || !d_out || !d_workspace) { set_error("gmsd: null pointer"); return M2T_E_ARG; }
    if (B < 1 || (colors != 1 && colors != 3) || H < 2 || W < 2 || !(data_range > 0.f)) {
        set_error("gmsd: bad B %d / colors %d / %dx%d / data_range %g", B, colors, H, W, (double)data_range);
        return M2T_E_ARG;
    }
    M2T_TRY(check_device());
    cudaStream_t s = (cudaStream_t)stream;
    int H2, W2, blocks;
    gm_geom(H, W, &H2, &W2, &blocks);
    const long n = (long)B * H2 * W2;
    float* la = static_cast<float*>(d_workspace);
    float* lb = la + n;
    double2* partials = reinterpret_cast<double2*>(static_cast<uint8_t*>(d_workspace) + align_up((size_t)2 * n * sizeof(float), 256));
    gmsd_pool_kernel<<<(unsigned)((n + GM_THREADS - 1) / GM_THREADS), GM_THREADS, 0, s>>>(d_x, d_y, la, lb, B, colors, H, W, H2, W2, 1.f / data_range);
    M2T_LAUNCH_CHECK("gmsd_pool_kernel");
    gmsd_map_kernel<<<(unsigned)(B * blocks), GM_THREADS, 0, s>>>(la, lb, partials, H2, W2, blocks);
    M2T_LAUNCH_CHECK("gmsd_map_kernel");
    gmsd_final_kernel<<<(unsigned)B, 1, 0, s>>>(partials, d_out, blocks, (double)H2 * W2);
    M2T_LAUNCH_CHECK("gmsd_final_kernel");
    return M2T_OK;
}

}  // extern "C"
