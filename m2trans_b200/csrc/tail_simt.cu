// Tail (ref M2Trans_network.py:40-56, :70-76), CUDA-core variant (M2T_VAR_SIMT_TAIL) + the border kernel
// shared by both variants.
//   tail_up        : 1x1 conv + bias -> PixelShuffle(r) -> exact (erff) GELU, fp16 NHWC at r x resolution
//   reflect_border : fills the 1-pixel ring of a [B][h+2][w+2][64] tensor with the reflection of its interior
//                    (row -1 = row 1, row h = row h-2, same for columns), i.e. padding_mode='reflect' of the
//                    last conv (ref :48/:55) materialised once so that its loader never needs index math
//   tail_out       : last 3x3 conv 64->3 without bias, clamp to [0, rgb_range] (ref :74), crop (ref :76),
//                    written straight into the caller's fp32 NCHW output
// PixelShuffle: out[c, r*y+u, r*x+v] = in[c*r*r + u*r + v, y, x]; the packed weight rows are already
// sub-pixel-major (row (u*r+v)*64 + c), see pack.cu.
#include "common.cuh"

namespace m2t {

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

constexpr int TU_PX = 32;

// blockDim.x = N = 64*r*r ; thread n owns packed output row n of the 1x1 conv for TU_PX pixels
__global__ void __launch_bounds__(576)
tail_up_simt_kernel(const __half* __restrict__ Ain, const __half* __restrict__ Wt, const float* __restrict__ bias,
                    __half* __restrict__ out, int h, int w, int r, int pad) {
    extern __shared__ __align__(16) uint8_t smem[];
    float* As = reinterpret_cast<float*>(smem);                          // [32][64]
    __half* Os = reinterpret_cast<__half*>(smem + TU_PX * NF * 4);       // [r][32 r][64]
    const int n = threadIdx.x, N = blockDim.x;
    const long pix0 = (long)blockIdx.x * TU_PX;                          // first pixel (linear in [B,h,w])
    for (int i = n; i < TU_PX * NF; i += N) As[i] = __half2float(Ain[pix0 * NF + i]);
    float wreg[NF];
    {
        const uint4* wp = reinterpret_cast<const uint4*>(Wt + (long)n * NF);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const uint4 u = __ldg(wp + c);
            const __half2* hp = reinterpret_cast<const __half2*>(&u);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = __half22float2(hp[i]);
                wreg[c * 8 + 2 * i] = f.x; wreg[c * 8 + 2 * i + 1] = f.y;
            }
        }
    }
    const float bn = bias[n];
    const int uv = n / NF, c = n - uv * NF, u = uv / r, v = uv - u * r;
    __syncthreads();
    for (int p = 0; p < TU_PX; ++p) {
        float acc = bn;
#pragma unroll
        for (int k = 0; k < NF; k += 4) {
            const float4 a = *reinterpret_cast<const float4*>(&As[p * NF + k]);
            acc = fmaf(a.x, wreg[k], acc); acc = fmaf(a.y, wreg[k + 1], acc);
            acc = fmaf(a.z, wreg[k + 2], acc); acc = fmaf(a.w, wreg[k + 3], acc);
        }
        Os[((long)(u * TU_PX * r + r * p + v)) * NF + c] = __float2half_rn(gelu_erf(acc));
    }
    __syncthreads();
    // coalesced copy-out: r output rows of 32 r pixels (128 B each)
    const int b = (int)(pix0 / ((long)h * w));
    const int rem = (int)(pix0 - (long)b * h * w);
    const int y = rem / w, x0 = rem - y * w;
    const int row_u4 = TU_PX * r * 8;                                    // uint4 per output row segment
    const long opitch = (long)(w * r + 2 * pad) * NF;
    const uint4* src = reinterpret_cast<const uint4*>(Os);
    for (int i = n; i < r * row_u4; i += N) {
        const int uu = i / row_u4, k = i - uu * row_u4;
        __half* dst = out + ((long)b * (h * r + 2 * pad) + (long)y * r + uu + pad) * opitch + ((long)x0 * r + pad) * NF;
        reinterpret_cast<uint4*>(dst)[k] = src[i];
    }
}

int launch_tail_up_simt(const __half* Ain, const __half* Wt, const float* bias, __half* out, int B, int h, int w, int r,
                        int pad, cudaStream_t s) {
    if (w % TU_PX) { set_error("tail_up: width %d not a multiple of %d", w, TU_PX); return M2T_E_ARG; }
    if (r < 2 || r > 3) { set_error("tail_up: shuffle factor %d", r); return M2T_E_UNSUPPORTED; }
    const size_t smem = (size_t)TU_PX * NF * 4 + (size_t)r * r * TU_PX * NF * 2;
    M2T_ENSURE_SMEM(tail_up_simt_kernel, 64 * 1024);
    const long npix = (long)B * h * w;
    tail_up_simt_kernel<<<(unsigned)(npix / TU_PX), NF * r * r, smem, s>>>(Ain, Wt, bias, out, h, w, r, pad);
    M2T_LAUNCH_CHECK("tail_up_simt_kernel");
    return M2T_OK;
}

// ---- reflected border ring --------------------------------------------------------------------------------
// One thread per (ring pixel, 16-byte chunk).  Ring pixels of a (h+2) x (w+2) frame: 2 (w+2) + 2 h.
__global__ void reflect_border_kernel(__half* __restrict__ T, int B, int h, int w) {
    const int ring = 2 * (w + 2) + 2 * h;
    pdl_trigger();
    pdl_wait();
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long)B * ring * 8) return;
    const int ch = (int)(idx & 7);
    const long pi = idx >> 3;
    const int b = (int)(pi / ring), k = (int)(pi - (long)b * ring);
    int py, px;                                   // coordinates in the padded frame
    if (k < w + 2) { py = 0; px = k; }
    else if (k < 2 * (w + 2)) { py = h + 1; px = k - (w + 2); }
    else if (k < 2 * (w + 2) + h) { py = k - 2 * (w + 2) + 1; px = 0; }
    else { py = k - 2 * (w + 2) - h + 1; px = w + 1; }
    const int iy = py - 1, ix = px - 1;           // interior coordinates in [-1, h] x [-1, w]
    const int sy = iy < 0 ? 1 : (iy >= h ? h - 2 : iy);
    const int sx = ix < 0 ? 1 : (ix >= w ? w - 2 : ix);
    const long pitch = (long)(w + 2) * NF;
    __half* base = T + (long)b * (h + 2) * pitch;
    *reinterpret_cast<uint4*>(base + (long)py * pitch + (long)px * NF + ch * 8) =
        *reinterpret_cast<const uint4*>(base + (long)(sy + 1) * pitch + (long)(sx + 1) * NF + ch * 8);
}

int launch_reflect_border(__half* T, int B, int h, int w, cudaStream_t s) {
    const long total = (long)B * (2 * (w + 2) + 2 * h) * 8;
    M2T_CUDA(launch_pdl(reflect_border_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, s, T, B, h, w));
    return M2T_OK;
}

// ---- final conv (CUDA cores) ----------------------------------------------------------------------------------
constexpr int TO_T = 16;                 // 16x16 output pixels per CTA
constexpr int TO_LD = NF + 8;            // halves per staged pixel
constexpr size_t TO_SMEM = (size_t)(TO_T + 2) * (TO_T + 2) * TO_LD * 2 + 9 * 3 * NF * 4;

// T: fp16 [Bc][hp+2][wp+2][64] with its reflected border ring filled; Wc packed fp16 [9][16][64] (rows 3.. zero)
__global__ void __launch_bounds__(TO_T* TO_T)
tail_out_simt_kernel(const __half* __restrict__ T, const __half* __restrict__ Wc, float* __restrict__ y, int hp, int wp,
                     int hout, int wout, int b0, float rgb_range) {
    extern __shared__ __align__(16) uint8_t smem[];
    __half* Ts = reinterpret_cast<__half*>(smem);
    float* Ws = reinterpret_cast<float*>(smem + (size_t)(TO_T + 2) * (TO_T + 2) * TO_LD * 2);  // [9][3][64]
    const int t = threadIdx.x;
    const int bl = blockIdx.z, y0 = blockIdx.y * TO_T, x0 = blockIdx.x * TO_T;
    if (y0 >= hout || x0 >= wout) return;            // tile entirely in the cropped-away padding
    for (int i = t; i < 9 * 3 * NF; i += TO_T * TO_T) {
        const int tap = i / (3 * NF), rem = i - tap * 3 * NF;
        Ws[i] = __half2float(Wc[(tap * 16 + rem / NF) * NF + rem % NF]);
    }
    const long pitch = (long)(wp + 2) * NF;
    for (int idx = t; idx < (TO_T + 2) * (TO_T + 2) * 8; idx += TO_T * TO_T) {
        const int p = idx >> 3, ch = idx & 7;
        const int py = y0 + p / (TO_T + 2), px = x0 + p % (TO_T + 2);       // padded-frame coordinates
        *reinterpret_cast<uint4*>(&Ts[p * TO_LD + ch * 8]) =
            *reinterpret_cast<const uint4*>(T + ((long)bl * (hp + 2) + py) * pitch + (long)px * NF + ch * 8);
    }
    __syncthreads();
    const int ty = t / TO_T, tx = t % TO_T;
    float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
        const __half* src = &Ts[((ty + tap / 3) * (TO_T + 2) + tx + tap % 3) * TO_LD];
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
            const uint4 u = *reinterpret_cast<const uint4*>(src + ch * 8);
            const __half2* hp2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = __half22float2(hp2[i]);
#pragma unroll
                for (int o = 0; o < 3; ++o) {
                    const float* wr = &Ws[(tap * 3 + o) * NF + ch * 8 + 2 * i];
                    acc[o] = fmaf(f.x, wr[0], acc[o]);
                    acc[o] = fmaf(f.y, wr[1], acc[o]);
                }
            }
        }
    }
    const int oy = y0 + ty, ox = x0 + tx;
    if (oy < hout && ox < wout) {
#pragma unroll
        for (int o = 0; o < 3; ++o)
            y[(((long)(b0 + bl) * 3 + o) * hout + oy) * wout + ox] = fminf(fmaxf(acc[o], 0.f), rgb_range);
    }
}

int launch_tail_out_simt(const __half* T, const __half* Wc, float* y, int Bc, int hp, int wp, int hout, int wout,
                         int b0, float rgb_range, cudaStream_t s) {
    M2T_ENSURE_SMEM(tail_out_simt_kernel, TO_SMEM);
    dim3 grid(wp / TO_T, hp / TO_T, Bc);
    tail_out_simt_kernel<<<grid, TO_T * TO_T, TO_SMEM, s>>>(T, Wc, y, hp, wp, hout, wout, b0, rgb_range);
    M2T_LAUNCH_CHECK("tail_out_simt_kernel");
    return M2T_OK;
}

}  // namespace m2t
