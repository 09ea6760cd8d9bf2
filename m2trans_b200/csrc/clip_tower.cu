// MedCLIP image-embedding pass (ref losses.py:53,68-77): bicubic resize to 224x224, Swin-T tower
// ('microsoft/swin-tiny-patch4-window7-224' as built by medclip's MedCLIPVisionModelViT; the architecture is cited from
// the Hugging Face implementation, transformers/models/swin/modeling_swin.py), pooled feature -> Linear(768,512) ->
// L2 normalise -> cosine logit against a text feature.
//
// Data flow per image (all buffers in HBM, token-major):
//   resize       fp32 [3][H][W]          -> bf16 patch rows [3136][48]   (the 4x4/4 patch conv's im2col, written directly)
//   patch embed  lin_umma K=48           -> fp32 X [3136][96], LayerNorm in place
//   12 x layer   LN -> bf16 | qkv GEMM -> bf16 [tokens][3C] | window attention (mma.sync, 49 tokens x 32 dims per head, shift and
//                window partition/reverse as index arithmetic) -> bf16 | proj GEMM, reduce-add into X | LN -> bf16 |
//                fc1 GEMM + GELU -> bf16 [tokens][4C] | fc2 GEMM, reduce-add into X  (stages 1-2: one fused kernel, mlp_umma.cu,
//                the [tokens][4C] tensor stays on the SM)
//   3 x merging  gather 2x2 + LN(4C) -> bf16 | reduction GEMM -> fp32 X' [tokens/4][2C]
//   final        LN(768) + mean over the 49 tokens + projection (8 CTAs per image, 64 outputs each), then L2 norm (+ logit)
// The residual stream X stays fp32; GEMM operands are bf16 with fp32 accumulation (north_star: "bf16 tcgen05 ViT forward").
#include <cuda_bf16.h>

#include "clip.cuh"

namespace m2t {

namespace {

constexpr int kDepth[CL_STAGES] = {2, 2, 6, 2};
constexpr int kHeads[CL_STAGES] = {3, 6, 12, 24};

// M2T_CLIP_UNFUSED_MLP=1: stages 1-2 run fc1 / fc2 as two Linear launches like stages 3-4 (A/B measurements)
bool clip_unfused_mlp() {
    static const bool v = [] { const char* e = getenv("M2T_CLIP_UNFUSED_MLP"); return e && e[0] == '1'; }();
    return v;
}

struct ClipBlockW {
    size_t ln1g, ln1b, wqkv, bqkv, rpb, wo, bo, ln2g, ln2b, w1, b1, w2, b2;
};
struct ClipMergeW {
    size_t ng, nb, red;
};
struct ClipLayout {
    size_t pew, peb, eng, enb;
    ClipBlockW blk[12];
    ClipMergeW mrg[3];
    size_t lng, lnb, projT;
    size_t total;
};

ClipLayout clip_layout() {
    ClipLayout L{};
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off = align_up(off + bytes, 256); return o; };
    L.pew = take(96 * 48 * 2); L.peb = take(96 * 4); L.eng = take(96 * 4); L.enb = take(96 * 4);
    int bi = 0;
    for (int s = 0; s < CL_STAGES; ++s) {
        const size_t C = (size_t)CL_EMBED << s;
        for (int b = 0; b < kDepth[s]; ++b, ++bi) {
            ClipBlockW& W = L.blk[bi];
            W.ln1g = take(C * 4); W.ln1b = take(C * 4);
            W.wqkv = take(3 * C * C * 2); W.bqkv = take(3 * C * 4);
            W.rpb = take((size_t)kHeads[s] * CL_WT * CL_RPB_PITCH * 4);
            W.wo = take(C * C * 2); W.bo = take(C * 4);
            W.ln2g = take(C * 4); W.ln2b = take(C * 4);
            W.w1 = take(4 * C * C * 2); W.b1 = take(4 * C * 4);
            W.w2 = take(4 * C * C * 2); W.b2 = take(C * 4);
        }
        if (s < CL_STAGES - 1) { L.mrg[s].ng = take(4 * C * 4); L.mrg[s].nb = take(4 * C * 4); L.mrg[s].red = take(8 * C * C * 2); }
    }
    L.lng = take(CL_FEAT * 4); L.lnb = take(CL_FEAT * 4); L.projT = take((size_t)CL_FEAT * CL_PROJ * 4);
    L.total = off;
    return L;
}

// ---- weight packing ------------------------------------------------------------------------------------
enum { CPK_F32 = 0, CPK_BF16 = 1, CPK_RPB = 2, CPK_PROJ_T = 3 };

__global__ void clip_pack_kernel(int mode, const float* __restrict__ src, void* __restrict__ dst, long n, int p0) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    switch (mode) {
        case CPK_F32: static_cast<float*>(dst)[i] = src[i]; break;
        case CPK_BF16: static_cast<__nv_bfloat16*>(dst)[i] = __float2bfloat16_rn(src[i]); break;
        case CPK_RPB: {   // dst [heads = p0][49][56] from the table [169][heads] (modeling_swin.py:398-428); pad columns -1e30
            const int j = (int)(i % CL_RPB_PITCH), q = (int)(i / CL_RPB_PITCH) % CL_WT, hd = (int)(i / (CL_WT * CL_RPB_PITCH));
            const int idx = (q / CL_WIN - j / CL_WIN + CL_WIN - 1) * (2 * CL_WIN - 1) + (q % CL_WIN - j % CL_WIN + CL_WIN - 1);
            static_cast<float*>(dst)[i] = j < CL_WT ? src[idx * p0 + hd] : -1.0e30f;
        } break;
        case CPK_PROJ_T: {  // dst [768][512] from the head's weight [512][768]
            const int nn = (int)(i % CL_PROJ), k = (int)(i / CL_PROJ);
            static_cast<float*>(dst)[i] = src[(long)nn * CL_FEAT + k];
        } break;
    }
}

int run_cpack(int mode, const float* src, void* dst, long n, int p0, cudaStream_t s) {
    clip_pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(mode, src, dst, n, p0);
    M2T_LAUNCH_CHECK("clip_pack_kernel");
    return M2T_OK;
}

// ---- bicubic resize (ref losses.py:53: F.interpolate(mode='bicubic', size=(224,224), align_corners=True)) ----------
// torch's upsample_bicubic2d: source = dst * (in-1)/(out-1), Keys kernel with A = -0.75, border taps clamped.
__device__ __forceinline__ void cubic_coeffs(float t, float c[4]) {
    constexpr float A = -0.75f;
    const float x0 = t + 1.f, x3 = 2.f - t, x2 = 1.f - t;
    c[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
    c[1] = ((A + 2.f) * t - (A + 3.f)) * t * t + 1.f;
    c[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
    c[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}

__global__ void __launch_bounds__(256)
clip_resize_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ rows, int B, int H, int W) {
    pdl_wait();
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)B * 3 * CL_IMG * CL_IMG) return;
    const int ox = (int)(i % CL_IMG), oy = (int)(i / CL_IMG) % CL_IMG, c = (int)(i / (CL_IMG * CL_IMG)) % 3;
    const long b = i / (3 * CL_IMG * CL_IMG);
    const float sy = (float)(H - 1) / (float)(CL_IMG - 1), sx = (float)(W - 1) / (float)(CL_IMG - 1);
    const float ry = sy * oy, rx = sx * ox;
    const int iy = (int)floorf(ry), ix = (int)floorf(rx);
    float cy[4], cx[4];
    cubic_coeffs(ry - iy, cy);
    cubic_coeffs(rx - ix, cx);
    const float* p = img + (b * 3 + c) * (long)H * W;
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int yy = min(max(iy - 1 + a, 0), H - 1);
        float r = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) r = fmaf(cx[e], __ldg(p + (long)yy * W + min(max(ix - 1 + e, 0), W - 1)), r);
        acc = fmaf(cy[a], r, acc);
    }
    const long row = (b * CL_GRID + oy / CL_PATCH) * CL_GRID + ox / CL_PATCH;
    rows[row * 48 + c * 16 + (oy % CL_PATCH) * 4 + (ox % CL_PATCH)] = __float2bfloat16_rn(acc);
}

// ---- LayerNorm: one warp per output row, the row held in registers ------------------------------------------------
// MERGE: output row (b, i, j) of the half-resolution grid is the concatenation of input tokens (2i,2j), (2i+1,2j),
// (2i,2j+1), (2i+1,2j+1) (modeling_swin.py:333-341), normalised over 4C.
__host__ __device__ constexpr int ln_rows_per_warp(int npl) { return npl <= 6 ? 4 : 2; }

template <int NPL, bool OUTF32, bool MERGE>
__global__ void __launch_bounds__(256)
clip_ln_kernel(const float* __restrict__ X, void* __restrict__ out, const float* __restrict__ g,
               const float* __restrict__ bta, long rows, int h, int w) {
    pdl_wait();
    constexpr int CT = NPL * 32;
    const int lane = threadIdx.x & 31;
    if constexpr (!MERGE && NPL <= 6) {
        // short rows (C = 96, 192): 4 rows per warp with all their loads in flight before the first reduction -- one
        // 384-byte row per warp left the kernel latency-bound at half of the copy bandwidth (16 -> 13 us at C = 96, batch 32;
        // C = 384 has too few rows per launch for it to matter and measured slightly worse)
        constexpr int R = ln_rows_per_warp(NPL);
        const long row0 = ((long)blockIdx.x * 8 + (threadIdx.x >> 5)) * R;
        float u[R][NPL];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int k = 0; k < NPL; ++k) u[r][k] = row0 + r < rows ? X[(row0 + r) * CT + lane + 32 * k] : 0.f;
        float gam[NPL], bet[NPL];
#pragma unroll
        for (int k = 0; k < NPL; ++k) { gam[k] = __ldg(g + lane + 32 * k); bet[k] = __ldg(bta + lane + 32 * k); }
        float mean[R], rstd[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < NPL; ++k) s += u[r][k];
            mean[r] = s;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1)
#pragma unroll
            for (int r = 0; r < R; ++r) mean[r] += __shfl_xor_sync(0xffffffffu, mean[r], o);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            mean[r] *= (1.f / CT);
            float q = 0.f;
#pragma unroll
            for (int k = 0; k < NPL; ++k) { const float d = u[r][k] - mean[r]; q = fmaf(d, d, q); }
            rstd[r] = q;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1)
#pragma unroll
            for (int r = 0; r < R; ++r) rstd[r] += __shfl_xor_sync(0xffffffffu, rstd[r], o);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (row0 + r >= rows) break;
            const float rs = rsqrtf(rstd[r] * (1.f / CT) + CL_LN_EPS);
#pragma unroll
            for (int k = 0; k < NPL; ++k) {
                const float y = fmaf((u[r][k] - mean[r]) * rs, gam[k], bet[k]);
                if constexpr (OUTF32) static_cast<float*>(out)[(row0 + r) * CT + lane + 32 * k] = y;
                else static_cast<__nv_bfloat16*>(out)[(row0 + r) * CT + lane + 32 * k] = __float2bfloat16_rn(y);
            }
        }
        return;
    }
    const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    float v[NPL];
    if constexpr (MERGE) {
        constexpr int CI = CT / 4;
        const int ho = h / 2, wo = w / 2;
        const long bimg = row / (ho * wo);
        const int rem = (int)(row - bimg * ho * wo), i = rem / wo, j = rem - i * wo;
#pragma unroll
        for (int k = 0; k < NPL; ++k) {
            const int e = lane + 32 * k, sg = e / CI, o = e - sg * CI;
            const long src = (bimg * h + 2 * i + (sg & 1)) * w + 2 * j + (sg >> 1);
            v[k] = X[src * CI + o];
        }
    } else {
#pragma unroll
        for (int k = 0; k < NPL; ++k) v[k] = X[row * CT + lane + 32 * k];
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NPL; ++k) s += v[k];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.f / CT);
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < NPL; ++k) { const float d = v[k] - mean; q = fmaf(d, d, q); }
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.f / CT) + CL_LN_EPS);
#pragma unroll
    for (int k = 0; k < NPL; ++k) {
        const int e = lane + 32 * k;
        const float y = fmaf((v[k] - mean) * rstd, __ldg(g + e), __ldg(bta + e));
        if constexpr (OUTF32) static_cast<float*>(out)[row * CT + e] = y;
        else static_cast<__nv_bfloat16*>(out)[row * CT + e] = __float2bfloat16_rn(y);
    }
}

template <int NPL, bool OUTF32, bool MERGE>
int launch_ln_t(const float* X, void* out, const float* g, const float* b, long rows, int h, int w, cudaStream_t s) {
    const long per_cta = 8L * ((!MERGE && NPL <= 6) ? ln_rows_per_warp(NPL) : 1);
    M2T_CUDA(launch_pdl(clip_ln_kernel<NPL, OUTF32, MERGE>, dim3((unsigned)((rows + per_cta - 1) / per_cta)), dim3(256), 0, s, X, out,
                        g, b, rows, h, w));
    return M2T_OK;
}

int launch_ln(const float* X, __nv_bfloat16* out, const float* g, const float* b, long rows, int C, cudaStream_t s) {
    switch (C) {
        case 96: return launch_ln_t<3, false, false>(X, out, g, b, rows, 0, 0, s);
        case 192: return launch_ln_t<6, false, false>(X, out, g, b, rows, 0, 0, s);
        case 384: return launch_ln_t<12, false, false>(X, out, g, b, rows, 0, 0, s);
        case 768: return launch_ln_t<24, false, false>(X, out, g, b, rows, 0, 0, s);
    }
    set_error("clip LN: channel count %d", C);
    return M2T_E_UNSUPPORTED;
}

int launch_ln_merge(const float* X, __nv_bfloat16* out, const float* g, const float* b, long rows_out, int C, int h, int w,
                    cudaStream_t s) {
    switch (C) {
        case 96: return launch_ln_t<12, false, true>(X, out, g, b, rows_out, h, w, s);
        case 192: return launch_ln_t<24, false, true>(X, out, g, b, rows_out, h, w, s);
        case 384: return launch_ln_t<48, false, true>(X, out, g, b, rows_out, h, w, s);
    }
    set_error("clip merge LN: channel count %d", C);
    return M2T_E_UNSUPPORTED;
}

// ---- final LN + mean pool + projection + L2 norm (+ logit) (modeling_swin.py:883-887; ref losses.py:71-77) ---------
__device__ __forceinline__ float block_sum_256(float v, float* red) {      // fixed order: deterministic
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k];
    return t;
}

constexpr int CL_PSLICE = 64;       // projection outputs per CTA: grid (B, 512 / 64); every CTA redoes the image's LN + pool

__global__ void __launch_bounds__(256)
clip_final_kernel(const float* __restrict__ X, const float* __restrict__ g, const float* __restrict__ bta,
                  const float* __restrict__ projT, float* __restrict__ embed, float* __restrict__ n2part) {
    __shared__ float part[8][CL_FEAT];
    __shared__ float pooled[CL_FEAT];
    __shared__ float red[8];
    pdl_wait();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long b = blockIdx.x;
    constexpr int NPL = CL_FEAT / 32;
    float acc[NPL];
#pragma unroll
    for (int k = 0; k < NPL; ++k) acc[k] = 0.f;
    for (int r = warp; r < CL_WT; r += 8) {
        const float* xr = X + (b * CL_WT + r) * CL_FEAT;
        float v[NPL];
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < NPL; ++k) { v[k] = xr[lane + 32 * k]; s += v[k]; }
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s * (1.f / CL_FEAT);
        float q = 0.f;
#pragma unroll
        for (int k = 0; k < NPL; ++k) { const float d = v[k] - mean; q = fmaf(d, d, q); }
#pragma unroll
        for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = rsqrtf(q * (1.f / CL_FEAT) + CL_LN_EPS);
#pragma unroll
        for (int k = 0; k < NPL; ++k) acc[k] += fmaf((v[k] - mean) * rstd, __ldg(g + lane + 32 * k), __ldg(bta + lane + 32 * k));
    }
#pragma unroll
    for (int k = 0; k < NPL; ++k) part[warp][lane + 32 * k] = acc[k];
    __syncthreads();
    for (int e = tid; e < CL_FEAT; e += 256) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += part[k][e];
        pooled[e] = t * (1.f / CL_WT);
    }
    __syncthreads();
    // projection: this CTA's 64 of the 512 outputs; thread = (4 outputs, 48 of the 768 inputs), float4 weight loads.  One CTA
    // per image with a 768-long dependent load loop per thread was latency-bound (40 us at batch 32 on 32 of the 148 SMs).
    const int c4 = tid & 15, ks = tid >> 4;
    float* pp = &part[0][0];                              // re-used: [16 k slices][64 outputs]
#pragma unroll 1
    for (int unit = blockIdx.y; unit < CL_PROJ / CL_PSLICE; unit += gridDim.y) {      // large batches: fewer, fatter CTAs
    const int c0 = unit * CL_PSLICE;
    float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int k = ks * 48; k < ks * 48 + 48; ++k) {
        const float pk = pooled[k];
        const float4 wv = __ldg(reinterpret_cast<const float4*>(projT + (long)k * CL_PROJ + c0) + c4);
        y.x = fmaf(pk, wv.x, y.x); y.y = fmaf(pk, wv.y, y.y); y.z = fmaf(pk, wv.z, y.z); y.w = fmaf(pk, wv.w, y.w);
    }
    __syncthreads();
    *reinterpret_cast<float4*>(pp + ks * CL_PSLICE + 4 * c4) = y;
    __syncthreads();
    float yo = 0.f;
    if (tid < CL_PSLICE) {
#pragma unroll
        for (int q = 0; q < 16; ++q) yo += pp[q * CL_PSLICE + tid];       // fixed order
        embed[b * CL_PROJ + c0 + tid] = yo;                                // un-normalised; clip_norm_logit_kernel finishes
    }
    const float n2 = block_sum_256(tid < CL_PSLICE ? yo * yo : 0.f, red);
    if (tid == 0) n2part[b * (CL_PROJ / CL_PSLICE) + unit] = n2;
    }
}

// embedding /= its L2 norm (partial sums of squares added in a fixed order); logit = embedding . text / |text|
__global__ void __launch_bounds__(CL_PROJ)
clip_norm_logit_kernel(float* __restrict__ embed, const float* __restrict__ n2part, const float* __restrict__ text,
                       float* __restrict__ logits) {
    __shared__ float red[2][CL_PROJ / 32];
    pdl_wait();
    const int tid = threadIdx.x;
    const long b = blockIdx.x;
    float n2 = 0.f;
#pragma unroll
    for (int c = 0; c < CL_PROJ / CL_PSLICE; ++c) n2 += n2part[b * (CL_PROJ / CL_PSLICE) + c];
    const float e = embed[b * CL_PROJ + tid] * (1.f / sqrtf(n2));
    embed[b * CL_PROJ + tid] = e;
    if (text == nullptr || logits == nullptr) return;
    const float t = __ldg(text + tid);
    float dot = e * t, tt = t * t;
#pragma unroll
    for (int o = 16; o; o >>= 1) { dot += __shfl_xor_sync(0xffffffffu, dot, o); tt += __shfl_xor_sync(0xffffffffu, tt, o); }
    if ((tid & 31) == 0) { red[0][tid >> 5] = dot; red[1][tid >> 5] = tt; }
    __syncthreads();
    if (tid == 0) {
        float d = 0.f, q = 0.f;
        for (int k = 0; k < CL_PROJ / 32; ++k) { d += red[0][k]; q += red[1][k]; }
        logits[b] = d / sqrtf(q);
    }
}

struct ClipWs {
    __nv_bfloat16 *Ape, *Hn, *QKV, *Ao, *G;
    float *X0, *X1, *n2part;
    size_t total;
};

ClipWs clip_ws(void* base, int B) {
    ClipWs w{};
    size_t off = 0;
    const size_t T = (size_t)B * CL_GRID * CL_GRID;
    auto take = [&](size_t bytes) { const size_t o = off; off = align_up(off + bytes, 1024); return static_cast<uint8_t*>(base) + o; };
    w.Ape = reinterpret_cast<__nv_bfloat16*>(take(T * 48 * 2));
    w.X0 = reinterpret_cast<float*>(take(T * CL_EMBED * 4));
    w.X1 = reinterpret_cast<float*>(take(T * CL_EMBED * 4 / 2));
    w.Hn = reinterpret_cast<__nv_bfloat16*>(take(T * CL_EMBED * 2));
    w.QKV = reinterpret_cast<__nv_bfloat16*>(take(T * 3 * CL_EMBED * 2));
    w.Ao = reinterpret_cast<__nv_bfloat16*>(take(T * CL_EMBED * 2));
    w.G = reinterpret_cast<__nv_bfloat16*>(take(T * 4 * CL_EMBED * 2));
    w.n2part = reinterpret_cast<float*>(take((size_t)B * (CL_PROJ / CL_PSLICE) * 4));
    w.total = off;
    return w;
}

}  // namespace
}  // namespace m2t

using namespace m2t;

extern "C" {

int m2t_clip_param_count(void) { return CL_NPARAMS; }

size_t m2t_clip_packed_bytes(void) { return clip_layout().total; }

size_t m2t_clip_workspace_bytes(int B) {
    if (B < 1) return 0;
    return clip_ws(nullptr, B).total;
}

int m2t_clip_pack_weights(const float* const* P, int n_params, void* d_packed, void* stream) {
    if (!P || !d_packed) { set_error("clip pack: null pointer"); return M2T_E_ARG; }
    if (n_params != CL_NPARAMS) { set_error("clip pack: expected %d tensors (Swin-T state_dict order + projection head), got %d", CL_NPARAMS, n_params); return M2T_E_ARG; }
    for (int i = 0; i < n_params; ++i)
        if (!P[i]) { set_error("clip pack: parameter %d is null", i); return M2T_E_ARG; }
    M2T_TRY(check_device());
    cudaStream_t s = (cudaStream_t)stream;
    const ClipLayout L = clip_layout();
    uint8_t* pk = static_cast<uint8_t*>(d_packed);
    M2T_TRY(run_cpack(CPK_BF16, P[0], pk + L.pew, 96 * 48, 0, s));      // [96][3][4][4] is already [96][c*16 + ky*4 + kx]
    M2T_TRY(run_cpack(CPK_F32, P[1], pk + L.peb, 96, 0, s));
    M2T_TRY(run_cpack(CPK_F32, P[2], pk + L.eng, 96, 0, s));
    M2T_TRY(run_cpack(CPK_F32, P[3], pk + L.enb, 96, 0, s));
    int pi = 4, bi = 0;
    for (int st = 0; st < CL_STAGES; ++st) {
        const long C = (long)CL_EMBED << st;
        for (int b = 0; b < kDepth[st]; ++b, ++bi, pi += 17) {
            const ClipBlockW& W = L.blk[bi];
            const float* const* Q = P + pi;   // ln1.w ln1.b table q.w q.b k.w k.b v.w v.b o.w o.b ln2.w ln2.b fc1.w fc1.b fc2.w fc2.b
            M2T_TRY(run_cpack(CPK_F32, Q[0], pk + W.ln1g, C, 0, s));
            M2T_TRY(run_cpack(CPK_F32, Q[1], pk + W.ln1b, C, 0, s));
            M2T_TRY(run_cpack(CPK_RPB, Q[2], pk + W.rpb, (long)kHeads[st] * CL_WT * CL_RPB_PITCH, kHeads[st], s));
            for (int j = 0; j < 3; ++j) {
                M2T_TRY(run_cpack(CPK_BF16, Q[3 + 2 * j], pk + W.wqkv + (size_t)j * C * C * 2, C * C, 0, s));
                M2T_TRY(run_cpack(CPK_F32, Q[4 + 2 * j], pk + W.bqkv + (size_t)j * C * 4, C, 0, s));
            }
            M2T_TRY(run_cpack(CPK_BF16, Q[9], pk + W.wo, C * C, 0, s));
            M2T_TRY(run_cpack(CPK_F32, Q[10], pk + W.bo, C, 0, s));
            M2T_TRY(run_cpack(CPK_F32, Q[11], pk + W.ln2g, C, 0, s));
            M2T_TRY(run_cpack(CPK_F32, Q[12], pk + W.ln2b, C, 0, s));
            M2T_TRY(run_cpack(CPK_BF16, Q[13], pk + W.w1, 4 * C * C, 0, s));
            M2T_TRY(run_cpack(CPK_F32, Q[14], pk + W.b1, 4 * C, 0, s));
            M2T_TRY(run_cpack(CPK_BF16, Q[15], pk + W.w2, 4 * C * C, 0, s));
            M2T_TRY(run_cpack(CPK_F32, Q[16], pk + W.b2, C, 0, s));
        }
        if (st < CL_STAGES - 1) {              // reduction.weight norm.weight norm.bias
            M2T_TRY(run_cpack(CPK_BF16, P[pi], pk + L.mrg[st].red, 8 * C * C, 0, s));
            M2T_TRY(run_cpack(CPK_F32, P[pi + 1], pk + L.mrg[st].ng, 4 * C, 0, s));
            M2T_TRY(run_cpack(CPK_F32, P[pi + 2], pk + L.mrg[st].nb, 4 * C, 0, s));
            pi += 3;
        }
    }
    M2T_TRY(run_cpack(CPK_F32, P[pi], pk + L.lng, CL_FEAT, 0, s));
    M2T_TRY(run_cpack(CPK_F32, P[pi + 1], pk + L.lnb, CL_FEAT, 0, s));
    M2T_TRY(run_cpack(CPK_PROJ_T, P[pi + 2], pk + L.projT, (long)CL_FEAT * CL_PROJ, 0, s));
    return M2T_OK;
}

int m2t_clip_stage_linear(int epilogue, const void* d_a, const void* d_w, const float* d_bias, void* d_out, int M, int N,
                          int K, void* stream) {
    if (!d_a || !d_w || !d_out) { set_error("clip linear: null pointer"); return M2T_E_ARG; }
    M2T_TRY(check_device());
    return launch_lin_umma(epilogue, d_a, d_w, d_bias, d_out, M, N, K, (cudaStream_t)stream);
}

int m2t_clip_stage_mlp(const void* d_a, const void* d_w1, const float* d_b1, const void* d_w2, const float* d_b2, float* d_x,
                       int M, int C, void* stream) {
    if (!d_a || !d_w1 || !d_b1 || !d_w2 || !d_b2 || !d_x) { set_error("clip mlp: null pointer"); return M2T_E_ARG; }
    M2T_TRY(check_device());
    return launch_mlp_umma(d_a, d_w1, d_b1, d_w2, d_b2, d_x, M, C, (cudaStream_t)stream);
}

int m2t_debug_lin_timing(long long* host128) {
    if (!host128) { set_error("lin timing: null pointer"); return M2T_E_ARG; }
    return read_lin_timing(host128);
}

int m2t_clip_stage_resize(const float* d_img, void* d_rows, int B, int H, int W, void* stream) {
    if (!d_img || !d_rows || B < 1 || H < 2 || W < 2) { set_error("clip resize: bad argument"); return M2T_E_ARG; }
    M2T_TRY(check_device());
    const long npx = (long)B * 3 * CL_IMG * CL_IMG;
    M2T_CUDA(launch_pdl(clip_resize_kernel, dim3((unsigned)((npx + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, d_img,
                        static_cast<__nv_bfloat16*>(d_rows), B, H, W));
    return M2T_OK;
}

int m2t_clip_stage_layernorm(const float* d_x, void* d_out, const float* d_gamma, const float* d_beta, int B, int h, int w,
                             int C, int merge, void* stream) {
    if (!d_x || !d_out || !d_gamma || !d_beta || B < 1 || h < 1 || w < 1) { set_error("clip layernorm: bad argument"); return M2T_E_ARG; }
    if (merge && ((h | w) & 1)) { set_error("clip layernorm: merging needs an even grid"); return M2T_E_ARG; }
    M2T_TRY(check_device());
    const long rows = (long)B * h * w;
    if (merge) return launch_ln_merge(d_x, static_cast<__nv_bfloat16*>(d_out), d_gamma, d_beta, rows / 4, C, h, w, (cudaStream_t)stream);
    return launch_ln(d_x, static_cast<__nv_bfloat16*>(d_out), d_gamma, d_beta, rows, C, (cudaStream_t)stream);
}

int m2t_clip_stage_attention(const void* d_qkv, void* d_out, const float* d_bias, int B, int h, int w, int C, int heads,
                             int shift, void* stream) {
    if (!d_qkv || !d_out || !d_bias || B < 1) { set_error("clip attention: bad argument"); return M2T_E_ARG; }
    if (h % CL_WIN || w % CL_WIN || heads * CL_HD != C || shift < 0 || shift >= CL_WIN) {
        set_error("clip attention: grid %d x %d, C %d, heads %d, shift %d", h, w, C, heads, shift);
        return M2T_E_UNSUPPORTED;
    }
    M2T_TRY(check_device());
    return launch_clip_attn(d_qkv, d_out, d_bias, B, h, w, C, heads, shift, (cudaStream_t)stream);
}

int m2t_clip_encode_image(const void* d_packed, const float* d_img, int B, int H, int W, float* d_embed,
                          const float* d_text, float* d_logits, void* d_workspace, void* stream) {
    if (!d_packed || !d_img || !d_embed || !d_workspace) { set_error("clip encode: null pointer"); return M2T_E_ARG; }
    if ((d_text == nullptr) != (d_logits == nullptr)) { set_error("clip encode: text feature and logits go together"); return M2T_E_ARG; }
    if (B < 1 || H < 2 || W < 2) { set_error("clip encode: bad shape B %d H %d W %d", B, H, W); return M2T_E_ARG; }
    M2T_TRY(check_device());
    cudaStream_t s = (cudaStream_t)stream;
    const ClipLayout L = clip_layout();
    const uint8_t* pk = static_cast<const uint8_t*>(d_packed);
    auto F = [&](size_t o) { return reinterpret_cast<const float*>(pk + o); };
    auto Hf = [&](size_t o) { return reinterpret_cast<const void*>(pk + o); };
    const ClipWs ws = clip_ws(d_workspace, B);

    const long npx = (long)B * 3 * CL_IMG * CL_IMG;
    M2T_CUDA(launch_pdl(clip_resize_kernel, dim3((unsigned)((npx + 255) / 256)), dim3(256), 0, s, d_img, ws.Ape, B, H, W));
    int h = CL_GRID, w = CL_GRID, C = CL_EMBED;
    long M = (long)B * h * w;
    M2T_TRY(launch_lin_umma(LIN_F32, ws.Ape, Hf(L.pew), F(L.peb), ws.X0, (int)M, C, 48, s));
    M2T_TRY((launch_ln_t<3, true, false>(ws.X0, ws.X0, F(L.eng), F(L.enb), M, 0, 0, s)));
    float* X = ws.X0;
    float* Xalt = ws.X1;
    int bi = 0;
    for (int st = 0; st < CL_STAGES; ++st) {
        for (int b = 0; b < kDepth[st]; ++b, ++bi) {
            const ClipBlockW& Wt = L.blk[bi];
            const int shift = (b % 2 == 1 && h > CL_WIN) ? CL_WIN / 2 : 0;      // modeling_swin.py:546-554, :1037
            M2T_TRY(launch_ln(X, ws.Hn, F(Wt.ln1g), F(Wt.ln1b), M, C, s));
            M2T_TRY(launch_lin_umma(LIN_BF16, ws.Hn, Hf(Wt.wqkv), F(Wt.bqkv), ws.QKV, (int)M, 3 * C, C, s));
            M2T_TRY(launch_clip_attn(ws.QKV, ws.Ao, F(Wt.rpb), B, h, w, C, kHeads[st], shift, s));
            M2T_TRY(launch_lin_umma(LIN_ADD_F32, ws.Ao, Hf(Wt.wo), F(Wt.bo), X, (int)M, C, C, s));
            M2T_TRY(launch_ln(X, ws.Hn, F(Wt.ln2g), F(Wt.ln2b), M, C, s));
            if (C <= 192 && !clip_unfused_mlp()) {       // stages 1-2 are HBM-bound: the [tokens][4C] tensor stays on the SM
                M2T_TRY(launch_mlp_umma(ws.Hn, Hf(Wt.w1), F(Wt.b1), Hf(Wt.w2), F(Wt.b2), X, (int)M, C, s));
            } else {
                M2T_TRY(launch_lin_umma(LIN_GELU_BF16, ws.Hn, Hf(Wt.w1), F(Wt.b1), ws.G, (int)M, 4 * C, C, s));
                M2T_TRY(launch_lin_umma(LIN_ADD_F32, ws.G, Hf(Wt.w2), F(Wt.b2), X, (int)M, C, 4 * C, s));
            }
        }
        if (st < CL_STAGES - 1) {
            M2T_TRY(launch_ln_merge(X, ws.Hn, F(L.mrg[st].ng), F(L.mrg[st].nb), M / 4, C, h, w, s));
            M2T_TRY(launch_lin_umma(LIN_F32, ws.Hn, Hf(L.mrg[st].red), nullptr, Xalt, (int)(M / 4), 2 * C, 4 * C, s));
            float* t = X; X = Xalt; Xalt = t;
            h /= 2; w /= 2; C *= 2; M /= 4;
        }
    }
    const int nsl = B >= 256 ? 1 : B >= 128 ? 2 : B >= 64 ? 4 : 8;      // every CTA re-reads its image's 150 KB: split only to fill the SMs
    M2T_CUDA(launch_pdl(clip_final_kernel, dim3((unsigned)B, (unsigned)nsl), dim3(256), 0, s, (const float*)X, F(L.lng),
                        F(L.lnb), F(L.projT), d_embed, ws.n2part));
    M2T_CUDA(launch_pdl(clip_norm_logit_kernel, dim3((unsigned)B), dim3(CL_PROJ), 0, s, d_embed, (const float*)ws.n2part, d_text,
                        d_logits));
    return M2T_OK;
}

}  // extern "C"
