// rlutrans.TransBlock (ref /root/reference/util/rlutrans.py:70-87; EffAttention :30-67; Mlp :11-27), SURVEY §8 a15.
//     x <- x + proj(BDAttn(qkv(reduce(LN1(x)))))          BDAttn: 8 heads of 8 channels, softmax inside chunks of
//     x <- x + fc2(ReLU(fc1(LN2(x))))                      floor(N/16) consecutive tokens (ref :53-63)
// Everything is token-local except the attention, which is chunk-local, so the block is three kernels:
//   tb_qkv   : LN1 -> reduce (64x64) -> qkv (64x192), 64 tokens per tile, weights resident in smem
//   tb_attn  : one CTA per (batch, chunk, head); K/V of the chunk staged in smem, one query row per thread
//   tb_out   : proj + bias + residual -> LN2 -> fc1 + ReLU -> fc2 + bias + residual
// fp32 CUDA-core arithmetic throughout: the contractions have K = 8 (attention) and K = 16/64 (linears), the whole
// block is 28 MFLOP per 256-token chunk against 128 KB of traffic, and the reference is fp32, so parity here is
// fp32-exact up to summation order (tests: max-abs <= 2e-5) instead of the fp16-operand bar of the SR path.
// Supported: dim = 64, num_heads = 8 (the constructor defaults, ref :72-73; nothing in the reference
// instantiates any other shape), any B >= 1, N >= 16.
#include "common.cuh"

namespace m2t {

constexpr int TB_DIM = 64;
constexpr int TB_HEADS = 8;
constexpr int TB_HD = 8;
constexpr int TB_HID = 16;          // Mlp hidden = dim / 4 (ref :81)
constexpr int TB_TOK = 64;          // tokens per tile
constexpr int TB_THREADS = 128;
constexpr int TB_PITCH = TB_DIM + 4;     // 68 = 4 (mod 32): float4 rows, see tile_linear
constexpr int TB_HPITCH = TB_HID + 4;
constexpr float TB_LN_EPS = 1e-5f;  // nn.LayerNorm default

// acc[t][j] += sum_k in[ty + 8 t][k] * wT[k][o0 + tx*4 + j]     (ty = tid / 16, tx = tid % 16)
// Token rows are interleaved (ty + 8 t) so the two rows a warp touches per t are adjacent: with a pitch of
// 4 (mod 32) words their float4 reads fall in different banks.  K multiple of 4, in_pitch multiple of 4.
template <int K>
__device__ __forceinline__ void tile_linear(const float* __restrict__ s_in, int in_pitch, const float* __restrict__ s_wT,
                                            int ldw, int o0, float acc[8][4]) {
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const float* a = s_in + ty * in_pitch;
    const float* w = s_wT + o0 + tx * 4;
#pragma unroll 2
    for (int k = 0; k < K; k += 4) {
        float4 wv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) wv[i] = *reinterpret_cast<const float4*>(w + (k + i) * ldw);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const float4 av = *reinterpret_cast<const float4*>(a + t * 8 * in_pitch + k);
            const float ak[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc[t][0] = fmaf(ak[i], wv[i].x, acc[t][0]);
                acc[t][1] = fmaf(ak[i], wv[i].y, acc[t][1]);
                acc[t][2] = fmaf(ak[i], wv[i].z, acc[t][2]);
                acc[t][3] = fmaf(ak[i], wv[i].w, acc[t][3]);
            }
        }
    }
}

// W [out][in] (nn.Linear layout) -> smem transposed [in][out]
__device__ __forceinline__ void load_wT(float* s_wT, const float* __restrict__ w, int out, int in) {
    for (int i = threadIdx.x; i < out * in; i += TB_THREADS) {
        const int o = i / in, k = i - o * in;
        s_wT[k * out + o] = __ldg(w + i);
    }
}

// tile of 64 tokens x 64 channels, global [tokens][64] -> smem [64][68]; rows past `total` are zero
__device__ __forceinline__ void load_tile(float* s, const float* __restrict__ g, long tok0, long total) {
    for (int i = threadIdx.x; i < TB_TOK * TB_DIM / 4; i += TB_THREADS) {
        const int t = i >> 4, c4 = (i & 15) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tok0 + t < total) v = *reinterpret_cast<const float4*>(g + (tok0 + t) * TB_DIM + c4);
        *reinterpret_cast<float4*>(s + t * TB_PITCH + c4) = v;
    }
}

// LayerNorm over the 64 channels of each of the 64 tokens, in place (dst may equal src); 4 warps x 16 tokens
__device__ __forceinline__ void layer_norm_tile(const float* src, float* dst, const float* __restrict__ gamma,
                                                const float* __restrict__ beta) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float g0 = __ldg(gamma + lane), g1 = __ldg(gamma + lane + 32);
    const float b0 = __ldg(beta + lane), b1 = __ldg(beta + lane + 32);
    for (int t = warp * 16; t < warp * 16 + 16; ++t) {
        const float x0 = src[t * TB_PITCH + lane], x1 = src[t * TB_PITCH + lane + 32];
        float s = x0 + x1;
#pragma unroll
        for (int m = 16; m; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
        const float mean = s * (1.f / TB_DIM);
        const float d0 = x0 - mean, d1 = x1 - mean;
        float v = d0 * d0 + d1 * d1;
#pragma unroll
        for (int m = 16; m; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
        const float rstd = rsqrtf(v * (1.f / TB_DIM) + TB_LN_EPS);
        dst[t * TB_PITCH + lane] = fmaf(d0 * rstd, g0, b0);
        dst[t * TB_PITCH + lane + 32] = fmaf(d1 * rstd, g1, b1);
    }
}

struct TbParams {           // TransBlock.state_dict() order
    const float* reduce_w;  // atten.reduce.weight [64][64]
    const float* qkv_w;     // atten.qkv.weight    [192][64]
    const float* proj_w;    // atten.proj.weight   [64][64]
    const float* proj_b;    // atten.proj.bias     [64]
    const float* ln1_w;     // norm1.weight / bias [64]
    const float* ln1_b;
    const float* fc1_w;     // mlp.fc1.weight [16][64], bias [16]
    const float* fc1_b;
    const float* fc2_w;     // mlp.fc2.weight [64][16], bias [64]
    const float* fc2_b;
    const float* ln2_w;     // norm2.weight / bias [64]
    const float* ln2_b;
};

constexpr int TBQ_SMEM = (TB_DIM * TB_DIM + TB_DIM * 3 * TB_DIM + 2 * TB_TOK * TB_PITCH) * 4;

__global__ void __launch_bounds__(TB_THREADS)
tb_qkv_kernel(const float* __restrict__ x, TbParams P, float* __restrict__ qkv, long total, int with_ln) {
    extern __shared__ __align__(16) float smf[];
    float* s_wr = smf;                               // [64][64]
    float* s_wq = s_wr + TB_DIM * TB_DIM;            // [64][192]
    float* s_x = s_wq + TB_DIM * 3 * TB_DIM;         // [64][68]
    float* s_r = s_x + TB_TOK * TB_PITCH;            // [64][68]
    load_wT(s_wr, P.reduce_w, TB_DIM, TB_DIM);
    load_wT(s_wq, P.qkv_w, 3 * TB_DIM, TB_DIM);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const long tiles = (total + TB_TOK - 1) / TB_TOK;
    for (long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long tok0 = tile * TB_TOK;
        __syncthreads();                             // weights ready / previous tile's s_x, s_r consumed
        load_tile(s_x, x, tok0, total);
        __syncthreads();
        if (with_ln) {                               // standalone EffAttention.forward (ref :47) starts at reduce
            layer_norm_tile(s_x, s_x, P.ln1_w, P.ln1_b); // ref :85 norm1
            __syncthreads();
        }
        float acc[8][4];
#pragma unroll
        for (int t = 0; t < 8; ++t) { acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f; }
        tile_linear<TB_DIM>(s_x, TB_PITCH, s_wr, TB_DIM, 0, acc);          // ref :48 reduce (no bias)
#pragma unroll
        for (int t = 0; t < 8; ++t)
            *reinterpret_cast<float4*>(s_r + (ty + 8 * t) * TB_PITCH + tx * 4) = make_float4(acc[t][0], acc[t][1], acc[t][2], acc[t][3]);
        __syncthreads();
#pragma unroll 1
        for (int slab = 0; slab < 3; ++slab) {                             // ref :50 qkv (no bias): q | k | v
#pragma unroll
            for (int t = 0; t < 8; ++t) { acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f; }
            tile_linear<TB_DIM>(s_r, TB_PITCH, s_wq, 3 * TB_DIM, slab * TB_DIM, acc);
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const long tok = tok0 + ty + 8 * t;
                if (tok < total)
                    *reinterpret_cast<float4*>(qkv + tok * (3 * TB_DIM) + slab * TB_DIM + tx * 4) =
                        make_float4(acc[t][0], acc[t][1], acc[t][2], acc[t][3]);
            }
        }
    }
}

// Block-diagonal attention (ref :53-63): chunk c of batch b holds tokens [c*n, min((c+1)*n, N)).
// One query row per thread, keys/values of the chunk staged in smem TBA_KT at a time, ONE pass with an online
// softmax over blocks of 8 keys (the running maximum only moves in the first few blocks, so the rescale branch
// is rarely taken).
constexpr int TBA_KT = 512;          // keys staged at a time (32 KB of K and V)
constexpr int TBA_THREADS = 256;

__global__ void __launch_bounds__(TBA_THREADS)
tb_attn_kernel(const float* __restrict__ qkv, float* __restrict__ o, int N, int n, int nchunks, float scale_log2e) {
    __shared__ float4 sk[TBA_KT][2];
    __shared__ float4 sv[TBA_KT][2];
    const int bc = blockIdx.x, head = blockIdx.y;
    const int b = bc / nchunks, c = bc - b * nchunks;
    const int t0 = c * n, len = min(n, N - t0);
    const float* base = qkv + ((long)b * N + t0) * (3 * TB_DIM) + head * TB_HD;
    // every thread walks the same number of query rounds so the staging barriers stay uniform
    const int rounds = (len + TBA_THREADS - 1) / TBA_THREADS;
    for (int r = 0; r < rounds; ++r) {
        const int qi = r * TBA_THREADS + threadIdx.x;
        const bool live = qi < len;
        float q[8];
        {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c4 = a;
            if (live) {
                a = *reinterpret_cast<const float4*>(base + (long)qi * (3 * TB_DIM));
                c4 = *reinterpret_cast<const float4*>(base + (long)qi * (3 * TB_DIM) + 4);
            }
            // softmax((q.k) * scale) = exp2((q.k) * scale * log2 e - max): fold the constant into q
            q[0] = a.x * scale_log2e; q[1] = a.y * scale_log2e; q[2] = a.z * scale_log2e; q[3] = a.w * scale_log2e;
            q[4] = c4.x * scale_log2e; q[5] = c4.y * scale_log2e; q[6] = c4.z * scale_log2e; q[7] = c4.w * scale_log2e;
        }
        float m = -INFINITY, l = 0.f, acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int k0 = 0; k0 < len; k0 += TBA_KT) {
            const int kn = min(TBA_KT, len - k0);
            __syncthreads();
            for (int i = threadIdx.x; i < kn * 2; i += TBA_THREADS) {
                const float* kp = base + (long)(k0 + (i >> 1)) * (3 * TB_DIM) + (i & 1) * 4;
                sk[i >> 1][i & 1] = *reinterpret_cast<const float4*>(kp + TB_DIM);
                sv[i >> 1][i & 1] = *reinterpret_cast<const float4*>(kp + 2 * TB_DIM);
            }
            __syncthreads();
            for (int j0 = 0; j0 < kn; j0 += 8) {
                float sc[8], bm = -INFINITY;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int j = min(j0 + u, kn - 1);          // tail keys are clamped here and masked below
                    const float4 k0v = sk[j][0], k1v = sk[j][1];
                    float sdot = q[0] * k0v.x;
                    sdot = fmaf(q[1], k0v.y, sdot); sdot = fmaf(q[2], k0v.z, sdot); sdot = fmaf(q[3], k0v.w, sdot);
                    sdot = fmaf(q[4], k1v.x, sdot); sdot = fmaf(q[5], k1v.y, sdot); sdot = fmaf(q[6], k1v.z, sdot);
                    sdot = fmaf(q[7], k1v.w, sdot);
                    sc[u] = sdot;
                    bm = fmaxf(bm, sdot);
                }
                if (bm > m) {                                    // rescale what has been accumulated so far
                    const float f = exp2f(m - bm);               // m = -inf on the first block: f = 0
                    l *= f;
#pragma unroll
                    for (int d = 0; d < 8; ++d) acc[d] *= f;
                    m = bm;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int j = min(j0 + u, kn - 1);
                    const float p = (j0 + u < kn) ? exp2f(sc[u] - m) : 0.f;
                    l += p;
                    const float4 v0 = sv[j][0], v1 = sv[j][1];
                    acc[0] = fmaf(p, v0.x, acc[0]); acc[1] = fmaf(p, v0.y, acc[1]);
                    acc[2] = fmaf(p, v0.z, acc[2]); acc[3] = fmaf(p, v0.w, acc[3]);
                    acc[4] = fmaf(p, v1.x, acc[4]); acc[5] = fmaf(p, v1.y, acc[5]);
                    acc[6] = fmaf(p, v1.z, acc[6]); acc[7] = fmaf(p, v1.w, acc[7]);
                }
            }
        }
        if (live) {
            const float inv = 1.f / l;
            float* op = o + ((long)b * N + t0 + qi) * TB_DIM + head * TB_HD;   // heads concatenated (ref :61-64)
            *reinterpret_cast<float4*>(op) = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
            *reinterpret_cast<float4*>(op + 4) = make_float4(acc[4] * inv, acc[5] * inv, acc[6] * inv, acc[7] * inv);
        }
    }
}

constexpr int TBO_SMEM = (TB_DIM * TB_DIM + TB_DIM * TB_HID + TB_HID * TB_DIM + 2 * TB_TOK * TB_PITCH + TB_TOK * TB_HPITCH) * 4;

__global__ void __launch_bounds__(TB_THREADS)
tb_out_kernel(const float* __restrict__ x, const float* __restrict__ o, TbParams P, float* __restrict__ y, long total, int mode) {
    // mode 0: the TransBlock tail (proj + residual, LN2, Mlp + residual);  mode 1: EffAttention.forward's tail, y = proj(o)
    // (ref :65);  mode 2: Mlp.forward, y = fc2(ReLU(fc1(x))) (ref :21-27; x comes in through `o`)
    extern __shared__ __align__(16) float smf[];
    float* s_wp = smf;                               // proj^T [64][64]
    float* s_w1 = s_wp + TB_DIM * TB_DIM;            // fc1^T  [64][16]
    float* s_w2 = s_w1 + TB_DIM * TB_HID;            // fc2^T  [16][64]
    float* s_a = s_w2 + TB_HID * TB_DIM;             // [64][68]
    float* s_x1 = s_a + TB_TOK * TB_PITCH;           // [64][68]
    float* s_h = s_x1 + TB_TOK * TB_PITCH;           // [64][20]
    if (mode != 2) load_wT(s_wp, P.proj_w, TB_DIM, TB_DIM);
    if (mode != 1) { load_wT(s_w1, P.fc1_w, TB_HID, TB_DIM); load_wT(s_w2, P.fc2_w, TB_DIM, TB_HID); }
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float4 bp = make_float4(0.f, 0.f, 0.f, 0.f), b2 = bp;
    if (mode != 2) bp = make_float4(__ldg(P.proj_b + tx * 4), __ldg(P.proj_b + tx * 4 + 1), __ldg(P.proj_b + tx * 4 + 2), __ldg(P.proj_b + tx * 4 + 3));
    if (mode != 1) b2 = make_float4(__ldg(P.fc2_b + tx * 4), __ldg(P.fc2_b + tx * 4 + 1), __ldg(P.fc2_b + tx * 4 + 2), __ldg(P.fc2_b + tx * 4 + 3));
    const long tiles = (total + TB_TOK - 1) / TB_TOK;
    for (long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long tok0 = tile * TB_TOK;
        __syncthreads();
        load_tile(s_a, o, tok0, total);
        __syncthreads();
        float x1[8][4];
#pragma unroll
        for (int t = 0; t < 8; ++t) { x1[t][0] = bp.x; x1[t][1] = bp.y; x1[t][2] = bp.z; x1[t][3] = bp.w; }
        if (mode != 2) {
            tile_linear<TB_DIM>(s_a, TB_PITCH, s_wp, TB_DIM, 0, x1);       // ref :65 proj
            if (mode == 1) {                                                // EffAttention.forward ends here
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const long tok = tok0 + ty + 8 * t;
                    if (tok < total) *reinterpret_cast<float4*>(y + tok * TB_DIM + tx * 4) = make_float4(x1[t][0], x1[t][1], x1[t][2], x1[t][3]);
                }
                continue;
            }
#pragma unroll
            for (int t = 0; t < 8; ++t) {                                   // ref :85 x + atten(...)
                const long tok = tok0 + ty + 8 * t;
                float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (tok < total) xv = *reinterpret_cast<const float4*>(x + tok * TB_DIM + tx * 4);
                x1[t][0] += xv.x; x1[t][1] += xv.y; x1[t][2] += xv.z; x1[t][3] += xv.w;
                *reinterpret_cast<float4*>(s_x1 + (ty + 8 * t) * TB_PITCH + tx * 4) = make_float4(x1[t][0], x1[t][1], x1[t][2], x1[t][3]);
            }
            __syncthreads();
            layer_norm_tile(s_x1, s_a, P.ln2_w, P.ln2_b);                   // ref :86 norm2
            __syncthreads();
        } else {
#pragma unroll
            for (int t = 0; t < 8; ++t) { x1[t][0] = x1[t][1] = x1[t][2] = x1[t][3] = 0.f; }   // no residual in Mlp.forward
        }
        {   // fc1 + ReLU (ref :22-23): thread -> token tid/2, hidden units (tid%2)*8 .. +7
            const int tok = threadIdx.x >> 1, h0 = (threadIdx.x & 1) * 8;
            float hacc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) hacc[j] = __ldg(P.fc1_b + h0 + j);
            for (int k = 0; k < TB_DIM; ++k) {
                const float av = s_a[tok * TB_PITCH + k];
                const float4 w0 = *reinterpret_cast<const float4*>(s_w1 + k * TB_HID + h0);
                const float4 w1 = *reinterpret_cast<const float4*>(s_w1 + k * TB_HID + h0 + 4);
                hacc[0] = fmaf(av, w0.x, hacc[0]); hacc[1] = fmaf(av, w0.y, hacc[1]);
                hacc[2] = fmaf(av, w0.z, hacc[2]); hacc[3] = fmaf(av, w0.w, hacc[3]);
                hacc[4] = fmaf(av, w1.x, hacc[4]); hacc[5] = fmaf(av, w1.y, hacc[5]);
                hacc[6] = fmaf(av, w1.z, hacc[6]); hacc[7] = fmaf(av, w1.w, hacc[7]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) s_h[tok * TB_HPITCH + h0 + j] = fmaxf(hacc[j], 0.f);
        }
        __syncthreads();
        float out[8][4];
#pragma unroll
        for (int t = 0; t < 8; ++t) { out[t][0] = b2.x; out[t][1] = b2.y; out[t][2] = b2.z; out[t][3] = b2.w; }
        tile_linear<TB_HID>(s_h, TB_HPITCH, s_w2, TB_DIM, 0, out);         // ref :25 fc2
#pragma unroll
        for (int t = 0; t < 8; ++t) {                                       // ref :86 x + mlp(...)
            const long tok = tok0 + ty + 8 * t;
            if (tok < total)
                *reinterpret_cast<float4*>(y + tok * TB_DIM + tx * 4) =
                    make_float4(out[t][0] + x1[t][0], out[t][1] + x1[t][1], out[t][2] + x1[t][2], out[t][3] + x1[t][3]);
        }
    }
}

}  // namespace m2t

using namespace m2t;

extern "C" {

size_t m2t_transblock_workspace_bytes(int B, int N, int dim) {
    if (B < 1 || N < 1 || dim != TB_DIM) return 0;
    return (size_t)B * N * (3 * TB_DIM + TB_DIM) * sizeof(float);      // QKV + attention output
}

int m2t_transblock_forward(const float* d_x, float* d_y, const float* const* d_params, int n_params, int B, int N,
                           int dim, int num_heads, void* d_workspace, void* stream) {
    if (!d_x || !d_y || !d_params || !d_workspace) { set_error("transblock: null pointer"); return M2T_E_ARG; }
    if (n_params != 12) { set_error("transblock: expected the 12 tensors of TransBlock.state_dict(), got %d", n_params); return M2T_E_ARG; }
    if (dim != TB_DIM || num_heads != TB_HEADS) {
        set_error("transblock: dim %d / num_heads %d: built for the constructor defaults 64 / 8", dim, num_heads);
        return M2T_E_UNSUPPORTED;
    }
    if (B < 1) { set_error("transblock: bad batch %d", B); return M2T_E_ARG; }
    if (N < 16) {      // ref :53: torch.split(q, N // 16) raises for a chunk length of 0
        set_error("transblock: N = %d < 16: the reference's chunk length N // 16 is 0", N);
        return M2T_E_ARG;
    }
    for (int i = 0; i < 12; ++i)
        if (!d_params[i]) { set_error("transblock: parameter %d is null", i); return M2T_E_ARG; }
    M2T_TRY(check_device());
    cudaStream_t s = (cudaStream_t)stream;
    TbParams P{d_params[0], d_params[1], d_params[2], d_params[3], d_params[4], d_params[5],
               d_params[6], d_params[7], d_params[8], d_params[9], d_params[10], d_params[11]};
    const long total = (long)B * N;
    float* qkv = static_cast<float*>(d_workspace);
    float* o = qkv + total * 3 * TB_DIM;
    const long tiles = (total + TB_TOK - 1) / TB_TOK;
    const int sms = device_sm_count();
    const int grid = (int)(tiles < 2L * sms ? tiles : 2L * sms);
    M2T_ENSURE_SMEM(tb_qkv_kernel, TBQ_SMEM);
    M2T_ENSURE_SMEM(tb_out_kernel, TBO_SMEM);
    tb_qkv_kernel<<<grid, TB_THREADS, TBQ_SMEM, s>>>(d_x, P, qkv, total, 1);
    M2T_LAUNCH_CHECK("tb_qkv_kernel");
    const int n = N / 16, nchunks = (N + n - 1) / n;
    const float scale = 1.0f / sqrtf((float)TB_HD);                      // head_dim ** -0.5 (ref :35)
    tb_attn_kernel<<<dim3((unsigned)(B * nchunks), TB_HEADS), TBA_THREADS, 0, s>>>(qkv, o, N, n, nchunks,
                                                                                 scale * 1.4426950408889634f);
    M2T_LAUNCH_CHECK("tb_attn_kernel");
    tb_out_kernel<<<grid, TB_THREADS, TBO_SMEM, s>>>(d_x, o, P, d_y, total, 0);
    M2T_LAUNCH_CHECK("tb_out_kernel");
    return M2T_OK;
}

// EffAttention.forward alone (ref util/rlutrans.py:47-66): proj(BDAttn(qkv(reduce(x)))).  d_params: reduce.weight,
// qkv.weight, proj.weight, proj.bias (EffAttention.state_dict() order); same workspace as the whole block.
int m2t_rlutrans_attention(const float* d_x, float* d_y, const float* const* d_params, int n_params, int B, int N, int dim,
                           int num_heads, void* d_workspace, void* stream) {
    if (!d_x || !d_y || !d_params || !d_workspace) { set_error("rlutrans attention: null pointer"); return M2T_E_ARG; }
    if (n_params != 4) { set_error("rlutrans attention: expected the 4 tensors of EffAttention.state_dict(), got %d", n_params); return M2T_E_ARG; }
    if (dim != TB_DIM || num_heads != TB_HEADS) { set_error("rlutrans attention: dim %d / num_heads %d: built for 64 / 8", dim, num_heads); return M2T_E_UNSUPPORTED; }
    if (B < 1 || N < 16) { set_error("rlutrans attention: B %d, N %d (the reference's chunk length N // 16 must be >= 1)", B, N); return M2T_E_ARG; }
    for (int i = 0; i < 4; ++i)
        if (!d_params[i]) { set_error("rlutrans attention: parameter %d is null", i); return M2T_E_ARG; }
    M2T_TRY(check_device());
    cudaStream_t s = (cudaStream_t)stream;
    TbParams P{};
    P.reduce_w = d_params[0]; P.qkv_w = d_params[1]; P.proj_w = d_params[2]; P.proj_b = d_params[3];
    const long total = (long)B * N;
    float* qkv = static_cast<float*>(d_workspace);
    float* o = qkv + total * 3 * TB_DIM;
    const long tiles = (total + TB_TOK - 1) / TB_TOK;
    const int sms = device_sm_count();
    const int grid = (int)(tiles < 2L * sms ? tiles : 2L * sms);
    M2T_ENSURE_SMEM(tb_qkv_kernel, TBQ_SMEM);
    M2T_ENSURE_SMEM(tb_out_kernel, TBO_SMEM);
    tb_qkv_kernel<<<grid, TB_THREADS, TBQ_SMEM, s>>>(d_x, P, qkv, total, 0);
    M2T_LAUNCH_CHECK("tb_qkv_kernel");
    const int n = N / 16, nchunks = (N + n - 1) / n;
    tb_attn_kernel<<<dim3((unsigned)(B * nchunks), TB_HEADS), TBA_THREADS, 0, s>>>(qkv, o, N, n, nchunks,
                                                                                 (1.0f / sqrtf((float)TB_HD)) * 1.4426950408889634f);
    M2T_LAUNCH_CHECK("tb_attn_kernel");
    tb_out_kernel<<<grid, TB_THREADS, TBO_SMEM, s>>>(d_x, o, P, d_y, total, 1);
    M2T_LAUNCH_CHECK("tb_out_kernel");
    return M2T_OK;
}

// Mlp.forward alone (ref util/rlutrans.py:21-27): fc2(ReLU(fc1(x))) on `tokens` rows of 64 channels.
// d_params: fc1.weight [16][64], fc1.bias, fc2.weight [64][16], fc2.bias (Mlp.state_dict() order).
int m2t_rlutrans_mlp(const float* d_x, float* d_y, const float* const* d_params, int n_params, long tokens, int dim,
                     int hidden, void* stream) {
    if (!d_x || !d_y || !d_params) { set_error("rlutrans mlp: null pointer"); return M2T_E_ARG; }
    if (n_params != 4) { set_error("rlutrans mlp: expected the 4 tensors of Mlp.state_dict(), got %d", n_params); return M2T_E_ARG; }
    if (dim != TB_DIM || hidden != TB_HID) { set_error("rlutrans mlp: %d -> %d -> %d: built for 64 -> 16 -> 64", dim, hidden, dim); return M2T_E_UNSUPPORTED; }
    if (tokens < 1) { set_error("rlutrans mlp: %ld tokens", tokens); return M2T_E_ARG; }
    for (int i = 0; i < 4; ++i)
        if (!d_params[i]) { set_error("rlutrans mlp: parameter %d is null", i); return M2T_E_ARG; }
    M2T_TRY(check_device());
    TbParams P{};
    P.fc1_w = d_params[0]; P.fc1_b = d_params[1]; P.fc2_w = d_params[2]; P.fc2_b = d_params[3];
    const long tiles = (tokens + TB_TOK - 1) / TB_TOK;
    const int sms = device_sm_count();
    const int grid = (int)(tiles < 2L * sms ? tiles : 2L * sms);
    M2T_ENSURE_SMEM(tb_out_kernel, TBO_SMEM);
    tb_out_kernel<<<grid, TB_THREADS, TBO_SMEM, (cudaStream_t)stream>>>(d_x, d_x, P, d_y, tokens, 2);
    M2T_LAUNCH_CHECK("tb_out_kernel");
    return M2T_OK;
}

}  // extern "C"
