// CUDA-core halo attention (M2T_VAR_SIMT_ATTN variant) -- ref M2Trans_network.py:310-332 with block 8,
// halo 1, one head.  One CTA per 8x8 query block; keys/values are the 10x10 neighbourhood, ZERO outside
// the frame (F.unfold padding, ref :313-317).  The relative-position terms are added to K in the
// reference (ref :322-325), also at the zero-padded keys; here q.(k + rel) is evaluated as
// q.k + q[:C/2].rel_h[row] + q[C/2:].rel_w[col], which is the same sum.  q arrives pre-scaled by C^-1/2
// (folded into the packed qkv weight; exact, the scale is a power of two).
// fp16 operands, fp32 accumulation, fp32 softmax.
#include "common.cuh"

namespace m2t {

constexpr int AT_THREADS = 128;
constexpr int S_LD = 101;
constexpr int AB_LD = 21;

template <int C>
struct AttnSmem {
    static constexpr int LD = C + 8;                       // halves per row (16 B pad)
    static constexpr size_t q_off = 0;
    static constexpr size_t k_off = q_off + (size_t)64 * LD * 2;
    static constexpr size_t v_off = k_off + (size_t)NKEY * LD * 2;
    static constexpr size_t rel_off = v_off + (size_t)NKEY * LD * 2;
    static constexpr size_t ab_off = rel_off + (size_t)20 * (C / 2) * 4;
    static constexpr size_t s_off = ab_off + (size_t)64 * AB_LD * 4;
    static constexpr size_t bytes = s_off + (size_t)64 * S_LD * 4;
};

__device__ __forceinline__ void h8_to_f(const uint4& u, float f[8]) {
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 v = __half22float2(h[i]);
        f[2 * i] = v.x; f[2 * i + 1] = v.y;
    }
}

template <int C>
__global__ void __launch_bounds__(AT_THREADS)
attn_simt_kernel(const __half* __restrict__ QKV, const float* __restrict__ relf, __half* __restrict__ O, int h,
                 int w) {
    using SM = AttnSmem<C>;
    constexpr int LD = SM::LD;
    constexpr int CH = C / 8;                              // 16-byte chunks per row
    extern __shared__ __align__(16) uint8_t smem[];
    __half* Qs = reinterpret_cast<__half*>(smem + SM::q_off);
    __half* Ks = reinterpret_cast<__half*>(smem + SM::k_off);
    __half* Vs = reinterpret_cast<__half*>(smem + SM::v_off);
    float* Rl = reinterpret_cast<float*>(smem + SM::rel_off);
    float* AB = reinterpret_cast<float*>(smem + SM::ab_off);
    float* Sm = reinterpret_cast<float*>(smem + SM::s_off);

    const int t = threadIdx.x;
    const int nwx = w / BLK;
    const int bx = blockIdx.x % nwx, by = blockIdx.x / nwx, b = blockIdx.y;
    const __half* base = QKV + (long)b * h * w * 3 * C;

    // ---- stage Q (64 x C), K and V (100 x C, zero outside the frame), rel tables
    for (int idx = t; idx < NKEY * CH; idx += AT_THREADS) {
        const int j = idx / CH, ch = idx - j * CH;
        const int r = j / WIN, s = j - r * WIN;
        const int y = by * BLK - 1 + r, x = bx * BLK - 1 + s;
        uint4 kv = make_uint4(0, 0, 0, 0), vv = kv;
        if (y >= 0 && y < h && x >= 0 && x < w) {
            const __half* p = base + ((long)y * w + x) * 3 * C + ch * 8;
            kv = *reinterpret_cast<const uint4*>(p + C);
            vv = *reinterpret_cast<const uint4*>(p + 2 * C);
            if (r >= 1 && r <= BLK && s >= 1 && s <= BLK)
                *reinterpret_cast<uint4*>(&Qs[((r - 1) * BLK + (s - 1)) * LD + ch * 8]) =
                    *reinterpret_cast<const uint4*>(p);
        }
        *reinterpret_cast<uint4*>(&Ks[j * LD + ch * 8]) = kv;
        *reinterpret_cast<uint4*>(&Vs[j * LD + ch * 8]) = vv;
    }
    for (int idx = t; idx < 20 * (C / 2); idx += AT_THREADS) Rl[idx] = relf[idx];
    __syncthreads();

    // ---- AB[i][rr] = q_i[:C/2].rel_h[rr] (rr<10) ; q_i[C/2:].rel_w[rr-10] (rr>=10)
    for (int idx = t; idx < 64 * 20; idx += AT_THREADS) {
        const int i = idx / 20, rr = idx - i * 20;
        const __half* qp = &Qs[i * LD + (rr < 10 ? 0 : C / 2)];
        const float* rp = &Rl[rr * (C / 2)];
        float acc = 0.f;
#pragma unroll 4
        for (int c = 0; c < C / 2; c += 2) {
            const float2 qv = __half22float2(*reinterpret_cast<const __half2*>(qp + c));
            acc = fmaf(qv.x, rp[c], acc);
            acc = fmaf(qv.y, rp[c + 1], acc);
        }
        AB[i * AB_LD + rr] = acc;
    }
    __syncthreads();

    // ---- S = Q K^T + rel terms
    {
        const int ti = t & 15, tj = t >> 4;
        float acc[4][13];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int bb = 0; bb < 13; ++bb) acc[a][bb] = 0.f;
        for (int ch = 0; ch < CH; ++ch) {
            float q[4][8];
#pragma unroll
            for (int a = 0; a < 4; ++a)
                h8_to_f(*reinterpret_cast<const uint4*>(&Qs[(ti + 16 * a) * LD + ch * 8]), q[a]);
#pragma unroll
            for (int bb = 0; bb < 13; ++bb) {
                const int j = tj + 8 * bb;
                if (j < NKEY) {
                    float kf[8];
                    h8_to_f(*reinterpret_cast<const uint4*>(&Ks[j * LD + ch * 8]), kf);
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int e = 0; e < 8; ++e) acc[a][bb] = fmaf(q[a][e], kf[e], acc[a][bb]);
                }
            }
        }
#pragma unroll
        for (int bb = 0; bb < 13; ++bb) {
            const int j = tj + 8 * bb;
            if (j < NKEY) {
                const int r = j / WIN, s = j - r * WIN;
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const int i = ti + 16 * a;
                    Sm[i * S_LD + j] = acc[a][bb] + AB[i * AB_LD + r] + AB[i * AB_LD + 10 + s];
                }
            }
        }
    }
    __syncthreads();

    // ---- row softmax (one warp per row, 16 rows per warp)
    {
        const int lane = t & 31, wid = t >> 5;
        for (int i = wid * 16; i < wid * 16 + 16; ++i) {
            float v[4], m = -INFINITY;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = lane + 32 * u;
                v[u] = j < NKEY ? Sm[i * S_LD + j] : -INFINITY;
                m = fmaxf(m, v[u]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            float sum = 0.f;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                v[u] = (lane + 32 * u) < NKEY ? expf(v[u] - m) : 0.f;
                sum += v[u];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float inv = 1.f / sum;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = lane + 32 * u;
                if (j < NKEY) Sm[i * S_LD + j] = v[u] * inv;
            }
        }
    }
    __syncthreads();

    // ---- O = P V, written back in window-reversed order (ref :332)
    {
        const int ti = t & 15, tc = t >> 4;
        for (int cc = tc; cc < CH; cc += 8) {
            float acc[4][8];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[a][e] = 0.f;
            for (int j = 0; j < NKEY; ++j) {
                float vf[8];
                h8_to_f(*reinterpret_cast<const uint4*>(&Vs[j * LD + cc * 8]), vf);
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const float p = Sm[(ti + 16 * a) * S_LD + j];
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[a][e] = fmaf(p, vf[e], acc[a][e]);
                }
            }
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int i = ti + 16 * a;
                const int y = by * BLK + (i >> 3), x = bx * BLK + (i & 7);
                __half2 hv[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) hv[e] = __floats2half2_rn(acc[a][2 * e], acc[a][2 * e + 1]);
                *reinterpret_cast<uint4*>(O + (((long)b * h + y) * w + x) * C + cc * 8) =
                    *reinterpret_cast<const uint4*>(hv);
            }
        }
    }
}

template <int C>
static int launch_attn_c(const __half* QKV, const float* relf, __half* O, int B, int h, int w, cudaStream_t s) {
    using SM = AttnSmem<C>;
    M2T_ENSURE_SMEM(attn_simt_kernel<C>, SM::bytes);
    dim3 grid((h / BLK) * (w / BLK), B);
    attn_simt_kernel<C><<<grid, AT_THREADS, SM::bytes, s>>>(QKV, relf, O, h, w);
    M2T_LAUNCH_CHECK("attn_simt_kernel");
    return M2T_OK;
}

int launch_attn_simt(int C, const __half* QKV, const float* relf, __half* O, int B, int h, int w, cudaStream_t s) {
    if (h % BLK || w % BLK) { set_error("attn: %dx%d is not a multiple of the 8x8 block", h, w); return M2T_E_ARG; }
    if (C == 16) return launch_attn_c<16>(QKV, relf, O, B, h, w, s);
    if (C == 64) return launch_attn_c<64>(QKV, relf, O, B, h, w, s);
    if (C == 256) return launch_attn_c<256>(QKV, relf, O, B, h, w, s);
    set_error("attn: unsupported channel count %d", C);
    return M2T_E_UNSUPPORTED;
}

}  // namespace m2t
