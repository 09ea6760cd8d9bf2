// Window attention of the MedCLIP image tower (Swin-T; modeling_swin.py:430-487, :598-640): 49 tokens x 32 dims per
// (window, head).  Two warps per (image, window, head), three heads per CTA; bf16 mma.sync.m16n8k16 with fp32 accumulators:
//   S = Q K^T (64 x 56 padded, 16 query rows at a time) -> + relative-position bias, - 100 across the shift regions ->
//   softmax on the accumulator fragments -> P re-used in registers as the A operand -> O = P V -> bf16
// The tiles are 49 x 32: far below the 128-row tcgen05 tile (two windows per tile would waste three quarters of S), and
// attention is 4 % of the tower's FLOPs; the Linear layers around it are the tcgen05 kernels (lin_umma.cu).
// The cyclic shift of the odd layers, the window partition and their inverses are index arithmetic on the token-major
// tensors; the region mask (:556-582) is recomputed from region ids.
#include <cuda_bf16.h>

#include "clip.cuh"

namespace m2t {

namespace {

constexpr int CA_WARPS = 3;                       // heads per CTA (3, 6, 12, 24 heads per stage)
constexpr int CA_ROWS = 64;                       // 49 tokens padded to 4 m16 tiles; keys to 7 n8 tiles / 4 k16 steps
constexpr int CA_MAT = CA_ROWS * 64;              // bytes per operand matrix: 64 rows x 32 bf16

// 64-byte rows, 16-byte chunk c of row r stored at chunk c ^ ((r >> 1) & 3): ldmatrix reads 8 rows conflict-free
__device__ __forceinline__ uint32_t ca_off(int r, int c) { return (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float c[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(CA_WARPS * 64, 6)
clip_attn_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, const float* __restrict__ rpb,
                 int h, int w, int C, int shift) {
    __shared__ __align__(128) uint8_t sm[CA_WARPS][3][CA_MAT];
    __shared__ int sRow[CA_ROWS];
    __shared__ int sId[CA_ROWS];
    // two warps per (window, head): each owns two of the four 16-row query tiles and loads half of the operand rows; the
    // kernel is latency-bound (ncu: issue slots 27 % busy at 18 warps per SM), so the same shared memory now feeds 36 warps
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 6, half = (tid >> 5) & 1;
    const int head = blockIdx.y * CA_WARPS + warp, nwx = w / CL_WIN, nW = (h / CL_WIN) * nwx;
    const int bimg = blockIdx.x / nW, wi = blockIdx.x - bimg * nW, wy = wi / nwx, wx = wi - wy * nwx;
    if (tid < CA_ROWS) {
        int row = 0, id = 0;
        if (tid < CL_WT) {
            const int y = wy * CL_WIN + tid / CL_WIN, x = wx * CL_WIN + tid % CL_WIN;     // coordinates in the shifted frame
            int gy = y + shift, gx = x + shift;
            if (gy >= h) gy -= h;
            if (gx >= w) gx -= w;
            row = (bimg * h + gy) * w + gx;
            id = (y < h - CL_WIN ? 0 : (y < h - shift ? 1 : 2)) * 3 + (x < w - CL_WIN ? 0 : (x < w - shift ? 1 : 2));
        }
        sRow[tid] = row;
        sId[tid] = id;
    }
    pdl_wait();
    __syncthreads();
    uint8_t* mq = sm[warp][0];
    // q | k | v rows of this head: 49 tokens x 4 chunks of 16 bytes each, rows 49..63 zero
    // (cp.async: all 24 chunks of a lane are in flight at once, no staging registers)
    const uint32_t sq = (uint32_t)__cvta_generic_to_shared(mq), sk = sq + CA_MAT, sv = sq + 2 * CA_MAT;
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        const __nv_bfloat16* src = qkv + m * C + head * CL_HD + (lane & 3) * 8;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int tok = (half * 4 + p) * 8 + (lane >> 2);
            const uint32_t dst = sq + m * CA_MAT + ca_off(tok, lane & 3);
            if (tok < CL_WT)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src + (long)sRow[tok] * 3 * C) : "memory");
            else
                asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" :: "r"(dst), "r"(0u) : "memory");
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    asm volatile("bar.sync %0, 64;" :: "r"(1 + warp) : "memory");      // the pair's rows have landed
    const int g = lane >> 2, t = lane & 3;
    const float* bh = rpb + (long)head * CL_WT * 56;
    constexpr float kScale = 0.17677669529663687f, kLog2e = 1.4426950408889634f;

#pragma unroll 1
    for (int mt = 2 * half; mt < 2 * half + 2; ++mt) {
        uint32_t qa[2][4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
            ldsm_x4(sq + ca_off(mt * 16 + (lane & 15), ks * 2 + (lane >> 4)), qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
        float s[7][4];
#pragma unroll
        for (int nt = 0; nt < 7; ++nt) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4(sk + ca_off(nt * 8 + (lane & 7), lane >> 3), b0, b1, b2, b3);
            s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
            mma_bf16(s[nt], qa[0][0], qa[0][1], qa[0][2], qa[0][3], b0, b1);
            mma_bf16(s[nt], qa[1][0], qa[1][1], qa[1][2], qa[1][3], b2, b3);
        }
        const int r0 = mt * 16 + g, r1 = r0 + 8;
        const bool v0 = r0 < CL_WT, v1 = r1 < CL_WT;
        const int id0 = sId[r0], id1 = sId[r1];
        float mx0 = -3.0e38f, mx1 = -3.0e38f;
#pragma unroll
        for (int nt = 0; nt < 7; ++nt) {
            const int j = nt * 8 + 2 * t;
            const float2 ba = v0 ? __ldg(reinterpret_cast<const float2*>(bh + r0 * 56 + j)) : make_float2(0.f, 0.f);
            const float2 bb = v1 ? __ldg(reinterpret_cast<const float2*>(bh + r1 * 56 + j)) : make_float2(0.f, 0.f);
            s[nt][0] = fmaf(s[nt][0], kScale, ba.x); s[nt][1] = fmaf(s[nt][1], kScale, ba.y);
            s[nt][2] = fmaf(s[nt][2], kScale, bb.x); s[nt][3] = fmaf(s[nt][3], kScale, bb.y);
            if (shift) {
                const int c0 = sId[j], c1 = sId[j + 1];
                if (c0 != id0) s[nt][0] -= 100.f;
                if (c1 != id0) s[nt][1] -= 100.f;
                if (c0 != id1) s[nt][2] -= 100.f;
                if (c1 != id1) s[nt][3] -= 100.f;
            }
            mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
            mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float o0 = -mx0 * kLog2e, o1 = -mx1 * kLog2e;
        float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < 7; ++nt) {          // the padded key columns carry a bias of -1e30: exp2 gives 0
            s[nt][0] = ex2(fmaf(s[nt][0], kLog2e, o0)); s[nt][1] = ex2(fmaf(s[nt][1], kLog2e, o0));
            s[nt][2] = ex2(fmaf(s[nt][2], kLog2e, o1)); s[nt][3] = ex2(fmaf(s[nt][3], kLog2e, o1));
            sum0 += s[nt][0] + s[nt][1];
            sum1 += s[nt][2] + s[nt][3];
        }
        sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
        sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
        float o[4][4];
#pragma unroll
        for (int dn = 0; dn < 4; ++dn) o[dn][0] = o[dn][1] = o[dn][2] = o[dn][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {          // 16 keys per step; the accumulator fragments of S are the A fragments
            const uint32_t a0 = pack2(s[2 * kk][0], s[2 * kk][1]), a1 = pack2(s[2 * kk][2], s[2 * kk][3]);
            uint32_t a2 = 0u, a3 = 0u;
            if (kk < 3) { a2 = pack2(s[2 * kk + 1][0], s[2 * kk + 1][1]); a3 = pack2(s[2 * kk + 1][2], s[2 * kk + 1][3]); }
#pragma unroll
            for (int dp = 0; dp < 2; ++dp) {
                uint32_t b0, b1, b2, b3;
                ldsm_x4_t(sv + ca_off(kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, dp * 2 + (lane >> 4)), b0, b1, b2, b3);
                mma_bf16(o[2 * dp], a0, a1, a2, a3, b0, b1);
                mma_bf16(o[2 * dp + 1], a0, a1, a2, a3, b2, b3);
            }
        }
        const float i0 = 1.f / sum0, i1 = 1.f / sum1;
        __syncwarp();                              // every lane has its Q fragments of this m tile: the rows become O
#pragma unroll
        for (int dn = 0; dn < 4; ++dn) {
            *reinterpret_cast<uint32_t*>(mq + ca_off(r0, dn) + 4 * t) = pack2(o[dn][0] * i0, o[dn][1] * i0);
            *reinterpret_cast<uint32_t*>(mq + ca_off(r1, dn) + 4 * t) = pack2(o[dn][2] * i1, o[dn][3] * i1);
        }
    }
    __syncwarp();
    for (int idx = lane; idx < 32 * 4; idx += 32) {          // each warp stores the 32 rows its two tiles produced
        const int tok = half * 32 + (idx >> 2), c = idx & 3;
        if (tok < CL_WT)
            *reinterpret_cast<uint4*>(out + (long)sRow[tok] * C + head * CL_HD + c * 8) =
                *reinterpret_cast<const uint4*>(mq + ca_off(tok, c));
    }
}

}  // namespace

// qkv bf16 [tokens][3C], out bf16 [tokens][C], rpb fp32 [heads][49][56] (columns 49..55 = -1e30)
int launch_clip_attn(const void* qkv, void* out, const float* rpb, int B, int h, int w, int C, int heads, int shift,
                     cudaStream_t s) {
    if (heads % CA_WARPS || heads * CL_HD != C || h % CL_WIN || w % CL_WIN) {
        set_error("clip attention: grid %d x %d, C %d, heads %d", h, w, C, heads);
        return M2T_E_UNSUPPORTED;
    }
    M2T_CUDA(launch_pdl(clip_attn_kernel, dim3((unsigned)(B * (h / CL_WIN) * (w / CL_WIN)), (unsigned)(heads / CA_WARPS)),
                        dim3(CA_WARPS * 64), 0, s, static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), rpb,
                        h, w, C, shift));
    return M2T_OK;
}

}  // namespace m2t
