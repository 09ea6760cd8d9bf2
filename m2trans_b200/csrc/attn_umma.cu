// tcgen05 halo attention (ref M2Trans_network.py:310-332; block 8, halo 1, one head) for C in {16, 64, 256}.
//
// One work item = a PAIR of 8x8 query windows, stacked into one M = 128 tile (rows 0-63 window A, 64-127
// window B), so every TMEM lane / epilogue thread owns one query row.
//   S  = [Qa;Qb] . [Ka;Kb]^T         M=128, N=208 (2 x 100 keys padded to 104), K = C     (block-diagonal use:
//                                     row m only reads the 100 columns of its own window)
//   S' = [Qa;Qb] . Rel^T              N = 32 extra columns: q[:C/2].rel_h[r] (cols 208..217) and
//                                     q[C/2:].rel_w[c] (cols 218..227): the reference adds rel to K, also at the
//                                     zero-padded keys (ref :322-325); q.(k+rel) = q.k + q.rel
//   P  = exp2((S + rel terms - max) * log2e), unnormalised, fp16, written K-major/128B-swizzled to smem
//   O  = P . [Va;Vb]                  M=128, N = C, K = 208 keys; rows scaled by 1/sum in the epilogue
// Keys/values outside the frame are TMA out-of-bounds zero fill = F.unfold's zero padding (ref :313-317);
// window partition / reverse (ref :310, :332) are TMA box coordinates and the store address.
// Operands stream through one ring in 64-channel blocks (Q 128 rows + K 208 rows, then V 208 rows), so the
// C = 256 case (288 KB of Q/K/V per pair) fits and loads overlap the previous pair's softmax / PV.
// Warp roles (192 threads): warps 0-3 softmax + epilogue (thread = query row), warp 4 TMA, warp 5 MMA issue.
//
// FUSE = true additionally folds the CFTM branch glue (ref :139-161) into the epilogue, see AttnFuse.
#include "common.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace m2t {

constexpr int WR = 104;                                       // key rows per window (100 + 4 zero rows)
constexpr int NKP = 2 * WR;                                   // 208 key rows / S columns per pair

template <int C>
struct AtCfg {
    static constexpr int CB = C < 64 ? C : 64;               // channels per streamed block
    static constexpr int NBLK = C / CB;
    static constexpr uint32_t ROWB = CB * 2;                  // bytes per row (32 or 128)
    static constexpr uint64_t LAYOUT = CB == 64 ? UMMA_LAYOUT_SW128 : UMMA_LAYOUT_SW32;
    static constexpr int TMA_SWZ = CB == 64 ? 3 : 1;
    static constexpr uint32_t SBO = 8 * ROWB;
    static constexpr uint32_t QB = 128 * ROWB;                // Q block (two windows x 64 rows)
    static constexpr uint32_t KVB = NKP * ROWB;               // K or V block
    static constexpr uint32_t WIN_B = WR * ROWB;              // offset of window B's keys
    static constexpr uint32_t STAGE = (QB + KVB + 1023) / 1024 * 1024;
    static constexpr int STAGES = C == 16 ? 4 : 3;
    static constexpr uint32_t REL_BLOCK = 32 * ROWB;
    static constexpr uint32_t OFF_P = STAGES * STAGE;
    static constexpr uint32_t OFF_REL = OFF_P + 4 * 16384;
    static constexpr uint32_t OFF_BAR = OFF_REL + (NBLK * REL_BLOCK + 1023) / 1024 * 1024;
    static constexpr uint32_t SMEM = 1024 + OFF_BAR + 256;
    static constexpr uint32_t TX_QK = 2 * 64 * ROWB + 2 * 100 * ROWB;
    static constexpr uint32_t TX_V = 2 * 100 * ROWB;
    // TMEM columns: S [0,208) | rel [208,240) | O.  C = 16 fits 256 columns, so two CTAs can share an SM.
    static constexpr uint32_t TM_REL = NKP;
    static constexpr uint32_t TM_O = C == 16 ? 240 : 256;
    static constexpr uint32_t TM_COLS = C == 16 ? 256 : 512;
    static constexpr int MIN_CTAS = C == 16 ? 2 : 1;
};

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct WinCoord { int b, y, x; };
__device__ __forceinline__ WinCoord win_coord(int wi, int nwx, int per_img) {
    WinCoord c;
    c.b = wi / per_img;
    const int r = wi - c.b * per_img;
    c.y = (r / nwx) * BLK;
    c.x = (r % nwx) * BLK;
    return c;
}

template <int G>
struct GluePre {       // prefetched operands of G sub-pixels of the fused glue
    uint4 tk[2 * G];
    float4 xv[4 * G];
};

template <int C, bool FUSE>
__global__ void __launch_bounds__(192, AtCfg<C>::MIN_CTAS)
attn_umma_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapKV,
                 const __grid_constant__ CUtensorMap mapR, __half* __restrict__ O, int h, int w, int nwin,
                 const AttnFuse fz) {
    using CF = AtCfg<C>;
    constexpr int CB = CF::CB, NBLK = CF::NBLK, STAGES = CF::STAGES;
    constexpr uint32_t ROWB = CF::ROWB;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + CF::OFF_BAR);
    uint64_t* full = bars;                         // [STAGES]
    uint64_t* empty = bars + STAGES;               // [STAGES]
    uint64_t* rfull = bars + 2 * STAGES;           // rel tables landed
    uint64_t* s_full = bars + 2 * STAGES + 1;      // S complete in TMEM
    uint64_t* p_ready = bars + 2 * STAGES + 2;     // P written to smem, S consumed
    uint64_t* o_full = bars + 2 * STAGES + 3;      // O complete in TMEM
    uint64_t* o_empty = bars + 2 * STAGES + 4;     // O consumed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 5);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nwx = w / BLK, per_img = (h / BLK) * nwx;
    const int npairs = (nwin + 1) / 2;

    // zero P (its off-diagonal / padding columns stay zero for ever) and the 4 padding key rows per window
    for (uint32_t i = tid * 16; i < 4 * 16384; i += 192 * 16) *reinterpret_cast<uint4*>(sm + CF::OFF_P + i) = make_uint4(0, 0, 0, 0);
    for (int s = 0; s < STAGES; ++s)
        for (int half = 0; half < 2; ++half) {
            uint8_t* pad = sm + s * CF::STAGE + CF::QB + half * CF::WIN_B + 100 * ROWB;
            for (uint32_t i = tid * 16; i < (WR - 100) * ROWB; i += 192 * 16) *reinterpret_cast<uint4*>(pad + i) = make_uint4(0, 0, 0, 0);
        }
    fence_proxy_async();
    if (warp == 5) tmem_alloc(tmem_slot, CF::TM_COLS);
    if (tid == 128) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(rfull, 1);
        mbar_init(s_full, 1);
        mbar_init(p_ready, 4);
        mbar_init(o_full, 1);
        mbar_init(o_empty, 4);
        mbar_fence_init();
        tma_prefetch_desc(&mapQ);
        tma_prefetch_desc(&mapKV);
        tma_prefetch_desc(&mapR);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t TM_S = 0, TM_REL = CF::TM_REL, TM_O = CF::TM_O;

    if (warp == 4) {
        if (lane == 0) {
            mbar_expect_tx(rfull, NBLK * CF::REL_BLOCK);
            for (int kb = 0; kb < NBLK; ++kb) tma_load_2d(sm + CF::OFF_REL + kb * CF::REL_BLOCK, &mapR, rfull, kb * CB, 0);
            uint32_t g = 0;
            for (int p = blockIdx.x; p < npairs; p += gridDim.x) {
                const int wa = 2 * p, wb = (2 * p + 1 < nwin) ? 2 * p + 1 : 2 * p;
                const WinCoord a = win_coord(wa, nwx, per_img), b = win_coord(wb, nwx, per_img);
                for (int kb = 0; kb < NBLK; ++kb, ++g) {
                    const uint32_t s = g % STAGES, ph = (g / STAGES) & 1;
                    uint8_t* st = sm + s * CF::STAGE;
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_expect_tx(&full[s], CF::TX_QK);
                    tma_load_4d(st, &mapQ, &full[s], kb * CB, a.x, a.y, a.b);
                    tma_load_4d(st + 64 * ROWB, &mapQ, &full[s], kb * CB, b.x, b.y, b.b);
                    tma_load_4d(st + CF::QB, &mapKV, &full[s], C + kb * CB, a.x - 1, a.y - 1, a.b);
                    tma_load_4d(st + CF::QB + CF::WIN_B, &mapKV, &full[s], C + kb * CB, b.x - 1, b.y - 1, b.b);
                }
                for (int nb = 0; nb < NBLK; ++nb, ++g) {
                    const uint32_t s = g % STAGES, ph = (g / STAGES) & 1;
                    uint8_t* st = sm + s * CF::STAGE;
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_expect_tx(&full[s], CF::TX_V);
                    tma_load_4d(st + CF::QB, &mapKV, &full[s], 2 * C + nb * CB, a.x - 1, a.y - 1, a.b);
                    tma_load_4d(st + CF::QB + CF::WIN_B, &mapKV, &full[s], 2 * C + nb * CB, b.x - 1, b.y - 1, b.b);
                }
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            constexpr uint32_t id_s = umma_idesc_f16(128, NKP), id_r = umma_idesc_f16(128, 32);
            constexpr uint32_t id_o = umma_idesc_f16(128, CB, 0, 1);
            mbar_wait(rfull, 0);
            uint32_t g = 0, it = 0;
            for (int p = blockIdx.x; p < npairs; p += gridDim.x, ++it) {
                // S and the rel columns
                for (int kb = 0; kb < NBLK; ++kb, ++g) {
                    const uint32_t s = g % STAGES, ph = (g / STAGES) & 1;
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t st = base + s * CF::STAGE;
#pragma unroll
                    for (int k = 0; k < CB / 16; ++k) {
                        const uint64_t dq = umma_smem_desc(st + k * 32, 16, CF::SBO, CF::LAYOUT);
                        const uint64_t dk = umma_smem_desc(st + CF::QB + k * 32, 16, CF::SBO, CF::LAYOUT);
                        const uint64_t dr = umma_smem_desc(base + CF::OFF_REL + kb * CF::REL_BLOCK + k * 32, 16, CF::SBO, CF::LAYOUT);
                        const uint32_t accum = (kb | k) ? 1u : 0u;
                        umma_f16_ss(tmem_base + TM_S, dq, dk, id_s, accum);
                        umma_f16_ss(tmem_base + TM_REL, dq, dr, id_r, accum);
                    }
                    umma_commit(&empty[s]);
                }
                umma_commit(s_full);
                // O = P . V
                mbar_wait(p_ready, it & 1);
                mbar_wait(o_empty, (it & 1) ^ 1);
                tc_fence_after();
                for (int nb = 0; nb < NBLK; ++nb, ++g) {
                    const uint32_t s = g % STAGES, ph = (g / STAGES) & 1;
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t vb = base + s * CF::STAGE + CF::QB;
#pragma unroll
                    for (int k = 0; k < NKP / 16; ++k) {
                        const uint64_t dp = umma_smem_desc(base + CF::OFF_P + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024, UMMA_LAYOUT_SW128);
                        const uint64_t dv = umma_smem_desc(vb + k * 16 * ROWB, 16, CF::SBO, CF::LAYOUT);
                        umma_f16_ss(tmem_base + TM_O + nb * CB, dp, dv, id_o, k ? 1u : 0u);
                    }
                    umma_commit(&empty[s]);
                }
                umma_commit(o_full);
            }
        }
    } else {
        const int m = tid, half = m >> 6, qi = m & 63;
        const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
        uint8_t* prow = sm + CF::OFF_P + m * 128;
        uint32_t it = 0;
        for (int p = blockIdx.x; p < npairs; p += gridDim.x, ++it) {
            mbar_wait(s_full, it & 1);
            tc_fence_after();
            float sv[104];
            uint32_t ab[24];
            {
                uint32_t* su = reinterpret_cast<uint32_t*>(sv);
                const uint32_t t0 = tmem_base + lane_sel + TM_S + half * WR;
                tmem_ld32(t0, su);
                tmem_ld32(t0 + 32, su + 32);
                tmem_ld32(t0 + 64, su + 64);
                tmem_ld8(t0 + 96, su + 96);
                tmem_ld16(tmem_base + lane_sel + TM_REL, ab);
                tmem_ld8(tmem_base + lane_sel + TM_REL + 16, ab + 16);
                tmem_ld_wait();
            }
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < NKEY; ++j) {
                sv[j] += __uint_as_float(ab[j / WIN]) + __uint_as_float(ab[10 + j % WIN]);
                mx = fmaxf(mx, sv[j]);
            }
            const float mxl = mx * 1.4426950408889634f;
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < NKEY; ++j) {
                sv[j] = fast_exp2(fmaf(sv[j], 1.4426950408889634f, -mxl));
                sum += sv[j];
            }
#pragma unroll
            for (int j = NKEY; j < 104; ++j) sv[j] = 0.f;
            // P row: 13 chunks of 8 keys starting at key column half*104
#pragma unroll
            for (int q = 0; q < 13; ++q) {
                const int cg = half * 13 + q;                     // 16-byte chunk index along the 256-key row
                uint4 u;
                uint32_t* pu = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const __half2 hv = __floats2half2_rn(sv[q * 8 + 2 * e], sv[q * 8 + 2 * e + 1]);
                    pu[e] = *reinterpret_cast<const uint32_t*>(&hv);
                }
                *reinterpret_cast<uint4*>(prow + (cg >> 3) * 16384 + (((cg & 7) ^ (m & 7)) << 4)) = u;
            }
            const float inv = 1.f / sum;
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_ready);

            const int wi = 2 * p + half;
            const bool valid = wi < nwin;
            const WinCoord wc = win_coord(valid ? wi : 2 * p, nwx, per_img);
            if constexpr (FUSE) {
                // Fused branch glue (ref :139-161).  With the Haar transforms folded into the qkv weights the
                // accumulator row IS IWT^L(attention) in space-to-depth order: column s*16+k belongs to pixel
                // s = dy*2^L + dx of this level pixel's 2^L x 2^L block.
                //   y_k = O' + t_k              -> Y[..., 16k..16k+15]                  (ref :139,:145,:153,:161)
                //   t_{k+1} = (n_{k+1} + y_k)/2 -> Tnext, next level's space-to-depth order (ref :141,:147,:155)
                // Operand loads are issued one group ahead (the first group before the PV MMAs finish).
                constexpr int LV = C == 16 ? 0 : (C == 64 ? 1 : 2);
                constexpr int S = 1 << LV, NSUB = S * S;
                constexpr int G = NSUB >= 2 ? 2 : 1, NG = NSUB / G;
                const int ly = wc.y + (qi >> 3), lx = wc.x + (qi & 7);
                const __half* trow = fz.T + (((long)wc.b * h + ly) * w + lx) * C;
                const int br = fz.branch;
                const bool has_next = fz.Tnext != nullptr;
                const int lvn = br == 0 ? 1 : 2, Sn = 1 << lvn, Cn = NB * Sn * Sn;
                float mu[NB], rs[NB];
                if (has_next) {
#pragma unroll
                    for (int e = 0; e < NB; ++e) {
                        const float2 mr = __ldg(&fz.munorm[wc.b * NF + NB * (br + 1) + e]);
                        mu[e] = mr.x; rs[e] = mr.y;
                    }
                }
                auto pix_of = [&](int s) -> long {
                    return ((long)wc.b * fz.Hp + (ly * S + s / S)) * fz.Wp + (lx * S + s % S);
                };
                auto load_group = [&](int g0, GluePre<G>& pr) {
#pragma unroll
                    for (int j = 0; j < G; ++j) {
                        const int s = g0 * G + j;
                        pr.tk[2 * j] = *reinterpret_cast<const uint4*>(trow + s * NB);
                        pr.tk[2 * j + 1] = *reinterpret_cast<const uint4*>(trow + s * NB + 8);
                        if (has_next) {
                            const float4* xp = reinterpret_cast<const float4*>(fz.X + pix_of(s) * NF + NB * (br + 1));
#pragma unroll
                            for (int v = 0; v < 4; ++v) pr.xv[4 * j + v] = xp[v];
                        }
                    }
                };
                GluePre<G> cur;
                load_group(0, cur);
                mbar_wait(o_full, it & 1);
                tc_fence_after();
#pragma unroll 1
                for (int g0 = 0; g0 < NG; ++g0) {
                    GluePre<G> nxt;
                    if (g0 + 1 < NG) load_group(g0 + 1, nxt);
#pragma unroll
                    for (int j = 0; j < G; ++j) {
                        const int s = g0 * G + j;
                        uint32_t r[16];
                        tmem_ld16(tmem_base + lane_sel + TM_O + s * NB, r);
                        tmem_ld_wait();
                        if (valid) {
                            const int fy = ly * S + s / S, fx = lx * S + s % S;
                            const long pix = ((long)wc.b * fz.Hp + fy) * fz.Wp + fx;
                            float yv[NB];
                            const __half2* th = reinterpret_cast<const __half2*>(&cur.tk[2 * j]);
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const float2 tf = __half22float2(th[e]);
                                yv[2 * e] = fmaf(__uint_as_float(r[2 * e]), inv, tf.x);
                                yv[2 * e + 1] = fmaf(__uint_as_float(r[2 * e + 1]), inv, tf.y);
                            }
                            uint4 yo[2];
                            __half2* yh = reinterpret_cast<__half2*>(yo);
#pragma unroll
                            for (int e = 0; e < 8; ++e) yh[e] = __floats2half2_rn(yv[2 * e], yv[2 * e + 1]);
                            __half* yp = fz.Y + pix * NF + NB * br;
                            *reinterpret_cast<uint4*>(yp) = yo[0];
                            *reinterpret_cast<uint4*>(yp + 8) = yo[1];
                            if (has_next) {
                                uint4 to[2];
                                __half2* tnh = reinterpret_cast<__half2*>(to);
#pragma unroll
                                for (int v = 0; v < 4; ++v) {
                                    const float4 xv = cur.xv[4 * j + v];
                                    const float t0 = 0.5f * ((xv.x - mu[4 * v]) * rs[4 * v] + yv[4 * v]);
                                    const float t1 = 0.5f * ((xv.y - mu[4 * v + 1]) * rs[4 * v + 1] + yv[4 * v + 1]);
                                    const float t2 = 0.5f * ((xv.z - mu[4 * v + 2]) * rs[4 * v + 2] + yv[4 * v + 2]);
                                    const float t3 = 0.5f * ((xv.w - mu[4 * v + 3]) * rs[4 * v + 3] + yv[4 * v + 3]);
                                    tnh[2 * v] = __floats2half2_rn(t0, t1);
                                    tnh[2 * v + 1] = __floats2half2_rn(t2, t3);
                                }
                                const int sn = (fy & (Sn - 1)) * Sn + (fx & (Sn - 1));
                                __half* tp = fz.Tnext + ((((long)wc.b * (fz.Hp >> lvn)) + (fy >> lvn)) * (fz.Wp >> lvn) + (fx >> lvn)) * Cn + sn * NB;
                                *reinterpret_cast<uint4*>(tp) = to[0];
                                *reinterpret_cast<uint4*>(tp + 8) = to[1];
                            }
                        }
                    }
                    if (g0 + 1 < NG) cur = nxt;
                }
            } else {
                mbar_wait(o_full, it & 1);
                tc_fence_after();
                __half* orow = O + (((long)wc.b * h + wc.y + (qi >> 3)) * w + wc.x + (qi & 7)) * C;
                constexpr int OCH = C < 32 ? C : 32;
#pragma unroll 1
                for (int c0 = 0; c0 < C; c0 += OCH) {
                    uint32_t r[OCH];
                    if constexpr (OCH == 32) tmem_ld32(tmem_base + lane_sel + TM_O + c0, r);
                    else tmem_ld16(tmem_base + lane_sel + TM_O + c0, r);
                    tmem_ld_wait();
                    if (valid) {
#pragma unroll
                        for (int v = 0; v < OCH / 8; ++v) {
                            uint4 u;
                            uint32_t* pu = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const __half2 hv = __floats2half2_rn(__uint_as_float(r[v * 8 + 2 * e]) * inv,
                                                                     __uint_as_float(r[v * 8 + 2 * e + 1]) * inv);
                                pu[e] = *reinterpret_cast<const uint32_t*>(&hv);
                            }
                            *reinterpret_cast<uint4*>(orow + c0 + v * 8) = u;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(o_empty);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, CF::TM_COLS);
}

template <int C>
static int launch_attn_umma_c(const __half* QKV, const __half* relx, __half* O, int B, int h, int w, cudaStream_t s,
                              const AttnFuse* fuse) {
    using CF = AtCfg<C>;
    CUtensorMap mapQ, mapKV, mapR;
    const uint64_t dims[4] = {(uint64_t)3 * C, (uint64_t)w, (uint64_t)h, (uint64_t)B};
    const uint64_t str[4] = {2, (uint64_t)3 * C * 2, (uint64_t)w * 3 * C * 2, (uint64_t)h * w * 3 * C * 2};
    {
        const uint32_t box[4] = {(uint32_t)CF::CB, BLK, BLK, 1};
        M2T_TRY(make_tensor_map(&mapQ, QKV, 2, 4, dims, str, box, CF::TMA_SWZ));
    }
    {
        const uint32_t box[4] = {(uint32_t)CF::CB, WIN, WIN, 1};
        M2T_TRY(make_tensor_map(&mapKV, QKV, 2, 4, dims, str, box, CF::TMA_SWZ));
    }
    {
        const uint64_t d2[2] = {(uint64_t)C, 32}, s2[2] = {2, (uint64_t)C * 2};
        const uint32_t box[2] = {(uint32_t)CF::CB, 32};
        M2T_TRY(make_tensor_map(&mapR, relx, 2, 2, d2, s2, box, CF::TMA_SWZ));
    }
    const int nwin = B * (h / BLK) * (w / BLK);
    const int npairs = (nwin + 1) / 2;
    const int cap = device_sm_count() * CF::MIN_CTAS;
    const int grid = npairs < cap ? npairs : cap;
    if (fuse != nullptr) {
        M2T_ENSURE_SMEM((attn_umma_kernel<C, true>), CF::SMEM);
        attn_umma_kernel<C, true><<<grid, 192, CF::SMEM, s>>>(mapQ, mapKV, mapR, O, h, w, nwin, *fuse);
    } else {
        M2T_ENSURE_SMEM((attn_umma_kernel<C, false>), CF::SMEM);
        attn_umma_kernel<C, false><<<grid, 192, CF::SMEM, s>>>(mapQ, mapKV, mapR, O, h, w, nwin, AttnFuse{});
    }
    M2T_LAUNCH_CHECK("attn_umma_kernel");
    return M2T_OK;
}

int launch_attn_umma(int C, const __half* QKV, const __half* relx, __half* O, int B, int h, int w, cudaStream_t s,
                     const AttnFuse* fuse) {
    if (h % BLK || w % BLK) { set_error("attn: %dx%d is not a multiple of the 8x8 block", h, w); return M2T_E_ARG; }
    if (C == 16) return launch_attn_umma_c<16>(QKV, relx, O, B, h, w, s, fuse);
    if (C == 64) return launch_attn_umma_c<64>(QKV, relx, O, B, h, w, s, fuse);
    if (C == 256) return launch_attn_umma_c<256>(QKV, relx, O, B, h, w, s, fuse);
    set_error("attn: unsupported channel count %d", C);
    return M2T_E_UNSUPPORTED;
}

}  // namespace m2t
