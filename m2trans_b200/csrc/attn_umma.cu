// tcgen05 halo attention (ref M2Trans_network.py:310-332; block 8, halo 1, one head) for C in {16, 64, 256}.
//
// One work item = a PAIR of 8x8 query windows.  Each window is an M = 64 tcgen05 accumulator; an M = 64
// accumulator occupies TMEM lanes 0-15 of every 32-lane quadrant, so window B is placed at lane offset 16 in the
// SAME columns (tests/test_probes.py::test_umma_m64_two_windows_interleaved): all 128 lanes / epilogue threads are
// busy, nothing is block-diagonal, and one pair needs only 144 + C TMEM columns.
//   S   = Q . K^T                  per window M=64, N=112 (100 keys + 12 unused), K = C
//   S'  = Q . Rel^T                N = 32 extra columns: q[:C/2].rel_h[r] (cols 112..121), q[C/2:].rel_w[c]
//                                  (122..131): the reference adds rel to K, also at the zero-padded keys
//                                  (ref :322-325); q.(k+rel) = q.k + q.rel
//   P   = exp2((S + rel terms - max) * log2e), unnormalised, fp16, K-major / 128B-swizzled in smem (zero in the
//         12 padding key columns)
//   O   = P . V                    per window M=64, N = C, K = 112; rows scaled by 1/sum in the epilogue
// Keys/values outside the frame are TMA out-of-bounds zero fill = F.unfold's zero padding (ref :313-317);
// window partition / reverse (ref :310, :332) are TMA box coordinates and the store address.
// Operands stream in 64-channel blocks through two rings (Q+K blocks, V blocks).  C = 16 / 64 CTAs need <= 256
// TMEM columns and <= 110 KB of shared memory, so two CTAs share an SM and hide each other's latencies.
// Warp roles (192 threads): warps 0-3 softmax + epilogue (thread = query row), warp 4 TMA, warp 5 MMA issue;
// C = 256 fused adds warps 6-9, which share the epilogue of the rows of warps 0-3 (the epilogue is half of that kernel's chain).
//
// FUSE = true additionally folds the CFTM branch glue (ref :139-161) into the epilogue, see AttnFuse.
#include "common.cuh"
#include "gelu.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace m2t {

constexpr int WR = 112;                                       // key rows per window (100 + 12 zero rows)

template <int C, bool FUSE>
struct AtCfg {
    static constexpr int CB = C < 64 ? C : 64;               // channels per streamed block
    static constexpr int NBLK = C / CB;
    static constexpr uint32_t ROWB = CB * 2;                  // bytes per row (32 or 128)
    static constexpr uint64_t LAYOUT = CB == 64 ? UMMA_LAYOUT_SW128 : UMMA_LAYOUT_SW32;
    static constexpr int TMA_SWZ = CB == 64 ? 3 : 1;
    static constexpr uint32_t SBO = 8 * ROWB;
    static constexpr uint32_t QB = 128 * ROWB;                // Q block: two windows x 64 rows
    static constexpr uint32_t KVB = 2 * WR * ROWB;            // K or V block: two windows x 112 rows
    static constexpr uint32_t WIN_B = WR * ROWB;              // offset of window B's keys
    static constexpr int SQ = C == 64 ? 1 : 2;                // Q+K ring depth
    static constexpr int SV = C == 64 ? 1 : 2;                // V ring depth
    static constexpr bool TSTAGE = FUSE && C == 256;          // t_k rows staged through smem by TMA
    static constexpr int ST = TSTAGE ? 2 : 0;
    static constexpr uint32_t QK_STAGE = (QB + KVB + 1023) / 1024 * 1024;
    static constexpr uint32_t V_STAGE = (KVB + 1023) / 1024 * 1024;
    static constexpr uint32_t T_STAGE = QB;
    static constexpr uint32_t REL_BLOCK = 32 * ROWB;
    static constexpr uint32_t OFF_V = SQ * QK_STAGE;
    static constexpr uint32_t OFF_P = OFF_V + SV * V_STAGE;
    static constexpr uint32_t OFF_T = OFF_P + 2 * 16384;      // P: per window two 64-key blocks of [64 rows][128 B]
    static constexpr uint32_t OFF_REL = OFF_T + ST * T_STAGE;
    static constexpr uint32_t OFF_BAR = OFF_REL + (NBLK * REL_BLOCK + 1023) / 1024 * 1024;
    // C = 256 fused: a second epilogue warpgroup (warps 6-9) takes half of the sub-pixels of every 64-channel block;
    // the per-row 1/sum travels through smem (double-buffered by pair parity)
    static constexpr bool SPLIT = FUSE && C == 256;
    static constexpr int THREADS = SPLIT ? 320 : 192;
    static constexpr uint32_t OFF_INV = OFF_BAR + 256;
    static constexpr uint32_t SMEM = 1024 + OFF_BAR + 256 + (SPLIT ? 1024 : 0);
    static constexpr uint32_t TX_QK = 2 * 64 * ROWB + 2 * 100 * ROWB;
    static constexpr uint32_t TX_V = 2 * 100 * ROWB;
    static constexpr uint32_t TX_T = 2 * 64 * ROWB;
    // TMEM columns: S [0,112) | rel [112,144) | O
    static constexpr uint32_t TM_REL = WR;
    static constexpr uint32_t TM_O = C <= 64 ? 144 : 256;
    static constexpr uint32_t TM_COLS = C <= 64 ? 256 : 512;
    static constexpr int MIN_CTAS = C <= 64 ? 2 : 1;
};

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Division by a runtime constant as multiply-high + shifts (Granlund-Montgomery round-up form, exact for every
// 32-bit numerator).  The window -> (image, y, x) decomposition runs once per pair in every epilogue thread; the
// three hardware-emulated integer divisions it used to contain cost ~300 cycles of a 3.9 K-cycle pair at C = 16.
struct FastDiv {
    uint32_t m, s1, s2, d;
    __device__ __forceinline__ explicit FastDiv(uint32_t div) : d(div) {
        uint32_t l = 0;
        while ((1u << l) < div) ++l;                         // ceil(log2 d), d >= 1
        m = (uint32_t)(((uint64_t(1) << 32) * ((uint64_t(1) << l) - div)) / div + 1);
        s1 = l < 1 ? l : 1;
        s2 = l > 1 ? l - 1 : 0;
    }
    __device__ __forceinline__ uint32_t div(uint32_t n) const {
        const uint32_t t = __umulhi(m, n);
        return (t + ((n - t) >> s1)) >> s2;
    }
};

struct WinCoord { int b, y, x; };
__device__ __forceinline__ WinCoord win_coord(int wi, const FastDiv& nwx, const FastDiv& per_img) {
    WinCoord c;
    c.b = (int)per_img.div((uint32_t)wi);
    const uint32_t r = (uint32_t)wi - (uint32_t)c.b * per_img.d;
    const uint32_t ry = nwx.div(r);
    c.y = (int)ry * BLK;
    c.x = (int)(r - ry * nwx.d) * BLK;
    return c;
}

#ifdef M2T_TIMING
__device__ long long g_attn_dbg[4 * 64];   // one 64-entry record per branch (fused) / record 0 (plain)
#define M2T_T(slot) do { if (blockIdx.x == 0 && tid == 0 && it < 6) g_attn_dbg[64 * fz.branch + (slot) + 8 * it] = clock64(); } while (0)
#else
#define M2T_T(slot) do { } while (0)
#endif

// LO: the residual path uses t_k = T + Tlo (AttnFuse; precise mode).  A template parameter, not a run-time test: the
// fast mode must not carry the extra registers and adds.
template <int C, bool FUSE, bool LO>
__global__ void __launch_bounds__(AtCfg<C, FUSE>::THREADS, AtCfg<C, FUSE>::MIN_CTAS)
attn_umma_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapKV,
                 const __grid_constant__ CUtensorMap mapR, const __grid_constant__ CUtensorMap mapT,
                 __half* __restrict__ O, int h, int w, int nwin, const AttnFuse fz) {
    using CF = AtCfg<C, FUSE>;
    constexpr int CB = CF::CB, NBLK = CF::NBLK, SQ = CF::SQ, SV = CF::SV;
    constexpr uint32_t ROWB = CF::ROWB;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + CF::OFF_BAR);
    uint64_t* q_full = bars;                 // [2]
    uint64_t* q_empty = bars + 2;            // [2]
    uint64_t* v_full = bars + 4;             // [2]
    uint64_t* v_empty = bars + 6;            // [2]
    uint64_t* t_full = bars + 8;             // [2]
    uint64_t* t_empty = bars + 10;           // [2]
    uint64_t* rfull = bars + 12;             // rel tables landed
    uint64_t* s_full = bars + 13;            // S complete in TMEM
    uint64_t* p_ready = bars + 14;           // P written to smem, S consumed
    uint64_t* o_full = bars + 15;            // O complete in TMEM
    uint64_t* o_empty = bars + 16;           // O consumed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const FastDiv nwx((uint32_t)(w / BLK)), per_img((uint32_t)((h / BLK) * (w / BLK)));
    const int npairs = (nwin + 1) / 2;
#ifdef M2T_TIMING
    if (blockIdx.x == 0 && tid == 0) g_attn_dbg[64 * fz.branch + 6] = clock64();
#endif

    // zero P (the 12 padding key columns stay zero for ever) and the 12 padding rows of every V stage
    for (uint32_t i = tid * 16; i < 2 * 16384; i += CF::THREADS * 16) *reinterpret_cast<uint4*>(sm + CF::OFF_P + i) = make_uint4(0, 0, 0, 0);
    for (int s = 0; s < SV; ++s)
        for (int win = 0; win < 2; ++win) {
            uint8_t* pad = sm + CF::OFF_V + s * CF::V_STAGE + win * CF::WIN_B + 100 * ROWB;
            for (uint32_t i = tid * 16; i < (WR - 100) * ROWB; i += CF::THREADS * 16) *reinterpret_cast<uint4*>(pad + i) = make_uint4(0, 0, 0, 0);
        }
    fence_proxy_async();
    if (warp == 5) tmem_alloc(tmem_slot, CF::TM_COLS);
    if (tid == 128) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1);
            mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
            mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], CF::SPLIT ? 8 : 4);
        }
        mbar_init(rfull, 1);
        mbar_init(s_full, 1);
        mbar_init(p_ready, 4);
        mbar_init(o_full, 1);
        mbar_init(o_empty, CF::SPLIT ? 8 : 4);
        mbar_fence_init();
        tma_prefetch_desc(&mapQ);
        tma_prefetch_desc(&mapKV);
        tma_prefetch_desc(&mapR);
        if constexpr (CF::TSTAGE) tma_prefetch_desc(&mapT);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t TM_S = 0, TM_REL = CF::TM_REL, TM_O = CF::TM_O;
    constexpr uint32_t LANE_B = 16u << 16;                      // TMEM lane offset of window B
    pdl_trigger();
#ifdef M2T_TIMING
    if (blockIdx.x == 0 && tid == 0) g_attn_dbg[64 * fz.branch + 7] = clock64();
#endif

    if (warp == 4) {
        // TMA producer: the whole warp runs the loops, one elected lane issues
        if (elect_one_sync()) {      // rel tables are constants: loaded while the previous kernel drains
            mbar_expect_tx(rfull, NBLK * CF::REL_BLOCK);
            for (int kb = 0; kb < NBLK; ++kb) tma_load_2d(sm + CF::OFF_REL + kb * CF::REL_BLOCK, &mapR, rfull, kb * CB, 0);
        }
        pdl_wait();
        uint32_t gq = 0, gv = 0, gt = 0;
        (void)gt;
        for (int p = blockIdx.x; p < npairs; p += gridDim.x) {
            const int wa = 2 * p, wb = (2 * p + 1 < nwin) ? 2 * p + 1 : 2 * p;
            const WinCoord a = win_coord(wa, nwx, per_img), b = win_coord(wb, nwx, per_img);
            for (int kb = 0; kb < NBLK; ++kb, ++gq) {
                const uint32_t s = gq % SQ, ph = (gq / SQ) & 1;
                uint8_t* st = sm + s * CF::QK_STAGE;
                mbar_wait(&q_empty[s], ph ^ 1);
                if (elect_one_sync()) {
                    mbar_expect_tx(&q_full[s], CF::TX_QK);
                    tma_load_4d(st, &mapQ, &q_full[s], kb * CB, a.x, a.y, a.b);
                    tma_load_4d(st + 64 * ROWB, &mapQ, &q_full[s], kb * CB, b.x, b.y, b.b);
                    tma_load_4d(st + CF::QB, &mapKV, &q_full[s], C + kb * CB, a.x - 1, a.y - 1, a.b);
                    tma_load_4d(st + CF::QB + CF::WIN_B, &mapKV, &q_full[s], C + kb * CB, b.x - 1, b.y - 1, b.b);
                }
                __syncwarp();
            }
            // t_k rows first (their stages were released by the previous pair's epilogue), then V
            auto load_t = [&](int nb) {
                if constexpr (CF::TSTAGE) {
                    const uint32_t s = gt & 1, ph = (gt >> 1) & 1;
                    uint8_t* st = sm + CF::OFF_T + s * CF::T_STAGE;
                    mbar_wait(&t_empty[s], ph ^ 1);
                    if (elect_one_sync()) {
                        mbar_expect_tx(&t_full[s], CF::TX_T);
                        tma_load_4d(st, &mapT, &t_full[s], nb * CB, a.x, a.y, a.b);
                        tma_load_4d(st + 64 * ROWB, &mapT, &t_full[s], nb * CB, b.x, b.y, b.b);
                    }
                    __syncwarp();
                    ++gt;
                }
            };
            if constexpr (CF::TSTAGE) { load_t(0); load_t(1); }
            for (int nb = 0; nb < NBLK; ++nb, ++gv) {
                const uint32_t s = gv % SV, ph = (gv / SV) & 1;
                uint8_t* st = sm + CF::OFF_V + s * CF::V_STAGE;
                mbar_wait(&v_empty[s], ph ^ 1);
                if (elect_one_sync()) {
                    mbar_expect_tx(&v_full[s], CF::TX_V);
                    tma_load_4d(st, &mapKV, &v_full[s], 2 * C + nb * CB, a.x - 1, a.y - 1, a.b);
                    tma_load_4d(st + CF::WIN_B, &mapKV, &v_full[s], 2 * C + nb * CB, b.x - 1, b.y - 1, b.b);
                }
                __syncwarp();
            }
            if constexpr (CF::TSTAGE) { load_t(2); load_t(3); }
        }
    } else if (warp == 5) {
        // MMA issuer: warp-uniform loops, one elected lane issues
        constexpr uint32_t id_s = umma_idesc_f16(64, WR), id_r = umma_idesc_f16(64, 32);
        constexpr uint32_t id_o = umma_idesc_f16(64, CB, 0, 1);
        constexpr uint64_t tmpl = umma_smem_desc(0, 16, CF::SBO, CF::LAYOUT);
        constexpr uint64_t tmpl_p = umma_smem_desc(0, 16, 1024, UMMA_LAYOUT_SW128);
        mbar_wait(rfull, 0);
        uint32_t gq = 0, gv = 0, it = 0;
        for (int p = blockIdx.x; p < npairs; p += gridDim.x, ++it) {
            // S and the rel columns, both windows
            for (int kb = 0; kb < NBLK; ++kb, ++gq) {
                const uint32_t s = gq % SQ, ph = (gq / SQ) & 1;
                mbar_wait(&q_full[s], ph);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t st = base + s * CF::QK_STAGE;
                    const uint64_t dq0 = umma_desc_at(tmpl, st), dk0 = umma_desc_at(tmpl, st + CF::QB);
                    const uint64_t dr0 = umma_desc_at(tmpl, base + CF::OFF_REL + kb * CF::REL_BLOCK);
#pragma unroll
                    for (int k = 0; k < CB / 16; ++k) {
                        const uint32_t accum = (kb | k) ? 1u : 0u;
#pragma unroll
                        for (int win = 0; win < 2; ++win) {
                            const uint64_t dq = dq0 + (uint64_t)((win * 64 * ROWB + k * 32) >> 4);
                            const uint64_t dk = dk0 + (uint64_t)((win * CF::WIN_B + k * 32) >> 4);
                            umma_f16_ss(tmem_base + TM_S + win * LANE_B, dq, dk, id_s, accum);
                            umma_f16_ss(tmem_base + TM_REL + win * LANE_B, dq, dr0 + 2 * k, id_r, accum);
                        }
                    }
                    umma_commit(&q_empty[s]);
                    if (kb == NBLK - 1) umma_commit(s_full);
                }
                __syncwarp();
            }
            // O = P . V
            mbar_wait(p_ready, it & 1);
            mbar_wait(o_empty, (it & 1) ^ 1);
            tc_fence_after();
            for (int nb = 0; nb < NBLK; ++nb, ++gv) {
                const uint32_t s = gv % SV, ph = (gv / SV) & 1;
                mbar_wait(&v_full[s], ph);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint64_t dv0 = umma_desc_at(tmpl, base + CF::OFF_V + s * CF::V_STAGE);
                    const uint64_t dp0 = umma_desc_at(tmpl_p, base + CF::OFF_P);
#pragma unroll
                    for (int win = 0; win < 2; ++win) {
#pragma unroll
                        for (int k = 0; k < WR / 16; ++k) {
                            const uint64_t dp = dp0 + (uint64_t)((win * 16384 + (k >> 2) * 8192 + (k & 3) * 32) >> 4);
                            const uint64_t dv = dv0 + (uint64_t)((win * CF::WIN_B + k * 16 * ROWB) >> 4);
                            umma_f16_ss(tmem_base + TM_O + nb * CB + win * LANE_B, dp, dv, id_o, k ? 1u : 0u);
                        }
                    }
                    umma_commit(&v_empty[s]);
                    if (nb == NBLK - 1) umma_commit(o_full);
                }
                __syncwarp();
            }
        }
    } else {
        // thread (warp, lane) owns TMEM lane 32*quad + lane: window (lane >> 4), query row 16*quad + (lane & 15).
        // Warps 0-3 do the softmax and (their share of) the epilogue; helper warps 6-9 (SPLIT only) share the rows of
        // the primary warp with the same quadrant and only run the epilogue.
        const int quad = warp & 3;
        const bool helper = CF::SPLIT && warp >= 6;
        const int win = lane >> 4, qi = quad * 16 + (lane & 15);
        const uint32_t lane_sel = (uint32_t)(quad * 32) << 16;
        float* sinv = reinterpret_cast<float*>(sm + CF::OFF_INV);
        (void)sinv;
        uint8_t* prow = sm + CF::OFF_P + win * 16384 + qi * 128;
        uint32_t it = 0, gt = 0;
        (void)gt;
        pdl_wait();
        for (int p = blockIdx.x; p < npairs; p += gridDim.x, ++it) {
            float inv;
            if (!helper) {
            M2T_T(0);
            mbar_wait(s_full, it & 1);
            tc_fence_after();
            M2T_T(1);
            float sv[104];
            uint32_t ab[24];
            {
                uint32_t* su = reinterpret_cast<uint32_t*>(sv);
                const uint32_t t0 = tmem_base + lane_sel + TM_S;
                tmem_ld32(t0, su);
                tmem_ld32(t0 + 32, su + 32);
                tmem_ld32(t0 + 64, su + 64);
                tmem_ld8(t0 + 96, su + 96);
                tmem_ld16(tmem_base + lane_sel + TM_REL, ab);
                tmem_ld8(tmem_base + lane_sel + TM_REL + 16, ab + 16);
                tmem_ld_wait();
            }
            // Softmax over the 100 keys, two keys per instruction on the packed fp32x2 pipe (keys 2i and 2i+1 lie in
            // the same key row because the row length is even, so they share the rel_h term): 1 FADD2 + 1/2 FMNMX3 +
            // 1/2 FFMA2 + 1 MUFU + 1/2 FADD2 + 1/2 F2FP per key instead of 6.5 scalar instructions.
            uint64_t rw2[WIN / 2];
#pragma unroll
            for (int c = 0; c < WIN / 2; ++c) rw2[c] = f2_pack(__uint_as_float(ab[10 + 2 * c]), __uint_as_float(ab[11 + 2 * c]));
            uint64_t s2[NKEY / 2];
            float mx = -INFINITY;
#pragma unroll
            for (int r = 0; r < WIN; ++r) {
                const uint64_t rh2 = f2_splat(__uint_as_float(ab[r]));
#pragma unroll
                for (int c = 0; c < WIN / 2; ++c) {
                    const int i = r * (WIN / 2) + c;
                    s2[i] = f2_add(f2_pack(sv[2 * i], sv[2 * i + 1]), f2_add(rw2[c], rh2));
                    float a0, a1;
                    f2_unpack(s2[i], a0, a1);
                    asm("max.f32 %0, %0, %1, %2;" : "+f"(mx) : "f"(a0), "f"(a1));
                }
            }
            const float mxl = mx * 1.4426950408889634f;
            const uint64_t l2e = f2_splat(1.4426950408889634f), nmx = f2_splat(-mxl);
            uint64_t sum2 = f2_splat(0.f);
            uint32_t ph[52];                               // P row as fp16 pairs; pairs 50, 51 = padding keys 100..103
#pragma unroll
            for (int i = 0; i < NKEY / 2; ++i) {
                float a0, a1;
                f2_unpack(f2_fma(s2[i], l2e, nmx), a0, a1);
                const float e0 = fast_exp2(a0), e1 = fast_exp2(a1);
                sum2 = f2_add(sum2, f2_pack(e0, e1));
                const __half2 hv = __floats2half2_rn(e0, e1);
                ph[i] = *reinterpret_cast<const uint32_t*>(&hv);
            }
            ph[50] = 0u; ph[51] = 0u;
            float sum;
            {
                float a0, a1;
                f2_unpack(sum2, a0, a1);
                sum = a0 + a1;
            }
            // P row: 13 chunks of 8 keys (keys 100..103 written as zeros, chunk 13 stays zero from the prologue)
#pragma unroll
            for (int q = 0; q < 13; ++q)
                *reinterpret_cast<uint4*>(prow + (q >> 3) * 8192 + (((q & 7) ^ (qi & 7)) << 4)) =
                    make_uint4(ph[4 * q], ph[4 * q + 1], ph[4 * q + 2], ph[4 * q + 3]);
            inv = 1.f / sum;
            if constexpr (CF::SPLIT) sinv[(it & 1) * 128 + quad * 32 + lane] = inv;
            M2T_T(2);
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_ready);
            } else {
                mbar_wait(p_ready, it & 1);        // the primaries have published 1/sum of this pair
                inv = sinv[(it & 1) * 128 + quad * 32 + lane];
            }

            const int wi = 2 * p + win;
            const bool valid = wi < nwin;
            const WinCoord wc = win_coord(valid ? wi : 2 * p, nwx, per_img);
            if constexpr (FUSE) {
                // Fused branch glue (ref :139-161).  With the Haar transforms folded into the qkv weights the
                // accumulator row IS IWT^L(attention) in space-to-depth order: column s*16+k belongs to pixel
                // s = dy*2^L + dx of this level pixel's 2^L x 2^L block.
                //   y_k = O' + t_k                 -> Y[..., 16k..16k+15]               (ref :139,:145,:153,:161)
                //   t_{k+1} = n_{k+1}/2 + y_k/2    -> Tnext, updated in place (branch_prep_all pre-filled n_{k+1}/2 in
                //                                     the next level's space-to-depth order)  (ref :141,:147,:155)
                // t_k rows come from registers (C <= 64, loaded before the PV MMAs finish) or from the TMA-fed T
                // stages (C = 256); the Tnext segments are prefetched one 64-channel block ahead.
                constexpr int LV = C == 16 ? 0 : (C == 64 ? 1 : 2);
                constexpr int S = 1 << LV;
                constexpr int SPB = CB / NB;                     // sub-pixels per 64-channel block: 1 or 4
                const int ly = wc.y + (qi >> 3), lx = wc.x + (qi & 7);
                const int br = fz.branch;
                const bool has_next = fz.Tnext != nullptr;
                const int lvn = br == 0 ? 1 : 2, Sn = 1 << lvn, Cn = NB * Sn * Sn;
                auto tnext_off = [&](int s) -> long {
                    const int fy = ly * S + s / S, fx = lx * S + s % S;
                    const int sn = (fy & (Sn - 1)) * Sn + (fx & (Sn - 1));
                    return ((((long)wc.b * (fz.Hp >> lvn)) + (fy >> lvn)) * (fz.Wp >> lvn) + (fx >> lvn)) * Cn + sn * NB;
                };
                const long trow_off = (((long)wc.b * h + ly) * w + lx) * C;
                uint4 tkr[CF::TSTAGE ? 1 : 2 * SPB];            // register copy of the t_k row (C <= 64)
                if constexpr (!CF::TSTAGE) {
                    const __half* trow = fz.T + trow_off;
#pragma unroll
                    for (int j = 0; j < SPB; ++j) ldg256(trow + j * 16, tkr[2 * j], tkr[2 * j + 1]);
                }
                constexpr int JN = CF::SPLIT ? SPB / 2 : SPB;    // sub-pixels of each block this thread handles
                const int j0 = helper ? SPB / 2 : 0;
                uint4 hcur[2 * JN], hnxt[2 * JN];               // n_{k+1}/2 segments of the current / next block
                auto load_h = [&](int nb, uint4* dst) {
#pragma unroll
                    for (int j = 0; j < JN; ++j) ldg256(fz.Tnext + tnext_off(nb * SPB + j0 + j), dst[2 * j], dst[2 * j + 1]);
                };
                // rounding residuals of t_k (AttnFuse::Tlo): the residual add below uses t_k = T + Tlo
                uint4 lcur[2 * JN], lnxt[2 * JN];
                constexpr bool has_lo = LO;
                auto load_l = [&](int nb, uint4* dst) {
#pragma unroll
                    for (int j = 0; j < JN; ++j) {
                        if (has_lo) ldg256(fz.Tlo + trow_off + (nb * SPB + j0 + j) * NB, dst[2 * j], dst[2 * j + 1]);
                        else { dst[2 * j] = make_uint4(0u, 0u, 0u, 0u); dst[2 * j + 1] = make_uint4(0u, 0u, 0u, 0u); }
                    }
                };
                load_l(0, lcur);
                if (has_next) load_h(0, hcur);
                M2T_T(3);
                mbar_wait(o_full, it & 1);
                tc_fence_after();
                M2T_T(4);
#pragma unroll 1
                for (int nb = 0; nb < NBLK; ++nb) {
                    if (nb + 1 < NBLK) {
                        load_l(nb + 1, lnxt);
                        if (has_next) load_h(nb + 1, hnxt);
                    }
                    const uint8_t* tst = nullptr;
                    if constexpr (CF::TSTAGE) {
                        mbar_wait(&t_full[gt & 1], (gt >> 1) & 1);
                        tst = sm + CF::OFF_T + (gt & 1) * CF::T_STAGE + win * 64 * ROWB + qi * 128;
                    }
#pragma unroll
                    for (int jj = 0; jj < JN; ++jj) {
                        const int j = j0 + jj;
                        const int s = nb * SPB + j;
                        uint32_t r[16];
                        tmem_ld16(tmem_base + lane_sel + TM_O + s * NB, r);
                        uint4 tk[2];
                        if constexpr (CF::TSTAGE) {
                            tk[0] = *reinterpret_cast<const uint4*>(tst + (((2 * j) ^ (qi & 7)) << 4));
                            tk[1] = *reinterpret_cast<const uint4*>(tst + (((2 * j + 1) ^ (qi & 7)) << 4));
                        } else {
                            tk[0] = tkr[2 * j]; tk[1] = tkr[2 * j + 1];
                        }
                        tmem_ld_wait();
                        if (valid) {
                            const int fy = ly * S + s / S, fx = lx * S + s % S;
                            const long pix = ((long)wc.b * fz.Hp + fy) * fz.Wp + fx;
                            float yv[NB];
                            const __half2* th = reinterpret_cast<const __half2*>(tk);
                            const __half2* tl = reinterpret_cast<const __half2*>(&lcur[2 * jj]);
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                float2 tf = __half22float2(th[e]);
                                if constexpr (LO) {
                                    const float2 lf = __half22float2(tl[e]);
                                    tf.x += lf.x; tf.y += lf.y;
                                }
                                yv[2 * e] = fmaf(__uint_as_float(r[2 * e]), inv, tf.x);
                                yv[2 * e + 1] = fmaf(__uint_as_float(r[2 * e + 1]), inv, tf.y);
                            }
                            uint4 yo[2];
                            __half2* yh = reinterpret_cast<__half2*>(yo);
#pragma unroll
                            for (int e = 0; e < 8; ++e) yh[e] = __floats2half2_rn(yv[2 * e], yv[2 * e + 1]);
                            __half* yp = fz.Y + pix * NF + NB * br;
                            stg256(yp, yo[0], yo[1]);
                            if constexpr (LO) {          // rounding residual of y_k * 2^11 for the split-precision ff conv
                                uint4 yl[2];
                                __half2* ylh = reinterpret_cast<__half2*>(yl);
#pragma unroll
                                for (int e = 0; e < 8; ++e) {
                                    const float2 yr = __half22float2(yh[e]);
                                    ylh[e] = __floats2half2_rn((yv[2 * e] - yr.x) * 2048.f, (yv[2 * e + 1] - yr.y) * 2048.f);
                                }
                                stg256(fz.Ylo + pix * NF + NB * br, yl[0], yl[1]);
                            }
                            if (has_next) {
                                uint4 to[2], tol[2];
                                __half2* tnh = reinterpret_cast<__half2*>(to);
                                __half2* tnl = reinterpret_cast<__half2*>(tol);
                                const __half2* hh = reinterpret_cast<const __half2*>(&hcur[2 * jj]);
#pragma unroll
                                for (int e = 0; e < 8; ++e) {
                                    const float2 hf = __half22float2(hh[e]);
                                    const float t0 = fmaf(0.5f, yv[2 * e], hf.x), t1 = fmaf(0.5f, yv[2 * e + 1], hf.y);
                                    tnh[e] = __floats2half2_rn(t0, t1);
                                    const float2 tr = __half22float2(tnh[e]);
                                    tnl[e] = __floats2half2_rn(t0 - tr.x, t1 - tr.y);
                                }
                                const long toff = tnext_off(s);
                                stg256(fz.Tnext + toff, to[0], to[1]);
                                if (has_lo) stg256(fz.Tnext_lo + toff, tol[0], tol[1]);
                            }
                        }
                    }
                    if constexpr (CF::TSTAGE) {
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&t_empty[gt & 1]);
                        ++gt;
                    }
                    if (nb + 1 < NBLK) {
#pragma unroll
                        for (int j = 0; j < 2 * JN; ++j) { lcur[j] = lnxt[j]; if (has_next) hcur[j] = hnxt[j]; }
                    }
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
            M2T_T(5);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, CF::TM_COLS);
}

template <int C, bool FUSE, bool LO>
static int launch_attn_umma_cf(const __half* QKV, const __half* relx, __half* O, int B, int h, int w, cudaStream_t s,
                               const AttnFuse& fz) {
    using CF = AtCfg<C, FUSE>;
    CUtensorMap mapQ, mapKV, mapR, mapT;
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
    if (CF::TSTAGE) {
        const uint64_t dt[4] = {(uint64_t)C, (uint64_t)w, (uint64_t)h, (uint64_t)B};
        const uint64_t st[4] = {2, (uint64_t)C * 2, (uint64_t)w * C * 2, (uint64_t)h * w * C * 2};
        const uint32_t box[4] = {(uint32_t)CF::CB, BLK, BLK, 1};
        M2T_TRY(make_tensor_map(&mapT, fz.T, 2, 4, dt, st, box, CF::TMA_SWZ));
    } else {
        mapT = mapQ;
    }
    const int nwin = B * (h / BLK) * (w / BLK);
    const int npairs = (nwin + 1) / 2;
    const int cap = device_sm_count() * CF::MIN_CTAS;
    const int grid = npairs < cap ? npairs : cap;
    M2T_ENSURE_SMEM((attn_umma_kernel<C, FUSE, LO>), CF::SMEM);
    M2T_CUDA(launch_pdl(attn_umma_kernel<C, FUSE, LO>, dim3(grid), dim3(CF::THREADS), CF::SMEM, s, mapQ, mapKV, mapR, mapT, O, h, w, nwin, fz));
    return M2T_OK;
}

template <int C>
static int launch_attn_umma_c(const __half* QKV, const __half* relx, __half* O, int B, int h, int w, cudaStream_t s,
                              const AttnFuse* fuse) {
    if (fuse != nullptr && fuse->Tlo != nullptr) return launch_attn_umma_cf<C, true, true>(QKV, relx, O, B, h, w, s, *fuse);
    if (fuse != nullptr) return launch_attn_umma_cf<C, true, false>(QKV, relx, O, B, h, w, s, *fuse);
    return launch_attn_umma_cf<C, false, false>(QKV, relx, O, B, h, w, s, AttnFuse{});
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

#ifdef M2T_TIMING
int read_attn_timing(long long* host64) {
    M2T_CUDA(cudaMemcpyFromSymbol(host64, g_attn_dbg, sizeof(long long) * 256));
    return M2T_OK;
}
#else
int read_attn_timing(long long* host64) { memset(host64, 0, sizeof(long long) * 256); return M2T_OK; }
#endif

}  // namespace m2t
