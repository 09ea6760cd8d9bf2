// Branch 1 of a CFTM in ONE kernel: qkv 1x1 conv + halo attention + branch glue (ref M2Trans_network.py:139, :307-332)
// for C = 16.  The separate qkv kernel wrote 96 B/px of QKV and the attention kernel read back ~150 B/px (halo
// included); here a window pair loads only its two 10 x 10-pixel tiles of t_1 (32 B/px), computes q, k, v for the 100
// halo pixels on the tensor core and keeps them in shared memory.  The price is the halo recompute (100 instead of 64
// pixels per window: 1.56x of a GEMM with K = 16).
//
// Per window pair:
//   1. TMA: the two t_1 tiles (box 10 x 10 x 16 ch at (8bx-1, 8by-1), out-of-frame pixels -> 0, so q = k = v = 0
//      there exactly like the reference's zero-padded unfold of k and v, ref :313-317).
//   2. MMA: QKV[128 rows (100 used)][48] = T . Wqkv^T per window (M128 x N48 x K16), accumulators in TMEM.
//   3. the four compute warps (thread = tile row) convert to fp16 and write the K and V operand tiles (all 100 rows)
//      and the Q operand tile (the 64 interior rows) in the 32-byte-swizzled layouts the attention MMAs expect.
//   4. from here on the pair runs exactly like attn_umma_kernel<16, true>: S = Q K^T (+ q.rel columns), softmax
//      thread-per-row from TMEM, P -> smem, O = P V, fused epilogue (y_1 into Y, t_2 completed in place).  The
//      t_1 row of the residual add comes from the tile already in shared memory.
// TMEM (256 columns, two CTAs per SM): S [0,112) | rel [112,144) | O [144,160) | QKV window A [160,208) | B [208,256).
#include "common.cuh"
#include "gelu.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace m2t {
namespace {

constexpr int AQ_C = 16;
constexpr int AQ_WR = 112;                                  // key rows per window (100 + 12 unused)
constexpr uint32_t AQ_ROWB = 32;                            // bytes per operand row (16 fp16 channels)
constexpr uint32_t AQ_SBO = 8 * AQ_ROWB;
constexpr uint32_t AQ_TWIN = 128 * AQ_ROWB;                 // one t tile as a 128-row A operand
constexpr uint32_t AQ_TSTAGE = 2 * AQ_TWIN;
constexpr uint32_t AQ_KWIN = AQ_WR * AQ_ROWB;               // 3584
constexpr uint32_t AQ_OFF_W = 0;                            // 48 x 32 B
constexpr uint32_t AQ_OFF_REL = 2048;                       // 32 x 32 B
constexpr uint32_t AQ_OFF_T = 4096;                         // 2 stages
constexpr uint32_t AQ_OFF_Q = AQ_OFF_T + 2 * AQ_TSTAGE;     // 128 rows
constexpr uint32_t AQ_OFF_K = AQ_OFF_Q + 128 * AQ_ROWB;     // 2 x 112 rows, padded to 8 KB
constexpr uint32_t AQ_OFF_V = AQ_OFF_K + 8192;
constexpr uint32_t AQ_OFF_P = AQ_OFF_V + 8192;              // per window two 64-key blocks of [64 rows][128 B]
constexpr uint32_t AQ_OFF_BAR = AQ_OFF_P + 2 * 16384;
constexpr uint32_t AQ_SMEM = 1024 + AQ_OFF_BAR + 256;
constexpr uint32_t AQ_TM_S = 0, AQ_TM_REL = 112, AQ_TM_O = 144, AQ_TM_QKV = 160;
constexpr uint32_t AQ_LANE_B = 16u << 16;                   // TMEM lane offset of window B (M = 64 accumulators)

__device__ __forceinline__ float aq_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct AqDiv {                                              // division by a runtime constant, see attn_umma.cu
    uint32_t m, s1, s2, d;
    __device__ __forceinline__ explicit AqDiv(uint32_t div) : d(div) {
        uint32_t l = 0;
        while ((1u << l) < div) ++l;
        m = (uint32_t)(((uint64_t(1) << 32) * ((uint64_t(1) << l) - div)) / div + 1);
        s1 = l < 1 ? l : 1;
        s2 = l > 1 ? l - 1 : 0;
    }
    __device__ __forceinline__ uint32_t div(uint32_t n) const {
        const uint32_t t = __umulhi(m, n);
        return (t + ((n - t) >> s1)) >> s2;
    }
};
struct AqCoord { int b, y, x; };
__device__ __forceinline__ AqCoord aq_coord(int wi, const AqDiv& nwx, const AqDiv& per_img) {
    AqCoord c;
    c.b = (int)per_img.div((uint32_t)wi);
    const uint32_t r = (uint32_t)wi - (uint32_t)c.b * per_img.d;
    const uint32_t ry = nwx.div(r);
    c.y = (int)ry * BLK;
    c.x = (int)(r - ry * nwx.d) * BLK;
    return c;
}
// byte offset of 16-byte chunk c (0 / 1) of row r in a 32-byte-swizzled tile of 32-byte rows
__device__ __forceinline__ uint32_t aq_sw32(uint32_t r, uint32_t c) { return r * 32u + ((c ^ ((r >> 2) & 1u)) << 4); }

template <bool LO>
__global__ void __launch_bounds__(192, 2)
attn16_qkv_kernel(const __grid_constant__ CUtensorMap mapT, const __grid_constant__ CUtensorMap mapW,
                  const __grid_constant__ CUtensorMap mapR, int h, int w, int nwin, const AttnFuse fz) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + AQ_OFF_BAR);
    uint64_t* wfull = bars;                  // weights + rel tables landed
    uint64_t* t_full = bars + 1;             // [2] t tiles landed
    uint64_t* t_empty = bars + 3;            // [2] ... no longer needed (qkv MMA and epilogue done)
    uint64_t* qkv_full = bars + 5;           // qkv accumulators complete
    uint64_t* qkv_empty = bars + 6;          // ... drained
    uint64_t* qk_ready = bars + 7;           // Q / K / V operand tiles written
    uint64_t* qk_free = bars + 8;            // S MMAs have read Q and K
    uint64_t* v_free = bars + 9;             // PV MMAs have read V (and P)
    uint64_t* s_full = bars + 10;
    uint64_t* p_ready = bars + 11;
    uint64_t* o_full = bars + 12;
    uint64_t* o_empty = bars + 13;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const AqDiv nwx((uint32_t)(w / BLK)), per_img((uint32_t)((h / BLK) * (w / BLK)));
    const int npairs = (nwin + 1) / 2;

    // zero P (its 12 padding key columns stay zero for ever) and the V tiles (their 12 padding rows likewise)
    for (uint32_t i = tid * 16; i < 2 * 16384; i += 192 * 16) *reinterpret_cast<uint4*>(sm + AQ_OFF_P + i) = make_uint4(0, 0, 0, 0);
    for (uint32_t i = tid * 16; i < 8192; i += 192 * 16) *reinterpret_cast<uint4*>(sm + AQ_OFF_V + i) = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    if (warp == 5) tmem_alloc(tmem_slot, 256);
    if (tid == 128) {
        mbar_init(wfull, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&t_full[s], 1); mbar_init(&t_empty[s], 5); }
        mbar_init(qkv_full, 1); mbar_init(qkv_empty, 4);
        mbar_init(qk_ready, 4); mbar_init(qk_free, 1); mbar_init(v_free, 1);
        mbar_init(s_full, 1); mbar_init(p_ready, 4); mbar_init(o_full, 1); mbar_init(o_empty, 4);
        mbar_fence_init();
        tma_prefetch_desc(&mapT);
        tma_prefetch_desc(&mapW);
        tma_prefetch_desc(&mapR);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();

    if (warp == 4) {
        // ---- TMA producer ---------------------------------------------------------------------------------
        if (elect_one_sync()) {              // constants: loaded while the previous kernel drains
            mbar_expect_tx(wfull, 48 * AQ_ROWB + 32 * AQ_ROWB);
            tma_load_2d(sm + AQ_OFF_W, &mapW, wfull, 0, 0);
            tma_load_2d(sm + AQ_OFF_REL, &mapR, wfull, 0, 0);
        }
        pdl_wait();
        uint32_t it = 0;
        for (int p = blockIdx.x; p < npairs; p += gridDim.x, ++it) {
            const int wa = 2 * p, wb = (2 * p + 1 < nwin) ? 2 * p + 1 : 2 * p;
            const AqCoord a = aq_coord(wa, nwx, per_img), b = aq_coord(wb, nwx, per_img);
            const uint32_t s = it & 1, ph = (it >> 1) & 1;
            mbar_wait(&t_empty[s], ph ^ 1);
            if (elect_one_sync()) {
                uint8_t* st = sm + AQ_OFF_T + s * AQ_TSTAGE;
                mbar_expect_tx(&t_full[s], 2 * 100 * AQ_ROWB);
                tma_load_4d(st, &mapT, &t_full[s], 0, a.x - 1, a.y - 1, a.b);
                tma_load_4d(st + AQ_TWIN, &mapT, &t_full[s], 0, b.x - 1, b.y - 1, b.b);
            }
            __syncwarp();
        }
    } else if (warp == 5) {
        // ---- MMA issuer -----------------------------------------------------------------------------------
        constexpr uint32_t id_qkv = umma_idesc_f16(128, 48);
        constexpr uint32_t id_s = umma_idesc_f16(64, AQ_WR), id_r = umma_idesc_f16(64, 32);
        constexpr uint32_t id_o = umma_idesc_f16(64, AQ_C, 0, 1);
        constexpr uint64_t tmpl = umma_smem_desc(0, 16, AQ_SBO, UMMA_LAYOUT_SW32);
        constexpr uint64_t tmpl_p = umma_smem_desc(0, 16, 1024, UMMA_LAYOUT_SW128);
        mbar_wait(wfull, 0);
        // qkv of pair j: issued one pair ahead so that it runs under the previous pair's softmax / epilogue
        auto issue_qkv = [&](uint32_t j) {
            const uint32_t s = j & 1, ph = (j >> 1) & 1;
            mbar_wait(&t_full[s], ph);
            mbar_wait(qkv_empty, (j & 1) ^ 1);
            tc_fence_after();
            if (elect_one_sync()) {
                const uint64_t dw = umma_desc_at(tmpl, base + AQ_OFF_W);
#pragma unroll
                for (int win = 0; win < 2; ++win)
                    umma_f16_ss(tmem_base + AQ_TM_QKV + win * 48, umma_desc_at(tmpl, base + AQ_OFF_T + s * AQ_TSTAGE + win * AQ_TWIN),
                                dw, id_qkv, 0u);
                umma_commit(qkv_full);
                umma_commit(&t_empty[s]);          // one of the five arrivals: the tile's MMA reads are done
            }
            __syncwarp();
        };
        uint32_t it = 0;
        if ((int)blockIdx.x < npairs) issue_qkv(0);
        for (int p = blockIdx.x; p < npairs; p += gridDim.x, ++it) {
            // S and the rel columns
            mbar_wait(qk_ready, it & 1);
            tc_fence_after();
            if (elect_one_sync()) {
                const uint64_t dr = umma_desc_at(tmpl, base + AQ_OFF_REL);
#pragma unroll
                for (int win = 0; win < 2; ++win) {
                    const uint64_t dq = umma_desc_at(tmpl, base + AQ_OFF_Q + win * 64 * AQ_ROWB);
                    const uint64_t dk = umma_desc_at(tmpl, base + AQ_OFF_K + win * AQ_KWIN);
                    umma_f16_ss(tmem_base + AQ_TM_S + win * AQ_LANE_B, dq, dk, id_s, 0u);
                    umma_f16_ss(tmem_base + AQ_TM_REL + win * AQ_LANE_B, dq, dr, id_r, 0u);
                }
                umma_commit(s_full);
                umma_commit(qk_free);
            }
            __syncwarp();
            if (p + (int)gridDim.x < npairs) issue_qkv(it + 1);
            // O = P . V
            mbar_wait(p_ready, it & 1);
            mbar_wait(o_empty, (it & 1) ^ 1);
            tc_fence_after();
            if (elect_one_sync()) {
                const uint64_t dv0 = umma_desc_at(tmpl, base + AQ_OFF_V);
                const uint64_t dp0 = umma_desc_at(tmpl_p, base + AQ_OFF_P);
#pragma unroll
                for (int win = 0; win < 2; ++win) {
#pragma unroll
                    for (int k = 0; k < AQ_WR / 16; ++k) {
                        const uint64_t dp = dp0 + (uint64_t)((win * 16384 + (k >> 2) * 8192 + (k & 3) * 32) >> 4);
                        const uint64_t dv = dv0 + (uint64_t)((win * AQ_KWIN + k * 16 * AQ_ROWB) >> 4);
                        umma_f16_ss(tmem_base + AQ_TM_O + win * AQ_LANE_B, dp, dv, id_o, k ? 1u : 0u);
                    }
                }
                umma_commit(o_full);
                umma_commit(v_free);
            }
            __syncwarp();
        }
    } else {
        // ---- compute warps --------------------------------------------------------------------------------
        // qkv phase: thread = tile row trow of BOTH windows.  Attention phase: thread owns TMEM lane 32*warp + lane
        // of the interleaved M = 64 accumulators: window (lane >> 4), query row 16*warp + (lane & 15).
        const int trow = warp * 32 + lane;
        const int thy = trow / WIN, thx = trow - thy * WIN;
        const bool t_used = trow < NKEY;
        const bool t_inner = t_used && thy >= 1 && thy <= BLK && thx >= 1 && thx <= BLK;
        const int t_q = (thy - 1) * BLK + (thx - 1);              // query index of an interior tile row
        const int win = lane >> 4, qi = warp * 16 + (lane & 15);
        const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
        uint8_t* prow = sm + AQ_OFF_P + win * 16384 + qi * 128;
        const uint32_t my_trow = (uint32_t)(((qi >> 3) + 1) * WIN + (qi & 7) + 1);   // this query's row in the t tile
        uint32_t it = 0;
        pdl_wait();
        for (int p = blockIdx.x; p < npairs; p += gridDim.x, ++it) {
            const uint32_t s = it & 1;
            // ---- q, k, v of this thread's tile row, both windows ------------------------------------------------
            mbar_wait(qkv_full, it & 1);
            tc_fence_after();
            uint32_t qa[48], qb[48];
            tmem_ld32(tmem_base + lane_sel + AQ_TM_QKV, qa);
            tmem_ld16(tmem_base + lane_sel + AQ_TM_QKV + 32, qa + 32);
            tmem_ld32(tmem_base + lane_sel + AQ_TM_QKV + 48, qb);
            tmem_ld16(tmem_base + lane_sel + AQ_TM_QKV + 80, qb + 32);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(qkv_empty);
            mbar_wait(qk_free, (it & 1) ^ 1);                      // the previous pair's MMAs have read Q, K ...
            mbar_wait(v_free, (it & 1) ^ 1);                       // ... and V
            if (t_used) {
#pragma unroll
                for (int wn = 0; wn < 2; ++wn) {
                    const uint32_t* r = wn ? qb : qa;
                    uint4 o[6];                                    // q0 q1 k0 k1 v0 v1 (16-byte chunks)
                    uint32_t* po = reinterpret_cast<uint32_t*>(o);
#pragma unroll
                    for (int e = 0; e < 24; ++e) {
                        const __half2 hv = __floats2half2_rn(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1]));
                        po[e] = *reinterpret_cast<const uint32_t*>(&hv);
                    }
                    uint8_t* kt = sm + AQ_OFF_K + wn * AQ_KWIN;
                    uint8_t* vt = sm + AQ_OFF_V + wn * AQ_KWIN;
                    *reinterpret_cast<uint4*>(kt + aq_sw32(trow, 0)) = o[2];
                    *reinterpret_cast<uint4*>(kt + aq_sw32(trow, 1)) = o[3];
                    *reinterpret_cast<uint4*>(vt + aq_sw32(trow, 0)) = o[4];
                    *reinterpret_cast<uint4*>(vt + aq_sw32(trow, 1)) = o[5];
                    if (t_inner) {
                        uint8_t* qt = sm + AQ_OFF_Q;
                        const uint32_t qr = (uint32_t)(wn * 64 + t_q);
                        *reinterpret_cast<uint4*>(qt + aq_sw32(qr, 0)) = o[0];
                        *reinterpret_cast<uint4*>(qt + aq_sw32(qr, 1)) = o[1];
                    }
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(qk_ready);

            // ---- softmax (identical to attn_umma_kernel) ---------------------------------------------------------
            mbar_wait(s_full, it & 1);
            tc_fence_after();
            float sv[104];
            uint32_t ab[24];
            {
                uint32_t* su = reinterpret_cast<uint32_t*>(sv);
                const uint32_t t0 = tmem_base + lane_sel + AQ_TM_S;
                tmem_ld32(t0, su);
                tmem_ld32(t0 + 32, su + 32);
                tmem_ld32(t0 + 64, su + 64);
                tmem_ld8(t0 + 96, su + 96);
                tmem_ld16(tmem_base + lane_sel + AQ_TM_REL, ab);
                tmem_ld8(tmem_base + lane_sel + AQ_TM_REL + 16, ab + 16);
                tmem_ld_wait();
            }
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
            uint32_t ph[52];
#pragma unroll
            for (int i = 0; i < NKEY / 2; ++i) {
                float a0, a1;
                f2_unpack(f2_fma(s2[i], l2e, nmx), a0, a1);
                const float e0 = aq_exp2(a0), e1 = aq_exp2(a1);
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
#pragma unroll
            for (int q = 0; q < 13; ++q)
                *reinterpret_cast<uint4*>(prow + (q >> 3) * 8192 + (((q & 7) ^ (qi & 7)) << 4)) =
                    make_uint4(ph[4 * q], ph[4 * q + 1], ph[4 * q + 2], ph[4 * q + 3]);
            const float inv = 1.f / sum;
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_ready);

            // ---- fused branch glue: y_1 = O' + t_1 -> Y;  t_2 = n_2/2 + y_1/2 in place (ref :139-141) ----------------
            const int wi = 2 * p + win;
            const bool valid = wi < nwin;
            const AqCoord wc = aq_coord(valid ? wi : 2 * p, nwx, per_img);
            const int ly = wc.y + (qi >> 3), lx = wc.x + (qi & 7);
            const long pix = ((long)wc.b * fz.Hp + ly) * fz.Wp + lx;
            const long toff = ((((long)wc.b * (fz.Hp >> 1)) + (ly >> 1)) * (fz.Wp >> 1) + (lx >> 1)) * 64 + ((ly & 1) * 2 + (lx & 1)) * NB;
            const bool has_next = fz.Tnext != nullptr;
            const uint8_t* tt = sm + AQ_OFF_T + s * AQ_TSTAGE + win * AQ_TWIN;
            uint4 tk[2], tlo[2], hc[2];
            tk[0] = *reinterpret_cast<const uint4*>(tt + aq_sw32(my_trow, 0));
            tk[1] = *reinterpret_cast<const uint4*>(tt + aq_sw32(my_trow, 1));
            if constexpr (LO) ldg256(fz.Tlo + pix * NB, tlo[0], tlo[1]);
            if (has_next) ldg256(fz.Tnext + toff, hc[0], hc[1]);
            mbar_wait(o_full, it & 1);
            tc_fence_after();
            uint32_t r[16];
            tmem_ld16(tmem_base + lane_sel + AQ_TM_O, r);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(o_empty); mbar_arrive(&t_empty[s]); }
            if (valid) {
                float yv[NB];
                const __half2* th = reinterpret_cast<const __half2*>(tk);
                const __half2* tl = reinterpret_cast<const __half2*>(tlo);
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
                stg256(fz.Y + pix * NF, yo[0], yo[1]);             // branch 1 = channels 0..15 of Y
                if constexpr (LO) {
                    uint4 yl[2];
                    __half2* ylh = reinterpret_cast<__half2*>(yl);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float2 yr = __half22float2(yh[e]);
                        ylh[e] = __floats2half2_rn((yv[2 * e] - yr.x) * 2048.f, (yv[2 * e + 1] - yr.y) * 2048.f);
                    }
                    stg256(fz.Ylo + pix * NF, yl[0], yl[1]);
                }
                if (has_next) {
                    uint4 to[2], tol[2];
                    __half2* tnh = reinterpret_cast<__half2*>(to);
                    __half2* tnl = reinterpret_cast<__half2*>(tol);
                    const __half2* hh = reinterpret_cast<const __half2*>(hc);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float2 hf = __half22float2(hh[e]);
                        const float t0 = fmaf(0.5f, yv[2 * e], hf.x), t1 = fmaf(0.5f, yv[2 * e + 1], hf.y);
                        tnh[e] = __floats2half2_rn(t0, t1);
                        const float2 tr = __half22float2(tnh[e]);
                        tnl[e] = __floats2half2_rn(t0 - tr.x, t1 - tr.y);
                    }
                    stg256(fz.Tnext + toff, to[0], to[1]);
                    if constexpr (LO) stg256(fz.Tnext_lo + toff, tol[0], tol[1]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, 256);
}

}  // namespace

// T: t_1 fp16 [B,h,w,16]; Wqkv: fp16 [48][16] (q rows pre-scaled); relx: fp16 [32][16]; fz as for launch_attn_umma
int launch_attn16_qkv(const __half* T, const __half* Wqkv, const __half* relx, int B, int h, int w, cudaStream_t s,
                      const AttnFuse& fz) {
    if (h % BLK || w % BLK) { set_error("attn16_qkv: %dx%d is not a multiple of the 8x8 block", h, w); return M2T_E_ARG; }
    CUtensorMap mapT, mapW, mapR;
    {
        const uint64_t dims[4] = {AQ_C, (uint64_t)w, (uint64_t)h, (uint64_t)B};
        const uint64_t str[4] = {2, AQ_C * 2, (uint64_t)w * AQ_C * 2, (uint64_t)h * w * AQ_C * 2};
        const uint32_t box[4] = {AQ_C, WIN, WIN, 1};
        M2T_TRY(make_tensor_map(&mapT, T, 2, 4, dims, str, box, 1));
    }
    {
        const uint64_t dims[2] = {AQ_C, 48}, str[2] = {2, AQ_C * 2};
        const uint32_t box[2] = {AQ_C, 48};
        M2T_TRY(make_tensor_map(&mapW, Wqkv, 2, 2, dims, str, box, 1));
    }
    {
        const uint64_t dims[2] = {AQ_C, 32}, str[2] = {2, AQ_C * 2};
        const uint32_t box[2] = {AQ_C, 32};
        M2T_TRY(make_tensor_map(&mapR, relx, 2, 2, dims, str, box, 1));
    }
    const int nwin = B * (h / BLK) * (w / BLK);
    const int npairs = (nwin + 1) / 2;
    const int cap = device_sm_count() * 2;
    const int grid = npairs < cap ? npairs : cap;
    if (fz.Tlo != nullptr) {
        M2T_ENSURE_SMEM(attn16_qkv_kernel<true>, AQ_SMEM);
        M2T_CUDA(launch_pdl(attn16_qkv_kernel<true>, dim3(grid), dim3(192), AQ_SMEM, s, mapT, mapW, mapR, h, w, nwin, fz));
    } else {
        M2T_ENSURE_SMEM(attn16_qkv_kernel<false>, AQ_SMEM);
        M2T_CUDA(launch_pdl(attn16_qkv_kernel<false>, dim3(grid), dim3(192), AQ_SMEM, s, mapT, mapW, mapR, h, w, nwin, fz));
    }
    return M2T_OK;
}

}  // namespace m2t
