// Branches 2-4 of a CFTM in ONE kernel each (ref M2Trans_network.py:143-161, :307-332): qkv 1x1 conv + halo attention
// + branch glue, for C = 64 (level 1) and C = 256 (level 2), WITHOUT ever forming q, k or v.
//
// The qkv conv has no bias (ref :281) and every step up to the softmax is linear in the branch input z = t_k
// (space-to-depth, Haar transforms folded into the weights by pack.cu), so the contractions re-associate:
//     S[i][j] = (Wq z_i) . (Wk z_j + rel_j) = z_i^T (Wq^T Wk) z_j + z_i^T (Wq^T rel_j)       MQ = [Wq^T rel | Wq^T Wk]
//     O[i]    = sum_j P[i][j] (Wv z_j)      = Wv (sum_j P[i][j] z_j)
// Per query the kernel computes A = z_i MQ (one GEMM, K = C), S = A . Z^T against the RAW halo tile of t_k,
// PZ = P . Z (the tile again, as an MN-major operand) and O = PZ . Wv^T.  Compared with qkv_umma + attn_umma this
//   * removes the QKV tensor (96 B/px written + ~150 B/px read per branch) and one launch per branch,
//   * needs no halo recompute: keys and values ARE the tile (zero outside the frame = the reference's zero-padded
//     unfold of k and v, ref :313-317; the rel terms are still added at padded keys, ref :322-325),
//   * costs FEWER MMA flops than the two-kernel form (per window 2*64*C*(C+32+100+100+C) vs 2*64*C*(3C+200)).
// Rounding points (fp16 operands, fp32 accumulate): MQ and Wv once at pack time, A, P (unnormalised exp) and
// PZ/sum; emulated on the CPU against the fp64 reference this is as accurate as the q/k/v form or slightly better.
//
// Work item = a VERTICAL pair of 8x8 windows = one M = 128 accumulator (row m = 8*(y-1) + (x-1) of the 18 x 10 tile).
//   tile     TMA box 10 (x) x 18 (y) x 64 ch per 64-channel chunk, rows t = 10 y + x of 128 B (128-byte swizzle);
//            rows 180..191 of every chunk stay zero (N = 192 key columns per MMA)
//   phase 1  [QR | A] = Zq . MQ^T      A operand = the tile itself: start at row 11, 8-row groups 10 rows apart
//                                      (SBO 1280 B, the ff conv's centre tap; tests/test_probes.py pins it)
//   phase 2  A -> fp16 -> K-major operand tile AQ in shared memory (epilogue warps, thread = query row)
//   phase 3  S = AQ . Z^T              N = 192: window A uses key columns 0..99, window B 80..179
//   phase 4  softmax thread-per-row from TMEM (+ q.rel terms from QR), P fp16 [128][192], zero outside the own window
//   phase 5  PZ = P . Z                B = the tile as MN-major operand, N = C in one instruction (LBO = chunk pitch)
//   phase 6  PZ / sum -> fp16 -> operand tile (same shared memory as AQ / P)
//   phase 7  O = PZ . Wv^T
//   phase 8  fused branch glue of attn_umma_kernel (y_k into Y, t_{k+1} completed in place); t_k rows of the
//            residual add come from the tile
// C = 256 streams MQ and Wv through a 4-slot ring of [128 rows][64 ch] boxes (278 KB per pair from L2, issued ahead
// of the tile so the first boxes land while the previous kernel drains); C = 64 keeps both resident (20 KB) and runs
// two CTAs per SM.  TMEM: QR [0,32) | A / PZ / O [32,32+C) | S (C = 64: aliases A, [32,224); C = 256: [288,480)).
// Warps: 0-3 epilogue / softmax (thread = query row = TMEM lane), 4 TMA, 5 MMA, C = 256 only: 6-9 share the
// conversions and the final epilogue of the rows of warps 0-3.
#include "common.cuh"
#include "gelu.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace m2t {
namespace {

constexpr int AZ_KEYS = 192;                           // key rows of a tile chunk as an MMA operand (180 loaded + 12 zero)
constexpr int AZ_TROWS = 180;                          // 18 x 10 pixels
constexpr uint32_t AZ_CHUNK = AZ_KEYS * 128;           // 24576 B: one 64-channel chunk of the tile
constexpr uint32_t AZ_QOFF = 11 * 128;                 // first query row of the tile (y = 1, x = 1)
constexpr uint32_t AZ_OPCH = 128 * 128;                // 16384 B: one 64-column chunk of a 128-row K-major operand
constexpr uint32_t AZ_SLOT = 16384;                    // ring slot: 128 weight rows x 64 channels

template <int C>
struct AzCfg {
    static constexpr int NBLK = C / 64;
    static constexpr bool RING = C == 256;
    static constexpr bool SPLIT = C == 256;            // second epilogue warpgroup
    static constexpr int NQ = C + 32;                  // MQ rows: 32 rel rows (20 used), then C rows of Wq^T Wk
    static constexpr int NSLOT = 4;
    static constexpr uint32_t OFF_TILE = 0;
    static constexpr uint32_t OFF_OPER = NBLK * AZ_CHUNK;
    static constexpr uint32_t OPER_BYTES = (NBLK > 3 ? NBLK : 3) * AZ_OPCH;       // AQ / PZ: NBLK chunks, P: 3 chunks
    static constexpr uint32_t OFF_W = OFF_OPER + OPER_BYTES;
    static constexpr uint32_t MQ_BYTES = NQ * 128;                                 // resident form (C = 64)
    static constexpr uint32_t W_BYTES = RING ? NSLOT * AZ_SLOT : (MQ_BYTES + C * 128);
    static constexpr uint32_t OFF_BAR = OFF_W + W_BYTES;
    static constexpr uint32_t OFF_INV = OFF_BAR + 256;
    static constexpr uint32_t SMEM = 1024 + OFF_BAR + 256 + 512;
    static constexpr int THREADS = SPLIT ? 320 : 192;
    static constexpr int NEPI = SPLIT ? 8 : 4;         // epilogue warps
    static constexpr int MIN_CTAS = C == 64 ? 2 : 1;
    static constexpr uint32_t TM_QR = 0, TM_A = 32;
    static constexpr uint32_t TM_S = C == 64 ? 32 : 288;
    static constexpr uint32_t TM_COLS = C == 64 ? 256 : 512;
    static constexpr uint32_t TX_TILE = NBLK * AZ_TROWS * 128;
};
static_assert(AzCfg<256>::SMEM <= 232448, "attn_z<256> exceeds the shared memory of an SM");
static_assert(2 * AzCfg<64>::SMEM <= 232448, "attn_z<64> must fit twice");

__device__ __forceinline__ float az_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct AzDiv {                                          // division by a runtime constant, see attn_umma.cu
    uint32_t m, s1, s2, d;
    __device__ __forceinline__ explicit AzDiv(uint32_t div) : d(div) {
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
// pair p -> image b, first row of the upper window (level pixels), first column
struct AzPair { int b, y, x; };
__device__ __forceinline__ AzPair az_pair(int p, const AzDiv& nwx, const AzDiv& per_img, int ystep) {
    AzPair c;
    c.b = (int)per_img.div((uint32_t)p);
    const uint32_t r = (uint32_t)p - (uint32_t)c.b * per_img.d;
    const uint32_t py = nwx.div(r);
    c.y = (int)py * ystep;
    c.x = (int)(r - py * nwx.d) * BLK;
    return c;
}

__device__ __forceinline__ void az_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// TMEM [this warp's 32 lanes][64 columns] fp32 (times scale) -> fp16 row m of a 128-byte-swizzled K-major chunk
__device__ __forceinline__ void az_cvt_chunk(uint32_t taddr, uint8_t* chunk, int m, float scale) {
    uint32_t r[64];
    tmem_ld32(taddr, r);
    tmem_ld32(taddr + 32, r + 32);
    tmem_ld_wait();
    uint8_t* row = chunk + m * 128;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        uint4 u;
        uint32_t* pu = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const __half2 hv = __floats2half2_rn(__uint_as_float(r[8 * q + 2 * e]) * scale, __uint_as_float(r[8 * q + 2 * e + 1]) * scale);
            pu[e] = *reinterpret_cast<const uint32_t*>(&hv);
        }
        *reinterpret_cast<uint4*>(row + ((q ^ (m & 7)) << 4)) = u;
    }
}

#ifdef M2T_TIMING
__device__ long long g_az_dbg[3 * 64];     // per branch 2..4: CTA 0, epilogue thread 0: stamps [10 it + slot], it < 6; [60], [61] prologue
#define M2T_ZT(slot) do { if (blockIdx.x == 0 && tid == 0 && it < 6) g_az_dbg[64 * (fz.branch - 1) + (slot) + 10 * it] = clock64(); } while (0)
#else
#define M2T_ZT(slot) do { } while (0)
#endif

template <int C, bool LO>
__global__ void __launch_bounds__(AzCfg<C>::THREADS, AzCfg<C>::MIN_CTAS)
attn_z_kernel(const __grid_constant__ CUtensorMap mapT, const __grid_constant__ CUtensorMap mapMQ,
              const __grid_constant__ CUtensorMap mapWV, const __grid_constant__ CUtensorMap mapTlo,
              const __grid_constant__ CUtensorMap mapTn, const __grid_constant__ CUtensorMap mapTnlo, int h, int w, int npairs,
              int single, const AttnFuse fz) {
    using CF = AzCfg<C>;
    constexpr int NBLK = CF::NBLK;
    constexpr bool RING = CF::RING;
    // Precise mode, C = 256: the glue's 32-byte pieces of Tlo / Tnext / Tnext_lo (one LSU sector per lane and 16-byte pass,
    // ~2 cycles each: 23 K cycles per pair in branch 3) travel as TMA boxes of whole 512-byte pixel rows instead.  After the
    // O MMAs the operand region and the weight ring are idle: the t_k residual tile lands in the former, the n_{k+1}/2 tile in
    // the latter, the epilogue threads update both in place (t_{k+1} and its residual) and one thread stores them.  The weight
    // ring then cannot run ahead into the next pair (branch 3 only; branch 4 has no next tensor and keeps the ring).
    // Fast mode, branch 3: the same for the n_4/2 tile alone, staged in the operand region (the ring keeps running ahead).
    // C = 64 (branch 2): its next tensor lives one level up, a window covers 4 x 4 of its pixels completely: boxes of
    // {64 ch, 4, 4} per 64-channel chunk, all staged in the operand region (48 KB: n_3/2 -> t_3 | t_2 residual | t_3 residual).
    // Precise mode only: -13 % for the kernel at cfg3.  In fast mode the C = 64 glue is 2 K cycles of a 12 K-cycle chain and
    // the exposed TMA round trip plus the extra block barrier cost more than the LSU pieces (169 vs 150 us at cfg2): not used.
    constexpr bool STG = true;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + CF::OFF_BAR);
    uint64_t* w_full = bars;                 // resident weights landed (C = 64)
    uint64_t* r_full = bars + 1;             // [4] ring slot landed
    uint64_t* r_empty = bars + 5;            // [4] ... consumed by the MMAs
    uint64_t* tile_full = bars + 9;
    uint64_t* a_full = bars + 10;            // [QR | A] complete in TMEM
    uint64_t* aq_ready = bars + 11;          // AQ written to smem
    uint64_t* s_full = bars + 12;
    uint64_t* p_ready = bars + 13;           // P written, S and QR consumed
    uint64_t* pz_full = bars + 14;
    uint64_t* pzs_ready = bars + 15;         // PZ / sum written to smem
    uint64_t* o_full = bars + 16;
    uint64_t* pair_done = bars + 17;         // O consumed and the tile no longer needed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
    uint64_t* g_full = bars + 20;            // STG: staged tiles landed
    uint64_t* stage_free = bars + 21;        // STG: the stores have read the ring region
    float* sinv = reinterpret_cast<float*>(sm + CF::OFF_INV);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef M2T_TIMING
    if (blockIdx.x == 0 && tid == 0) g_az_dbg[64 * (fz.branch - 1) + 60] = clock64();
#endif
    // single != 0 (small inputs, see launch_attn_z): every window is the upper window of its own pair, the lower one is a
    // phantom like the one below an odd last row -- its rows are computed and dropped
    const int nwx_i = w / BLK, nwy = h / BLK, npy = single ? nwy : (nwy + 1) / 2, ystep = single ? BLK : 2 * BLK;
    const AzDiv nwx((uint32_t)nwx_i), per_img((uint32_t)(npy * nwx_i));

    // zero the 12 padding key rows of every tile chunk (TMA never writes them; the S and PZ MMAs read them)
    for (int kb = 0; kb < NBLK; ++kb)
        for (uint32_t i = tid * 16; i < (AZ_KEYS - AZ_TROWS) * 128; i += CF::THREADS * 16)
            *reinterpret_cast<uint4*>(sm + CF::OFF_TILE + kb * AZ_CHUNK + AZ_TROWS * 128 + i) = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    if (warp == 5) tmem_alloc(tmem_slot, CF::TM_COLS);
    if (tid == 128) {
        mbar_init(w_full, 1);
        for (int s = 0; s < CF::NSLOT; ++s) { mbar_init(&r_full[s], 1); mbar_init(&r_empty[s], 1); }
        mbar_init(tile_full, 1);
        mbar_init(a_full, 1);
        mbar_init(aq_ready, CF::NEPI);
        mbar_init(s_full, 1);
        mbar_init(p_ready, 4);
        mbar_init(pz_full, 1);
        mbar_init(pzs_ready, CF::NEPI);
        mbar_init(o_full, 1);
        mbar_init(pair_done, CF::NEPI);
        if constexpr (STG) { mbar_init(g_full, 1); mbar_init(stage_free, 1); }
        mbar_fence_init();
        tma_prefetch_desc(&mapT);
        tma_prefetch_desc(&mapMQ);
        tma_prefetch_desc(&mapWV);
        if constexpr (STG) { tma_prefetch_desc(&mapTlo); tma_prefetch_desc(&mapTn); tma_prefetch_desc(&mapTnlo); }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
#ifdef M2T_TIMING
    if (blockIdx.x == 0 && tid == 0) g_az_dbg[64 * (fz.branch - 1) + 61] = clock64();
#endif

    if (warp == 4) {
        // ---- TMA producer ------------------------------------------------------------------------------------
        uint32_t g = 0;                                   // ring boxes issued so far
        auto ring_load = [&](const CUtensorMap* map, int kb, int row0) {
            const uint32_t s = g % CF::NSLOT, ph = (g / CF::NSLOT) & 1;
            mbar_wait(&r_empty[s], ph ^ 1);
            if (elect_one_sync()) {
                mbar_expect_tx(&r_full[s], AZ_SLOT);
                tma_load_2d(sm + CF::OFF_W + s * AZ_SLOT, map, &r_full[s], kb * 64, row0);
            }
            __syncwarp();
            ++g;
        };
        // MQ box i = 0..11 is (K chunk i / 3, 128-row slab i % 3).  (Four extra slots in the operand region, idle during
        // phase 1, were tried: eight boxes in flight instead of four did not shorten the phase and were removed again.  The
        // phase is not bound by the L2 port -- tools/probes/tma_stream_probe.cu: 64 B/clk per SM with >= 2 boxes in flight,
        // alone or with 148 SMs pulling -- but by shared-memory bandwidth: per 16 KB box TMA writes 16 KB and the four N = 128
        // MMAs read 16 KB of weights + 16 KB of query rows = 384 cycles at 128 B/clk, which is the ~32 B/clk observed.)
        auto mq_box = [&](int i) { ring_load(&mapMQ, i / 3, (i % 3) * 128); };
        if constexpr (!RING) {
            if (elect_one_sync()) {                      // constants: loaded while the previous kernel drains
                mbar_expect_tx(w_full, CF::W_BYTES);
                tma_load_2d(sm + CF::OFF_W, &mapMQ, w_full, 0, 0);
                tma_load_2d(sm + CF::OFF_W + CF::MQ_BYTES, &mapWV, w_full, 0, 0);
            }
            __syncwarp();
        }
        uint32_t it = 0;
        for (int p = blockIdx.x; p < npairs; p += gridDim.x, ++it) {
            const AzPair pc = az_pair(p, nwx, per_img, ystep);
            // the first four weight boxes do not wait for this pair's MMAs (the ring has four slots), the tile does
            // not wait for the ring: issuing in this order cannot deadlock and lets the weights run ahead
            if constexpr (RING && LO) {
                if (fz.Tnext != nullptr && it > 0) mbar_wait(stage_free, (it - 1) & 1);   // the ring region held the n/2 tile
            }
            if constexpr (RING) { mq_box(0); mq_box(1); mq_box(2); mq_box(3); }
            if (it == 0) pdl_wait();
            mbar_wait(pair_done, (it & 1) ^ 1);
            if (elect_one_sync()) {
                mbar_expect_tx(tile_full, CF::TX_TILE);
                for (int kb = 0; kb < NBLK; ++kb)
                    tma_load_4d(sm + CF::OFF_TILE + kb * AZ_CHUNK, &mapT, tile_full, kb * 64, pc.x - 1, pc.y - 1, pc.b);
            }
            __syncwarp();
            if constexpr (RING) {
                for (int i = 4; i < 3 * NBLK; ++i) mq_box(i);
                for (int i = 0; i < 2 * NBLK; ++i) ring_load(&mapWV, i / 2, (i % 2) * 128);
            }
        }
    } else if (warp == 5) {
        // ---- MMA issuer: warp-uniform loops, one elected lane issues -------------------------------------------
        constexpr uint64_t tmpl = umma_smem_desc(0, 16, 1024, UMMA_LAYOUT_SW128);           // K-major, 8-row groups dense
        constexpr uint64_t tmpl_q = umma_smem_desc(0, 16, 1280, UMMA_LAYOUT_SW128);         // query rows of the tile
        constexpr uint64_t tmpl_z = umma_smem_desc(0, AZ_CHUNK, 1024, UMMA_LAYOUT_SW128);   // tile as MN-major B
        constexpr uint32_t id_s = umma_idesc_f16(128, AZ_KEYS);
        constexpr uint32_t id_pz = umma_idesc_f16(128, C, 0, 1);
        if constexpr (!RING) mbar_wait(w_full, 0);
        uint32_t g = 0, it = 0;
        for (int p = blockIdx.x; p < npairs; p += gridDim.x, ++it) {
            mbar_wait(tile_full, it & 1);
            mbar_wait(pair_done, (it & 1) ^ 1);            // the previous pair's O has been read: [0, 32 + C) is free
            tc_fence_after();
            // phase 1: [QR | A] = Zq . MQ^T
            if constexpr (RING) {
                for (int i = 0; i < 3 * NBLK; ++i, ++g) {
                    const int kb = i / 3, sl = i % 3;
                    const uint32_t s = g % CF::NSLOT, ph = (g / CF::NSLOT) & 1;
                    mbar_wait(&r_full[s], ph);
                    tc_fence_after();
                    if (elect_one_sync()) {
                        const uint64_t da0 = umma_desc_at(tmpl_q, base + CF::OFF_TILE + kb * AZ_CHUNK + AZ_QOFF);
                        const uint64_t db0 = umma_desc_at(tmpl, base + CF::OFF_W + s * AZ_SLOT);
                        const uint32_t idesc = sl < 2 ? umma_idesc_f16(128, 128) : umma_idesc_f16(128, 32);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_f16_ss(tmem_base + sl * 128, da0 + 2 * k, db0 + 2 * k, idesc, (kb | k) ? 1u : 0u);
                        umma_commit(&r_empty[s]);
                        if (i == 3 * NBLK - 1) umma_commit(a_full);
                    }
                    __syncwarp();
                }
            } else {
                if (elect_one_sync()) {
                    const uint64_t da0 = umma_desc_at(tmpl_q, base + CF::OFF_TILE + AZ_QOFF);
                    const uint64_t db0 = umma_desc_at(tmpl, base + CF::OFF_W);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_f16_ss(tmem_base, da0 + 2 * k, db0 + 2 * k, umma_idesc_f16(128, CF::NQ), k ? 1u : 0u);
                    umma_commit(a_full);
                }
                __syncwarp();
            }
            // phase 3: S = AQ . Z^T over all 192 key rows of the tile
            mbar_wait(aq_ready, it & 1);
            tc_fence_after();
            if (elect_one_sync()) {
#pragma unroll
                for (int kb = 0; kb < NBLK; ++kb) {
                    const uint64_t da0 = umma_desc_at(tmpl, base + CF::OFF_OPER + kb * AZ_OPCH);
                    const uint64_t db0 = umma_desc_at(tmpl, base + CF::OFF_TILE + kb * AZ_CHUNK);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_f16_ss(tmem_base + CF::TM_S, da0 + 2 * k, db0 + 2 * k, id_s, (kb | k) ? 1u : 0u);
                }
                umma_commit(s_full);
            }
            __syncwarp();
            // phase 5: PZ = P . Z (12 key steps of 16; the tile is the MN-major B operand, all C columns at once)
            mbar_wait(p_ready, it & 1);
            tc_fence_after();
            if (elect_one_sync()) {
                const uint64_t dp0 = umma_desc_at(tmpl, base + CF::OFF_OPER);
                const uint64_t dz0 = umma_desc_at(tmpl_z, base + CF::OFF_TILE);
#pragma unroll
                for (int k = 0; k < AZ_KEYS / 16; ++k) {
                    const uint64_t dp = dp0 + (uint64_t)(((k >> 2) * AZ_OPCH + (k & 3) * 32) >> 4);
                    const uint64_t dz = dz0 + (uint64_t)((k * 16 * 128) >> 4);
                    umma_f16_ss(tmem_base + CF::TM_A, dp, dz, id_pz, k ? 1u : 0u);
                }
                umma_commit(pz_full);
            }
            __syncwarp();
            // phase 7: O = (PZ / sum) . Wv^T
            mbar_wait(pzs_ready, it & 1);
            tc_fence_after();
            if constexpr (RING) {
                for (int i = 0; i < 2 * NBLK; ++i, ++g) {
                    const int kb = i / 2, sl = i % 2;
                    const uint32_t s = g % CF::NSLOT, ph = (g / CF::NSLOT) & 1;
                    mbar_wait(&r_full[s], ph);
                    tc_fence_after();
                    if (elect_one_sync()) {
                        const uint64_t da0 = umma_desc_at(tmpl, base + CF::OFF_OPER + kb * AZ_OPCH);
                        const uint64_t db0 = umma_desc_at(tmpl, base + CF::OFF_W + s * AZ_SLOT);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_f16_ss(tmem_base + CF::TM_A + sl * 128, da0 + 2 * k, db0 + 2 * k, umma_idesc_f16(128, 128), (kb | k) ? 1u : 0u);
                        umma_commit(&r_empty[s]);
                        if (i == 2 * NBLK - 1) umma_commit(o_full);
                    }
                    __syncwarp();
                }
            } else {
                if (elect_one_sync()) {
                    const uint64_t da0 = umma_desc_at(tmpl, base + CF::OFF_OPER);
                    const uint64_t db0 = umma_desc_at(tmpl, base + CF::OFF_W + CF::MQ_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_f16_ss(tmem_base + CF::TM_A, da0 + 2 * k, db0 + 2 * k, umma_idesc_f16(128, C), k ? 1u : 0u);
                    umma_commit(o_full);
                }
                __syncwarp();
            }
        }
    } else {
        // ---- epilogue warps: thread = accumulator row m = TMEM lane; window m >> 6, query (m & 63) ---------------
        const int quad = warp & 3;
        const bool helper = CF::SPLIT && warp >= 6;
        const int m = quad * 32 + lane;
        const int win = m >> 6, qi = m & 63;
        const uint32_t lane_sel = (uint32_t)(quad * 32) << 16;
        const uint32_t trow = (uint32_t)(((qi >> 3) + 1 + BLK * win) * WIN + (qi & 7) + 1);   // this query's tile row
        // chunks of the conversions this thread handles
        constexpr int CH_N = CF::SPLIT ? NBLK / 2 : NBLK;
        const int ch0 = helper ? NBLK / 2 : 0;
        uint8_t* oper = sm + CF::OFF_OPER;
        uint32_t it = 0;
        pdl_wait();
        for (int p = blockIdx.x; p < npairs; p += gridDim.x, ++it) {
            // ---- phase 2: A -> fp16 operand tile ---------------------------------------------------------------------
            M2T_ZT(0);
            mbar_wait(a_full, it & 1);
            if constexpr (RING != LO) {      // the previous pair's t_{k+1} tile(s) have left the operand region
                if (fz.Tnext != nullptr && it > 0) mbar_wait(stage_free, (it - 1) & 1);
            }
            tc_fence_after();
            M2T_ZT(1);
#pragma unroll 1
            for (int c = ch0; c < ch0 + CH_N; ++c)
                az_cvt_chunk(tmem_base + lane_sel + CF::TM_A + c * 64, oper + c * AZ_OPCH, m, 1.f);
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(aq_ready);
            M2T_ZT(2);

            // ---- phase 4: softmax over this window's 100 keys ------------------------------------------------------------
            float inv;
            if (!helper) {
                mbar_wait(s_full, it & 1);
                tc_fence_after();
                M2T_ZT(3);
                float sv[104];
                uint32_t ab[24];
                {
                    uint32_t* su = reinterpret_cast<uint32_t*>(sv);
                    const uint32_t t0 = tmem_base + lane_sel + CF::TM_S + (uint32_t)(80 * win);
                    tmem_ld32(t0, su);
                    tmem_ld32(t0 + 32, su + 32);
                    tmem_ld32(t0 + 64, su + 64);
                    tmem_ld8(t0 + 96, su + 96);
                    tmem_ld16(tmem_base + lane_sel + CF::TM_QR, ab);
                    tmem_ld8(tmem_base + lane_sel + CF::TM_QR + 16, ab + 16);
                    tmem_ld_wait();
                }
                // two keys per instruction on the packed fp32x2 pipe, as in attn_umma_kernel: keys 2i and 2i+1 share
                // their key row (rel_h term); key j of the window is tile pixel (j / 10, j % 10)
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
                uint32_t ph[52];                               // P row as fp16 pairs; pairs 50, 51 = zeros
#pragma unroll
                for (int i = 0; i < NKEY / 2; ++i) {
                    float a0, a1;
                    f2_unpack(f2_fma(s2[i], l2e, nmx), a0, a1);
                    const float e0 = az_exp2(a0), e1 = az_exp2(a1);
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
                // P row m: 24 chunks of 8 keys over the pair's 192 key columns; the own window's 100 keys start at
                // column 80 * win (chunk 10 * win), everything else is written as zeros (the buffer held AQ before)
                uint8_t* prow = oper + m * 128;
                auto pstore = [&](int gq, const uint4& v) {
                    *reinterpret_cast<uint4*>(prow + (gq >> 3) * AZ_OPCH + (((gq & 7) ^ (m & 7)) << 4)) = v;
                };
                const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
                if (win == 0) {
#pragma unroll
                    for (int q = 0; q < 13; ++q) pstore(q, make_uint4(ph[4 * q], ph[4 * q + 1], ph[4 * q + 2], ph[4 * q + 3]));
#pragma unroll
                    for (int gq = 13; gq < 24; ++gq) pstore(gq, z4);
                } else {
#pragma unroll
                    for (int gq = 0; gq < 10; ++gq) pstore(gq, z4);
#pragma unroll
                    for (int q = 0; q < 13; ++q) pstore(q + 10, make_uint4(ph[4 * q], ph[4 * q + 1], ph[4 * q + 2], ph[4 * q + 3]));
                    pstore(23, z4);
                }
                inv = 1.f / sum;
                if constexpr (CF::SPLIT) sinv[m] = inv;
                fence_proxy_async();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(p_ready);
                M2T_ZT(4);
            } else {
                mbar_wait(p_ready, it & 1);                 // the primaries have published 1/sum of this pair
                inv = sinv[m];
            }

            // ---- glue set-up, issued early: the n_{k+1}/2 segments of Tnext (and the first t_k residuals) depend only on
            //      earlier kernels, so their loads go out now and land under the PZ / O MMAs (issued just before the O
            //      accumulator was read they cost 1-2 K cycles per 64-channel block, 6-10 K per pair in branch 3) -----------
            // With the Haar transforms folded into the weights the accumulator row IS IWT^L(attention) in space-to-depth
            // order: column s*16+k belongs to pixel s = dy*2^L + dx of this level pixel's 2^L x 2^L block.
            //   y_k = O' + t_k -> Y[..., 16k..16k+15];   t_{k+1} = n_{k+1}/2 (pre-filled) + y_k/2 -> Tnext in place
            const AzPair pc = az_pair(p, nwx, per_img, ystep);
            const int wy = pc.y + BLK * win;                  // first row of this thread's window
            const bool valid = wy < h && !(single && win);    // false for a phantom lower window (odd last row, single mode)
            constexpr int LV = C == 64 ? 1 : 2;
            constexpr int S = 1 << LV;
            constexpr int SPB = 4;                            // 16-channel sub-pixels per 64-channel block
            const int ly = wy + (qi >> 3), lx = pc.x + (qi & 7);
            const int br = fz.branch;
            const bool has_next = fz.Tnext != nullptr;
            constexpr int lvn = 2, Sn = 4, Cn = NB * Sn * Sn;   // the next branch is always a level-2 branch
            auto tnext_off = [&](int s) -> long {
                const int fy = ly * S + s / S, fx = lx * S + s % S;
                const int sn = (fy & (Sn - 1)) * Sn + (fx & (Sn - 1));
                return ((((long)pc.b * (fz.Hp >> lvn)) + (fy >> lvn)) * (fz.Wp >> lvn) + (fx >> lvn)) * Cn + sn * NB;
            };
            const long trow_off = (((long)pc.b * h + ly) * w + lx) * C;
            constexpr int JN = CF::SPLIT ? SPB / 2 : SPB;     // sub-pixels of each block this thread handles
            const int j0 = helper ? SPB / 2 : 0;
            uint4 hcur[2 * JN], hnxt[2 * JN];                 // n_{k+1}/2 segments of the current / next block
            auto load_h = [&](int nb, uint4* dst) {
#pragma unroll
                for (int j = 0; j < JN; ++j) ldg256(fz.Tnext + tnext_off(nb * SPB + j0 + j), dst[2 * j], dst[2 * j + 1]);
            };
            uint4 lcur[2 * JN], lnxt[2 * JN];                 // rounding residuals of t_k (precise mode), one block ahead
            auto load_l = [&](int nb, uint4* dst) {
#pragma unroll
                for (int j = 0; j < JN; ++j) {
                    if (LO) ldg256(fz.Tlo + trow_off + (nb * SPB + j0 + j) * NB, dst[2 * j], dst[2 * j + 1]);
                    else { dst[2 * j] = make_uint4(0u, 0u, 0u, 0u); dst[2 * j + 1] = make_uint4(0u, 0u, 0u, 0u); }
                }
            };
            if (valid) {
                // Tnext was written by branch_prep_all several kernels ago and has usually left L2: pull every segment
                // this thread will touch back into L2 now (no registers held), block 0 also into registers
                const bool staged = RING ? (LO || has_next) : LO;   // this pair's tiles travel by TMA (phase 8)
                if (has_next) {
#pragma unroll
                    for (int nb = 0; nb < NBLK; ++nb)
#pragma unroll
                        for (int j = 0; j < JN; ++j)
                            if (staged || nb > 0) az_prefetch_l2(fz.Tnext + tnext_off(nb * SPB + j0 + j));
                    if (!staged) load_h(0, hcur);
                }
                if constexpr (LO) {
#pragma unroll
                    for (int nb = 0; nb < NBLK; ++nb)
#pragma unroll
                        for (int j = 0; j < JN; ++j) az_prefetch_l2(fz.Tlo + trow_off + (nb * SPB + j0 + j) * NB);
                } else {
                    if (!staged) load_l(0, lcur);                 // zeros for the unstaged loop below
                }
            }

            // ---- phase 6: PZ / sum -> fp16 operand tile ---------------------------------------------------------------
            mbar_wait(pz_full, it & 1);
            tc_fence_after();
            M2T_ZT(5);
#pragma unroll 1
            for (int c = ch0; c < ch0 + CH_N; ++c)
                az_cvt_chunk(tmem_base + lane_sel + CF::TM_A + c * 64, oper + c * AZ_OPCH, m, inv);
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(pzs_ready);
            M2T_ZT(6);

            // ---- phase 8: fused branch glue (ref :143-161), as in attn_umma_kernel ---------------------------------------
            mbar_wait(o_full, it & 1);
            mbar_wait(tile_full, it & 1);                     // completed long ago: makes the TMA-written t_k rows visible here
            tc_fence_after();
            M2T_ZT(7);
            if constexpr (!RING && LO) {
              {
                // C = 64: thread = level-1 pixel (win, qi) with its 2 x 2 sub-pixels s = dy * 2 + dx.  In the next tensor the
                // sub-pixel is piece (fx & 3) of chunk (fy & 3) of level-2 pixel (ly / 2, lx / 2).
                uint8_t* sth = sm + CF::OFF_OPER;                 // n_3/2 -> t_3 in place: [chunk][window][4 x 4 px][128 B]
                uint8_t* stl = sm + CF::OFF_OPER + 16384;         // t_2 residual tile: [window * 64 + qi][128 B]
                uint8_t* stn = sm + CF::OFF_OPER + 32768;         // t_3 residual, layout of sth
                const int nv = (pc.y + BLK < h && !single) ? 2 : 1;
                if (tid == 0) {
                    mbar_expect_tx(g_full, (uint32_t)(nv * 8192 * ((LO ? 1 : 0) + (has_next ? 1 : 0))));
                    for (int wv = 0; wv < nv; ++wv) {
                        if constexpr (LO) tma_load_4d(stl + wv * 8192, &mapTlo, g_full, 0, pc.x, pc.y + BLK * wv, pc.b);
                        if (has_next)
                            for (int c2 = 0; c2 < 4; ++c2)
                                tma_load_4d(sth + c2 * 4096 + wv * 2048, &mapTn, g_full, c2 * 64, pc.x >> 1, (pc.y + BLK * wv) >> 1, pc.b);
                    }
                }
                mbar_wait(g_full, it & 1);
                const int lyl = qi >> 3, lxl = qi & 7;
                const uint32_t r2 = (uint32_t)((lyl >> 1) * 4 + (lxl >> 1));
                const uint8_t* tst = sm + CF::OFF_TILE + trow * 128;
                const uint8_t* lrow = stl + m * 128;
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    uint32_t r[16];
                    tmem_ld16(tmem_base + lane_sel + CF::TM_A + s * NB, r);
                    uint4 tk[2], tl[2];
                    tk[0] = *reinterpret_cast<const uint4*>(tst + (((2 * s) ^ (trow & 7)) << 4));
                    tk[1] = *reinterpret_cast<const uint4*>(tst + (((2 * s + 1) ^ (trow & 7)) << 4));
                    if constexpr (LO) {
                        tl[0] = *reinterpret_cast<const uint4*>(lrow + (((2 * s) ^ (m & 7)) << 4));
                        tl[1] = *reinterpret_cast<const uint4*>(lrow + (((2 * s + 1) ^ (m & 7)) << 4));
                    } else {
                        tl[0] = make_uint4(0u, 0u, 0u, 0u); tl[1] = make_uint4(0u, 0u, 0u, 0u);
                    }
                    tmem_ld_wait();
                    if (valid) {
                        const int dy = s >> 1, dx = s & 1;
                        const int fy = ly * 2 + dy, fx = lx * 2 + dx;
                        const long pix = ((long)pc.b * fz.Hp + fy) * fz.Wp + fx;
                        float yv[NB];
                        const __half2* th = reinterpret_cast<const __half2*>(tk);
                        const __half2* tlh = reinterpret_cast<const __half2*>(tl);
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            float2 tf = __half22float2(th[e]);
                            if constexpr (LO) {
                                const float2 lf = __half22float2(tlh[e]);
                                tf.x += lf.x; tf.y += lf.y;
                            }
                            yv[2 * e] = __uint_as_float(r[2 * e]) + tf.x;
                            yv[2 * e + 1] = __uint_as_float(r[2 * e + 1]) + tf.y;
                        }
                        uint4 yo[2], yl[2];
                        __half2* yh = reinterpret_cast<__half2*>(yo);
                        __half2* ylh = reinterpret_cast<__half2*>(yl);
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            yh[e] = __floats2half2_rn(yv[2 * e], yv[2 * e + 1]);
                            const float2 yr = __half22float2(yh[e]);
                            ylh[e] = __floats2half2_rn((yv[2 * e] - yr.x) * 2048.f, (yv[2 * e + 1] - yr.y) * 2048.f);
                        }
                        stg256(fz.Y + pix * NF + NB * br, yo[0], yo[1]);
                        if constexpr (LO) stg256(fz.Ylo + pix * NF + NB * br, yl[0], yl[1]);
                        if (has_next) {
                            const uint32_t c2 = (uint32_t)(2 * (lyl & 1) + dy), pc2 = (uint32_t)(2 * (lxl & 1) + dx);
                            const uint32_t off = c2 * 4096 + (uint32_t)win * 2048 + r2 * 128;
                            const uint32_t c0o = ((2 * pc2) ^ (r2 & 7)) << 4, c1o = ((2 * pc2 + 1) ^ (r2 & 7)) << 4;
                            uint4 hc[2], to[2], tol[2];
                            hc[0] = *reinterpret_cast<const uint4*>(sth + off + c0o);
                            hc[1] = *reinterpret_cast<const uint4*>(sth + off + c1o);
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
                            *reinterpret_cast<uint4*>(sth + off + c0o) = to[0];
                            *reinterpret_cast<uint4*>(sth + off + c1o) = to[1];
                            if constexpr (LO) {
                                *reinterpret_cast<uint4*>(stn + off + c0o) = tol[0];
                                *reinterpret_cast<uint4*>(stn + off + c1o) = tol[1];
                            }
                        }
                    }
                }
                tc_fence_before();
                if (has_next) {
                    fence_proxy_async();
                    asm volatile("bar.sync 1, %0;" ::"n"(CF::NEPI * 32) : "memory");
                    if (tid == 0) {
                        for (int wv = 0; wv < nv; ++wv)
                            for (int c2 = 0; c2 < 4; ++c2) {
                                tma_store_4d(&mapTn, sth + c2 * 4096 + wv * 2048, c2 * 64, pc.x >> 1, (pc.y + BLK * wv) >> 1, pc.b);
                                if constexpr (LO)
                                    tma_store_4d(&mapTnlo, stn + c2 * 4096 + wv * 2048, c2 * 64, pc.x >> 1, (pc.y + BLK * wv) >> 1, pc.b);
                            }
                        tma_store_commit();
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(pair_done);
                if (has_next && tid == 0) {
                    tma_store_wait_read();
                    mbar_arrive(stage_free);
                }
                __syncwarp();
                M2T_ZT(8);
                continue;
              }
            }
            if (RING && (LO || has_next)) {
                uint8_t* stl = sm + CF::OFF_OPER;                 // t_k residual tile -> t_{k+1} residual in place (precise mode)
                // n_{k+1}/2 tile -> t_{k+1} in place: the weight ring in precise mode, the operand region in fast mode
                uint8_t* sth = sm + (LO ? CF::OFF_W : CF::OFF_OPER);
                const int nv = (pc.y + BLK < h && !single) ? 2 : 1;   // real windows of this pair
                if (tid == 0) {
                    mbar_expect_tx(g_full, (uint32_t)(nv * NBLK * 8192 * ((LO ? 1 : 0) + (has_next ? 1 : 0))));
                    for (int wv = 0; wv < nv; ++wv)
                        for (int nb = 0; nb < NBLK; ++nb) {
                            if constexpr (LO)
                                tma_load_4d(stl + nb * AZ_OPCH + wv * 8192, &mapTlo, g_full, nb * 64, pc.x, pc.y + BLK * wv, pc.b);
                            if (has_next)
                                tma_load_4d(sth + nb * AZ_SLOT + wv * 8192, &mapTn, g_full, nb * 64, pc.x, pc.y + BLK * wv, pc.b);
                        }
                }
                mbar_wait(g_full, it & 1);
#pragma unroll 1
                for (int nb = 0; nb < NBLK; ++nb) {
                    const uint8_t* tst = sm + CF::OFF_TILE + nb * AZ_CHUNK + trow * 128;
                    uint8_t* lrow = stl + nb * AZ_OPCH + m * 128;
                    uint8_t* hrow = sth + nb * AZ_SLOT + m * 128;
#pragma unroll
                    for (int jj = 0; jj < JN; ++jj) {
                        const int j = j0 + jj;
                        const int s = nb * SPB + j;
                        uint32_t r[16];
                        tmem_ld16(tmem_base + lane_sel + CF::TM_A + s * NB, r);
                        const uint32_t c0o = (uint32_t)(((2 * j) ^ (m & 7)) << 4), c1o = (uint32_t)(((2 * j + 1) ^ (m & 7)) << 4);
                        uint4 tk[2], tl[2];
                        tk[0] = *reinterpret_cast<const uint4*>(tst + (((2 * j) ^ (trow & 7)) << 4));
                        tk[1] = *reinterpret_cast<const uint4*>(tst + (((2 * j + 1) ^ (trow & 7)) << 4));
                        if constexpr (LO) {
                            tl[0] = *reinterpret_cast<const uint4*>(lrow + c0o);
                            tl[1] = *reinterpret_cast<const uint4*>(lrow + c1o);
                        } else {
                            tl[0] = make_uint4(0u, 0u, 0u, 0u); tl[1] = make_uint4(0u, 0u, 0u, 0u);
                        }
                        tmem_ld_wait();
                        if (valid) {
                            const int fy = ly * S + s / S, fx = lx * S + s % S;
                            const long pix = ((long)pc.b * fz.Hp + fy) * fz.Wp + fx;
                            float yv[NB];
                            const __half2* th = reinterpret_cast<const __half2*>(tk);
                            const __half2* tlh = reinterpret_cast<const __half2*>(tl);
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                float2 tf = __half22float2(th[e]);
                                if constexpr (LO) {
                                    const float2 lf = __half22float2(tlh[e]);
                                    tf.x += lf.x; tf.y += lf.y;
                                }
                                yv[2 * e] = __uint_as_float(r[2 * e]) + tf.x;
                                yv[2 * e + 1] = __uint_as_float(r[2 * e + 1]) + tf.y;
                            }
                            uint4 yo[2], yl[2];
                            __half2* yh = reinterpret_cast<__half2*>(yo);
                            __half2* ylh = reinterpret_cast<__half2*>(yl);
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                yh[e] = __floats2half2_rn(yv[2 * e], yv[2 * e + 1]);
                                const float2 yr = __half22float2(yh[e]);
                                ylh[e] = __floats2half2_rn((yv[2 * e] - yr.x) * 2048.f, (yv[2 * e + 1] - yr.y) * 2048.f);
                            }
                            stg256(fz.Y + pix * NF + NB * br, yo[0], yo[1]);
                            if constexpr (LO) stg256(fz.Ylo + pix * NF + NB * br, yl[0], yl[1]);
                            if (has_next) {
                                uint4 hc[2], to[2], tol[2];
                                hc[0] = *reinterpret_cast<const uint4*>(hrow + c0o);
                                hc[1] = *reinterpret_cast<const uint4*>(hrow + c1o);
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
                                *reinterpret_cast<uint4*>(hrow + c0o) = to[0];
                                *reinterpret_cast<uint4*>(hrow + c1o) = to[1];
                                if constexpr (LO) {
                                    *reinterpret_cast<uint4*>(lrow + c0o) = tol[0];
                                    *reinterpret_cast<uint4*>(lrow + c1o) = tol[1];
                                }
                            }
                        }
                    }
                }
                tc_fence_before();
                if (has_next) {
                    fence_proxy_async();                          // in-place updates -> the TMA stores below
                    asm volatile("bar.sync 1, %0;" ::"n"(CF::NEPI * 32) : "memory");
                    if (tid == 0) {
                        for (int wv = 0; wv < nv; ++wv)
                            for (int nb = 0; nb < NBLK; ++nb) {
                                tma_store_4d(&mapTn, sth + nb * AZ_SLOT + wv * 8192, nb * 64, pc.x, pc.y + BLK * wv, pc.b);
                                if constexpr (LO)
                                    tma_store_4d(&mapTnlo, stl + nb * AZ_OPCH + wv * 8192, nb * 64, pc.x, pc.y + BLK * wv, pc.b);
                            }
                        tma_store_commit();
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(pair_done);
                if (has_next && tid == 0) {                       // after the arrive: the next tile need not wait for this
                    tma_store_wait_read();
                    mbar_arrive(stage_free);
                }
                __syncwarp();
                M2T_ZT(8);
                continue;
            }
            if constexpr (!LO) {      // fast mode without a next tensor (branch 4): nothing to stage, y pieces only
#pragma unroll 1
            for (int nb = 0; nb < NBLK; ++nb) {
                if (valid && nb + 1 < NBLK) {
                    load_l(nb + 1, lnxt);
                    if (has_next) load_h(nb + 1, hnxt);
                }
                const uint8_t* tst = sm + CF::OFF_TILE + nb * AZ_CHUNK + trow * 128;
#pragma unroll
                for (int jj = 0; jj < JN; ++jj) {
                    const int j = j0 + jj;
                    const int s = nb * SPB + j;
                    uint32_t r[16];
                    tmem_ld16(tmem_base + lane_sel + CF::TM_A + s * NB, r);
                    uint4 tk[2];
                    tk[0] = *reinterpret_cast<const uint4*>(tst + (((2 * j) ^ (trow & 7)) << 4));
                    tk[1] = *reinterpret_cast<const uint4*>(tst + (((2 * j + 1) ^ (trow & 7)) << 4));
                    tmem_ld_wait();
                    if (valid) {
                        const int fy = ly * S + s / S, fx = lx * S + s % S;
                        const long pix = ((long)pc.b * fz.Hp + fy) * fz.Wp + fx;
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
                            yv[2 * e] = __uint_as_float(r[2 * e]) + tf.x;
                            yv[2 * e + 1] = __uint_as_float(r[2 * e + 1]) + tf.y;
                        }
                        uint4 yo[2];
                        __half2* yh = reinterpret_cast<__half2*>(yo);
#pragma unroll
                        for (int e = 0; e < 8; ++e) yh[e] = __floats2half2_rn(yv[2 * e], yv[2 * e + 1]);
                        stg256(fz.Y + pix * NF + NB * br, yo[0], yo[1]);
                        if constexpr (LO) {                 // rounding residual of y_k * 2^11 for the split-precision ff conv
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
                            if constexpr (LO) stg256(fz.Tnext_lo + toff, tol[0], tol[1]);
                        }
                    }
                }
                if (nb + 1 < NBLK) {
#pragma unroll
                    for (int j = 0; j < 2 * JN; ++j) { lcur[j] = lnxt[j]; hcur[j] = hnxt[j]; }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(pair_done);
            M2T_ZT(8);
            }
        }
    }
    if constexpr (STG) { if (tid == 0) tma_store_wait_all(); }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, CF::TM_COLS);
}

template <int C, bool LO>
int launch_attn_z_c(const __half* T, const __half* MQ, const __half* WV, int B, int h, int w, cudaStream_t s, bool paired_only,
                    const AttnFuse& fz) {
    using CF = AzCfg<C>;
    CUtensorMap mapT, mapMQ, mapWV;
    {
        const uint64_t dims[4] = {(uint64_t)C, (uint64_t)w, (uint64_t)h, (uint64_t)B};
        const uint64_t str[4] = {2, (uint64_t)C * 2, (uint64_t)w * C * 2, (uint64_t)h * w * C * 2};
        const uint32_t box[4] = {64, WIN, 2 * BLK + 2, 1};
        M2T_TRY(make_tensor_map(&mapT, T, 2, 4, dims, str, box, 3));
    }
    {
        const uint64_t dims[2] = {(uint64_t)C, (uint64_t)CF::NQ}, str[2] = {2, (uint64_t)C * 2};
        const uint32_t box[2] = {64, (uint32_t)(CF::RING ? 128 : CF::NQ)};
        M2T_TRY(make_tensor_map(&mapMQ, MQ, 2, 2, dims, str, box, 3));
    }
    {
        const uint64_t dims[2] = {(uint64_t)C, (uint64_t)C}, str[2] = {2, (uint64_t)C * 2};
        const uint32_t box[2] = {64, (uint32_t)(CF::RING ? 128 : C)};
        M2T_TRY(make_tensor_map(&mapWV, WV, 2, 2, dims, str, box, 3));
    }
    const int nwy = h / BLK, nwx = w / BLK;
    const int cap = device_sm_count() * CF::MIN_CTAS;
    // small inputs (every window finds a CTA slot of its own in one wave): one window per CTA.  The chain of a CTA is then one
    // window's glue long instead of two (the glue is bound by the SM's load/store unit: one 32-byte sector per sub-pixel and
    // tensor), and twice the SMs work; the weights come from L2 either way.
    const int single = !paired_only && B * nwy * nwx <= cap ? 1 : 0;
    const int npairs = single ? B * nwy * nwx : B * ((nwy + 1) / 2) * nwx;
    const int grid = npairs < cap ? npairs : cap;
    CUtensorMap mapTlo = mapT, mapTn = mapT, mapTnlo = mapT;      // used by the staged glue only (precise mode, C = 256)
    if (!CF::RING) {                                              // C = 64: own level for Tlo, one level up for the next tensor
        const uint64_t dims[4] = {(uint64_t)C, (uint64_t)w, (uint64_t)h, (uint64_t)B};
        const uint64_t str[4] = {2, (uint64_t)C * 2, (uint64_t)w * C * 2, (uint64_t)h * w * C * 2};
        const uint32_t box[4] = {64, BLK, BLK, 1};
        if (LO) M2T_TRY(make_tensor_map(&mapTlo, fz.Tlo, 2, 4, dims, str, box, 3));
        if (fz.Tnext != nullptr) {
            if (LO && fz.Tnext_lo == nullptr) { set_error("attn_z: precise mode needs Tnext_lo with Tnext"); return M2T_E_ARG; }
            const uint64_t h2 = (uint64_t)h / 2, w2 = (uint64_t)w / 2;
            const uint64_t dims2[4] = {256, w2, h2, (uint64_t)B};
            const uint64_t str2[4] = {2, 512, w2 * 512, h2 * w2 * 512};
            const uint32_t box2[4] = {64, BLK / 2, BLK / 2, 1};
            M2T_TRY(make_tensor_map(&mapTn, fz.Tnext, 2, 4, dims2, str2, box2, 3));
            if (LO) M2T_TRY(make_tensor_map(&mapTnlo, fz.Tnext_lo, 2, 4, dims2, str2, box2, 3));
        }
    }
    if (CF::RING) {
        const uint64_t dims[4] = {(uint64_t)C, (uint64_t)w, (uint64_t)h, (uint64_t)B};
        const uint64_t str[4] = {2, (uint64_t)C * 2, (uint64_t)w * C * 2, (uint64_t)h * w * C * 2};
        const uint32_t box[4] = {64, BLK, BLK, 1};
        if (LO) M2T_TRY(make_tensor_map(&mapTlo, fz.Tlo, 2, 4, dims, str, box, 3));
        if (fz.Tnext != nullptr) {
            if (LO && fz.Tnext_lo == nullptr) { set_error("attn_z: precise mode needs Tnext_lo with Tnext"); return M2T_E_ARG; }
            M2T_TRY(make_tensor_map(&mapTn, fz.Tnext, 2, 4, dims, str, box, 3));
            if (LO) M2T_TRY(make_tensor_map(&mapTnlo, fz.Tnext_lo, 2, 4, dims, str, box, 3));
        }
    }
    M2T_ENSURE_SMEM((attn_z_kernel<C, LO>), CF::SMEM);
    M2T_CUDA(launch_pdl(attn_z_kernel<C, LO>, dim3(grid), dim3(CF::THREADS), CF::SMEM, s, mapT, mapMQ, mapWV, mapTlo, mapTn, mapTnlo,
                        h, w, npairs, single, fz));
    return M2T_OK;
}

}  // namespace

#ifdef M2T_TIMING
int read_az_timing(long long* host64) {
    M2T_CUDA(cudaMemcpyFromSymbol(host64, g_az_dbg, sizeof(long long) * 192));
    return M2T_OK;
}
#else
int read_az_timing(long long* host64) { memset(host64, 0, sizeof(long long) * 192); return M2T_OK; }
#endif

// T: t_k space-to-depth fp16 [B,h,w,C]; MQ: fp16 [C+32][C] (AttnW::mq); WV: fp16 [C][C] (the v rows of AttnW::wqkv_f)
// paired_only (M2T_VAR_AZ_PAIRED): two windows per CTA at every size (the results are bit-identical either way)
int launch_attn_z(int C, const __half* T, const __half* MQ, const __half* WV, int B, int h, int w, cudaStream_t s,
                  const AttnFuse& fz, bool paired_only) {
    if (h % BLK || w % BLK) { set_error("attn_z: %dx%d is not a multiple of the 8x8 block", h, w); return M2T_E_ARG; }
    if (fz.Y == nullptr || fz.T != T) { set_error("attn_z: the fused branch glue is not optional"); return M2T_E_ARG; }
    const bool lo = fz.Tlo != nullptr;
    if (C == 64) return lo ? launch_attn_z_c<64, true>(T, MQ, WV, B, h, w, s, paired_only, fz) : launch_attn_z_c<64, false>(T, MQ, WV, B, h, w, s, paired_only, fz);
    if (C == 256) return lo ? launch_attn_z_c<256, true>(T, MQ, WV, B, h, w, s, paired_only, fz) : launch_attn_z_c<256, false>(T, MQ, WV, B, h, w, s, paired_only, fz);
    set_error("attn_z: unsupported channel count %d", C);
    return M2T_E_UNSUPPORTED;
}

}  // namespace m2t
