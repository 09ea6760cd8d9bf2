// Fused last stage of the tail for a x2 PixelShuffle (ref M2Trans_network.py:46-55, :72-76):
//     A [B,h,w,64] fp16  --1x1 conv 64->256 + bias--> PixelShuffle(2) --GELU--> U [B,2h,2w,64]
//                        --3x3 reflect conv 64->3 (no bias)--> clamp --> crop --> y fp32 NCHW
// in ONE kernel, so U (2 048 B per LR pixel at x4: the largest tensor of the forward) never leaves the SM.
// x4 runs it on T1 (after the first tail_up); x2 runs it on XR directly.  x3 keeps the two-kernel path.
//
// One work item = a 16 x 12 tile of A pixels plus a 1-pixel halo (18 x 14 = 252 rows, ONE TMA box, OOB -> 0):
//   1. GEMM-1 on tcgen05: 2 M-tiles x 2 N-halves (sub-pixel rows) of M128 x N128 x K64, accumulators
//      double-buffered in TMEM.
//   2. epilogue 1 (16 warps): TMEM -> bias + GELU (packed fp32x2) -> fp16 -> the 36 x 28-pixel U tile in shared
//      memory, 128-B-swizzled rows (PixelShuffle = the row address).  Pixels on rows/cols 1 and size-2 are also
//      stored at -1 / size: the reflected ring the conv needs is built in place.
//   3. conv as ONE GEMM per 128 U pixels: D[px][tap*3+c] = sum_k U[px][k] * W[tap][c][k]  (N = 27 -> 32, K = 64).
//      Every U row is read from shared memory once (the N = 16 implicit GEMM of tail_out re-reads it 9 times,
//      which makes that kernel shared-memory-bandwidth-bound).
//   4. epilogue 2: D -> fp32 planes [27][1024] in the shared memory the U tile occupied (it is dead once the
//      conv MMAs have completed), then each output pixel gathers its 9 taps: y[c] = sum_tap D[px+shift(tap)][tap*3+c].
// TMEM: 2 x 128 (GEMM-1) + 8 x 32 (conv) = 512 columns.  Shared memory: 32 KB W1 + 32 KB A + 128 KB U + 4 KB Wc.
#include "common.cuh"
#include "gelu.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace m2t {

constexpr int TF_TH = 16, TF_TW = 12;                  // interior tile of A pixels
constexpr int TF_AH = TF_TH + 2, TF_AW = TF_TW + 2;    // with halo
constexpr int TF_APX = TF_AH * TF_AW;                  // 252 rows of the GEMM-1 A operand
constexpr int TF_UW = 2 * TF_AW;                       // U tile: 28 px wide, 36 tall
constexpr int TF_UPX = 1024;                           // 1008 U pixels padded to 8 M-tiles
constexpr int TF_NC = 32;                              // conv GEMM N
constexpr uint32_t TF_OFF_W1 = 0;
constexpr uint32_t TF_OFF_A = 32768;
constexpr uint32_t TF_OFF_U = 65536;
constexpr uint32_t TF_OFF_WC = TF_OFF_U + TF_UPX * 128;
constexpr uint32_t TF_OFF_BIAS = TF_OFF_WC + TF_NC * 128;
constexpr uint32_t TF_OFF_BAR = TF_OFF_BIAS + 256 * 4;
constexpr uint32_t TF_SMEM = 1024 + TF_OFF_BAR + 256;
constexpr int TF_EPI = 512;                            // 16 epilogue warps: the GELU epilogue is latency-bound with fewer
constexpr int TF_THREADS = TF_EPI + 64;                // + warp 16 TMA, warp 17 MMA
constexpr uint32_t TF_COL_D = 256;                     // TMEM column of the conv accumulators

__device__ __forceinline__ void tf_epi_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

#ifdef M2T_TIMING
__device__ long long g_tail_dbg[64];   // CTA 0, first 8 tiles: 6 clock64 stamps per tile (m2t_debug_attn_timing record 4)
#define M2T_TT(slot) do { if (blockIdx.x == 0 && tid == 0 && it < 8) g_tail_dbg[(slot) + 8 * it] = clock64(); } while (0)
#else
#define M2T_TT(slot) do { } while (0)
#endif

__global__ void __launch_bounds__(TF_THREADS, 1)
tail_fused_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW,
                  const float* __restrict__ bias, const __half* __restrict__ wc, float* __restrict__ y,
                  int Bc, int h, int w, int hout, int wout, int b0, float rgb_range) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    float* sbias = reinterpret_cast<float*>(sm + TF_OFF_BIAS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + TF_OFF_BAR);
    uint64_t* wfull = bars;
    uint64_t* afull = bars + 1;
    uint64_t* aempty = bars + 2;
    uint64_t* accfull = bars + 3;     // [2]
    uint64_t* accempty = bars + 5;    // [2]
    uint64_t* ufull = bars + 7;
    uint64_t* dfull = bars + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_x = (wout + 2 * TF_TW - 1) / (2 * TF_TW), tiles_y = (hout + 2 * TF_TH - 1) / (2 * TF_TH);
    const int per_img = tiles_x * tiles_y;
    const int ntiles = Bc * per_img;

    if (warp == 17) tmem_alloc(tmem_slot, 512);
    if (tid == TF_EPI) {
        mbar_init(wfull, 1); mbar_init(afull, 1); mbar_init(aempty, 1);
        for (int a = 0; a < 2; ++a) { mbar_init(&accfull[a], 1); mbar_init(&accempty[a], TF_EPI / 32); }
        mbar_init(ufull, 1); mbar_init(dfull, 1);
        mbar_fence_init();
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapW);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 16) {
        // ---- TMA producer ---------------------------------------------------------------------------------
        if (elect_one_sync()) {
            mbar_expect_tx(wfull, 256 * 128);
            tma_load_2d(sm + TF_OFF_W1, &mapW, wfull, 0, 0);
        }
        pdl_wait();
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int bl = tile / per_img, r = tile - bl * per_img;
            const int y0 = (r / tiles_x) * TF_TH, x0 = (r % tiles_x) * TF_TW;
            mbar_wait(aempty, (it & 1) ^ 1);
            if (elect_one_sync()) {
                mbar_expect_tx(afull, TF_APX * 128);
                tma_load_4d(sm + TF_OFF_A, &mapA, afull, 0, x0 - 1, y0 - 1, bl);
            }
            __syncwarp();
        }
    } else if (warp == 17) {
        // ---- MMA issuer -----------------------------------------------------------------------------------
        constexpr uint32_t idesc1 = umma_idesc_f16(128, 128);
        constexpr uint32_t idesc2 = umma_idesc_f16(128, TF_NC);
        constexpr uint64_t tmpl = umma_smem_desc(0, 16, 1024, UMMA_LAYOUT_SW128);
        mbar_wait(wfull, 0);
        uint32_t it = 0, un = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            mbar_wait(afull, it & 1);
            tc_fence_after();
            for (int u = 0; u < 4; ++u, ++un) {            // unit = (M-tile u/2, sub-pixel row u%2)
                const uint32_t acc = un & 1, aph = (un >> 1) & 1;
                mbar_wait(&accempty[acc], aph ^ 1);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint64_t da0 = umma_desc_at(tmpl, base + TF_OFF_A + (u >> 1) * 16384);
                    const uint64_t db0 = umma_desc_at(tmpl, base + TF_OFF_W1 + (u & 1) * 16384);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base + acc * 128, da0 + 2 * k, db0 + 2 * k, idesc1, k ? 1u : 0u);
                    umma_commit(&accfull[acc]);
                    if (u == 3) umma_commit(aempty);
                }
                __syncwarp();
            }
            mbar_wait(ufull, it & 1);
            tc_fence_after();
            if (elect_one_sync()) {
                const uint64_t db0 = umma_desc_at(tmpl, base + TF_OFF_WC);
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    const uint64_t da0 = umma_desc_at(tmpl, base + TF_OFF_U + m * 16384);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_f16_ss(tmem_base + TF_COL_D + m * TF_NC, da0 + 2 * k, db0 + 2 * k, idesc2, k ? 1u : 0u);
                }
                umma_commit(dfull);
            }
            __syncwarp();
        }
    } else {
        // ---- epilogue warps -------------------------------------------------------------------------------
        // warpgroup wg serves sub-pixel column wg/2 and channels (wg%2)*32.. of every accumulator unit
        const int quad = warp & 3, wg = warp >> 2;
        const int et = tid;                                          // 0..511
        const uint32_t lanef = (uint32_t)(quad * 32) << 16;
        {   // conv weights [9][16][64] (rows 0..2 of each tap real) -> [32][64] rows tap*3+c, 128-B swizzled
            if (et < 256) {
                const int row = et >> 3, ch = et & 7;
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (row < 27) v = *reinterpret_cast<const uint4*>(wc + ((row / 3) * 16 + (row % 3)) * NF + ch * 8);
                *reinterpret_cast<uint4*>(sm + TF_OFF_WC + row * 128 + ((ch ^ (row & 7)) << 4)) = v;
                sbias[et] = bias[et];
            }
        }
        fence_proxy_async();
        tf_epi_sync();
        pdl_wait();
        const int H2 = 2 * h, W2 = 2 * w;
        const long plane = (long)hout * wout;
        float* planes = reinterpret_cast<float*>(sm + TF_OFF_U);
        uint32_t it = 0, un = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int bl = tile / per_img, r = tile - bl * per_img;
            const int y0 = (r / tiles_x) * TF_TH, x0 = (r % tiles_x) * TF_TW;
            M2T_TT(0);
            // ---- epilogue 1: GEMM-1 accumulators -> GELU -> U tile ------------------------------------------
            for (int u = 0; u < 4; ++u, ++un) {
                const uint32_t acc = un & 1, aph = (un >> 1) & 1;
                const int p = (u >> 1) * 128 + quad * 32 + lane;     // row of the A tile
                const int ty = p / TF_AW, tx = p - ty * TF_AW;
                const int gy = y0 - 1 + ty, gx = x0 - 1 + tx;
                const bool valid = p < TF_APX && gy >= 0 && gy < h && gx >= 0 && gx < w;
                const int uu = u & 1, vv = wg >> 1, c0 = (wg & 1) * 32; // sub-pixel (row, col), first channel
                const int Y = 2 * gy + uu, X = 2 * gx + vv;
                int dyr = Y == 1 ? -2 : (Y == H2 - 2 ? 2 : 0);        // reflected copy: -1 <- 1, H2 <- H2-2
                int dxr = X == 1 ? -2 : (X == W2 - 2 ? 2 : 0);
                if ((unsigned)(2 * ty + uu + dyr) >= (unsigned)(2 * TF_AH)) dyr = 0;   // target outside this tile: the
                if ((unsigned)(2 * tx + vv + dxr) >= (unsigned)TF_UW) dxr = 0;        // tile that needs it writes it
                const int R00 = (2 * ty + uu) * TF_UW + 2 * tx + vv;
                const float* bs = sbias + (uu * 2 + vv) * NF;
                mbar_wait(&accfull[acc], aph);
                tc_fence_after();
                {
                    uint32_t rr[32];
                    tmem_ld32(tmem_base + acc * 128 + vv * NF + c0 + lanef, rr);
                    tmem_ld_wait();
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        uint4 q;
                        uint32_t* pq = reinterpret_cast<uint32_t*>(&q);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int c = c0 + v * 8 + 2 * e;
                            pq[e] = gelu_pair_h2(f2_pack(__uint_as_float(rr[v * 8 + 2 * e]), __uint_as_float(rr[v * 8 + 2 * e + 1])),
                                                 *reinterpret_cast<const uint64_t*>(bs + c));
                        }
                        if (valid) {
                            const int ch = (c0 >> 3) + v;
#pragma unroll
                            for (int a = 0; a < 2; ++a) {
                                if (a == 1 && dyr == 0) continue;
#pragma unroll
                                for (int b = 0; b < 2; ++b) {
                                    if (b == 1 && dxr == 0) continue;
                                    const int R = R00 + (a ? dyr * TF_UW : 0) + (b ? dxr : 0);
                                    *reinterpret_cast<uint4*>(sm + TF_OFF_U + R * 128 + ((ch ^ (R & 7)) << 4)) = q;
                                }
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&accempty[acc]);
            }
            M2T_TT(1);
            fence_proxy_async();           // U tile written by the generic proxy, read by the tensor core
            tf_epi_sync();
            if (et == 0) mbar_arrive(ufull);
            M2T_TT(2);
            // ---- epilogue 2: conv accumulators -> planes -> 9-tap gather -------------------------------------
            mbar_wait(dfull, it & 1);
            tc_fence_after();
            M2T_TT(3);
#pragma unroll 1
            for (int jj = 0; jj < 2; ++jj) {
                const int m = wg + 4 * jj;
                uint32_t rr[32];
                tmem_ld32(tmem_base + TF_COL_D + m * TF_NC + lanef, rr);
                tmem_ld_wait();
                float* pp = planes + m * 128 + quad * 32 + lane;
#pragma unroll
                for (int c = 0; c < 27; ++c) pp[c * TF_UPX] = __uint_as_float(rr[c]);
            }
            tc_fence_before();
            tf_epi_sync();
            M2T_TT(4);
#pragma unroll 1
            for (int o = et; o < 4 * TF_TH * TF_TW; o += TF_EPI) {    // 32 x 24 output pixels of the tile
                const int oyl = o / (2 * TF_TW), oxl = o - oyl * (2 * TF_TW);
                const int Yo = 2 * y0 + oyl, Xo = 2 * x0 + oxl;
                if (Yo < hout && Xo < wout) {
                    const float* pi = planes + (oyl + 2) * TF_UW + (oxl + 2);
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const int off = (tap / 3 - 1) * TF_UW + (tap % 3 - 1);
                        s0 += pi[(tap * 3 + 0) * TF_UPX + off];
                        s1 += pi[(tap * 3 + 1) * TF_UPX + off];
                        s2 += pi[(tap * 3 + 2) * TF_UPX + off];
                    }
                    float* yp = y + (long)(b0 + bl) * 3 * plane + (long)Yo * wout + Xo;
                    yp[0] = fminf(fmaxf(s0, 0.f), rgb_range);
                    yp[plane] = fminf(fmaxf(s1, 0.f), rgb_range);
                    yp[2 * plane] = fminf(fmaxf(s2, 0.f), rgb_range);
                }
            }
            tf_epi_sync();                 // planes consumed: the next tile may overwrite the U tile
            M2T_TT(5);
        }
    }
    pdl_trigger();
    tc_fence_before();
    __syncthreads();
    if (warp == 17) tmem_dealloc(tmem_base, 512);
}

#ifdef M2T_TIMING
int read_tail_timing(long long* host64) {
    M2T_CUDA(cudaMemcpyFromSymbol(host64, g_tail_dbg, sizeof(long long) * 64));
    return M2T_OK;
}
#else
int read_tail_timing(long long* host64) { memset(host64, 0, sizeof(long long) * 64); return M2T_OK; }
#endif

// A: fp16 [Bc][h][w][64]; W1: fp16 [256][64] sub-pixel-major; bias fp32 [256]; Wc: fp16 [9][16][64];
// y: fp32 NCHW, images b0.. cropped to hout x wout (hout <= 2h, wout <= 2w)
int launch_tail_fused(const __half* A, const __half* W1, const float* bias, const __half* Wc, float* y, int Bc, int h,
                      int w, int hout, int wout, int b0, float rgb_range, cudaStream_t s) {
    if (h % TF_TH) { set_error("tail_fused: %d rows is not a multiple of %d", h, TF_TH); return M2T_E_ARG; }
    CUtensorMap mapA, mapW;
    {
        const uint64_t dims[4] = {NF, (uint64_t)w, (uint64_t)h, (uint64_t)Bc};
        const uint64_t str[4] = {2, NF * 2, (uint64_t)w * NF * 2, (uint64_t)h * w * NF * 2};
        const uint32_t box[4] = {NF, TF_AW, TF_AH, 1};
        M2T_TRY(make_tensor_map(&mapA, A, 2, 4, dims, str, box, 3));
    }
    {
        const uint64_t dims[2] = {NF, 256}, str[2] = {2, NF * 2};
        const uint32_t box[2] = {NF, 256};
        M2T_TRY(make_tensor_map(&mapW, W1, 2, 2, dims, str, box, 3));
    }
    M2T_ENSURE_SMEM(tail_fused_kernel, TF_SMEM);
    const int ntiles = Bc * ((hout + 2 * TF_TH - 1) / (2 * TF_TH)) * ((wout + 2 * TF_TW - 1) / (2 * TF_TW));
    const int grid = ntiles < device_sm_count() ? ntiles : device_sm_count();
    M2T_CUDA(launch_pdl(tail_fused_kernel, dim3(grid), dim3(TF_THREADS), TF_SMEM, s, mapA, mapW, bias, Wc, y, Bc, h, w,
                        hout, wout, b0, rgb_range));
    return M2T_OK;
}

}  // namespace m2t
