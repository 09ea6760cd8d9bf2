// tcgen05 implicit-GEMM form of CFTM.feed_forward + residual (ref M2Trans_network.py:124-126, :164):
//   Xout[p][o] = sum_{tap,c} Y[p + tap][c] * W[tap][o][c] + bias[o] + Xin[p][o]        (zero padding)
// One output tile = 16 rows x 8 pixels (M = 128), N = 64 output channels, K = 9 taps x 64 channels.
//
// No im2col: TMA loads ONE 18 x 10 pixel halo tile of Y (fp16 NHWC, 128 B per pixel = one 128-byte swizzle
// row; out-of-frame pixels arrive as zeros = the conv's zero padding).  The A operand of tap (dy,dx) is the
// same shared-memory tile addressed through a descriptor whose start is shifted by (dy*10+dx) rows and whose
// 8-row groups are 10 rows apart (SBO = 1280 B): the 128-byte swizzle phase follows the absolute
// shared-memory address, so shifted descriptors read exactly what TMA wrote (tests/test_probes.py pins this).
// 36 MMAs (M=128, N=64, K=16) accumulate one tile in TMEM; two accumulators let the epilogue of tile i
// overlap the MMAs of tile i+1.  The 9 x 64 x 64 weights stay resident in shared memory (72 KB).
//
// Warp roles (192 threads): warps 0-3 epilogue (TMEM -> smem staging -> coalesced +bias +residual store and
// InstanceNorm partial sums), warp 4 TMA producer, warp 5 MMA issuer.
// The stage is memory-bound (reads 128 B Y + 256 B X, writes 256 B X per pixel); the tensor pipe is ~half idle.
#include "common.cuh"
#include "epilogue.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace m2t {

constexpr int CU_TH = 16, CU_TW = 8;                                  // output tile
constexpr int CU_HW = CU_TW + 2;                                      // halo tile width (10)
constexpr uint32_t CU_TILE_BYTES = (CU_TH + 2) * CU_HW * 128;         // 23040
constexpr uint32_t CU_STAGE = 23 * 1024;                              // 1024-aligned stage pitch
constexpr int CU_STAGES = 3;
constexpr uint32_t CU_W_BYTES = 9 * NF * 128;                         // 73728
constexpr uint32_t CU_OFF_A = CU_W_BYTES;
constexpr uint32_t CU_OFF_O = CU_OFF_A + CU_STAGES * CU_STAGE;
constexpr uint32_t CU_O_BYTES = 128 * EPI_LD * 4;                      // one fp32 staging tile per epilogue warpgroup
constexpr uint32_t CU_OFF_RED = CU_OFF_O + 2 * CU_O_BYTES;            // 2 x float [4][2][64]
constexpr uint32_t CU_OFF_BAR = CU_OFF_RED + 2 * 4 * 2 * NF * 4;
constexpr uint32_t CU_SMEM = 1024 + CU_OFF_BAR + 256;
#ifndef M2T_CONV_WGS
#define M2T_CONV_WGS 1
#endif
constexpr int CU_WGS = M2T_CONV_WGS;   // epilogue warpgroups (A/B switch for tuning)
constexpr int CU_THREADS = CU_WGS == 2 ? 320 : 192;   // warps 0-3: epilogue (even tiles), 4: TMA, 5: MMA, 6-9: epilogue of odd tiles

__global__ void __launch_bounds__(CU_THREADS, 1)
ffconv_umma_kernel(const __grid_constant__ CUtensorMap mapY, const __grid_constant__ CUtensorMap mapW,
                   const float* __restrict__ bias, const float* Xin, float* Xout, double* __restrict__ stats, int B,
                   int Hp, int Wp, const float* __restrict__ res, __half* __restrict__ xr) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + CU_OFF_BAR);
    uint64_t* full = bars;                       // [STAGES]
    uint64_t* empty = bars + CU_STAGES;          // [STAGES]
    uint64_t* wfull = bars + 2 * CU_STAGES;
    uint64_t* tfull = bars + 2 * CU_STAGES + 1;  // [2]
    uint64_t* tempty = bars + 2 * CU_STAGES + 3; // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * CU_STAGES + 5);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_x = Wp / CU_TW, tiles_y = Hp / CU_TH;
    const int per_img = tiles_x * tiles_y;
    const int ntiles = B * per_img;
    // contiguous tile range per CTA: neighbouring tiles share halo rows in L2 and mostly belong to one image,
    // so the InstanceNorm partial sums are flushed once or twice per CTA instead of once per tile
    const int tile_lo = (int)((long)blockIdx.x * ntiles / gridDim.x);
    const int tile_hi = (int)((long)(blockIdx.x + 1) * ntiles / gridDim.x);

    if (warp == 5) tmem_alloc(tmem_slot, 128);
    if (tid == 128) {
        for (int s = 0; s < CU_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(wfull, 1);
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
        mbar_fence_init();
        tma_prefetch_desc(&mapY);
        tma_prefetch_desc(&mapW);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();                              // the next kernel may start its own prologue

    if (warp == 4) {
        // TMA producer: the whole warp runs the loop, one elected lane issues
        if (elect_one_sync()) {
            mbar_expect_tx(wfull, CU_W_BYTES);   // weights are constants: loaded while the previous kernel drains
            for (int tap = 0; tap < 9; ++tap) tma_load_2d(sm + tap * NF * 128, &mapW, wfull, 0, tap * NF);
        }
        pdl_wait();
        uint32_t it = 0;
        for (int tile = tile_lo; tile < tile_hi; ++tile, ++it) {
            const int b = tile / per_img, r = tile - b * per_img;
            const int y0 = (r / tiles_x) * CU_TH, x0 = (r % tiles_x) * CU_TW;
            const uint32_t s = it % CU_STAGES, ph = (it / CU_STAGES) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            if (elect_one_sync()) {
                mbar_expect_tx(&full[s], CU_TILE_BYTES);
                tma_load_4d(sm + CU_OFF_A + s * CU_STAGE, &mapY, &full[s], 0, x0 - 1, y0 - 1, b);
            }
            __syncwarp();
        }
    } else if (warp == 5) {
        // MMA issuer: warp-uniform loop, one elected lane issues the 36 MMAs of a tile and the two commits
        constexpr uint32_t idesc = umma_idesc_f16(128, NF);
        constexpr uint64_t tmpl_a = umma_smem_desc(0, 16, CU_HW * 128, UMMA_LAYOUT_SW128);
        constexpr uint64_t tmpl_b = umma_smem_desc(0, 16, 1024, UMMA_LAYOUT_SW128);
        mbar_wait(wfull, 0);
        uint32_t it = 0;
        for (int tile = tile_lo; tile < tile_hi; ++tile, ++it) {
            const uint32_t s = it % CU_STAGES, ph = (it / CU_STAGES) & 1;
            const uint32_t acc = it & 1, aph = (it >> 1) & 1;
            mbar_wait(&tempty[acc], aph ^ 1);
            mbar_wait(&full[s], ph);
            tc_fence_after();
            if (elect_one_sync()) {
                const uint64_t da0 = umma_desc_at(tmpl_a, base + CU_OFF_A + s * CU_STAGE);
                const uint64_t db0 = umma_desc_at(tmpl_b, base);
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t da = da0 + (uint64_t)((((tap / 3) * CU_HW + (tap % 3)) * 128 + k * 32) >> 4);
                        const uint64_t db = db0 + (uint64_t)((tap * NF * 128 + k * 32) >> 4);
                        umma_f16_ss(tmem_base + acc * NF, da, db, idesc, (tap | k) ? 1u : 0u);
                    }
                }
                umma_commit(&empty[s]);
                umma_commit(&tfull[acc]);
            }
            __syncwarp();
        }
    } else {
        // Two epilogue warpgroups: wg 0 (warps 0-3) drains accumulator 0 = even tiles, wg 1 (warps 6-9) accumulator 1
        // = odd tiles, each with its own staging tile and named barrier.  The residual rows of a tile are requested
        // before its accumulator is complete, so up to 64 KB of loads per SM overlap the MMAs and the other group.
        const int wg = warp >= 6 ? 1 : 0, quad = warp & 3;
        const int t = quad * 32 + lane;                         // 0..127 inside the warpgroup = TMEM lane = tile pixel
        float* Os = reinterpret_cast<float*>(sm + CU_OFF_O + wg * CU_O_BYTES);
        float (*red)[2][NF] = reinterpret_cast<float (*)[2][NF]>(sm + CU_OFF_RED + wg * 4 * 2 * NF * 4);
        pdl_wait();
        EpiStats st;
        st.clear();
        int cur_b = -1;
        uint32_t it = 0;
        for (int tile = tile_lo; tile < tile_hi; ++tile, ++it) {
            if (CU_WGS == 2 && (int)(it & 1) != wg) continue;
            const int b = tile / per_img, r = tile - b * per_img;
            const int y0 = (r / tiles_x) * CU_TH, x0 = (r % tiles_x) * CU_TW;
            const uint32_t acc = it & 1, aph = (it >> 1) & 1;
            float4 xi[16];
            epilogue_load_residual<CU_TW>(t, Xin, b, y0, x0, Hp, Wp, xi);
            if (b != cur_b) {                                  // image changed: publish the finished image's sums
                if (cur_b >= 0) { if (wg == 0) epilogue_flush_stats<1>(red, t, st, stats, cur_b); else epilogue_flush_stats<2>(red, t, st, stats, cur_b); }
                cur_b = b;
            }
            mbar_wait(&tfull[acc], aph);
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < NF; c0 += 32) {
                uint32_t rr[32];
                tmem_ld32(tmem_base + acc * NF + c0 + ((uint32_t)(quad * 32) << 16), rr);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) Os[t * EPI_LD + c0 + i] = __uint_as_float(rr[i]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);          // accumulator drained: MMA may reuse it
            if (wg == 0) epi_sync<1>(); else epi_sync<2>();    // all 128 staged rows visible
            epilogue_apply<CU_TW>(Os, t, xi, bias, Xout, st, b, y0, x0, Hp, Wp, res, xr);
            if (wg == 0) epi_sync<1>(); else epi_sync<2>();    // Os free for the next tile
        }
        if (cur_b >= 0) { if (wg == 0) epilogue_flush_stats<1>(red, t, st, stats, cur_b); else epilogue_flush_stats<2>(red, t, st, stats, cur_b); }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, 128);
}

int launch_ffconv_umma(const __half* Y, const __half* Wpk, const float* bias, const float* Xin, float* Xout,
                       double* stats, const Geom& g, cudaStream_t s, const float* res, __half* xr) {
    CUtensorMap mapY, mapW;
    {
        const uint64_t dims[4] = {NF, (uint64_t)g.Wp, (uint64_t)g.Hp, (uint64_t)g.B};
        const uint64_t str[4] = {2, NF * 2, (uint64_t)g.Wp * NF * 2, (uint64_t)g.Hp * g.Wp * NF * 2};
        const uint32_t box[4] = {NF, CU_HW, CU_TH + 2, 1};
        M2T_TRY(make_tensor_map(&mapY, Y, 2, 4, dims, str, box, 3));
    }
    {
        const uint64_t dims[2] = {NF, 9 * NF}, str[2] = {2, NF * 2};
        const uint32_t box[2] = {NF, NF};
        M2T_TRY(make_tensor_map(&mapW, Wpk, 2, 2, dims, str, box, 3));
    }
    M2T_ENSURE_SMEM(ffconv_umma_kernel, CU_SMEM);
    const int ntiles = g.B * (g.Hp / CU_TH) * (g.Wp / CU_TW);
    const int grid = ntiles < device_sm_count() ? ntiles : device_sm_count();
    M2T_CUDA(launch_pdl(ffconv_umma_kernel, dim3(grid), dim3(CU_THREADS), CU_SMEM, s, mapY, mapW, bias, Xin, Xout, stats, g.B, g.Hp,
                        g.Wp, res, xr));
    return M2T_OK;
}

}  // namespace m2t
