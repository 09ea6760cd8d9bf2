// Feed-forward 3x3 conv of a CFTM in precise mode (split-precision weights AND input, ref M2Trans_network.py:163-164),
// as a CTA PAIR on two SMs (tcgen05 cta_group::2, thread-block cluster of 2).
//
// ffconv_umma_kernel<W2 = true> (conv_umma.cu) gives each of two CTAs 32 output channels of the SAME tile: the halo tiles
// of Y and of its rounding residual are fetched and, worse, read from shared memory as the A operand twice per tile (the
// kernel is bound by shared-memory bandwidth, see conv_umma.cu).  Here each CTA of a pair owns a DIFFERENT tile (M = 256
// over the pair) and half of the weight rows: CTA r keeps rows 32r..32r+31 of the fp16 weights Wh and of their residuals
// Wl (the same 72 KB slab BlockW::ffw2 holds for channel half r), and one cta_group::2 MMA with N = 64 reads rows 0..31
// from CTA 0 and 32..63 from CTA 1.  Per tile:
//     D_hi[128 px][64]  = sum over taps, k:  Yh . Wh^T                         36 MMAs (M 256, N 64, K 16)
//     D_lo[128 px][64]  = sum:  Yh . Wl^T  +  Yl . Wh^T                        72 MMAs
//     x_out = x_in + bias + D_hi + D_lo * 2^-11
// i.e. exactly the products and accumulation order of the W2 kernel.  Shared-memory traffic per tile: A 3 x 36 x 4 KB +
// B 108 x 1 KB (each CTA serves only its own half of B) against 2 x (72 x 4 KB + 108 KB) before.
//
// Measured (B200): bit-identical outputs, and NO gain -- 485 vs 476 us per launch at cfg3 (1.7 M pixels), 56 vs 47 us on
// one 300 x 400 frame, 17 vs 13 us at cfg1: the A-operand re-reads this saves were not what bounds the precise-mode conv
// (10 K cycles per tile and SM against 3.5 K of MMAs and ~6 K of shared-memory traffic in either form), and a pair has half
// as many independent work items on small inputs.  Kept as an opt-in variant (M2T_VAR_W2_PAIR) and as the repo's worked
// example of the cta_group::2 protocol; the default stays conv_umma.cu's two-CTAs-per-tile form.
//
// Protocol (leader = CTA 0 of the pair): both producers load their own tiles with the cta_group::2 form of the TMA load,
// whose transaction bytes count on the LEADER's `full` barrier; the leader's MMA warp issues for both; tcgen05.commit
// multicasts `empty` (stage free) and `tfull` (accumulator complete) to both CTAs; the epilogue warps of both CTAs arrive
// on the leader's `tempty`.  Everything else (fp32 residual tile ring, in-place epilogue, statistics, TMA store) is local
// and as in conv_umma.cu, with the 64 channels of the tile handled as two 32-channel half tiles one after the other.
#include "common.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace m2t {
namespace {

constexpr int CP_TH = 16, CP_TW = 8, CP_HW = CP_TW + 2;
constexpr uint32_t CP_TILE_BYTES = (CP_TH + 2) * CP_HW * 128;          // 23040
constexpr uint32_t CP_STAGE = 23 * 1024;
constexpr uint32_t CP_ASTAGE = 2 * CP_STAGE;                            // Y halo tile | residual halo tile
constexpr int CP_XSTAGES = 3;
constexpr uint32_t CP_XHALF = 128 * 128;                                // 128 pixels x 32 channels x 4 B
constexpr uint32_t CP_W_BYTES = 9 * NF * 128;                           // 73728: per tap [32 Wh rows | 32 Wl rows]
constexpr uint32_t CP_OFF_A = CP_W_BYTES;
constexpr uint32_t CP_OFF_X = CP_OFF_A + 2 * CP_ASTAGE;
constexpr uint32_t CP_OFF_BIAS = CP_OFF_X + CP_XSTAGES * CP_XHALF;
constexpr uint32_t CP_OFF_BAR = CP_OFF_BIAS + NF * 4;
constexpr uint32_t CP_SMEM = 1024 + CP_OFF_BAR + 256;
constexpr int CP_THREADS = 224;
constexpr uint32_t CP_PEER_MASK = 0xFEFFFFFFu;                          // shared::cluster address -> the pair's even CTA
static_assert(CP_SMEM <= 232448, "conv_pair exceeds the shared memory of an SM");

__device__ __forceinline__ uint32_t cp_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cp_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cp_tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void cp_tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void cp_mma2(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs of the pair once all MMAs issued so far are complete
__device__ __forceinline__ void cp_commit2(uint64_t* bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
// TMA loads into THIS CTA's shared memory whose bytes count on the leader's barrier
__device__ __forceinline__ void cp_tma2_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"((uint64_t)m), "r"(smem_u32(bar) & CP_PEER_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void cp_tma2_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"((uint64_t)m), "r"(smem_u32(bar) & CP_PEER_MASK), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void cp_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & CP_PEER_MASK) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CP_THREADS, 1)
ffconv_pair_kernel(const __grid_constant__ CUtensorMap mapY, const __grid_constant__ CUtensorMap mapYlo,
                   const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapXin,
                   const __grid_constant__ CUtensorMap mapXout, const float* __restrict__ bias, double* __restrict__ stats,
                   int B, int Hp, int Wp, const float* __restrict__ res, __half* __restrict__ xr) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    float* sbias = reinterpret_cast<float*>(sm + CP_OFF_BIAS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + CP_OFF_BAR);
    uint64_t* full = bars;                        // [2]  LEADER: the halo tiles of both CTAs landed
    uint64_t* empty = bars + 2;                   // [2]  each CTA: its stage was consumed by the MMAs (multicast commit)
    uint64_t* wfull = bars + 4;                   //      LEADER: both weight slabs landed
    uint64_t* tfull = bars + 5;                   // [2]  each CTA: accumulator complete (multicast commit)
    uint64_t* tempty = bars + 7;                  // [2]  LEADER: drained by the epilogue warps of both CTAs
    uint64_t* xfull = bars + 9;                   // [3]  local: residual half tile landed
    uint64_t* xout = bars + 12;                   // [3]  local: result half tile complete in smem
    uint64_t* xempty = bars + 15;                 // [3]  local: buffer free
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = (int)cp_ctarank();
    const int tiles_x = Wp / CP_TW, tiles_y = Hp / CP_TH;
    const int per_img = tiles_x * tiles_y;
    const int ntiles = B * per_img;
    const int pair = (int)(blockIdx.x >> 1), npairs = (int)(gridDim.x >> 1);
    const int tile_lo = (int)((long)pair * ntiles / npairs);
    const int tile_hi = (int)((long)(pair + 1) * ntiles / npairs);
    const int nsteps = (tile_hi - tile_lo + 1) / 2;           // step j: CTA r works on tile tile_lo + 2 j + r (if < tile_hi)

    if (tid < NF) sbias[tid] = bias[tid];
    if (warp == 5) cp_tmem_alloc2(tmem_slot, 256);
    if (tid == 128) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&full[s], 1); mbar_init(&empty[s], 1);
            mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 8);
        }
        mbar_init(wfull, 1);
        for (int s = 0; s < CP_XSTAGES; ++s) { mbar_init(&xfull[s], 1); mbar_init(&xout[s], 1); mbar_init(&xempty[s], 5); }
        mbar_fence_init();
        tma_prefetch_desc(&mapY);
        tma_prefetch_desc(&mapYlo);
        tma_prefetch_desc(&mapW);
        tma_prefetch_desc(&mapXin);
        tma_prefetch_desc(&mapXout);
    }
    tc_fence_before();
    __syncthreads();
    cp_cluster_sync();                          // the peer's barriers exist before anything is signalled across the pair
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();

    if (warp == 4) {
        // TMA producer
        if (elect_one_sync()) {
            if (rank == 0) mbar_expect_tx(wfull, 2 * CP_W_BYTES);
            for (int tap = 0; tap < 9; ++tap) cp_tma2_load_2d(sm + tap * NF * 128, &mapW, wfull, 0, (rank * 9 + tap) * NF);
        }
        pdl_wait();
        uint32_t xi = 0;
        for (int it = 0; it < nsteps; ++it) {
            const int tile = tile_lo + 2 * it + rank;
            const bool valid = tile < tile_hi;
            // a missing last tile is loaded from beyond the batch (all zeros): the pair's MMAs always cover both CTAs
            const int b = valid ? tile / per_img : B, r = valid ? tile - b * per_img : 0;
            const int y0 = (r / tiles_x) * CP_TH, x0 = (r % tiles_x) * CP_TW;
            const uint32_t s = it & 1, ph = (it >> 1) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            if (elect_one_sync()) {
                if (rank == 0) mbar_expect_tx(&full[s], 4 * CP_TILE_BYTES);
                cp_tma2_load_4d(sm + CP_OFF_A + s * CP_ASTAGE, &mapY, &full[s], 0, x0 - 1, y0 - 1, b);
                cp_tma2_load_4d(sm + CP_OFF_A + s * CP_ASTAGE + CP_STAGE, &mapYlo, &full[s], 0, x0 - 1, y0 - 1, b);
            }
            __syncwarp();
            if (valid) {
                for (int hh = 0; hh < 2; ++hh, ++xi) {
                    const uint32_t xs = xi % CP_XSTAGES, xph = (xi / CP_XSTAGES) & 1;
                    mbar_wait(&xempty[xs], xph ^ 1);
                    if (elect_one_sync()) {
                        mbar_expect_tx(&xfull[xs], CP_XHALF);
                        tma_load_4d(sm + CP_OFF_X + xs * CP_XHALF, &mapXin, &xfull[xs], 32 * hh, x0, y0, b);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 5) {
        // MMA issuer of the pair: the leader's warp only
        if (rank == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(256, NF);
            constexpr uint64_t tmpl_a = umma_smem_desc(0, 16, CP_HW * 128, UMMA_LAYOUT_SW128);
            constexpr uint64_t tmpl_b = umma_smem_desc(0, 16, 1024, UMMA_LAYOUT_SW128);
            mbar_wait(wfull, 0);
            for (int it = 0; it < nsteps; ++it) {
                const uint32_t s = it & 1, ph = (it >> 1) & 1;
                mbar_wait(&tempty[s], ph ^ 1);
                mbar_wait(&full[s], ph);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint64_t dh0 = umma_desc_at(tmpl_a, base + CP_OFF_A + s * CP_ASTAGE);
                    const uint64_t dl0 = umma_desc_at(tmpl_a, base + CP_OFF_A + s * CP_ASTAGE + CP_STAGE);
                    const uint64_t db0 = umma_desc_at(tmpl_b, base);
                    const uint32_t d_hi = tmem_base + s * 128, d_lo = d_hi + 64;
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t da = dh0 + (uint64_t)((((tap / 3) * CP_HW + (tap % 3)) * 128 + k * 32) >> 4);
                            const uint64_t db = db0 + (uint64_t)((tap * NF * 128 + k * 32) >> 4);
                            cp_mma2(d_hi, da, db, idesc, (tap | k) ? 1u : 0u);
                        }
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t da = dh0 + (uint64_t)((((tap / 3) * CP_HW + (tap % 3)) * 128 + k * 32) >> 4);
                            const uint64_t db = db0 + (uint64_t)((tap * NF * 128 + 32 * 128 + k * 32) >> 4);
                            cp_mma2(d_lo, da, db, idesc, (tap | k) ? 1u : 0u);
                        }
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t da = dl0 + (uint64_t)((((tap / 3) * CP_HW + (tap % 3)) * 128 + k * 32) >> 4);
                            const uint64_t db = db0 + (uint64_t)((tap * NF * 128 + k * 32) >> 4);
                            cp_mma2(d_lo, da, db, idesc, 1u);
                        }
                    cp_commit2(&empty[s]);
                    cp_commit2(&tfull[s]);
                }
                __syncwarp();
            }
        }
    } else if (warp == 6) {
        // TMA store warp
        pdl_wait();
        uint32_t xi = 0;
        for (int it = 0; it < nsteps; ++it) {
            const int tile = tile_lo + 2 * it + rank;
            if (tile >= tile_hi) break;
            const int b = tile / per_img, r = tile - b * per_img;
            const int y0 = (r / tiles_x) * CP_TH, x0 = (r % tiles_x) * CP_TW;
            for (int hh = 0; hh < 2; ++hh, ++xi) {
                const uint32_t xs = xi % CP_XSTAGES, xph = (xi / CP_XSTAGES) & 1;
                mbar_wait(&xout[xs], xph);
                if (elect_one_sync()) {
                    tma_store_4d(&mapXout, sm + CP_OFF_X + xs * CP_XHALF, 32 * hh, x0, y0, b);
                    tma_store_commit();
                    tma_store_wait_read();
                    mbar_arrive(&xempty[xs]);
                }
                __syncwarp();
            }
        }
        tma_store_wait_all();
    } else {
        // Epilogue: thread t = TMEM lane = tile pixel; statistics pass: thread = (channel of the half, quarter of the pixels)
        const int t = warp * 32 + lane;
        const uint32_t lanef = (uint32_t)(warp * 32) << 16;
        const int sc = t & 31, spart = t >> 5;
        const uint32_t sc_off = (uint32_t)(sc & 3) * 4, sc_chunk = (uint32_t)sc >> 2;
        pdl_wait();
        double dsum[2] = {0.0, 0.0}, dsq[2] = {0.0, 0.0};
        int cur_b = -1;
        uint32_t xi = 0;
        for (int it = 0; it < nsteps; ++it) {
            const int tile = tile_lo + 2 * it + rank;
            const bool valid = tile < tile_hi;
            const uint32_t acc = it & 1, aph = (it >> 1) & 1;
            if (!valid) {                                       // the pair's last step without a tile for this CTA
                mbar_wait(&tfull[acc], aph);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) cp_arrive_leader(&tempty[acc]);
                continue;
            }
            const int b = tile / per_img, r = tile - b * per_img;
            const int y0 = (r / tiles_x) * CP_TH, x0 = (r % tiles_x) * CP_TW;
            const long pix = ((long)b * Hp + (y0 + (t >> 3))) * Wp + (x0 + (t & 7));
            uint4 rv[16];                                       // last CFTM: this pixel's row of the head output
            if (xr != nullptr) {
#pragma unroll
                for (int j = 0; j < 8; ++j) ldg256(res + pix * NF + 8 * j, rv[2 * j], rv[2 * j + 1]);
            }
            if (b != cur_b) {                                   // image changed: publish the finished image's sums
                if (cur_b >= 0) {
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        atomicAdd(&stats[((long)cur_b * NF + 32 * hh + sc) * 2], dsum[hh]);
                        atomicAdd(&stats[((long)cur_b * NF + 32 * hh + sc) * 2 + 1], dsq[hh]);
                        dsum[hh] = 0.0; dsq[hh] = 0.0;
                    }
                }
                cur_b = b;
            }
            mbar_wait(&tfull[acc], aph);
            tc_fence_after();
#pragma unroll
            for (int hh = 0; hh < 2; ++hh, ++xi) {
                const uint32_t xs = xi % CP_XSTAGES, xph = (xi / CP_XSTAGES) & 1;
                uint8_t* xt = sm + CP_OFF_X + xs * CP_XHALF;
                uint32_t rr[32], rl[32];
                tmem_ld32(tmem_base + acc * 128 + hh * 32 + lanef, rr);
                tmem_ld32(tmem_base + acc * 128 + 64 + hh * 32 + lanef, rl);
                tmem_ld_wait();
                if (hh == 1) {                                  // both halves are in registers: the MMAs may refill
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) cp_arrive_leader(&tempty[acc]);
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) rr[i] = __float_as_uint(fmaf(__uint_as_float(rl[i]), 1.f / 2048.f, __uint_as_float(rr[i])));
                mbar_wait(&xfull[xs], xph);
                uint8_t* row = xt + t * 128;
                uint32_t hx[16];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4* cell = reinterpret_cast<float4*>(row + ((j ^ (t & 7)) << 4));
                    const float4 xv = *cell;
                    const float4 bv = *reinterpret_cast<const float4*>(sbias + hh * 32 + 4 * j);
                    float4 v;
                    v.x = __uint_as_float(rr[4 * j]) + bv.x + xv.x;
                    v.y = __uint_as_float(rr[4 * j + 1]) + bv.y + xv.y;
                    v.z = __uint_as_float(rr[4 * j + 2]) + bv.z + xv.z;
                    v.w = __uint_as_float(rr[4 * j + 3]) + bv.w + xv.w;
                    *cell = v;
                    if (xr != nullptr) {
                        const uint4 rq = rv[hh * 8 + j];
                        const __half2 h0 = __floats2half2_rn(v.x + __uint_as_float(rq.x), v.y + __uint_as_float(rq.y));
                        const __half2 h1 = __floats2half2_rn(v.z + __uint_as_float(rq.z), v.w + __uint_as_float(rq.w));
                        hx[2 * j] = *reinterpret_cast<const uint32_t*>(&h0);
                        hx[2 * j + 1] = *reinterpret_cast<const uint32_t*>(&h1);
                    }
                }
                if (xr != nullptr) {
                    __half* xp = xr + pix * NF + hh * 32;
                    stg256(xp, make_uint4(hx[0], hx[1], hx[2], hx[3]), make_uint4(hx[4], hx[5], hx[6], hx[7]));
                    stg256(xp + 16, make_uint4(hx[8], hx[9], hx[10], hx[11]), make_uint4(hx[12], hx[13], hx[14], hx[15]));
                }
                fence_proxy_async();                            // the half tile is read next by the TMA store
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (t == 0) mbar_arrive(&xout[xs]);
                // statistics of the finished half tile (fp32 per tile in a fixed order, fp64 across tiles)
                float ssum = 0.f, ssq = 0.f;
#pragma unroll 8
                for (int p = spart * 32; p < spart * 32 + 32; ++p) {
                    const float v = *reinterpret_cast<const float*>(xt + sc_off + p * 128 + ((sc_chunk ^ (uint32_t)(p & 7)) << 4));
                    ssum += v;
                    ssq = fmaf(v, v, ssq);
                }
                dsum[hh] += (double)ssum;
                dsq[hh] += (double)ssq;
                __syncwarp();
                if (lane == 0) mbar_arrive(&xempty[xs]);
            }
        }
        if (cur_b >= 0) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                atomicAdd(&stats[((long)cur_b * NF + 32 * hh + sc) * 2], dsum[hh]);
                atomicAdd(&stats[((long)cur_b * NF + 32 * hh + sc) * 2 + 1], dsq[hh]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cp_cluster_sync();                          // neither CTA leaves while the other can still signal it or read its operands
    if (warp == 5) cp_tmem_dealloc2(tmem_base, 256);
}

}  // namespace

// Same contract as launch_ffconv_umma_w2 (conv_umma.cu): Wpk2 = BlockW::ffw2, [2 halves][9 taps][32 Wh rows | 32 Wl rows][64]
int launch_ffconv_pair(const __half* Y, const __half* Ylo, const __half* Wpk2, const float* bias, const float* Xin,
                       float* Xout, double* stats, const Geom& g, cudaStream_t s, const float* res, __half* xr) {
    CUtensorMap mapY, mapYlo, mapW, mapXin, mapXout;
    {
        const uint64_t dims[4] = {NF, (uint64_t)g.Wp, (uint64_t)g.Hp, (uint64_t)g.B};
        const uint64_t str[4] = {2, NF * 2, (uint64_t)g.Wp * NF * 2, (uint64_t)g.Hp * g.Wp * NF * 2};
        const uint32_t box[4] = {NF, CP_HW, CP_TH + 2, 1};
        M2T_TRY(make_tensor_map(&mapY, Y, 2, 4, dims, str, box, 3));
        M2T_TRY(make_tensor_map(&mapYlo, Ylo, 2, 4, dims, str, box, 3));
    }
    {
        const uint64_t dims[2] = {NF, (uint64_t)2 * 9 * NF}, str[2] = {2, NF * 2};
        const uint32_t box[2] = {NF, NF};
        M2T_TRY(make_tensor_map(&mapW, Wpk2, 2, 2, dims, str, box, 3));
    }
    {
        const uint64_t dims[4] = {NF, (uint64_t)g.Wp, (uint64_t)g.Hp, (uint64_t)g.B};
        const uint64_t str[4] = {4, NF * 4, (uint64_t)g.Wp * NF * 4, (uint64_t)g.Hp * g.Wp * NF * 4};
        const uint32_t box[4] = {32, CP_TW, CP_TH, 1};
        M2T_TRY(make_tensor_map(&mapXin, Xin, 4, 4, dims, str, box, 3));
        M2T_TRY(make_tensor_map(&mapXout, Xout, 4, 4, dims, str, box, 3));
    }
    M2T_ENSURE_SMEM(ffconv_pair_kernel, CP_SMEM);
    const int ntiles = g.B * (g.Hp / CP_TH) * (g.Wp / CP_TW);
    int grid = device_sm_count() & ~1;
    const int want = 2 * ((ntiles + 1) / 2);
    if (grid > want) grid = want;
    M2T_CUDA(launch_pdl(ffconv_pair_kernel, dim3(grid), dim3(CP_THREADS), CP_SMEM, s, mapY, mapYlo, mapW, mapXin, mapXout, bias,
                        stats, g.B, g.Hp, g.Wp, res, xr));
    return M2T_OK;
}

}  // namespace m2t
