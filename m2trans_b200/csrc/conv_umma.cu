// tcgen05 implicit-GEMM form of CFTM.feed_forward + residual (ref M2Trans_network.py:124-126, :164):
//   Xout[p][o] = sum_{tap,c} Y[p + tap][c] * W[tap][o][c] + bias[o] + Xin[p][o]        (zero padding)
// One output tile = 16 rows x 8 pixels (M = 128), N = 64 accumulator columns, K = 9 taps x 64 channels.
//
// No im2col: TMA loads ONE 18 x 10 pixel halo tile of Y (fp16 NHWC, 128 B per pixel = one 128-byte swizzle
// row; out-of-frame pixels arrive as zeros = the conv's zero padding).  The A operand of tap (dy,dx) is the
// same shared-memory tile addressed through a descriptor whose start is shifted by (dy*10+dx) rows and whose
// 8-row groups are 10 rows apart (SBO = 1280 B): the 128-byte swizzle phase follows the absolute
// shared-memory address, so shifted descriptors read exactly what TMA wrote (tests/test_probes.py pins this).
// 36 MMAs (M=128, N=64, K=16) accumulate one tile in TMEM; two accumulators let the epilogue of tile i
// overlap the MMAs of tile i+1.  The 9 x 64 x 64 weight rows stay resident in shared memory (72 KB).
//
// The fp32 residual stream moves by TMA in BOTH directions.  The first version loaded the residual rows with LDG and
// stored the result with STG from the epilogue threads: with one epilogue warp per SM sub-partition the per-SM
// load/store queue, not HBM, set the pace (measured: 7 K cycles per tile, the tensor pipe idle 70 % of the time, DRAM at
// half its bandwidth).  Now the fp32 tile of Xin lands in shared memory as 128-byte-swizzled half tiles of 32 channels
// (3-stage ring), thread = pixel = TMEM lane adds accumulator + bias IN PLACE (conflict-free 16-byte accesses thanks
// to the swizzle), and a dedicated warp sends the tile back with a TMA store.  The InstanceNorm partial sums are a
// second pass over the finished tile in shared memory with thread = channel.
//
// Two weight formats (template parameter W2):
//   W2 = false  the 64 accumulator columns are the 64 output channels, fp16 weights [9][64][64]
//   W2 = true   split-precision weights (BlockW::ffw2): a CTA owns 32 output channels; accumulator columns 0..31 use
//               the fp16 weights, columns 32..63 their rounding residuals * 2^11, and the epilogue adds
//               hi + residual * 2^-11.  The A operand is split the same way: a second halo tile holds the rounding
//               residual of Y * 2^11 (written by the attention epilogue) and 36 more MMAs (N = 32, residual tile x fp16
//               weights) accumulate into columns 32..63.  fp16 rounding of the conv's weights and input were the two
//               largest error terms left after the residual-path fix (emulated at x3: 9e-5 and 8e-5 rms of 1.6e-4);
//               the tensor pipe had the room (29 % busy).
//
// Warp roles (224 threads): warps 0-3 epilogue, warp 4 TMA loads, warp 5 MMA issuer, warp 6 TMA stores.
//
// What bounds it (round 2, clock64 stamps + ncu at cfg2): shared-memory bandwidth.  Per 128-pixel tile the 36 MMAs read
// 147 KB of A (the halo tile, once per tap and K step) and 74 KB of B, TMA writes 55 KB (Y halo + X tile), the epilogue
// reads and rewrites the 32 KB X tile, the statistics pass reads it again and the TMA store reads it once more: ~400 KB
// = 3.2 K cycles at 128 B/clk, against 3.4-3.6 K cycles measured per tile (MMA issue 1.15 K, the MMA warp waits 1.3-2.2 K
// per tile for a free accumulator).  A second epilogue warpgroup (channels 32..63 of every pixel) was built and
// measured: the accumulate phase fell from 1.2 K to 0.75 K cycles, the statistics pass rose from 1.4 K to 2.0 K, the tile
// time and the forward (1.301 vs 1.310 ms) did not move -- not kept.  L2 evict_last hints on the fp32 residual stream
// (67 MB at cfg2, re-read by branch_prep_all and by the next ff conv) changed nothing either: at cfg2 the kernel runs at
// 41 % of DRAM bandwidth, at cfg4 (79 % of the copy bandwidth on algorithmic bytes) the stream cannot stay in L2.
// W2 at cfg3 (stamps): the MMA warp never waits and takes 3.5 K cycles per half tile = 48 cycles per MMA = its operand
// bytes (4 KB of A + 1-2 KB of B) at 128 B/clk.  conv_pair.cu (CTA pair, cta_group::2, each CTA its own tile and half of
// the weight rows) halves the A reads per tile, gives bit-identical results and is no faster: its M = 256 instructions
// take about twice as long each.
#include "common.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace m2t {

constexpr int CU_TH = 16, CU_TW = 8;                                  // output tile
constexpr int CU_HW = CU_TW + 2;                                      // halo tile width (10)
constexpr uint32_t CU_TILE_BYTES = (CU_TH + 2) * CU_HW * 128;         // 23040
constexpr uint32_t CU_STAGE = 23 * 1024;                              // 1024-aligned stage pitch
constexpr int CU_STAGES = 2;                                          // Y halo tiles (the MMAs run ahead of the epilogue)
constexpr int CU_XSTAGES = 3;                                         // fp32 residual tiles
constexpr uint32_t CU_XHALF = 128 * 128;                              // 128 pixels x 32 channels x 4 B
constexpr uint32_t CU_W_BYTES = 9 * NF * 128;                         // 73728
constexpr uint32_t CU_OFF_A = CU_W_BYTES;
constexpr int CU_THREADS = 224;

template <bool W2>
struct CuCfg {
    static constexpr int NCH = W2 ? 32 : 64;                          // output channels per CTA
    static constexpr int NHALF = NCH / 32;                            // 32-channel half tiles per residual tile
    static constexpr uint32_t XSTAGE = NHALF * CU_XHALF;
    static constexpr uint32_t ASTAGE = (W2 ? 2 : 1) * CU_STAGE;   // W2: the halo tile of Y and of its rounding residual
    static constexpr uint32_t OFF_X = CU_OFF_A + CU_STAGES * ASTAGE;
    static constexpr uint32_t OFF_BIAS = OFF_X + CU_XSTAGES * XSTAGE;
    static constexpr uint32_t OFF_BAR = OFF_BIAS + NF * 4;
    static constexpr uint32_t SMEM = 1024 + OFF_BAR + 256;
};

#ifdef M2T_TIMING
__device__ long long g_conv_dbg[64];   // CTA 0: epilogue thread 0 stamps [8i+0..4], MMA warp stamps [8i+5..7], tiles i < 8
#define M2T_CT(slot) do { if (xr == nullptr && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 5) && it < 8) g_conv_dbg[(slot) + 8 * it] = clock64(); } while (0)
#else
#define M2T_CT(slot) do { } while (0)
#endif

template <bool W2>
__global__ void __launch_bounds__(CU_THREADS, 1)
ffconv_umma_kernel(const __grid_constant__ CUtensorMap mapY, const __grid_constant__ CUtensorMap mapYlo,
                   const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapXin, const __grid_constant__ CUtensorMap mapXout,
                   const float* __restrict__ bias, double* __restrict__ stats, int B, int Hp, int Wp,
                   const float* __restrict__ res, __half* __restrict__ xr) {
    using CF = CuCfg<W2>;
    constexpr int NCH = CF::NCH, NHALF = CF::NHALF;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    float* sbias = reinterpret_cast<float*>(sm + CF::OFF_BIAS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + CF::OFF_BAR);
    uint64_t* full = bars;                        // [2]  Y halo tile landed
    uint64_t* empty = bars + 2;                   // [2]  ... consumed by the MMAs
    uint64_t* wfull = bars + 4;
    uint64_t* tfull = bars + 5;                   // [2]  accumulator complete
    uint64_t* tempty = bars + 7;                  // [2]  ... drained
    uint64_t* xfull = bars + 9;                   // [3]  residual tile landed
    uint64_t* xout = bars + 12;                   // [3]  result tile complete in smem (store may start)
    uint64_t* xempty = bars + 15;                 // [3]  tile buffer free (store has read it, stats pass done)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_x = Wp / CU_TW, tiles_y = Hp / CU_TH;
    const int per_img = tiles_x * tiles_y;
    const int ntiles = B * per_img;
    // W2: CTA c owns output channels (c & 1) * 32 .. and shares the tiles with the other CTAs of its parity
    const int chalf = W2 ? (int)(blockIdx.x & 1) : 0;
    const int c0 = chalf * 32;                                     // first output channel of this CTA
    const int part = W2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int nparts = W2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    // contiguous tile range per CTA: neighbouring tiles share halo rows in L2 and mostly belong to one image,
    // so the InstanceNorm partial sums are flushed once or twice per CTA instead of once per tile
    const int tile_lo = (int)((long)part * ntiles / nparts);
    const int tile_hi = (int)((long)(part + 1) * ntiles / nparts);

    if (tid < NF) sbias[tid] = bias[tid];
    if (warp == 5) tmem_alloc(tmem_slot, 128);
    if (tid == 128) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&full[s], 1); mbar_init(&empty[s], 1);
            mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 4);
        }
        mbar_init(wfull, 1);
        for (int s = 0; s < CU_XSTAGES; ++s) { mbar_init(&xfull[s], 1); mbar_init(&xout[s], 1); mbar_init(&xempty[s], 5); }
        mbar_fence_init();
        tma_prefetch_desc(&mapY);
        if (W2) tma_prefetch_desc(&mapYlo);
        tma_prefetch_desc(&mapW);
        tma_prefetch_desc(&mapXin);
        tma_prefetch_desc(&mapXout);
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
            for (int tap = 0; tap < 9; ++tap) tma_load_2d(sm + tap * NF * 128, &mapW, wfull, 0, (chalf * 9 + tap) * NF);
        }
        pdl_wait();
        uint32_t it = 0;
        for (int tile = tile_lo; tile < tile_hi; ++tile, ++it) {
            const int b = tile / per_img, r = tile - b * per_img;
            const int y0 = (r / tiles_x) * CU_TH, x0 = (r % tiles_x) * CU_TW;
            const uint32_t s = it & 1, ph = (it >> 1) & 1;
            const uint32_t xs = it % CU_XSTAGES, xph = (it / CU_XSTAGES) & 1;
            // Y first: its stage frees early (the MMAs run ahead), the residual stage only when the epilogue of
            // tile it-3 is done, and that wait must not hold back the tile the MMAs need next
            mbar_wait(&empty[s], ph ^ 1);
            if (elect_one_sync()) {
                mbar_expect_tx(&full[s], (W2 ? 2 : 1) * CU_TILE_BYTES);
                tma_load_4d(sm + CU_OFF_A + s * CF::ASTAGE, &mapY, &full[s], 0, x0 - 1, y0 - 1, b);
                if (W2) tma_load_4d(sm + CU_OFF_A + s * CF::ASTAGE + CU_STAGE, &mapYlo, &full[s], 0, x0 - 1, y0 - 1, b);
            }
            __syncwarp();
            mbar_wait(&xempty[xs], xph ^ 1);
            if (elect_one_sync()) {
                uint8_t* xt = sm + CF::OFF_X + xs * CF::XSTAGE;
                mbar_expect_tx(&xfull[xs], CF::XSTAGE);
                for (int hh = 0; hh < NHALF; ++hh) tma_load_4d(xt + hh * CU_XHALF, &mapXin, &xfull[xs], c0 + 32 * hh, x0, y0, b);
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
            const uint32_t s = it & 1, ph = (it >> 1) & 1;
            const uint32_t acc = it & 1, aph = (it >> 1) & 1;
            M2T_CT(5);
            mbar_wait(&tempty[acc], aph ^ 1);
            M2T_CT(6);
            mbar_wait(&full[s], ph);
            M2T_CT(7);
            tc_fence_after();
            if (elect_one_sync()) {
                const uint64_t da0 = umma_desc_at(tmpl_a, base + CU_OFF_A + s * CF::ASTAGE);
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
                if constexpr (W2) {   // residual tile x fp16 weight rows (the first 32 of each tap) -> columns 32..63
                    constexpr uint32_t idesc_lo = umma_idesc_f16(128, 32);
                    const uint64_t dl0 = umma_desc_at(tmpl_a, base + CU_OFF_A + s * CF::ASTAGE + CU_STAGE);
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t da = dl0 + (uint64_t)((((tap / 3) * CU_HW + (tap % 3)) * 128 + k * 32) >> 4);
                            const uint64_t db = db0 + (uint64_t)((tap * NF * 128 + k * 32) >> 4);
                            umma_f16_ss(tmem_base + acc * NF + 32, da, db, idesc_lo, 1u);
                        }
                    }
                }
                umma_commit(&empty[s]);
                umma_commit(&tfull[acc]);
            }
            __syncwarp();
        }
    } else if (warp == 6) {
        // TMA store warp: sends each finished tile back and releases its buffer once the store has read it
        pdl_wait();
        uint32_t it = 0;
        for (int tile = tile_lo; tile < tile_hi; ++tile, ++it) {
            const int b = tile / per_img, r = tile - b * per_img;
            const int y0 = (r / tiles_x) * CU_TH, x0 = (r % tiles_x) * CU_TW;
            const uint32_t xs = it % CU_XSTAGES, xph = (it / CU_XSTAGES) & 1;
            mbar_wait(&xout[xs], xph);
            if (elect_one_sync()) {
                const uint8_t* xt = sm + CF::OFF_X + xs * CF::XSTAGE;
                for (int hh = 0; hh < NHALF; ++hh) tma_store_4d(&mapXout, xt + hh * CU_XHALF, c0 + 32 * hh, x0, y0, b);
                tma_store_commit();
                tma_store_wait_read();
                mbar_arrive(&xempty[xs]);
            }
            __syncwarp();
        }
        tma_store_wait_all();                      // a no-op for lanes that issued nothing
    } else {
        // Epilogue: thread t = TMEM lane = tile pixel t (row t of every half tile)
        const int t = warp * 32 + lane;
        const uint32_t lanef = (uint32_t)(warp * 32) << 16;
        // statistics pass: thread = (channel, part of the pixels)
        constexpr int SPX = 128 * NCH / 128;                    // pixels per thread: 64 (NCH = 64) or 32
        const int sc = t % NCH, spart = t / NCH;
        const uint32_t sc_off = (uint32_t)(sc >> 5) * CU_XHALF + (uint32_t)(sc & 3) * 4;
        const uint32_t sc_chunk = (uint32_t)(sc & 31) >> 2;
        pdl_wait();
        // per-tile partial sums in fp32 (fixed order), accumulated across tiles in fp64: the statistics then do not depend
        // on which tiles a CTA happens to own, i.e. on the batch size (frames stay bit-independent of their batch)
        double dsum = 0.0, dsq = 0.0;
        int cur_b = -1;
        uint32_t it = 0;
        for (int tile = tile_lo; tile < tile_hi; ++tile, ++it) {
            const int b = tile / per_img, r = tile - b * per_img;
            const int y0 = (r / tiles_x) * CU_TH, x0 = (r % tiles_x) * CU_TW;
            const uint32_t acc = it & 1, aph = (it >> 1) & 1;
            const uint32_t xs = it % CU_XSTAGES, xph = (it / CU_XSTAGES) & 1;
            uint8_t* xt = sm + CF::OFF_X + xs * CF::XSTAGE;
            const long pix = ((long)b * Hp + (y0 + (t >> 3))) * Wp + (x0 + (t & 7));
            M2T_CT(0);
            uint4 rv[NCH / 4];                                  // last CFTM: this pixel's row of the head output
            if (xr != nullptr) {
#pragma unroll
                for (int j = 0; j < NCH / 8; ++j) ldg256(res + pix * NF + c0 + 8 * j, rv[2 * j], rv[2 * j + 1]);
            }
            if (b != cur_b) {                                   // image changed: publish the finished image's sums
                if (cur_b >= 0) {
                    atomicAdd(&stats[((long)cur_b * NF + c0 + sc) * 2], dsum);
                    atomicAdd(&stats[((long)cur_b * NF + c0 + sc) * 2 + 1], dsq);
                    dsum = 0.0; dsq = 0.0;
                }
                cur_b = b;
            }
            M2T_CT(1);
            mbar_wait(&tfull[acc], aph);
            mbar_wait(&xfull[xs], xph);
            tc_fence_after();
            M2T_CT(2);
#pragma unroll
            for (int hh = 0; hh < NHALF; ++hh) {
                uint32_t rr[32];
                tmem_ld32(tmem_base + acc * NF + hh * 32 + lanef, rr);
                if constexpr (W2) {                            // columns 32..63: the same channels through the residual rows
                    uint32_t rl[32];
                    tmem_ld32(tmem_base + acc * NF + 32 + lanef, rl);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) rr[i] = __float_as_uint(fmaf(__uint_as_float(rl[i]), 1.f / 2048.f, __uint_as_float(rr[i])));
                } else {
                    tmem_ld_wait();
                }
                uint8_t* row = xt + hh * CU_XHALF + t * 128;
                uint32_t hx[16];                               // fp16(res + x) of these 32 channels (last CFTM)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4* cell = reinterpret_cast<float4*>(row + ((j ^ (t & 7)) << 4));
                    const float4 xv = *cell;
                    const float4 bv = *reinterpret_cast<const float4*>(sbias + c0 + hh * 32 + 4 * j);
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
                if (xr != nullptr) {   // fp16(res + x), the tail's first GEMM operand (ref :70): 64 B per half
                    __half* xp = xr + pix * NF + c0 + hh * 32;
                    stg256(xp, make_uint4(hx[0], hx[1], hx[2], hx[3]), make_uint4(hx[4], hx[5], hx[6], hx[7]));
                    stg256(xp + 16, make_uint4(hx[8], hx[9], hx[10], hx[11]), make_uint4(hx[12], hx[13], hx[14], hx[15]));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);          // accumulator drained: MMA may reuse it
            fence_proxy_async();                               // the tile is read next by the TMA store
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (t == 0) mbar_arrive(&xout[xs]);
            M2T_CT(3);
            // statistics of the finished tile: conflict-free row reads (a warp reads 32 channels of one pixel)
            float ssum = 0.f, ssq = 0.f;
#pragma unroll 8
            for (int p = spart * SPX; p < spart * SPX + SPX; ++p) {
                const float v = *reinterpret_cast<const float*>(xt + sc_off + p * 128 + ((sc_chunk ^ (uint32_t)(p & 7)) << 4));
                ssum += v;
                ssq = fmaf(v, v, ssq);
            }
            dsum += (double)ssum;
            dsq += (double)ssq;
            __syncwarp();
            if (lane == 0) mbar_arrive(&xempty[xs]);
            M2T_CT(4);
        }
        if (cur_b >= 0) {
            atomicAdd(&stats[((long)cur_b * NF + c0 + sc) * 2], dsum);
            atomicAdd(&stats[((long)cur_b * NF + c0 + sc) * 2 + 1], dsq);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, 128);
}

#ifdef M2T_TIMING
int read_conv_timing(long long* host64) {
    M2T_CUDA(cudaMemcpyFromSymbol(host64, g_conv_dbg, sizeof(long long) * 64));
    return M2T_OK;
}
#else
int read_conv_timing(long long* host64) { memset(host64, 0, sizeof(long long) * 64); return M2T_OK; }
#endif

template <bool W2>
static int launch_ffconv_umma_t(const __half* Y, const __half* Ylo, const __half* Wpk, const float* bias, const float* Xin,
                                float* Xout, double* stats, const Geom& g, cudaStream_t s, const float* res, __half* xr) {
    using CF = CuCfg<W2>;
    CUtensorMap mapY, mapYlo, mapW, mapXin, mapXout;
    {
        const uint64_t dims[4] = {NF, (uint64_t)g.Wp, (uint64_t)g.Hp, (uint64_t)g.B};
        const uint64_t str[4] = {2, NF * 2, (uint64_t)g.Wp * NF * 2, (uint64_t)g.Hp * g.Wp * NF * 2};
        const uint32_t box[4] = {NF, CU_HW, CU_TH + 2, 1};
        M2T_TRY(make_tensor_map(&mapY, Y, 2, 4, dims, str, box, 3));
        M2T_TRY(make_tensor_map(&mapYlo, W2 ? Ylo : Y, 2, 4, dims, str, box, 3));
    }
    {   // weight rows: [9][64] (W2: [2][9][64]) x 64 input channels
        const uint64_t dims[2] = {NF, (uint64_t)(W2 ? 2 : 1) * 9 * NF}, str[2] = {2, NF * 2};
        const uint32_t box[2] = {NF, NF};
        M2T_TRY(make_tensor_map(&mapW, Wpk, 2, 2, dims, str, box, 3));
    }
    {   // fp32 residual stream: 32-channel (128-byte) boxes of 16 x 8 pixels
        const uint64_t dims[4] = {NF, (uint64_t)g.Wp, (uint64_t)g.Hp, (uint64_t)g.B};
        const uint64_t str[4] = {4, NF * 4, (uint64_t)g.Wp * NF * 4, (uint64_t)g.Hp * g.Wp * NF * 4};
        const uint32_t box[4] = {32, CU_TW, CU_TH, 1};
        M2T_TRY(make_tensor_map(&mapXin, Xin, 4, 4, dims, str, box, 3));
        M2T_TRY(make_tensor_map(&mapXout, Xout, 4, 4, dims, str, box, 3));
    }
    M2T_ENSURE_SMEM(ffconv_umma_kernel<W2>, CF::SMEM);
    const int ntiles = g.B * (g.Hp / CU_TH) * (g.Wp / CU_TW);
    int grid = device_sm_count();
    if (W2) {
        grid &= ~1;                                           // CTA pairs: one per channel half
        if (grid > 2 * ntiles) grid = 2 * ntiles;
    } else if (grid > ntiles) {
        grid = ntiles;
    }
    M2T_CUDA(launch_pdl(ffconv_umma_kernel<W2>, dim3(grid), dim3(CU_THREADS), CF::SMEM, s, mapY, mapYlo, mapW, mapXin, mapXout, bias,
                        stats, g.B, g.Hp, g.Wp, res, xr));
    return M2T_OK;
}

int launch_ffconv_umma(const __half* Y, const __half* Wpk, const float* bias, const float* Xin, float* Xout,
                       double* stats, const Geom& g, cudaStream_t s, const float* res, __half* xr) {
    return launch_ffconv_umma_t<false>(Y, nullptr, Wpk, bias, Xin, Xout, stats, g, s, res, xr);
}

int launch_ffconv_umma_w2(const __half* Y, const __half* Ylo, const __half* Wpk2, const float* bias, const float* Xin,
                          float* Xout, double* stats, const Geom& g, cudaStream_t s, const float* res, __half* xr) {
    return launch_ffconv_umma_t<true>(Y, Ylo, Wpk2, bias, Xin, Xout, stats, g, s, res, xr);
}

}  // namespace m2t
