// Fused last stage of the tail for a x2 PixelShuffle (ref M2Trans_network.py:46-55, :72-76), strip-marching form:
//     A [B,h,w,64] fp16  --1x1 conv 64->256 + bias--> PixelShuffle(2) --GELU--> U [B,2h,2w,64]
//                        --3x3 reflect conv 64->3 (no bias)--> clamp --> crop --> y fp32 NCHW
// Same contract as tail_fused.cu (which tiles the frame in 16 x 12 blocks, recomputes a 1.31x halo of GELUs and runs
// its GEMM / GELU / conv / gather phases one after the other: 48 % FMA-pipe utilisation in ncu).  Here a CTA walks DOWN
// a strip of 30 A columns (+1 halo column each side = 32 = one 128-byte-swizzled TMA row block) in steps of 4 A rows:
//   * no vertical halo: U rows are produced once, the only recompute is the 2/30 column halo (1.07x);
//   * the phases of consecutive steps overlap, each on its own warps and its own pipe:
//       warp 20       TMA: A tile of the step (4 x 32 px x 128 B), two stages
//       warp 21       MMA: GEMM-1 as two N = 128 halves (even / odd U rows) into two TMEM accumulators, then the conv
//                     GEMM of the PREVIOUS step: D[px][tap*3+c] = U[px][:] . Wc[tap][c][:]  (N = 32, every U pixel read once)
//       warps 0-15    E1: accumulator -> +bias -> GELU (packed fp32x2, the FMA-pipe-bound part) -> fp16 -> U buffer
//                     (two buffers of 8 U rows x 64 px; PixelShuffle is the store address; reflected border copies)
//       warps 16-19   E2: D from TMEM -> 3-tap horizontal sums by warp shuffles -> a 12-row ring of 9 partial-sum planes
//                     in shared memory -> vertical 3-row sum one row behind -> clamp -> crop -> y
//     The U buffer layout puts the left and right 32 columns of a U row into different M-tiles at the SAME TMEM lanes,
//     so one E2 warp owns whole U rows and the horizontal neighbours are a shuffle away.
// TMEM: GEMM-1 halves [0,128) [128,256) | D of step parity 0 / 1 [256,384) [384,512).
// Shared memory: W1 32 KB | A 2 x 16 KB | U 2 x 64 KB | Wc 4 KB | ring 27 KB | bias 1 KB = 225 KB.
#include "common.cuh"
#include "gelu.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace m2t {
namespace {

constexpr int TS_TW = 30;                               // interior A columns of a strip
constexpr int TS_AW = 32;                               // with halo
constexpr int TS_AR = 4;                                // A rows per step (M = 128)
constexpr int TS_RING = 12;                             // rows of the partial-sum ring
constexpr uint32_t TS_OFF_W1 = 0;
constexpr uint32_t TS_OFF_A = 32768;
constexpr uint32_t TS_ASTAGE = 16384;
constexpr uint32_t TS_OFF_U = 65536;
constexpr uint32_t TS_UBUF = 65536;                     // 4 M-tiles x 16 KB
constexpr uint32_t TS_OFF_WC = TS_OFF_U + 2 * TS_UBUF;  // 196608
constexpr uint32_t TS_OFF_RING = TS_OFF_WC + 4096;      // 200704
constexpr uint32_t TS_RING_ROW = 9 * 64 * 4;            // 2304 B per U row
constexpr uint32_t TS_OFF_BIAS = TS_OFF_RING + TS_RING * TS_RING_ROW;   // 228352
constexpr uint32_t TS_OFF_BAR = TS_OFF_BIAS + 1024;
constexpr uint32_t TS_SMEM = 1024 + TS_OFF_BAR + 256;
static_assert(TS_SMEM <= 232448, "tail_strip exceeds the shared memory of an SM");
constexpr int TS_E1 = 512, TS_E2 = 128;
constexpr int TS_THREADS = TS_E1 + TS_E2 + 64;          // + warp 20 TMA, warp 21 MMA
constexpr uint32_t TS_COL_D = 256;
#ifndef TS_IDLE_NS
#define TS_IDLE_NS 200
#endif
#ifndef TS_MMA_NS
#define TS_MMA_NS 200                                    // the MMA warp's polls (E1 waits for its GEMM-1 issues)
#endif

struct TsItem { int bl, x0, ya, nsteps, yhi; };

// item -> image, strip, row segment.  hn = A rows that are needed at all (crop), SH = segment height (multiple of 4)
__device__ __forceinline__ TsItem ts_item(int item, int nstrips, int nseg, int SH, int hn, int hout) {
    TsItem t;
    const int per_img = nstrips * nseg;
    t.bl = item / per_img;
    const int r = item - t.bl * per_img;
    const int seg = r / nstrips;
    t.x0 = (r - seg * nstrips) * TS_TW;
    t.ya = seg * SH;
    const int she = hn - t.ya < SH ? hn - t.ya : SH;
    t.nsteps = she / TS_AR + 1;
    const int yh = 2 * (t.ya + she);
    t.yhi = yh < hout ? yh : hout;
    return t;
}

__global__ void __launch_bounds__(TS_THREADS, 1)
tail_strip_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW,
                  const float* __restrict__ bias, const __half* __restrict__ wc, float* __restrict__ y,
                  int Bc, int h, int w, int hout, int wout, int b0, float rgb_range, int nstrips, int nseg, int SH, int hn) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    float* sbias = reinterpret_cast<float*>(sm + TS_OFF_BIAS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + TS_OFF_BAR);
    uint64_t* w_full = bars;
    uint64_t* a_full = bars + 1;      // [2]
    uint64_t* a_empty = bars + 3;     // [2]
    uint64_t* acc_full = bars + 5;    // [2] one per N half
    uint64_t* acc_empty = bars + 7;   // [2]
    uint64_t* u_full = bars + 9;      // [2]
    uint64_t* u_empty = bars + 11;    // [2]
    uint64_t* d_full = bars + 13;     // [2]
    uint64_t* d_empty = bars + 15;    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nitems = Bc * nstrips * nseg;

    // constants: conv weights [9][16][64] (rows 0..2 of each tap real) -> [32][64] rows tap*3+c, 128-B swizzled; bias
    if (tid < 256) {
        const int row = tid >> 3, ch = tid & 7;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (row < 27) v = *reinterpret_cast<const uint4*>(wc + ((row / 3) * 16 + (row % 3)) * NF + ch * 8);
        *reinterpret_cast<uint4*>(sm + TS_OFF_WC + row * 128 + ((ch ^ (row & 7)) << 4)) = v;
        sbias[tid] = bias[tid];
    }
    // the U buffers start as zeros: pixels outside the frame are never written and must stay finite
    for (uint32_t i = tid * 16; i < 2 * TS_UBUF; i += TS_THREADS * 16) *reinterpret_cast<uint4*>(sm + TS_OFF_U + i) = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    if (warp == 21) tmem_alloc(tmem_slot, 512);
    if (tid == TS_E1 + TS_E2) {
        mbar_init(w_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1);
            mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], TS_E1 / 32);
            mbar_init(&u_full[s], TS_E1 / 32); mbar_init(&u_empty[s], 1);
            mbar_init(&d_full[s], 1); mbar_init(&d_empty[s], TS_E2 / 32);
        }
        mbar_fence_init();
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapW);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();                              // persistent one-wave grid: whatever follows may be launched now, it cannot take our SMs

    if (warp == 20) {
        // ---- TMA producer ---------------------------------------------------------------------------------
        if (elect_one_sync()) {
            mbar_expect_tx(w_full, 256 * 128);
            tma_load_2d(sm + TS_OFF_W1, &mapW, w_full, 0, 0);
        }
        pdl_wait();
        uint32_t g = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const TsItem ti = ts_item(item, nstrips, nseg, SH, hn, hout);
            for (int s = 0; s < ti.nsteps; ++s, ++g) {
                const uint32_t st = g & 1;
                mbar_wait_idle(&a_empty[st], ((g >> 1) & 1) ^ 1, TS_IDLE_NS);
                if (elect_one_sync()) {
                    mbar_expect_tx(&a_full[st], TS_ASTAGE);
                    tma_load_4d(sm + TS_OFF_A + st * TS_ASTAGE, &mapA, &a_full[st], 0, ti.x0 - 1, ti.ya + TS_AR * s - 1, ti.bl);
                }
                __syncwarp();
            }
        }
    } else if (warp == 21) {
        // ---- MMA issuer -----------------------------------------------------------------------------------
        constexpr uint32_t idesc1 = umma_idesc_f16(128, 128);
        constexpr uint32_t idesc2 = umma_idesc_f16(128, 32);
        constexpr uint64_t tmpl = umma_smem_desc(0, 16, 1024, UMMA_LAYOUT_SW128);
        auto conv = [&](uint32_t j) {               // conv GEMM of step j: D[j & 1] = U[j & 1] . Wc^T
            const uint32_t buf = j & 1, ph = (j >> 1) & 1;
            mbar_wait_idle(&u_full[buf], ph, TS_MMA_NS);
            mbar_wait_idle(&d_empty[buf], ph ^ 1, TS_MMA_NS);
            tc_fence_after();
            if (elect_one_sync()) {
                const uint64_t db0 = umma_desc_at(tmpl, base + TS_OFF_WC);
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    const uint64_t da0 = umma_desc_at(tmpl, base + TS_OFF_U + buf * TS_UBUF + m * 16384);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_f16_ss(tmem_base + TS_COL_D + buf * 128 + m * 32, da0 + 2 * k, db0 + 2 * k, idesc2, k ? 1u : 0u);
                }
                umma_commit(&d_full[buf]);
                umma_commit(&u_empty[buf]);
            }
            __syncwarp();
        };
        mbar_wait(w_full, 0);
        uint32_t g = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const TsItem ti = ts_item(item, nstrips, nseg, SH, hn, hout);
            for (int s = 0; s < ti.nsteps; ++s, ++g) {
                const uint32_t st = g & 1;
                mbar_wait_idle(&a_full[st], (g >> 1) & 1, TS_MMA_NS);
                for (int hf = 0; hf < 2; ++hf) {
                    mbar_wait_idle(&acc_empty[hf], (g & 1) ^ 1, TS_MMA_NS);
                    tc_fence_after();
                    if (elect_one_sync()) {
                        const uint64_t da0 = umma_desc_at(tmpl, base + TS_OFF_A + st * TS_ASTAGE);
                        const uint64_t db0 = umma_desc_at(tmpl, base + TS_OFF_W1 + hf * 16384);
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base + hf * 128, da0 + 2 * k, db0 + 2 * k, idesc1, k ? 1u : 0u);
                        umma_commit(&acc_full[hf]);
                        if (hf == 1) umma_commit(&a_empty[st]);
                    }
                    __syncwarp();
                }
                if (g > 0) conv(g - 1);
            }
        }
        if (g > 0) conv(g - 1);
    } else if (warp < 16) {
        // ---- E1: GEMM-1 accumulator -> bias + GELU -> U buffer ------------------------------------------------------
        // thread = A pixel (row quad, column lane) of the step; warpgroup wg serves sub-pixel column vv and 32 channels
        const int quad = warp & 3, wg = warp >> 2;
        const int vv = wg >> 1, c0 = (wg & 1) * 32;
        const uint32_t lanef = (uint32_t)(quad * 32) << 16;
        const int H2 = 2 * h, W2 = 2 * w;
        // U pixel (urow, ucol) of a step -> byte offset in its buffer: M-tile (urow / 4, ucol / 32), row (urow % 4) * 32 + p(ucol % 32)
        // with p(u) = u / 2 + 16 (u & 1): even columns first, then the odd ones.  A warp writes the columns 2 lane + vv, i.e.
        // rows (lane & 15) + 16 vv: eight consecutive lanes have eight different swizzle keys (row & 7) and their 16-byte stores
        // hit eight different bank groups (with p = identity they shared four: a 2-way conflict on every store, ncu).
        auto uoff = [](int urow, int ucol) -> uint32_t {
            const int u = ucol & 31;
            return (uint32_t)(((urow >> 2) * 2 + (ucol >> 5)) * 16384 + ((urow & 3) * 32 + (u >> 1) + 16 * (u & 1)) * 128);
        };
        uint32_t g = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const TsItem ti = ts_item(item, nstrips, nseg, SH, hn, hout);
            const int ax = ti.x0 - 1 + lane;
            const int X = 2 * ax + vv, ucol = 2 * lane + vv;
            int dxr = X == 1 ? -2 : (X == W2 - 2 ? 2 : 0);                 // reflected copy: -1 <- 1, W2 <- W2-2
            if ((unsigned)(ucol + dxr) >= 64u) dxr = 0;
            for (int s = 0; s < ti.nsteps; ++s, ++g) {
                const int ay = ti.ya + TS_AR * s - 1 + quad;
                const bool valid = ay >= 0 && ay < h && ax >= 0 && ax < w;
                const uint32_t buf = g & 1;
                uint8_t* ub = sm + TS_OFF_U + buf * TS_UBUF;
                // Measured and not kept: starting half of the warps on the odd rows (220 vs 213 us at cfg2) and reading the
                // accumulator in 16-column chunks one chunk ahead of the GELU (248 us).  ncu: the E1 warps sit in the GELU
                // code 96 % of the time with `math` (FMA-pipe throttle) and `not selected` as their stalls; the kernel is
                // bound by instruction issue (316 K warp instructions per sub-partition at cfg2, 70 % issue utilisation).
#pragma unroll 1
                for (int hf = 0; hf < 2; ++hf) {                               // hf = sub-pixel row uu
                    mbar_wait(&acc_full[hf], g & 1);
                    if (hf == 0) mbar_wait(&u_empty[buf], ((g >> 1) & 1) ^ 1);
                    tc_fence_after();
                    uint32_t rr[32];
                    tmem_ld32(tmem_base + hf * 128 + vv * NF + c0 + lanef, rr);
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[hf]);               // values are in registers: the MMA may refill
                    const int urow = 2 * quad + hf, Y = 2 * ay + hf;
                    int dyr = Y == 1 ? -2 : (Y == H2 - 2 ? 2 : 0);
                    if ((unsigned)(urow + dyr) >= 8u) dyr = 0;
                    // own pixel and (frame border rows / columns only) its reflected copies: base addresses and swizzle keys
                    const uint32_t o00 = uoff(urow, ucol);
                    uint8_t* p00 = ub + o00;
                    const uint32_t k00 = (o00 >> 7) & 7;
                    const bool refl = valid && (dyr | dxr) != 0;
                    const uint32_t ch0 = (uint32_t)(c0 >> 3);
                    const float* bs = sbias + (hf * 2 + vv) * NF + c0;
                    uint4 q[4];
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        uint32_t* pq = reinterpret_cast<uint32_t*>(&q[v]);
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            pq[e] = gelu_pair_h2(f2_pack(__uint_as_float(rr[v * 8 + 2 * e]), __uint_as_float(rr[v * 8 + 2 * e + 1])),
                                                 *reinterpret_cast<const uint64_t*>(bs + v * 8 + 2 * e));
                        if (valid) *reinterpret_cast<uint4*>(p00 + (((ch0 + v) ^ k00) << 4)) = q[v];
                    }
                    if (refl) {
#pragma unroll 1
                        for (int t = 1; t < 4; ++t) {                 // t bit 0: row copy, bit 1: column copy
                            if (((t & 1) && dyr == 0) || ((t & 2) && dxr == 0)) continue;
                            const uint32_t o = uoff(urow + ((t & 1) ? dyr : 0), ucol + ((t & 2) ? dxr : 0));
                            const uint32_t k = (o >> 7) & 7;
#pragma unroll
                            for (int v = 0; v < 4; ++v) *reinterpret_cast<uint4*>(ub + o + (((ch0 + v) ^ k) << 4)) = q[v];
                        }
                    }
                }
                fence_proxy_async();            // U written by the generic proxy, read by the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive(&u_full[buf]);
            }
        }
    } else {
        // ---- E2: conv accumulators -> horizontal sums -> ring -> vertical sums -> y ---------------------------------
        // warp quad owns U rows quad and quad + 4 of every step, both 32-column halves (same TMEM lanes, two M-tiles)
        const int quad = warp & 3;
        const uint32_t lanef = (uint32_t)(quad * 32) << 16;
        float* ring = reinterpret_cast<float*>(sm + TS_OFF_RING);
        auto slot = [](int R) -> int { return (int)((unsigned)(R + 2 * TS_RING) % (unsigned)TS_RING) * (9 * 64); };
        const long plane = (long)hout * wout;
        // TMEM lane j of an M-tile holds U column u(j) of its 32-column half (see uoff): even columns in lanes 0..15, odd in 16..31
        const int ucl = lane < 16 ? 2 * lane : 2 * (lane - 16) + 1;
        const int src_l = lane < 16 ? (lane + 15) : lane - 16;          // lane holding column u - 1 (lane 0: none, patched below)
        const int src_r = lane < 16 ? lane + 16 : (lane - 15) & 31;     // lane holding column u + 1 (lane 31: none, patched below)
        uint32_t g = 0;
        pdl_wait();
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const TsItem ti = ts_item(item, nstrips, nseg, SH, hn, hout);
            const int ylo = 2 * ti.ya;
            float* yimg = y + (long)(b0 + ti.bl) * 3 * plane;
            for (int s = 0; s < ti.nsteps; ++s, ++g) {
                const uint32_t buf = g & 1;
                const int R0 = 2 * (ti.ya + TS_AR * s - 1);                    // U row of this step's first row
                mbar_wait_idle(&d_full[buf], (g >> 1) & 1, TS_IDLE_NS);
                tc_fence_after();
#pragma unroll 1
                for (int r2 = 0; r2 < 2; ++r2) {
                    const int urow = quad + 4 * r2;
                    uint32_t dl[32], dr[32];
                    const uint32_t t0 = tmem_base + TS_COL_D + buf * 128 + (uint32_t)(r2 * 2) * 32 + lanef;
                    tmem_ld32(t0, dl);
                    tmem_ld32(t0 + 32, dr);
                    tmem_ld_wait();
                    float* rw = ring + slot(R0 + urow);
#pragma unroll
                    for (int dyi = 0; dyi < 3; ++dyi) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const float l0 = __uint_as_float(dl[(dyi * 3 + 0) * 3 + c]), l1 = __uint_as_float(dl[(dyi * 3 + 1) * 3 + c]);
                            const float l2 = __uint_as_float(dl[(dyi * 3 + 2) * 3 + c]);
                            const float r0 = __uint_as_float(dr[(dyi * 3 + 0) * 3 + c]), r1 = __uint_as_float(dr[(dyi * 3 + 1) * 3 + c]);
                            const float r2v = __uint_as_float(dr[(dyi * 3 + 2) * 3 + c]);
                            // out[X] sums D[X-1][tap dx=-1] + D[X][tap dx=0] + D[X+1][tap dx=+1]
                            const float fl_l = __shfl_sync(0xffffffffu, l0, src_l);
                            float fr_l = __shfl_sync(0xffffffffu, l2, src_r);
                            const float r2_first = __shfl_sync(0xffffffffu, r2v, 0);       // column 32 = column 0 of the right half
                            if (lane == 31) fr_l = r2_first;
                            float fl_r = __shfl_sync(0xffffffffu, r0, src_l);
                            const float l0_last = __shfl_sync(0xffffffffu, l0, 31);        // column 31 of the left half
                            if (lane == 0) fl_r = l0_last;
                            const float fr_r = __shfl_sync(0xffffffffu, r2v, src_r);
                            rw[(dyi * 3 + c) * 64 + ucl] = fl_l + l1 + fr_l;
                            rw[(dyi * 3 + c) * 64 + 32 + ucl] = fl_r + r1 + fr_r;
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&d_empty[buf]);
                asm volatile("bar.sync 2, 128;" ::: "memory");
                // output rows R0-1 .. R0+6 (one row behind the newest partial sums)
#pragma unroll 1
                for (int j = 0; j < 2; ++j) {
                    const int Y = R0 - 1 + 4 * j + quad;
                    if (Y >= ylo && Y < ti.yhi) {
                        const float* ra = ring + slot(Y - 1);
                        const float* rb = ring + slot(Y);
                        const float* rc = ring + slot(Y + 1);
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            const int uc = hh * 32 + lane;
                            const int Xg = 2 * (ti.x0 - 1) + uc;
                            if (uc >= 2 && uc < 62 && Xg < wout) {
                                float* yp = yimg + (long)Y * wout + Xg;
#pragma unroll
                                for (int c = 0; c < 3; ++c) {
                                    const float v = ra[c * 64 + uc] + rb[(3 + c) * 64 + uc] + rc[(6 + c) * 64 + uc];
                                    yp[c * plane] = fminf(fmaxf(v, 0.f), rgb_range);
                                }
                            }
                        }
                    }
                }
                asm volatile("bar.sync 2, 128;" ::: "memory");     // ring rows of this step are consumed before the next step writes
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 21) tmem_dealloc(tmem_base, 512);
}

}  // namespace

// A: fp16 [Bc][h][w][64]; W1: fp16 [256][64] sub-pixel-major; bias fp32 [256]; Wc: fp16 [9][16][64];
// y: fp32 NCHW, images b0.. cropped to hout x wout (hout <= 2h, wout <= 2w)
int launch_tail_strip(const __half* A, const __half* W1, const float* bias, const __half* Wc, float* y, int Bc, int h,
                      int w, int hout, int wout, int b0, float rgb_range, cudaStream_t s) {
    if (h % TS_AR || hout > 2 * h || wout > 2 * w || hout < 1 || wout < 1) {
        set_error("tail_strip: %dx%d -> %dx%d is not a x2 stage on a multiple of %d rows", h, w, hout, wout, TS_AR);
        return M2T_E_ARG;
    }
    CUtensorMap mapA, mapW;
    {
        const uint64_t dims[4] = {NF, (uint64_t)w, (uint64_t)h, (uint64_t)Bc};
        const uint64_t str[4] = {2, NF * 2, (uint64_t)w * NF * 2, (uint64_t)h * w * NF * 2};
        const uint32_t box[4] = {NF, TS_AW, TS_AR, 1};
        M2T_TRY(make_tensor_map(&mapA, A, 2, 4, dims, str, box, 3));
    }
    {
        const uint64_t dims[2] = {NF, 256}, str[2] = {2, NF * 2};
        const uint32_t box[2] = {NF, 256};
        M2T_TRY(make_tensor_map(&mapW, W1, 2, 2, dims, str, box, 3));
    }
    const int nstrips = (wout + 2 * TS_TW - 1) / (2 * TS_TW);
    int hn = ((hout + 1) / 2 + TS_AR - 1) / TS_AR * TS_AR;          // A rows whose outputs survive the crop
    if (hn > h) hn = h;
    // row segments: whole columns when there are enough strips to fill the SMs, else shorter ones (each extra segment
    // costs one step of halo rows)
    const int sms = device_sm_count();
    int nseg = 1;
    while (Bc * nstrips * nseg < sms && hn / (nseg * 2) >= 4 * TS_AR) nseg *= 2;
    int SH = ((hn + nseg - 1) / nseg + TS_AR - 1) / TS_AR * TS_AR;
    nseg = (hn + SH - 1) / SH;
    const int nitems = Bc * nstrips * nseg;
    M2T_ENSURE_SMEM(tail_strip_kernel, TS_SMEM);
    const int grid = nitems < sms ? nitems : sms;
    M2T_CUDA(launch_pdl(tail_strip_kernel, dim3(grid), dim3(TS_THREADS), TS_SMEM, s, mapA, mapW, bias, Wc, y, Bc, h, w,
                        hout, wout, b0, rgb_range, nstrips, nseg, SH, hn));
    return M2T_OK;
}

}  // namespace m2t
