// tcgen05 bf16 GEMM for the Linear layers of the MedCLIP image tower (Swin-T):
//   Y[m][n] = epilogue( sum_k A[m][k] * W[n][k] + bias[n] )      A bf16 [M][K], W bf16 [N][K] (nn.Linear layout)
// Persistent, warp-specialised CTAs (320 threads), tiles of 128 rows x 128 columns, K in blocks of 64 (128-byte swizzle):
//   warp 8   : TMA producer, ring of LG_STAGES {A block, W block} stages.  K, M and N need not be multiples of the tile:
//              the tensor maps zero-fill out-of-range elements (K = 96 -> a 64 block and a 32 + 32 zero block; only the
//              MMAs that touch real columns are issued), and the TMA stores clip rows >= M and columns >= N
//   warp 9   : single-thread tcgen05.mma issue, M=128 x N=128 x K=16, fp32 accumulators in TMEM (two of 128 columns: the
//              epilogue of tile i overlaps the MMAs of tile i+1)
//   warps 0-7: epilogue, two warpgroups splitting the tile's column blocks -- tcgen05.ld, + bias, optional GELU,
//              conversion, 128-byte swizzled staging rows, TMA store.  (Keeping the weight slab resident for K <= 384 was
//              measured: 3-7 % on the stage-1 shapes, nothing on the pass; not kept.)  The
//              residual form (X += Y) is a TMA reduce-add store of the fp32 tile: the read-modify-write happens in L2 and
//              every element receives exactly one addend per GEMM, so the result does not depend on scheduling.
#include <cuda_bf16.h>

#include "clip.cuh"
#include "gelu.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace m2t {

#ifndef LG_STAGES_OVERRIDE
#define LG_STAGES_OVERRIDE 4
#endif
constexpr int LG_BN = 128, LG_STAGES = LG_STAGES_OVERRIDE;
constexpr int LG_NWG = 2;                        // epilogue warpgroups: column blocks blk % LG_NWG == wg
constexpr int LG_EPI_WARPS = 4 * LG_NWG, LG_THREADS = 32 * (LG_EPI_WARPS + 2);
constexpr uint32_t LG_A = 128 * 128, LG_B = LG_BN * 128, LG_STAGE = LG_A + LG_B;
constexpr uint32_t LG_OUT = LG_NWG * 2 * 16384;
constexpr uint32_t LG_OFF_OUT = LG_STAGES * LG_STAGE;
constexpr uint32_t LG_SMEM = 1024 + LG_STAGES * LG_STAGE + LG_OUT + 256;

__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                 :: "l"(m), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}

// exact-erf GELU (HF ACT2FN["gelu"]) in the erfc form of gelu.cuh, fp32 results, two elements at a time with the packed
// fp32x2 instructions (one issue slot per pair for the 6 FMA-pipe operations): the GELU epilogue is issue-bound
// (tools/lin_timing.py: +16 us on a 23 us kernel at N = 384, K = 96 with a scalar form)
__device__ __forceinline__ void gelu_erfc2(float& x0, float& x1) { f2_unpack(gelu_pair(f2_pack(x0, x1)), x0, x1); }

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&v);
}

#ifdef M2T_TIMING
// clock64 stamps of CTA 0 of the last LIN_BF16 launch, tiles 2..9 of that CTA (slot = 16 * (tile - 2) + k):
// k 0 producer: ring slot free for the tile's first K block; k 1..4 MMA warp: before / after the accumulator-free wait,
// first K block landed, last MMAs issued; k 5..10 epilogue thread 0: before / after the accumulator-ready wait, staging
// buffer free (the store two blocks back has read it), block staged, TMA store issued, accumulator released
__device__ long long g_lin_dbg[128];
#define M2T_LT(tile, k) do { if (EPI == LIN_BF16 && blockIdx.x == 0 && (tile) >= 2 && (tile) < 10) g_lin_dbg[16 * ((tile) - 2) + (k)] = clock64(); } while (0)
int read_lin_timing(long long* host128) {
    M2T_CUDA(cudaMemcpyFromSymbol(host128, g_lin_dbg, sizeof(long long) * 128));
    return M2T_OK;
}
#else
#define M2T_LT(tile, k) do { } while (0)
int read_lin_timing(long long* host128) { memset(host128, 0, sizeof(long long) * 128); return M2T_OK; }
#endif

template <int EPI>
__global__ void __launch_bounds__(LG_THREADS, 1)
lin_umma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW,
                const __grid_constant__ CUtensorMap mapO, const float* __restrict__ bias, int M, int N, int K) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + LG_OFF_OUT + LG_OUT);
    uint64_t* full = bars;                        // [STAGES]
    uint64_t* empty = bars + LG_STAGES;           // [STAGES]
    uint64_t* tfull = bars + 2 * LG_STAGES;       // [2] accumulator ready
    uint64_t* tempty = bars + 2 * LG_STAGES + 2;  // [2] accumulator drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * LG_STAGES + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int num_mt = (M + 127) / 128, num_nt = (N + LG_BN - 1) / LG_BN, num_t = num_mt * num_nt;
    const int KB = (K + 63) / 64;

    constexpr int W_TMA = LG_EPI_WARPS, W_MMA = LG_EPI_WARPS + 1;
    if (warp == W_MMA) tmem_alloc(tmem_slot, 256);
    if (tid == W_TMA * 32) {
        for (int s = 0; s < LG_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], LG_EPI_WARPS); }
        mbar_fence_init();
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapW);
        tma_prefetch_desc(&mapO);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();

    if (warp == W_TMA) {
        pdl_wait();
        uint32_t it = 0, tlp = 0;
        for (int t = blockIdx.x; t < num_t; t += gridDim.x, ++tlp) {
            const int mt = t / num_nt, nt = t - mt * num_nt;
            for (int kb = 0; kb < KB; ++kb, ++it) {
                const uint32_t s = it % LG_STAGES, ph = (it / LG_STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                if (kb == 0 && lane == 0) M2T_LT(tlp, 0);
                if (elect_one_sync()) {
                    mbar_expect_tx(&full[s], LG_STAGE);
                    tma_load_2d(sm + s * LG_STAGE, &mapA, &full[s], kb * 64, mt * 128);
                    tma_load_2d(sm + s * LG_STAGE + LG_A, &mapW, &full[s], kb * 64, nt * LG_BN);
                }
                __syncwarp();
            }
        }
    } else if (warp == W_MMA) {
        constexpr uint32_t idesc = umma_idesc_f16(128, LG_BN, 0, 0, 1);
        constexpr uint64_t tmpl = umma_smem_desc(0, 16, 1024, UMMA_LAYOUT_SW128);
        uint32_t it = 0, tl = 0;
        for (int t = blockIdx.x; t < num_t; t += gridDim.x, ++tl) {
            const uint32_t acc = tl & 1, aph = (tl >> 1) & 1;
            if (lane == 0) M2T_LT(tl, 1);
            mbar_wait(&tempty[acc], aph ^ 1);
            if (lane == 0) M2T_LT(tl, 2);
            tc_fence_after();
            for (int kb = 0; kb < KB; ++kb, ++it) {
                const uint32_t s = it % LG_STAGES, ph = (it / LG_STAGES) & 1;
                mbar_wait(&full[s], ph);
                if (kb == 0 && lane == 0) M2T_LT(tl, 3);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint64_t da0 = umma_desc_at(tmpl, base + s * LG_STAGE);
                    const uint64_t db0 = umma_desc_at(tmpl, base + s * LG_STAGE + LG_A);
                    const int rem = K - kb * 64, ks = rem >= 64 ? 4 : (rem + 15) / 16;
                    for (int k = 0; k < ks; ++k)
                        umma_f16_ss(tmem_base + acc * LG_BN, da0 + 2 * k, db0 + 2 * k, idesc, (kb | k) ? 1u : 0u);
                    umma_commit(&empty[s]);
                    if (kb == KB - 1) { umma_commit(&tfull[acc]); M2T_LT(tl, 4); }
                }
                __syncwarp();
            }
        }
    } else {
        pdl_wait();
        constexpr bool F32 = EPI == LIN_ADD_F32 || EPI == LIN_F32;
        constexpr int BLKC = F32 ? 32 : 64;              // columns per 128-byte staging row
        const int wg = warp >> 2, wq = warp & 3;         // TMEM lanes 32 wq .. 32 wq + 31 belong to warps with id % 4 == wq
        const int trow = wq * 32 + lane;
        const bool issuer = (tid & 127) == 0;
        const uint32_t bar_id = 1 + wg;
        // TMEM reads run at 64 B/clk per SM (B300_MICROARCH.md): a 128 x 128 fp32 accumulator takes >= 1024 cycles to read,
        // more than its MMAs at K <= 256, and tools/lin_stamps.py shows ~1800 cycles of epilogue per tile however the work is
        // arranged: the two warpgroups splitting each tile's column blocks (below) and the two taking whole tiles alternately
        // (one accumulator each, measured) come out the same.
        uint32_t tl = 0, nblk = 0;
        for (int t = blockIdx.x; t < num_t; t += gridDim.x, ++tl) {
            const int mt = t / num_nt, nt = t - mt * num_nt;
            const uint32_t acc = tl & 1, aph = (tl >> 1) & 1;
            if (tid == 0) M2T_LT(tl, 5);
            mbar_wait(&tfull[acc], aph);
            if (tid == 0) M2T_LT(tl, 6);
            tc_fence_after();
#pragma unroll 1
            for (int blk = wg; blk < LG_BN / BLKC; blk += LG_NWG) {
                const int col0 = nt * LG_BN + blk * BLKC;
                if (col0 >= N) break;                    // uniform within the warpgroup
                uint8_t* ob = sm + LG_OFF_OUT + (wg * 2 + (nblk & 1)) * 16384;
                if (issuer) tma_store_wait_read1();      // the store that last used this buffer has read it
                asm volatile("bar.sync %0, 128;" :: "r"(bar_id) : "memory");
                if (tid == 0) M2T_LT(tl, 7);
                // the epilogue's per-tile chain is the kernel's critical path (tools/lin_stamps.py): both TMEM reads of the
                // block are issued together and the bias loads, which do not depend on the accumulator, go out before the wait
                uint32_t racc[BLKC];
                float4 b4[BLKC / 4];
#pragma unroll
                for (int hh = 0; hh < BLKC / 32; ++hh)
                    tmem_ld32(tmem_base + acc * LG_BN + blk * BLKC + hh * 32 + ((uint32_t)(wq * 32) << 16), racc + 32 * hh);
#pragma unroll
                for (int q = 0; q < BLKC / 4; ++q) {     // N is a multiple of 32: a 32-column group is all in or all out
                    b4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (bias != nullptr && col0 + (q / 8) * 32 < N) b4[q] = __ldg(reinterpret_cast<const float4*>(bias + col0) + q);
                }
                tmem_ld_wait();
#pragma unroll
                for (int hh = 0; hh < BLKC / 32; ++hh) {
                    const uint32_t* r = racc + 32 * hh;
                    float v[32];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 bq = b4[hh * 8 + q];
                        v[4 * q + 0] = __uint_as_float(r[4 * q + 0]) + bq.x;
                        v[4 * q + 1] = __uint_as_float(r[4 * q + 1]) + bq.y;
                        v[4 * q + 2] = __uint_as_float(r[4 * q + 2]) + bq.z;
                        v[4 * q + 3] = __uint_as_float(r[4 * q + 3]) + bq.w;
                    }
                    if constexpr (EPI == LIN_GELU_BF16) {
#pragma unroll
                        for (int e = 0; e < 32; e += 2) gelu_erfc2(v[e], v[e + 1]);
                    }
                    if constexpr (F32) {
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            *reinterpret_cast<float4*>(ob + trow * 128 + ((q ^ (trow & 7)) << 4)) =
                                make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            uint4 u;
                            u.x = pack_bf16x2(v[8 * q + 0], v[8 * q + 1]);
                            u.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
                            u.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]);
                            u.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
                            *reinterpret_cast<uint4*>(ob + trow * 128 + (((hh * 4 + q) ^ (trow & 7)) << 4)) = u;
                        }
                    }
                }
                if (tid == 0) M2T_LT(tl, 8);
                fence_proxy_async();
                asm volatile("bar.sync %0, 128;" :: "r"(bar_id) : "memory");
                if (issuer) {
                    if constexpr (EPI == LIN_ADD_F32) tma_reduce_add_2d(&mapO, ob, col0, mt * 128);
                    else tma_store_2d(&mapO, ob, col0, mt * 128);
                    tma_store_commit();
                }
                if (tid == 0) M2T_LT(tl, 9);
                ++nblk;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (tid == 0) M2T_LT(tl, 10);
        }
        if (issuer) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) tmem_dealloc(tmem_base, 256);
}

template <int EPI>
static int launch_lin_epi(const void* A, const void* W, const float* bias, void* out, int M, int N, int K, cudaStream_t s) {
    CUtensorMap mapA, mapW, mapO;
    {
        const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M}, str[2] = {2, (uint64_t)K * 2};
        const uint32_t box[2] = {64, 128};
        M2T_TRY(make_tensor_map(&mapA, A, 2, 2, dims, str, box, 3));
    }
    {
        const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N}, str[2] = {2, (uint64_t)K * 2};
        const uint32_t box[2] = {64, (uint32_t)LG_BN};
        M2T_TRY(make_tensor_map(&mapW, W, 2, 2, dims, str, box, 3));
    }
    {
        constexpr bool F32 = EPI == LIN_ADD_F32 || EPI == LIN_F32;
        const int eb = F32 ? 4 : 2;
        const uint64_t dims[2] = {(uint64_t)N, (uint64_t)M}, str[2] = {(uint64_t)eb, (uint64_t)N * eb};
        const uint32_t box[2] = {(uint32_t)(F32 ? 32 : 64), 128};
        M2T_TRY(make_tensor_map(&mapO, out, eb, 2, dims, str, box, 3));
    }
    M2T_ENSURE_SMEM(lin_umma_kernel<EPI>, LG_SMEM);
    const int num_t = cdiv(M, 128) * cdiv(N, LG_BN);
    int grid = device_sm_count();
    if (grid > num_t) grid = num_t;
    M2T_CUDA(launch_pdl(lin_umma_kernel<EPI>, dim3(grid), dim3(LG_THREADS), LG_SMEM, s, mapA, mapW, mapO, bias, M, N, K));
    return M2T_OK;
}

// A bf16 [M][K], W bf16 [N][K], bias fp32 [N] or null; out: bf16 [M][N] (LIN_BF16, LIN_GELU_BF16) or fp32 [M][N]
// (LIN_F32 overwrites, LIN_ADD_F32 accumulates).  K a multiple of 8 (16-byte rows), N a multiple of 32.
int launch_lin_umma(int epi, const void* A, const void* W, const float* bias, void* out, int M, int N, int K,
                    cudaStream_t s) {
    if (M < 1 || N < 32 || N % 32 || K < 16 || K % 8) { set_error("lin_umma: bad shape M %d N %d K %d", M, N, K); return M2T_E_ARG; }
    switch (epi) {
        case LIN_BF16: return launch_lin_epi<LIN_BF16>(A, W, bias, out, M, N, K, s);
        case LIN_GELU_BF16: return launch_lin_epi<LIN_GELU_BF16>(A, W, bias, out, M, N, K, s);
        case LIN_ADD_F32: return launch_lin_epi<LIN_ADD_F32>(A, W, bias, out, M, N, K, s);
        case LIN_F32: return launch_lin_epi<LIN_F32>(A, W, bias, out, M, N, K, s);
    }
    set_error("lin_umma: unknown epilogue %d", epi);
    return M2T_E_ARG;
}

}  // namespace m2t
