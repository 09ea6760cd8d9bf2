// Fused MLP of the MedCLIP image tower's first two stages (Swin-T, C = 96 / 192; modeling_swin.py:490-533, :648-655):
//   X[m][:] += W2 . gelu(W1 . A[m][:] + b1) + b2        A bf16 [M][C] (LayerNorm output), W1 bf16 [4C][C], W2 bf16 [C][4C]
// The [tokens][4C] intermediate never leaves the SM: per 128-token tile the hidden dimension is walked in chunks of 128;
// GEMM-1 of chunk j (M128 x N128 x K=C, accumulator in TMEM) is converted by the epilogue warps (+ b1, GELU, bf16) into a
// 128-byte-swizzled shared-memory tile that is the A operand of GEMM-2 (M128 x N=C x K128, accumulating over the chunks in a
// second TMEM accumulator).  The separate fc1 / fc2 launches write and re-read 1536 B per token for that tensor, a quarter
// of the layer's traffic, and these stages are bound by HBM.
// Warp roles (320 threads, one CTA per SM, persistent over row tiles):
//   warps 0-7: epilogue warpgroups 0 / 1 (columns 0-63 / 64-127 of a chunk; alternating 32-column blocks of the output)
//   warp 8   : TMA producer: the tile's A blocks (resident for the tile), then W1 / W2 blocks through one ring in the
//              order the MMA warp consumes them: W1(0) | W1(1) W2(0) | W1(2) W2(1) | ... | W2(last)
//   warp 9   : tcgen05.mma issue: GEMM-1(j) then GEMM-2(j-1), so the conversion of chunk j-1 overlaps GEMM-1 of chunk j
#include <cuda_bf16.h>

#include "clip.cuh"
#include "gelu.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace m2t {

namespace {

template <int C>
struct MlpCfg {
    static constexpr int KB1 = (C + 63) / 64;                 // K blocks of GEMM-1 (C = 96: the second one is half zero-filled)
    static constexpr int NCH = 4 * C / 128;                   // hidden chunks
    static constexpr uint32_t ABLK = 128 * 128;               // one 128-row x 64-column bf16 block
    static constexpr uint32_t W2BLK = C * 128;                // W2 block: C rows x 64 hidden columns
    static constexpr uint32_t SLOT = W2BLK > ABLK ? W2BLK : ABLK;
    static constexpr int SLOTS = C == 96 ? 6 : 4;
    static constexpr uint32_t OFF_G = KB1 * ABLK;             // two G buffers of two blocks
    static constexpr uint32_t OFF_RING = OFF_G + 4 * ABLK;
    static constexpr uint32_t OFF_BAR = OFF_RING + SLOTS * SLOT;
    static constexpr uint32_t SMEM = 1024 + OFF_BAR + 512;
    static constexpr uint32_t ACC2 = 256;                     // TMEM column of the output accumulator (acc1: 0 and 128)
};

__device__ __forceinline__ void gelu2(float& x0, float& x1) { f2_unpack(gelu_pair(f2_pack(x0, x1)), x0, x1); }

__device__ __forceinline__ uint32_t bf16x2(float a, float b) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&v);
}

__device__ __forceinline__ void reduce_add_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                 :: "l"(m), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}

template <int C>
__global__ void __launch_bounds__(320, 1)
mlp_umma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW1,
                const __grid_constant__ CUtensorMap mapW2, const __grid_constant__ CUtensorMap mapX,
                const float* __restrict__ b1, const float* __restrict__ b2, int M) {
    using CF = MlpCfg<C>;
    constexpr int KB1 = CF::KB1, NCH = CF::NCH, S = CF::SLOTS;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + CF::OFF_BAR);
    uint64_t* rfull = bars;                 // [S] ring slot landed
    uint64_t* rempty = bars + S;            // [S] ring slot consumed
    uint64_t* a_full = bars + 2 * S;        // A blocks of the tile landed
    uint64_t* a_empty = a_full + 1;         // GEMM-1s of the tile done with A
    uint64_t* acc1_full = a_full + 2;       // [2]
    uint64_t* acc1_empty = a_full + 4;      // [2]
    uint64_t* g_full = a_full + 6;          // [2] G buffer written (8 warps)
    uint64_t* g_empty = a_full + 8;         // [2] GEMM-2 done with the G buffer
    uint64_t* acc2_full = a_full + 10;
    uint64_t* acc2_empty = a_full + 11;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 12);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int num_mt = (M + 127) / 128;

    if (warp == 9) tmem_alloc(tmem_slot, 512);
    if (tid == 256) {
        for (int s = 0; s < S; ++s) { mbar_init(&rfull[s], 1); mbar_init(&rempty[s], 1); }
        mbar_init(a_full, 1); mbar_init(a_empty, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc1_full[i], 1); mbar_init(&acc1_empty[i], 8);
            mbar_init(&g_full[i], 8); mbar_init(&g_empty[i], 1);
        }
        mbar_init(acc2_full, 1); mbar_init(acc2_empty, 8);
        mbar_fence_init();
        tma_prefetch_desc(&mapA); tma_prefetch_desc(&mapW1); tma_prefetch_desc(&mapW2); tma_prefetch_desc(&mapX);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();

    if (warp == 8) {
        // ---- TMA producer ----
        pdl_wait();
        uint32_t it = 0, tlc = 0;
        auto ring_load = [&](const CUtensorMap* map, int c0, int c1, uint32_t bytes) {
            const uint32_t s = it % S, ph = (it / S) & 1;
            mbar_wait(&rempty[s], ph ^ 1);
            if (elect_one_sync()) {
                mbar_expect_tx(&rfull[s], bytes);
                tma_load_2d(sm + CF::OFF_RING + s * CF::SLOT, map, &rfull[s], c0, c1);
            }
            __syncwarp();
            ++it;
        };
        for (int mt = blockIdx.x; mt < num_mt; mt += gridDim.x, ++tlc) {
            mbar_wait(a_empty, (tlc & 1) ^ 1);
            if (elect_one_sync()) {
                mbar_expect_tx(a_full, KB1 * CF::ABLK);
                for (int kb = 0; kb < KB1; ++kb) tma_load_2d(sm + kb * CF::ABLK, &mapA, a_full, kb * 64, mt * 128);
            }
            __syncwarp();
            for (int j = 0; j <= NCH; ++j) {
                if (j < NCH)
                    for (int kb = 0; kb < KB1; ++kb) ring_load(&mapW1, kb * 64, j * 128, CF::ABLK);
                if (j >= 1)
                    for (int kb2 = 0; kb2 < 2; ++kb2) ring_load(&mapW2, (j - 1) * 128 + kb2 * 64, 0, CF::W2BLK);
            }
        }
    } else if (warp == 9) {
        // ---- MMA issue ----
        constexpr uint32_t idesc1 = umma_idesc_f16(128, 128, 0, 0, 1), idesc2 = umma_idesc_f16(128, C, 0, 0, 1);
        constexpr uint64_t tmpl = umma_smem_desc(0, 16, 1024, UMMA_LAYOUT_SW128);
        uint32_t it = 0, tlc = 0, cc = 0;
        auto gemm2 = [&](uint32_t c, bool first) {
            const uint32_t pb = c & 1, pu = c >> 1;
            if (first) { mbar_wait(acc2_empty, (tlc & 1) ^ 1); }
            mbar_wait(&g_full[pb], pu & 1);
            tc_fence_after();
            for (int kb2 = 0; kb2 < 2; ++kb2, ++it) {
                const uint32_t s = it % S, ph = (it / S) & 1;
                mbar_wait(&rfull[s], ph);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint64_t da0 = umma_desc_at(tmpl, base + CF::OFF_G + (pb * 2 + kb2) * CF::ABLK);
                    const uint64_t db0 = umma_desc_at(tmpl, base + CF::OFF_RING + s * CF::SLOT);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_f16_ss(tmem_base + CF::ACC2, da0 + 2 * k, db0 + 2 * k, idesc2, (first && kb2 == 0 && k == 0) ? 0u : 1u);
                    umma_commit(&rempty[s]);
                    if (kb2 == 1) umma_commit(&g_empty[pb]);
                }
                __syncwarp();
            }
        };
        for (int mt = blockIdx.x; mt < num_mt; mt += gridDim.x, ++tlc) {
            mbar_wait(a_full, tlc & 1);
            tc_fence_after();
            for (int j = 0; j < NCH; ++j, ++cc) {
                const uint32_t buf = cc & 1, u = cc >> 1;
                mbar_wait(&acc1_empty[buf], (u & 1) ^ 1);
                tc_fence_after();
                for (int kb = 0; kb < KB1; ++kb, ++it) {
                    const uint32_t s = it % S, ph = (it / S) & 1;
                    mbar_wait(&rfull[s], ph);
                    tc_fence_after();
                    if (elect_one_sync()) {
                        const uint64_t da0 = umma_desc_at(tmpl, base + kb * CF::ABLK);
                        const uint64_t db0 = umma_desc_at(tmpl, base + CF::OFF_RING + s * CF::SLOT);
                        const int rem = C - kb * 64, ks = rem >= 64 ? 4 : (rem + 15) / 16;
                        for (int k = 0; k < ks; ++k)
                            umma_f16_ss(tmem_base + buf * 128, da0 + 2 * k, db0 + 2 * k, idesc1, (kb | k) ? 1u : 0u);
                        umma_commit(&rempty[s]);
                        if (kb == KB1 - 1) {
                            umma_commit(&acc1_full[buf]);
                            if (j == NCH - 1) umma_commit(a_empty);        // the tile's last GEMM-1: A may be overwritten
                        }
                    }
                    __syncwarp();
                }
                if (j >= 1) gemm2(cc - 1, j == 1);
            }
            gemm2(cc - 1, NCH == 1);
            if (elect_one_sync()) umma_commit(acc2_full);
            __syncwarp();
        }
    } else {
        // ---- epilogue warpgroups ----
        pdl_wait();
        const int wg = warp >> 2, wq = warp & 3, trow = wq * 32 + lane;
        const bool issuer = (tid & 127) == 0;
        const uint32_t bar_id = 1 + wg, lane_sel = (uint32_t)(wq * 32) << 16;
        uint32_t cc = 0, tlc = 0, nst = 0;
        for (int mt = blockIdx.x; mt < num_mt; mt += gridDim.x, ++tlc) {
            for (int j = 0; j < NCH; ++j, ++cc) {
                const uint32_t buf = cc & 1, u = cc >> 1;
                // bias of this warpgroup's 64 hidden columns: independent of the accumulator, loaded before the waits
                float4 bb[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) bb[q] = __ldg(reinterpret_cast<const float4*>(b1 + j * 128 + wg * 64) + q);
                mbar_wait(&acc1_full[buf], u & 1);
                tc_fence_after();
                uint32_t r[64];
                tmem_ld32(tmem_base + buf * 128 + wg * 64 + lane_sel, r);
                tmem_ld32(tmem_base + buf * 128 + wg * 64 + 32 + lane_sel, r + 32);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc1_empty[buf]);
                mbar_wait(&g_empty[buf], (u & 1) ^ 1);        // the GEMM-2 that read this buffer two chunks ago is done
                uint8_t* gb = sm + CF::OFF_G + (buf * 2 + wg) * CF::ABLK + trow * 128;
#pragma unroll
                for (int q = 0; q < 8; ++q) {                 // 8 columns -> one 16-byte chunk
                    float v[8];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float4 b4 = bb[2 * q + e];
                        v[4 * e + 0] = __uint_as_float(r[8 * q + 4 * e + 0]) + b4.x;
                        v[4 * e + 1] = __uint_as_float(r[8 * q + 4 * e + 1]) + b4.y;
                        v[4 * e + 2] = __uint_as_float(r[8 * q + 4 * e + 2]) + b4.z;
                        v[4 * e + 3] = __uint_as_float(r[8 * q + 4 * e + 3]) + b4.w;
                    }
#pragma unroll
                    for (int e = 0; e < 8; e += 2) gelu2(v[e], v[e + 1]);
                    uint4 o;
                    o.x = bf16x2(v[0], v[1]); o.y = bf16x2(v[2], v[3]); o.z = bf16x2(v[4], v[5]); o.w = bf16x2(v[6], v[7]);
                    *reinterpret_cast<uint4*>(gb + ((q ^ (trow & 7)) << 4)) = o;
                }
                fence_proxy_async();                           // generic-proxy writes -> visible to the MMA's async-proxy reads
                __syncwarp();
                if (lane == 0) mbar_arrive(&g_full[buf]);
            }
            // output: acc2 [128][C] + b2 -> fp32 reduce-add into X, 32-column blocks alternating between the warpgroups;
            // staging in this warpgroup's blocks of the G buffers (every GEMM-2 of the tile has completed: acc2_full)
            mbar_wait(acc2_full, tlc & 1);
            tc_fence_after();
#pragma unroll 1
            for (int blk = wg; blk < C / 32; blk += 2, ++nst) {
                uint8_t* ob = sm + CF::OFF_G + ((nst & 1) * 2 + wg) * CF::ABLK;
                if (issuer) tma_store_wait_read1();
                asm volatile("bar.sync %0, 128;" :: "r"(bar_id) : "memory");
                uint32_t r[32];
                tmem_ld32(tmem_base + CF::ACC2 + blk * 32 + lane_sel, r);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(b2 + blk * 32) + q);
                    *reinterpret_cast<float4*>(ob + trow * 128 + ((q ^ (trow & 7)) << 4)) =
                        make_float4(__uint_as_float(r[4 * q]) + b4.x, __uint_as_float(r[4 * q + 1]) + b4.y,
                                    __uint_as_float(r[4 * q + 2]) + b4.z, __uint_as_float(r[4 * q + 3]) + b4.w);
                }
                fence_proxy_async();
                asm volatile("bar.sync %0, 128;" :: "r"(bar_id) : "memory");
                if (issuer) { reduce_add_2d(&mapX, ob, blk * 32, mt * 128); tma_store_commit(); }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc2_empty);
            // the next tile's first chunks overwrite the staging blocks: their stores must have read them
            if (issuer) tma_store_wait_read();
            asm volatile("bar.sync %0, 128;" :: "r"(bar_id) : "memory");
        }
        if (issuer) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem_base, 512);
}

template <int C>
int launch_mlp_c(const void* A, const void* W1, const float* b1, const void* W2, const float* b2, float* X, int M, cudaStream_t s) {
    using CF = MlpCfg<C>;
    CUtensorMap mapA, mapW1, mapW2, mapX;
    {
        const uint64_t dims[2] = {(uint64_t)C, (uint64_t)M}, str[2] = {2, (uint64_t)C * 2};
        const uint32_t box[2] = {64, 128};
        M2T_TRY(make_tensor_map(&mapA, A, 2, 2, dims, str, box, 3));
    }
    {
        const uint64_t dims[2] = {(uint64_t)C, (uint64_t)4 * C}, str[2] = {2, (uint64_t)C * 2};
        const uint32_t box[2] = {64, 128};
        M2T_TRY(make_tensor_map(&mapW1, W1, 2, 2, dims, str, box, 3));
    }
    {
        const uint64_t dims[2] = {(uint64_t)4 * C, (uint64_t)C}, str[2] = {2, (uint64_t)4 * C * 2};
        const uint32_t box[2] = {64, (uint32_t)C};
        M2T_TRY(make_tensor_map(&mapW2, W2, 2, 2, dims, str, box, 3));
    }
    {
        const uint64_t dims[2] = {(uint64_t)C, (uint64_t)M}, str[2] = {4, (uint64_t)C * 4};
        const uint32_t box[2] = {32, 128};
        M2T_TRY(make_tensor_map(&mapX, X, 4, 2, dims, str, box, 3));
    }
    M2T_ENSURE_SMEM(mlp_umma_kernel<C>, CF::SMEM);
    const int num_mt = cdiv(M, 128);
    int grid = device_sm_count();
    if (grid > num_mt) grid = num_mt;
    M2T_CUDA(launch_pdl(mlp_umma_kernel<C>, dim3(grid), dim3(320), CF::SMEM, s, mapA, mapW1, mapW2, mapX, b1, b2, M));
    return M2T_OK;
}

}  // namespace

// X fp32 [M][C] += W2 . gelu(W1 . A + b1) + b2;  A bf16 [M][C], W1 bf16 [4C][C], W2 bf16 [C][4C];  C = 96 or 192
int launch_mlp_umma(const void* A, const void* W1, const float* b1, const void* W2, const float* b2, float* X, int M, int C,
                    cudaStream_t s) {
    if (M < 1) { set_error("mlp_umma: bad row count %d", M); return M2T_E_ARG; }
    if (C == 96) return launch_mlp_c<96>(A, W1, b1, W2, b2, X, M, s);
    if (C == 192) return launch_mlp_c<192>(A, W1, b1, W2, b2, X, M, s);
    set_error("mlp_umma: built for C = 96 and 192, got %d", C);
    return M2T_E_UNSUPPORTED;
}

}  // namespace m2t
