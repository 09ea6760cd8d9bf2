// tcgen05 GEMM for the TBlock qkv 1x1 conv (ref M2Trans_network.py:307) at C = 16, 64 and 256:
//   QKV[m][n] = sum_k Z[m][k] * Wqkv[n][k]      Z fp16 [M][C], Wqkv fp16 [3C][C], QKV fp16 [M][3C]
// Persistent, warp-specialised CTAs (192 threads):
//   warp 4  : TMA producer -- the CTA's weight slab (NT x C, resident for the CTA's lifetime), then a ring of
//             128-row activation tiles in 64-channel K blocks (128-byte swizzle; 32-byte swizzle for C = 16)
//   warp 5  : single-thread tcgen05.mma issue, M=128 x N=NT x K=16 per instruction, fp32 accumulators in TMEM
//             (two accumulators of 256 columns: the epilogue of tile i overlaps the MMAs of tile i+1)
//   warps 0-3: epilogue -- tcgen05.ld, fp32 -> fp16, 16-byte stores
// C = 256: the 768 output channels are split into q / k / v slabs of NT = 256; CTA c owns slab c % 3, so the
// 128 KB slab is read from L2 once per CTA and every 64 KB activation tile feeds 16 x 128 cycles of MMA
// (32 B/clk/SM, under the ~42 B/clk/SM L2->SM ceiling of B300_MICROARCH.md).  C = 64 / 16: one slab (192 / 48).
#include "common.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace m2t {

template <int C>
struct QkvCfg {
    static constexpr int CB = C < 64 ? C : 64;               // channels per K block
    static constexpr int KB = C / CB;
    static constexpr int KSTEPS = CB / 16;
    static constexpr uint32_t ROWB = CB * 2;                  // bytes per row: 32 or 128
    static constexpr uint64_t LAYOUT = CB == 64 ? UMMA_LAYOUT_SW128 : UMMA_LAYOUT_SW32;
    static constexpr int TMA_SWZ = CB == 64 ? 3 : 1;
    static constexpr uint32_t SBO = 8 * ROWB;
    static constexpr int NT = C == 256 ? 256 : 3 * C;
    static constexpr int NCHUNK = 3 * C / NT;
    static constexpr int OCH = NT % 32 == 0 ? 32 : 16;       // epilogue column chunk
    static constexpr int STAGES = 4;
    static constexpr uint32_t A_STAGE = 128 * ROWB;
    static constexpr uint32_t B_BLOCK = NT * ROWB;
    static constexpr uint32_t B_BYTES = (KB * B_BLOCK + 1023) / 1024 * 1024;
    // C >= 64: the epilogue stages the fp16 tile in two 128-row x 64-column (128-byte-swizzled) buffers and sends it out
    // with TMA stores; per-thread row stores made the kernel epilogue-bound (4 K cycles per tile against 1.4 K of MMA)
    static constexpr bool TSTORE = C >= 64;
    static constexpr uint32_t OUT_BYTES = TSTORE ? 2 * 128 * 128 : 0;
    static constexpr uint32_t OFF_OUT = B_BYTES + STAGES * A_STAGE;
    static constexpr uint32_t SMEM_MIN = 1024 + B_BYTES + STAGES * A_STAGE + OUT_BYTES + 256;
    // every CTA allocates all 512 TMEM columns: never let two share an SM
    static constexpr uint32_t SMEM = SMEM_MIN < 120 * 1024 ? 120 * 1024 : SMEM_MIN;
};

#ifdef M2T_TIMING
__device__ long long g_qkv_dbg[64];   // C = 256, CTA 0: [0] entry, [1] prologue done, [2] weight slab landed; per tile t < 6 at
                                      // [8+8t..]: MMA warp before waits, first A block landed, MMAs issued; epilogue: acc ready, done
#define M2T_QT(slot) do { if (C == 256 && blockIdx.x == 0 && lane == 0) g_qkv_dbg[(slot)] = clock64(); } while (0)
#else
#define M2T_QT(slot) do { } while (0)
#endif

template <int C>
__global__ void __launch_bounds__(192, 1)
qkv_umma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW,
                const __grid_constant__ CUtensorMap mapO, __half* __restrict__ out, int M) {
    using CF = QkvCfg<C>;
    constexpr int NT = CF::NT, KB = CF::KB, STAGES = CF::STAGES, CB = CF::CB;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    uint8_t* sB = sm;
    uint8_t* sA = sm + CF::B_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + CF::B_BYTES + STAGES * CF::A_STAGE + CF::OUT_BYTES);
    uint64_t* full = bars;                 // [STAGES]
    uint64_t* empty = bars + STAGES;       // [STAGES]
    uint64_t* bfull = bars + 2 * STAGES;   // weights landed
    uint64_t* tfull = bars + 2 * STAGES + 1;   // [2] accumulator ready
    uint64_t* tempty = bars + 2 * STAGES + 3;  // [2] accumulator drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 5);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int chunk = blockIdx.x % CF::NCHUNK;
    const int first = blockIdx.x / CF::NCHUNK, stride = gridDim.x / CF::NCHUNK;
    const int num_mt = (M + 127) / 128;
    if (warp == 0) M2T_QT(0);

    if (warp == 5) tmem_alloc(tmem_slot, 512);
    if (tid == 128) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(bfull, 1);
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
        mbar_fence_init();
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapW);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    if (warp == 0) M2T_QT(1);

    if (warp == 4) {
        // TMA producer: warp-uniform loop, one elected lane issues
        if (elect_one_sync()) {      // the weight slab is constant: loaded while the previous kernel drains
            mbar_expect_tx(bfull, KB * CF::B_BLOCK);
            for (int kb = 0; kb < KB; ++kb) tma_load_2d(sB + kb * CF::B_BLOCK, &mapW, bfull, kb * CB, chunk * NT);
        }
        pdl_wait();
        uint32_t it = 0;
        for (int mt = first; mt < num_mt; mt += stride) {
            for (int kb = 0; kb < KB; ++kb, ++it) {
                const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                if (elect_one_sync()) {
                    mbar_expect_tx(&full[s], CF::A_STAGE);
                    tma_load_2d(sA + s * CF::A_STAGE, &mapA, &full[s], kb * CB, mt * 128);
                }
                __syncwarp();
            }
        }
    } else if (warp == 5) {
        // MMA issuer: warp-uniform loop, one elected lane issues
        constexpr uint32_t idesc = umma_idesc_f16(128, NT);
        constexpr uint64_t tmpl = umma_smem_desc(0, 16, CF::SBO, CF::LAYOUT);
        mbar_wait(bfull, 0);
        M2T_QT(2);
        uint32_t it = 0, t = 0;
        for (int mt = first; mt < num_mt; mt += stride, ++t) {
            const uint32_t acc = t & 1, aph = (t >> 1) & 1;
            if (t < 6) M2T_QT(8 + 8 * t);
            mbar_wait(&tempty[acc], aph ^ 1);
            tc_fence_after();
            for (int kb = 0; kb < KB; ++kb, ++it) {
                const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(&full[s], ph);
                if (kb == 0 && t < 6) M2T_QT(8 + 8 * t + 1);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint64_t da0 = umma_desc_at(tmpl, base + CF::B_BYTES + s * CF::A_STAGE);
                    const uint64_t db0 = umma_desc_at(tmpl, base + kb * CF::B_BLOCK);
#pragma unroll
                    for (int k = 0; k < CF::KSTEPS; ++k)
                        umma_f16_ss(tmem_base + acc * 256, da0 + 2 * k, db0 + 2 * k, idesc, (kb | k) ? 1u : 0u);
                    umma_commit(&empty[s]);
                    if (kb == KB - 1) umma_commit(&tfull[acc]);
                }
                __syncwarp();
            }
            if (t < 6) M2T_QT(8 + 8 * t + 2);
        }
    } else {
        pdl_wait();
        uint32_t t = 0, nblk = 0;
        (void)nblk;
        for (int mt = first; mt < num_mt; mt += stride, ++t) {
            const uint32_t acc = t & 1, aph = (t >> 1) & 1;
            mbar_wait(&tfull[acc], aph);
            tc_fence_after();
            if (warp == 0 && t < 6) M2T_QT(8 + 8 * t + 3);
            const int row = mt * 128 + warp * 32 + lane;
            if constexpr (CF::TSTORE) {
                // 64-column blocks through two staging buffers: TMEM -> fp16 -> swizzled smem row -> TMA store
                const int trow = warp * 32 + lane;
#pragma unroll 1
                for (int blk = 0; blk < NT / 64; ++blk, ++nblk) {
                    uint8_t* ob = sm + CF::OFF_OUT + (nblk & 1) * 16384;
                    if (tid == 0) tma_store_wait_read1();          // the store that last used this buffer has read it
                    asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        uint32_t r[32];
                        tmem_ld32(tmem_base + acc * 256 + blk * 64 + hh * 32 + ((uint32_t)(warp * 32) << 16), r);
                        tmem_ld_wait();
#pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            uint4 u;
                            uint32_t* pu = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const __half2 h = __floats2half2_rn(__uint_as_float(r[v * 8 + 2 * e]), __uint_as_float(r[v * 8 + 2 * e + 1]));
                                pu[e] = *reinterpret_cast<const uint32_t*>(&h);
                            }
                            *reinterpret_cast<uint4*>(ob + trow * 128 + (((hh * 4 + v) ^ (trow & 7)) << 4)) = u;
                        }
                    }
                    fence_proxy_async();
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (tid == 0) {
                        tma_store_2d(&mapO, ob, chunk * NT + blk * 64, mt * 128);     // rows past M are clipped
                        tma_store_commit();
                    }
                }
            } else {
            __half* orow = out + (long)row * (3 * C) + chunk * NT;
            constexpr int OCH = CF::OCH;
#pragma unroll 1
            for (int c0 = 0; c0 < NT; c0 += OCH) {
                uint32_t r[OCH];
                if constexpr (OCH == 32) tmem_ld32(tmem_base + acc * 256 + c0 + ((uint32_t)(warp * 32) << 16), r);
                else tmem_ld16(tmem_base + acc * 256 + c0 + ((uint32_t)(warp * 32) << 16), r);
                tmem_ld_wait();
                if (row < M) {
#pragma unroll
                    for (int v = 0; v < OCH / 16; ++v) {       // 32 B per store: one sector transaction, not two
                        uint4 u[2];
                        uint32_t* pu = reinterpret_cast<uint32_t*>(u);
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const __half2 h = __floats2half2_rn(__uint_as_float(r[v * 16 + 2 * e]),
                                                                __uint_as_float(r[v * 16 + 2 * e + 1]));
                            pu[e] = *reinterpret_cast<const uint32_t*>(&h);
                        }
                        stg256(orow + c0 + v * 16, u[0], u[1]);
                    }
                }
            }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (warp == 0 && t < 6) M2T_QT(8 + 8 * t + 4);
        }
        if (CF::TSTORE && tid == 0) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, 512);
}

#ifdef M2T_TIMING
int read_qkv_timing(long long* host64) {
    M2T_CUDA(cudaMemcpyFromSymbol(host64, g_qkv_dbg, sizeof(long long) * 64));
    return M2T_OK;
}
#else
int read_qkv_timing(long long* host64) { memset(host64, 0, sizeof(long long) * 64); return M2T_OK; }
#endif

template <int C>
static int launch_qkv_umma_c(const __half* Z, const __half* Wqkv, __half* QKV, int M, cudaStream_t s) {
    using CF = QkvCfg<C>;
    CUtensorMap mapA, mapW;
    {
        const uint64_t dims[2] = {(uint64_t)C, (uint64_t)M}, str[2] = {2, (uint64_t)C * 2};
        const uint32_t box[2] = {(uint32_t)CF::CB, 128};
        M2T_TRY(make_tensor_map(&mapA, Z, 2, 2, dims, str, box, CF::TMA_SWZ));
    }
    {
        const uint64_t dims[2] = {(uint64_t)C, (uint64_t)3 * C}, str[2] = {2, (uint64_t)C * 2};
        const uint32_t box[2] = {(uint32_t)CF::CB, (uint32_t)CF::NT};
        M2T_TRY(make_tensor_map(&mapW, Wqkv, 2, 2, dims, str, box, CF::TMA_SWZ));
    }
    CUtensorMap mapO = mapA;
    if (CF::TSTORE) {
        const uint64_t dims[2] = {(uint64_t)3 * C, (uint64_t)M}, str[2] = {2, (uint64_t)3 * C * 2};
        const uint32_t box[2] = {64, 128};
        M2T_TRY(make_tensor_map(&mapO, QKV, 2, 2, dims, str, box, 3));
    }
    M2T_ENSURE_SMEM(qkv_umma_kernel<C>, CF::SMEM);
    const int num_mt = (M + 127) / 128;
    int per_chunk = device_sm_count() / CF::NCHUNK;
    if (per_chunk > num_mt) per_chunk = num_mt;
    if (per_chunk < 1) per_chunk = 1;
    M2T_CUDA(launch_pdl(qkv_umma_kernel<C>, dim3(per_chunk * CF::NCHUNK), dim3(192), CF::SMEM, s, mapA, mapW, mapO, QKV, M));
    return M2T_OK;
}

int launch_qkv_umma(const __half* Z, const __half* Wqkv, __half* QKV, int M, int C, cudaStream_t s) {
    if (C == 16) return launch_qkv_umma_c<16>(Z, Wqkv, QKV, M, s);
    if (C == 64) return launch_qkv_umma_c<64>(Z, Wqkv, QKV, M, s);
    if (C == 256) return launch_qkv_umma_c<256>(Z, Wqkv, QKV, M, s);
    set_error("qkv_umma: unsupported channel count %d", C);
    return M2T_E_UNSUPPORTED;
}

}  // namespace m2t
