// tcgen05 tail (ref M2Trans_network.py:40-56, :72-76).
//   tail_up_umma<R> : 1x1 conv (GEMM M = 128 pixels, N = 64 R^2, K = 64) + bias + PixelShuffle(R) + GELU in the
//                     epilogue, fp16 NHWC output at R x resolution.  The weight rows are packed sub-pixel-major
//                     (row uv*64 + c  <-  reference row c*R*R + uv), so the 64 accumulator columns of one
//                     sub-pixel are the 64 contiguous channels of one output pixel: PixelShuffle is a store address.
//   tail_out_umma   : last 3x3 reflect conv 64 -> 3 (N padded to 16) as the same TMA-halo-tile implicit GEMM as the
//                     ff conv, reading a tensor that carries a 1-pixel reflected border (so TMA never leaves it),
//                     + clamp + crop, written straight into the caller's NCHW fp32 output.
// GELU is the exact erf form of nn.GELU(), evaluated two at a time on the packed fp32x2 pipe (gelu.cuh).
// A variant of tail_up that staged the output in shared memory and sent it out with TMA stores (one per 32-pixel segment
// and output row) was measured SLOWER (cfg2 +20 us, cfg3 +3 %): the per-unit barrier and the wait for the store to read
// the staging tile cost more than the per-thread 32-byte stores they replaced.
// Round 2: a per-warp staging tile with only __syncwarp and 16-byte stores that cover four complete 128-byte lines per
// instruction (instead of 32 lanes x 32 B in 32 lines) was slower as well (51 vs 45 us at cfg2, 844 vs 704 us at cfg3): the
// epilogue is bound by instruction issue, not by the LSU's sector rate.
#include "common.cuh"
#include "gelu.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace m2t {

template <int R>
struct TuCfg {
    static constexpr int N = NF * R * R;
    static constexpr int NT = R == 2 ? 256 : 192;
    static constexpr int NCH = N / NT;
    static constexpr int SUB = NT / NF;                       // sub-pixels per N chunk
    static constexpr int STAGES = 4;
    static constexpr uint32_t A_STAGE = 128 * 128;
    static constexpr uint32_t W_BYTES = N * 128;
    static constexpr uint32_t OFF_A = W_BYTES;
    static constexpr uint32_t OFF_BIAS = OFF_A + STAGES * A_STAGE;
    static constexpr uint32_t OFF_BAR = OFF_BIAS + N * 4;
    static constexpr uint32_t SMEM = 1024 + OFF_BAR + 256;
};

constexpr int TU_THREADS = 320;   // warps 0-3 + 6-9 epilogue, warp 4 TMA, warp 5 MMA

template <int R>
__global__ void __launch_bounds__(TU_THREADS, 1)
tail_up_umma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW,
                    const float* __restrict__ bias, __half* __restrict__ out, int M, int h, int w, int pad) {
    using CF = TuCfg<R>;
    constexpr int NT = CF::NT, NCH = CF::NCH, SUB = CF::SUB, STAGES = CF::STAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    float* sbias = reinterpret_cast<float*>(sm + CF::OFF_BIAS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + CF::OFF_BAR);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* wfull = bars + 2 * STAGES;
    uint64_t* tfull = bars + 2 * STAGES + 1;     // [2]
    uint64_t* tempty = bars + 2 * STAGES + 3;    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 5);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int num_mt = M / 128;
    for (int i = tid; i < CF::N; i += TU_THREADS) sbias[i] = bias[i];
    if (warp == 5) tmem_alloc(tmem_slot, 512);
    if (tid == 128) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(wfull, 1);
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 8); }
        mbar_fence_init();
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapW);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();

    if (warp == 4) {
        // TMA producer: warp-uniform loop, one elected lane issues
        if (elect_one_sync()) {
            mbar_expect_tx(wfull, CF::W_BYTES);
            for (int ch = 0; ch < NCH; ++ch) tma_load_2d(sm + ch * NT * 128, &mapW, wfull, 0, ch * NT);
        }
        pdl_wait();
        uint32_t it = 0;
        for (int mt = blockIdx.x; mt < num_mt; mt += gridDim.x, ++it) {
            const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            if (elect_one_sync()) {
                mbar_expect_tx(&full[s], CF::A_STAGE);
                tma_load_2d(sm + CF::OFF_A + s * CF::A_STAGE, &mapA, &full[s], 0, mt * 128);
            }
            __syncwarp();
        }
    } else if (warp == 5) {
        // MMA issuer: warp-uniform loop, one elected lane issues
        constexpr uint32_t idesc = umma_idesc_f16(128, NT);
        constexpr uint64_t tmpl = umma_smem_desc(0, 16, 1024, UMMA_LAYOUT_SW128);
        mbar_wait(wfull, 0);
        uint32_t it = 0, u = 0;
        for (int mt = blockIdx.x; mt < num_mt; mt += gridDim.x, ++it) {
            const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
            mbar_wait(&full[s], ph);
            tc_fence_after();
            const uint64_t da0 = umma_desc_at(tmpl, base + CF::OFF_A + s * CF::A_STAGE);
            for (int ch = 0; ch < NCH; ++ch, ++u) {
                const uint32_t acc = u & 1, aph = (u >> 1) & 1;
                mbar_wait(&tempty[acc], aph ^ 1);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint64_t db0 = umma_desc_at(tmpl, base + ch * NT * 128);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base + acc * 256, da0 + 2 * k, db0 + 2 * k, idesc, k ? 1u : 0u);
                    umma_commit(&tfull[acc]);
                    if (ch == NCH - 1) umma_commit(&empty[s]);
                }
                __syncwarp();
            }
        }
    } else {
        // Two epilogue warpgroups (warps 0-3 and 6-9) share every accumulator: the GELU epilogue is issue-bound,
        // so each TMEM lane quadrant (warp % 4) is served by two warps that split the sub-pixels between them.
        const int hr = h * R, wr = w * R;
        const long orow_pitch = (long)(wr + 2 * pad) * NF;
        const int quad = warp & 3, wg = warp >= 6 ? 1 : 0;
        constexpr int SPLIT = (SUB + 1) / 2;
        const int sp_lo = wg == 0 ? 0 : SPLIT, sp_hi = wg == 0 ? SPLIT : SUB;
        pdl_wait();
        uint32_t u = 0;
        for (int mt = blockIdx.x; mt < num_mt; mt += gridDim.x) {
            const int gp = mt * 128 + quad * 32 + lane;
            const int b = gp / (h * w), rem = gp - b * (h * w);
            const int y = rem / w, x = rem - y * w;
            __half* obase = out + ((long)b * (hr + 2 * pad) + (long)y * R + pad) * orow_pitch + ((long)x * R + pad) * NF;
            for (int ch = 0; ch < NCH; ++ch, ++u) {
                const uint32_t acc = u & 1, aph = (u >> 1) & 1;
                mbar_wait(&tfull[acc], aph);
                tc_fence_after();
#pragma unroll 1
                for (int sp = sp_lo; sp < sp_hi; ++sp) {
                    const int uv = ch * SUB + sp;
                    __half* op = obase + (long)(uv / R) * orow_pitch + (uv % R) * NF;
                    const float* bs = sbias + uv * NF;
#pragma unroll
                    for (int c0 = 0; c0 < NF; c0 += 32) {
                        uint32_t r[32];
                        tmem_ld32(tmem_base + acc * 256 + sp * NF + c0 + ((uint32_t)(quad * 32) << 16), r);
                        tmem_ld_wait();
#pragma unroll
                        for (int v = 0; v < 2; ++v) {            // 32 B per store: one sector transaction, not two
                            uint4 q[2];
                            uint32_t* pq = reinterpret_cast<uint32_t*>(q);
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const int c = c0 + v * 16 + 2 * e;
                                pq[e] = gelu_pair_h2(f2_pack(__uint_as_float(r[v * 16 + 2 * e]), __uint_as_float(r[v * 16 + 2 * e + 1])),
                                                     *reinterpret_cast<const uint64_t*>(bs + c));
                            }
                            stg256(op + c0 + v * 16, q[0], q[1]);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, 512);
}

template <int R>
static int launch_tail_up_umma_r(const __half* A, const __half* Wt, const float* bias, __half* out, int B, int h, int w,
                                 int pad, cudaStream_t s) {
    using CF = TuCfg<R>;
    const int M = B * h * w;
    CUtensorMap mapA, mapW;
    {
        const uint64_t dims[2] = {NF, (uint64_t)M}, str[2] = {2, NF * 2};
        const uint32_t box[2] = {NF, 128};
        M2T_TRY(make_tensor_map(&mapA, A, 2, 2, dims, str, box, 3));
    }
    {
        const uint64_t dims[2] = {NF, (uint64_t)CF::N}, str[2] = {2, NF * 2};
        const uint32_t box[2] = {NF, (uint32_t)CF::NT};
        M2T_TRY(make_tensor_map(&mapW, Wt, 2, 2, dims, str, box, 3));
    }
    M2T_ENSURE_SMEM(tail_up_umma_kernel<R>, CF::SMEM);
    const int num_mt = M / 128;
    const int grid = num_mt < device_sm_count() ? num_mt : device_sm_count();
    M2T_CUDA(launch_pdl(tail_up_umma_kernel<R>, dim3(grid), dim3(TU_THREADS), CF::SMEM, s, mapA, mapW, bias, out, M, h, w, pad));
    return M2T_OK;
}

int launch_tail_up_umma(const __half* A, const __half* Wt, const float* bias, __half* out, int B, int h, int w, int r,
                        int pad, cudaStream_t s) {
    if ((h * w) % 128) { set_error("tail_up: %dx%d pixels per image is not a multiple of 128", h, w); return M2T_E_ARG; }
    if (r == 2) return launch_tail_up_umma_r<2>(A, Wt, bias, out, B, h, w, pad, s);
    if (r == 3) return launch_tail_up_umma_r<3>(A, Wt, bias, out, B, h, w, pad, s);
    set_error("tail_up: shuffle factor %d", r);
    return M2T_E_UNSUPPORTED;
}

// ---- final 3x3 conv ------------------------------------------------------------------------------------------
constexpr int TO_TH = 16, TO_TW = 8, TO_HW = TO_TW + 2;
constexpr uint32_t TO_TILE_BYTES = (TO_TH + 2) * TO_HW * 128;       // 23040
constexpr uint32_t TO_STAGE = 23 * 1024;
constexpr int TO_STAGES = 4;
constexpr int TO_N = 16;                                            // 3 output channels padded to the MMA minimum
constexpr uint32_t TO_W_BYTES = 9 * TO_N * 128;                      // 18432
constexpr uint32_t TO_OFF_A = TO_W_BYTES;
constexpr uint32_t TO_OFF_BAR = TO_OFF_A + TO_STAGES * TO_STAGE;
constexpr uint32_t TO_SMEM_U = 1024 + TO_OFF_BAR + 256;

__global__ void __launch_bounds__(192, 1)
tail_out_umma_kernel(const __grid_constant__ CUtensorMap mapT, const __grid_constant__ CUtensorMap mapW,
                     float* __restrict__ y, int Bc, int hout, int wout, int b0, float rgb_range) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + TO_OFF_BAR);
    uint64_t* full = bars;
    uint64_t* empty = bars + TO_STAGES;
    uint64_t* wfull = bars + 2 * TO_STAGES;
    uint64_t* tfull = bars + 2 * TO_STAGES + 1;
    uint64_t* tempty = bars + 2 * TO_STAGES + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TO_STAGES + 5);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_x = (wout + TO_TW - 1) / TO_TW, tiles_y = (hout + TO_TH - 1) / TO_TH;   // crop: skip padding tiles
    const int per_img = tiles_x * tiles_y;
    const int ntiles = Bc * per_img;

    if (warp == 5) tmem_alloc(tmem_slot, 32);
    if (tid == 128) {
        for (int s = 0; s < TO_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(wfull, 1);
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
        mbar_fence_init();
        tma_prefetch_desc(&mapT);
        tma_prefetch_desc(&mapW);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();

    if (warp == 4) {
        // TMA producer: warp-uniform loop, one elected lane issues
        if (elect_one_sync()) {
            mbar_expect_tx(wfull, TO_W_BYTES);
            tma_load_2d(sm, &mapW, wfull, 0, 0);
        }
        pdl_wait();
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int bl = tile / per_img, r = tile - bl * per_img;
            const int y0 = (r / tiles_x) * TO_TH, x0 = (r % tiles_x) * TO_TW;
            const uint32_t s = it % TO_STAGES, ph = (it / TO_STAGES) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            if (elect_one_sync()) {
                mbar_expect_tx(&full[s], TO_TILE_BYTES);
                // the tensor has a 1-pixel border: halo origin (y0-1, x0-1) is (y0, x0) in its coordinates
                tma_load_4d(sm + TO_OFF_A + s * TO_STAGE, &mapT, &full[s], 0, x0, y0, bl);
            }
            __syncwarp();
        }
    } else if (warp == 5) {
        // MMA issuer: warp-uniform loop, one elected lane issues
        constexpr uint32_t idesc = umma_idesc_f16(128, TO_N);
        constexpr uint64_t tmpl_a = umma_smem_desc(0, 16, TO_HW * 128, UMMA_LAYOUT_SW128);
        constexpr uint64_t tmpl_b = umma_smem_desc(0, 16, 1024, UMMA_LAYOUT_SW128);
        mbar_wait(wfull, 0);
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const uint32_t s = it % TO_STAGES, ph = (it / TO_STAGES) & 1;
            const uint32_t acc = it & 1, aph = (it >> 1) & 1;
            mbar_wait(&tempty[acc], aph ^ 1);
            mbar_wait(&full[s], ph);
            tc_fence_after();
            if (elect_one_sync()) {
                const uint64_t da0 = umma_desc_at(tmpl_a, base + TO_OFF_A + s * TO_STAGE);
                const uint64_t db0 = umma_desc_at(tmpl_b, base);
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t da = da0 + (uint64_t)((((tap / 3) * TO_HW + (tap % 3)) * 128 + k * 32) >> 4);
                        const uint64_t db = db0 + (uint64_t)((tap * TO_N * 128 + k * 32) >> 4);
                        umma_f16_ss(tmem_base + acc * TO_N, da, db, idesc, (tap | k) ? 1u : 0u);
                    }
                }
                umma_commit(&empty[s]);
                umma_commit(&tfull[acc]);
            }
            __syncwarp();
        }
    } else {
        pdl_wait();
        uint32_t it = 0;
        const long plane = (long)hout * wout;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int bl = tile / per_img, r = tile - bl * per_img;
            const int oy = (r / tiles_x) * TO_TH + (tid >> 3), ox = (r % tiles_x) * TO_TW + (tid & 7);
            const uint32_t acc = it & 1, aph = (it >> 1) & 1;
            mbar_wait(&tfull[acc], aph);
            tc_fence_after();
            uint32_t rr[8];
            tmem_ld8(tmem_base + acc * TO_N + ((uint32_t)(warp * 32) << 16), rr);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (oy < hout && ox < wout) {
                float* yp = y + (long)(b0 + bl) * 3 * plane + (long)oy * wout + ox;
#pragma unroll
                for (int o = 0; o < 3; ++o)     // columns 3..5: the weights' fp16 rounding residual x 2^11 (pack.cu)
                    yp[o * plane] = fminf(fmaxf(__uint_as_float(rr[o]) + __uint_as_float(rr[3 + o]) * (1.f / 2048.f),
                                                0.f), rgb_range);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) tmem_dealloc(tmem_base, 32);
}

// T: fp16 [Bc][hp+2][wp+2][64] with the reflected 1-pixel border already filled
int launch_tail_out_umma(const __half* T, const __half* Wc, float* y, int Bc, int hp, int wp, int hout, int wout, int b0,
                         float rgb_range, cudaStream_t s) {
    CUtensorMap mapT, mapW;
    {
        const uint64_t dims[4] = {NF, (uint64_t)wp + 2, (uint64_t)hp + 2, (uint64_t)Bc};
        const uint64_t str[4] = {2, NF * 2, (uint64_t)(wp + 2) * NF * 2, (uint64_t)(hp + 2) * (wp + 2) * NF * 2};
        const uint32_t box[4] = {NF, TO_HW, TO_TH + 2, 1};
        M2T_TRY(make_tensor_map(&mapT, T, 2, 4, dims, str, box, 3));
    }
    {
        const uint64_t dims[2] = {NF, 9 * TO_N}, str[2] = {2, NF * 2};
        const uint32_t box[2] = {NF, 9 * TO_N};
        M2T_TRY(make_tensor_map(&mapW, Wc, 2, 2, dims, str, box, 3));
    }
    M2T_ENSURE_SMEM(tail_out_umma_kernel, TO_SMEM_U);
    const int ntiles = Bc * ((hout + TO_TH - 1) / TO_TH) * ((wout + TO_TW - 1) / TO_TW);
    const int grid = ntiles < device_sm_count() ? ntiles : device_sm_count();
    M2T_CUDA(launch_pdl(tail_out_umma_kernel, dim3(grid), dim3(192), TO_SMEM_U, s, mapT, mapW, y, Bc, hout, wout, b0, rgb_range));
    return M2T_OK;
}

}  // namespace m2t
