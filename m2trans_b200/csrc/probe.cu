// Hardware probes: tiny single-CTA kernels that expose tcgen05.mma and TMA behaviour to the Python tests
// (tests/test_probes.py).  They pin down, on the real B200, the descriptor conventions the tensor-core
// kernels rely on (swizzle phase vs. start address, stride fields, MN-major operands, OOB zero fill), so
// that a wrong assumption shows up as a failing probe and not as a subtly wrong SR image.
#include "common.cuh"
#include "tma.cuh"
#include "umma.cuh"

namespace m2t {

constexpr uint32_t PROBE_MAX_IMG = 96 * 1024;

__global__ void __launch_bounds__(128, 1)
probe_umma_kernel(const uint8_t* __restrict__ a_img, uint32_t a_bytes, const uint8_t* __restrict__ b_img,
                  uint32_t b_bytes, uint64_t a_desc, uint64_t b_desc, uint32_t a_step, uint32_t b_step, int k_steps,
                  uint32_t idesc, int n_cols, float* __restrict__ out, uint32_t d_lane) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t b_off = (a_bytes + 1023u) & ~1023u;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    for (uint32_t i = t * 16; i < a_bytes; i += 128 * 16)
        *reinterpret_cast<uint4*>(sm + i) = *reinterpret_cast<const uint4*>(a_img + i);
    for (uint32_t i = t * 16; i < b_bytes; i += 128 * 16)
        *reinterpret_cast<uint4*>(sm + b_off + i) = *reinterpret_cast<const uint4*>(b_img + i);
    fence_proxy_async();
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    if (t == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (t == 0) {
        uint64_t da = umma_desc_advance(a_desc, base);
        uint64_t db = umma_desc_advance(b_desc, base + b_off);
        for (int k = 0; k < k_steps; ++k) {
            umma_f16_ss(tmem_base + (d_lane << 16), da, db, idesc, k > 0 ? 1u : 0u);
            da = umma_desc_advance(da, a_step);
            db = umma_desc_advance(db, b_step);
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < n_cols; c0 += 8) {
        uint32_t r[8];
        tmem_ld8(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (c0 + i < n_cols) out[(long)(warp * 32 + lane) * n_cols + c0 + i] = __uint_as_float(r[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

__global__ void __launch_bounds__(128, 1)
probe_tma_kernel(const __grid_constant__ CUtensorMap map, int rank, int c0, int c1, int c2, int c3, int c4,
                 uint32_t box_bytes, uint8_t* __restrict__ out) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
    const int t = threadIdx.x;
    // poison so that bytes TMA does not write are recognisable
    for (uint32_t i = t * 4; i < box_bytes; i += 128 * 4) *reinterpret_cast<uint32_t*>(sm + i) = 0x7E7E7E7Eu;
    if (t == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    fence_proxy_async();
    __syncthreads();
    if (t == 0) {
        mbar_expect_tx(&bar, box_bytes);
        if (rank == 2) tma_load_2d(sm, &map, &bar, c0, c1);
        else if (rank == 3) tma_load_3d(sm, &map, &bar, c0, c1, c2);
        else if (rank == 4) tma_load_4d(sm, &map, &bar, c0, c1, c2, c3);
        else tma_load_5d(sm, &map, &bar, c0, c1, c2, c3, c4);
    }
    mbar_wait(&bar, 0);
    for (uint32_t i = t * 4; i < box_bytes; i += 128 * 4)
        *reinterpret_cast<uint32_t*>(out + i) = *reinterpret_cast<const uint32_t*>(sm + i);
}

}  // namespace m2t

using namespace m2t;

extern "C" int m2t_probe_umma(const void* d_a_image, uint32_t a_bytes, const void* d_b_image, uint32_t b_bytes,
                              uint64_t a_desc, uint64_t b_desc, uint32_t a_step, uint32_t b_step, int k_steps,
                              uint32_t idesc, int n_cols, float* d_out, void* stream) {
    if (!d_a_image || !d_b_image || !d_out) { set_error("probe_umma: null pointer"); return M2T_E_ARG; }
    if (a_bytes % 16 || b_bytes % 16 || a_bytes > PROBE_MAX_IMG || b_bytes > PROBE_MAX_IMG) {
        set_error("probe_umma: images must be multiples of 16 bytes and at most %u bytes", PROBE_MAX_IMG);
        return M2T_E_ARG;
    }
    const uint32_t d_lane = ((uint32_t)n_cols >> 16) & 0xFFu;   // bits 16..23 of n_cols: TMEM lane offset of D
    n_cols &= 0xFFFF;
    if (n_cols < 8 || n_cols > 512 || k_steps < 1 || k_steps > 64) { set_error("probe_umma: bad n_cols/k_steps"); return M2T_E_ARG; }
    // always the maximum, so that a wrong stride hypothesis reads stale shared memory instead of faulting
    const size_t smem = 200 * 1024;
    M2T_ENSURE_SMEM(probe_umma_kernel, 200 * 1024);
    probe_umma_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(
        static_cast<const uint8_t*>(d_a_image), a_bytes, static_cast<const uint8_t*>(d_b_image), b_bytes, a_desc, b_desc,
        a_step, b_step, k_steps, idesc, n_cols, d_out, d_lane);
    M2T_LAUNCH_CHECK("probe_umma_kernel");
    return M2T_OK;
}

extern "C" int m2t_probe_tma(const void* d_tensor, int elem_bytes, int rank, const uint64_t* dims,
                             const uint64_t* strides_bytes, const uint32_t* box, int swizzle, const int32_t* coords,
                             void* d_out, uint32_t out_bytes, void* stream) {
    if (!d_tensor || !dims || !strides_bytes || !box || !coords || !d_out) { set_error("probe_tma: null pointer"); return M2T_E_ARG; }
    if (rank < 2 || rank > 5) { set_error("probe_tma: rank %d", rank); return M2T_E_ARG; }
    uint64_t box_bytes = (uint64_t)elem_bytes;
    for (int i = 0; i < rank; ++i) box_bytes *= box[i];
    if (box_bytes != out_bytes || box_bytes > 200 * 1024 || box_bytes % 16) { set_error("probe_tma: box is %llu bytes, out %u", (unsigned long long)box_bytes, out_bytes); return M2T_E_ARG; }
    CUtensorMap map;
    M2T_TRY(make_tensor_map(&map, d_tensor, elem_bytes, rank, dims, strides_bytes, box, swizzle));
    M2T_ENSURE_SMEM(probe_tma_kernel, 204 * 1024);
    probe_tma_kernel<<<1, 128, box_bytes + 1024, (cudaStream_t)stream>>>(map, rank, coords[0], coords[1],
                                                                         rank > 2 ? coords[2] : 0, rank > 3 ? coords[3] : 0,
                                                                         rank > 4 ? coords[4] : 0, (uint32_t)box_bytes,
                                                                         static_cast<uint8_t*>(d_out));
    M2T_LAUNCH_CHECK("probe_tma_kernel");
    return M2T_OK;
}
