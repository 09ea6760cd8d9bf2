// Probe (development aid, not part of the library): how fast does one SM move a window pair's worth of 32-byte pieces
// (128 level-2 pixels x 16 sub-pixels x 16 fp16 channels = 64 KB) between shared memory and a [H][W][64] fp16 tensor
//   (a) with 16 TMA box stores / loads of {16 ch, 1 dx, 8 lx, 1 dy, 16 ly} (rows of 32 bytes), and
//   (b) with 256-bit st.global / ld.global, one 32-byte sector per lane and instruction, as the attention glue does?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_bin/tma_rows_probe tools/probes/tma_rows_probe.cu
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(m), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

constexpr int HP = 512, WP = 512, REP = 64;

__global__ void __launch_bounds__(256) probe(const __grid_constant__ CUtensorMap map, __half* Y, long long* out, int mode) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar;
    const int tid = threadIdx.x;
    const int p = blockIdx.x;                       // pair: 16 windows across, 8 pairs down
    const int lx0 = 8 * (p % 16), ly0 = 16 * (p / 16);
    for (int i = tid; i < 65536 / 16; i += 256) reinterpret_cast<uint4*>(sm)[i] = make_uint4(i, p, 0x3c003c00u, 0x3c003c00u);
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    long long t0 = 0, t1 = 0, t2 = 0;
    if (mode == 0 || mode == 1) {                   // TMA stores: until the source may be reused / until the data is out
        if (tid == 0) {
            t0 = clock64();
            for (int r = 0; r < REP; ++r) {
                for (int s = 0; s < 16; ++s) tma_store_5d(&map, sm + s * 4096, 0, s & 3, lx0 / 1, s >> 2, ly0);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if (mode == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                else asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            }
            t1 = clock64();
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            t2 = clock64();
        }
    } else if (mode == 2) {                         // TMA loads
        if (tid == 0) {
            t0 = clock64();
            for (int r = 0; r < REP; ++r) {
                mbar_expect(&bar, 65536);
                for (int s = 0; s < 16; ++s) tma_load_5d(sm + s * 4096, &map, &bar, 0, s & 3, lx0, s >> 2, ly0);
                mbar_wait(&bar, r & 1);
            }
            t1 = t2 = clock64();
        }
    } else {                                        // LSU: thread = (row m, half): 8 sub-pixels x 32 B each, like the glue
        const int m = tid & 127, half = tid >> 7;
        const int ly = ly0 + (m >> 3), lx = lx0 + (m & 7);
        __syncthreads();
        t0 = clock64();
        uint4 acc = make_uint4(0, 0, 0, 0);
        for (int r = 0; r < REP; ++r) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int s = half * 8 + j;
                __half* g = Y + (((long)(4 * ly + (s >> 2)) * WP) + 4 * lx + (s & 3)) * 64;
                if (mode == 3) {
                    const uint4 v0 = make_uint4(r, s, m, 1), v1 = make_uint4(r, s, m, 2);
                    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(g), "r"(v0.x), "r"(v0.y), "r"(v0.z), "r"(v0.w), "r"(v1.x), "r"(v1.y), "r"(v1.z), "r"(v1.w) : "memory");
                } else {
                    uint4 v0, v1;
                    asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];" : "=r"(v0.x), "=r"(v0.y), "=r"(v0.z), "=r"(v0.w), "=r"(v1.x), "=r"(v1.y), "=r"(v1.z), "=r"(v1.w) : "l"(g) : "memory");
                    acc.x ^= v0.x ^ v1.w; acc.y += v0.y + v1.z;
                }
            }
        }
        if (mode == 4 && acc.x == 0x12345u && acc.y == 77u) Y[0] = __float2half(1.f);
        __syncthreads();
        t1 = t2 = clock64();
    }
    if (tid == 0) { out[3 * blockIdx.x] = t1 - t0; out[3 * blockIdx.x + 1] = t2 - t0; }
}

int main() {
    __half* Y;
    cudaMalloc(&Y, (size_t)HP * WP * 64 * 2);
    cudaMemset(Y, 0, (size_t)HP * WP * 64 * 2);
    long long* out;
    cudaMallocManaged(&out, 3 * 148 * sizeof(long long));
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    CUtensorMap map;
    // (16 ch of branch 2) x dx x lx x dy x ly
    const cuuint64_t dims[5] = {16, 4, WP / 4, 4, HP / 4};
    const cuuint64_t str[4] = {128, 512, (cuuint64_t)WP * 128, (cuuint64_t)4 * WP * 128};
    const cuuint32_t box[5] = {16, 1, 8, 1, 16}, es[5] = {1, 1, 1, 1, 1};
    CUresult r = ((EncodeFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, Y + 32, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 1024);
    const char* names[5] = {"TMA store, source reusable", "TMA store, complete", "TMA load", "LSU 256-bit stores", "LSU 256-bit loads"};
    for (int grid : {1, 128}) {
        for (int mode = 0; mode < 5; ++mode) {
            for (int rep = 0; rep < 2; ++rep) {
                probe<<<grid, 256, 65536 + 1024>>>(map, Y, out, mode);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
            }
            long long mx = 0, mn = 1LL << 60;
            for (int b = 0; b < grid; ++b) { mx = out[3 * b] > mx ? out[3 * b] : mx; mn = out[3 * b] < mn ? out[3 * b] : mn; }
            printf("grid %3d  %-28s %7.0f .. %7.0f clk per 64 KB pair tile (2048 pieces of 32 B)\n", grid, names[mode], (double)mn / REP, (double)mx / REP);
        }
    }
    // correctness of the box mapping: one store, then check a few pieces
    cudaMemset(Y, 0, (size_t)HP * WP * 64 * 2);
    probe<<<128, 256, 65536 + 1024>>>(map, Y, out, 1);
    cudaDeviceSynchronize();
    uint32_t h[8];
    // pair 17 (lx0 = 8, ly0 = 16), sub-pixel s = 6 (dy 1, dx 2), row m = 9 (ly' 1, lx' 1): staging offset s*4096 + m*32
    const long px = ((long)(4 * (16 + 1) + 1) * WP + 4 * (8 + 1) + 2);
    cudaMemcpy(h, Y + px * 64 + 32, 32, cudaMemcpyDeviceToHost);
    printf("piece check: got i=%u p=%u (want i=%d p=17)\n", h[0], h[1], (6 * 4096 + 9 * 32) / 16);
    return 0;
}
