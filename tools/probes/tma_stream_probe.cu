// Probe (development aid): how fast can ONE SM stream an L2-resident weight matrix through shared memory with TMA boxes of
// 128 rows x 128 B (16 KB, 128-byte swizzle), as a function of the number of boxes in flight?  (attn_z<256> streams 278 KB of
// weights per window pair through a 4-slot ring and measures ~32 B/clk.)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_bin/tma_stream_probe tools/probes/tma_stream_probe.cu
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
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
constexpr int ROWS = 288 + 256;         // MQ [288][256] + WV [256][256]: 17 boxes of 128 x 64 ch... (4 column chunks each)
constexpr int NBOX = (ROWS / 128) * 4;  // 16 boxes of 16 KB = 256 KB per pass
constexpr int REP = 32;

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap map, long long* out, int depth) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t full[12];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 12; ++i) mbar_init(&full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        const int total = NBOX * REP;
        int issued = 0;
        for (; issued < depth; ++issued) {
            const int bx = issued % NBOX;
            mbar_expect(&full[issued % depth], 16384);
            tma_load_2d(sm + (issued % depth) * 16384, &map, &full[issued % depth], (bx & 3) * 64, (bx >> 2) * 128);
        }
        for (int i = 0; i < total; ++i) {
            const int s = i % depth;
            mbar_wait(&full[s], (i / depth) & 1);
            if (issued < total) {                 // slot consumed at once: refill (an MMA would hold it ~130 clk longer)
                const int bx = issued % NBOX;
                mbar_expect(&full[s], 16384);
                tma_load_2d(sm + s * 16384, &map, &full[s], (bx & 3) * 64, (bx >> 2) * 128);
                ++issued;
            }
        }
        out[blockIdx.x] = clock64() - t0;
    }
}

int main() {
    __half* W;
    cudaMalloc(&W, (size_t)ROWS * 256 * 2);
    cudaMemset(W, 0, (size_t)ROWS * 256 * 2);
    long long* out;
    cudaMallocManaged(&out, 148 * sizeof(long long));
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    CUtensorMap map;
    const cuuint64_t dims[2] = {256, ROWS};
    const cuuint64_t str[1] = {512};
    const cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
    CUresult r = ((EncodeFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, W, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * 16384 + 1024);
    for (int grid : {1, 2, 128, 148}) {
        for (int depth : {1, 2, 4, 6, 8, 12}) {
            for (int rep = 0; rep < 2; ++rep) {
                probe<<<grid, 128, 12 * 16384 + 1024>>>(map, out, depth);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("depth %d: %s\n", depth, cudaGetErrorString(e)); return 1; }
            }
            long long mx = 0;
            for (int b = 0; b < grid; ++b) mx = out[b] > mx ? out[b] : mx;
            printf("grid %3d depth %2d: %6.1f B/clk per SM (%7.0f clk per 256 KB pass)\n", grid, depth, 262144.0 * REP / mx, (double)mx / REP);
        }
    }
    return 0;
}
