// CUDA-core GEMM used by the M2T_VAR_SIMT_QKV variant of the TBlock qkv 1x1 conv (ref
// M2Trans_network.py:307): out[m][n] = sum_k A[m][k] * Wt[n][k], fp16 operands, fp32 accumulate,
// fp16 result -- the same operand/accumulator precision as the tcgen05 path, so the two variants
// differ only in summation order.
#include "common.cuh"

namespace m2t {

constexpr int GBM = 64, GBN = 64, GBK = 16, GLD = GBK + 8;  // row stride 48 B: conflict-free 16 B reads

__device__ __forceinline__ void fma8(float& acc, const uint4& a, const uint4& b) {
    const __half2* ha = reinterpret_cast<const __half2*>(&a);
    const __half2* hb = reinterpret_cast<const __half2*>(&b);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 fa = __half22float2(ha[i]), fb = __half22float2(hb[i]);
        acc = fmaf(fa.x, fb.x, acc);
        acc = fmaf(fa.y, fb.y, acc);
    }
}

__global__ void __launch_bounds__(256)
gemm_simt_kernel(const __half* __restrict__ A, const __half* __restrict__ Wt, __half* __restrict__ out, int M,
                 int N, int K) {
    __shared__ __align__(16) __half As[GBM * GLD];
    __shared__ __align__(16) __half Ws[GBN * GLD];
    const int t = threadIdx.x;
    const int m0 = blockIdx.x * GBM, n0 = blockIdx.y * GBN;
    const int tx = t & 15, ty = t >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += GBK) {
        {
            const int tt = t & 127, row = tt >> 1, kc = (tt & 1) * 8;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (t < 128) {
                if (m0 + row < M) v = *reinterpret_cast<const uint4*>(A + (long)(m0 + row) * K + k0 + kc);
                *reinterpret_cast<uint4*>(&As[row * GLD + kc]) = v;
            } else {
                if (n0 + row < N) v = *reinterpret_cast<const uint4*>(Wt + (long)(n0 + row) * K + k0 + kc);
                *reinterpret_cast<uint4*>(&Ws[row * GLD + kc]) = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kc = 0; kc < GBK; kc += 8) {
            uint4 a[4], w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const uint4*>(&As[(ty * 4 + i) * GLD + kc]);
#pragma unroll
            for (int j = 0; j < 4; ++j) w[j] = *reinterpret_cast<const uint4*>(&Ws[(tx + 16 * j) * GLD + kc]);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) fma8(acc[i][j], a[i], w[j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx + 16 * j;
            if (n < N) out[(long)m * N + n] = __float2half_rn(acc[i][j]);
        }
    }
}

int launch_gemm_simt(const __half* A, const __half* Wt, __half* out, int M, int N, int K, cudaStream_t s) {
    if (K % GBK != 0) { set_error("gemm_simt: K=%d not a multiple of %d", K, GBK); return M2T_E_ARG; }
    dim3 grid(cdiv(M, GBM), cdiv(N, GBN));
    gemm_simt_kernel<<<grid, 256, 0, s>>>(A, Wt, out, M, N, K);
    M2T_LAUNCH_CHECK("gemm_simt_kernel");
    return M2T_OK;
}

}  // namespace m2t
