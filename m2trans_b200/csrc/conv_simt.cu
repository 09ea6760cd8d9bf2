// CUDA-core variant (M2T_VAR_SIMT_CONV) of CFTM.feed_forward + residual (ref M2Trans_network.py:124-126,
// :164): Xout = conv3x3(zero_pad1(Y)) + bias + Xin, Y = cat[y1..y4] as fp16 NHWC, X fp32 NHWC.  The
// epilogue also accumulates the per-(image,channel) sum / sum-of-squares that the NEXT block's
// InstanceNorm needs (ref :135), so the norm never costs a pass of its own.
// Persistent CTAs: the 9x64x64 weight block (fp32 in shared memory) is loaded once per CTA.
#include "common.cuh"
#include "epilogue.cuh"

namespace m2t {

constexpr int CT_H = 8, CT_W = 16;                 // output tile (pixels)
constexpr int CT_TLD = NF + 8;                     // halves per staged input pixel (144 B: conflict-free)
constexpr size_t CS_W = 0;                                               // fp32 [9][64][64]
constexpr size_t CS_T = CS_W + (size_t)9 * NF * NF * 4;                  // fp16 [10][18][72]
constexpr size_t CS_O = CS_T + (size_t)(CT_H + 2) * (CT_W + 2) * CT_TLD * 2;  // fp32 [128][65]
constexpr size_t CS_BYTES = CS_O + (size_t)CT_H * CT_W * EPI_LD * 4;

__global__ void __launch_bounds__(128, 1)
ffconv_simt_kernel(const __half* __restrict__ Y, const __half* __restrict__ Wp, const float* __restrict__ bias,
                   const float* Xin, float* Xout, double* __restrict__ stats, int B,
                   int Hp, int Wpx, const float* __restrict__ res, __half* __restrict__ xr) {
    extern __shared__ __align__(16) uint8_t smem[];
    float* Ws = reinterpret_cast<float*>(smem + CS_W);
    __half* Ts = reinterpret_cast<__half*>(smem + CS_T);
    float* Os = reinterpret_cast<float*>(smem + CS_O);
    __shared__ float red[4][2][NF];
    const int t = threadIdx.x;
    for (int i = t; i < 9 * NF * NF; i += 128) Ws[i] = __half2float(Wp[i]);

    const int tiles_x = Wpx / CT_W, tiles_y = Hp / CT_H;
    const int ntiles = B * tiles_y * tiles_x;
    const int ty = t >> 4, tx = t & 15;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int b = tile / (tiles_y * tiles_x);
        const int r = tile - b * tiles_y * tiles_x;
        const int y0 = (r / tiles_x) * CT_H, x0 = (r % tiles_x) * CT_W;
        __syncthreads();   // previous tile's epilogue is done with Os / Ts; weights visible
        for (int idx = t; idx < (CT_H + 2) * (CT_W + 2) * 8; idx += 128) {
            const int p = idx >> 3, ch = idx & 7;
            const int py = y0 - 1 + p / (CT_W + 2), px = x0 - 1 + p % (CT_W + 2);
            uint4 v = make_uint4(0, 0, 0, 0);
            if (py >= 0 && py < Hp && px >= 0 && px < Wpx)
                v = *reinterpret_cast<const uint4*>(Y + (((long)b * Hp + py) * Wpx + px) * NF + ch * 8);
            *reinterpret_cast<uint4*>(&Ts[p * CT_TLD + ch * 8]) = v;
        }
        __syncthreads();
        float acc[NF];
#pragma unroll
        for (int o = 0; o < NF; ++o) acc[o] = 0.f;
        for (int tap = 0; tap < 9; ++tap) {
            const __half* src = &Ts[((ty + tap / 3) * (CT_W + 2) + tx + tap % 3) * CT_TLD];
            const float* wt = &Ws[tap * NF * NF];
#pragma unroll 2
            for (int ch = 0; ch < 8; ++ch) {
                const uint4 u = *reinterpret_cast<const uint4*>(src + ch * 8);
                const __half2* hp = reinterpret_cast<const __half2*>(&u);
                float in[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f = __half22float2(hp[i]);
                    in[2 * i] = f.x; in[2 * i + 1] = f.y;
                }
#pragma unroll
                for (int o = 0; o < NF; ++o) {
                    const float4 w0 = *reinterpret_cast<const float4*>(&wt[o * NF + ch * 8]);
                    const float4 w1 = *reinterpret_cast<const float4*>(&wt[o * NF + ch * 8 + 4]);
                    float a = acc[o];
                    a = fmaf(in[0], w0.x, a); a = fmaf(in[1], w0.y, a);
                    a = fmaf(in[2], w0.z, a); a = fmaf(in[3], w0.w, a);
                    a = fmaf(in[4], w1.x, a); a = fmaf(in[5], w1.y, a);
                    a = fmaf(in[6], w1.z, a); a = fmaf(in[7], w1.w, a);
                    acc[o] = a;
                }
            }
        }
#pragma unroll
        for (int o = 0; o < NF; ++o) Os[t * EPI_LD + o] = acc[o];
        __syncthreads();
        uint4 xi[16], rv[16];
        EpiStats st;
        st.clear();
        epilogue_load_residual<CT_W>(t, Xin, b, y0, x0, Hp, Wpx, xi, xr != nullptr ? res : nullptr, rv);
        epilogue_apply<CT_W>(Os, t, xi, bias, Xout, st, b, y0, x0, Hp, Wpx, rv, xr);
        epilogue_flush_stats<0>(red, t, st, stats, b);
    }
}

int launch_ffconv_simt(const __half* Y, const __half* Wp, const float* bias, const float* Xin, float* Xout,
                       double* stats, const Geom& g, cudaStream_t s, const float* res, __half* xr) {
    M2T_ENSURE_SMEM(ffconv_simt_kernel, CS_BYTES);
    const int ntiles = g.B * (g.Hp / CT_H) * (g.Wp / CT_W);
    const int grid = ntiles < device_sm_count() ? ntiles : device_sm_count();
    ffconv_simt_kernel<<<grid, 128, CS_BYTES, s>>>(Y, Wp, bias, Xin, Xout, stats, g.B, g.Hp, g.Wp, res, xr);
    M2T_LAUNCH_CHECK("ffconv_simt_kernel");
    return M2T_OK;
}

}  // namespace m2t
