// uint8 HWC image -> fp32 CHW / 255 on the device (ref datas/benchmark.py:62-69: ndarray2tensor(lr) / 255.,
// ref utils.py:237-240 ndarray2tensor = HWC -> CHW, float).  The loader keeps images as uint8 arrays in RAM; sending those
// bytes and converting here moves 4x fewer bytes over PCIe than the reference's float tensors (3 B instead of 12 B per
// pixel).  Bit-exact: fp32(u8) / 255 with IEEE division, the same operation torch performs.
#include "common.cuh"

namespace m2t {
namespace {

// one thread per 4 consecutive pixels of a row: 12 bytes in (three 32-bit loads when aligned), three float4 stores
__global__ void __launch_bounds__(256)
u8hwc_to_f32chw_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, long npx4, long plane, float denom) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npx4) return;
    const long px = i * 4, b = px / plane, o = px - b * plane;      // plane is a multiple of 4 (checked by the caller)
    const uint32_t* p = reinterpret_cast<const uint32_t*>(src + px * 3);
    const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
    const uint8_t v[12] = {(uint8_t)w0, (uint8_t)(w0 >> 8), (uint8_t)(w0 >> 16), (uint8_t)(w0 >> 24),
                           (uint8_t)w1, (uint8_t)(w1 >> 8), (uint8_t)(w1 >> 16), (uint8_t)(w1 >> 24),
                           (uint8_t)w2, (uint8_t)(w2 >> 8), (uint8_t)(w2 >> 16), (uint8_t)(w2 >> 24)};
    float* d = dst + b * 3 * plane + o;
#pragma unroll
    for (int c = 0; c < 3; ++c)
        *reinterpret_cast<float4*>(d + c * plane) = make_float4(__fdiv_rn((float)v[c], denom), __fdiv_rn((float)v[3 + c], denom),
                                                                __fdiv_rn((float)v[6 + c], denom), __fdiv_rn((float)v[9 + c], denom));
}

__global__ void __launch_bounds__(256)
u8hwc_to_f32chw_scalar_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, long npx, long plane, int colors,
                              float denom) {
    const long px = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (px >= npx) return;
    const long b = px / plane, o = px - b * plane;
    for (int c = 0; c < colors; ++c) dst[(b * colors + c) * plane + o] = __fdiv_rn((float)__ldg(src + px * colors + c), denom);
}

// fp32 CHW in [0, 1] -> uint8 HWC, the inverse conversion: what writing the SR batch to image files does
// (round-to-nearest-even of x * scale, clamped to [0, 255]).  One thread per pixel: three strided 4-byte loads
// (coalesced across the warp), three 1-byte stores into 3 consecutive bytes.
__global__ void __launch_bounds__(256)
f32chw_to_u8hwc_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, long npx, long plane, int colors, float scale) {
    const long px = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (px >= npx) return;
    const long b = px / plane, o = px - b * plane;
    for (int c = 0; c < colors; ++c) {
        const int v = __float2int_rn(src[(b * colors + c) * plane + o] * scale);
        dst[px * colors + c] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
}

}  // namespace
}  // namespace m2t

using namespace m2t;

extern "C" int m2t_f32chw_to_u8hwc(const float* d_src, void* d_dst, int B, int H, int W, int colors, float scale, void* stream) {
    if (!d_src || !d_dst) { set_error("u8 writer: null pointer"); return M2T_E_ARG; }
    if (B < 1 || H < 1 || W < 1 || (colors != 1 && colors != 3) || !(scale > 0.f)) {
        set_error("u8 writer: bad B %d H %d W %d colors %d scale %g", B, H, W, colors, (double)scale);
        return M2T_E_ARG;
    }
    M2T_TRY(check_device());
    const long plane = (long)H * W, npx = plane * B;
    f32chw_to_u8hwc_kernel<<<(unsigned)((npx + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_src, static_cast<uint8_t*>(d_dst), npx, plane, colors, scale);
    M2T_LAUNCH_CHECK("f32chw_to_u8hwc");
    return M2T_OK;
}

extern "C" int m2t_u8hwc_to_f32chw(const void* d_src, float* d_dst, int B, int H, int W, int colors, float denom, void* stream) {
    if (!d_src || !d_dst) { set_error("u8 loader: null pointer"); return M2T_E_ARG; }
    if (B < 1 || H < 1 || W < 1 || (colors != 1 && colors != 3) || !(denom > 0.f)) {
        set_error("u8 loader: bad B %d H %d W %d colors %d denom %g", B, H, W, colors, (double)denom);
        return M2T_E_ARG;
    }
    M2T_TRY(check_device());
    cudaStream_t s = (cudaStream_t)stream;
    const long plane = (long)H * W, npx = plane * B;
    if (colors == 3 && plane % 4 == 0 && (reinterpret_cast<uintptr_t>(d_src) & 3) == 0 && (reinterpret_cast<uintptr_t>(d_dst) & 15) == 0) {
        u8hwc_to_f32chw_kernel<<<(unsigned)((npx / 4 + 255) / 256), 256, 0, s>>>(static_cast<const uint8_t*>(d_src), d_dst, npx / 4, plane, denom);
    } else {
        u8hwc_to_f32chw_scalar_kernel<<<(unsigned)((npx + 255) / 256), 256, 0, s>>>(static_cast<const uint8_t*>(d_src), d_dst, npx, plane, colors, denom);
    }
    M2T_LAUNCH_CHECK("u8hwc_to_f32chw");
    return M2T_OK;
}
