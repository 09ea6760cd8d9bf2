// Shared declarations for the m2trans_b200 sm_100a engine.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/m2trans_b200.h"

namespace m2t {

constexpr int NF = 64;          // n_feats (ref configs/M2Trans_x*.yml: n_feats 64)
constexpr int NB = 16;          // channels per CFTM branch = nf/4 (ref M2Trans_network.py:137)
constexpr int BLK = 8;          // TBlock block_size (ref :119-122)
constexpr int WIN = 10;         // block + 2*halo
constexpr int NKEY = WIN * WIN; // 100 keys per window
constexpr float IN_EPS = 1e-5f; // nn.InstanceNorm2d eps (ref :127)

// ---- error plumbing ---------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define M2T_CUDA(call)                                                        \
    do {                                                                      \
        cudaError_t e__ = (call);                                             \
        if (e__ != cudaSuccess) return m2t::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

#define M2T_LAUNCH_CHECK(name)                                                \
    do {                                                                      \
        cudaError_t e__ = cudaGetLastError();                                 \
        if (e__ != cudaSuccess) return m2t::cuda_fail(e__, name, __FILE__, __LINE__); \
    } while (0)

#define M2T_TRY(expr)                         \
    do {                                      \
        int rc__ = (expr);                    \
        if (rc__ != M2T_OK) return rc__;      \
    } while (0)

// opt a kernel in to > 48 KB of dynamic shared memory, once per device
#define M2T_ENSURE_SMEM(fn, bytes)                                                              \
    do {                                                                                        \
        static unsigned char done__[64];                                                        \
        int dev__ = 0;                                                                          \
        M2T_CUDA(cudaGetDevice(&dev__));                                                        \
        if (dev__ < 0 || dev__ >= 64 || !done__[dev__]) {                                       \
            M2T_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
            if (dev__ >= 0 && dev__ < 64) done__[dev__] = 1;                                    \
        }                                                                                       \
    } while (0)

int device_sm_count();
int check_device();   // M2T_OK on an sm_100 device, M2T_E_DEVICE otherwise (api.cu)

// ---- programmatic dependent launch ---------------------------------------------------------------------
// Every kernel of the forward is launched with programmatic stream serialisation: its CTAs may start while
// the previous kernel drains, run their prologue (barrier init, TMEM alloc, weight loads: nothing that depends
// on the previous kernel) and then block in pdl_wait() until the previous grid has completed and flushed.
// Rule: no thread reads activations or writes ANY global memory before pdl_wait().
bool pdl_enabled();
bool pdl_in_graph();
bool prof_active();                                   // m2t_debug_profile_forward is running on this thread
void prof_mark(cudaStream_t s, const void* kernel);   // records an event after a launch
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                              Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (pdl_enabled() && !prof_active()) ? 1 : 0;
    // Measured on B200 (cfg2, final kernels): graph 1.745 ms, graph + PDL 1.718 ms, eager + PDL 1.725 ms (device
    // time with the host running ahead).  An earlier measurement had PDL hurting inside graphs; that was caused by
    // multi-wave kernels triggering at their START (dependents stole SM slots), fixed by triggering at the end.
    // M2T_GRAPH_PDL=0 drops the attribute during capture for A/B runs.
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cfg.numAttrs && !pdl_in_graph() && cudaStreamIsCapturing(s, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone) cfg.numAttrs = 0;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
    if (prof_active()) prof_mark(s, reinterpret_cast<const void*>(kernel));
    return e;
}
// CTAs of `kernel` that are resident on the device at once (occupancy x SMs); cached per kernel
template <typename K>
inline int resident_ctas(K kernel, int block, size_t smem) {
    static int cached = 0;                         // one instance per kernel type and call site pattern (K is the function type)
    static const void* cached_for = nullptr;
    const void* key = reinterpret_cast<const void*>(kernel);
    if (cached_for != key) {
        int per_sm = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
        cached = per_sm * device_sm_count();
        cached_for = key;
    }
    return cached;
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// Multi-wave grids: the CTAs of the LAST wave let the dependent kernel's launch proceed when they start (all earlier CTAs
// have left by then, so nothing can take slots this grid still needs); the launch latency of the dependent then overlaps
// the last wave instead of following it.  `resident` = resident_ctas(kernel, block, smem).
__device__ __forceinline__ void pdl_trigger_last_wave(int resident) {
    const long id = (long)blockIdx.x + (long)gridDim.x * ((long)blockIdx.y + (long)gridDim.y * blockIdx.z);
    const long n = (long)gridDim.x * gridDim.y * gridDim.z;
    if (n - id <= (long)resident) pdl_trigger();
}

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256).  A thread that owns 32 contiguous, 32-byte-aligned bytes
// touches its sector once instead of twice: scattered per-pixel epilogues are bound by sector transactions.
__device__ __forceinline__ void ldg256(const void* p, uint4& a, uint4& b) {
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint4& a, const uint4& b) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
#endif

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---- packed weights ----------------------------------------------------------------
// Byte offsets into the packed weight blob (all 256-byte aligned).
struct AttnW {
    size_t wqkv;   // fp16 [3C][C]   rows 0..C-1 (q) pre-multiplied by C^-1/2 (exact: power of two)
    size_t wqkv_f; // fp16 [3C][C]   the same with the Haar DWT folded into the columns and the IWT into the v rows
    size_t relf;   // fp32 [20][C/2] rows 0..9 rel_h, rows 10..19 rel_w
    size_t relx;   // fp16 [32][C]   rows 0..9 [rel_h|0], 10..19 [0|rel_w], 20..31 zero (MMA operand form)
    size_t mq;     // fp16 [32+C][C] attn_z.cu: rows 0..19 = Wq'^T rel (10 rel_h, 10 rel_w terms), 20..31 zero, rows 32.. =
                   //                (Wq'^T Wk')^T, i.e. mq[32+n][k] = sum_c Wq'[c][k] Wk'[c][n]; Wq', Wk' = the folded, q-scaled
                   //                rows of wqkv_f before their fp16 rounding
};
struct BlockW {
    AttnW attn[4];
    size_t ffw;    // fp16 [9][64 out][64 in]
    size_t ffw2;   // fp16 [2 halves][9][64 rows][64 in]: rows 0..31 hi, 32..63 residual * 2^11 (split-precision conv)
    size_t ffb;    // fp32 [64]
};
struct PackedLayout {
    int scale, n_blocks;
    size_t head_w;   // fp32 [27][64]  index (c*9 + ky*3 + kx)*64 + o
    size_t head_b;   // fp32 [64]
    BlockW blk[64];
    size_t t0w, t0b; // fp16 [N0][64], fp32 [N0]      N0 = 64*r0^2
    size_t t3w, t3b; // x4 only: fp16 [256][64], fp32 [256]
    size_t tcw;      // fp16 [9][16][64] final 3x3 conv, rows 3..15 zero
    size_t fold_scratch;
    size_t fold_scratch2;   // fp32 [2C][C]: folded q (scaled) and k rows for the mq products
    size_t total;
};
int make_packed_layout(int scale, int n_blocks, PackedLayout* out);
static inline int branch_ch(int a) { return a == 0 ? 16 : (a == 1 ? 64 : 256); }
static inline int branch_level(int a) { return a == 0 ? 0 : (a == 1 ? 1 : 2); }

// ---- stage launchers (each returns M2T_OK or an error code) ------------------------
struct Geom {
    int B, H, W;     // LR input
    int Hp, Wp;      // padded to multiples of 32 (ref :78-86)
    int scale;
};

// head.cu
int launch_head(const float* x, const float* w, const float* b, float* res, double* stats,
                const Geom& g, cudaStream_t s);
int launch_stats_finalize(const double* stats, float2* munorm, int B, int npix, cudaStream_t s);

// branch.cu
int launch_branch_prep(int level, int branch, const float* X, const float2* munorm, const __half* Y,
                       __half* Z, const Geom& g, cudaStream_t s);
int launch_branch_post(int level, int branch, const __half* O, const float* X, const float2* munorm,
                       __half* Y, const Geom& g, cudaStream_t s);
int launch_dwt_nchw(const float* in, float* out, int B, int C, int H, int W, cudaStream_t s);
int launch_iwt_nchw(const float* in, float* out, int B, int C4, int H, int W, cudaStream_t s);
int launch_nchw_to_nhwc_half(const float* in, __half* out, int B, int C, int H, int W, cudaStream_t s);
int launch_nhwc_half_to_nchw(const __half* in, float* out, int B, int C, int H, int W, cudaStream_t s);
int launch_nhwc_to_nchw_f32(const float* in, float* out, int B, int C, int H, int W, cudaStream_t s);

// gemm_simt.cu : out[m][n] = sum_k A[m][k] * Wt[n][k]   (fp16 in, fp32 accumulate, fp16 out)
int launch_gemm_simt(const __half* A, const __half* Wt, __half* out, int M, int N, int K, cudaStream_t s);

// gemm_umma.cu : same contract on tcgen05 (C = 64 or 256; N = 3C, K = C)
int launch_qkv_umma(const __half* Z, const __half* Wqkv, __half* QKV, int M, int C, cudaStream_t s);

// attn_simt.cu : halo attention over QKV [B,h,w,3C] -> O [B,h,w,C]
int launch_attn_simt(int C, const __half* QKV, const float* relf, __half* O, int B, int h, int w,
                     cudaStream_t s);

// attn_umma.cu : same contract on tcgen05 (relx: fp16 [32][C] MMA-operand form of the rel tables)
// Optional fused branch glue for the tensor-core attention epilogue (ref :139-161): instead of storing O it
// writes y_k = O' + t_k into Y and completes the next branch's input in place,
// Tnext = n_{k+1}/2 (pre-filled by branch_prep_all) + y_k/2 = t_{k+1}.  Requires the Haar-folded qkv weights
// (pack.cu) and space-to-depth T tensors.
struct AttnFuse {
    const __half* T;        // this branch's input t_k, space-to-depth fp16 [B,h,w,C]
    __half* Y;              // cat[y1..y4] fp16 NHWC [B,Hp,Wp,64]
    __half* Tnext;          // next branch's space-to-depth tensor holding n_{k+1}/2, or nullptr after branch 4
    // t_k = T + Tlo: the fp16 rounding residual travels beside t_k for the RESIDUAL path (y_k = attention + t_k); the
    // qkv GEMM reads T alone.  Rounding t_k before that add was the largest single contributor to the output error.
    const __half* Tlo;      // same layout as T
    __half* Tnext_lo;       // same layout as Tnext, written here
    __half* Ylo;            // (y_k - fp16(y_k)) * 2^11, same layout as Y: the split-precision ff conv's second A operand
    int branch;             // 0..3
    int Hp, Wp;
};
int launch_branch_prep_all(const float* X, const double* stats, __half* T1, __half* T1lo, __half* H2, __half* H3, __half* H4,
                           const Geom& g, cudaStream_t s);
int launch_attn_umma(int C, const __half* QKV, const __half* relx, __half* O, int B, int h, int w,
                     cudaStream_t s, const AttnFuse* fuse = nullptr);

// attn16_qkv.cu : branch 1 with the qkv conv inside the attention kernel (fused glue only)
int launch_attn16_qkv(const __half* T, const __half* Wqkv, const __half* relx, int B, int h, int w, cudaStream_t s,
                      const AttnFuse& fz);

// attn_z.cu : branches 2-4 (C = 64 / 256) with the qkv conv inside the attention kernel, contractions re-associated so
// that q, k, v never exist (fused glue only).  MQ = AttnW::mq, WV = the v rows of AttnW::wqkv_f.
int launch_attn_z(int C, const __half* T, const __half* MQ, const __half* WV, int B, int h, int w, cudaStream_t s,
                  const AttnFuse& fz, bool paired_only = false);

int read_attn_timing(long long* host64);   // development aid, zeros unless built with -DM2T_TIMING (256 values)
int read_tail_timing(long long* host64);   // the same for the fused tail kernel (64 values)
int read_conv_timing(long long* host64);   // the same for the tcgen05 ff conv (64 values)
int read_qkv_timing(long long* host64);
int read_az_timing(long long* host64);     // the same for attn_z (192 values: 64 per branch 2..4)    // the same for the 256-channel qkv GEMM (64 values)

// conv_simt.cu : X_out = conv3x3_zero(Y) + bias + X_in, plus InstanceNorm partial sums
// res/xr (optional): also write xr = fp16(Xout + res), the tail's first GEMM operand (ref :70)
int launch_ffconv_simt(const __half* Y, const __half* Wp, const float* bias, const float* Xin, float* Xout,
                       double* stats, const Geom& g, cudaStream_t s, const float* res = nullptr,
                       __half* xr = nullptr);

// conv_umma.cu : same contract as an implicit GEMM on tcgen05 fed by TMA
int launch_ffconv_umma(const __half* Y, const __half* Wp, const float* bias, const float* Xin, float* Xout,
                       double* stats, const Geom& g, cudaStream_t s, const float* res = nullptr,
                       __half* xr = nullptr);
// split-precision weights (BlockW::ffw2): every CTA computes 32 output channels with hi and residual weight rows
// and split-precision input: Ylo = fp16 rounding residual of Y * 2^11, same layout as Y
int launch_ffconv_umma_w2(const __half* Y, const __half* Ylo, const __half* Wp2, const float* bias, const float* Xin,
                          float* Xout, double* stats, const Geom& g, cudaStream_t s, const float* res = nullptr,
                          __half* xr = nullptr);
// conv_pair.cu: the same contract as a CTA pair per two tiles (tcgen05 cta_group::2, cluster of 2)
int launch_ffconv_pair(const __half* Y, const __half* Ylo, const __half* Wp2, const float* bias, const float* Xin,
                          float* Xout, double* stats, const Geom& g, cudaStream_t s, const float* res = nullptr,
                          __half* xr = nullptr);

// tail_simt.cu
//   tail_up : A fp16 [B,h,w,64] -> out fp16 [B, r h + 2 pad, r w + 2 pad, 64] (interior only; pad = 0 or 1)
//   reflect_border : fills the pad ring of a [B][h+2][w+2][64] tensor
//   tail_out: T fp16 [Bc][hp+2][wp+2][64] (ring filled) -> y fp32 NCHW images b0.. cropped to hout x wout
int launch_tail_up_simt(const __half* Ain, const __half* Wt, const float* bias, __half* out, int B, int h, int w,
                        int r, int pad, cudaStream_t s);
int launch_reflect_border(__half* T, int B, int h, int w, cudaStream_t s);
int launch_tail_out_simt(const __half* T, const __half* Wc, float* y, int Bc, int hp, int wp, int hout, int wout,
                         int b0, float rgb_range, cudaStream_t s);
// tail_umma.cu : the same two stages on tcgen05
int launch_tail_up_umma(const __half* A, const __half* Wt, const float* bias, __half* out, int B, int h, int w,
                        int r, int pad, cudaStream_t s);
int launch_tail_out_umma(const __half* T, const __half* Wc, float* y, int Bc, int hp, int wp, int hout, int wout,
                         int b0, float rgb_range, cudaStream_t s);

// tail_fused.cu : last PixelShuffle(2) stage (1x1 conv + bias + shuffle + GELU) + 3x3 reflect conv + clamp + crop
int launch_tail_fused(const __half* A, const __half* W1, const float* bias, const __half* Wc, float* y, int Bc, int h,
                      int w, int hout, int wout, int b0, float rgb_range, cudaStream_t s);

// tail_strip.cu : the same stage as tail_fused, strip-marching with warp-specialised GELU / conv-epilogue roles
int launch_tail_strip(const __half* A, const __half* W1, const float* bias, const __half* Wc, float* y, int Bc, int h,
                      int w, int hout, int wout, int b0, float rgb_range, cudaStream_t s);

// pack.cu
int pack_weights_impl(const PackedLayout& L, const float* const* params, int n_params, uint8_t* packed,
                      cudaStream_t s);

}  // namespace m2t
