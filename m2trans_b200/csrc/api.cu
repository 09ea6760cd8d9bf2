// C ABI (include/m2trans_b200.h): plan, weight packing, the forward orchestration and the per-stage
// entry points.  The forward is a fixed sequence of kernel launches on the caller's stream; it never
// synchronises, allocates or touches host memory, so the Python side can capture it in a CUDA graph.
#include <stdarg.h>
#include <stdlib.h>

#include <new>

#include <vector>

#include "common.cuh"

namespace m2t {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) at %s [%s:%d]", (int)e, cudaGetErrorString(e), what, file, line);
    return M2T_E_CUDA;
}

int device_sm_count() {
    static int cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

bool pdl_enabled() {
    static int cached = -1;
    if (cached < 0) {
        const char* e = getenv("M2T_NO_PDL");      // development switch for A/B timing
        cached = (e && e[0] == '1') ? 0 : 1;
    }
    return cached == 1;
}
bool pdl_in_graph() {
    static int cached = -1;
    if (cached < 0) {
        const char* e = getenv("M2T_GRAPH_PDL");   // development switch: "0" drops the PDL attribute during stream capture
        cached = (e && e[0] == '0') ? 0 : 1;
    }
    return cached == 1;
}

// ---- per-launch profiling of one eager forward (development aid) ------------------------------------------
struct ProfRec { std::vector<cudaEvent_t> ev; std::vector<const void*> fn; };
static thread_local ProfRec* g_prof = nullptr;
bool prof_active() { return g_prof != nullptr; }
void prof_mark(cudaStream_t s, const void* kernel) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, s);
    g_prof->ev.push_back(e);
    g_prof->fn.push_back(kernel);
}

int check_device() {
    int dev = 0, major = 0, minor = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { set_error("no CUDA device: %s", cudaGetErrorString(e)); return M2T_E_DEVICE; }
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    if (major != 10) {
        set_error("device %d is sm_%d%d; this library contains sm_100a code only (no fallback)", dev, major, minor);
        return M2T_E_DEVICE;
    }
    return M2T_OK;
}

}  // namespace m2t

using namespace m2t;

struct m2t_plan {
    m2t_cfg cfg;
    Geom g;
    PackedLayout L;
    int tail_chunk;            // images per tail pass
    // workspace byte offsets
    size_t o_res, o_x, o_y, o_ylo, o_z, o_qkv, o_o, o_h3, o_h4, o_lo[4], o_stats, o_munorm, o_xr, o_t1, ws_bytes;
    int n_launches;
};

static inline int tail_r0(int scale) { return scale == 4 ? 2 : scale; }
// The tensor the last 3x3 conv reads carries a 1-pixel reflected border ring (T1 for x2/x3, T2 for x4).
static inline size_t tail_t1_bytes(int scale, int B, int Hp, int Wp) {
    const int r0 = tail_r0(scale), pad = scale == 4 ? 0 : 1;
    return align_up((size_t)B * (Hp * r0 + 2 * pad) * (Wp * r0 + 2 * pad) * NF * 2, 256);
}
static inline size_t tail_t2_bytes(int scale, int B, int Hp, int Wp) {
    return scale == 4 ? align_up((size_t)B * (Hp * 4 + 2) * (Wp * 4 + 2) * NF * 2, 256) : 0;
}

extern "C" {

const char* m2t_last_error(void) { return g_err; }
#ifndef M2T_SRC_HASH
#define M2T_SRC_HASH "unhashed-build"
#endif
// the build script compiles the sha256 of the sources in (m2trans_b200/build.py: source_hash / built_hash)
const char* m2t_version(void) { return "m2trans_b200 0.2 (sm_100a) m2t-src-hash:" M2T_SRC_HASH; }

int m2t_query_device(int* sm_major, int* sm_minor, int* sm_count) {
    int dev = 0;
    M2T_CUDA(cudaGetDevice(&dev));
    int v = 0;
    if (sm_major) { M2T_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev)); *sm_major = v; }
    if (sm_minor) { M2T_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev)); *sm_minor = v; }
    if (sm_count) { M2T_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev)); *sm_count = v; }
    return M2T_OK;
}

int m2t_num_params(int scale, int n_blocks) { return 6 + 14 * n_blocks + (scale == 4 ? 5 : 3); }

size_t m2t_packed_weight_bytes(int scale, int n_blocks) {
    PackedLayout L;
    if (make_packed_layout(scale, n_blocks, &L) != M2T_OK) return 0;
    return L.total;
}

size_t m2t_packed_offset(int scale, int n_blocks, const char* name) {
    PackedLayout L;
    if (!name || make_packed_layout(scale, n_blocks, &L) != M2T_OK) return (size_t)-1;
    if (!strcmp(name, "head_w")) return L.head_w;
    if (!strcmp(name, "head_b")) return L.head_b;
    if (!strcmp(name, "t0w")) return L.t0w;
    if (!strcmp(name, "t0b")) return L.t0b;
    if (!strcmp(name, "tcw")) return L.tcw;
    if (scale == 4 && !strcmp(name, "t3w")) return L.t3w;
    if (scale == 4 && !strcmp(name, "t3b")) return L.t3b;
    int i = -1, a = -1;
    char what[32];
    if (sscanf(name, "body.%d.attn%d.%31s", &i, &a, what) == 3 && i >= 0 && i < n_blocks && a >= 1 && a <= 4) {
        const AttnW& A = L.blk[i].attn[a - 1];
        if (!strcmp(what, "wqkv")) return A.wqkv;
        if (!strcmp(what, "wqkv_f")) return A.wqkv_f;
        if (!strcmp(what, "relf")) return A.relf;
        if (!strcmp(what, "relx")) return A.relx;
        if (!strcmp(what, "mq")) return A.mq;
        return (size_t)-1;
    }
    if (sscanf(name, "body.%d.%31s", &i, what) == 2 && i >= 0 && i < n_blocks) {
        if (!strcmp(what, "ffw")) return L.blk[i].ffw;
        if (!strcmp(what, "ffw2")) return L.blk[i].ffw2;
        if (!strcmp(what, "ffb")) return L.blk[i].ffb;
    }
    return (size_t)-1;
}

int m2t_pack_weights(int scale, int n_blocks, const float* const* d_params, int n_params, void* d_packed,
                     void* stream) {
    if (!d_params || !d_packed) { set_error("pack_weights: null pointer"); return M2T_E_ARG; }
    M2T_TRY(check_device());
    PackedLayout L;
    M2T_TRY(make_packed_layout(scale, n_blocks, &L));
    return pack_weights_impl(L, d_params, n_params, static_cast<uint8_t*>(d_packed), (cudaStream_t)stream);
}

size_t m2t_tail_scratch_bytes(int scale, int B, int Hp, int Wp) {
    return tail_t1_bytes(scale, B, Hp, Wp) + tail_t2_bytes(scale, B, Hp, Wp);
}

int m2t_plan_create(const m2t_cfg* cfg, m2t_plan** out) {
    if (!cfg || !out) { set_error("plan_create: null pointer"); return M2T_E_ARG; }
    *out = nullptr;
    if (cfg->n_feats != NF) { set_error("n_feats %d: the engine is built for 64", cfg->n_feats); return M2T_E_UNSUPPORTED; }
    if (cfg->colors != 3) { set_error("colors %d: the engine is built for 3", cfg->colors); return M2T_E_UNSUPPORTED; }
    if (cfg->batch < 1 || cfg->height < 1 || cfg->width < 1) { set_error("bad input shape"); return M2T_E_ARG; }
    m2t_plan* p = new (std::nothrow) m2t_plan();
    if (!p) { set_error("out of host memory"); return M2T_E_ARG; }
    p->cfg = *cfg;
    int rc = make_packed_layout(cfg->scale, cfg->n_blocks, &p->L);
    if (rc != M2T_OK) { delete p; return rc; }
    Geom& g = p->g;
    g.B = cfg->batch; g.H = cfg->height; g.W = cfg->width; g.scale = cfg->scale;
    g.Hp = (g.H + 31) / 32 * 32; g.Wp = (g.W + 31) / 32 * 32;
    // F.pad(mode='reflect') requires pad < dim (ref :85 raises otherwise)
    if (g.Hp - g.H >= g.H || g.Wp - g.W >= g.W) {
        set_error("reflect padding %dx%d -> %dx%d needs pad < size (the reference raises here too)", g.H, g.W, g.Hp, g.Wp);
        delete p;
        return M2T_E_UNSUPPORTED;
    }
    const size_t P = (size_t)g.B * g.Hp * g.Wp;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    p->o_res = take(P * NF * 4);
    p->o_x = take(P * NF * 4);
    p->o_y = take(P * NF * 2);
    p->o_z = take(P * NB * 2);
    p->o_qkv = take(P * NB * 3 * 2);
    p->o_o = take(P * NB * 2);
    p->o_h3 = take(P * NB * 2);
    p->o_h4 = take(P * NB * 2);
    p->o_ylo = take(P * NF * 2);                                  // rounding residual of Y * 2^11 (precise mode)
    for (int a = 0; a < 4; ++a) p->o_lo[a] = take(P * NB * 2);   // fp16 rounding residuals of t_1..t_4 (residual path)
    p->o_stats = take((size_t)(cfg->n_blocks + 1) * g.B * NF * 2 * sizeof(double));
    p->o_munorm = take((size_t)g.B * NF * sizeof(float2));
    // tail scratch: only what the selected tail variant really writes (the fused x2 tail needs none, the fused x4 tail
    // only T1); images go through the tail in equal chunks of at most ~2 GiB of intermediates (32-bit element indices)
    const bool tail_fused = cfg->scale != 3 && !(cfg->variant & (M2T_VAR_SIMT_TAIL | M2T_VAR_UNFUSED_TAIL));
    const size_t per_img = tail_fused ? (cfg->scale == 4 ? tail_t1_bytes(4, 1, g.Hp, g.Wp) : 0)
                                      : m2t_tail_scratch_bytes(cfg->scale, 1, g.Hp, g.Wp);
    int chunk = per_img ? (int)(((size_t)2 << 30) / per_img) : g.B;
    if (chunk < 1) chunk = 1;
    if (chunk > g.B) chunk = g.B;
    const int tail_passes = (g.B + chunk - 1) / chunk;
    chunk = (g.B + tail_passes - 1) / tail_passes;                // equal passes
    p->tail_chunk = chunk;
    p->o_xr = take(P * NF * 2);
    p->o_t1 = take(tail_fused ? per_img * chunk + 256 : m2t_tail_scratch_bytes(cfg->scale, chunk, g.Hp, g.Wp));
    p->ws_bytes = off;
    const bool qkv16_fused = !(cfg->variant & (M2T_VAR_SIMT_ATTN | M2T_VAR_SIMT_QKV | M2T_VAR_SPLIT_QKV16));
    const bool qkv_fused = !(cfg->variant & (M2T_VAR_SIMT_ATTN | M2T_VAR_SIMT_QKV | M2T_VAR_SPLIT_QKV));
    const int per_block = (cfg->variant & M2T_VAR_SIMT_ATTN) ? (1 + 4 * 4 + 1)
                                                             : (1 + 4 * 2 + 1 - (qkv16_fused ? 1 : 0) - (qkv_fused ? 3 : 0));
    // tail per image chunk: x3 and the unfused variants run up [, up], border, out; otherwise [up,] fused
    const int per_tail = tail_fused ? (cfg->scale == 4 ? 2 : 1) : (cfg->scale == 4 ? 4 : 3);
    p->n_launches = 1 /*head*/ + cfg->n_blocks * per_block + tail_passes * per_tail;
    *out = p;
    return M2T_OK;
}

void m2t_plan_destroy(m2t_plan* plan) { delete plan; }
size_t m2t_workspace_bytes(const m2t_plan* plan) { return plan ? plan->ws_bytes : 0; }
int m2t_plan_num_launches(const m2t_plan* plan) { return plan ? plan->n_launches : 0; }
int m2t_plan_padded(const m2t_plan* plan, int* Hp, int* Wp) {
    if (!plan) { set_error("null plan"); return M2T_E_ARG; }
    if (Hp) *Hp = plan->g.Hp;
    if (Wp) *Wp = plan->g.Wp;
    return M2T_OK;
}
size_t m2t_workspace_offset(const m2t_plan* plan, const char* name) {
    if (!plan || !name) return (size_t)-1;
    if (!strcmp(name, "res")) return plan->o_res;
    if (!strcmp(name, "x")) return plan->o_x;
    if (!strcmp(name, "y")) return plan->o_y;
    if (!strcmp(name, "stats")) return plan->o_stats;      // fp64 [n_blocks + 1][B][64][2] InstanceNorm sums (sum, sum of squares)
    return (size_t)-1;
}

// ---- stage dispatch ------------------------------------------------------------------------------
static int run_qkv(uint32_t variant, const __half* Z, const __half* wqkv, __half* QKV, int M, int C, cudaStream_t s) {
    if (!(variant & M2T_VAR_SIMT_QKV)) return launch_qkv_umma(Z, wqkv, QKV, M, C, s);
    return launch_gemm_simt(Z, wqkv, QKV, M, 3 * C, C, s);
}
static int run_attn(uint32_t variant, int C, const __half* QKV, const float* relf, const __half* relx, __half* O,
                    int B, int h, int w, cudaStream_t s) {
    if (!(variant & M2T_VAR_SIMT_ATTN)) return launch_attn_umma(C, QKV, relx, O, B, h, w, s);
    return launch_attn_simt(C, QKV, relf, O, B, h, w, s);
}
static int run_ffconv(uint32_t variant, const __half* Y, const __half* ffw, const float* ffb, const float* Xin,
                      float* Xout, double* stats, const Geom& g, cudaStream_t s, const float* res = nullptr,
                      __half* xr = nullptr) {
    if (!(variant & M2T_VAR_SIMT_CONV)) return launch_ffconv_umma(Y, ffw, ffb, Xin, Xout, stats, g, s, res, xr);
    return launch_ffconv_simt(Y, ffw, ffb, Xin, Xout, stats, g, s, res, xr);
}
// XR = fp16(res + x) of B images (ref :70), written by the last ff conv's epilogue
static int run_tail(uint32_t variant, int scale, const PackedLayout& L, const uint8_t* W, const __half* XR, float* y,
                    int B, int b0, const Geom& g, float rgb_range, uint8_t* scratch, cudaStream_t s) {
    const bool tc = !(variant & M2T_VAR_SIMT_TAIL);
    const int r0 = tail_r0(scale);
    __half* T1 = reinterpret_cast<__half*>(scratch);
    __half* T2 = reinterpret_cast<__half*>(scratch + tail_t1_bytes(scale, B, g.Hp, g.Wp));
    const int hout = g.H * scale, wout = g.W * scale;
    const __half* w0 = reinterpret_cast<const __half*>(W + L.t0w);
    const float* b0p = reinterpret_cast<const float*>(W + L.t0b);
    const __half* wc = reinterpret_cast<const __half*>(W + L.tcw);
    const int pad1 = scale == 4 ? 0 : 1;
    if (tc && scale != 3 && !(variant & M2T_VAR_UNFUSED_TAIL)) {
        // last PixelShuffle(2) stage + GELU + 3x3 conv + clamp + crop in one kernel (tail_fused.cu)
        const bool tiled = (variant & M2T_VAR_TILE_TAIL) != 0;
        if (scale == 2) return tiled ? launch_tail_fused(XR, w0, b0p, wc, y, B, g.Hp, g.Wp, hout, wout, b0, rgb_range, s)
                                     : launch_tail_strip(XR, w0, b0p, wc, y, B, g.Hp, g.Wp, hout, wout, b0, rgb_range, s);
        M2T_TRY(launch_tail_up_umma(XR, w0, b0p, T1, B, g.Hp, g.Wp, 2, 0, s));
        const __half* w3 = reinterpret_cast<const __half*>(W + L.t3w);
        const float* b3 = reinterpret_cast<const float*>(W + L.t3b);
        return tiled ? launch_tail_fused(T1, w3, b3, wc, y, B, 2 * g.Hp, 2 * g.Wp, hout, wout, b0, rgb_range, s)
                     : launch_tail_strip(T1, w3, b3, wc, y, B, 2 * g.Hp, 2 * g.Wp, hout, wout, b0, rgb_range, s);
    }
    if (tc) M2T_TRY(launch_tail_up_umma(XR, w0, b0p, T1, B, g.Hp, g.Wp, r0, pad1, s));
    else M2T_TRY(launch_tail_up_simt(XR, w0, b0p, T1, B, g.Hp, g.Wp, r0, pad1, s));
    __half* Tl = T1;
    int hl = g.Hp * r0, wl = g.Wp * r0;
    if (scale == 4) {
        const __half* w3 = reinterpret_cast<const __half*>(W + L.t3w);
        const float* b3 = reinterpret_cast<const float*>(W + L.t3b);
        if (tc) M2T_TRY(launch_tail_up_umma(T1, w3, b3, T2, B, hl, wl, 2, 1, s));
        else M2T_TRY(launch_tail_up_simt(T1, w3, b3, T2, B, hl, wl, 2, 1, s));
        Tl = T2; hl *= 2; wl *= 2;
    }
    M2T_TRY(launch_reflect_border(Tl, B, hl, wl, s));
    if (tc) return launch_tail_out_umma(Tl, wc, y, B, hl, wl, hout, wout, b0, rgb_range, s);
    return launch_tail_out_simt(Tl, wc, y, B, hl, wl, hout, wout, b0, rgb_range, s);
}

int m2t_forward(const m2t_plan* plan, const void* d_packed, const float* d_x, float* d_y, void* d_workspace,
                void* stream) {
    return m2t_forward_phases(plan, d_packed, d_x, d_y, d_workspace, stream, M2T_PHASE_ALL);
}

int m2t_forward_phases(const m2t_plan* plan, const void* d_packed, const float* d_x, float* d_y, void* d_workspace,
                       void* stream, uint32_t phases) {
    if (!plan || !d_packed || !d_workspace) { set_error("forward: null pointer"); return M2T_E_ARG; }
    if (((phases & M2T_PHASE_HEAD) && !d_x) || ((phases & M2T_PHASE_TAIL) && !d_y)) { set_error("forward: null pointer"); return M2T_E_ARG; }
    if (!(phases & M2T_PHASE_ALL)) { set_error("forward: no phase selected"); return M2T_E_ARG; }
    if (((uintptr_t)d_workspace & 255) || ((uintptr_t)d_packed & 255)) {
        set_error("forward: workspace / packed weights must be 256-byte aligned");
        return M2T_E_ARG;
    }
    M2T_TRY(check_device());
    cudaStream_t s = (cudaStream_t)stream;
    const Geom& g = plan->g;
    const PackedLayout& L = plan->L;
    const uint32_t var = plan->cfg.variant;
    const uint8_t* W = static_cast<const uint8_t*>(d_packed);
    uint8_t* ws = static_cast<uint8_t*>(d_workspace);
    float* res = reinterpret_cast<float*>(ws + plan->o_res);
    float* X = reinterpret_cast<float*>(ws + plan->o_x);
    __half* Y = reinterpret_cast<__half*>(ws + plan->o_y);
    __half* Z = reinterpret_cast<__half*>(ws + plan->o_z);
    __half* QKV = reinterpret_cast<__half*>(ws + plan->o_qkv);
    __half* O = reinterpret_cast<__half*>(ws + plan->o_o);
    __half* XR = reinterpret_cast<__half*>(ws + plan->o_xr);
    double* stats = reinterpret_cast<double*>(ws + plan->o_stats);
    float2* munorm = reinterpret_cast<float2*>(ws + plan->o_munorm);
    const size_t stat_stride = (size_t)g.B * NF * 2;
    const int npix = g.Hp * g.Wp;

    if (phases & M2T_PHASE_HEAD) {
        M2T_CUDA(cudaMemsetAsync(stats, 0, (size_t)(plan->cfg.n_blocks + 1) * stat_stride * sizeof(double), s));
        M2T_TRY(launch_head(d_x, reinterpret_cast<const float*>(W + L.head_w), reinterpret_cast<const float*>(W + L.head_b),
                            res, stats, g, s));
    }
    const float* Xin = res;
    // Default path: Haar-folded weights + branch glue fused into the attention epilogue (11 launches per CFTM).
    // The CUDA-core attention variant keeps the explicit prep / post kernels (18 launches per CFTM).
    const bool fused = !(var & M2T_VAR_SIMT_ATTN);
    // precise mode (M2T_VAR_PRECISE_*): t_k residuals on the residual path + split-precision ff-conv weights;
    // default on for x2 / x3, off for x4
    const bool precise = (var & M2T_VAR_PRECISE_ON) || (!(var & M2T_VAR_PRECISE_OFF) && plan->cfg.scale != 4);
    for (int i = 0; (phases & M2T_PHASE_BODY) && i < plan->cfg.n_blocks; ++i) {
        if (!fused) M2T_TRY(launch_stats_finalize(stats + i * stat_stride, munorm, g.B, npix, s));
        if (fused) {
            // t_1 = n_1 and n_k/2 (k = 2..4) in their consumers' space-to-depth layouts, one pass over X (ref :135-137)
            __half* Tb[4] = {Z, O, reinterpret_cast<__half*>(ws + plan->o_h3), reinterpret_cast<__half*>(ws + plan->o_h4)};
            __half* Lb[4];
            for (int a = 0; a < 4; ++a) Lb[a] = precise ? reinterpret_cast<__half*>(ws + plan->o_lo[a]) : nullptr;
            M2T_TRY(launch_branch_prep_all(Xin, stats + i * stat_stride, Tb[0], Lb[0], Tb[1], Tb[2], Tb[3], g, s));
            for (int a = 0; a < 4; ++a) {
                const int lv = branch_level(a), C = branch_ch(a);
                const int h = g.Hp >> lv, w = g.Wp >> lv;
                const AttnW& A = L.blk[i].attn[a];
                AttnFuse fz;
                fz.T = Tb[a]; fz.Y = Y; fz.Tnext = a < 3 ? Tb[a + 1] : nullptr;
                fz.Tlo = Lb[a]; fz.Tnext_lo = a < 3 ? Lb[a + 1] : nullptr;
                fz.Ylo = precise ? reinterpret_cast<__half*>(ws + plan->o_ylo) : nullptr;
                fz.branch = a; fz.Hp = g.Hp; fz.Wp = g.Wp;
                if (a == 0 && !(var & (M2T_VAR_SIMT_QKV | M2T_VAR_SPLIT_QKV16))) {
                    // branch 1: qkv conv inside the attention kernel (attn16_qkv.cu)
                    M2T_TRY(launch_attn16_qkv(Tb[a], reinterpret_cast<const __half*>(W + A.wqkv_f),
                                              reinterpret_cast<const __half*>(W + A.relx), g.B, h, w, s, fz));
                    continue;
                }
                if (a > 0 && !(var & (M2T_VAR_SIMT_QKV | M2T_VAR_SPLIT_QKV))) {
                    // branches 2-4: qkv conv inside the attention kernel, q / k / v never formed (attn_z.cu)
                    M2T_TRY(launch_attn_z(C, Tb[a], reinterpret_cast<const __half*>(W + A.mq),
                                          reinterpret_cast<const __half*>(W + A.wqkv_f) + (size_t)2 * C * C, g.B, h, w, s, fz,
                                          (var & M2T_VAR_AZ_PAIRED) != 0));
                    continue;
                }
                M2T_TRY(run_qkv(var, Tb[a], reinterpret_cast<const __half*>(W + A.wqkv_f), QKV, g.B * h * w, C, s));
                M2T_TRY(launch_attn_umma(C, QKV, reinterpret_cast<const __half*>(W + A.relx), nullptr, g.B, h, w, s, &fz));
            }
        } else
        for (int a = 0; a < 4; ++a) {
            const int lv = branch_level(a), C = branch_ch(a);
            const int h = g.Hp >> lv, w = g.Wp >> lv;
            const AttnW& A = L.blk[i].attn[a];
            M2T_TRY(launch_branch_prep(lv, a, Xin, munorm, Y, Z, g, s));
            M2T_TRY(run_qkv(var, Z, reinterpret_cast<const __half*>(W + A.wqkv), QKV, g.B * h * w, C, s));
            M2T_TRY(run_attn(var, C, QKV, reinterpret_cast<const float*>(W + A.relf),
                             reinterpret_cast<const __half*>(W + A.relx), O, g.B, h, w, s));
            M2T_TRY(launch_branch_post(lv, a, O, Xin, munorm, Y, g, s));
        }
        const bool last = i == plan->cfg.n_blocks - 1;      // the last block also emits fp16(res + x) for the tail
        if (precise && !(var & M2T_VAR_SIMT_CONV))
            M2T_TRY(((var & M2T_VAR_W2_PAIR) ? launch_ffconv_pair : launch_ffconv_umma_w2)(Y, reinterpret_cast<const __half*>(ws + plan->o_ylo),
                                          reinterpret_cast<const __half*>(W + L.blk[i].ffw2),
                                          reinterpret_cast<const float*>(W + L.blk[i].ffb), Xin, X,
                                          stats + (i + 1) * stat_stride, g, s, last ? res : nullptr, last ? XR : nullptr));
        else
        M2T_TRY(run_ffconv(var, Y, reinterpret_cast<const __half*>(W + L.blk[i].ffw),
                           reinterpret_cast<const float*>(W + L.blk[i].ffb), Xin, X, stats + (i + 1) * stat_stride, g, s,
                           last ? res : nullptr, last ? XR : nullptr));
        Xin = X;
    }
    const size_t img_stride = (size_t)npix * NF;
    for (int b0 = 0; (phases & M2T_PHASE_TAIL) && b0 < g.B; b0 += plan->tail_chunk) {
        const int nb = g.B - b0 < plan->tail_chunk ? g.B - b0 : plan->tail_chunk;
        M2T_TRY(run_tail(var, plan->cfg.scale, L, W, XR + b0 * img_stride, d_y, nb, b0, g, plan->cfg.rgb_range,
                         ws + plan->o_t1, s));
    }
    return M2T_OK;
}

int m2t_debug_profile_forward(const m2t_plan* plan, const void* d_packed, const float* d_x, float* d_y,
                              void* d_workspace, void* stream, char* text, size_t cap) {
    if (!text || cap < 64) { set_error("profile_forward: text buffer too small"); return M2T_E_ARG; }
    ProfRec rec;
    cudaStream_t s = (cudaStream_t)stream;
    g_prof = &rec;
    prof_mark(s, nullptr);                                   // start marker
    const int rc = m2t_forward(plan, d_packed, d_x, d_y, d_workspace, stream);
    g_prof = nullptr;
    if (rc != M2T_OK) return rc;
    M2T_CUDA(cudaStreamSynchronize(s));
    struct Agg { const void* fn; int n; float ms; };
    std::vector<Agg> agg;
    float total = 0.f;
    for (size_t i = 1; i < rec.ev.size(); ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, rec.ev[i - 1], rec.ev[i]);
        total += ms;
        bool found = false;
        for (auto& a : agg) if (a.fn == rec.fn[i]) { a.n++; a.ms += ms; found = true; break; }
        if (!found) agg.push_back({rec.fn[i], 1, ms});
    }
    for (cudaEvent_t e : rec.ev) cudaEventDestroy(e);
    size_t off = 0;
    for (const auto& a : agg) {
        const char* name = "?";
        cudaFuncGetName(&name, a.fn);
        const int w = snprintf(text + off, cap - off, "%9.1f us %5.1f%%  x%-3d %.150s\n", a.ms * 1e3f, 100.f * a.ms / total, a.n, name);
        if (w < 0 || (size_t)w >= cap - off) break;
        off += (size_t)w;
    }
    snprintf(text + off, cap - off, "total %.1f us (eager, one event per launch, no PDL overlap, warm L2)\n", total * 1e3f);
    return M2T_OK;
}

// ---- per-stage entry points -------------------------------------------------------------------------
static int make_geom(Geom* g, int B, int Hp, int Wp) {
    if (B < 1 || Hp < 32 || Wp < 32 || Hp % 32 || Wp % 32) { set_error("stage: padded frame %dx%dx%d must be multiples of 32", B, Hp, Wp); return M2T_E_ARG; }
    g->B = B; g->H = Hp; g->W = Wp; g->Hp = Hp; g->Wp = Wp; g->scale = 0;
    return M2T_OK;
}

int m2t_stage_head(const float* d_x, const float* d_head_w, const float* d_head_b, float* d_res, double* d_stats,
                   int B, int H, int W, void* stream) {
    M2T_TRY(check_device());
    Geom g;
    g.B = B; g.H = H; g.W = W; g.Hp = (H + 31) / 32 * 32; g.Wp = (W + 31) / 32 * 32; g.scale = 0;
    if (g.Hp - H >= H || g.Wp - W >= W) { set_error("stage_head: reflect padding undefined for %dx%d", H, W); return M2T_E_UNSUPPORTED; }
    return launch_head(d_x, d_head_w, d_head_b, d_res, d_stats, g, (cudaStream_t)stream);
}

int m2t_stage_stats_finalize(const double* d_stats, float* d_munorm, int B, int npix, void* stream) {
    M2T_TRY(check_device());
    return launch_stats_finalize(d_stats, reinterpret_cast<float2*>(d_munorm), B, npix, (cudaStream_t)stream);
}

int m2t_stage_branch_prep(int branch, const float* d_X, const float* d_munorm, const void* d_Y, void* d_Z, int B,
                          int Hp, int Wp, void* stream) {
    M2T_TRY(check_device());
    if (branch < 0 || branch > 3) { set_error("branch %d", branch); return M2T_E_ARG; }
    Geom g;
    M2T_TRY(make_geom(&g, B, Hp, Wp));
    return launch_branch_prep(branch_level(branch), branch, d_X, reinterpret_cast<const float2*>(d_munorm),
                              static_cast<const __half*>(d_Y), static_cast<__half*>(d_Z), g, (cudaStream_t)stream);
}

int m2t_stage_branch_post(int branch, const void* d_O, const float* d_X, const float* d_munorm, void* d_Y, int B,
                          int Hp, int Wp, void* stream) {
    M2T_TRY(check_device());
    if (branch < 0 || branch > 3) { set_error("branch %d", branch); return M2T_E_ARG; }
    Geom g;
    M2T_TRY(make_geom(&g, B, Hp, Wp));
    return launch_branch_post(branch_level(branch), branch, static_cast<const __half*>(d_O), d_X,
                              reinterpret_cast<const float2*>(d_munorm), static_cast<__half*>(d_Y), g,
                              (cudaStream_t)stream);
}

int m2t_stage_qkv(uint32_t variant, const void* d_Z, const void* d_wqkv, void* d_QKV, int M, int C, void* stream) {
    M2T_TRY(check_device());
    if (C != 16 && C != 64 && C != 256) { set_error("qkv: C=%d", C); return M2T_E_UNSUPPORTED; }
    if (M % 64) { set_error("qkv: M=%d must be a multiple of 64 (whole 8x8 blocks)", M); return M2T_E_ARG; }
    return run_qkv(variant, static_cast<const __half*>(d_Z), static_cast<const __half*>(d_wqkv),
                   static_cast<__half*>(d_QKV), M, C, (cudaStream_t)stream);
}

int m2t_stage_attn(uint32_t variant, int C, const void* d_QKV, const float* d_relf, const void* d_relx, void* d_O,
                   int B, int h, int w, void* stream) {
    M2T_TRY(check_device());
    return run_attn(variant, C, static_cast<const __half*>(d_QKV), d_relf, static_cast<const __half*>(d_relx),
                    static_cast<__half*>(d_O), B, h, w, (cudaStream_t)stream);
}

int m2t_stage_attn_z(int C, const void* d_T, const void* d_mq, const void* d_wv, void* d_Y, void* d_Tnext, int branch,
                     int B, int h, int w, void* stream) {
    if (!d_T || !d_mq || !d_wv || !d_Y) { set_error("stage_attn_z: null pointer"); return M2T_E_ARG; }
    if (C != 64 && C != 256) { set_error("stage_attn_z: C=%d", C); return M2T_E_UNSUPPORTED; }
    if (branch < 1 || branch > 3 || B < 1) { set_error("stage_attn_z: branch %d, B %d", branch, B); return M2T_E_ARG; }
    M2T_TRY(check_device());
    const int lv = C == 64 ? 1 : 2;
    AttnFuse fz{};
    fz.T = static_cast<const __half*>(d_T); fz.Y = static_cast<__half*>(d_Y); fz.Tnext = static_cast<__half*>(d_Tnext);
    fz.branch = branch; fz.Hp = h << lv; fz.Wp = w << lv;
    return launch_attn_z(C, fz.T, static_cast<const __half*>(d_mq), static_cast<const __half*>(d_wv), B, h, w,
                         (cudaStream_t)stream, fz);
}

int m2t_stage_ffconv(uint32_t variant, const void* d_Y, const void* d_ffw, const float* d_ffb, const float* d_Xin,
                     float* d_Xout, double* d_stats, int B, int Hp, int Wp, void* stream) {
    M2T_TRY(check_device());
    Geom g;
    M2T_TRY(make_geom(&g, B, Hp, Wp));
    return run_ffconv(variant, static_cast<const __half*>(d_Y), static_cast<const __half*>(d_ffw), d_ffb, d_Xin,
                      d_Xout, d_stats, g, (cudaStream_t)stream);
}

int m2t_stage_tail(uint32_t variant, int scale, int n_blocks, const void* d_packed, const void* d_XR, float* d_y,
                   int B, int b0, int H, int W, float rgb_range, void* d_scratch, void* stream) {
    M2T_TRY(check_device());
    if (!d_packed || !d_XR || !d_y || !d_scratch) { set_error("stage_tail: null pointer"); return M2T_E_ARG; }
    PackedLayout L;
    M2T_TRY(make_packed_layout(scale, n_blocks, &L));
    Geom g;
    g.B = B; g.H = H; g.W = W; g.Hp = (H + 31) / 32 * 32; g.Wp = (W + 31) / 32 * 32; g.scale = scale;
    return run_tail(variant, scale, L, static_cast<const uint8_t*>(d_packed), static_cast<const __half*>(d_XR), d_y, B,
                    b0, g, rgb_range, static_cast<uint8_t*>(d_scratch), (cudaStream_t)stream);
}

int m2t_debug_attn_timing(long long* host64) {
    if (!host64) { set_error("null pointer"); return M2T_E_ARG; }
    M2T_TRY(read_attn_timing(host64));
    M2T_TRY(read_tail_timing(host64 + 256));
    M2T_TRY(read_conv_timing(host64 + 320));
    return read_qkv_timing(host64 + 384);
}

int m2t_debug_az_timing(long long* host192) {
    if (!host192) { set_error("null pointer"); return M2T_E_ARG; }
    return read_az_timing(host192);
}

}  // extern "C"
