// Weight packing: reference state_dict tensors (fp32, ref M2Trans_network.py registration order, see
// SURVEY.md appendix B.2) -> the engine's operand layouts (common.cuh PackedLayout).  Runs once per
// load_state_dict on the device; nothing here is on the per-forward path.
#include "common.cuh"

namespace m2t {

int make_packed_layout(int scale, int n_blocks, PackedLayout* L) {
    if (scale < 2 || scale > 4) { set_error("scale %d not in {2,3,4}", scale); return M2T_E_UNSUPPORTED; }
    if (n_blocks < 1 || n_blocks > 64) { set_error("n_blocks %d not in 1..64", n_blocks); return M2T_E_UNSUPPORTED; }
    memset(L, 0, sizeof(*L));
    L->scale = scale; L->n_blocks = n_blocks;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    L->head_w = take(27 * NF * 4);
    L->head_b = take(NF * 4);
    for (int i = 0; i < n_blocks; ++i) {
        for (int a = 0; a < 4; ++a) {
            const int C = branch_ch(a);
            L->blk[i].attn[a].wqkv = take((size_t)3 * C * C * 2);
            L->blk[i].attn[a].wqkv_f = take((size_t)3 * C * C * 2);
            L->blk[i].attn[a].relf = take((size_t)20 * (C / 2) * 4);
            L->blk[i].attn[a].relx = take((size_t)32 * C * 2);
            L->blk[i].attn[a].mq = take((size_t)(32 + C) * C * 2);
        }
        L->blk[i].ffw = take((size_t)9 * NF * NF * 2);
        L->blk[i].ffw2 = take((size_t)2 * 9 * NF * NF * 2);
        L->blk[i].ffb = take(NF * 4);
    }
    const int r0 = scale == 4 ? 2 : scale;
    const int N0 = NF * r0 * r0;
    L->t0w = take((size_t)N0 * NF * 2);
    L->t0b = take((size_t)N0 * 4);
    if (scale == 4) {
        L->t3w = take((size_t)256 * NF * 2);
        L->t3b = take((size_t)256 * 4);
    }
    L->tcw = take((size_t)9 * 16 * NF * 2);
    L->fold_scratch = take((size_t)768 * 256 * 4);   // fp32 staging for the Haar folding at pack time
    L->fold_scratch2 = take((size_t)512 * 256 * 4);
    L->total = off;
    return M2T_OK;
}

enum PackMode {
    PK_COPY_F32 = 0,   // dst[i] = src[i]
    PK_CVT_F16,        // dst[i] = half(src[i] * (i < n_scaled ? scale : 1))
    PK_HEAD_W,         // src [64][3][3][3] -> dst [(c*9+tap)][64]
    PK_CONV3_W,        // src [O][64][3][3] -> dst [tap][Opad][64] fp16, rows >= O zero
    PK_RELX,           // src rel_h [10][C/2] (+ rel_w passed as src2) -> dst fp16 [32][C]
    PK_UP_W,           // tail 1x1 conv: src [64 r^2][64] -> dst fp16 row (uv*64 + c) <- src row (c*r^2 + uv)
    PK_UP_B,           // its bias with the same row permutation (fp32)
    PK_CONV3_W2        // src [64][64][3][3] -> dst [half][tap][64 rows][64] fp16: rows 0..31 = hi of output channels
                       // half*32.., rows 32..63 = their rounding residual * 2^11 (split-precision ff conv)
};

__global__ void pack_kernel(int mode, const float* __restrict__ src, const float* __restrict__ src2, void* dstv,
                            int n, int p0, int p1, float scale) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    switch (mode) {
        case PK_COPY_F32: reinterpret_cast<float*>(dstv)[i] = src[i]; break;
        case PK_CVT_F16:
            reinterpret_cast<__half*>(dstv)[i] = __float2half_rn(src[i] * (i < p0 ? scale : 1.f));
            break;
        case PK_HEAD_W: {   // i over dst [27][64]
            const int o = i % NF, k = i / NF;
            reinterpret_cast<float*>(dstv)[i] = src[o * 27 + k];
        } break;
        case PK_CONV3_W: {  // i over dst [9][Opad=p1][64]; p0 = real O
            const int c = i % NF, o = (i / NF) % p1, tap = i / (NF * p1);
            // scale != 0 (the tail's final conv, p0 = 3 of 16 rows): rows p0..2*p0-1 carry the fp16 rounding residual of
            // rows 0..p0-1 times `scale` (2^11); tail_out_umma adds them back, the other consumers read rows < p0 only
            __half v = __half(0);
            if (o < p0) v = __float2half_rn(src[(o * NF + c) * 9 + tap]);
            else if (scale != 0.f && o < 2 * p0) {
                const float wv = src[((o - p0) * NF + c) * 9 + tap];
                v = __float2half_rn((wv - __half2float(__float2half_rn(wv))) * scale);
            }
            reinterpret_cast<__half*>(dstv)[i] = v;
        } break;
        case PK_CONV3_W2: { // i over dst [2][9][64][64]
            const int c = i % NF, row = (i / NF) % NF, tap = (i / (NF * NF)) % 9, hh = i / (NF * NF * 9);
            const int o = hh * 32 + (row & 31);
            const float wv = src[(o * NF + c) * 9 + tap];
            const __half hi = __float2half_rn(wv);
            reinterpret_cast<__half*>(dstv)[i] = row < 32 ? hi : __float2half_rn((wv - __half2float(hi)) * 2048.f);
        } break;
        case PK_RELX: {     // i over dst [32][C=p0]
            const int C = p0, hc = C / 2, c = i % C, row = i / C;
            float v = 0.f;
            if (row < 10 && c < hc) v = src[row * hc + c];
            else if (row >= 10 && row < 20 && c >= hc) v = src2[(row - 10) * hc + (c - hc)];
            reinterpret_cast<__half*>(dstv)[i] = __float2half_rn(v);
        } break;
        case PK_UP_W: {     // i over dst [64 r^2][64]; p0 = r^2
            const int k = i % NF, np = i / NF, uv = np / NF, c = np % NF;
            reinterpret_cast<__half*>(dstv)[i] = __float2half_rn(src[(c * p0 + uv) * NF + k]);
        } break;
        case PK_UP_B: {     // i over dst [64 r^2]
            const int uv = i / NF, c = i % NF;
            reinterpret_cast<float*>(dstv)[i] = src[c * p0 + uv];
        } break;
    }
}

// ---- Haar folding --------------------------------------------------------------------------------------------
// DWT and IWT (ref :203-209, :223-232) are fixed orthogonal maps on the 4^L pixels of a 2^L x 2^L block, so they
// commute into the 1x1 qkv conv: with T the "space-to-depth" tensor (channel s*16+k = pixel s of the block,
// s = dy*2^L + dx, base channel k) and Z = DWT^L(t) = G^T T,
//     q,k = W_{q,k} Z = (W_{q,k} G^T) T          v' = G v = (G W_v G^T) T
// and the attention output computed from v' is IWT^L(attention output) in space-to-depth order.  The engine
// therefore never runs a Haar butterfly: g[s][band] = Hf[b2][P(s)] * Hf[b1][p(s)] is folded into the weights here
// (in fp32, one fp16 rounding at the end).
__device__ __forceinline__ float haar_g(int L, int s, int band) {
    // Hf[band][pos], pos = dy + 2*dx (a,b,c,d of ref :204-207), without the 1/2
    const int sgn[4][4] = {{1, 1, 1, 1}, {-1, -1, 1, 1}, {-1, 1, -1, 1}, {1, -1, -1, 1}};
    if (L == 0) return 1.f;
    if (L == 1) { const int dy = s >> 1, dx = s & 1; return 0.5f * sgn[band][dy + 2 * dx]; }
    const int dy = s >> 2, dx = s & 3;
    const int P = (dy >> 1) + 2 * (dx >> 1), p = (dy & 1) + 2 * (dx & 1);
    return 0.25f * sgn[band >> 2][P] * sgn[band & 3][p];
}

// step 1: rows of v: tmp[(s_o,k_o)][j] = sum_band g[s_o][band] * W[2C + band*16 + k_o][j]; q,k rows copied
__global__ void fold_rows_kernel(const float* __restrict__ W, float* __restrict__ tmp, int C, int L) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * C * C) return;
    const int n = i / C, j = i - n * C;
    if (n < 2 * C) { tmp[i] = W[i]; return; }
    const int nb = C / NB, so = (n - 2 * C) / NB, ko = (n - 2 * C) % NB;
    float acc = 0.f;
    for (int band = 0; band < nb; ++band) acc = fmaf(haar_g(L, so, band), W[(long)(2 * C + band * NB + ko) * C + j], acc);
    tmp[i] = acc;
}
// step 2: columns of all rows: out[n][(s,k)] = sum_band tmp[n][band*16 + k] * g[s][band]; q rows scaled
__global__ void fold_cols_kernel(const float* __restrict__ tmp, __half* __restrict__ out, float* __restrict__ out32,
                                 int C, int L, float qscale) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * C * C) return;
    const int n = i / C, j = i - n * C, nb = C / NB, s = j / NB, k = j % NB;
    float acc = 0.f;
    for (int band = 0; band < nb; ++band) acc = fmaf(tmp[(long)n * C + band * NB + k], haar_g(L, s, band), acc);
    const float v = acc * (n < C ? qscale : 1.f);
    out[i] = __float2half_rn(v);
    if (n < 2 * C) out32[i] = v;                    // unrounded q / k rows for mq_kernel
}

// attn_z.cu operand (AttnW::mq): fp16 [32 + C][C].  qk = fp32 [2C][C]: rows 0..C-1 = Wq' (folded, scaled), C..2C-1 = Wk'.
//   row r < 10        mq[r][k] = sum_{c < C/2}  Wq'[c][k] rel_h[r][c]              (ref :322: rel_h on the first half of k)
//   row 10 <= r < 20  mq[r][k] = sum_{c >= C/2} Wq'[c][k] rel_w[r-10][c-C/2]       (ref :323)
//   row 32 + n        mq[32+n][k] = sum_c Wq'[c][k] Wk'[c][n]
// Sums in fp64, one fp16 rounding at the end.
__global__ void mq_kernel(const float* __restrict__ qk, const float* __restrict__ relh, const float* __restrict__ relw,
                          __half* __restrict__ out, int C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (32 + C) * C) return;
    const int n = i / C, k = i - n * C, hc = C / 2;
    double acc = 0.0;
    if (n < 10) {
        for (int c = 0; c < hc; ++c) acc += (double)qk[(long)c * C + k] * (double)relh[n * hc + c];
    } else if (n < 20) {
        for (int c = hc; c < C; ++c) acc += (double)qk[(long)c * C + k] * (double)relw[(n - 10) * hc + (c - hc)];
    } else if (n >= 32) {
        const float* wk = qk + (long)C * C;
        for (int c = 0; c < C; ++c) acc += (double)qk[(long)c * C + k] * (double)wk[(long)c * C + (n - 32)];
    }
    out[i] = __float2half_rn((float)acc);
}

static int run_pack(int mode, const float* src, const float* src2, void* dst, int n, int p0, int p1, float scale,
                    cudaStream_t s) {
    pack_kernel<<<cdiv(n, 256), 256, 0, s>>>(mode, src, src2, dst, n, p0, p1, scale);
    M2T_LAUNCH_CHECK("pack_kernel");
    return M2T_OK;
}

int pack_weights_impl(const PackedLayout& L, const float* const* P, int n_params, uint8_t* packed, cudaStream_t s) {
    const int want = 6 + 14 * L.n_blocks + (L.scale == 4 ? 5 : 3);
    if (n_params != want) {
        set_error("pack_weights: got %d tensors, the x%d state_dict with %d blocks has %d", n_params, L.scale,
                  L.n_blocks, want);
        return M2T_E_ARG;
    }
    for (int i = 4; i < n_params; ++i)
        if (P[i] == nullptr) { set_error("pack_weights: tensor %d is null", i); return M2T_E_ARG; }
    // P[0..3] = sub_mean / add_mean: present in checkpoints, never used by forward (ref :30-31, :58-76)
    M2T_TRY(run_pack(PK_HEAD_W, P[4], nullptr, packed + L.head_w, 27 * NF, 0, 0, 1.f, s));
    M2T_TRY(run_pack(PK_COPY_F32, P[5], nullptr, packed + L.head_b, NF, 0, 0, 1.f, s));
    for (int i = 0; i < L.n_blocks; ++i) {
        const int base = 6 + 14 * i;
        for (int a = 0; a < 4; ++a) {
            const int C = branch_ch(a), hc = C / 2;
            const float* relh = P[base + 3 * a];
            const float* relw = P[base + 3 * a + 1];
            const float* wq = P[base + 3 * a + 2];
            const AttnW& A = L.blk[i].attn[a];
            const float qscale = 1.0f / sqrtf((float)C);    // ref :311 q * head_ch^-0.5 ; C in {16,64,256}
            M2T_TRY(run_pack(PK_CVT_F16, wq, nullptr, packed + A.wqkv, 3 * C * C, C * C, 0, qscale, s));
            {   // Haar-folded copy (stream order makes the shared fp32 scratch safe to reuse)
                float* tmp = reinterpret_cast<float*>(packed + L.fold_scratch);
                fold_rows_kernel<<<cdiv(3 * C * C, 256), 256, 0, s>>>(wq, tmp, C, branch_level(a));
                M2T_LAUNCH_CHECK("fold_rows_kernel");
                float* qk32 = reinterpret_cast<float*>(packed + L.fold_scratch2);
                fold_cols_kernel<<<cdiv(3 * C * C, 256), 256, 0, s>>>(tmp, reinterpret_cast<__half*>(packed + A.wqkv_f), qk32, C,
                                                                     branch_level(a), qscale);
                M2T_LAUNCH_CHECK("fold_cols_kernel");
                mq_kernel<<<cdiv((32 + C) * C, 256), 256, 0, s>>>(qk32, relh, relw, reinterpret_cast<__half*>(packed + A.mq), C);
                M2T_LAUNCH_CHECK("mq_kernel");
            }
            M2T_TRY(run_pack(PK_COPY_F32, relh, nullptr, packed + A.relf, 10 * hc, 0, 0, 1.f, s));
            M2T_TRY(run_pack(PK_COPY_F32, relw, nullptr, packed + A.relf + (size_t)10 * hc * 4, 10 * hc, 0, 0, 1.f, s));
            M2T_TRY(run_pack(PK_RELX, relh, relw, packed + A.relx, 32 * C, C, 0, 1.f, s));
        }
        M2T_TRY(run_pack(PK_CONV3_W, P[base + 12], nullptr, packed + L.blk[i].ffw, 9 * NF * NF, NF, NF, 0.f, s));
        M2T_TRY(run_pack(PK_CONV3_W2, P[base + 12], nullptr, packed + L.blk[i].ffw2, 2 * 9 * NF * NF, 0, 0, 1.f, s));
        M2T_TRY(run_pack(PK_COPY_F32, P[base + 13], nullptr, packed + L.blk[i].ffb, NF, 0, 0, 1.f, s));
    }
    const int tb = 6 + 14 * L.n_blocks;
    const int r0 = L.scale == 4 ? 2 : L.scale, N0 = NF * r0 * r0;
    M2T_TRY(run_pack(PK_UP_W, P[tb], nullptr, packed + L.t0w, N0 * NF, r0 * r0, 0, 1.f, s));
    M2T_TRY(run_pack(PK_UP_B, P[tb + 1], nullptr, packed + L.t0b, N0, r0 * r0, 0, 1.f, s));
    if (L.scale == 4) {
        M2T_TRY(run_pack(PK_UP_W, P[tb + 2], nullptr, packed + L.t3w, 256 * NF, 4, 0, 1.f, s));
        M2T_TRY(run_pack(PK_UP_B, P[tb + 3], nullptr, packed + L.t3b, 256, 4, 0, 1.f, s));
        M2T_TRY(run_pack(PK_CONV3_W, P[tb + 4], nullptr, packed + L.tcw, 9 * 16 * NF, 3, 16, 2048.f, s));
    } else {
        M2T_TRY(run_pack(PK_CONV3_W, P[tb + 2], nullptr, packed + L.tcw, 9 * 16 * NF, 3, 16, 2048.f, s));
    }
    return M2T_OK;
}

}  // namespace m2t
