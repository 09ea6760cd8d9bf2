/*
 * m2trans_b200 -- C ABI of the B200 (sm_100a) engine for the M2Trans forward path.
 *
 * The reference (eezkni/M2Trans) is pure Python: its "operator interface" for this
 * path is the nn.Module `M2Trans` of models/M2Trans_network.py, reached through the
 * plugin hook `utils.import_module('models.{}_network').create_model(args)`
 * (ref train.py:69-70, utils.py:175-176) or directly (ref test.py:66,90).  There is
 * no FFI in the reference; this header is the boundary a binding would sit on
 * (INTEGRATION.md shows the ctypes stub).  Every entry point cites the reference
 * code it replaces.
 *
 * Conventions
 *   - plain C types only; pointers named d_* are DEVICE pointers on the current
 *     CUDA device; `stream` is a cudaStream_t passed as void* (NULL = legacy stream)
 *   - every int function returns 0 (M2T_OK) or a negative M2T_E* code; the message
 *     of the last failure on the calling thread is m2t_last_error()
 *   - nothing here synchronises the device, allocates device memory, or falls back
 *     to the CPU; a device that is not sm_100 is M2T_E_DEVICE
 *   - re-entrant: a plan is immutable after creation; concurrent m2t_forward calls
 *     are allowed when they use different workspaces
 *
 * Internal tensor layout: NHWC; residual stream fp32; branch tensors
 * and GEMM operands fp16 with fp32 accumulation; weights pre-packed by
 * m2t_pack_weights.
 */
#ifndef M2TRANS_B200_H
#define M2TRANS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define M2T_OK             0
#define M2T_E_ARG         -1  /* bad argument (null pointer, bad shape)                  */
#define M2T_E_DEVICE      -2  /* no CUDA device, or the device is not sm_100             */
#define M2T_E_CUDA        -3  /* a CUDA runtime/driver call or a kernel launch failed    */
#define M2T_E_UNSUPPORTED -4  /* configuration outside what the reference constructs     */

/* kernel-variant bits for m2t_cfg.variant (0 = the default product path).  The SIMT
 * variants are CUDA-core kernels with the same operand precision; they exist as an
 * on-device cross-check for the tcgen05 kernels, not as a fallback: nothing selects
 * them automatically. */
#define M2T_VAR_DEFAULT   0u
#define M2T_VAR_SIMT_CONV (1u << 0)  /* ff 3x3 conv on CUDA cores instead of tcgen05   */
#define M2T_VAR_SIMT_QKV  (1u << 1)  /* qkv 1x1 conv on CUDA cores instead of tcgen05  */
#define M2T_VAR_SIMT_TAIL (1u << 2)  /* tail on CUDA cores instead of tcgen05          */
#define M2T_VAR_SIMT_ATTN (1u << 3)  /* attention on CUDA cores instead of tcgen05     */
#define M2T_VAR_SIMT_ALL  0xFu
#define M2T_VAR_UNFUSED_TAIL (1u << 4) /* x2/x4: tail_up + border + tail_out instead of the fused last stage */
/* Precise mode.  (1) The fp16 rounding residual of t_k travels beside t_k and is added back where
 * y_k = attention + t_k is formed (two more 32-byte segments per pixel and branch).  (2) The ff conv uses
 * split-precision weights: fp16 weight + fp16 residual * 2^11 in extra accumulator columns, 32 output channels per CTA.
 * Together they remove the two largest error terms of the fp16 operand format (emulated at x3 on a speckle frame:
 * max-abs 2.7e-3 -> 1.7e-3) for ~10 % of the forward time.  Default: on for x2 / x3, whose single-stage tail leaves
 * less margin under the 2e-3 bar, off for x4.  These bits force it either way. */
#define M2T_VAR_SPLIT_QKV16  (1u << 7) /* branch 1: separate qkv kernel + attention instead of the fused attn16_qkv kernel */
#define M2T_VAR_SPLIT_QKV    (1u << 8) /* branches 2-4: separate qkv GEMM + attention kernels (QKV through HBM) instead of attn_z */
#define M2T_VAR_TILE_TAIL    (1u << 9) /* x2/x4: the tiled fused tail (tail_fused.cu) instead of the strip-marching one (tail_strip.cu) */
#define M2T_VAR_AZ_PAIRED    (1u << 10) /* attn_z: two windows per CTA even for small inputs (default: one window per CTA when all fit in one wave) */
#define M2T_VAR_W2_PAIR      (1u << 11) /* precise mode: ff conv as a CTA pair per two tiles (conv_pair.cu, tcgen05 cta_group::2) instead of two 32-channel CTAs per tile; bit-identical, not faster */
#define M2T_VAR_PRECISE_ON   (1u << 5)
#define M2T_VAR_PRECISE_OFF  (1u << 6)

typedef struct m2t_plan m2t_plan;

/* The hyper-parameters M2Trans.__init__ reads (ref M2Trans_network.py:21-25,34) plus
 * the input geometry of one forward call (ref :60). */
typedef struct m2t_cfg {
    int32_t  scale;      /* args.scale: 2, 3 or 4                                    */
    int32_t  n_feats;    /* args.n_feats: must be 64 (ref configs/M2Trans_x*.yml)    */
    int32_t  n_blocks;   /* args.n_blocks: number of CFTM blocks, 1..64 (ref: 8)     */
    int32_t  colors;     /* args.colors: must be 3                                   */
    int32_t  batch;      /* B >= 1                                                   */
    int32_t  height;     /* LR H; reflect padding to a multiple of 32 must be defined */
    int32_t  width;      /* LR W                                                     */
    uint32_t variant;    /* M2T_VAR_* bits                                           */
    float    rgb_range;  /* args.rgb_range: upper clamp (ref :74)                    */
} m2t_cfg;

/* ---- device / library ------------------------------------------------------------ */
int         m2t_query_device(int* sm_major, int* sm_minor, int* sm_count);
const char* m2t_last_error(void);
const char* m2t_version(void);

/* ---- weights ----------------------------------------------------------------------
 * Replaces nn.Module parameter storage + load_state_dict (ref :88-112, test.py:70).
 * `d_params` is a HOST array of n_params DEVICE pointers to the fp32 tensors of
 * M2Trans.state_dict() in registration order (123 tensors for
 * x4, 121 for x2/x3).  sub_mean/add_mean (entries 0..3) are accepted and ignored
 * because the reference never calls them in forward (ref :58-76).  `d_packed`
 * (m2t_packed_weight_bytes, 256-byte aligned) receives the engine's packed copy. */
int    m2t_num_params(int scale, int n_blocks);
size_t m2t_packed_weight_bytes(int scale, int n_blocks);
int    m2t_pack_weights(int scale, int n_blocks, const float* const* d_params, int n_params,
                        void* d_packed, void* stream);
/* byte offset of a packed tensor, for tests: name is e.g. "head_w", "head_b",
 * "body.3.attn2.wqkv", "body.3.attn2.relf", "body.3.attn2.relx", "body.3.ffw", "body.3.ffb",
 * "t0w", "t0b", "t3w", "t3b", "tcw".  Returns (size_t)-1 for an unknown name. */
size_t m2t_packed_offset(int scale, int n_blocks, const char* name);

/* ---- plan ------------------------------------------------------------------------- */
int    m2t_plan_create(const m2t_cfg* cfg, m2t_plan** out);
void   m2t_plan_destroy(m2t_plan* plan);
size_t m2t_workspace_bytes(const m2t_plan* plan);
/* padded frame size (multiples of 32, ref :78-86) */
int    m2t_plan_padded(const m2t_plan* plan, int* Hp, int* Wp);
/* kernel launches one m2t_forward issues (bench.py's gpu_launches) */
int    m2t_plan_num_launches(const m2t_plan* plan);
/* byte offset of an internal NHWC tensor inside the workspace, for tests:
 * "res" (head output, fp32 [B,Hp,Wp,64]), "x" (residual stream after the last CFTM,
 * fp32), "y" (cat[y1..y4] of the last CFTM, fp16).  (size_t)-1 for an unknown name. */
size_t m2t_workspace_offset(const m2t_plan* plan, const char* name);

/* ---- the hot path -------------------------------------------------------------------
 * Replaces M2Trans.forward (ref :58-76): check_image_size reflect pad (:78-86), head
 * conv (:34,:63), n_blocks x CFTM (:132-164: InstanceNorm :135, four chained TBlocks
 * :290-340 over Haar DWT/IWT pyramids :198-237, 3x3 feed_forward + residual :164),
 * global residual (:70), tail (:40-56), clamp (:74) and crop (:76).
 *   d_x  [B,3,H,W]      fp32 NCHW contiguous, values in [0, rgb_range]
 *   d_y  [B,3,H*s,W*s]  fp32 NCHW contiguous
 *   d_workspace         m2t_workspace_bytes(plan) bytes, 256-byte aligned          */
int m2t_forward(const m2t_plan* plan, const void* d_packed, const float* d_x, float* d_y,
                void* d_workspace, void* stream);
/* The same forward in up to three parts (bit mask): M2T_PHASE_HEAD reads d_x (frame pad + head conv), M2T_PHASE_BODY
 * runs the CFTM blocks on the workspace only, M2T_PHASE_TAIL writes d_y.  The parts of one forward must run in this
 * order on one stream with the same workspace.  Lets a host capture the pointer-independent BODY in a CUDA graph and
 * launch HEAD / TAIL directly on the caller's tensors (no staging copies).  d_x / d_y may be NULL when their phase is
 * not selected. */
#define M2T_PHASE_HEAD 1u
#define M2T_PHASE_BODY 2u
#define M2T_PHASE_TAIL 4u
#define M2T_PHASE_ALL  7u
int m2t_forward_phases(const m2t_plan* plan, const void* d_packed, const float* d_x, float* d_y,
                       void* d_workspace, void* stream, uint32_t phases);

/* ---- per-stage entry points (unit tests, ncu) ------------------------------------------
 * The same kernels m2t_forward launches, on the engine's native NHWC tensors.
 * Geometry: B images of Hp x Wp padded LR pixels (multiples of 32). */

/* head: reflect pad + 3->64 conv + bias (ref :34,:63,:78-86); also accumulates the
 * InstanceNorm sums of the first CFTM.  d_stats: double [B][64][2], zeroed by caller. */
int m2t_stage_head(const float* d_x, const float* d_head_w, const float* d_head_b, float* d_res,
                   double* d_stats, int B, int H, int W, void* stream);
/* (sum, sumsq) -> (mean, rstd): nn.InstanceNorm2d statistics (ref :127,:135) */
int m2t_stage_stats_finalize(const double* d_stats, float* d_munorm /* float2 [B][64] */, int B,
                             int npix, void* stream);
/* branch glue (ref :135-161): prep writes Z = DWT^L((n_k + y_{k-1})/2) as fp16
 * [B,Hp>>L,Wp>>L,16*4^L]; post writes y_k = IWT^L(O) + t_k into channels 16k.. of Y. */
int m2t_stage_branch_prep(int branch, const float* d_X, const float* d_munorm, const void* d_Y,
                          void* d_Z, int B, int Hp, int Wp, void* stream);
int m2t_stage_branch_post(int branch, const void* d_O, const float* d_X, const float* d_munorm,
                          void* d_Y, int B, int Hp, int Wp, void* stream);
/* TBlock qkv 1x1 conv (ref :307): QKV[m][3C] = Z[m][C] . Wqkv^T, fp16 */
int m2t_stage_qkv(uint32_t variant, const void* d_Z, const void* d_wqkv, void* d_QKV, int M, int C,
                  void* stream);
/* TBlock attention core (ref :310-332) on QKV [B,h,w,3C] -> O [B,h,w,C], fp16 */
int m2t_stage_attn(uint32_t variant, int C, const void* d_QKV, const float* d_relf,
                   const void* d_relx, void* d_O, int B, int h, int w, void* stream);
/* One whole CFTM branch k = branch+1 in {2,3,4} (ref :143-161 with :307-332 inside) as ONE kernel: qkv conv, halo
 * attention and the branch glue, with the contractions re-associated so that q, k, v are never formed (attn_z.cu).
 *   d_T      t_k, fp16 space-to-depth [B, h, w, C]  (C = 64: level 1, h = Hp/2;  C = 256: level 2, h = Hp/4)
 *   d_mq     fp16 [32+C][C]  (m2t_packed_offset "body.i.attnK.mq"),  d_wv  fp16 [C][C] (the v rows of "wqkv_f")
 *   d_Y      fp16 NHWC [B, Hp, Wp, 64]: channels 16*branch.. receive y_k = attention + t_k
 *   d_Tnext  fp16 level-2 space-to-depth [B, Hp/4, Wp/4, 256] holding n_{k+1}/2; completed in place to t_{k+1}; or NULL */
int m2t_stage_attn_z(int C, const void* d_T, const void* d_mq, const void* d_wv, void* d_Y, void* d_Tnext,
                     int branch, int B, int h, int w, void* stream);
/* CFTM.feed_forward + residual (ref :124-126,:164) + next block's InstanceNorm sums */
int m2t_stage_ffconv(uint32_t variant, const void* d_Y, const void* d_ffw, const float* d_ffb,
                     const float* d_Xin, float* d_Xout, double* d_stats, int B, int Hp, int Wp,
                     void* stream);
/* tail (ref :40-56,:72-76) on B images; d_XR is fp16 NHWC [B,Hp,Wp,64] = res + x (ref :70; in
 * m2t_forward the last ff conv's epilogue writes it); output images b0.. of d_y [*,3,H*s,W*s];
 * d_scratch holds m2t_tail_scratch_bytes(scale, B, Hp, Wp) bytes; d_packed is the blob of
 * m2t_pack_weights(scale, n_blocks, ...). */
size_t m2t_tail_scratch_bytes(int scale, int B, int Hp, int Wp);
int m2t_stage_tail(uint32_t variant, int scale, int n_blocks, const void* d_packed, const void* d_XR,
                   float* d_y, int B, int b0, int H, int W, float rgb_range, void* d_scratch,
                   void* stream);

/* ---- rlutrans.TransBlock -------------------------------------------------
 * Replaces util/rlutrans.py TransBlock.forward (ref util/rlutrans.py:82-87; EffAttention.forward
 * :46-66; Mlp.forward :20-27):  y = x1 + fc2(ReLU(fc1(LN2(x1)))),  x1 = x + proj(attn(qkv(reduce(LN1(x))))),
 * attention with 8 heads and softmax restricted to chunks of N/16 consecutive tokens.
 * d_x, d_y: fp32 [B][N][dim] (d_y may not alias d_x); d_params: HOST array of the 12 DEVICE fp32
 * tensors of TransBlock.state_dict() in registration order (atten.reduce.weight, atten.qkv.weight,
 * atten.proj.weight, atten.proj.bias, norm1.weight, norm1.bias, mlp.fc1.weight, mlp.fc1.bias,
 * mlp.fc2.weight, mlp.fc2.bias, norm2.weight, norm2.bias).  dim must be 64 and num_heads 8 (the
 * constructor defaults; M2T_E_UNSUPPORTED otherwise); N < 16 is M2T_E_ARG (the reference raises).
 * d_workspace: m2t_transblock_workspace_bytes(B, N, dim) bytes, 16-byte aligned. */
size_t m2t_transblock_workspace_bytes(int B, int N, int dim);
int m2t_transblock_forward(const float* d_x, float* d_y, const float* const* d_params, int n_params,
                           int B, int N, int dim, int num_heads, void* d_workspace, void* stream);
/* The two sub-modules of util/rlutrans.py on their own (the reference only calls them from TransBlock.forward, but they
 * are public classes): EffAttention.forward (ref :47-66; d_params = reduce.weight, qkv.weight, proj.weight, proj.bias;
 * workspace as for the block) and Mlp.forward (ref :21-27; d_params = fc1.weight, fc1.bias, fc2.weight, fc2.bias). */
int m2t_rlutrans_attention(const float* d_x, float* d_y, const float* const* d_params, int n_params, int B, int N, int dim,
                           int num_heads, void* d_workspace, void* stream);
int m2t_rlutrans_mlp(const float* d_x, float* d_y, const float* const* d_params, int n_params, long tokens, int dim,
                     int hidden, void* stream);

/* ---- MedCLIP image-embedding pass ------------------------------------------
 * Replaces the image side of SemanticLoss.__call__ (ref losses.py:53-54 bicubic resize to 224x224 with
 * align_corners=True, :68-69 medmodel.encode_image, :71-72 L2 normalise, :76-77 dot with the normalised text
 * feature).  encode_image is third-party (medclip's MedCLIPVisionModelViT = Swin-T
 * 'microsoft/swin-tiny-patch4-window7-224' -> pooler_output [B,768] -> Linear(768,512,bias=False)); the
 * architecture follows transformers/models/swin/modeling_swin.py.  bf16 tensor-core GEMMs, fp32 residual stream.
 * Parameters: HOST array of m2t_clip_param_count() = 220 DEVICE fp32 tensors: SwinModel.state_dict() order without
 * the relative_position_index buffers (patch projection weight/bias, embeddings norm; per block layernorm_before,
 * relative_position_bias_table, query/key/value weight+bias, attention.output.dense, layernorm_after,
 * intermediate.dense, output.dense; per stage downsample.reduction.weight, downsample.norm; final layernorm), then
 * projection_head.weight [512][768].
 * m2t_clip_encode_image: d_img fp32 [B][3][H][W] (values as the SR network returns them, no mean/std
 * normalisation: the reference applies none); d_embed fp32 [B][512] L2-normalised; d_text fp32 [512] (any norm)
 * and d_logits fp32 [B] are both given or both NULL; d_workspace m2t_clip_workspace_bytes(B) bytes, 256-byte
 * aligned; d_packed m2t_clip_packed_bytes() bytes, 256-byte aligned.  Asynchronous on `stream`; no allocation, no
 * synchronisation (capturable in a CUDA graph). */
int m2t_clip_param_count(void);
size_t m2t_clip_packed_bytes(void);
size_t m2t_clip_workspace_bytes(int B);
int m2t_clip_pack_weights(const float* const* d_params, int n_params, void* d_packed, void* stream);
int m2t_clip_encode_image(const void* d_packed, const float* d_img, int B, int H, int W, float* d_embed,
                          const float* d_text, float* d_logits, void* d_workspace, void* stream);
/* One Linear of the tower on its own (stage-level tests): out = epilogue(A W^T + bias), A bf16 [M][K], W bf16 [N][K]
 * (nn.Linear layout), bias fp32 [N] or NULL.  epilogue 0: bf16 [M][N]; 1: GELU, bf16; 2: fp32 [M][N] += (the residual
 * form); 3: fp32 [M][N] =.  K a multiple of 8, N a multiple of 32; neither needs to be a multiple of the 128 x 128 x 64
 * tile. */
int m2t_clip_stage_linear(int epilogue, const void* d_a, const void* d_w, const float* d_bias, void* d_out,
                          int M, int N, int K, void* stream);
/* The fused MLP of stages 1-2 on its own: d_x fp32 [M][C] += W2 . gelu(W1 . A + b1) + b2 with A bf16 [M][C], W1 bf16 [4C][C],
 * W2 bf16 [C][4C], biases fp32; C = 96 or 192 (M2T_E_UNSUPPORTED otherwise).  The hidden [M][4C] tensor is rounded to bf16 as
 * in the two-launch form but never written to memory. */
int m2t_clip_stage_mlp(const void* d_a, const void* d_w1, const float* d_b1, const void* d_w2, const float* d_b2,
                       float* d_x, int M, int C, void* stream);
/* The other kernels of the tower on their own (stage-level tests; the end-to-end embedding of a random-weight tower
 * only resolves errors above ~1e-3, these resolve an indexing slip exactly).
 * resize: d_img fp32 [B][3][H][W] -> d_rows bf16 [B*56*56][48], the 4x4 patch rows (column c*16 + ky*4 + kx) of the
 *   bicubic, align_corners=True, 224x224 image (ref losses.py:53).
 * layernorm: d_x fp32 tokens [B*h*w][C] -> bf16; merge = 0: LayerNorm(C) per token; merge = 1: the 2x2 patch-merging
 *   gather, [B*h*w/4][4C], LayerNorm(4C) (modeling_swin.py:333-343).
 * attention: d_qkv bf16 [B*h*w][3C] (q | k | v, head-major 32-wide slices) -> d_out bf16 [B*h*w][C]; d_bias fp32
 *   [heads][49][56]: the gathered relative-position table with rows padded to 56 and -1e30 in columns 49..55 (the
 *   kernel's key tiles are 7 x 8 wide; the pad masks the 7 phantom keys); shift 0 or 3 (cyclic shift + region mask,
 *   :556-582, :615); heads a multiple of 3. */
int m2t_clip_stage_resize(const float* d_img, void* d_rows, int B, int H, int W, void* stream);
/* Development aid: 128 clock64 stamps of CTA 0 of the last epilogue-0 Linear launch (library built with M2T_TIMING=1;
 * zeros otherwise); layout in csrc/lin_umma.cu. */
int m2t_debug_lin_timing(long long* host128);
int m2t_clip_stage_layernorm(const float* d_x, void* d_out, const float* d_gamma, const float* d_beta, int B,
                             int h, int w, int C, int merge, void* stream);
int m2t_clip_stage_attention(const void* d_qkv, void* d_out, const float* d_bias, int B, int h, int w, int C,
                             int heads, int shift, void* stream);

/* ---- evaluation metrics of the test loop -----------------------------------
 * Replaces ref test.py:103-116 for one batch: Y channel of SR and HR (ref utils.py:119-146; colors == 1 skips it),
 * `shave` = args.scale border pixels cropped (test.py:109-110), x 255 when rgb_range == 1 (:111-112),
 * utils.calc_psnr (utils.py:179-184) and utils.calc_ssim = pytorch_msssim.ssim(size_average=True) with its defaults
 * (utils.py:232-234: 11-tap Gaussian, sigma 1.5, valid region, data_range 255).
 * d_sr, d_hr: fp32 [B][colors][H][W]; d_out: fp32 [2 (B + 1)]: {psnr_b, ssim_b} per image, then {psnr, ssim} of the whole
 * batch tensor as the reference computes them (for B = 1, the loader's batch size, the two coincide).
 * d_workspace: m2t_metrics_workspace_bytes(B, H, W, shave) bytes, 16-byte aligned.  No host synchronisation. */
size_t m2t_metrics_workspace_bytes(int B, int H, int W, int shave);
int m2t_eval_psnr_ssim(const float* d_sr, const float* d_hr, int B, int colors, int H, int W, int shave,
                       float rgb_range, float* d_out, void* d_workspace, void* stream);
/* GMSD of the test loop (ref test.py:98-99: piq.gmsd(hr, sr, data_range=1., reduction='none')); piq is third-party and
 * absent offline, so this follows its published algorithm (csrc/gmsd.cu): d_out[B] fp32, one value per image.  d_x, d_y:
 * fp32 [B, colors, H, W] device tensors (colors 1 or 3), H, W >= 2. */
size_t m2t_gmsd_workspace_bytes(int B, int H, int W);
int m2t_eval_gmsd(const float* d_x, const float* d_y, int B, int colors, int H, int W, float data_range, float* d_out,
                  void* d_workspace, void* stream);

/* ---- loader conversion (device side) -----------------------------------------
 * Replaces ndarray2tensor(img) / 255. of the benchmark loader (ref datas/benchmark.py:66-69, utils.py:237-240) on the
 * device: d_src uint8 [B][H][W][colors] (the image arrays the loader keeps in RAM) -> d_dst fp32 [B][colors][H][W] =
 * float(u8) / denom (IEEE division: bit-identical to torch).  colors 1 or 3.  3 bytes per pixel cross PCIe instead of 12. */
int m2t_u8hwc_to_f32chw(const void* d_src, float* d_dst, int B, int H, int W, int colors, float denom, void* stream);
/* The inverse, for writing SR batches out as images: fp32 CHW -> uint8 HWC, dst = clamp(rint(src * scale), 0, 255)
 * (round half to even, like torch.round).  4x fewer bytes over PCIe than the fp32 batch. */
int m2t_f32chw_to_u8hwc(const float* d_src, void* d_dst, int B, int H, int W, int colors, float scale, void* stream);

/* ---- hardware probes (development aids; tests/test_probes.py) ---------------------------
 * m2t_probe_umma: copies two raw shared-memory images (A, B operands), issues k_steps
 * tcgen05.mma (kind::f16, cta_group::1, M=128) with the given 64-bit shared-memory
 * descriptors (the 14-bit start-address field is added to the image base; each k step
 * adds a_step/b_step to the descriptor's low word) and instruction descriptor, and
 * dumps the fp32 accumulator [128 lanes][n_cols]. */
int m2t_probe_umma(const void* d_a_image, uint32_t a_bytes, const void* d_b_image, uint32_t b_bytes,
                   uint64_t a_desc, uint64_t b_desc, uint32_t a_step, uint32_t b_step, int k_steps,
                   uint32_t idesc, int n_cols, float* d_out, void* stream);
/* m2t_debug_profile_forward: runs m2t_forward once eagerly with a CUDA event after every launch (programmatic
 * dependent launch off, so launches do not overlap), synchronises the stream and writes a per-kernel table
 * (time, share, launches, mangled name) into `text` (HOST buffer of `cap` bytes).  Development aid: unlike ncu it
 * sees the kernels with the caches as the previous kernel left them. */
int m2t_debug_profile_forward(const m2t_plan* plan, const void* d_packed, const float* d_x, float* d_y,
                              void* d_workspace, void* stream, char* text, size_t cap);
/* m2t_debug_attn_timing: host64 holds 7 x 64 values.  Record 6: CTA 0 of the last 256-channel qkv GEMM launch ([0] entry,
 * [1] prologue done, [2] weight slab landed, then per tile t < 6 at [8+8t+0..4]: MMA warp before its waits, first A
 * block landed, MMAs issued; epilogue: accumulator ready, tile stored).  Record 5: CTA 0 of the last tcgen05 ff-conv launch, per tile i < 8:
 * epilogue [8i+0..4] tile start, residual loads issued, accumulator ready, staged, stored; MMA warp [8i+5..7] before
 * / after the accumulator-free wait and after the halo-tile wait.  Record 4: CTA 0 of the last fused-tail launch, per tile i < 8 at
 * [8i+0..5]: tile start, GELU epilogue done, U tile published, conv accumulators ready, planes written, gather done.
 * Records 0..3: clock64 stamps of CTA 0 of the last tcgen05 attention launch of each branch (per record:
 * [6] kernel entry, [7] prologue done, then per pair i < 6 at [8i+0..5]: before S wait, S ready,
 * softmax done, glue operands issued, O ready, epilogue done).  All zero unless the library was
 * built with M2T_TIMING=1 (development builds only). */
int m2t_debug_attn_timing(long long* host64);
/* attn_z.cu stamps: 3 x 64 values (branches 2..4); CTA 0, epilogue thread 0, pairs it < 6 at [10 it + 0..8]: pair start, A ready,
 * AQ written, S ready, P written, PZ ready, PZ written, O ready, pair done; [60], [61] kernel entry / prologue done.  Zeros
 * unless the library was built with M2T_TIMING=1. */
int m2t_debug_az_timing(long long* host192);
/* m2t_probe_tma: builds a tiled tensor map over d_tensor (dims / strides_bytes / box given
 * innermost first, HOST arrays of `rank` entries; swizzle 0 none, 1 32B, 2 64B, 3 128B), loads
 * one box at `coords` (may be negative / out of bounds: zero fill) into 1024-byte-aligned
 * shared memory and copies the raw shared-memory image (out_bytes = box bytes) to d_out. */
int m2t_probe_tma(const void* d_tensor, int elem_bytes, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box, int swizzle,
                  const int32_t* coords, void* d_out, uint32_t out_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* M2TRANS_B200_H */
