// MedCLIP image tower (Swin-T + projection head): shared declarations.
#pragma once
#include "common.cuh"

namespace m2t {

// epilogues of launch_lin_umma (lin_umma.cu)
enum { LIN_BF16 = 0, LIN_GELU_BF16 = 1, LIN_ADD_F32 = 2, LIN_F32 = 3 };
int launch_lin_umma(int epi, const void* A, const void* W, const float* bias, void* out, int M, int N, int K,
                    cudaStream_t s);

// fused MLP (mlp_umma.cu): X fp32 [M][C] += W2 . gelu(W1 . A + b1) + b2, C = 96 or 192; the hidden tensor stays on the SM
int launch_mlp_umma(const void* A, const void* W1, const float* b1, const void* W2, const float* b2, float* X, int M, int C,
                    cudaStream_t s);
int read_lin_timing(long long* host128);      // clock64 stamps of the last LIN_BF16 launch (M2T_TIMING builds; zeros otherwise)

// window attention (clip_attn.cu): qkv bf16 [tokens][3C], out bf16 [tokens][C], rpb fp32 [heads][49][CL_RPB_PITCH]
int launch_clip_attn(const void* qkv, void* out, const float* rpb, int B, int h, int w, int C, int heads, int shift,
                     cudaStream_t s);
constexpr int CL_RPB_PITCH = 56;     // bias rows padded to the 7 n8 key tiles; the pad columns hold -1e30

constexpr int CL_IMG = 224, CL_PATCH = 4, CL_GRID = 56, CL_EMBED = 96, CL_WIN = 7, CL_WT = 49, CL_HD = 32;
constexpr int CL_STAGES = 4, CL_FEAT = 768, CL_PROJ = 512, CL_NPARAMS = 220;
constexpr float CL_LN_EPS = 1e-5f;

}  // namespace m2t
