"""Drop-in mirror of the reference's `util/rlutrans.py`.

Same public surface as the reference file (ref util/rlutrans.py): Mlp :11, EffAttention :30,
TransBlock :70, with identical constructor arguments, parameter names, shapes, registration order
and default initialisation (nn.Linear / nn.LayerNorm defaults), so a TransBlock state_dict moves
between the two unchanged.

`TransBlock.forward` is the engine's fused three-kernel path (include/m2trans_b200.h,
m2t_transblock_forward): CUDA fp32 in, CUDA fp32 out, no CPU or PyTorch fallback.  The reference only
ever calls Mlp and EffAttention from TransBlock.forward (ref :85-86), where their arithmetic lives inside
the fused kernels; called on their own they run the same kernels in a standalone mode
(m2t_rlutrans_mlp / m2t_rlutrans_attention).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from ._lib import M2TError
from ._params import state_tensors

__all__ = ["Mlp", "EffAttention", "TransBlock"]


def _check_tokens(x, what, dim):
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise M2TError(f"{what}: expected a CUDA tensor; the B200 engine has no CPU path")
    if x.dtype != torch.float32:
        raise M2TError(f"{what}: expected float32 (the reference's dtype), got {x.dtype}")
    if x.dim() < 2 or x.shape[-1] != dim:
        raise M2TError(f"{what}: expected [..., {dim}], got {tuple(x.shape)}")


def _device_params(module, x, what):
    params = [p.detach() for p in state_tensors(module)]          # also right on DataParallel replicas
    for p in params:
        if p.device != x.device or p.dtype != torch.float32:
            raise M2TError(f"{what}: parameters must be float32 on the input's device")
    return [p.contiguous() for p in params]


class Mlp(nn.Module):
    """dim -> dim//4 -> dim with ReLU (ref :11-27).  Parameter container; see TransBlock.forward."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.ReLU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features // 4
        if act_layer is not nn.ReLU or drop != 0.:
            raise M2TError("Mlp: the engine implements ReLU without dropout (the TransBlock defaults, ref :72-74)")
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    @torch.no_grad()
    def forward(self, x):
        """x [..., 64] fp32 CUDA -> same shape: fc2(ReLU(fc1(x))) (ref :21-27)."""
        _check_tokens(x, "Mlp.forward", self.fc1.in_features)
        if (self.fc1.in_features, self.fc1.out_features, self.fc2.out_features) != (64, 16, 64):
            raise M2TError("Mlp.forward: the engine implements 64 -> 16 -> 64 (the TransBlock defaults, ref :81)")
        lib = _lib.load()
        params = _device_params(self, x, "Mlp.forward")
        with torch.cuda.device(x.device):
            xc = x.contiguous()
            y = torch.empty_like(xc)
            ptrs = (C.c_void_p * 4)(*[p.data_ptr() for p in params])
            _lib.check(lib.m2t_rlutrans_mlp(xc.data_ptr(), y.data_ptr(), ptrs, 4, xc.numel() // 64, 64, 16,
                                            torch.cuda.current_stream(x.device).cuda_stream), "m2t_rlutrans_mlp")
        return y


class EffAttention(nn.Module):
    """reduce -> qkv -> 8-head attention inside chunks of N//16 tokens -> proj (ref :30-67).
    Parameter container; see TransBlock.forward."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        if qkv_bias or qk_scale is not None or attn_drop != 0. or proj_drop != 0.:
            raise M2TError("EffAttention: the engine implements qkv_bias=False, qk_scale=None, no dropout "
                           "(what TransBlock passes, ref :77-78)")
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = head_dim ** -0.5
        self.reduce = nn.Linear(dim, dim, bias=qkv_bias)
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        self.attn_drop = nn.Dropout(attn_drop)

    @torch.no_grad()
    def forward(self, x):
        """x [B, N, 64] fp32 CUDA, N >= 16 -> same shape: proj(chunked attention(qkv(reduce(x)))) (ref :47-66)."""
        _check_tokens(x, "EffAttention.forward", self.reduce.in_features)
        if x.dim() != 3 or self.reduce.in_features != 64 or self.num_heads != 8:
            raise M2TError(f"EffAttention.forward: expected [B, N, 64] with 8 heads, got {tuple(x.shape)}")
        b, n, _ = x.shape
        lib = _lib.load()
        params = _device_params(self, x, "EffAttention.forward")
        with torch.cuda.device(x.device):
            xc = x.contiguous()
            y = torch.empty_like(xc)
            ws = torch.empty(max(int(lib.m2t_transblock_workspace_bytes(b, n, 64)), 16), dtype=torch.uint8, device=x.device)
            ptrs = (C.c_void_p * 4)(*[p.data_ptr() for p in params])
            _lib.check(lib.m2t_rlutrans_attention(xc.data_ptr(), y.data_ptr(), ptrs, 4, b, n, 64, self.num_heads, ws.data_ptr(),
                                                  torch.cuda.current_stream(x.device).cuda_stream), "m2t_rlutrans_attention")
        return y


class TransBlock(nn.Module):
    """x + atten(norm1(x)), then x + mlp(norm2(x)) (ref :70-87)."""

    def __init__(self, n_feat=64, dim=64, num_heads=8, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0.,
                 attn_drop=0., drop_path=0., act_layer=nn.ReLU, norm_layer=nn.LayerNorm):
        super().__init__()
        if dim != 64 or num_heads != 8:
            raise M2TError(f"TransBlock: dim {dim} / num_heads {num_heads}: the engine is built for the defaults 64 / 8")
        self.dim = dim
        # the reference ignores its own qkv_bias / qk_scale / attn_drop arguments here (ref :77-78)
        self.atten = EffAttention(self.dim, num_heads=num_heads, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.)
        self.norm1 = nn.LayerNorm(self.dim)
        self.mlp = Mlp(in_features=dim, hidden_features=dim // 4, act_layer=act_layer, drop=drop)
        self.norm2 = nn.LayerNorm(self.dim)

    @torch.no_grad()
    def forward(self, x):
        """x [B, N, 64] fp32 CUDA, N >= 16 -> same shape (ref :82-87)."""
        if not isinstance(x, torch.Tensor) or not x.is_cuda:
            raise M2TError("TransBlock.forward: expected a CUDA tensor; the B200 engine has no CPU path")
        if x.dtype != torch.float32:
            raise M2TError(f"TransBlock.forward: expected float32 (the reference's dtype), got {x.dtype}")
        if x.dim() != 3 or x.shape[2] != self.dim:
            raise M2TError(f"TransBlock.forward: expected [B, N, {self.dim}], got {tuple(x.shape)}")
        b, n, _ = x.shape
        lib = _lib.load()
        params = [p.detach() for p in state_tensors(self)]          # also right on DataParallel replicas
        for p in params:
            if p.device != x.device or p.dtype != torch.float32:
                raise M2TError("TransBlock.forward: parameters must be float32 on the input's device")
        params = [p.contiguous() for p in params]
        with torch.cuda.device(x.device):
            xc = x.contiguous()
            y = torch.empty_like(xc)
            ws = torch.empty(max(int(lib.m2t_transblock_workspace_bytes(b, n, self.dim)), 16), dtype=torch.uint8,
                             device=x.device)
            ptrs = (C.c_void_p * len(params))(*[p.data_ptr() for p in params])
            _lib.check(lib.m2t_transblock_forward(xc.data_ptr(), y.data_ptr(), ptrs, len(params), b, n, self.dim,
                                                  self.atten.num_heads, ws.data_ptr(),
                                                  torch.cuda.current_stream(x.device).cuda_stream),
                       "m2t_transblock_forward")
        return y
