"""PSNR / SSIM of the reference's test loop on the device (ref test.py:103-116).

The reference converts SR and HR to YCbCr (utils.rgb_to_ycbcr, ref utils.py:119-146), keeps Y, shaves `scale` border
pixels, multiplies by 255 when rgb_range == 1, and calls utils.calc_psnr / utils.calc_ssim (ref utils.py:179-184,
:232-234), each ending in a `float(...)` host synchronisation.  `psnr_ssim` does all of that in one pass over the two
tensors and returns device scalars, so an evaluation loop only synchronises when it prints.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import M2TError

__all__ = ["psnr_ssim", "calc_psnr_ssim", "gmsd"]


@torch.no_grad()
def psnr_ssim(sr: torch.Tensor, hr: torch.Tensor, scale: int, rgb_range: float = 1.0):
    """sr, hr: fp32 CUDA [B, colors, H, W] (colors 3 or 1) as `model(lr)` and the loader return them.
    Returns (per_image [B, 2], batch [2]) fp32 CUDA tensors of (PSNR dB, SSIM); `batch` follows the reference's
    arithmetic on the whole batch tensor (one MSE over all images, mean SSIM)."""
    for t, n in ((sr, "sr"), (hr, "hr")):
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise M2TError(f"psnr_ssim: {n} must be a CUDA tensor; the B200 engine has no CPU path")
        if t.dtype != torch.float32 or t.dim() != 4:
            raise M2TError(f"psnr_ssim: {n} must be float32 [B,C,H,W], got {t.dtype} {tuple(t.shape)}")
    if sr.shape != hr.shape:
        raise M2TError(f"psnr_ssim: shapes differ: {tuple(sr.shape)} vs {tuple(hr.shape)}")      # ref test.py:94 asserts
    b, c, h, w = sr.shape
    if c not in (1, 3):
        raise M2TError(f"psnr_ssim: colors must be 1 or 3, got {c}")
    lib = _lib.load()
    with torch.cuda.device(sr.device):
        nbytes = int(lib.m2t_metrics_workspace_bytes(b, h, w, int(scale)))
        if nbytes == 0:
            raise M2TError(f"psnr_ssim: {h}x{w} shaved by {scale} is smaller than the 11x11 SSIM window")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=sr.device)
        out = torch.empty(2 * (b + 1), dtype=torch.float32, device=sr.device)
        s, r = sr.contiguous(), hr.contiguous()
        _lib.check(lib.m2t_eval_psnr_ssim(s.data_ptr(), r.data_ptr(), b, c, h, w, int(scale), float(rgb_range), out.data_ptr(),
                                          ws.data_ptr(), torch.cuda.current_stream(sr.device).cuda_stream),
                   "m2t_eval_psnr_ssim")
    return out[:2 * b].view(b, 2), out[2 * b:]


def calc_psnr_ssim(sr, hr, scale, rgb_range=1.0):
    """(psnr, ssim) as Python floats: what ref test.py:113-114 adds to its running sums (one synchronisation)."""
    _, batch = psnr_ssim(sr, hr, scale, rgb_range)
    p, s = batch.tolist()
    return p, s


@torch.no_grad()
def gmsd(x: torch.Tensor, y: torch.Tensor, data_range: float = 1.0) -> torch.Tensor:
    """piq.gmsd(x, y, data_range=data_range, reduction='none') of ref test.py:98-99 on the device: fp32 CUDA [B, C, H, W]
    (C = 1 or 3) -> fp32 CUDA [B].  piq is absent offline: the kernel follows its published algorithm (parity unpinned)."""
    for t, n in ((x, "x"), (y, "y")):
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise M2TError(f"gmsd: {n} must be a CUDA tensor; the B200 engine has no CPU path")
        if t.dtype != torch.float32 or t.dim() != 4:
            raise M2TError(f"gmsd: {n} must be float32 [B,C,H,W], got {t.dtype} {tuple(t.shape)}")
    if x.shape != y.shape or x.shape[1] not in (1, 3) or x.shape[2] < 2 or x.shape[3] < 2:
        raise M2TError(f"gmsd: need equal shapes [B, 1|3, H>=2, W>=2], got {tuple(x.shape)} and {tuple(y.shape)}")
    b, c, h, w = x.shape
    lib = _lib.load()
    with torch.cuda.device(x.device):
        ws = torch.empty(int(lib.m2t_gmsd_workspace_bytes(b, h, w)) + 256, dtype=torch.uint8, device=x.device)
        out = torch.empty(b, dtype=torch.float32, device=x.device)
        off = (-ws.data_ptr()) % 256
        _lib.check(lib.m2t_eval_gmsd(x.contiguous().data_ptr(), y.contiguous().data_ptr(), b, c, h, w, float(data_range), out.data_ptr(),
                                     ws.data_ptr() + off, torch.cuda.current_stream(x.device).cuda_stream), "m2t_eval_gmsd")
    return out
