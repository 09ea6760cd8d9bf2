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

__all__ = ["psnr_ssim", "calc_psnr_ssim"]


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
