"""PSNR / SSIM of the reference's test loop on the device (ref test.py:103-116).

The reference converts SR and HR to YCbCr (utils.rgb_to_ycbcr, ref utils.py:119-146), keeps Y, shaves `scale` border
pixels, multiplies by 255 when rgb_range == 1, and calls utils.calc_psnr / utils.calc_ssim (ref utils.py:179-184,
:232-234), each ending in a `float(...)` host synchronisation.  `psnr_ssim` does all of that in one pass over the two
tensors and returns device scalars, so an evaluation loop only synchronises when it prints.
"""
from __future__ import annotations

import math

import torch

from . import _lib
from ._lib import M2TError

__all__ = ["psnr_ssim", "calc_psnr_ssim", "gmsd", "fsim"]


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


# ---- FSIM (ref test.py:95-96: piq.fsim(hr, sr, data_range=1., reduction='none')) --------------------------------------
# Zhang et al., "FSIM: a feature similarity index for image quality assessment" (IEEE TIP 2011) with piq's defaults.  Unlike
# the other metrics this one is not a hand-written kernel: its cost is 2 + 2 x 16 two-dimensional FFTs per image pair, which
# go to torch.fft (cuFFT); the filter bank is cached per (device, H, W, dtype).  `piq` is absent offline: parity unpinned.
_FSIM_BANK: dict = {}


def _fsim_bank(h: int, w: int, device, dtype):
    key = (str(device), h, w, dtype)
    hit = _FSIM_BANK.get(key)
    if hit is not None:
        return hit
    scales, orients, min_len, mult, sigma_f, d_theta = 4, 4, 6, 2, 0.55, 1.2

    def axis(n):
        a = torch.arange(n, dtype=torch.float64, device=device)
        return (a - (n - 1) / 2) / (n - 1) if n % 2 else (a - n / 2) / n
    fy, fx = torch.meshgrid(axis(h), axis(w), indexing="ij")          # first axis = rows, as piq builds it
    rad = torch.fft.ifftshift(torch.hypot(fy, fx))
    ang = torch.fft.ifftshift(torch.atan2(-fx, fy))
    lowpass = 1.0 / (1.0 + (rad / 0.45) ** 30)
    rad[0, 0] = 1.0
    f0 = 1.0 / (min_len * mult ** torch.arange(scales, dtype=torch.float64, device=device))
    radial = torch.exp(-(torch.log(rad[None] / f0[:, None, None]) ** 2) / (2 * math.log(sigma_f) ** 2)) * lowpass[None]
    radial[:, 0, 0] = 0.0
    centre = torch.arange(orients, dtype=torch.float64, device=device) * (math.pi / orients)
    dth = torch.atan2(torch.sin(ang[None] - centre[:, None, None]), torch.cos(ang[None] - centre[:, None, None])).abs()
    angular = torch.exp(-(dth ** 2) / (2 * (math.pi / (orients * d_theta)) ** 2))
    bank = angular[:, None] * radial[None]                            # [O,S,H,W] float64
    # noise model constants of every orientation: sum of squared smallest-scale filter, sums over the spatial filters
    spatial = torch.fft.ifft2(bank).real * math.sqrt(h * w)
    em_n = (bank[:, 0] ** 2).sum((-2, -1))
    an2 = (spatial ** 2).sum((1, 2, 3))
    aiaj = torch.zeros_like(an2)
    for s in range(scales - 1):
        aiaj = aiaj + (spatial[:, s:s + 1] * spatial[:, s + 1:]).sum((1, 2, 3))
    hit = (bank.to(dtype), (em_n.to(dtype), an2.to(dtype), aiaj.to(dtype)))
    if len(_FSIM_BANK) >= 8:
        _FSIM_BANK.pop(next(iter(_FSIM_BANK)))
    _FSIM_BANK[key] = hit
    return hit


def _fsim_pc(lum: torch.Tensor, k: float = 2.0) -> torch.Tensor:
    n, _, h, w = lum.shape
    bank, (em_n, an2, aiaj) = _fsim_bank(h, w, lum.device, lum.dtype)
    eps = torch.finfo(lum.dtype).eps
    resp = torch.fft.ifft2(torch.fft.fft2(lum)[:, :, None] * bank[None])          # [N,O,S,H,W]
    even, odd = resp.real, resp.imag
    amp = resp.abs()
    se, so = even.sum(2, keepdim=True), odd.sum(2, keepdim=True)
    norm = torch.sqrt(se * se + so * so) + eps
    me, mo = se / norm, so / norm
    energy = (even * me + odd * mo - (even * mo - odd * me).abs()).sum(2)          # [N,O,H,W]
    med = (amp[:, :, 0] ** 2).flatten(-2).median(dim=-1).values                    # [N,O]
    noise_power = (-med / math.log(0.5)) / em_n[None]
    tau = torch.sqrt((2 * noise_power * an2[None] + 4 * noise_power * aiaj[None]) / 2)
    thr = (tau * math.sqrt(math.pi / 2) + k * torch.sqrt((2 - math.pi / 2) * tau * tau)) / 1.7
    energy = torch.clamp(energy - thr[..., None, None], min=0.0)
    return ((energy.sum(1) + eps) / (amp.sum((1, 2)) + eps))[:, None]


@torch.no_grad()
def fsim(x: torch.Tensor, y: torch.Tensor, data_range: float = 1.0, chromatic: bool = True, dtype=torch.float64) -> torch.Tensor:
    """Per-image FSIM (FSIMc for RGB) of two CUDA batches [B,C,H,W], C = 1 or 3, like piq.fsim(..., reduction='none').
    Computed in `dtype` (float64 by default: the value the reference's float32 evaluation approximates)."""
    for t, n in ((x, "x"), (y, "y")):
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise M2TError(f"fsim: {n} must be a CUDA tensor; the B200 engine has no CPU path")
        if t.dim() != 4 or t.shape[1] not in (1, 3) or not t.is_floating_point():
            raise M2TError(f"fsim: {n} must be a floating-point [B,C,H,W] tensor with 1 or 3 channels, got {tuple(t.shape)}")
    if x.shape != y.shape:
        raise M2TError(f"fsim: shapes differ: {tuple(x.shape)} vs {tuple(y.shape)}")
    with torch.cuda.device(x.device):
        a, b = (t.to(dtype) * (255.0 / float(data_range)) for t in (x, y))
        ks = max(1, round(min(a.shape[-2:]) / 256))
        if ks > 1:
            a, b = torch.nn.functional.avg_pool2d(a, ks), torch.nn.functional.avg_pool2d(b, ks)
        if a.shape[1] == 3:
            m = torch.tensor([[0.299, 0.587, 0.114], [0.5959, -0.2746, -0.3213], [0.2115, -0.5227, 0.3112]], dtype=dtype, device=a.device)
            a, b = (torch.einsum("kc,nchw->nkhw", m, t) for t in (a, b))
        else:
            chromatic = False
        la, lb = a[:, :1], b[:, :1]
        pa, pb = _fsim_pc(la), _fsim_pc(lb)
        sch = torch.tensor([[-3.0, 0.0, 3.0], [-10.0, 0.0, 10.0], [-3.0, 0.0, 3.0]], dtype=dtype, device=a.device) / 16.0
        kern = torch.stack((sch, sch.t()))[:, None]
        ga, gb = (torch.sqrt((torch.nn.functional.conv2d(t, kern, padding=1) ** 2).sum(1, keepdim=True)) for t in (la, lb))

        def sim(p, q, c):
            return (2 * p * q + c) / (p * p + q * q + c)
        pmax = torch.maximum(pa, pb)
        score = sim(ga, gb, 160.0) * sim(pa, pb, 0.85) * pmax
        if chromatic:
            score = score * (sim(a[:, 1:2], b[:, 1:2], 200.0) * sim(a[:, 2:3], b[:, 2:3], 200.0)).abs() ** 0.03
        return (score.sum((1, 2, 3)) / pmax.sum((1, 2, 3))).float()
