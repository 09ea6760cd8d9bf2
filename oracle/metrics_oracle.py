"""CPU oracle for the evaluation metrics of the reference's test loop.  TEST INFRASTRUCTURE ONLY.

PSNR side: restates ref utils.py:119-146 (rgb_to_ycbcr), test.py:103-112 (Y, shave, x255) and utils.py:179-184
(calc_psnr); PINNED on fixtures produced by those reference functions themselves (oracle/make_golden_metrics.py imports
/root/reference/utils.py with its unavailable imports stubbed; tests/golden/metrics_*.npz).

SSIM side: ref utils.py:232-234 calls `pytorch_msssim.ssim(sr, hr, size_average=True)`; the package (pinned
pytorch-msssim==1.0.0, ref environment.yml:134) is third-party and absent, so its published algorithm is restated here
(Gaussian window 11, sigma 1.5, separable valid convolution, data_range 255, K = (0.01, 0.03), mean of the SSIM map per
channel, then over the batch) -- PARITY UNPINNED for SSIM.  `ssim(..., dtype=torch.float64)` is the exact-arithmetic value;
in float32 (what the reference runs) the 4080 offset of the x255 Y channel makes E[x^2] - mu^2 noisy at the 1e-4 level.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def rgb_to_ycbcr(image: torch.Tensor) -> torch.Tensor:
    """ref utils.py:119-146."""
    image = image / 255.
    r, g, b = image[..., 0, :, :], image[..., 1, :, :], image[..., 2, :, :]
    y = 65.481 * r + 128.553 * g + 24.966 * b + 16.0
    cb = -37.797 * r + -74.203 * g + 112.0 * b + 128.0
    cr = 112.0 * r + -93.786 * g + -18.214 * b + 128.0
    return torch.stack((y, cb, cr), -3)


def prepare(sr, hr, scale, rgb_range=1.0, colors=3):
    """ref test.py:103-112."""
    if colors == 3:
        hr = rgb_to_ycbcr(hr)[:, 0:1, :, :]
        sr = rgb_to_ycbcr(sr)[:, 0:1, :, :]
    hr = hr[:, :, scale:-scale, scale:-scale]
    sr = sr[:, :, scale:-scale, scale:-scale]
    if rgb_range == 1:
        hr, sr = hr * 255., sr * 255.
    return sr, hr


def calc_psnr(sr, hr):
    """ref utils.py:179-184."""
    sr, hr = sr.double(), hr.double()
    diff = (sr - hr) / 255.00
    mse = diff.pow(2).mean()
    return float(-10 * math.log10(mse))


def _gauss(size=11, sigma=1.5, dtype=torch.float32):
    c = torch.arange(size, dtype=dtype) - size // 2
    g = torch.exp(-(c ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def ssim(x, y, data_range=255.0, dtype=torch.float32):
    """pytorch_msssim.ssim(X, Y, data_range=255, size_average=True, win_size=11, win_sigma=1.5, K=(0.01, 0.03))."""
    x, y = x.to(dtype), y.to(dtype)
    c = x.shape[1]
    g = _gauss(dtype=dtype)
    wh, ww = g.view(1, 1, -1, 1).repeat(c, 1, 1, 1), g.view(1, 1, 1, -1).repeat(c, 1, 1, 1)

    def blur(t):
        return F.conv2d(F.conv2d(t, wh, groups=c), ww, groups=c)

    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    mu1, mu2 = blur(x), blur(y)
    s1, s2, s12 = blur(x * x) - mu1 * mu1, blur(y * y) - mu2 * mu2, blur(x * y) - mu1 * mu2
    cs = (2 * s12 + c2) / (s1 + s2 + c2)
    m = ((2 * mu1 * mu2 + c1) / (mu1 * mu1 + mu2 * mu2 + c1)) * cs
    return float(m.flatten(2).mean(-1).mean())


def test_loop_metrics(sr, hr, scale, rgb_range=1.0, colors=3, dtype=torch.float32):
    """(psnr, ssim) of one batch as ref test.py:103-116 computes them."""
    s, h = prepare(sr, hr, scale, rgb_range, colors)
    return calc_psnr(s, h), ssim(s, h, dtype=dtype)


def gmsd(x, y, data_range=1.0):
    """piq.gmsd(x, y, data_range, reduction='none') (ref test.py:98-99).  `piq` (pinned nowhere: `pip install piq`, ref
    README) is absent offline -- PARITY UNPINNED; this restates the published algorithm (Xue et al., IEEE TIP 2014) the way
    piq implements it: luma of x / data_range, zero pad bottom / right by max(H % 2, W % 2), 2 x 2 average pooling,
    Prewitt / 3 gradient magnitudes with zero padding, GMS with c = 170 / 255^2, population standard deviation per image."""
    import torch.nn.functional as F

    def luma(t):
        t = t.double() / data_range
        return (0.299 * t[:, 0:1] + 0.587 * t[:, 1:2] + 0.114 * t[:, 2:3]) if t.shape[1] == 3 else t
    a, b = luma(x), luma(y)
    p = max(a.shape[2] % 2, a.shape[3] % 2)
    a, b = (F.avg_pool2d(F.pad(t, (0, p, 0, p)), 2, 2) for t in (a, b))
    k = torch.tensor([[1.0, 0.0, -1.0]] * 3, dtype=torch.float64) / 3.0
    kern = torch.stack((k, k.t()))[:, None]
    ga, gb = (torch.sqrt((F.conv2d(t, kern, padding=1) ** 2).sum(1, keepdim=True) + 1e-12) for t in (a, b))
    c = 170.0 / 255.0 ** 2
    gms = (2 * ga * gb + c) / (ga ** 2 + gb ** 2 + c)
    return gms.flatten(1).std(dim=1, unbiased=False)
