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


# ---------------------------------------------------------------------------------------------------------------------
# FSIM (ref test.py:95-96: piq.fsim(hr, sr, data_range=1., reduction='none')).  `piq` is absent offline -- PARITY UNPINNED.
# Restated from the published algorithm (Zhang, Zhang, Mou, Zhang: "FSIM: a feature similarity index for image quality
# assessment", IEEE TIP 2011, with Kovesi's phase congruency PC_2 as in the authors' phasecong2) with piq's defaults:
# 4 scales, 4 orientations, min wavelength 6, mult 2, sigma_f 0.55, delta_theta 1.2, k 2, chromatic (FSIMc) for RGB.
def _fsim_meshgrid(h, w):
    def axis(n):
        if n % 2:
            return torch.arange(-(n - 1) / 2, n / 2, dtype=torch.float64) / (n - 1)
        return torch.arange(-n / 2, n / 2, dtype=torch.float64) / n
    return torch.meshgrid(axis(h), axis(w), indexing="ij")


def _fsim_filters(h, w, scales=4, orientations=4, min_length=6, mult=2, sigma_f=0.55, delta_theta=1.2):
    gx, gy = _fsim_meshgrid(h, w)
    radius = torch.fft.ifftshift(torch.sqrt(gx ** 2 + gy ** 2))
    theta = torch.fft.ifftshift(torch.atan2(-gy, gx))
    lowpass = 1.0 / (1.0 + (radius / 0.45) ** 30)          # Butterworth, cutoff 0.45, order 15 (radius already shifted)
    radius[0, 0] = 1.0
    sin_t, cos_t = torch.sin(theta), torch.cos(theta)
    log_gabor = []
    for s in range(scales):
        f0 = 1.0 / (min_length * mult ** s)
        g = torch.exp(-(torch.log(radius / f0) ** 2) / (2 * math.log(sigma_f) ** 2)) * lowpass
        g[0, 0] = 0.0
        log_gabor.append(g)
    theta_sigma = math.pi / (orientations * delta_theta)
    spread = []
    for o in range(orientations):
        a = o * math.pi / orientations
        ds = sin_t * math.cos(a) - cos_t * math.sin(a)
        dc = cos_t * math.cos(a) + sin_t * math.sin(a)
        spread.append(torch.exp(-(torch.atan2(ds, dc).abs() ** 2) / (2 * theta_sigma ** 2)))
    # [orientations, scales, h, w]
    return torch.stack(spread)[:, None] * torch.stack(log_gabor)[None]


def _phase_congruency(lum, k=2.0):
    """lum [N,1,H,W] float64 in [0,255] -> PC_2 map [N,1,H,W]."""
    n, _, h, w = lum.shape
    filt = _fsim_filters(h, w)                                         # [O,S,H,W]
    o_n, s_n = filt.shape[:2]
    eps = torch.finfo(torch.float64).eps
    resp = torch.fft.ifft2(torch.fft.fft2(lum)[:, :, None] * filt[None])   # [N,O,S,H,W] complex: even + i odd
    even, odd = resp.real, resp.imag
    an = resp.abs()
    sum_e, sum_o = even.sum(2, keepdim=True), odd.sum(2, keepdim=True)
    x_energy = torch.sqrt(sum_e ** 2 + sum_o ** 2) + eps
    mean_e, mean_o = sum_e / x_energy, sum_o / x_energy
    energy = (even * mean_e + odd * mean_o - (even * mean_o - odd * mean_e).abs()).sum(2, keepdim=True)
    # noise threshold from the smallest scale (Rayleigh statistics of the filter response to noise)
    em_n = (filt[:, :1] ** 2).sum((-2, -1), keepdim=True)[None]        # [1,O,1,1,1]
    median_e2n = (an[:, :, :1] ** 2).flatten(-2).median(dim=-1, keepdim=True).values[..., None]   # lower median, like torch
    noise_power = (-median_e2n / math.log(0.5)) / em_n
    f_ifft = torch.fft.ifft2(filt).real * math.sqrt(h * w)             # [O,S,H,W]
    sum_an2 = (f_ifft ** 2).sum(1, keepdim=True).sum((-2, -1), keepdim=True)[None]
    sum_aiaj = torch.zeros_like(sum_an2)
    for s in range(s_n - 1):
        sum_aiaj = sum_aiaj + (f_ifft[:, s:s + 1] * f_ifft[:, s + 1:]).sum(1, keepdim=True).sum((-2, -1), keepdim=True)[None]
    noise_energy2 = 2 * noise_power * sum_an2 + 4 * noise_power * sum_aiaj
    tau = torch.sqrt(noise_energy2 / 2)
    thr = (tau * math.sqrt(math.pi / 2) + k * torch.sqrt((2 - math.pi / 2) * tau ** 2)) / 1.7
    energy = torch.clamp(energy - thr, min=0.0)
    return ((energy.sum((1, 2)) + eps) / (an.sum((1, 2)) + eps))[:, None]


def fsim(x, y, data_range=1.0, chromatic=True):
    """Per-image FSIM (FSIMc for 3-channel inputs) of [N,C,H,W] tensors, float64 arithmetic."""
    import torch.nn.functional as F
    x, y = (t.double() / float(data_range) * 255.0 for t in (x, y))
    ks = max(1, round(min(x.shape[-2:]) / 256))
    x, y = F.avg_pool2d(x, ks), F.avg_pool2d(y, ks)
    if x.shape[1] == 3:
        m = torch.tensor([[0.299, 0.587, 0.114], [0.5959, -0.2746, -0.3213], [0.2115, -0.5227, 0.3112]], dtype=torch.float64)
        xq, yq = (torch.einsum("kc,nchw->nkhw", m, t) for t in (x, y))
        xl, yl = xq[:, :1], yq[:, :1]
    else:
        xl, yl, chromatic = x, y, False
    pcx, pcy = _phase_congruency(xl), _phase_congruency(yl)
    sch = torch.tensor([[-3.0, 0.0, 3.0], [-10.0, 0.0, 10.0], [-3.0, 0.0, 3.0]], dtype=torch.float64) / 16.0
    kern = torch.stack((sch, sch.t()))[:, None]
    gx, gy = (torch.sqrt((F.conv2d(t, kern, padding=1) ** 2).sum(1, keepdim=True)) for t in (xl, yl))

    def sim(a, b, c):
        return (2 * a * b + c) / (a ** 2 + b ** 2 + c)
    pc_max = torch.maximum(pcx, pcy)
    score = sim(gx, gy, 160.0) * sim(pcx, pcy, 0.85) * pc_max
    if chromatic:
        score = score * (sim(xq[:, 1:2], yq[:, 1:2], 200.0) * sim(xq[:, 2:3], yq[:, 2:3], 200.0)).abs() ** 0.03
    return score.sum((1, 2, 3)) / pc_max.sum((1, 2, 3))
