"""Device PSNR / SSIM of the test loop against the oracle and the reference-made fixtures."""
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "metrics_*.npz")))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_psnr_matches_reference_fixture(path):
    from oracle import metrics_oracle as M
    from m2trans_b200.metrics import psnr_ssim
    z = np.load(path)
    sr, hr, scale = torch.from_numpy(z["sr"]), torch.from_numpy(z["hr"]), int(z["scale"])
    per, batch = psnr_ssim(sr.cuda(), hr.cuda(), scale)
    per, batch = per.cpu(), batch.cpu()
    assert abs(float(batch[0]) - float(z["psnr"])) <= 2e-4           # dB; fp32 Y and squared error, fp64 sums
    assert np.abs(per[:, 0].numpy() - z["psnr_per_image"]).max() <= 2e-4
    s, h = M.prepare(sr, hr, scale)
    for i in range(sr.shape[0]):
        assert abs(float(per[i, 1]) - M.ssim(s[i:i + 1], h[i:i + 1], dtype=torch.float64)) <= 2e-5
    assert abs(float(batch[1]) - M.ssim(s, h, dtype=torch.float64)) <= 2e-5


@pytest.mark.parametrize("b,c,h,w,scale,rr", [(1, 3, 512, 512, 4, 1.0), (3, 3, 270, 481, 3, 1.0), (2, 1, 64, 75, 2, 1.0),
                                              (1, 3, 96, 43, 4, 255.0), (2, 3, 19, 19, 4, 1.0)])
def test_psnr_ssim_against_oracle(b, c, h, w, scale, rr):
    from oracle import metrics_oracle as M
    from m2trans_b200.metrics import calc_psnr_ssim, psnr_ssim
    g = torch.Generator().manual_seed(h + w)
    hr = torch.rand(b, c, h, w, generator=g)
    hr = torch.nn.functional.avg_pool2d(hr, 5, 1, 2)                    # some spatial structure
    hr = (hr - hr.min()) / (hr.max() - hr.min()) * rr
    sr = (hr + 0.02 * rr * torch.randn(hr.shape, generator=g)).clamp(0, rr)
    per, batch = psnr_ssim(sr.cuda(), hr.cuda(), scale, rr)
    p_ref, s_ref = M.test_loop_metrics(sr, hr, scale, rr, c, dtype=torch.float64)
    p32, s32 = M.test_loop_metrics(sr, hr, scale, rr, c, dtype=torch.float32)
    print(f"{(b, c, h, w)} x{scale}: psnr {float(batch[0]):.4f} (oracle {p_ref:.4f})  ssim {float(batch[1]):.6f} "
          f"(oracle fp64 {s_ref:.6f}, fp32 {s32:.6f})")
    assert abs(float(batch[0]) - p_ref) <= 2e-4
    assert abs(float(batch[1]) - s_ref) <= 2e-5
    if (h - 2 * scale - 10) * (w - 2 * scale - 10) >= 1000:      # the reference's own float32 arithmetic scatters: 1e-2 per pixel,
        assert abs(float(batch[1]) - s32) <= 2e-3               # averaging out over the valid region
    assert calc_psnr_ssim(sr.cuda(), hr.cuda(), scale, rr) == tuple(batch.tolist())
    for i in range(b):
        pi, si = M.test_loop_metrics(sr[i:i + 1], hr[i:i + 1], scale, rr, c, dtype=torch.float64)
        assert abs(float(per[i, 0]) - pi) <= 2e-4 and abs(float(per[i, 1]) - si) <= 2e-5


def test_identical_images_and_errors():
    from m2trans_b200._lib import M2TError
    from m2trans_b200.metrics import psnr_ssim
    x = torch.rand(1, 3, 40, 40, device="cuda")
    per, _ = psnr_ssim(x, x.clone(), 2)
    assert float(per[0, 1]) == pytest.approx(1.0, abs=1e-6) and torch.isinf(per[0, 0])      # mse 0: the reference raises on log10(0)
    with pytest.raises(M2TError):
        psnr_ssim(x, x[:, :, :30], 2)
    with pytest.raises(M2TError):
        psnr_ssim(x[:, :, :14, :14].contiguous(), x[:, :, :14, :14].contiguous(), 2)        # 10 x 10 after shaving
    with pytest.raises(M2TError):
        psnr_ssim(x.cpu(), x.cpu(), 2)


@pytest.mark.parametrize("shape", [(2, 3, 64, 96), (1, 3, 37, 50), (3, 1, 33, 33), (1, 3, 512, 512)])
def test_gmsd_matches_oracle(shape):
    """piq.gmsd of ref test.py:98-99 on the device against the oracle's fp64 restatement of the published algorithm."""
    from oracle import metrics_oracle as MO
    from m2trans_b200.metrics import gmsd
    g = torch.Generator().manual_seed(shape[2] + shape[3])
    hr = torch.rand(shape, generator=g)
    sr = (hr + 0.05 * torch.randn(shape, generator=g)).clamp(0, 1)
    got = gmsd(hr.cuda(), sr.cuda()).cpu().double()
    want = MO.gmsd(hr, sr)
    print(f"gmsd {shape}: {got.tolist()} vs {want.tolist()}")
    assert got.shape == (shape[0],) and float((got - want).abs().max()) <= 2e-5
    assert float(gmsd(hr.cuda(), hr.cuda()).abs().max()) <= 1e-6          # identical images: GMS == 1 everywhere


@pytest.mark.parametrize("shape", [(2, 3, 75, 98), (1, 1, 64, 80), (1, 3, 400, 390)])
def test_fsim_matches_oracle(shape):
    """Device FSIM (torch.fft on the GPU, filter bank cached) against the oracle's float64 restatement: float64 path to
    1e-9, float32 path (what piq computes in) to 2e-4; wrong shapes / channel counts are refused."""
    import torch
    from m2trans_b200 import metrics
    from m2trans_b200._lib import M2TError
    from oracle import metrics_oracle as MO
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.nn.functional.avg_pool2d(torch.rand(shape, generator=g), 3, 1, 1)
    y = (x + 0.04 * torch.randn(shape, generator=g)).clamp(0, 1)
    want = MO.fsim(x, y)
    got64 = metrics.fsim(x.cuda(), y.cuda())
    got32 = metrics.fsim(x.cuda(), y.cuda(), dtype=torch.float32)
    print(shape, want.tolist(), got64.tolist(), got32.tolist())
    assert got64.shape == (shape[0],) and got64.dtype == torch.float32
    assert float((got64.double().cpu() - want).abs().max()) <= 1e-6       # float32 result of a float64 evaluation
    assert float((got32.double().cpu() - want).abs().max()) <= 2e-4
    assert torch.allclose(metrics.fsim(x.cuda(), x.cuda()), torch.ones(shape[0], device="cuda"), atol=1e-6)
    with pytest.raises(M2TError):
        metrics.fsim(x.cuda(), y.cuda()[:, :, :-1])
    with pytest.raises(M2TError):
        metrics.fsim(torch.rand(1, 2, 32, 32, device="cuda"), torch.rand(1, 2, 32, 32, device="cuda"))
