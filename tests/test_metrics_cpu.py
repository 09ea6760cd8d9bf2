"""Oracle for the test-loop metrics against fixtures made by the reference's own utils.py."""
import glob
import os

import numpy as np
import pytest
import torch

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "metrics_*.npz")))


def test_fixtures_present():
    assert len(GOLD) == 3


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_oracle_matches_reference_fixture(path):
    from oracle import metrics_oracle as M
    z = np.load(path)
    sr, hr, scale = torch.from_numpy(z["sr"]), torch.from_numpy(z["hr"]), int(z["scale"])
    s, h = M.prepare(sr, hr, scale)
    assert torch.equal(s, torch.from_numpy(z["sr_y"])) and torch.equal(h, torch.from_numpy(z["hr_y"]))
    assert M.calc_psnr(s, h) == float(z["psnr"])
    for i, p in enumerate(z["psnr_per_image"]):
        assert M.calc_psnr(s[i:i + 1], h[i:i + 1]) == float(p)


def test_ssim_known_answers():
    from oracle import metrics_oracle as M
    g = torch.Generator().manual_seed(0)
    x = torch.rand(2, 1, 40, 52, generator=g) * 219 + 4080
    assert abs(M.ssim(x, x, dtype=torch.float64) - 1.0) < 1e-12
    y = x + 5 * torch.randn(x.shape, generator=g)
    a, b = M.ssim(x, y, dtype=torch.float64), M.ssim(x, y, dtype=torch.float32)
    assert 0.0 < a < 1.0
    assert abs(a - b) < 2e-3            # the reference's float32 path is only this good on the 4080-offset Y channel
    # shift invariance of the structure term: removing the offset changes only the (near 1) luminance term
    c = M.ssim(x - 4080, y - 4080, dtype=torch.float64)
    assert abs(a - c) < 0.05
    assert M._gauss().sum().item() == pytest.approx(1.0, abs=1e-6)


def test_gmsd_oracle_properties():
    """The GMSD restatement (piq absent: parity unpinned): 0 for identical images, symmetric, grows with distortion, odd
    sizes padded like piq (both dimensions by max(H % 2, W % 2))."""
    import torch
    from oracle import metrics_oracle as MO
    g = torch.Generator().manual_seed(3)
    a = torch.rand(2, 3, 37, 50, generator=g)
    n = torch.randn(a.shape, generator=g)
    small, big = (a + 0.02 * n).clamp(0, 1), (a + 0.2 * n).clamp(0, 1)
    assert float(MO.gmsd(a, a).abs().max()) == 0.0
    assert torch.allclose(MO.gmsd(a, small), MO.gmsd(small, a))
    assert (MO.gmsd(a, big) > MO.gmsd(a, small)).all() and float(MO.gmsd(a, big).max()) < 0.35
    gray = a[:, :1]
    assert MO.gmsd(gray, gray.flip(3)).shape == (2,)


def test_fsim_oracle_properties():
    """FSIM restated from the published algorithm (piq absent: unpinned): 1 for identical images, falls monotonically with
    noise, handles grayscale (no chromatic term), odd sizes and the average-pooling path of large frames (min side >= 384)."""
    import torch
    from oracle import metrics_oracle as MO
    g = torch.Generator().manual_seed(3)
    x = torch.nn.functional.avg_pool2d(torch.rand(2, 3, 75, 98, generator=g), 3, 1, 1)
    assert torch.allclose(MO.fsim(x, x), torch.ones(2, dtype=torch.float64), atol=1e-12)
    prev = 1.0
    for sigma in (0.01, 0.03, 0.1, 0.3):
        v = float(MO.fsim(x, (x + sigma * torch.randn(x.shape, generator=g)).clamp(0, 1)).mean())
        assert 0.0 < v < prev
        prev = v
    gray = x[:, :1]
    assert torch.allclose(MO.fsim(gray, gray), torch.ones(2, dtype=torch.float64), atol=1e-12)
    assert float(MO.fsim(gray, gray * 0.8).min()) < 1.0
    big = torch.rand(1, 3, 400, 390, generator=g)
    v = MO.fsim(big, (big + 0.05 * torch.randn(big.shape, generator=g)).clamp(0, 1))
    assert v.shape == (1,) and 0.5 < float(v) < 1.0
    # data_range only rescales
    assert torch.allclose(MO.fsim(x * 255.0, (x * 0.9) * 255.0, data_range=255.0), MO.fsim(x, x * 0.9), atol=1e-12)


def test_fsim_argument_errors():
    import pytest
    import torch
    from m2trans_b200 import metrics
    from m2trans_b200._lib import M2TError
    with pytest.raises(M2TError):
        metrics.fsim(torch.rand(1, 3, 32, 32), torch.rand(1, 3, 32, 32))          # no CPU path
