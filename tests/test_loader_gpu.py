"""uint8 HWC -> fp32 CHW / 255 on the device is bit-identical to the reference loader's CPU expression."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _reference(img):            # ref datas/benchmark.py:66-69 with utils.ndarray2tensor (utils.py:237-240)
    return torch.from_numpy(np.ascontiguousarray(img.transpose((2, 0, 1)))).float() / 255.


@pytest.mark.parametrize("shape", [(1, 64, 64, 3), (3, 200, 266, 3), (2, 37, 51, 3), (2, 33, 40, 1), (1, 1080, 1920, 3)])
def test_images_to_device_bit_exact(shape):
    from m2trans_b200.loader import images_to_device
    rng = np.random.default_rng(shape[1])
    imgs = rng.integers(0, 256, size=shape, dtype=np.uint8)
    out = images_to_device(imgs).cpu()
    ref = torch.stack([_reference(im) for im in imgs])
    assert out.shape == ref.shape and torch.equal(out, ref)
    one = images_to_device(torch.from_numpy(imgs[0]).pin_memory()).cpu()
    assert torch.equal(one[0], ref[0])


def test_images_to_device_all_byte_values_and_errors():
    from m2trans_b200._lib import M2TError
    from m2trans_b200.loader import images_to_device
    ramp = np.arange(256, dtype=np.uint8).reshape(1, 16, 16, 1).repeat(3, axis=3)
    assert torch.equal(images_to_device(ramp).cpu()[0, 0].flatten(), torch.arange(256).float() / 255.)
    with pytest.raises(M2TError):
        images_to_device(np.zeros((4, 4, 3), dtype=np.float32))
    with pytest.raises(M2TError):
        images_to_device(np.zeros((4, 4, 2), dtype=np.uint8))
    with pytest.raises(M2TError):
        images_to_device(np.zeros((4, 4, 3), dtype=np.uint8), device="cpu")


@pytest.mark.parametrize("shape", [(1, 3, 64, 64), (2, 3, 37, 51), (2, 1, 33, 40), (1, 3, 512, 512)])
def test_images_from_device_matches_torch(shape):
    """fp32 CHW -> uint8 HWC on the device (what writing the SR batch to image files does) equals the torch expression."""
    from m2trans_b200.loader import images_from_device, images_to_device
    g = torch.Generator().manual_seed(shape[2])
    sr = (torch.rand(shape, generator=g) * 1.2 - 0.1).cuda()                       # includes values outside [0, 1]
    sr[0, 0, 0, :8] = torch.tensor([0.5 / 255, 1.5 / 255, 2.5 / 255, 0.0, 1.0, -3.0, 7.0, 254.5 / 255]).cuda()   # ties, ends
    want = (sr * 255.0).round().clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()
    got = images_from_device(sr)
    assert got.shape == want.shape and got.dtype == torch.uint8 and torch.equal(got, want)
    back = images_to_device(got.cpu().numpy())                                      # round trip through the loader kernel
    assert torch.equal(back.cpu(), want.permute(0, 3, 1, 2).cpu().float() / 255.)        # the loader's expression, on the CPU


def test_eval_loop_uint8_to_metrics_matches_oracle():
    """The reference's test loop (test.py:87-116) end to end on the device: uint8 LR/HR arrays -> loader conversion ->
    model(lr) -> Y-channel PSNR / SSIM, against the CPU oracle of each piece chained the same way."""
    import types
    from oracle import m2trans_oracle as O
    from oracle import metrics_oracle as M
    from m2trans_b200.M2Trans_network import M2Trans
    from m2trans_b200.loader import images_to_device
    from m2trans_b200.metrics import psnr_ssim
    from m2trans_b200.synthetic import synthetic_input, synthetic_state_dict
    scale = 2
    hr_f = synthetic_input(1, 96, 128, seed=21)
    hr_u8 = (hr_f[0].permute(1, 2, 0) * 255).round().to(torch.uint8).numpy()
    lr_u8 = np.ascontiguousarray(hr_u8[::scale, ::scale])                     # a stand-in for LR_bicubic
    model = M2Trans(types.SimpleNamespace(scale=scale, rgb_range=1.0, colors=3, n_feats=64, n_blocks=8)).cuda()
    sd = synthetic_state_dict(scale, 0)
    model.load_state_dict(sd)
    lr, hr = images_to_device(lr_u8), images_to_device(hr_u8)
    sr = model(lr)
    assert sr.shape == hr.shape
    _, batch = psnr_ssim(sr, hr, scale)
    sr_ref = O.forward(sd, _reference(lr_u8)[None])
    p_ref, s_ref = M.test_loop_metrics(sr_ref, _reference(hr_u8)[None], scale, dtype=torch.float64)
    print(f"eval loop: psnr {float(batch[0]):.4f} (oracle {p_ref:.4f}) ssim {float(batch[1]):.6f} (oracle {s_ref:.6f})")
    assert abs(float(batch[0]) - p_ref) <= 0.02 and abs(float(batch[1]) - s_ref) <= 1e-3
