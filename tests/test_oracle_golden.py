"""The CPU oracle (oracle/m2trans_oracle.py) against fixtures produced by the REAL
reference module (oracle/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import m2trans_oracle as O
from m2trans_b200.synthetic import state_dict_spec, synthetic_state_dict

torch.set_grad_enabled(False)


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _wsum(sd):
    return np.array([sum(float(v.double().sum()) for v in sd.values()),
                     sum(float(v.double().abs().sum()) for v in sd.values())])


def test_dwt_iwt(golden_dir):
    u = _load(golden_dir, "units.npz")
    np.testing.assert_allclose(O.dwt(torch.from_numpy(u["dwt_in"])).numpy(), u["dwt_out"], atol=1e-6)
    np.testing.assert_allclose(O.iwt(torch.from_numpy(u["iwt_in"])).numpy(), u["iwt_out"], atol=1e-6)
    x = torch.randn(1, 4, 16, 8)
    np.testing.assert_allclose(O.iwt(O.dwt(x)).numpy(), x.numpy(), atol=1e-5)


@pytest.mark.parametrize("ch", [16, 64, 256])
def test_tblock(golden_dir, ch):
    u = _load(golden_dir, "units.npz")
    out = O.tblock(torch.from_numpy(u[f"tb{ch}_in"]), torch.from_numpy(u[f"tb{ch}_wqkv"]),
                   torch.from_numpy(u[f"tb{ch}_relh"]), torch.from_numpy(u[f"tb{ch}_relw"]))
    np.testing.assert_allclose(out.numpy(), u[f"tb{ch}_out"], atol=2e-5, rtol=1e-5)


def test_cftm(golden_dir):
    u = _load(golden_dir, "units.npz")
    sd = synthetic_state_dict(2, 5, n_blocks=1)
    np.testing.assert_allclose(_wsum(sd), u["cftm_wsum"], rtol=1e-12)
    out = O.cftm(sd, 0, torch.from_numpy(u["cftm_in"]))
    np.testing.assert_allclose(out.numpy(), u["cftm_out"], atol=2e-5, rtol=1e-5)


FWD = ["fwd_x2_64x64", "fwd_x3_40x50", "fwd_x4_24x40", "fwd_x4_32x32_sharp", "fwd_x4_b2_32x32_speckle"]
# reference outputs of checkpoints whose last conv is rescaled / offset so that ~95 % of the SR pixels lie strictly inside
# (0, 1): with the plain initialisation half of every output is clamped to 0, where any implementation is exact
UNCLAMPED = ["unc_x2_64x64", "unc_x3_40x50", "unc_x4_24x40", "unc_x4_b2_32x32_speckle", "unc_x4_32x32_g125_speckle",
             "unc_x3_64x40_g125_speckle", "unc_x4_32x32_sharp_speckle", "unc_x3_64x40_sharp_speckle", "unc_x2_48x64_flat",
             "unc_x4_128x128_cfg2_frame"]


def golden_state_dict(g):
    """The seeded checkpoint a fixture was generated with (older fixtures carry no out_gain / out_shift)."""
    kw = {"qkv_gain": float(g["qkv_gain"])}
    if "out_gain" in g.files:
        kw.update(out_gain=float(g["out_gain"]), out_shift=float(g["out_shift"]))
    return synthetic_state_dict(int(g["scale"]), int(g["seed"]), **kw)


@pytest.mark.parametrize("name", FWD + UNCLAMPED)
def test_forward(golden_dir, name):
    g = _load(golden_dir, name + ".npz")
    sd = golden_state_dict(g)
    # the seeded weights are the ones the fixture was generated with
    np.testing.assert_allclose(_wsum(sd), g["wsum"], rtol=1e-12)
    y = O.forward(sd, torch.from_numpy(g["x"]))
    assert tuple(y.shape) == g["y"].shape
    ref = torch.from_numpy(g["y"])
    assert O.max_abs(y, ref) <= 2e-5
    assert O.psnr(y, ref) >= 90.0
    if name in UNCLAMPED:                              # the fixture really is unclamped
        clamped = float(((ref <= 0.0) | (ref >= 1.0)).float().mean())
        assert clamped <= (0.16 if ("sharp" in name or "g125" in name) else 0.05), clamped


def test_forward_accepts_dataparallel_prefix(golden_dir):
    g = _load(golden_dir, "fwd_x4_24x40.npz")
    sd = synthetic_state_dict(4, 2)
    y = O.forward({"module." + k: v for k, v in sd.items()}, torch.from_numpy(g["x"]))
    assert O.max_abs(y, torch.from_numpy(g["y"])) <= 2e-5


@pytest.mark.parametrize("scale", [2, 3, 4])
def test_state_dict_manifest(golden_dir, scale):
    """Our key/shape spec equals the reference's state_dict() (121/121/123 tensors)."""
    want = []
    for line in open(os.path.join(golden_dir, "state_dict_manifest.txt")):
        s, key, shape, dtype = line.split()
        if s == f"x{scale}":
            want.append((key, tuple(int(t) for t in shape.split("x")), dtype))
    got = [(k, tuple(s), "float32") for k, s in state_dict_spec(scale)]
    assert got == want
    assert len(got) == (123 if scale == 4 else 121)
    sd = synthetic_state_dict(scale, 0)
    assert [(k, tuple(v.shape)) for k, v in sd.items()] == [(k, s) for k, s, _ in want]
