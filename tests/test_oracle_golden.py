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


@pytest.mark.parametrize("name", FWD)
def test_forward(golden_dir, name):
    g = _load(golden_dir, name + ".npz")
    scale, seed, gain = int(g["scale"]), int(g["seed"]), float(g["qkv_gain"])
    sd = synthetic_state_dict(scale, seed, qkv_gain=gain)
    # the seeded weights are the ones the fixture was generated with
    np.testing.assert_allclose(_wsum(sd), g["wsum"], rtol=1e-12)
    y = O.forward(sd, torch.from_numpy(g["x"]))
    assert tuple(y.shape) == g["y"].shape
    ref = torch.from_numpy(g["y"])
    assert O.max_abs(y, ref) <= 2e-5
    assert O.psnr(y, ref) >= 90.0


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
