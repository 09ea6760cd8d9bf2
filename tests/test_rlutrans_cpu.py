"""rlutrans.TransBlock on the CPU: the oracle against fixtures produced by the REAL reference
module (oracle/make_golden_rlutrans.py), and the host mirror's state-dict surface / error behaviour."""
import os

import numpy as np
import pytest
import torch

from oracle import rlutrans_oracle as R
from m2trans_b200 import rlutrans as ours
from m2trans_b200._lib import M2TError
from m2trans_b200.synthetic import synthetic_transblock_state_dict, transblock_state_dict_spec

torch.set_grad_enabled(False)
CASES = [("rlutrans_b2_n256", 0), ("rlutrans_b3_n100", 1), ("rlutrans_b1_n16", 2), ("rlutrans_b1_n4500", 3)]


@pytest.mark.parametrize("name,seed", CASES)
def test_oracle_matches_reference_fixture(golden_dir, name, seed):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    sd = synthetic_transblock_state_dict(seed)
    assert abs(sum(float(v.double().sum()) for v in sd.values()) - float(z["wsum"])) < 1e-6, "weight generator drifted"
    y = R.transblock(sd, torch.from_numpy(z["x"]))
    np.testing.assert_allclose(y.numpy(), z["y"], atol=5e-6, rtol=0)


def test_oracle_chunking_is_block_diagonal():
    """Tokens of one chunk never see another chunk: perturbing chunk 3 leaves every other chunk's output intact."""
    sd = synthetic_transblock_state_dict(0)
    x = torch.randn(1, 160, 64)
    y0 = R.transblock(sd, x)
    x2 = x.clone()
    x2[:, 30:40] += 1.0                       # chunk length 10 -> chunk 3
    y1 = R.transblock(sd, x2)
    d = (y1 - y0).abs().amax(dim=(0, 2))
    assert float(d[:30].max()) == 0.0 and float(d[40:].max()) == 0.0 and float(d[30:40].max()) > 1e-3


def test_oracle_rejects_short_sequences():
    with pytest.raises(ValueError):
        R.transblock(synthetic_transblock_state_dict(0), torch.randn(1, 15, 64))


def test_mirror_state_dict_surface(golden_dir):
    m = ours.TransBlock()
    want = [ln.split(" ", 1) for ln in open(os.path.join(golden_dir, "rlutrans_state_dict_manifest.txt")).read().splitlines()]
    got = [[k, f"{tuple(v.shape)} {v.dtype}"] for k, v in m.state_dict().items()]
    assert got == want
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == transblock_state_dict_spec()
    assert sum(v.numel() for v in m.state_dict().values()) == 22928
    m.load_state_dict(synthetic_transblock_state_dict(1), strict=True)


def test_mirror_has_no_cpu_path_and_rejects_other_shapes():
    m = ours.TransBlock()
    with pytest.raises(M2TError):
        m(torch.randn(1, 64, 64))              # CPU tensor
    with pytest.raises(M2TError):
        ours.TransBlock(dim=32)
    with pytest.raises(M2TError):
        ours.TransBlock(num_heads=4)
    with pytest.raises(M2TError):
        m.mlp(torch.randn(1, 64, 64))
