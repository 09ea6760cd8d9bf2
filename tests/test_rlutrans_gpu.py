"""rlutrans.TransBlock on the B200 through the C ABI (m2t_transblock_forward), against the
reference-generated fixtures and the CPU oracle.  fp32 arithmetic on both sides: the bar is max-abs <= 2e-5
(outputs are O(5)), i.e. summation-order noise only."""
import os

import numpy as np
import pytest
import torch

from oracle import rlutrans_oracle as R
from m2trans_b200 import rlutrans as ours
from m2trans_b200._lib import M2TError
from m2trans_b200.synthetic import synthetic_tokens, synthetic_transblock_state_dict

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
TOL = 2e-5
CASES = [("rlutrans_b2_n256", 0), ("rlutrans_b3_n100", 1), ("rlutrans_b1_n16", 2), ("rlutrans_b1_n4500", 3)]


def _block(seed):
    m = ours.TransBlock()
    m.load_state_dict(synthetic_transblock_state_dict(seed), strict=True)
    return m.cuda().eval()


@pytest.mark.parametrize("name,seed", CASES)
def test_golden(golden_dir, name, seed):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    y = _block(seed)(torch.from_numpy(z["x"]).cuda()).cpu().numpy()
    err = float(np.abs(y - z["y"]).max())
    print(f"{name}: max-abs {err:.2e}")
    assert err <= TOL


@pytest.mark.parametrize("b,n", [(16, 4096), (1, 17), (2, 1000), (5, 31)])
def test_against_oracle(b, n):
    sd = synthetic_transblock_state_dict(7)
    x = synthetic_tokens(b, n, seed=5)
    y = _block(7)(x.cuda()).cpu()
    ref = R.transblock(sd, x)
    err = float((y - ref).abs().max())
    print(f"B={b} N={n}: max-abs {err:.2e}")
    assert err <= TOL


def test_chunks_are_independent_on_device():
    m = _block(0)
    x = synthetic_tokens(1, 160, seed=9).cuda()
    y0 = m(x)
    x2 = x.clone()
    x2[:, 30:40] += 1.0
    d = (m(x2) - y0).abs().amax(dim=(0, 2)).cpu()
    assert float(d[:30].max()) == 0.0 and float(d[40:].max()) == 0.0 and float(d[30:40].max()) > 1e-3


def test_non_contiguous_input_and_errors():
    m = _block(1)
    xb = synthetic_tokens(2, 64, seed=3).cuda()
    xt = xb.transpose(0, 1).contiguous().transpose(0, 1)          # same values, non-contiguous
    assert not xt.is_contiguous()
    assert torch.equal(m(xt), m(xb))
    with pytest.raises(M2TError):
        m(torch.randn(1, 15, 64, device="cuda"))                  # the reference raises too (chunk length 0)
    with pytest.raises(M2TError):
        m(torch.randn(1, 64, 64, device="cuda").half())
    with pytest.raises(M2TError):
        m(torch.randn(1, 64, 32, device="cuda"))


@pytest.mark.parametrize("b,n", [(2, 256), (1, 17), (3, 1000)])
def test_submodules_run_on_their_own(b, n):
    """Mlp.forward (ref :21-27) and EffAttention.forward (ref :47-66) are public classes of the reference file: called
    directly they run the block's kernels in a standalone mode and match the oracle's restatement of each."""
    sd = synthetic_transblock_state_dict(4)
    m = _block(4)
    x = synthetic_tokens(b, n, seed=11)
    ya = m.atten(x.cuda()).cpu()
    ym = m.mlp(x.cuda()).cpu()
    ea = float((ya - R.eff_attention(sd, x)).abs().max())
    em = float((ym - R.mlp(sd, x)).abs().max())
    print(f"B={b} N={n}: EffAttention max-abs {ea:.2e}, Mlp max-abs {em:.2e}")
    assert ea <= TOL and em <= TOL
    assert ym.shape == x.shape and m.mlp(x.cuda().reshape(-1, 64)).shape == (b * n, 64)     # Mlp takes any [..., 64]
    with pytest.raises(M2TError):
        m.atten(torch.randn(1, 15, 64, device="cuda"))
    with pytest.raises(M2TError):
        m.mlp(torch.randn(4, 32, device="cuda"))
