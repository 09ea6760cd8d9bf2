"""CPU fp32 oracle for rlutrans.TransBlock.  TEST INFRASTRUCTURE ONLY.

Restatement in plain torch fp32 of `util/rlutrans.py` of the reference (eezkni/M2Trans): TransBlock.forward
:82-87, EffAttention.forward :46-66, Mlp.forward :20-27.  It works on a plain state dict and never on the
reference's classes.  Only `tests/` may import it; the product package `m2trans_b200` never does.

Parity pin: `oracle/make_golden_rlutrans.py` runs the real reference module in the build container on seeded
inputs and commits the results as tests/golden/rlutrans_*.npz; tests/test_oracle_golden.py holds this file to
those vectors.  The reference has no test that touches this module (nothing imports it), so the generated
fixtures are the only pin.

All `ref:` citations are into /root/reference/util/rlutrans.py.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

NUM_HEADS = 8   # ref :73 default; TransBlock passes it through (:77)


def block_diagonal_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, chunk: int, scale: float) -> torch.Tensor:
    """q, k, v [B, heads, N, hd] -> [B, N, heads*hd]; softmax restricted to runs of `chunk` tokens (ref :53-64).

    Written as ONE masked attention instead of the reference's per-chunk loop: token i attends to token j iff
    i // chunk == j // chunk.  Masked logits are -inf, so their probabilities are exactly zero."""
    b, h, n, hd = q.shape
    ids = torch.arange(n) // chunk
    mask = ids[:, None] == ids[None, :]
    logits = torch.einsum("bhid,bhjd->bhij", q, k) * scale
    logits = logits.masked_fill(~mask, float("-inf"))
    p = torch.softmax(logits, dim=-1)
    o = torch.einsum("bhij,bhjd->bhid", p, v)
    return o.permute(0, 2, 1, 3).reshape(b, n, h * hd)


def eff_attention(sd: Dict[str, torch.Tensor], x: torch.Tensor, prefix: str = "atten.") -> torch.Tensor:
    """ref :46-66.  reduce and qkv have no bias (TransBlock constructs them with qkv_bias=False, :77)."""
    b, n, _ = x.shape
    if n < 16:
        raise ValueError("N < 16: the reference's chunk length N // 16 is 0 and torch.split raises (ref :53)")
    r = F.linear(x, sd[prefix + "reduce.weight"])
    c = r.shape[-1]
    hd = c // NUM_HEADS
    qkv = F.linear(r, sd[prefix + "qkv.weight"]).reshape(b, n, 3, NUM_HEADS, hd).permute(2, 0, 3, 1, 4)
    o = block_diagonal_attention(qkv[0], qkv[1], qkv[2], n // 16, hd ** -0.5)
    return F.linear(o, sd[prefix + "proj.weight"], sd[prefix + "proj.bias"])


def mlp(sd: Dict[str, torch.Tensor], x: torch.Tensor, prefix: str = "mlp.") -> torch.Tensor:
    """ref :20-27 with ReLU and p = 0 dropout."""
    return F.linear(torch.relu(F.linear(x, sd[prefix + "fc1.weight"], sd[prefix + "fc1.bias"])),
                    sd[prefix + "fc2.weight"], sd[prefix + "fc2.bias"])


def transblock(sd: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """ref :82-87.  x [B, N, dim] fp32."""
    dim = x.shape[-1]
    x = x + eff_attention(sd, F.layer_norm(x, (dim,), sd["norm1.weight"], sd["norm1.bias"], 1e-5))
    return x + mlp(sd, F.layer_norm(x, (dim,), sd["norm2.weight"], sd["norm2.bias"], 1e-5))
