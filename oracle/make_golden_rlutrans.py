"""Generate tests/golden/rlutrans_*.npz by running the REAL reference `util/rlutrans.py`.

Build container only (needs /root/reference); nothing at test or bench time imports this script.
The reference module is imported unmodified and loaded, through its own load_state_dict(strict=True),
with the seeded synthetic parameters of m2trans_b200.synthetic.synthetic_transblock_state_dict.

    python oracle/make_golden_rlutrans.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

REF = os.environ.get("M2T_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
torch.set_grad_enabled(False)

from util.rlutrans import TransBlock  # noqa: E402  (the reference)
from m2trans_b200.synthetic import synthetic_transblock_state_dict, synthetic_tokens  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
# (name, B, N, seed): 16 | N, 16 does not divide N (17 ragged chunks), the minimum N, a chunk longer than one key pass
CASES = [("rlutrans_b2_n256", 2, 256, 0), ("rlutrans_b3_n100", 3, 100, 1), ("rlutrans_b1_n16", 1, 16, 2),
         ("rlutrans_b1_n4500", 1, 4500, 3)]


def main():
    for name, b, n, seed in CASES:
        sd = synthetic_transblock_state_dict(seed)
        m = TransBlock().eval()
        m.load_state_dict(sd, strict=True)
        x = synthetic_tokens(b, n, seed=33 + seed)
        y = m(x)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), x=x.numpy(), y=y.numpy(), seed=np.int64(seed),
                            wsum=np.float64(sum(float(v.double().sum()) for v in sd.values())))
        print(name, tuple(y.shape), float(y.abs().max()))
    with open(os.path.join(OUT, "rlutrans_state_dict_manifest.txt"), "w") as f:
        for k, v in TransBlock().state_dict().items():
            f.write(f"{k} {tuple(v.shape)} {v.dtype}\n")


if __name__ == "__main__":
    main()
