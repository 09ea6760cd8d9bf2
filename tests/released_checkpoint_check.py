"""Re-run the precision table of DESIGN.md section 3 on the reference's RELEASED checkpoints, if they are present
(checkpoints/model_x{2,3,4}.pt or $M2T_CHECKPOINTS; verified by git blob SHA-1).  Uses the oracle, so it
lives under tests/.   usage: python tests/released_checkpoint_check.py [dir]"""
import sys
import types

import torch

sys.path.insert(0, ".")
from m2trans_b200 import checkpoints as CK  # noqa: E402


def table(directory=None, sizes=None, kinds=("uniform", "speckle", "flat")):
    from m2trans_b200.M2Trans_network import M2Trans
    from m2trans_b200.synthetic import synthetic_input
    from oracle import m2trans_oracle as O
    found = CK.find_released(directory)
    rows = []
    for scale, path in sorted(found.items()):
        sd = CK.load_model_state_dict(path)
        m = M2Trans(types.SimpleNamespace(scale=scale, rgb_range=1.0, colors=3, n_feats=64, n_blocks=8)).cuda()
        m.load_state_dict(sd, strict=True)
        for (h, w) in sizes or [(64, 64), (96, 120)]:
            for kind in kinds:
                x = synthetic_input(1, h, w, seed=33, kind=kind)
                y = m(x.cuda()).cpu()
                ref = O.forward(sd, x, scale=scale)
                rows.append((scale, h, w, kind, O.psnr(y, ref), O.max_abs(y, ref), float((ref <= 0).float().mean() + (ref >= 1).float().mean())))
    return found, rows


if __name__ == "__main__":
    torch.set_num_threads(16)
    found, rows = table(sys.argv[1] if len(sys.argv) > 1 else None)
    if not found:
        print("no released checkpoint found (expected checkpoints/model_x{2,3,4}.pt with blob SHA-1", CK.RELEASED_SHA1, ")")
    for scale, h, w, kind, p, e, clamped in rows:
        print(f"x{scale} {h}x{w} {kind:8s}: PSNR {p:.1f} dB  max-abs {e:.2e}  clamped {100 * clamped:.1f} %")
