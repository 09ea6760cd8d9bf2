"""Development aid: parity of the default path against the CPU oracle over several seeded checkpoints and inputs.
usage: python tests/parity_sweep.py [scale] [H] [W] [n_seeds]"""
import sys
import types

import torch

sys.path.insert(0, ".")
from m2trans_b200.M2Trans_network import M2Trans  # noqa: E402
from m2trans_b200.synthetic import synthetic_input, synthetic_state_dict  # noqa: E402
from oracle import m2trans_oracle as O  # noqa: E402

scale = int(sys.argv[1]) if len(sys.argv) > 1 else 3
H = int(sys.argv[2]) if len(sys.argv) > 2 else 200
W = int(sys.argv[3]) if len(sys.argv) > 3 else 266
n = int(sys.argv[4]) if len(sys.argv) > 4 else 5
torch.set_num_threads(16)
worst = (0.0, 1e9)
for seed in range(n):
    for kind in ("uniform", "speckle"):
        sd = synthetic_state_dict(scale, seed)
        m = M2Trans(types.SimpleNamespace(scale=scale, rgb_range=1.0, colors=3, n_feats=64, n_blocks=8)).cuda()
        m.load_state_dict(sd)
        x = synthetic_input(1, H, W, seed=100 + seed, kind=kind)
        y = m(x.cuda()).cpu()
        ref = O.forward(sd, x)
        p, e = O.psnr(y, ref), O.max_abs(y, ref)
        worst = (max(worst[0], e), min(worst[1], p))
        print(f"x{scale} {H}x{W} seed {seed} {kind:8s}: PSNR {p:.1f} dB  max-abs {e:.2e}", flush=True)
print(f"worst: max-abs {worst[0]:.2e}, PSNR {worst[1]:.1f} dB")
