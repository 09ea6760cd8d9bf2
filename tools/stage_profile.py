"""Development aid: per-kernel times of one eager forward measured with CUDA events between launches (warm caches), then
the time of the replayed (graph) forward.
usage: python tools/stage_profile.py [cfg2|cfg1|cfg3|cfg4|frame2|frame4]"""
import sys
import types

import torch

sys.path.insert(0, ".")
from m2trans_b200.M2Trans_network import M2Trans  # noqa: E402
from m2trans_b200.synthetic import synthetic_input, synthetic_state_dict  # noqa: E402

WORK = {"cfg1": (2, 1, 64, 64), "cfg2": (4, 16, 128, 128), "cfg3": (3, 32, 200, 266), "cfg4": (4, 64, 270, 480),
        "frame2": (2, 1, 300, 400), "frame4": (4, 1, 152, 200)}       # one evaluation frame, the way test.py feeds them
scale, B, H, W = WORK[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
m = M2Trans(types.SimpleNamespace(scale=scale, rgb_range=1.0, colors=3, n_feats=64, n_blocks=8)).cuda()
m.load_state_dict(synthetic_state_dict(scale, 0))
x = synthetic_input(B, H, W).cuda()
for _ in range(3):
    m(x)
torch.cuda.synchronize()
for _ in range(2):
    txt = m.profile_forward(x)
lines = txt.strip().splitlines()
print("\n".join(sorted(lines[:-1], key=lambda l: -float(l.split()[0]))))
print(lines[-1])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    m(x)
e1.record()
torch.cuda.synchronize()
print(f"replayed forward: {e0.elapsed_time(e1) / 20:.4f} ms")
