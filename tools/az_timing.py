"""Development aid: phase timing (SM clocks) of CTA 0 of attn_z (branches 2-4), one record per branch.
Needs a library built with M2T_TIMING=1 (M2T_TIMING=1 python -m m2trans_b200.build --force).
usage: python tools/az_timing.py [cfg2|cfg1|cfg4]"""
import ctypes as C
import sys
import types

import torch

sys.path.insert(0, ".")
from m2trans_b200 import _lib  # noqa: E402
from m2trans_b200.M2Trans_network import M2Trans  # noqa: E402
from m2trans_b200.synthetic import synthetic_input, synthetic_state_dict  # noqa: E402

WORK = {"cfg1": (2, 1, 64, 64), "cfg2": (4, 16, 128, 128), "cfg3": (3, 32, 200, 266), "cfg4": (4, 64, 270, 480)}
scale, B, H, W = WORK[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
lib = _lib.load()
m = M2Trans(types.SimpleNamespace(scale=scale, rgb_range=1.0, colors=3, n_feats=64, n_blocks=1, cuda_graph=False)).cuda()
m.load_state_dict(synthetic_state_dict(scale, 0, n_blocks=1))
x = synthetic_input(B, H, W).cuda()
for _ in range(3):
    m(x)
torch.cuda.synchronize()
buf = (C.c_longlong * 192)()
_lib.check(lib.m2t_debug_az_timing(buf), "timing")
names = ["wait A", "cvt A", "wait S", "softmax", "wait PZ", "cvt PZ", "wait O", "glue"]
for br in range(3):
    t = list(buf)[64 * br: 64 * br + 64]
    if t[60] == 0:
        continue
    print(f"branch {br + 2}: prologue {t[61] - t[60]} clk")
    for i in range(6):
        r = t[10 * i: 10 * i + 9]
        if r[8] == 0 or r[8] < t[60]:
            break
        print(f"   pair {i}: start+{r[0] - t[60]:7d} | " + " | ".join(f"{n} {r[k + 1] - r[k]:6d}" for k, n in enumerate(names)) + f" | total {r[8] - r[0]:6d}")
