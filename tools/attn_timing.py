"""Development aid: phase timing (SM clocks) of CTA 0 of the tcgen05 attention kernel, one record per branch.
Needs a library built with M2T_TIMING=1 (M2T_TIMING=1 python -m m2trans_b200.build --force)."""
import ctypes as C
import sys
import types

import torch

sys.path.insert(0, ".")
from m2trans_b200 import _lib  # noqa: E402
from m2trans_b200.M2Trans_network import M2Trans  # noqa: E402
from m2trans_b200.synthetic import synthetic_input, synthetic_state_dict  # noqa: E402

lib = _lib.load()
lib.m2t_debug_attn_timing.argtypes = [C.POINTER(C.c_longlong)]


def show(tag):
    torch.cuda.synchronize()
    buf = (C.c_longlong * 448)()
    _lib.check(lib.m2t_debug_attn_timing(buf), "timing")
    for br in range(4):
        t = list(buf)[64 * br: 64 * br + 64]
        t0 = t[6]
        if t0 == 0:
            continue
        print(f"[{tag}] branch {br + 1}: prologue {t[7] - t0} clk")
        for i in range(6):
            r = t[8 * i: 8 * i + 6]
            if r[5] == 0 or r[5] < t0:
                break
            print(f"   pair {i}: start+{r[0] - t0:7d} | wait S {r[1] - r[0]:6d} | softmax {r[2] - r[1]:6d} | glue-issue {r[3] - r[2]:6d}"
                  f" | wait O {r[4] - r[3]:6d} | epilogue {r[5] - r[4]:6d} | end+{r[5] - t0:7d}")


args = types.SimpleNamespace(scale=4, rgb_range=1.0, colors=3, n_feats=64, n_blocks=1)
m = M2Trans(args).cuda()
m.load_state_dict(synthetic_state_dict(4, 0, n_blocks=1))
x = synthetic_input(16, 128, 128).cuda()
for _ in range(3):
    m(x)
show("fused forward, cfg2")
