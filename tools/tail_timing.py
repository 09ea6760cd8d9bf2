"""Development aid: phase timing (SM clocks) of CTA 0 of the fused tail kernel.
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
args = types.SimpleNamespace(scale=4, rgb_range=1.0, colors=3, n_feats=64, n_blocks=1)
m = M2Trans(args).cuda()
m.cuda_graph = False
m.load_state_dict(synthetic_state_dict(4, 0, n_blocks=1))
x = synthetic_input(16, 128, 128).cuda()
for _ in range(3):
    m(x)
torch.cuda.synchronize()
buf = (C.c_longlong * 448)()
_lib.check(lib.m2t_debug_attn_timing(buf), "timing")
t = list(buf)[256:320]
prev_end = None
for i in range(8):
    r = t[8 * i: 8 * i + 6]
    if r[5] == 0:
        break
    gap = "" if prev_end is None else f"gap {r[0] - prev_end:5d} | "
    print(f"tile {i}: {gap}gelu-epilogue {r[1] - r[0]:6d} | publish {r[2] - r[1]:5d} | wait conv {r[3] - r[2]:6d} | planes {r[4] - r[3]:5d}"
          f" | gather {r[5] - r[4]:5d} | total {r[5] - r[0]:6d} clk")
    prev_end = r[5]
