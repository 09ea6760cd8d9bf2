"""Development aid: phase timing (SM clocks) of CTA 0 of the 256-channel qkv GEMM.  Needs M2T_TIMING=1 build."""
import ctypes as C
import sys
import types

import torch

sys.path.insert(0, ".")
from m2trans_b200 import _lib  # noqa: E402
from m2trans_b200.M2Trans_network import M2Trans  # noqa: E402
from m2trans_b200.synthetic import synthetic_input, synthetic_state_dict  # noqa: E402

lib = _lib.load()
m = M2Trans(types.SimpleNamespace(scale=4, rgb_range=1.0, colors=3, n_feats=64, n_blocks=1)).cuda()
m.cuda_graph = False
m.load_state_dict(synthetic_state_dict(4, 0, n_blocks=1))
x = synthetic_input(16, 128, 128).cuda()
for _ in range(3):
    m(x)
torch.cuda.synchronize()
buf = (C.c_longlong * 448)()
_lib.check(lib.m2t_debug_attn_timing(buf), "timing")
t = list(buf)[384:448]
t0 = t[0]
print(f"prologue {t[1] - t0} | weight slab landed +{t[2] - t0}")
for i in range(6):
    r = t[8 + 8 * i: 8 + 8 * i + 5]
    if r[4] == 0:
        break
    print(f"tile {i}: mma warp at +{r[0] - t0:6d} | A landed +{r[1] - t0:6d} | MMAs issued +{r[2] - t0:6d} | acc ready +{r[3] - t0:6d} | stored +{r[4] - t0:6d}")
