"""Development aid: phase timing (SM clocks) of CTA 0 of the tcgen05 ff conv.  Needs M2T_TIMING=1 build."""
import ctypes as C
import sys
import types

import torch

sys.path.insert(0, ".")
from m2trans_b200 import _lib  # noqa: E402
from m2trans_b200.M2Trans_network import M2Trans  # noqa: E402
from m2trans_b200.synthetic import synthetic_input, synthetic_state_dict  # noqa: E402

lib = _lib.load()
# usage: python tools/conv_timing.py [cfg2|cfg3|cfg4]   (cfg3 = x3: the precise-mode kernel ffconv_umma_kernel<W2>)
WORK = {"cfg2": (4, 16, 128, 128), "cfg3": (3, 32, 200, 266), "cfg4": (4, 64, 270, 480)}
scale, B, H, W = WORK[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
args = types.SimpleNamespace(scale=scale, rgb_range=1.0, colors=3, n_feats=64, n_blocks=2)
m = M2Trans(args).cuda()
m.cuda_graph = False
m.load_state_dict(synthetic_state_dict(scale, 0, n_blocks=2))
x = synthetic_input(B, H, W).cuda()
for _ in range(3):
    m(x)
torch.cuda.synchronize()
buf = (C.c_longlong * 448)()
_lib.check(lib.m2t_debug_attn_timing(buf), "timing")
t = list(buf)[320:384]
t0 = t[0]
for i in range(8):
    r = t[8 * i: 8 * i + 8]
    if r[4] == 0:
        break
    print(f"tile {i}: epi start+{r[0] - t0:6d} | issue loads {r[1] - r[0]:5d} | wait acc {r[2] - r[1]:5d} | stage {r[3] - r[2]:5d} | apply {r[4] - r[3]:5d}"
          f" || mma: at+{r[5] - t0:6d} wait-free {r[6] - r[5]:5d} wait-tile {r[7] - r[6]:5d}")
