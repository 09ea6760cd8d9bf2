"""Development aid: phase timing (SM clocks) of CTA 0 of the tcgen05 attention kernel.
Needs a library built with M2T_TIMING=1 (python -m m2trans_b200.build --force)."""
import ctypes as C
import sys
import types

import torch

sys.path.insert(0, ".")
from m2trans_b200 import _lib  # noqa: E402

lib = _lib.load()


def show(tag):
    torch.cuda.synchronize()
    buf = (C.c_longlong * 64)()
    _lib.check(lib.m2t_debug_attn_timing(buf), "timing")
    t = list(buf)
    t0 = t[6]
    print(f"[{tag}] prologue {t[7] - t0} clk")
    for i in range(6):
        r = t[8 * i: 8 * i + 6]
        if r[5] == 0 or r[5] < t0:
            break
        print(f"   pair {i}: start+{r[0] - t0:7d} | wait S {r[1] - r[0]:6d} | softmax {r[2] - r[1]:6d} | glue-issue {r[3] - r[2]:6d}"
              f" | wait O {r[4] - r[3]:6d} | epilogue {r[5] - r[4]:6d} | end+{r[5] - t0:7d}")


for Cc, (B, h, w) in ((16, (16, 128, 128)), (64, (16, 64, 64)), (256, (16, 32, 32))):
    qkv = torch.randn(B, h, w, 3 * Cc, device="cuda").half()
    relf = torch.randn(20, Cc // 2, device="cuda")
    relx = torch.zeros(32, Cc, dtype=torch.float16, device="cuda")
    o = torch.empty(B, h, w, Cc, dtype=torch.float16, device="cuda")
    for _ in range(3):
        _lib.check(lib.m2t_stage_attn(0, Cc, qkv.data_ptr(), relf.data_ptr(), relx.data_ptr(), o.data_ptr(), B, h, w, None), "attn")
    show(f"plain attention C={Cc} cfg2")

from m2trans_b200.M2Trans_network import M2Trans  # noqa: E402
from m2trans_b200.synthetic import synthetic_input, synthetic_state_dict  # noqa: E402

for nb in (1,):
    args = types.SimpleNamespace(scale=4, rgb_range=1.0, colors=3, n_feats=64, n_blocks=nb)
    m = M2Trans(args).cuda()
    m.load_state_dict(synthetic_state_dict(4, 0, n_blocks=nb))
    x = synthetic_input(16, 128, 128).cuda()
    for _ in range(3):
        m(x)
    show("fused, last launch = branch 4 (C=256, no Tnext), cfg2")
