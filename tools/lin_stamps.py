"""clock64 stamps of the tower's Linear kernel (library built with M2T_TIMING=1): one epilogue-0 launch, CTA 0, tiles 2..9."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from m2trans_b200 import _lib  # noqa: E402

lib = _lib.load()
M, N, K = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (100352, 384, 96)
a = torch.randn(M, K, device="cuda").bfloat16()
w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
bias = torch.randn(N, device="cuda")
out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
s = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    _lib.check(lib.m2t_clip_stage_linear(0, a.data_ptr(), w.data_ptr(), bias.data_ptr(), out.data_ptr(), M, N, K, s), "lin")
torch.cuda.synchronize()
buf = (C.c_longlong * 128)()
_lib.check(lib.m2t_debug_lin_timing(buf), "timing")
v = list(buf)
t0 = min(x for x in v if x > 0)
names = ["ring free", "mma: before acc wait", "mma: acc free", "mma: K0 landed", "mma: issued", "epi: before ready wait", "epi: acc ready",
         "epi: buffer free", "epi: staged", "epi: store issued", "epi: released"]
print(f"M {M} N {N} K {K}; cycles relative to the first stamp")
for t in range(8):
    row = v[16 * t:16 * t + 11]
    print(f"tile {t + 2}: " + "  ".join(f"{n} {x - t0 if x else -1}" for n, x in zip(names, row)))
