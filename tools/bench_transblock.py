"""Time rlutrans.TransBlock on one B200: B=16, N=4096, dim=64 (the synthetic check size of §8).
Prints one JSON line: us per forward, tokens/s, algorithmic GB/s (x in + y out) and GFLOP/s."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from m2trans_b200 import rlutrans as ours                       # noqa: E402
from m2trans_b200.synthetic import synthetic_tokens, synthetic_transblock_state_dict  # noqa: E402

B, N = 16, 4096
m = ours.TransBlock()
m.load_state_dict(synthetic_transblock_state_dict(0))
m = m.cuda().eval()
xs = [synthetic_tokens(B, N, seed=s).cuda() for s in range(3)]
flush = torch.empty(768 << 20, dtype=torch.uint8, device="cuda")
for i in range(5):
    m(xs[i % 3])
torch.cuda.synchronize()
ts = []
for i in range(20):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    m(xs[i % 3])
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort()
us = ts[len(ts) // 2]
n = N // 16
flop = B * N * 2 * (64 * 64 + 64 * 192 + 64 * 64 + 2 * 64 * 16) + B * 16 * 8 * (n * n * 8 * 2 * 2)
print(json.dumps({"kernel": "rlutrans.TransBlock B=16 N=4096", "us": round(us, 1), "tokens_per_s": B * N / us * 1e6,
                  "alg_GBps": B * N * 64 * 4 * 2 / us * 1e-3, "gflops": flop / us * 1e-3,
                  "note": "includes the Python wrapper (state_dict walk, workspace alloc); L2 flushed between runs"}))
