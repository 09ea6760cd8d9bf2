"""Development aid: host-side cost of one M2Trans.forward call (enqueue only, no synchronisation)."""
import cProfile
import pstats
import sys
import time
import types

import torch

sys.path.insert(0, ".")
from m2trans_b200.M2Trans_network import M2Trans  # noqa: E402
from m2trans_b200.synthetic import synthetic_input, synthetic_state_dict  # noqa: E402

m = M2Trans(types.SimpleNamespace(scale=4, rgb_range=1.0, colors=3, n_feats=64, n_blocks=8)).cuda()
m.load_state_dict(synthetic_state_dict(4, 0))
x = synthetic_input(16, 128, 128).cuda()
for _ in range(5):
    m(x)
torch.cuda.synchronize()
n = 50
t0 = time.perf_counter()
for _ in range(n):
    m(x)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3 * (t1 - t0) / n:.3f} ms/forward; with drain {1e3 * (t2 - t0) / n:.3f} ms/forward")
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    m(x)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
