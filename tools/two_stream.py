"""Experiment: split the batch across 2/4 concurrent streams (independent images) to fill latency bubbles."""
import sys, types, time
import torch
sys.path.insert(0, ".")
from m2trans_b200.M2Trans_network import M2Trans
from m2trans_b200.synthetic import synthetic_input, synthetic_state_dict

args = types.SimpleNamespace(scale=4, rgb_range=1.0, colors=3, n_feats=64, n_blocks=8)
sd = synthetic_state_dict(4, 0)
x = synthetic_input(16, 128, 128).cuda()

def bench(nsplit, iters=20):
    models = []
    for _ in range(nsplit):
        m = M2Trans(args).cuda(); m.load_state_dict(sd); models.append(m)
    streams = [torch.cuda.Stream() for _ in range(nsplit)]
    parts = x.chunk(nsplit)
    def step():
        ev = torch.cuda.Event(); ev.record()
        outs = []
        for m, s, p in zip(models, streams, parts):
            s.wait_event(ev)
            with torch.cuda.stream(s):
                outs.append(m(p))
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        return outs
    for _ in range(5): step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): step()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters

for n in (1, 2, 4):
    print(f"streams={n}: {bench(n):.3f} ms per 16-image batch")
