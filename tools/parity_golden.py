"""Parity of the engine against every reference-made full-forward fixture under tests/golden, per kernel variant.
usage: python tools/parity_golden.py [substring]"""
import glob
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, ".")
from m2trans_b200 import _lib  # noqa: E402
from m2trans_b200.M2Trans_network import M2Trans  # noqa: E402
from m2trans_b200.synthetic import reference_checkpoint  # noqa: E402

VARIANTS = [("default", 0), ("precise", _lib.VAR_PRECISE_ON), ("fast", _lib.VAR_PRECISE_OFF), ("split_qkv", _lib.VAR_SPLIT_QKV),
            ("simt", 0xF)]
sub = sys.argv[1] if len(sys.argv) > 1 else ""
for path in sorted(glob.glob("tests/golden/*.npz")):
    name = os.path.basename(path)[:-4]
    g = np.load(path)
    if "scale" not in g.files or "x" not in g.files or sub not in name:
        continue
    scale, seed = int(g["scale"]), int(g["seed"])
    kw = {"qkv_gain": float(g["qkv_gain"])}
    if "out_gain" in g.files:
        kw.update(out_gain=float(g["out_gain"]), out_shift=float(g["out_shift"]))
    ref = torch.from_numpy(g["y"]).double()
    x = torch.from_numpy(g["x"]).cuda()
    inside = float(((ref > 0) & (ref < 1)).float().mean())
    line = f"{name:32s} x{scale} unclamped {100 * inside:5.1f}% |"
    for vn, var in VARIANTS:
        args = types.SimpleNamespace(scale=scale, rgb_range=1.0, colors=3, n_feats=64, n_blocks=8, kernel_variant=var)
        m = torch.nn.DataParallel(M2Trans(args), device_ids=[0]).cuda()
        m.load_state_dict(reference_checkpoint(scale, seed, **kw)["model_state_dict"], strict=True)
        y = m.eval()(x).double().cpu()
        mse = float(((y - ref) ** 2).mean())
        line += f" {vn} {10 * np.log10(1 / mse):5.1f} dB {float((y - ref).abs().max()):.2e} |"
    print(line, flush=True)
