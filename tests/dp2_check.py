"""nn.DataParallel over two GPUs exactly as ref test.py:68-90 does on a multi-GPU box: replicas, scatter, gather."""
import sys
import types

import torch

sys.path.insert(0, ".")
from m2trans_b200.M2Trans_network import M2Trans  # noqa: E402
from m2trans_b200.synthetic import reference_checkpoint, synthetic_input, synthetic_state_dict  # noqa: E402
from oracle import m2trans_oracle as O  # noqa: E402

assert torch.cuda.device_count() >= 2
for scale in (4, 3):
    args = types.SimpleNamespace(scale=scale, rgb_range=1.0, colors=3, n_feats=64, n_blocks=8)
    model = torch.nn.DataParallel(M2Trans(args)).to("cuda:0")            # all visible devices
    model.load_state_dict(reference_checkpoint(scale, 0)["model_state_dict"], strict=True)
    model.eval()
    x = synthetic_input(4, 40, 56, seed=3)
    with torch.no_grad():
        y = model(x.to("cuda:0"))
        y2 = model(x.to("cuda:0"))
    ref = O.forward(synthetic_state_dict(scale, 0), x)
    single = torch.nn.DataParallel(M2Trans(args), device_ids=[0]).to("cuda:0")
    single.load_state_dict(reference_checkpoint(scale, 0)["model_state_dict"], strict=True)
    y1 = single.eval()(x.to("cuda:0"))
    print(f"x{scale}: devices {model.device_ids} out {tuple(y.shape)} on {y.device} | vs oracle PSNR {O.psnr(y.cpu(), ref):.1f} dB "
          f"max-abs {O.max_abs(y.cpu(), ref):.2e} | vs single-GPU {float((y - y1).abs().max()):.2e} | repeat {float((y - y2).abs().max()):.2e}")
