"""Device PSNR/SSIM pass on B x [3,1080,1920] SR/HR pairs: ms and achieved HBM GB/s (24 B per pixel)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from m2trans_b200.metrics import psnr_ssim  # noqa: E402

B, H, W = int(sys.argv[1]) if len(sys.argv) > 1 else 16, 1080, 1920
hr = torch.rand(B, 3, H, W, device="cuda")
sr = (hr + 0.02 * torch.randn_like(hr)).clamp(0, 1)
for _ in range(3):
    psnr_ssim(sr, hr, 4)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
ev[0].record()
for i in range(10):
    psnr_ssim(sr, hr, 4)
    ev[i + 1].record()
torch.cuda.synchronize()
ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(10))[5]
print(json.dumps({"metric": "psnr_ssim_pass", "batch": B, "frame": [3, H, W], "ms": ms, "mpix_per_s": B * H * W / ms / 1e3,
                  "hbm_gbs": B * H * W * 24 / ms / 1e6}))
