"""BASELINE configs[4] on one GPU: MedCLIP image-embedding pass over x4 SR outputs ([B,3,512,512] -> 224x224 -> Swin-T ->
[B,512] -> logits), timed with CUDA events.  The 8-GPU figure of the config shards the batch by image (no collective).

    python tools/bench_clip.py [--batch 32] [--steps 10] [--warmup 3] [--profile]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from m2trans_b200.medclip_image import MedCLIPVisionModelViT, synthetic_state_dict  # noqa: E402

GFLOP_PER_IMAGE = 8.98          # 4.49 GMAC per 224x224 image


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)      # 256 images over 8 GPUs
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    tower = MedCLIPVisionModelViT()
    tower.load_state_dict(synthetic_state_dict(0), strict=False)
    tower = tower.cuda()
    x = torch.rand(args.batch, 3, args.size, args.size, device="cuda")
    text = torch.randn(512, device="cuda")
    for _ in range(args.warmup):
        tower.encode_image(x, text)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for i in range(args.steps):
        tower.encode_image(x, text)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    ms.sort()
    med = ms[len(ms) // 2]
    print(json.dumps({"metric": "medclip_image_pass", "batch": args.batch, "input": [3, args.size, args.size],
                      "ms_per_step": med, "ms_min": ms[0], "ms_max": ms[-1], "images_per_s": args.batch / med * 1e3,
                      "tflops": args.batch * GFLOP_PER_IMAGE / med, "dtype": "bf16"}))


if __name__ == "__main__":
    main()
