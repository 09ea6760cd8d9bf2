"""BASELINE configs[4]: MedCLIP image-embedding pass over x4 SR outputs ([B,3,512,512] -> 224x224 -> Swin-T -> [B,512] ->
logits), timed with CUDA events.  The pass shards by image with no collective on the data path: under torchrun every rank
takes `--batch` images (weak scaling, as bench.py), the timed region is bracketed by barriers and the time is the max over
ranks.

    python tools/bench_clip.py [--batch 32] [--steps 10] [--warmup 3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/bench_clip.py --batch 32          # the config's 256 images over 8 GPUs
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
    ap.add_argument("--batch", type=int, default=32, help="images per GPU (256 images over 8 GPUs)")
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-graph", action="store_true", help="eager launches even for small inputs")
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tower = MedCLIPVisionModelViT()
    tower.load_state_dict(synthetic_state_dict(0), strict=False)
    tower = tower.cuda()
    tower.cuda_graph = not args.no_graph
    x = torch.rand(args.batch, 3, args.size, args.size, device="cuda", generator=torch.Generator("cuda").manual_seed(rank))
    text = torch.randn(512, device="cuda")
    for _ in range(args.warmup):
        tower.encode_image(x, text)

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    fence()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for i in range(args.steps):
        tower.encode_image(x, text)
        ev[i + 1].record()
    fence()
    ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps))
    total = torch.tensor([ev[0].elapsed_time(ev[args.steps]) / args.steps], device="cuda")
    if world > 1:
        dist.all_reduce(total, op=dist.ReduceOp.MAX)
    mean_ms = float(total)
    if rank == 0:
        print(json.dumps({"metric": "medclip_image_pass", "n_gpus": world, "batch_per_gpu": args.batch,
                          "input": [3, args.size, args.size], "ms_per_step": mean_ms, "ms_median_rank0": ms[len(ms) // 2],
                          "ms_min_rank0": ms[0], "ms_max_rank0": ms[-1], "images_per_s": world * args.batch / mean_ms * 1e3,
                          "tflops": world * args.batch * GFLOP_PER_IMAGE / mean_ms, "scaling": "weak", "dtype": "bf16", "cuda_graph": bool(tower._graphs),
                          "timing": "CUDA events, mean over the timed steps, max over ranks"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
