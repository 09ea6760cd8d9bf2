#!/usr/bin/env python
"""bench.py -- SR output megapixels/s of the M2Trans forward on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg1..cfg5]

One "step" = one forward of the hot path over one batch of synthetic LR frames.  The default workload is BASELINE.json
configs[1] (the config the metric is quoted on): M2Trans x4, model_x4 (seeded synthetic checkpoint in the reference's
format; the released file is not available offline), batch 16 of 3x128x128.

Workloads and multi-GPU (one process per GPU under torchrun; images are independent, so there is no collective on the
data path: torch.distributed / NCCL is used for the barrier and the max-reduce of the timings only):
  cfg1 / cfg2   every rank runs the named batch (WEAK scaling; cfg2 is what the driver's 1/2/4/8 sweep runs)
  cfg3 / cfg4   the NAMED batch (32 / 64 frames) is split over the ranks with sharding.shard_range (STRONG scaling:
                32 -> 4 frames per GPU at N = 8, 64 -> 8 per GPU); value = all frames / max-over-ranks time
  cfg5          MedCLIP image-embedding pass, 256 SR outputs split over the ranks (STRONG), images/s

Printed JSON line (rank 0):
  value         whole-job throughput with inputs resident in HBM (CUDA events, max over ranks)
  e2e           the same metric through the public nn.Module call with HOST buffers, H2D and D2H inside the timed
                region: pinned fp32 LR batch -> device, model(x), SR batch -> uint8 HWC on the device (what writing the
                images does) -> pinned host.  K pipelined steps on three streams, CUDA-event timed (`value`), and the
                host wall clock of the same region (`wall_ms_per_step`).
  e2e_fp32      the same with the fp32 SR tensor copied to the host (12 B per output pixel; at N = 8 the eight 50 MB
                copies per step saturate the host's memory system, which is why the uint8 form is the headline)
  e2e_eval      the reference's eval loop (test.py:87-116) on the device: uint8 LR + HR batches H2D, loader conversion,
                forward, Y-PSNR / SSIM, D2H of the two scalars
  roofline      the dominant kernel (largest summed time per forward in `kernels`), timed alone with CUDA events over
                back-to-back launches on rotating buffers, against MEASURED_PEAKS.json; `traffic` = DRAM bytes per launch
                from the committed ncu capture of this command (profiles/r02_ncu_traffic.json)
  kernels       per kernel class: launches per forward, summed microseconds inside one eager forward (CUDA event after
                every launch: includes the ~3 us launch gap, warm caches), algorithmic FLOPs and bytes, fraction of the
                peak that bounds it; largest_launch names the longest single launch
  cpu_baseline  the CPU oracle (a torch fp32 port of the reference forward) on the host cores, bounded sample
`--impl reference` runs that CPU port alone (rank 0) with the same `config`, metric and unit.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (scale, batch, H, W, scaling)   -- BASELINE.json configs[0..3]
    "cfg1": (2, 1, 64, 64, "weak"),
    "cfg2": (4, 16, 128, 128, "weak"),
    "cfg3": (3, 32, 200, 266, "strong"),
    "cfg4": (4, 64, 270, 480, "strong"),
}
FLOP_PER_PX = {2: 1299328, 3: 1357568, 4: 1471872}     # algorithmic FLOPs per padded LR px
TAIL_FLOP_PER_PX = {2: 46592, 3: 104832, 4: 219136}

# Algorithmic work per padded LR pixel of every kernel class.
# bytes = the HBM traffic the dataflow needs (fp32 residual stream, fp16 branch tensors), not what ncu counts.
# key = substring of the kernel name; value = (label, bound, FLOP/px per launch, B/px per launch)
KERNELS = [
    ("head_conv", "head 3x3 conv + IN sums", "hbm", 3456, 12 + 256),
    ("branch_prep_all", "InstanceNorm apply + branch inputs", "hbm", 0, 256 + 128),
    ("attn16_qkv", "branch 1: qkv + attention + glue", "hbm", 1536 + 6400, 32 + 32 + 64),
    ("attn_z_kernelILi64E", "branch 2: qkv + attention + glue (attn_z<64>)", "hbm", 6144 + 6400, 32 + 32 + 64),
    ("attn_z_kernelILi256E", "branches 3-4: qkv + attention + glue (attn_z<256>)", "tensor", 24576 + 6400, 32 + 32 + 32),
    ("ffconv_umma", "ff 3x3 conv + residual + IN sums", "hbm", 73728, 128 + 256 + 256),
    ("tail_up_umma", "tail 1x1 conv + PixelShuffle + GELU", "hbm", 2 * 64 * 256, 128 + 512),
    ("tail_strip", "tail last stage: 1x1 + PS + GELU + 3x3 + clamp", "fma", 0, 0),
    ("tail_fused", "tail last stage (tiled variant)", "fma", 0, 0),
    ("tail_out_umma", "tail 3x3 conv + clamp", "hbm", 0, 0),
    ("reflect_border", "tail border ring", "hbm", 0, 0),
]


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "tflops_burst": p["bf16_tflops"], "tflops_sustained": p["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


def workload_config(name):
    """The `config` object both arms print (identical on purpose: the driver compares them)."""
    if name == "cfg5":
        return {"workload": "cfg5: MedCLIP image-embedding pass over x4 SR outputs, 256x3x512x512 -> 224x224 -> Swin-T -> "
                            "[256,512] -> cosine logits, images split over the GPUs, synthetic tower weights"}
    scale, B, H, W, scaling = WORKLOADS[name]
    per = "per GPU" if scaling == "weak" else "in total, split over the GPUs by image"
    return {"workload": f"{name}: M2Trans x{scale}, {B}x3x{H}x{W} LR {per}, synthetic model_x{scale} checkpoint seed 0"}


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md recipe).  Polls NVML every 10 ms (the
    timed region of the default run lasts ~40 ms; one nvidia-smi call takes longer than that) and falls back to
    nvidia-smi when the NVML bindings are missing."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.sm, self.mx, self.reasons, self._stop_evt = index, [], [], set(), threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll_nvml(self):
        n = self.nvml
        self.sm.append(int(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
        self.mx.append(int(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)))
        try:
            r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        except Exception:
            r = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        self.reasons.update(k for k, v in bits.items() if r & v)

    def _poll_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        r = [c.strip() for c in out.split(",")]
        if len(r) >= 6 and r[0].isdigit():
            self.sm.append(int(r[0]))
            self.mx.append(int(r[1]))
            self.reasons.update(self.NAMES[i] for i in range(4) if r[2 + i].lower().startswith("active"))

    def run(self):
        while not self._stop_evt.is_set():
            try:
                if self.nvml is not None:
                    self._poll_nvml()
                else:
                    self._poll_smi()
            except Exception:
                pass
            self._stop_evt.wait(0.01 if self.nvml is not None else 0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def model_args(scale):
    return types.SimpleNamespace(scale=scale, rgb_range=1.0, colors=3, n_feats=64, num_heads=4, n_blocks=8)


def metric_name(scale):
    return "x4 SR output megapixels/sec" if scale == 4 else f"x{scale} SR output megapixels/sec"


def cpu_reference_pass(scale, x, threads):
    """The CPU leg: oracle.forward (torch fp32 port of ref M2Trans_network.py:58-76) on the host cores."""
    import torch
    from oracle import m2trans_oracle as O
    from m2trans_b200.synthetic import synthetic_state_dict
    torch.set_num_threads(threads)
    sd = synthetic_state_dict(scale, 0)
    with torch.no_grad():
        O.forward(sd, x[:1])                                   # warm-up (allocator, threads)
        t0 = time.perf_counter()
        y = O.forward(sd, x)
        dt = time.perf_counter() - t0
    return dt, y


def run_reference(args):
    """--impl reference: the reference's own algorithm on the box's host cores (oracle port; the reference is Python and
    /root/reference does not exist on the GPU box).  Rank 0 only.  Same `config`, metric and unit as the GPU arm; every
    step processes a bounded sample of the workload's frames (all 16 for cfg2)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "cfg5":
        print(json.dumps({"impl": "reference", "unavailable": "the MedCLIP image tower is a third-party package that is absent; its stand-in oracle is exercised by tests/test_clip_cpu.py only"}), flush=True)
        return
    import torch
    from m2trans_b200.synthetic import synthetic_input
    scale, B, H, W, scaling = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    # bounded sample: the whole batch when steps + warmup passes over it finish in a few minutes (cfg1, cfg2), else a slice
    per_img = 0.45 * (H * W) / (128 * 128)
    budget = 200.0 / max(1, args.steps + args.warmup)
    nimg = max(1, min(B, int(budget / per_img)))
    x = synthetic_input(B, H, W, seed=33)[:nimg]
    times = []
    for i in range(args.warmup + args.steps):
        dt, _ = cpu_reference_pass(scale, x, threads)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    mp = nimg * (H * scale) * (W * scale) / 1e6
    val = mp / (ms / 1e3)
    sample = f"{nimg} of {B} frames of {args.workload} per step, torch fp32 port of the reference forward, {threads} threads"
    print(json.dumps({
        "impl": "reference", "metric": metric_name(scale), "value": val, "unit": "MP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.workload),
        "cpu_baseline": {"value": val, "unit": "MP/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def dist_setup():
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path (use --impl reference for the CPU leg)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    return rank, world, local, dev, barrier


def run_clip(args):
    """cfg5: MedCLIP image-embedding pass over x4 SR outputs, 256 images split over the ranks (strong scaling)."""
    import torch
    import torch.distributed as dist
    from m2trans_b200.medclip_image import MedCLIPVisionModelViT, synthetic_state_dict
    from m2trans_b200.sharding import max_over_ranks, shard_range
    rank, world, local, dev, barrier = dist_setup()
    total, size = 256, 512
    lo, hi = shard_range(total, rank, world)
    nloc = hi - lo
    tower = MedCLIPVisionModelViT()
    tower.load_state_dict(synthetic_state_dict(0), strict=False)
    tower = tower.to(dev)
    g = torch.Generator().manual_seed(33 + rank)
    xs = [torch.rand(nloc, 3, size, size, generator=g).to(dev) for _ in range(2)]
    text = torch.randn(512, generator=g).to(dev)
    for i in range(max(3, args.warmup)):
        tower.encode_image(xs[i % 2], text)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    torch.cuda._sleep(40_000_000)
    for i in range(args.steps):
        ev[i][0].record()
        tower.encode_image(xs[i % 2], text)
        ev[i][1].record()
    barrier()
    clocks = sampler.stop()
    ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    (ms,) = max_over_ranks([ms], device=dev)
    pk = peaks()
    if rank == 0:
        tflops = total * 8.98 / ms / world            # per GPU: 8.98 GFLOP per 224x224 image
        print(json.dumps({
            "metric": "MedCLIP image embeddings per second", "value": total / ms * 1e3, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config("cfg5"), "clocks": clocks,
            "gpu_launches": 94 * args.steps,
            "timing": "CUDA events per step, max over ranks; 2 rotating input batches (>= 100 MB of input and ~1 GB of intermediates per step per GPU at N = 1, far above L2)",
            "roofline": {"kernel": "whole pass (tcgen05 bf16 Linear / MLP kernels dominate)", "bound": "tensor", "achieved": tflops,
                         "peak": pk["tflops_sustained"], "unit": "TFLOP/s", "frac": tflops / pk["tflops_sustained"], "traffic": None,
                         "peak_source": pk["source"]},
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


def kernel_table(net, x_dev, scale, P, pk):
    """Per kernel class of ONE eager forward (m2t_debug_profile_forward: a CUDA event after every launch)."""
    rows, largest = [], None
    for line in net.profile_forward(x_dev).strip().splitlines():
        parts = line.split()
        if len(parts) < 5 or parts[1] != "us":
            continue
        us, n, name = float(parts[0]), int(parts[3][1:]), parts[4]
        label, bound, flop, byt = name[:40], None, 0, 0
        for key, lab, bnd, f, b in KERNELS:
            if key in name:
                label, bound, flop, byt = lab, bnd, f, b
                break
        if "tail_strip" in name or "tail_fused" in name:
            # last tail stage: the tensor FLOPs of the stage (1x1 conv 64 -> 256 at the stage's input resolution + the
            # 3x3 conv at twice that) and, as the quantity that bounds it, the GELU evaluations on the FMA pipe
            flop = (2 * 64 * 256 + 4 * 2 * 64 * 3 * 9) * (4 if scale == 4 else 1)
        row = {"kernel": label, "launches": n, "us_total": us, "us_per_launch": us / n}
        if bound == "tensor" and flop:
            row.update(bound="tensor", achieved=flop * P / (us / n * 1e-6) / 1e12, unit="TFLOP/s")
            row["frac"] = row["achieved"] / pk["tflops_burst"]
        elif bound == "hbm" and byt:
            row.update(bound="hbm", achieved=byt * P / (us / n * 1e-6) / 1e9, unit="GB/s")
            row["frac"] = row["achieved"] / pk["hbm_gbs"]
            if flop:
                row["tflops"] = flop * P / (us / n * 1e-6) / 1e12
        elif bound == "fma":
            gelus = 64 * 4 * (4 if scale == 4 else 1) * P          # GELU evaluations of the stage (output channels x sub-pixels)
            # 12 FMA-pipe cycles per GELU and lane (gelu.cuh): 128 lanes per SM and clock
            floor_us = gelus * 12 / (128 * 148) / 1.965e9 * 1e6
            row.update(bound="fma-pipe (exact-erf GELU, 12 pipe cycles each)", floor_us=floor_us, frac=floor_us / (us / n),
                       tflops=flop * P / (us / n * 1e-6) / 1e12)
        rows.append(row)
        if largest is None or us / n > largest["us_per_launch"]:
            largest = {"kernel": label, "us_per_launch": us / n}
    return rows, largest


def run_ours(args):
    import torch
    import torch.distributed as dist
    from m2trans_b200 import _lib
    from m2trans_b200.M2Trans_network import M2Trans
    from m2trans_b200.loader import images_from_device, images_to_device
    from m2trans_b200.metrics import psnr_ssim
    from m2trans_b200.sharding import gather_counts, max_over_ranks, shard_range
    from m2trans_b200.synthetic import reference_checkpoint, synthetic_input

    rank, world, local, dev, barrier = dist_setup()
    scale, Bnamed, H, W, scaling = WORKLOADS[args.workload]
    lib = _lib.load()

    model = torch.nn.DataParallel(M2Trans(model_args(scale)), device_ids=[local]).to(dev)
    model.load_state_dict(reference_checkpoint(scale, 0)["model_state_dict"], strict=True)
    model.eval()
    net = model.module

    # weak: every rank owns its own batch of the named size (different seeds); strong: the named batch is split by image
    if scaling == "weak":
        B, seeds = Bnamed, [33 + 17 * rank + i for i in range(3)]
        take = lambda t: t
    else:
        lo, hi = shard_range(Bnamed, rank, world)
        B, seeds = hi - lo, [33 + i for i in range(3)]
        take = lambda t: t[lo:hi].contiguous()
    n_rot = 3
    xs_host = [take(synthetic_input(Bnamed, H, W, seed=s)).pin_memory() for s in seeds]
    xs_dev = [x.to(dev) for x in xs_host]
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.float32, device=dev)     # 768 MB > 126 MB L2
    out_mp_local = B * (H * scale) * (W * scale) / 1e6
    out_mp_total = sum(gather_counts(B, device=dev)) * (H * scale) * (W * scale) / 1e6
    if B == 0:
        raise SystemExit(f"bench.py: rank {rank} has no frames ({Bnamed} frames over {world} ranks)")

    # ---- device-resident timing --------------------------------------------------------------------
    for i in range(args.warmup):
        net(xs_dev[i % n_rot])
    barrier()
    sampler = ClockSampler(local)
    sampler.start()

    def timed_region():
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        barrier()
        t0 = time.perf_counter()
        # Park the GPU for ~20 ms so the host enqueues all K steps ahead of it: the per-step CUDA events then measure
        # device time only, not host launch jitter of a busy box (observed: 1.8 vs 2.8 ms for the same binary).
        torch.cuda._sleep(40_000_000)
        for i in range(args.steps):
            flush.zero_()                               # evict L2 between timed iterations (not timed)
            ev[i][0].record()
            net(xs_dev[i % n_rot])
            ev[i][1].record()
        barrier()
        return [a.elapsed_time(b) for a, b in ev], time.perf_counter() - t0

    step_ms, t_wall = timed_region()
    remeasured = None
    if max(step_ms) > 2.0 * sorted(step_ms)[len(step_ms) // 2]:
        # One step several times slower than the median of the same K steps is a stall of the box, not of the kernels
        # (seen once: one 108 ms step among nine of 1.55 ms, clocks at maximum, no throttle reason).  Like a throttled run,
        # the region is measured once more and both are reported.
        first = {"ms_per_step": sum(step_ms) / len(step_ms), "max_step_ms": max(step_ms)}
        step2, t_wall2 = timed_region()
        if sum(step2) < sum(step_ms):
            step_ms, t_wall = step2, t_wall2
        remeasured = {"reason": "a step slower than 2x the median step", "first_region": first}
    ms = sum(step_ms) / len(step_ms)
    step_sorted = sorted(step_ms)
    launches = net.last_launches * args.steps

    # ---- end to end through the public call with host buffers -----------------------------------------
    # Every step: pinned H2D of the LR batch, model(x) (the call a user of the reference makes, ref test.py:90), the SR
    # batch to the host.  Copies run on their own streams and are double-buffered, so step i+1's upload and step i-1's
    # download overlap step i's compute (plain torch stream / event API).  mode "u8": the SR batch is converted to uint8
    # HWC on the device first (m2t_f32chw_to_u8hwc: the bytes an image writer stores); mode "fp32": the raw fp32 tensor.
    Hs, Ws = H * scale, W * scale
    s_in, s_out, s_main = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.current_stream(dev)

    def e2e_measure(mode):
        if mode == "u8":
            y_host = [torch.empty((B, Hs, Ws, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
            y_dev8 = [torch.empty((B, Hs, Ws, 3), dtype=torch.uint8, device=dev) for _ in range(2)]
        else:
            y_host = [torch.empty((B, 3, Hs, Ws), dtype=torch.float32).pin_memory() for _ in range(2)]
        x_dev = [torch.empty_like(xs_dev[0]) for _ in range(2)]
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_done = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(2)]

        def step(i):
            k = i % 2
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_done[k])                 # x_dev[k] is free once step i-2 has consumed it
                x_dev[k].copy_(xs_host[i % n_rot], non_blocking=True)
                ev_in[k].record(s_in)
            s_main.wait_event(ev_in[k])
            yd = model(x_dev[k])
            if mode == "u8":
                s_main.wait_event(ev_out[k])                # y_dev8[k] has been copied out (step i-2)
                images_from_device(yd, out=y_dev8[k])
                src = y_dev8[k]
            else:
                src = yd
            ev_done[k].record(s_main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_done[k])
                y_host[k].copy_(src, non_blocking=True)
                ev_out[k].record(s_out)
                src.record_stream(s_out)
        for i in range(3):
            step(i)
        barrier()
        # Timed on the device like the leg above: the three streams are parked behind one event so that the host can
        # queue all K steps ahead (a shared box can stall the Python thread for several ms, which wall-clock timing of
        # 10 steps turns into a 2x error).  The region still contains every H2D copy, every forward through the public
        # call and every D2H copy of the K steps.  Three regions, median reported, all kept; the wall clock of the
        # median region is reported beside it (it includes the ~20 ms of parking, spread over the K steps).
        runs = []
        for _rep in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            torch.cuda._sleep(40_000_000)
            a.record(s_main)
            s_in.wait_event(a)
            s_out.wait_event(a)
            for i in range(args.steps):
                step(i)
            s_out.wait_stream(s_main)
            b.record(s_out)
            torch.cuda.synchronize()
            runs.append((a.elapsed_time(b) / args.steps, 1e3 * (time.perf_counter() - t0) / args.steps))
        runs.sort()
        # the same K steps without parking, host clock from the first call to the final synchronize (what a caller sees when
        # the Python thread is not stalled by a shared box); median of 3
        walls = []
        for _rep in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(args.steps):
                step(i)
            s_out.wait_stream(s_main)
            torch.cuda.synchronize()
            walls.append(1e3 * (time.perf_counter() - t0) / args.steps)
        walls.sort()
        return runs[1][0], runs[1][1], [r[0] for r in runs], walls[1]

    e2e_ms, e2e_wall, e2e_all, e2e_wall_free = e2e_measure("u8")
    e2e32_ms, e2e32_wall, e2e32_all, _ = e2e_measure("fp32")

    # ---- the reference's eval loop on the device (ref test.py:87-116): uint8 LR + HR in, two scalars out --------------
    g8 = torch.Generator().manual_seed(5 + rank)
    lr8 = [torch.randint(0, 256, (B, H, W, 3), generator=g8, dtype=torch.uint8).pin_memory() for _ in range(2)]
    hr8 = [torch.randint(0, 256, (B, Hs, Ws, 3), generator=g8, dtype=torch.uint8).pin_memory() for _ in range(2)]
    res_host = torch.empty(B, 2, dtype=torch.float32).pin_memory()

    def eval_step(i):
        x = images_to_device(lr8[i % 2], dev)
        hr = images_to_device(hr8[i % 2], dev)
        sr = model(x)
        per_image, _batch = psnr_ssim(sr, hr, scale)
        res_host.copy_(per_image, non_blocking=True)
    try:
        for i in range(3):
            eval_step(i)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(40_000_000)
        a.record()
        for i in range(args.steps):
            eval_step(i)
        b.record()
        torch.cuda.synchronize()
        eval_ms = a.elapsed_time(b) / args.steps
    except Exception as e:  # noqa: BLE001  (a side measurement must not take the headline line down)
        eval_ms, eval_err = None, repr(e)[:200]
    clocks = sampler.stop()

    # ---- per-kernel table of one eager forward + the dominant kernel alone ---------------------------------------------
    hp, wp = (H + 31) // 32 * 32, (W + 31) // 32 * 32
    P = B * hp * wp
    pk = peaks()
    kernels, largest = kernel_table(net, xs_dev[0], scale, P, pk)
    dominant = max(kernels, key=lambda r: r["us_total"]) if kernels else None
    st = torch.cuda.current_stream(dev).cuda_stream

    def time_alone(fn, n_k=10):
        for i in range(4):
            fn(i)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(20_000_000)            # let the host queue all launches first
        a.record()
        for i in range(n_k):
            fn(i)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n_k

    traffic_file = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    traffic = json.load(open(traffic_file)) if os.path.exists(traffic_file) else {}

    def roofline_ffconv():
        sets = [(torch.randn(B, hp, wp, 64, device=dev).half(), torch.randn(B, hp, wp, 64, device=dev)) for _ in range(2)]
        ffw = torch.randn(9, 64, 64, device=dev).half() * 0.05
        ffb = torch.randn(64, device=dev)
        stats = torch.zeros(B, 64, 2, dtype=torch.float64, device=dev)

        def ffconv(i):      # in place on the fp32 stream, as inside the forward (Xin == Xout from the second CFTM on)
            Yk, Xk = sets[i % 2]
            _lib.check(lib.m2t_stage_ffconv(0, Yk.data_ptr(), ffw.data_ptr(), ffb.data_ptr(), Xk.data_ptr(), Xk.data_ptr(),
                                            stats.data_ptr(), B, hp, wp, st), "m2t_stage_ffconv")
        k_ms = time_alone(ffconv)
        byt, flop = (128 + 256 + 256) * P, 73728 * P
        t = traffic.get("ffconv_umma", {}) if args.workload == "cfg2" else {}
        return {"kernel": "ffconv_umma (CFTM 3x3 feed-forward conv + residual + norm stats)", "bound": "hbm",
                "achieved": byt / (k_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": byt / (k_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                "traffic": t.get("dram_bytes_per_launch"), "traffic_source": t.get("source"), "algorithmic_bytes_per_launch": byt,
                "tflops": flop / (k_ms * 1e-3) / 1e12, "frac_of_tensor_burst": flop / (k_ms * 1e-3) / 1e12 / pk["tflops_burst"],
                "peak_source": pk["source"], "ms_per_launch": k_ms,
                "timing": "10 launches back to back between one CUDA-event pair, two rotating buffer sets"}

    def roofline_attn_z256():
        h4, w4 = hp // 4, wp // 4
        sets = [(torch.randn(B, h4, w4, 256, device=dev).half(), torch.randn(B, h4, w4, 256, device=dev).half() * 0.5,
                 torch.empty(B, hp, wp, 64, dtype=torch.float16, device=dev)) for _ in range(2)]
        mq = (torch.randn(288, 256, device=dev) * (0.5 / 256)).half()
        wv = (torch.randn(256, 256, device=dev) / 16).half()

        def az(i):
            T, Hn, Y = sets[i % 2]
            _lib.check(lib.m2t_stage_attn_z(256, T.data_ptr(), mq.data_ptr(), wv.data_ptr(), Y.data_ptr(), Hn.data_ptr(), 2,
                                            B, h4, w4, st), "m2t_stage_attn_z")
        k_ms = time_alone(az)
        flop, byt = (24576 + 6400) * P, (32 + 32 + 32 + 32) * P
        t = traffic.get("attn_z256", {}) if args.workload == "cfg2" else {}
        return {"kernel": "attn_z_kernel<256> (CFTM branch 3: qkv conv + halo attention + branch glue in one kernel)", "bound": "tensor",
                "achieved": flop / (k_ms * 1e-3) / 1e12, "peak": pk["tflops_burst"], "unit": "TFLOP/s",
                "frac": flop / (k_ms * 1e-3) / 1e12 / pk["tflops_burst"], "traffic": t.get("dram_bytes_per_launch"),
                "traffic_source": t.get("source"), "algorithmic_flops_per_launch": flop, "algorithmic_bytes_per_launch": byt,
                "hbm_gbs": byt / (k_ms * 1e-3) / 1e9, "peak_source": pk["source"], "ms_per_launch": k_ms,
                "note": "algorithmic FLOPs of the reference's qkv conv + attention (SURVEY appendix C); the kernel executes fewer (contractions re-associated)",
                "timing": "10 launches back to back between one CUDA-event pair, two rotating buffer sets"}

    if dominant is not None and "attn_z<256>" in dominant["kernel"]:
        roof, roof2 = roofline_attn_z256(), roofline_ffconv()
    else:
        roof, roof2 = roofline_ffconv(), roofline_attn_z256()
    fwd_tflops = FLOP_PER_PX[scale] * P / (ms * 1e-3) / 1e12

    # ---- max over ranks ------------------------------------------------------------------------------------
    vals = [ms, e2e_ms, e2e32_ms, e2e_wall, eval_ms if eval_ms is not None else 0.0, e2e_wall_free]
    ms_g, e2e_g, e2e32_g, wall_g, eval_g, wall_free_g = max_over_ranks(vals, device=dev)
    total_mp = out_mp_total if scaling == "strong" else world * out_mp_local

    if rank == 0:
        cfg = workload_config(args.workload)
        line = {
            "metric": metric_name(scale), "value": total_mp / (ms_g * 1e-3), "unit": "MP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_g, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f16", "data": "synthetic", "config": cfg,
            "notes": {"frames_per_gpu": B, "l2": "768 MB buffer rewritten between timed iterations; 3 rotating input batches",
                      "sharding": "images; no collective on the data path",
                      "precision": "fp16 GEMM operands, fp32 accumulate (TMEM), fp32 residual stream / softmax / norm statistics; "
                                   + ("fast mode (plain fp16 operands)" if scale == 4 else "precise mode (fp16 hi + residual pairs on the residual path and in the ff conv)")},
            "clocks": clocks,
            "e2e": {"value": total_mp / (e2e_g * 1e-3), "unit": "MP/s", "h2d_bytes_per_step": B * 3 * H * W * 4,
                    "d2h_bytes_per_step": B * 3 * Hs * Ws, "ms_per_step": e2e_g, "wall_ms_per_step": wall_g,
                    "wall_unparked_ms_per_step": wall_free_g, "wall_unparked_value": total_mp / (wall_free_g * 1e-3),
                    "result": "SR batch as uint8 HWC (converted on the device), pinned host buffers",
                    "timing": "CUDA events over K pipelined steps (H2D + forward + uint8 conversion + D2H each), streams parked so the "
                              "host queues ahead; median of 3 such regions; wall_ms_per_step = host clock of the same region incl. ~20 ms parking / K; "
                              "wall_unparked_* = host clock of K steps issued without parking, first call to final synchronize, median of 3",
                    "ms_per_step_all": e2e_all},
            "e2e_fp32": {"value": total_mp / (e2e32_g * 1e-3), "unit": "MP/s", "h2d_bytes_per_step": B * 3 * H * W * 4,
                         "d2h_bytes_per_step": B * 3 * Hs * Ws * 4, "ms_per_step": e2e32_g, "ms_per_step_all": e2e32_all,
                         "result": "SR batch as the fp32 tensor the module returns"},
            "e2e_eval": ({"value": total_mp / (eval_g * 1e-3), "unit": "MP/s", "ms_per_step": eval_g,
                          "h2d_bytes_per_step": B * 3 * (H * W + Hs * Ws), "d2h_bytes_per_step": 8 * B,
                          "what": "ref test.py:87-116 on the device: uint8 LR + HR upload, loader conversion, forward, Y-PSNR + SSIM, scalars to the host"}
                         if eval_ms is not None else {"error": eval_err}),
            "gpu_launches": launches,
            "remeasured": remeasured,
            "step_ms": {"min": step_sorted[0], "median": step_sorted[len(step_sorted) // 2], "max": step_sorted[-1]},
            "roofline": roof,
            "roofline_second": roof2,
            "roofline_forward": {"bound": "tensor", "achieved": fwd_tflops, "peak": pk["tflops_sustained"], "unit": "TFLOP/s",
                                 "frac": fwd_tflops / pk["tflops_sustained"], "peak_source": pk["source"]},
            "kernels": kernels, "largest_launch": largest,
            "wall_s_timed_region": t_wall,
        }
        if world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            nimg = max(1, min(B, int(30.0 / (0.45 * (H * W) / (128 * 128)))))
            dt, yref = cpu_reference_pass(scale, xs_host[0][:nimg], threads)
            from oracle import m2trans_oracle as O
            yo = net(xs_dev[0])[:nimg].cpu()
            line["cpu_baseline"] = {"value": nimg * Hs * Ws / 1e6 / dt, "unit": "MP/s", "cores": threads,
                                    "kind": "port", "sample": f"{nimg} of {B} frames of {args.workload}, one pass after a 1-frame warm-up, torch fp32 port of the reference forward"}
            line["parity"] = {"psnr_db": O.psnr(yo, yref), "max_abs": O.max_abs(yo, yref), "frames": nimg}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS) + ["cfg5"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "cfg5":
        run_clip(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
