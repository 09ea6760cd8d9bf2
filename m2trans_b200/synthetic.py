"""Seeded synthetic checkpoints in the reference's on-disk format.

The released `checkpoints/model_x{2,3,4}.pt` are not available offline, so parity
and throughput are measured on checkpoints that have the reference's exact keys,
shapes, dtype and container format (ref train.py:343-349: a dict with
`model_state_dict` saved from an nn.DataParallel wrapper, hence the `module.`
prefix) and the reference's initial-value distributions:

  * nn.Conv2d default init: U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias
  * qkv_conv.weight: kaiming_normal_(fan_out, relu)   (ref M2Trans_network.py:343)
  * rel_h / rel_w: N(0, 1)                            (ref M2Trans_network.py:344-345)
  * sub_mean / add_mean: frozen MeanShift values      (ref M2Trans_network.py:370-379)

Values come from a CPU torch.Generator, so the same (scale, seed) gives the same
tensors in the build container and on the GPU box (same image, same torch).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict

import torch

N_FEATS = 64
RGB_MEAN = (0.4488, 0.4371, 0.4040)


def state_dict_spec(scale: int, n_blocks: int = 8, n_feats: int = N_FEATS, colors: int = 3):
    """(key, shape) in the reference's registration order."""
    nf = n_feats
    spec = [
        ("sub_mean.weight", (3, 3, 1, 1)), ("sub_mean.bias", (3,)),
        ("add_mean.weight", (3, 3, 1, 1)), ("add_mean.bias", (3,)),
        ("head.weight", (nf, colors, 3, 3)), ("head.bias", (nf,)),
    ]
    for i in range(n_blocks):
        for name, ch in (("attn1", nf // 4), ("attn2", nf), ("attn3", nf * 4), ("attn4", nf * 4)):
            p = f"body.{i}.{name}."
            spec += [(p + "rel_h", (1, 10, 1, ch // 2)), (p + "rel_w", (1, 1, 10, ch // 2)),
                     (p + "qkv_conv.weight", (3 * ch, ch, 1, 1))]
        spec += [(f"body.{i}.feed_forward.0.weight", (nf, nf, 3, 3)),
                 (f"body.{i}.feed_forward.0.bias", (nf,))]
    if scale == 4:
        spec += [("tail.0.weight", (nf * 4, nf, 1, 1)), ("tail.0.bias", (nf * 4,)),
                 ("tail.3.weight", (nf * 4, nf, 1, 1)), ("tail.3.bias", (nf * 4,)),
                 ("tail.6.weight", (3, nf, 3, 3))]
    else:
        spec += [("tail.0.weight", (nf * scale * scale, nf, 1, 1)), ("tail.0.bias", (nf * scale * scale,)),
                 ("tail.3.weight", (3, nf, 3, 3))]
    return spec


# (out_gain, out_shift) per scale for which the SR output of the seeded checkpoints lies inside (0, 1) for ~95 % of
# the pixels: with the plain initialisation about half of every output is clamped to 0, which hides errors.
UNCLAMPED_TAIL = {2: (0.7, 0.004), 3: (0.7, 0.004), 4: (1.5, 0.02)}


def synthetic_state_dict(scale: int, seed: int = 0, n_blocks: int = 8, qkv_gain: float = 1.0,
                         rgb_range: float = 1.0, out_gain: float = 1.0, out_shift: float = 0.0) -> "OrderedDict[str, torch.Tensor]":
    """Reference-shaped fp32 state dict.  `qkv_gain` > 1 gives the "sharp softmax"
    stress checkpoint of SURVEY.md section 8c.  `out_gain` / `out_shift` rescale and offset the weights of the last
    3x3 conv (w * gain + shift): its inputs are GELU outputs with a positive mean, so a positive shift moves the SR
    image into the interior of [0, 1] like a trained network's output (see UNCLAMPED_TAIL)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1000 * scale + seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    spec = state_dict_spec(scale, n_blocks)
    fan_in_of: Dict[str, int] = {}
    for key, shape in spec:
        if key.endswith(".weight") and len(shape) == 4:
            fan_in_of[key[: -len("weight")]] = shape[1] * shape[2] * shape[3]
    for key, shape in spec:
        if key.startswith(("sub_mean", "add_mean")):
            sign = -1.0 if key.startswith("sub_mean") else 1.0
            if key.endswith("weight"):
                t = torch.eye(3).view(3, 3, 1, 1).clone()
            else:
                t = sign * rgb_range * torch.tensor(RGB_MEAN)
        elif key.endswith(("rel_h", "rel_w")):
            t = torch.randn(shape, generator=g)
        elif key.endswith("qkv_conv.weight"):
            fan_out = shape[0] * shape[2] * shape[3]
            t = torch.randn(shape, generator=g) * (math.sqrt(2.0 / fan_out) * qkv_gain)
        else:
            prefix = key[: key.rfind(".") + 1]
            bound = 1.0 / math.sqrt(fan_in_of[prefix])
            t = (torch.rand(shape, generator=g) * 2.0 - 1.0) * bound
        if key == ("tail.6.weight" if scale == 4 else "tail.3.weight") and (out_gain != 1.0 or out_shift != 0.0):
            t = t * out_gain + out_shift
        sd[key] = t.float().contiguous()
    return sd


def reference_checkpoint(scale: int, seed: int = 0, **kw) -> dict:
    """The dict `torch.save`d by ref train.py:343-349 (DataParallel key prefix)."""
    sd = synthetic_state_dict(scale, seed, **kw)
    return {
        "epoch": 1,
        "model_state_dict": OrderedDict(("module." + k, v) for k, v in sd.items()),
        "optimizer_state_dict": {"state": {}, "param_groups": []},
        "scheduler_state_dict": {},
        "stat_dict": {"epochs": 1, "losses": []},
    }


def save_reference_checkpoint(path: str, scale: int, seed: int = 0, **kw) -> None:
    torch.save(reference_checkpoint(scale, seed, **kw), path)


def synthetic_input(batch: int, h: int, w: int, seed: int = 33, kind: str = "uniform") -> torch.Tensor:
    """LR frames in [0,1] like `lr/255.` (ref datas/benchmark.py:69).  Seed 33 is
    the reference's own (ref test.py:35).  kind='speckle' gives a smooth field
    with multiplicative noise, closer to ultrasound statistics."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    if kind == "uniform":
        return torch.rand(batch, 3, h, w, generator=g, dtype=torch.float32)
    if kind == "speckle":
        lo = torch.rand(batch, 1, max(2, h // 8 + 1), max(2, w // 8 + 1), generator=g)
        base = torch.nn.functional.interpolate(lo, size=(h, w), mode="bilinear", align_corners=True)
        sp = torch.rand(batch, 3, h, w, generator=g)
        return (base * (0.35 + 0.65 * sp)).clamp(0.0, 1.0).float()
    if kind == "flat":        # low contrast: InstanceNorm divides by a small sigma and amplifies whatever noise there is
        return (0.5 + 0.02 * (torch.rand(batch, 3, h, w, generator=g) - 0.5)).float()
    raise ValueError(kind)


# ---- rlutrans.TransBlock -------------------------------------------------
def transblock_state_dict_spec(dim: int = 64):
    """(key, shape) in TransBlock.state_dict() registration order (ref util/rlutrans.py:70-81)."""
    return [("atten.reduce.weight", (dim, dim)), ("atten.qkv.weight", (3 * dim, dim)),
            ("atten.proj.weight", (dim, dim)), ("atten.proj.bias", (dim,)),
            ("norm1.weight", (dim,)), ("norm1.bias", (dim,)),
            ("mlp.fc1.weight", (dim // 4, dim)), ("mlp.fc1.bias", (dim // 4,)),
            ("mlp.fc2.weight", (dim, dim // 4)), ("mlp.fc2.bias", (dim,)),
            ("norm2.weight", (dim,)), ("norm2.bias", (dim,))]


def synthetic_transblock_state_dict(seed: int = 0, dim: int = 64, qkv_gain: float = 3.0) -> "OrderedDict[str, torch.Tensor]":
    """Seeded TransBlock parameters.  Linear weights/biases follow nn.Linear's default U(-1/sqrt(fan_in), +);
    LayerNorm affines are perturbed away from (1, 0) and the qkv weight is scaled by `qkv_gain` so that the
    softmax is not near-uniform (at default init the logits have std ~0.05 and every chunk is an average)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1000 + seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for key, shape in transblock_state_dict_spec(dim):
        if key.startswith("norm"):
            t = torch.randn(shape, generator=g) * (0.2 if key.endswith("weight") else 0.1)
            if key.endswith("weight"):
                t = t + 1.0
        else:
            fan_in = shape[1] if len(shape) == 2 else {"atten.proj.bias": dim, "mlp.fc1.bias": dim, "mlp.fc2.bias": dim // 4}[key]
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2.0 - 1.0) * bound
            if key == "atten.qkv.weight":
                t = t * qkv_gain
        sd[key] = t.float().contiguous()
    return sd


def synthetic_tokens(batch: int, n: int, dim: int = 64, seed: int = 33) -> torch.Tensor:
    """Token features [B, N, dim] ~ N(0.3, 1): a non-zero mean so LayerNorm's centring is exercised."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return (torch.randn(batch, n, dim, generator=g) + 0.3).float()
