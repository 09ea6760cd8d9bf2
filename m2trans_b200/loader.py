"""Device side of the benchmark loader (ref datas/benchmark.py:62-69).

The reference keeps every image as a uint8 HWC array in RAM and returns `ndarray2tensor(img) / 255.` (ref utils.py:237-240),
i.e. converts on the CPU and ships 12 bytes per pixel to the GPU.  `images_to_device` ships the uint8 bytes (3 per pixel,
from pinned memory when given) and does the HWC -> CHW permutation, the float conversion and the division on the device,
bit-identically.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import M2TError

__all__ = ["images_to_device", "images_from_device"]


@torch.no_grad()
def images_to_device(images, device="cuda", denom: float = 255.0, non_blocking: bool = True) -> torch.Tensor:
    """images: uint8 [B,H,W,C] or [H,W,C] (numpy array or torch tensor, C = 1 or 3) -> fp32 [B,C,H,W] on `device`,
    equal to torch.from_numpy(img.transpose(2, 0, 1)).float() / denom of the reference loader."""
    t = torch.from_numpy(np.ascontiguousarray(images)) if isinstance(images, np.ndarray) else images
    if t.dtype != torch.uint8:
        raise M2TError(f"images_to_device: expected uint8 images, got {t.dtype}")
    if t.dim() == 3:
        t = t.unsqueeze(0)
    if t.dim() != 4 or t.shape[3] not in (1, 3):
        raise M2TError(f"images_to_device: expected [B,H,W,C] with C in (1, 3), got {tuple(t.shape)}")
    dev = torch.device(device)
    if dev.type != "cuda":
        raise M2TError("images_to_device: the B200 engine has no CPU path")
    b, h, w, c = t.shape
    lib = _lib.load()
    with torch.cuda.device(dev):
        src = t.contiguous().to(dev, non_blocking=non_blocking)
        dst = torch.empty(b, c, h, w, dtype=torch.float32, device=dev)
        _lib.check(lib.m2t_u8hwc_to_f32chw(src.data_ptr(), dst.data_ptr(), b, h, w, c, float(denom),
                                           torch.cuda.current_stream(dev).cuda_stream), "m2t_u8hwc_to_f32chw")
    return dst


@torch.no_grad()
def images_from_device(sr: torch.Tensor, out: torch.Tensor = None, scale: float = 255.0) -> torch.Tensor:
    """sr: fp32 [B,C,H,W] on a CUDA device, values in [0, 1] -> uint8 [B,H,W,C] on the same device (or into `out`),
    equal to (sr * scale).round().clamp(0, 255).byte().permute(0, 2, 3, 1): the bytes an image writer would store.
    Copying THIS to the host moves 3 bytes per pixel instead of 12."""
    if not sr.is_cuda or sr.dtype != torch.float32 or sr.dim() != 4 or sr.shape[1] not in (1, 3):
        raise M2TError(f"images_from_device: expected CUDA fp32 [B,C,H,W] with C in (1, 3), got {sr.dtype} {tuple(sr.shape)}")
    b, c, h, w = sr.shape
    sr = sr.contiguous()
    lib = _lib.load()
    with torch.cuda.device(sr.device):
        if out is None:
            out = torch.empty(b, h, w, c, dtype=torch.uint8, device=sr.device)
        _lib.check(lib.m2t_f32chw_to_u8hwc(sr.data_ptr(), out.data_ptr(), b, h, w, c, float(scale),
                                           torch.cuda.current_stream(sr.device).cuda_stream), "m2t_f32chw_to_u8hwc")
    return out
