"""MedCLIP image-embedding pass on the B200 engine (ref losses.py:42-81, image side).

The reference builds `MedCLIPModel(vision_cls=MedCLIPVisionModelViT)` (ref losses.py:22-24) and calls
`encode_image` on 224x224 views of the SR and HR images (:53-54, :68-69), L2-normalises (:71-72) and takes the dot
product with the normalised text feature (:76-77).  `medclip` is a third-party package that is not installed here;
what it builds for the ViT variant is the Hugging Face Swin-T (`microsoft/swin-tiny-patch4-window7-224`) followed by
`projection_head = nn.Linear(768, 512, bias=False)`.

`MedCLIPVisionModelViT` below is a parameter container with exactly those state_dict keys (`model.<SwinModel keys>`,
`projection_head.weight`), so an upstream vision-tower checkpoint loads with `load_state_dict(strict=True)`; its
arithmetic is the engine's `m2t_clip_encode_image` (include/m2trans_b200.h): bf16 tcgen05 GEMMs, fp32 residual stream,
CUDA only, no PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from ._lib import M2TError

__all__ = ["MedCLIPVisionModelViT", "swin_param_spec", "synthetic_state_dict", "semantic_distance"]

DEPTHS = (2, 2, 6, 2)
HEADS = (3, 6, 12, 24)
EMBED, WINDOW, IMG, FEAT, PROJ = 96, 7, 224, 768, 512


def swin_param_spec():
    """[(name, shape)] in the engine's parameter order: SwinModel.state_dict() order (modeling_swin.py) without the
    relative_position_index buffers, then the projection head."""
    spec = [("model.embeddings.patch_embeddings.projection.weight", (EMBED, 3, 4, 4)),
            ("model.embeddings.patch_embeddings.projection.bias", (EMBED,)),
            ("model.embeddings.norm.weight", (EMBED,)), ("model.embeddings.norm.bias", (EMBED,))]
    for s, depth in enumerate(DEPTHS):
        c = EMBED << s
        for b in range(depth):
            p = f"model.encoder.layers.{s}.blocks.{b}."
            spec += [(p + "layernorm_before.weight", (c,)), (p + "layernorm_before.bias", (c,)),
                     (p + "attention.self.relative_position_bias_table", ((2 * WINDOW - 1) ** 2, HEADS[s]))]
            for n in ("query", "key", "value"):
                spec += [(p + f"attention.self.{n}.weight", (c, c)), (p + f"attention.self.{n}.bias", (c,))]
            spec += [(p + "attention.output.dense.weight", (c, c)), (p + "attention.output.dense.bias", (c,)),
                     (p + "layernorm_after.weight", (c,)), (p + "layernorm_after.bias", (c,)),
                     (p + "intermediate.dense.weight", (4 * c, c)), (p + "intermediate.dense.bias", (4 * c,)),
                     (p + "output.dense.weight", (c, 4 * c)), (p + "output.dense.bias", (c,))]
        if s < len(DEPTHS) - 1:
            p = f"model.encoder.layers.{s}.downsample."
            spec += [(p + "reduction.weight", (2 * c, 4 * c)), (p + "norm.weight", (4 * c,)), (p + "norm.bias", (4 * c,))]
    spec += [("model.layernorm.weight", (FEAT,)), ("model.layernorm.bias", (FEAT,)),
             ("projection_head.weight", (PROJ, FEAT))]
    return spec


def synthetic_state_dict(seed: int = 0, gain: float = 1.0):
    """Random weights of the right shapes (there is no network for the released ones).  Linear weights are drawn at
    N(0, gain^2 / fan_in) and the norm / bias / table entries are perturbed, so every branch contributes at the scale of
    the residual stream and a wrong window, shift, mask or bias index moves the embedding visibly."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in swin_param_spec():
        if name.endswith("norm.weight") or "layernorm" in name and name.endswith(".weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif "relative_position_bias_table" in name:
            t = 0.5 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            t = 0.1 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = torch.randn(shape, generator=g) * (gain / fan_in ** 0.5)
        sd[name] = t.float()
    return sd


def _relative_position_index(ws: int = WINDOW) -> torch.Tensor:
    c = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")).flatten(1)
    rel = (c[:, :, None] - c[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def _child(root: nn.Module, path):
    for part in path:
        if part not in root._modules:
            root.add_module(part, nn.Module())
        root = root._modules[part]
    return root


class MedCLIPVisionModelViT(nn.Module):
    """Swin-T + projection head with upstream's parameter names.  `forward` / `encode_image` return the L2-normalised
    [B,512] embedding `MedCLIPModel.encode_image` hands to ref losses.py:68-72 (upstream's tower returns the projection
    before the normalisation; the reference never sees that intermediate)."""

    def __init__(self, checkpoint=None, medclip_checkpoint=None):
        super().__init__()
        if checkpoint is not None or medclip_checkpoint is not None:
            raise M2TError("MedCLIPVisionModelViT: load weights with load_state_dict; nothing is downloaded here")
        for name, shape in swin_param_spec():
            *path, leaf = name.split(".")
            init = torch.ones(shape) if leaf == "weight" and len(shape) == 1 else torch.zeros(shape)
            if len(shape) > 1 and "table" not in leaf:
                init = torch.randn(shape) * 0.02
            _child(self, path).register_parameter(leaf, nn.Parameter(init, requires_grad=False))
            if leaf == "relative_position_bias_table":
                _child(self, path).register_buffer("relative_position_index", _relative_position_index())
        self._packed = None
        self._packed_key = None
        self._graphs = {}
        self.cuda_graph = True               # replay a captured graph for small inputs (see _encode_graphed)
        self.graph_max_pixels = 4 << 20      # B*H*W above which the input copy costs more than the launches

    def _params(self):
        sd = dict(self.named_parameters())
        return [sd[name].detach() for name, _ in swin_param_spec()]

    def invalidate(self) -> None:
        """Forget the packed weights and the captured graphs.  Call after editing parameters through `.data` (such edits keep
        the tensor's address and version counter, so the (data_ptr, _version) key cannot see them); load_state_dict and
        .to() / .cuda() call it themselves."""
        self._packed = None
        self._packed_key = None
        self._graphs = {}

    repack = invalidate

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        if hasattr(self, "_graphs"):
            self.invalidate()
        return out

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate()
        return out

    def _pack(self, device):
        lib = _lib.load()
        params = self._params()
        key = tuple((p.data_ptr(), p._version) for p in params)
        if self._packed is not None and self._packed_key == key and self._packed.device == device:
            return self._packed
        if len(params) != lib.m2t_clip_param_count():
            raise M2TError("MedCLIPVisionModelViT: parameter count does not match the engine's")
        for p, (name, shape) in zip(params, swin_param_spec()):
            if tuple(p.shape) != tuple(shape) or p.dtype != torch.float32 or p.device != device:
                raise M2TError(f"MedCLIPVisionModelViT: {name} must be float32 {shape} on {device}")
        params = [p.contiguous() for p in params]
        packed = torch.empty(int(lib.m2t_clip_packed_bytes()), dtype=torch.uint8, device=device)
        ptrs = (C.c_void_p * len(params))(*[p.data_ptr() for p in params])
        _lib.check(lib.m2t_clip_pack_weights(ptrs, len(params), packed.data_ptr(),
                                             torch.cuda.current_stream(device).cuda_stream), "m2t_clip_pack_weights")
        self._packed, self._packed_key = packed, key
        self._graphs = {}                    # captured graphs point at the previous blob
        return packed

    @torch.no_grad()
    def encode_image(self, pixel_values, text_features=None):
        """pixel_values [B,3,H,W] fp32 CUDA in [0,1]; H, W other than 224 are resized as ref losses.py:53 does
        (bicubic, align_corners=True).  Returns the normalised embedding [B,512], or (embedding, logits [B]) when a
        text feature [512] or [1,512] is given (ref losses.py:73-77)."""
        x = pixel_values
        if not isinstance(x, torch.Tensor) or not x.is_cuda:
            raise M2TError("MedCLIPVisionModelViT: expected a CUDA tensor; the B200 engine has no CPU path")
        if x.dtype != torch.float32 or x.dim() != 4 or x.shape[1] != 3:
            raise M2TError(f"MedCLIPVisionModelViT: expected float32 [B,3,H,W], got {x.dtype} {tuple(x.shape)}")
        b, _, h, w = x.shape
        lib = _lib.load()
        if self.cuda_graph and b * h * w <= self.graph_max_pixels and not torch.cuda.is_current_stream_capturing():
            return self._encode_graphed(x, text_features)
        with torch.cuda.device(x.device):
            packed = self._pack(x.device)
            xc = x.contiguous()
            embed = torch.empty(b, PROJ, dtype=torch.float32, device=x.device)
            text = logits = None
            if text_features is not None:
                text = text_features.to(device=x.device, dtype=torch.float32).reshape(-1).contiguous()
                if text.numel() != PROJ:
                    raise M2TError(f"MedCLIPVisionModelViT: text feature must have {PROJ} elements")
                logits = torch.empty(b, dtype=torch.float32, device=x.device)
            ws = torch.empty(int(lib.m2t_clip_workspace_bytes(b)), dtype=torch.uint8, device=x.device)
            _lib.check(lib.m2t_clip_encode_image(packed.data_ptr(), xc.data_ptr(), b, h, w, embed.data_ptr(),
                                                 text.data_ptr() if text is not None else None,
                                                 logits.data_ptr() if logits is not None else None, ws.data_ptr(),
                                                 torch.cuda.current_stream(x.device).cuda_stream),
                       "m2t_clip_encode_image")
        return embed if logits is None else (embed, logits)

    def _encode_graphed(self, x, text_features):
        """Small batches (the reference encodes one image per call, ref losses.py:45-46, :68): the 94 launches and their
        tensor-map encodes cost more host time than the GPU needs, so the pass is captured once per (device, B, H, W,
        with/without text) into a CUDA graph over static buffers and replayed; inputs are copied in, results cloned out."""
        dev = x.device
        with torch.cuda.device(dev):
            packed = self._pack(dev)                  # drops the cached graphs when the weights changed
            has_text = text_features is not None
            t = None
            if has_text:
                t = text_features.to(device=dev, dtype=torch.float32).reshape(-1)
                if t.numel() != PROJ:
                    raise M2TError(f"MedCLIPVisionModelViT: text feature must have {PROJ} elements")
            key = (dev.index, tuple(x.shape), has_text)
            entry = self._graphs.get(key)
            if entry is None:
                sx = torch.empty_like(x, memory_format=torch.contiguous_format)
                stext = torch.empty(PROJ, dtype=torch.float32, device=dev) if has_text else None
                sx.copy_(x)
                if has_text:
                    stext.copy_(t)
                self.cuda_graph = False
                try:
                    side = torch.cuda.Stream(dev)
                    side.wait_stream(torch.cuda.current_stream(dev))
                    with torch.cuda.stream(side):
                        self.encode_image(sx, stext)                      # warm-up outside the capture
                    torch.cuda.current_stream(dev).wait_stream(side)
                    graph = torch.cuda.CUDAGraph()
                    try:
                        # thread_local: CUDA calls of other threads (a DataLoader's pin_memory thread, another model)
                        # must not invalidate the capture; the region only enqueues kernels on static buffers
                        with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                            out = self.encode_image(sx, stext)
                    except RuntimeError:
                        torch.cuda.synchronize(dev)
                        return self.encode_image(x, text_features)       # eager (cuda_graph is False here)
                finally:
                    self.cuda_graph = True
                entry = (graph, sx, stext, out, packed)
                self._graphs[key] = entry
            graph, sx, stext, out, _ = entry
            sx.copy_(x)
            if has_text:
                stext.copy_(t)
            graph.replay()
            return (out[0].clone(), out[1].clone()) if has_text else out.clone()

    def forward(self, pixel_values, **kwargs):
        return self.encode_image(pixel_values)


@torch.no_grad()
def semantic_distance(tower: MedCLIPVisionModelViT, sr, hr, text_features, n_patches: int = 3):
    """|logit_sr - logit_hr| / N per image pair (ref losses.py:76-79) for batches of 224x224 views or full images."""
    _, ls = tower.encode_image(sr, text_features)
    _, lh = tower.encode_image(hr, text_features)
    return (ls - lh).abs() / float(n_patches)
