"""CPU oracle for the MedCLIP image-embedding pass (ref losses.py:42-81).  TEST INFRASTRUCTURE ONLY:
imported by tests/ and __graft_entry__.smoke() only; never by the product path.

PARITY UNPINNED against the real thing: the arithmetic of `encode_image` lives in the third-party `medclip` package
(unpinned `pip install medclip`, ref README.md:42) on top of transformers==4.24.0 (ref environment.yml:176) with weights
fetched from the network (ref pretrained/medclip-vit/readme.md:3); neither package nor weights exist offline.  What
upstream builds is `AutoModel('microsoft/swin-tiny-patch4-window7-224')` -> pooler_output [B,768] ->
Linear(768,512,bias=False) -> L2 normalise.  This file restates that published architecture in
plain fp32 torch ops, following the Hugging Face implementation that IS in this image
(transformers/models/swin/modeling_swin.py, cited per function), and tests/test_clip_cpu.py pins the restatement against
`transformers.SwinModel(SwinConfig())` itself on random weights.

Parameter names are the Hugging Face ones (`SwinModel.state_dict()`), plus `projection_head.weight` [512,768].
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

DEPTHS = (2, 2, 6, 2)
HEADS = (3, 6, 12, 24)
EMBED = 96
WINDOW = 7
IMG = 224
PROJ = 512
LN_EPS = 1e-5

# emulate_bf16(True): round to bf16 exactly where the engine stores bf16 (GEMM operands: the resized image, every
# LayerNorm output, q/k/v, the attention output, the GELU output, all matrix weights); sums stay fp32.  The engine is
# compared against this mode tightly (what remains is summation order) and against the plain fp32 mode at the bar.
_EMULATE = False


def emulate_bf16(on: bool) -> None:
    global _EMULATE
    _EMULATE = bool(on)


def _r(t: torch.Tensor) -> torch.Tensor:
    return t.bfloat16().float() if _EMULATE else t


def _lin(x, w, b=None):
    return F.linear(x, _r(w), b)


def param_names():
    """The engine's parameter order: SwinModel.state_dict() order without the relative_position_index buffers, then the
    projection head."""
    names = ["embeddings.patch_embeddings.projection.weight", "embeddings.patch_embeddings.projection.bias",
             "embeddings.norm.weight", "embeddings.norm.bias"]
    for s, depth in enumerate(DEPTHS):
        for b in range(depth):
            p = f"encoder.layers.{s}.blocks.{b}."
            names += [p + "layernorm_before.weight", p + "layernorm_before.bias",
                      p + "attention.self.relative_position_bias_table",
                      p + "attention.self.query.weight", p + "attention.self.query.bias",
                      p + "attention.self.key.weight", p + "attention.self.key.bias",
                      p + "attention.self.value.weight", p + "attention.self.value.bias",
                      p + "attention.output.dense.weight", p + "attention.output.dense.bias",
                      p + "layernorm_after.weight", p + "layernorm_after.bias",
                      p + "intermediate.dense.weight", p + "intermediate.dense.bias",
                      p + "output.dense.weight", p + "output.dense.bias"]
        if s < len(DEPTHS) - 1:
            p = f"encoder.layers.{s}.downsample."
            names += [p + "reduction.weight", p + "norm.weight", p + "norm.bias"]
    names += ["layernorm.weight", "layernorm.bias", "projection_head.weight"]
    return names


def relative_position_index(ws: int = WINDOW) -> torch.Tensor:
    """[ws*ws, ws*ws] index into the (2ws-1)^2-row bias table (modeling_swin.py:398-413)."""
    c = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")).flatten(1)
    rel = (c[:, :, None] - c[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def shift_mask(h: int, w: int, ws: int, shift: int) -> torch.Tensor:
    """[nW, ws*ws, ws*ws] additive mask of the shifted layers: -100 between tokens that come from different sides of
    the cyclic wrap (modeling_swin.py:556-582)."""
    img = torch.zeros(h, w)
    cnt = 0
    for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[hs, wsl] = cnt
            cnt += 1
    win = img.view(h // ws, ws, w // ws, ws).permute(0, 2, 1, 3).reshape(-1, ws * ws)
    d = win[:, None, :] - win[:, :, None]
    return torch.where(d != 0, torch.full_like(d, -100.0), torch.zeros_like(d))


def _block(x, P, pre, h, w, heads, shift):
    """One SwinLayer (modeling_swin.py:598-660): LN, (shifted) window attention, residual, LN, GELU MLP, residual."""
    b, n, c = x.shape
    ws = WINDOW
    if min(h, w) <= ws:                      # modeling_swin.py:546-554: a 7x7 map is one window, never shifted
        shift, ws = 0, min(h, w)
    t = _r(F.layer_norm(x, (c,), P[pre + "layernorm_before.weight"], P[pre + "layernorm_before.bias"], LN_EPS)).view(b, h, w, c)
    if shift:
        t = torch.roll(t, (-shift, -shift), (1, 2))
    win = t.view(b, h // ws, ws, w // ws, ws, c).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, c)
    hd = c // heads

    def proj(name):
        y = _r(_lin(win, P[pre + f"attention.self.{name}.weight"], P[pre + f"attention.self.{name}.bias"]))
        return y.view(-1, ws * ws, heads, hd).transpose(1, 2)

    q, k, v = proj("query"), proj("key"), proj("value")
    att = q @ k.transpose(-1, -2) / hd ** 0.5
    table = P[pre + "attention.self.relative_position_bias_table"]
    bias = table[relative_position_index(ws).view(-1)].view(ws * ws, ws * ws, heads).permute(2, 0, 1)
    att = att + bias[None]
    if shift:
        m = shift_mask(h, w, ws, shift)
        att = (att.view(b, -1, heads, ws * ws, ws * ws) + m[None, :, None]).view(-1, heads, ws * ws, ws * ws)
    att = att.softmax(-1)
    o = _r((att @ v).transpose(1, 2).reshape(-1, ws * ws, c))
    o = _lin(o, P[pre + "attention.output.dense.weight"], P[pre + "attention.output.dense.bias"])
    o = o.view(b, h // ws, w // ws, ws, ws, c).permute(0, 1, 3, 2, 4, 5).reshape(b, h, w, c)
    if shift:
        o = torch.roll(o, (shift, shift), (1, 2))
    x = x + o.view(b, n, c)
    t = _r(F.layer_norm(x, (c,), P[pre + "layernorm_after.weight"], P[pre + "layernorm_after.bias"], LN_EPS))
    t = _r(F.gelu(_lin(t, P[pre + "intermediate.dense.weight"], P[pre + "intermediate.dense.bias"])))
    return x + _lin(t, P[pre + "output.dense.weight"], P[pre + "output.dense.bias"])


def _merge(x, P, pre, h, w):
    """SwinPatchMerging (modeling_swin.py:298-347): 2x2 neighbours concatenated in the order (0,0),(1,0),(0,1),(1,1),
    LN(4C), Linear(4C -> 2C, no bias)."""
    b, n, c = x.shape
    t = x.view(b, h, w, c)
    t = torch.cat([t[:, 0::2, 0::2], t[:, 1::2, 0::2], t[:, 0::2, 1::2], t[:, 1::2, 1::2]], -1).view(b, -1, 4 * c)
    t = _r(F.layer_norm(t, (4 * c,), P[pre + "norm.weight"], P[pre + "norm.bias"], LN_EPS))
    return _lin(t, P[pre + "reduction.weight"])


def swin_pooled(pixels: torch.Tensor, P: dict) -> torch.Tensor:
    """pixels [B,3,224,224] fp32 -> pooler_output [B,768] (SwinModel.forward, modeling_swin.py:849-900)."""
    x = F.conv2d(_r(pixels), _r(P["embeddings.patch_embeddings.projection.weight"]),
                 P["embeddings.patch_embeddings.projection.bias"], stride=4)
    b, c, h, w = x.shape
    x = x.flatten(2).transpose(1, 2)
    x = F.layer_norm(x, (c,), P["embeddings.norm.weight"], P["embeddings.norm.bias"], LN_EPS)
    for s, depth in enumerate(DEPTHS):
        for blk in range(depth):
            x = _block(x, P, f"encoder.layers.{s}.blocks.{blk}.", h, w, HEADS[s], 0 if blk % 2 == 0 else WINDOW // 2)
        if s < len(DEPTHS) - 1:
            x = _merge(x, P, f"encoder.layers.{s}.downsample.", h, w)
            h, w = h // 2, w // 2
    x = F.layer_norm(x, (x.shape[-1],), P["layernorm.weight"], P["layernorm.bias"], LN_EPS)
    return x.mean(1)


def resize224(img: torch.Tensor) -> torch.Tensor:
    """ref losses.py:53-54: bicubic, align_corners=True, no antialiasing, raw [0,1] values (no mean/std)."""
    return F.interpolate(img, mode="bicubic", size=(IMG, IMG), align_corners=True)


def encode_image(img: torch.Tensor, P: dict) -> torch.Tensor:
    """img [B,3,H,W] in [0,1] -> L2-normalised embedding [B,512] (ref losses.py:53,68,71)."""
    e = F.linear(swin_pooled(resize224(img.float()), P), P["projection_head.weight"])
    return e / e.norm(dim=-1, keepdim=True)


def image_logits(img: torch.Tensor, P: dict, text_feature: torch.Tensor) -> torch.Tensor:
    """cosine logit of every image against one text feature [512] (ref losses.py:73-77)."""
    t = text_feature.float() / text_feature.float().norm()
    return encode_image(img, P) @ t


def semantic_distance(sr: torch.Tensor, hr: torch.Tensor, P: dict, text_feature: torch.Tensor, n_patches: int = 3):
    """|logit_sr - logit_hr| / N per image pair (ref losses.py:79).  The reference's loop at :67-69 encodes every
    element of [resized image] + [N-1 random 224x224 crops] but keeps only the LAST one's embedding; which 224x224
    view is fed is the caller's choice here (a 224x224 input passes through the align_corners resize unchanged)."""
    return (image_logits(sr, P, text_feature) - image_logits(hr, P, text_feature)).abs() / float(n_patches)
