"""CPU fp32 oracle for the M2Trans forward path.  TEST INFRASTRUCTURE ONLY.

This file is a restatement, in plain torch fp32 on the CPU, of the algorithm of
the reference `models/M2Trans_network.py` (eezkni/M2Trans).  It exists so that
the CUDA path can be checked on a box where `/root/reference` is not present.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl
reference` legs of `bench.py` may import it.  The product package
`m2trans_b200` never imports anything from `oracle/`.

Parity pin: `oracle/make_golden.py` imports the real reference module (in the
build container, where `/root/reference` exists), runs it on seeded inputs with
seeded synthetic checkpoints and commits the results under `tests/golden/`;
`tests/test_oracle_golden.py` checks this file against those vectors.  The
reference has no tests or golden vectors of its own, so
reference-generated fixtures are the only pin available.

The functions work on a plain `state_dict` (reference key names, with or
without the DataParallel `module.` prefix), never on reference classes.

All citations `ref:` are into /root/reference/models/M2Trans_network.py.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

WINDOW_LCM = 32      # lcm(8, 16, 32), ref: 23, 78-83
BLOCK = 8            # TBlock block_size, ref: 119-122
HALO = 1             # TBlock halo_size,  ref: 119-122
WIN = BLOCK + 2 * HALO
IN_EPS = 1e-5        # nn.InstanceNorm2d default eps, ref: 127


def strip_module_prefix(sd: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """Checkpoints are saved from an nn.DataParallel wrapper (ref train.py:73,345)."""
    if all(k.startswith("module.") for k in sd):
        return {k[len("module."):]: v for k, v in sd.items()}
    return dict(sd)


def infer_scale(sd: Dict[str, Tensor]) -> int:
    """x4 has tail.0/tail.3/tail.6, x2/x3 have tail.0/tail.3 (ref: 40-56)."""
    if "tail.6.weight" in sd:
        return 4
    r2 = sd["tail.0.weight"].shape[0] // sd["tail.0.weight"].shape[1]
    return int(round(math.sqrt(r2)))


def pad_to_window(x: Tensor) -> Tensor:
    """ref: 78-86 check_image_size: reflect pad right/bottom to a multiple of 32."""
    h, w = x.shape[-2:]
    ph = (WINDOW_LCM - h % WINDOW_LCM) % WINDOW_LCM
    pw = (WINDOW_LCM - w % WINDOW_LCM) % WINDOW_LCM
    if ph == 0 and pw == 0:
        return x
    return F.pad(x, (0, pw, 0, ph), mode="reflect")


def conv3x3_reflect(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    """ref: 34 (head) and 48/55 (last tail conv): padding=1, padding_mode='reflect'."""
    return F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), w, b)


def instance_norm(x: Tensor) -> Tensor:
    """ref: 127,135 nn.InstanceNorm2d(nf): no affine, no running stats, biased var."""
    mu = x.mean(dim=(2, 3), keepdim=True)
    var = x.var(dim=(2, 3), unbiased=False, keepdim=True)
    return (x - mu) / torch.sqrt(var + IN_EPS)


def dwt(x: Tensor) -> Tensor:
    """ref: 203-209 Haar analysis; bands LL,HL,LH,HH concatenated along channels."""
    a = x[:, :, 0::2, 0::2]
    b = x[:, :, 1::2, 0::2]
    c = x[:, :, 0::2, 1::2]
    d = x[:, :, 1::2, 1::2]
    ll = 0.5 * (a + b + c + d)
    hl = 0.5 * (-a - b + c + d)
    lh = 0.5 * (-a + b - c + d)
    hh = 0.5 * (a - b - c + d)
    return torch.cat((ll, hl, lh, hh), dim=1)


def iwt(x: Tensor) -> Tensor:
    """ref: 219-234 Haar synthesis (the reference's `.cuda()` at :223 is a device
    placement, not arithmetic)."""
    n, c4, h, w = x.shape
    c = c4 // 4
    ll, hl, lh, hh = x[:, 0:c], x[:, c:2 * c], x[:, 2 * c:3 * c], x[:, 3 * c:4 * c]
    out = x.new_zeros((n, c, 2 * h, 2 * w))
    out[:, :, 0::2, 0::2] = 0.5 * (ll - hl - lh + hh)
    out[:, :, 1::2, 0::2] = 0.5 * (ll - hl + lh - hh)
    out[:, :, 0::2, 1::2] = 0.5 * (ll + hl - lh - hh)
    out[:, :, 1::2, 1::2] = 0.5 * (ll + hl + lh + hh)
    return out


def tblock(x: Tensor, w_qkv: Tensor, rel_h: Tensor, rel_w: Tensor) -> Tensor:
    """Blocked local ("halo") attention, ref: 290-340, with block=8, halo=1,
    heads=1, sr=1 (the only instantiation, ref: 119-122).

    x      [B,C,h,w] with h,w multiples of 8
    w_qkv  [3C,C,1,1] (no bias)     rel_h [1,10,1,C/2]     rel_w [1,1,10,C/2]
    """
    bsz, ch, h, w = x.shape
    assert h % BLOCK == 0 and w % BLOCK == 0
    nh, nw = h // BLOCK, w // BLOCK
    qkv = F.conv2d(x, w_qkv)                                             # ref: 307
    q, k, v = qkv[:, :ch], qkv[:, ch:2 * ch], qkv[:, 2 * ch:]            # ref: 308
    # queries: one 8x8 block per window, row-major inside the block      # ref: 310-311
    q = q.reshape(bsz, ch, nh, BLOCK, nw, BLOCK).permute(0, 2, 4, 3, 5, 1)
    q = q.reshape(bsz * nh * nw, BLOCK * BLOCK, ch) * (ch ** -0.5)
    # keys/values: 10x10 neighbourhood, zero outside the frame           # ref: 313-317
    def neigh(t: Tensor) -> Tensor:
        u = F.unfold(t, kernel_size=WIN, stride=BLOCK, padding=HALO)    # [B, C*100, L]
        u = u.reshape(bsz, ch, WIN * WIN, nh * nw).permute(0, 3, 2, 1)
        return u.reshape(bsz * nh * nw, WIN, WIN, ch)
    k = neigh(k)
    v = neigh(v).reshape(bsz * nh * nw, WIN * WIN, ch)
    # relative position terms are ADDED TO K (also at zero-padded keys)  # ref: 322-325
    half = ch // 2
    k = torch.cat((k[..., :half] + rel_h, k[..., half:] + rel_w), dim=-1)
    k = k.reshape(bsz * nh * nw, WIN * WIN, ch)
    sim = torch.bmm(q, k.transpose(1, 2))                                # ref: 328
    attn = torch.softmax(sim, dim=-1)                                    # ref: 329
    out = torch.bmm(attn, v)                                             # ref: 331
    out = out.reshape(bsz, nh, nw, BLOCK, BLOCK, ch).permute(0, 5, 1, 3, 2, 4)
    return out.reshape(bsz, ch, h, w)                                    # ref: 332


def _attn(sd: Dict[str, Tensor], prefix: str, x: Tensor) -> Tensor:
    return tblock(x, sd[prefix + "qkv_conv.weight"], sd[prefix + "rel_h"], sd[prefix + "rel_w"])


def cftm(sd: Dict[str, Tensor], i: int, x: Tensor) -> Tensor:
    """ref: 132-164, the is_norm=True path (the only one constructed, ref: 37)."""
    p = f"body.{i}."
    n = instance_norm(x)                                                 # ref: 135
    n1, n2, n3, n4 = torch.chunk(n, 4, dim=1)                            # ref: 137
    y1 = _attn(sd, p + "attn1.", n1) + n1                                # ref: 139
    t2 = (n2 + y1) / 2.0                                                 # ref: 141
    y2 = iwt(_attn(sd, p + "attn2.", dwt(t2))) + t2                      # ref: 143-145
    t3 = (n3 + y2) / 2.0                                                 # ref: 147
    y3 = iwt(iwt(_attn(sd, p + "attn3.", dwt(dwt(t3))))) + t3            # ref: 149-153
    t4 = (n4 + y3) / 2.0                                                 # ref: 155
    y4 = iwt(iwt(_attn(sd, p + "attn4.", dwt(dwt(t4))))) + t4            # ref: 157-161
    xc = torch.cat((y1, y2, y3, y4), dim=1)                              # ref: 163
    ff = F.conv2d(xc, sd[p + "feed_forward.0.weight"], sd[p + "feed_forward.0.bias"], padding=1)
    return ff + x                                                        # ref: 164


def tail(sd: Dict[str, Tensor], x: Tensor, scale: int) -> Tensor:
    """ref: 40-56. 1x1 conv(+bias) -> PixelShuffle -> exact GELU [x4: twice] ->
    3x3 reflect conv without bias."""
    if scale == 4:
        x = F.gelu(F.pixel_shuffle(F.conv2d(x, sd["tail.0.weight"], sd["tail.0.bias"]), 2))
        x = F.gelu(F.pixel_shuffle(F.conv2d(x, sd["tail.3.weight"], sd["tail.3.bias"]), 2))
        return conv3x3_reflect(x, sd["tail.6.weight"], None)
    x = F.gelu(F.pixel_shuffle(F.conv2d(x, sd["tail.0.weight"], sd["tail.0.bias"]), scale))
    return conv3x3_reflect(x, sd["tail.3.weight"], None)


def forward(sd: Dict[str, Tensor], x: Tensor, scale: Optional[int] = None,
            rgb_range: float = 1.0, n_blocks: Optional[int] = None,
            return_intermediates: bool = False):
    """ref: 58-76 M2Trans.forward.  x [B,3,H,W] fp32 -> [B,3,H*s,W*s] fp32."""
    sd = strip_module_prefix(sd)
    if scale is None:
        scale = infer_scale(sd)
    if n_blocks is None:
        n_blocks = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("body."))
    h, w = x.shape[-2:]
    xp = pad_to_window(x.float())                                        # ref: 61
    res = conv3x3_reflect(xp, sd["head.weight"], sd["head.bias"])        # ref: 63
    inter = {"res": res}
    y = res
    for i in range(n_blocks):                                            # ref: 67-68
        y = cftm(sd, i, y)
        if return_intermediates:
            inter[f"body{i}"] = y
    y = res + y                                                          # ref: 70
    y = tail(sd, y, scale)                                               # ref: 72
    y = torch.clamp(y, min=0.0, max=rgb_range)                           # ref: 74
    y = y[:, :, : h * scale, : w * scale]                                # ref: 76
    if return_intermediates:
        return y, inter
    return y


# --------------------------------------------------------------------------- #
# helpers shared by tests / bench (still test infrastructure)
# --------------------------------------------------------------------------- #
def psnr(a: Tensor, b: Tensor, peak: float = 1.0) -> float:
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    if mse == 0.0:
        return float("inf")
    return 10.0 * math.log10(peak * peak / mse)


def max_abs(a: Tensor, b: Tensor) -> float:
    return (a.double() - b.double()).abs().max().item()
