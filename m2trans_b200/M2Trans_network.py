"""Drop-in mirror of the reference's `models/M2Trans_network.py` for the inference path.

Same public surface as the reference file (ref M2Trans_network.py):
  create_model(args) :12, M2Trans :16 (forward :58, check_image_size :78, load_state_dict :88),
  CFTM :114, DWT :198, IWT :214, TBlock :267, MeanShift :370
with identical parameter names, shapes, dtypes, registration order and initialisation, so
`checkpoints/model_x{2,3,4}.pt` load unchanged (also through nn.DataParallel, ref test.py:68-70).

The arithmetic is NOT torch: `forward` hands raw device pointers to the sm_100a engine
(include/m2trans_b200.h) through ctypes.  Inputs must be CUDA fp32 tensors on a B200; anything
else raises -- there is no CPU or PyTorch fallback.  Inference only (no autograd graph).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import threading
from typing import Dict, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.nn.init as init

from . import _lib
from ._lib import M2TError, m2t_cfg

__all__ = ["create_model", "M2Trans", "CFTM", "DWT", "IWT", "TBlock", "MeanShift", "M2TError"]


def create_model(args):
    """Plugin hook (ref :12; called by ref train.py:69-70 through utils.import_module)."""
    return M2Trans(args)


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _require_cuda_f32(x: torch.Tensor, what: str) -> None:
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise M2TError(f"{what}: expected a CUDA tensor; the B200 engine has no CPU path")
    if x.dtype != torch.float32:
        raise M2TError(f"{what}: expected float32 (the reference's dtype), got {x.dtype}")


class MeanShift(nn.Conv2d):
    """Frozen 1x1 conv kept only for state_dict compatibility: the reference constructs
    sub_mean/add_mean (ref :30-31, :370-379) but never calls them in forward (ref :58-76)."""

    def __init__(self, rgb_range, rgb_mean=(0.4488, 0.4371, 0.4040), rgb_std=(1.0, 1.0, 1.0), sign=-1):
        super().__init__(3, 3, kernel_size=1)
        std = torch.tensor(rgb_std, dtype=torch.float32)
        self.weight.data = torch.eye(3).view(3, 3, 1, 1) / std.view(3, 1, 1, 1)
        self.bias.data = sign * rgb_range * torch.tensor(rgb_mean, dtype=torch.float32) / std
        for p in self.parameters():
            p.requires_grad = False


class DWT(nn.Module):
    """Haar analysis (ref :198-212).  Inside M2Trans the transform is fused into the branch
    kernels; this standalone module evaluates the same butterflies with torch slicing."""

    def forward(self, x):
        a, b = x[:, :, 0::2, 0::2], x[:, :, 1::2, 0::2]
        c, d = x[:, :, 0::2, 1::2], x[:, :, 1::2, 1::2]
        return torch.cat((0.5 * (a + b + c + d), 0.5 * (-a - b + c + d),
                          0.5 * (-a + b - c + d), 0.5 * (a - b - c + d)), 1)


class IWT(nn.Module):
    """Haar synthesis (ref :214-237); device-agnostic (the reference hard-codes .cuda(), ref :223)."""

    def forward(self, x):
        n, c4, h, w = x.shape
        c = c4 // 4
        ll, hl, lh, hh = x[:, :c], x[:, c:2 * c], x[:, 2 * c:3 * c], x[:, 3 * c:]
        out = x.new_zeros((n, c, 2 * h, 2 * w))
        out[:, :, 0::2, 0::2] = 0.5 * (ll - hl - lh + hh)
        out[:, :, 1::2, 0::2] = 0.5 * (ll - hl + lh - hh)
        out[:, :, 0::2, 1::2] = 0.5 * (ll + hl - lh - hh)
        out[:, :, 1::2, 1::2] = 0.5 * (ll + hl + lh + hh)
        return out


class TBlock(nn.Module):
    """Blocked local attention (ref :267-340) with the reference's parameters.  Only the
    configuration the reference instantiates is supported: block 8, halo 1, one head, sr 1
    (ref :119-122) and ch in {16, 64, 256}."""

    def __init__(self, ch, block_size=8, halo_size=1, num_heads=1, bias=False, sr=1):
        super().__init__()
        if (block_size, halo_size, num_heads, bias, sr) != (8, 1, 1, False, 1):
            raise M2TError("TBlock: the engine implements block 8 / halo 1 / 1 head / no bias / sr 1 "
                           "(the only configuration M2Trans constructs, ref :119-122)")
        self.block_size, self.halo_size, self.num_heads = block_size, halo_size, num_heads
        self.head_ch = ch // num_heads
        self.sr = sr
        self.rel_h = nn.Parameter(torch.randn(1, block_size + 2 * halo_size, 1, self.head_ch // 2), requires_grad=True)
        self.rel_w = nn.Parameter(torch.randn(1, 1, block_size + 2 * halo_size, self.head_ch // 2), requires_grad=True)
        self.qkv_conv = nn.Conv2d(ch, ch * 3, kernel_size=1, bias=bias)
        self.reset_parameters()

    def reset_parameters(self):                                  # ref :342-345
        init.kaiming_normal_(self.qkv_conv.weight, mode="fan_out", nonlinearity="relu")
        init.normal_(self.rel_h, 0, 1)
        init.normal_(self.rel_w, 0, 1)

    @torch.no_grad()
    def forward(self, x, variant: int = _lib.VAR_DEFAULT):
        """Standalone TBlock on the engine's qkv + attention kernels: x [B,C,h,w] fp32 CUDA with
        h, w multiples of 8 -> [B,C,h,w] fp32.  (The reference pads other sizes, ref :297-302;
        inside M2Trans that never triggers.)"""
        _require_cuda_f32(x, "TBlock.forward")
        if x.dim() != 4 or x.shape[1] not in (16, 64, 256) or x.shape[1] != 2 * self.rel_h.shape[-1]:
            raise M2TError(f"TBlock.forward: unsupported shape {tuple(x.shape)} for a block of {2 * self.rel_h.shape[-1]} channels")
        h0, w0 = x.shape[2:]
        pad_r, pad_b = (8 - w0 % 8) % 8, (8 - h0 % 8) % 8
        if pad_r or pad_b:                                   # ref :297-302: reflect pad right / bottom to a multiple of the block
            if pad_r >= w0 or pad_b >= h0:
                raise M2TError(f"TBlock.forward: reflect padding {h0}x{w0} to multiples of 8 needs pad < size (the reference raises too)")
            x = F.pad(x, (0, pad_r, 0, pad_b), mode="reflect")
        b, c, h, w = x.shape
        lib = _lib.load()
        with torch.cuda.device(x.device):
            st = _stream_ptr(x.device)
            z = x.permute(0, 2, 3, 1).contiguous().half()                       # NHWC fp16 operand
            scale = float(c) ** -0.5
            wq = self.qkv_conv.weight.detach().reshape(3 * c, c).float().clone()
            wq[:c] *= scale                                                      # ref :311, exact (power of two)
            wq = wq.half().contiguous()
            relf = torch.cat((self.rel_h.detach().reshape(10, c // 2), self.rel_w.detach().reshape(10, c // 2)), 0)
            relf = relf.float().contiguous()
            relx = torch.zeros(32, c, dtype=torch.float16, device=x.device)
            relx[:10, : c // 2] = relf[:10].half()
            relx[10:20, c // 2:] = relf[10:].half()
            qkv = torch.empty(b, h, w, 3 * c, dtype=torch.float16, device=x.device)
            o = torch.empty(b, h, w, c, dtype=torch.float16, device=x.device)
            _lib.check(lib.m2t_stage_qkv(variant, z.data_ptr(), wq.data_ptr(), qkv.data_ptr(), b * h * w, c, st),
                       "m2t_stage_qkv")
            _lib.check(lib.m2t_stage_attn(variant, c, qkv.data_ptr(), relf.data_ptr(), relx.data_ptr(), o.data_ptr(),
                                          b, h, w, st), "m2t_stage_attn")
            out = o.float().permute(0, 3, 1, 2)
            if pad_r or pad_b:
                out = out[:, :, :h0, :w0]                    # ref :339
            return out.contiguous()


class CFTM(nn.Module):
    """Parameter container of one coarse-to-fine block (ref :114-130).  Its arithmetic
    (ref :132-164) runs inside M2Trans.forward's fused kernel sequence."""

    def __init__(self, nf, block_size=8, halo_size=1, norm=True):
        super().__init__()
        if nf != 64 or not norm:
            raise M2TError("CFTM: the engine implements nf=64, norm=True (ref :37)")
        self.is_norm = norm
        self.attn1 = TBlock(nf // 4, block_size=8, halo_size=1, num_heads=1, bias=False)
        self.attn2 = TBlock(nf * 1, block_size=8, halo_size=1, num_heads=1, bias=False)
        self.attn3 = TBlock(nf * 4, block_size=8, halo_size=1, num_heads=1, bias=False)
        self.attn4 = TBlock(nf * 4, block_size=8, halo_size=1, num_heads=1, bias=False)
        self.feed_forward = nn.Sequential(nn.Conv2d(nf, nf, kernel_size=3, stride=1, padding=1, bias=True))
        self.norm = nn.InstanceNorm2d(nf)
        self.down = DWT()
        self.up = IWT()

    @torch.no_grad()
    def forward(self, x):
        """ref :132-164 on the engine, for calling a block on its own (`model.body[i](x)`): x [B,64,H,W] fp32 CUDA with H, W
        multiples of 32 (what M2Trans.forward hands its blocks after check_image_size; the two Haar levels and the 8x8
        attention blocks need it) -> same shape.  Runs the same kernels as M2Trans.forward through a one-block plan: the
        input goes into the plan's fp32 NHWC residual stream, its InstanceNorm sums are formed here (inside M2Trans the
        producing kernel's epilogue supplies them), and the block output is read back from the stream."""
        _require_cuda_f32(x, "CFTM.forward")
        if x.dim() != 4 or x.shape[1] != 64 or x.shape[2] % 32 or x.shape[3] % 32:
            raise M2TError(f"CFTM.forward: expected [B,64,H,W] with H, W multiples of 32, got {tuple(x.shape)}")
        from ._params import state_tensors
        lib = _lib.load()
        b, _, h, w = x.shape
        dev = x.device
        own = [p.detach() for p in state_tensors(self)]
        if len(own) != 14 or any(p.device != dev or p.dtype != torch.float32 for p in own):
            raise M2TError("CFTM.forward: the block's 14 parameters must be float32 on the input's device")
        with torch.cuda.device(dev):
            st = self.__dict__.setdefault("_m2t_block", {})
            key = tuple((p.data_ptr(), p._version) for p in own)
            if st.get("key") != key or st.get("dev") != dev:
                # a one-block x2 model around this block: head / tail / mean-shift tensors are never touched by the body
                z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=dev)
                full = [z(3, 3, 1, 1), z(3), z(3, 3, 1, 1), z(3), z(64, 3, 3, 3), z(64)] + [p.contiguous() for p in own] + \
                       [z(256, 64, 1, 1), z(256), z(3, 64, 3, 3)]
                ptrs = (C.c_void_p * len(full))(*[t.data_ptr() for t in full])
                packed = torch.empty(lib.m2t_packed_weight_bytes(2, 1) + 256, dtype=torch.uint8, device=dev)
                _lib.check(lib.m2t_pack_weights(2, 1, ptrs, len(full), _aligned_ptr(packed), _stream_ptr(dev)), "m2t_pack_weights")
                st.update(key=key, dev=dev, packed=packed, plans={})
            plan = st["plans"].get((b, h, w))
            if plan is None:
                cfg = m2t_cfg(2, 64, 1, 3, b, h, w, _lib.VAR_DEFAULT, 1.0)
                handle = C.c_void_p()
                _lib.check(lib.m2t_plan_create(C.byref(cfg), C.byref(handle)), "m2t_plan_create")
                ws = torch.empty(lib.m2t_workspace_bytes(handle) + 256, dtype=torch.uint8, device=dev)
                if len(st["plans"]) >= 4:
                    lib.m2t_plan_destroy(st["plans"].pop(next(iter(st["plans"])))["handle"])
                plan = st["plans"][(b, h, w)] = {"handle": handle, "ws": ws}
            ws, base = plan["ws"], _aligned_offset(plan["ws"])

            def view(name, nbytes, dtype):
                off = base + lib.m2t_workspace_offset(plan["handle"], name.encode())
                return ws[off: off + nbytes].view(dtype)
            n = b * h * w * 64
            xin = x.permute(0, 2, 3, 1).contiguous()                                   # NHWC
            view("res", n * 4, torch.float32).copy_(xin.reshape(-1))
            stats = view("stats", 2 * b * 64 * 2 * 8, torch.float64).view(2, b, 64, 2)
            xd = xin.double().reshape(b, h * w, 64)
            stats.zero_()
            stats[0, :, :, 0] = xd.sum(1)
            stats[0, :, :, 1] = (xd * xd).sum(1)
            _lib.check(lib.m2t_forward_phases(plan["handle"], _aligned_ptr(st["packed"]), None, None, _aligned_ptr(ws),
                                              _stream_ptr(dev), _lib.PHASE_BODY), "m2t_forward_phases")
            out = view("x", n * 4, torch.float32).view(b, h, w, 64).permute(0, 3, 1, 2).contiguous()
        return out


_MAX_PLANS = 12


class _DeviceState:
    """Per-device engine state shared by DataParallel replicas: packed weights + plans."""

    def __init__(self):
        self.packed = None          # uint8 tensor
        self.packed_key = None      # (data_ptr, _version) of every parameter at pack time
        self.plans: Dict[Tuple[int, int, int, int], dict] = {}


class _EngineState:
    """Lock + per-device state.  Copies (deepcopy / pickle of the module) start empty."""

    def __init__(self):
        self.lock = threading.Lock()
        self.per_device: Dict[int, _DeviceState] = {}

    def __deepcopy__(self, memo):
        return _EngineState()

    def __reduce__(self):
        return (_EngineState, ())


class M2Trans(nn.Module):
    def __init__(self, args):
        super().__init__()
        n_feats = args.n_feats
        self.scale = args.scale
        self.window_sizes = [8, 16, 32]
        self.rgb_range = args.rgb_range
        self.n_blocks = args.n_blocks
        self.colors = getattr(args, "colors", 3)
        if n_feats != 64 or self.colors != 3 or self.scale not in (2, 3, 4):
            raise M2TError("M2Trans: the engine implements n_feats=64, colors=3, scale in {2,3,4} "
                           "(every configs/M2Trans_x*.yml of the reference)")
        self.kernel_variant = int(getattr(args, "kernel_variant", _lib.VAR_DEFAULT))
        # replay the forward's launch sequence from a CUDA graph (set False, or M2T_CUDA_GRAPH=0, for eager launches)
        self.cuda_graph = bool(getattr(args, "cuda_graph", os.environ.get("M2T_CUDA_GRAPH", "1") != "0"))
        # concurrent image groups inside the graph (0 = automatic, see _image_groups)
        self.image_groups = int(getattr(args, "image_groups", os.environ.get("M2T_IMAGE_GROUPS", "0")))

        rgb_mean = (0.4488, 0.4371, 0.4040)
        rgb_std = (1.0, 1.0, 1.0)
        self.sub_mean = MeanShift(args.rgb_range, rgb_mean, rgb_std, -1)
        self.add_mean = MeanShift(args.rgb_range, rgb_mean, rgb_std, 1)
        self.head = nn.Conv2d(self.colors, n_feats, kernel_size=3, bias=True, stride=1, padding=1,
                              padding_mode="reflect")
        self.body = nn.ModuleList([CFTM(nf=n_feats, block_size=8, halo_size=1, norm=True)
                                   for _ in range(args.n_blocks)])
        if self.scale == 4:
            self.tail = nn.Sequential(
                nn.Conv2d(n_feats, n_feats * 4, kernel_size=1, bias=True, padding_mode="reflect"),
                nn.PixelShuffle(2), nn.GELU(),
                nn.Conv2d(n_feats, n_feats * 4, kernel_size=1, bias=True, padding_mode="reflect"),
                nn.PixelShuffle(2), nn.GELU(),
                nn.Conv2d(n_feats, 3, kernel_size=3, bias=False, stride=1, padding=1, padding_mode="reflect"))
        else:
            self.tail = nn.Sequential(
                nn.Conv2d(n_feats, n_feats * self.scale * self.scale, kernel_size=1, bias=True,
                          padding_mode="reflect"),
                nn.PixelShuffle(self.scale), nn.GELU(),
                nn.Conv2d(n_feats, 3, kernel_size=3, bias=False, stride=1, padding=1, padding_mode="reflect"))
        # engine state; plain attributes so that DataParallel replicas share them by reference
        self._m2t = _EngineState()

    # ------------------------------------------------------------------ reference surface
    def check_image_size(self, x):
        """ref :78-86 -- reflect pad right/bottom to a multiple of lcm(8,16,32)=32.  forward()
        does this inside the head kernel; the method is kept for API compatibility."""
        _, _, h, w = x.size()
        wsize = self.window_sizes[0]
        for i in range(1, len(self.window_sizes)):
            wsize = wsize * self.window_sizes[i] // math.gcd(wsize, self.window_sizes[i])
        return F.pad(x, (0, (wsize - w % wsize) % wsize, 0, (wsize - h % wsize) % wsize), "reflect")

    def load_state_dict(self, state_dict, strict=False):
        """Same contract as the reference override (ref :88-112): copy by name; a shape mismatch
        is tolerated (with the reference's message) only for `tail` entries; with strict=True
        unexpected non-tail keys and missing keys raise KeyError.  Default strict=False, as in
        the reference (nn.DataParallel.load_state_dict bypasses this, ref test.py:70)."""
        own = self.state_dict()
        for name, param in state_dict.items():
            if name in own:
                if isinstance(param, nn.Parameter):
                    param = param.data
                try:
                    own[name].copy_(param)
                except Exception:
                    if "tail" in name:
                        print("Replace pre-trained upsampler to new one...")
                    else:
                        raise RuntimeError(
                            "While copying the parameter named {}, whose dimensions in the model are {} and "
                            "whose dimensions in the checkpoint are {}.".format(name, own[name].size(), param.size()))
            elif strict and "tail" not in name:
                raise KeyError('unexpected key "{}" in state_dict'.format(name))
        self.invalidate()                                   # own[name].copy_() bumps _version; belt and braces
        if strict:
            missing = set(own.keys()) - set(state_dict.keys())
            if missing:
                raise KeyError('missing keys in state_dict: "{}"'.format(missing))

    # ------------------------------------------------------------------ engine plumbing
    def _param_list(self):
        # registration order == reference state_dict order.  The (owner dict, name) slots are
        # resolved once: walking state_dict() through the 125 submodules cost ~0.15 ms of host time per forward.  The
        # tensors are looked up through the slots on every call, so replaced Parameters (.to(), .half()) are seen.
        # The cache remembers which module it belongs to: DataParallel replicas are shallow __dict__ copies whose
        # parameter dicts are new objects, so a replica rebuilds its own slots.
        cached = self.__dict__.get("_m2t_slots")
        slots = cached[1] if cached is not None and cached[0] == id(self) else None
        if slots is None:
            # A DataParallel replica (torch.nn.parallel.replicate) has an EMPTY _parameters dict: its broadcast copies
            # are plain attributes, listed per module in _former_parameters in registration order.
            replica = "_former_parameters" in self.__dict__
            slots, got = [], []
            for prefix, mod in self.named_modules():
                if replica:
                    src = mod.__dict__.get("_former_parameters", {})
                    own = [(src, n) for n in src]
                else:
                    own = [(mod._parameters, n) for n, v in mod._parameters.items() if v is not None]
                own += [(mod._buffers, n) for n, v in mod._buffers.items()
                        if v is not None and n not in mod._non_persistent_buffers_set]
                slots += own
                got += [(prefix + "." if prefix else "") + n for _, n in own]
            if not replica:
                want = list(self.state_dict(keep_vars=True))
                if got != want:                           # never expected; fall back to the slow, always-correct walk
                    return [p for _, p in self.state_dict(keep_vars=True).items()]
            self.__dict__["_m2t_slots"] = (id(self), slots)
        return [d[name] for d, name in slots]

    def _device_state(self, device: torch.device) -> _DeviceState:
        idx = device.index if device.index is not None else torch.cuda.current_device()
        st = self._m2t.per_device.get(idx)
        if st is None:
            st = self._m2t.per_device[idx] = _DeviceState()
        return st

    def _packed_weights(self, st: _DeviceState, device) -> torch.Tensor:
        params = self._param_list()
        # Replicas get fresh broadcast copies on every DataParallel call; the caching allocator may hand out the same
        # addresses again with _version 0 even after the source weights changed, so (data_ptr, _version) proves nothing
        # there: a replica re-packs on every forward, into the same buffer so that the captured graph stays valid.
        replica = "_former_parameters" in self.__dict__
        key = None if replica else tuple((p.data_ptr(), p._version) for p in params)
        if not replica and st.packed is not None and st.packed_key == key:
            return st.packed
        lib = _lib.load()
        for p in params:
            if p.device != device or p.dtype != torch.float32:
                raise M2TError(f"parameters must be float32 on {device} (found {p.dtype} on {p.device}); "
                               "move the model with .to(device) before calling it")
        n = len(params)
        want = lib.m2t_num_params(self.scale, self.n_blocks)
        if n != want:
            raise M2TError(f"state_dict has {n} tensors, expected {want}")
        keep = [p.detach().contiguous() for p in params]
        ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in keep])
        nbytes = lib.m2t_packed_weight_bytes(self.scale, self.n_blocks)
        if replica and st.packed is not None and st.packed.numel() == nbytes + 256:
            packed = st.packed                                    # same buffer: graphs captured on it stay valid
        else:
            packed = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
        _lib.check(lib.m2t_pack_weights(self.scale, self.n_blocks, ptrs, n, _aligned_ptr(packed), _stream_ptr(device)),
                   "m2t_pack_weights")
        st.packed, st.packed_key = packed, key
        return packed

    def _plan(self, st: _DeviceState, b: int, h: int, w: int, device, group: int = 0) -> dict:
        key = (b, h, w, self.kernel_variant, group)
        plan = st.plans.get(key)
        if plan is None:
            lib = _lib.load()
            cfg = m2t_cfg(self.scale, 64, self.n_blocks, self.colors, b, h, w, self.kernel_variant,
                          float(self.rgb_range))
            handle = C.c_void_p()
            _lib.check(lib.m2t_plan_create(C.byref(cfg), C.byref(handle)), "m2t_plan_create")
            ws = torch.empty(lib.m2t_workspace_bytes(handle) + 256, dtype=torch.uint8, device=device)
            plan = {"handle": handle, "ws": ws, "launches": lib.m2t_plan_num_launches(handle), "uses": 0}
            while len(st.plans) >= _MAX_PLANS:          # bound the workspace cache: evict the least recently used
                old = st.plans.pop(next(iter(st.plans)))
                lib.m2t_plan_destroy(old["handle"])
        else:
            del st.plans[key]                           # re-insert: dict order = recency
        st.plans[key] = plan
        return plan

    def invalidate(self) -> None:
        """Drop the packed weights and captured graphs; the next forward re-packs from the current parameters.
        Needed after IN-PLACE edits through `.data` (p.data.mul_(2), p.data.copy_(...)): those keep the tensor's
        address and version counter, so the (data_ptr, _version) cache key cannot see them.  load_state_dict, .to(),
        .cuda(), .half()/.float() and optimiser-style in-place ops on the Parameters themselves are detected."""
        with self._m2t.lock:
            for st in self._m2t.per_device.values():
                st.packed_key = None
                for plan in st.plans.values():
                    plan.pop("graph", None)
                    plan.pop("graph_key", None)

    repack = invalidate

    def _apply(self, fn, *args, **kwargs):              # .to() / .cuda() / .float(): parameters are replaced or rewritten
        out = super()._apply(fn, *args, **kwargs)
        if "_m2t" in self.__dict__:
            self.invalidate()
        return out

    def _image_groups(self, b: int, h: int, w: int) -> int:
        """Frames are independent (InstanceNorm is per image), so a batch may run as G concurrent chains of b/G frames
        on G streams inside one CUDA graph (args.image_groups / M2T_IMAGE_GROUPS).  Measured on B200 it does not pay:
        cfg2 1.414 ms with one chain, 1.450 with two, 1.63 with four (the big-shared-memory kernels of two chains
        cannot share an SM, so the chains mostly serialise and pay twice the fixed costs); cfg3 gains 3 %.  Default 1."""
        g = self.image_groups
        return max(1, min(g if g > 0 else 1, b))

    @torch.no_grad()
    def forward(self, x):
        """ref :58-76.  x [B,3,H,W] fp32 CUDA in [0, rgb_range] -> [B,3,H*s,W*s] fp32."""
        _require_cuda_f32(x, "M2Trans.forward")
        if x.dim() != 4 or x.shape[1] != self.colors:
            raise M2TError(f"M2Trans.forward: expected [B,{self.colors},H,W], got {tuple(x.shape)}")
        lib = _lib.load()
        x = x.contiguous()
        b, _, h, w = x.shape
        device = x.device
        with torch.cuda.device(device), self._m2t.lock:
            st = self._device_state(device)
            packed = self._packed_weights(st, device)
            s = self.scale
            replica = "_former_parameters" in self.__dict__
            stream = torch.cuda.current_stream(device)
            y = torch.empty((b, 3, h * s, w * s), dtype=torch.float32, device=device)
            # image groups: contiguous chunks of the batch, each with its own plan / workspace
            ng = 1 if (replica or not self.cuda_graph) else self._image_groups(b, h, w)
            bounds = [(b * i) // ng for i in range(ng + 1)]
            plans = [self._plan(st, bounds[i + 1] - bounds[i], h, w, device, group=i) for i in range(ng)]
            self.last_launches = sum(p["launches"] for p in plans)
            in_px, out_px = 3 * h * w, 3 * h * s * w * s

            def phases(plan, mask, gi):
                xin = x.data_ptr() + 4 * in_px * bounds[gi] if mask & _lib.PHASE_HEAD else None
                yout = y.data_ptr() + 4 * out_px * bounds[gi] if mask & _lib.PHASE_TAIL else None
                _lib.check(lib.m2t_forward_phases(plan["handle"], _aligned_ptr(packed), xin, yout,
                                                  _aligned_ptr(plan["ws"]), _stream_ptr(device), mask), "m2t_forward_phases")

            # A plan's workspace is shared by every forward of that shape: a forward issued on another stream must wait
            # for the previous user of the workspace (multi-stream pipelining, per-thread default streams).
            for plan in plans:
                last = plan.get("last")
                if last is not None and last[0] != stream.cuda_stream:
                    stream.wait_event(last[1])
            head = plans[0]
            head["uses"] += 1
            # DataParallel replicas run concurrently in one thread per device and re-pack on every call: they launch
            # eagerly.  So does the FIRST forward of a shape (one-off shapes, e.g. an eval loop over frames of varying
            # size, never pay for a capture), and any forward issued while the caller is capturing a graph itself.
            eager = (not self.cuda_graph or replica or torch.cuda.is_current_stream_capturing()
                     or (head.get("graph_key") != packed.data_ptr() and head["uses"] < 2))
            if not eager and head.get("graph_key") != packed.data_ptr():
                # CUDA-graph capture of the CFTM blocks only: they touch nothing but the plan's workspace and the packed
                # weights, so the graph is independent of the caller's tensors.  The head conv (reads x) and the tail
                # (writes y) are launched directly on the caller's tensors around the replay.
                try:
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                        main = torch.cuda.current_stream(device)
                        if ng > 1:
                            fork = torch.cuda.Event()
                            fork.record(main)
                            side = head.setdefault("side_streams", [torch.cuda.Stream(device) for _ in range(ng - 1)])
                            joins = []
                            for gi in range(1, ng):
                                side[gi - 1].wait_event(fork)
                                with torch.cuda.stream(side[gi - 1]):
                                    phases(plans[gi], _lib.PHASE_BODY, gi)
                                    ev = torch.cuda.Event()
                                    ev.record(side[gi - 1])
                                    joins.append(ev)
                        phases(plans[0], _lib.PHASE_BODY, 0)
                        for ev in (joins if ng > 1 else []):
                            main.wait_event(ev)
                    head.update(graph=graph, graph_key=packed.data_ptr())
                except Exception:                        # capture refused (another capture in flight, ...): stay eager
                    head.pop("graph", None)
                    head["graph_key"] = None
                    head["uses"] = 0
                    eager = True
            if eager:
                for gi, plan in enumerate(plans):
                    phases(plan, _lib.PHASE_ALL, gi)
            else:
                for gi, plan in enumerate(plans):
                    phases(plan, _lib.PHASE_HEAD, gi)
                head["graph"].replay()
                for gi, plan in enumerate(plans):
                    phases(plan, _lib.PHASE_TAIL, gi)
            if not torch.cuda.is_current_stream_capturing():
                done = torch.cuda.Event()
                done.record(stream)
                for plan in plans:
                    plan["last"] = (stream.cuda_stream, done)
            return y

    @torch.no_grad()
    def profile_forward(self, x) -> str:
        """Development aid: one eager forward with a CUDA event after every launch; returns a per-kernel table
        (m2t_debug_profile_forward).  Unlike ncu it times the kernels with the caches as their predecessor left them."""
        _require_cuda_f32(x, "M2Trans.profile_forward")
        lib = _lib.load()
        x = x.contiguous()
        b, _, h, w = x.shape
        device = x.device
        with torch.cuda.device(device), self._m2t.lock:
            st = self._device_state(device)
            packed = self._packed_weights(st, device)
            plan = self._plan(st, b, h, w, device)
            y = torch.empty((b, 3, h * self.scale, w * self.scale), dtype=torch.float32, device=device)
            buf = C.create_string_buffer(1 << 14)
            _lib.check(lib.m2t_debug_profile_forward(plan["handle"], _aligned_ptr(packed), x.data_ptr(), y.data_ptr(),
                                                     _aligned_ptr(plan["ws"]), _stream_ptr(device), buf, len(buf)),
                       "m2t_debug_profile_forward")
        return buf.value.decode()

    def engine_tensor(self, x_shape, name: str) -> torch.Tensor:
        """Test hook: copy of an internal NHWC tensor ('res', 'x' fp32; 'y' fp16) of the plan(s) that served the
        last forward of shape `x_shape` on the current device (one plan per image group, concatenated)."""
        lib = _lib.load()
        b, _, h, w = x_shape
        st = self._device_state(torch.device("cuda", torch.cuda.current_device()))
        replica = "_former_parameters" in self.__dict__
        ng = 1 if (replica or not self.cuda_graph) else self._image_groups(b, h, w)
        parts = []
        for gi in range(ng):
            bg = (b * (gi + 1)) // ng - (b * gi) // ng
            plan = st.plans[(bg, h, w, self.kernel_variant, gi)]
            hp, wp = C.c_int(), C.c_int()
            _lib.check(lib.m2t_plan_padded(plan["handle"], C.byref(hp), C.byref(wp)), "m2t_plan_padded")
            off = lib.m2t_workspace_offset(plan["handle"], name.encode())
            if off == C.c_size_t(-1).value:
                raise KeyError(name)
            base = _aligned_offset(plan["ws"]) + off
            n = bg * hp.value * wp.value * 64
            if name == "y":
                parts.append(plan["ws"][base: base + n * 2].view(torch.float16).view(bg, hp.value, wp.value, 64))
            else:
                parts.append(plan["ws"][base: base + n * 4].view(torch.float32).view(bg, hp.value, wp.value, 64))
        return torch.cat(parts, 0) if len(parts) > 1 else parts[0]


def _aligned_offset(t: torch.Tensor) -> int:
    return (-t.data_ptr()) % 256


def _aligned_ptr(t: torch.Tensor) -> int:
    return t.data_ptr() + _aligned_offset(t)
