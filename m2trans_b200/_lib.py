"""ctypes binding of include/m2trans_b200.h (the engine's C ABI).

There is no fallback: if the shared library cannot be loaded the import of the
engine fails loudly, and every call checks the returned status and raises with
the library's own error message.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libm2trans_b200.so")

M2T_OK = 0
VAR_DEFAULT = 0
VAR_SIMT_CONV = 1 << 0
VAR_SIMT_QKV = 1 << 1
VAR_SIMT_TAIL = 1 << 2
VAR_SIMT_ATTN = 1 << 3
VAR_SIMT_ALL = 0xF
VAR_UNFUSED_TAIL = 1 << 4
VAR_PRECISE_ON, VAR_PRECISE_OFF = 1 << 5, 1 << 6
VAR_SPLIT_QKV16 = 1 << 7
VAR_SPLIT_QKV = 1 << 8
VAR_TILE_TAIL = 1 << 9
VAR_AZ_PAIRED = 1 << 10
VAR_W2_PAIR = 1 << 11
PHASE_HEAD, PHASE_BODY, PHASE_TAIL, PHASE_ALL = 1, 2, 4, 7


class M2TError(RuntimeError):
    pass


class m2t_cfg(C.Structure):
    _fields_ = [
        ("scale", C.c_int32), ("n_feats", C.c_int32), ("n_blocks", C.c_int32), ("colors", C.c_int32),
        ("batch", C.c_int32), ("height", C.c_int32), ("width", C.c_int32), ("variant", C.c_uint32),
        ("rgb_range", C.c_float),
    ]


_vp, _i, _u32, _u64, _sz, _f = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.c_size_t, C.c_float

# name -> (restype, argtypes): every symbol include/m2trans_b200.h declares
SIGNATURES = {
    "m2t_query_device": (_i, [C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "m2t_last_error": (C.c_char_p, []),
    "m2t_version": (C.c_char_p, []),
    "m2t_num_params": (_i, [_i, _i]),
    "m2t_packed_weight_bytes": (_sz, [_i, _i]),
    "m2t_pack_weights": (_i, [_i, _i, C.POINTER(_vp), _i, _vp, _vp]),
    "m2t_packed_offset": (_sz, [_i, _i, C.c_char_p]),
    "m2t_plan_create": (_i, [C.POINTER(m2t_cfg), C.POINTER(_vp)]),
    "m2t_plan_destroy": (None, [_vp]),
    "m2t_workspace_bytes": (_sz, [_vp]),
    "m2t_plan_padded": (_i, [_vp, C.POINTER(_i), C.POINTER(_i)]),
    "m2t_plan_num_launches": (_i, [_vp]),
    "m2t_workspace_offset": (_sz, [_vp, C.c_char_p]),
    "m2t_forward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "m2t_forward_phases": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _u32]),
    "m2t_stage_head": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "m2t_stage_stats_finalize": (_i, [_vp, _vp, _i, _i, _vp]),
    "m2t_stage_branch_prep": (_i, [_i, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "m2t_stage_branch_post": (_i, [_i, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "m2t_stage_qkv": (_i, [_u32, _vp, _vp, _vp, _i, _i, _vp]),
    "m2t_stage_attn": (_i, [_u32, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "m2t_stage_attn_z": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "m2t_stage_ffconv": (_i, [_u32, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "m2t_tail_scratch_bytes": (_sz, [_i, _i, _i, _i]),
    "m2t_stage_tail": (_i, [_u32, _i, _i, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp, _vp]),
    "m2t_transblock_workspace_bytes": (_sz, [_i, _i, _i]),
    "m2t_transblock_forward": (_i, [_vp, _vp, C.POINTER(_vp), _i, _i, _i, _i, _i, _vp, _vp]),
    "m2t_rlutrans_attention": (_i, [_vp, _vp, C.POINTER(_vp), _i, _i, _i, _i, _i, _vp, _vp]),
    "m2t_rlutrans_mlp": (_i, [_vp, _vp, C.POINTER(_vp), _i, C.c_long, _i, _i, _vp]),
    "m2t_clip_param_count": (_i, []),
    "m2t_clip_packed_bytes": (_sz, []),
    "m2t_clip_workspace_bytes": (_sz, [_i]),
    "m2t_clip_pack_weights": (_i, [C.POINTER(_vp), _i, _vp, _vp]),
    "m2t_clip_encode_image": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "m2t_clip_stage_linear": (_i, [_i, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "m2t_clip_stage_mlp": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "m2t_clip_stage_resize": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "m2t_debug_lin_timing": (_i, [C.POINTER(C.c_longlong)]),
    "m2t_clip_stage_layernorm": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "m2t_clip_stage_attention": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "m2t_metrics_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "m2t_eval_psnr_ssim": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp]),
    "m2t_gmsd_workspace_bytes": (_sz, [_i, _i, _i]),
    "m2t_eval_gmsd": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _vp]),
    "m2t_u8hwc_to_f32chw": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _vp]),
    "m2t_f32chw_to_u8hwc": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _vp]),
    "m2t_probe_umma": (_i, [_vp, _u32, _vp, _u32, _u64, _u64, _u32, _u32, _i, _u32, _i, _vp, _vp]),
    "m2t_debug_attn_timing": (_i, [C.POINTER(C.c_longlong)]),
    "m2t_debug_az_timing": (_i, [C.POINTER(C.c_longlong)]),
    "m2t_debug_profile_forward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_char_p, _sz]),
    "m2t_probe_tma": (_i, [_vp, _i, _i, C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u32), _i,
                           C.POINTER(C.c_int32), _vp, _u32, _vp]),
}

_lock = threading.Lock()
_lib = None


def load() -> C.CDLL:
    """Load the engine library (building nothing: use `python -m m2trans_b200.build`)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"m2trans_b200: {LIB_PATH} is missing. Build it with `python -m m2trans_b200.build` "
                "(nvcc, sm_100a). There is no CPU or PyTorch fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def last_error() -> str:
    return load().m2t_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != M2T_OK:
        raise M2TError(f"{what} failed (code {rc}): {last_error()}")
