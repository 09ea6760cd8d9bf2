"""Argument validation of the later C entry points (image pass, metrics, loader) happens before any device work, so it can
be checked without a GPU: bad arguments return M2T_E_ARG with a message, and the size queries answer on the CPU."""
import ctypes as C

from m2trans_b200 import _lib

E_ARG = -1


def _err(lib):
    return lib.m2t_last_error().decode()


def test_clip_sizes_and_argument_errors():
    lib = _lib.load()
    assert lib.m2t_clip_param_count() == 220
    packed = lib.m2t_clip_packed_bytes()
    assert 2 * 27_500_000 < packed < 2 * 27_500_000 + 8_000_000        # bf16 matrices + fp32 vectors, tables, projection
    per_image = lib.m2t_clip_workspace_bytes(2) - lib.m2t_clip_workspace_bytes(1)
    assert 7_000_000 < per_image < 8_500_000 and lib.m2t_clip_workspace_bytes(0) == 0
    ptrs = (C.c_void_p * 220)()
    assert lib.m2t_clip_pack_weights(ptrs, 219, 1, None) == E_ARG and "220" in _err(lib)
    assert lib.m2t_clip_pack_weights(ptrs, 220, 1, None) == E_ARG and "null" in _err(lib)
    assert lib.m2t_clip_encode_image(None, 1, 1, 224, 224, 1, None, None, 1, None) == E_ARG
    assert lib.m2t_clip_encode_image(1, 1, 1, 224, 224, 1, 1, None, 1, None) == E_ARG and "together" in _err(lib)
    assert lib.m2t_clip_encode_image(1, 1, 0, 224, 224, 1, None, None, 1, None) == E_ARG
    assert lib.m2t_clip_stage_resize(1, 1, 1, 1, 224, None) == E_ARG
    assert lib.m2t_clip_stage_layernorm(1, 1, 1, 1, 1, 7, 7, 96, 1, None) == E_ARG and "even" in _err(lib)


def test_metrics_and_loader_argument_errors():
    lib = _lib.load()
    assert lib.m2t_metrics_workspace_bytes(1, 64, 64, 4) == (2 * 2 + 1) * 16       # 46 x 46 valid -> 2 x 2 tiles, + 1 image sum
    assert lib.m2t_metrics_workspace_bytes(1, 18, 64, 4) == 0                      # 10 rows left: below the 11-tap window
    assert lib.m2t_eval_psnr_ssim(None, 1, 1, 3, 64, 64, 4, 1.0, 1, 1, None) == E_ARG
    assert lib.m2t_eval_psnr_ssim(1, 1, 1, 2, 64, 64, 4, 1.0, 1, 1, None) == E_ARG and "colors" in _err(lib)
    assert lib.m2t_eval_psnr_ssim(1, 1, 1, 3, 18, 64, 4, 1.0, 1, 1, None) == E_ARG and "window" in _err(lib)
    assert lib.m2t_u8hwc_to_f32chw(None, 1, 1, 4, 4, 3, 255.0, None) == E_ARG
    assert lib.m2t_u8hwc_to_f32chw(1, 1, 1, 4, 4, 2, 255.0, None) == E_ARG and "colors" in _err(lib)
    assert lib.m2t_u8hwc_to_f32chw(1, 1, 1, 4, 4, 3, 0.0, None) == E_ARG
