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
    assert lib.m2t_f32chw_to_u8hwc(None, 1, 1, 4, 4, 3, 255.0, None) == E_ARG
    assert lib.m2t_f32chw_to_u8hwc(1, 1, 1, 4, 4, 2, 255.0, None) == E_ARG and "colors" in _err(lib)
    assert lib.m2t_f32chw_to_u8hwc(1, 1, 0, 4, 4, 3, 255.0, None) == E_ARG


def test_attn_z_stage_argument_errors():
    lib = _lib.load()
    assert lib.m2t_stage_attn_z(64, None, 1, 1, 1, None, 1, 1, 16, 16, None) == E_ARG
    assert lib.m2t_stage_attn_z(128, 1, 1, 1, 1, None, 1, 1, 16, 16, None) != 0 and "C=128" in _err(lib)
    assert lib.m2t_stage_attn_z(64, 1, 1, 1, 1, None, 0, 1, 16, 16, None) == E_ARG and "branch" in _err(lib)


def test_bench_arms_print_the_same_config():
    """The driver compares the `config` of the GPU arm and of `--impl reference`: one function builds both."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(__file__)), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for name in ("cfg1", "cfg2", "cfg3", "cfg4", "cfg5"):
        cfg = bench.workload_config(name)
        assert list(cfg) == ["workload"] and cfg["workload"].startswith(name + ":")
    assert bench.WORKLOADS["cfg3"][4] == "strong" and bench.WORKLOADS["cfg4"][:2] == (4, 64) and bench.WORKLOADS["cfg2"][4] == "weak"
    from m2trans_b200.sharding import shard_range
    assert [shard_range(32, r, 8) for r in (0, 7)] == [(0, 4), (28, 32)]         # cfg3 at 8 GPUs: 4 frames each
    assert shard_range(64, 3, 8) == (24, 32) and shard_range(256, 7, 8) == (224, 256)


def test_gmsd_argument_errors():
    lib = _lib.load()
    assert lib.m2t_gmsd_workspace_bytes(1, 1, 64) == 0 and lib.m2t_gmsd_workspace_bytes(2, 37, 50) > 2 * 2 * 19 * 25 * 4
    assert lib.m2t_eval_gmsd(None, 1, 1, 3, 64, 64, 1.0, 1, 1, None) == E_ARG
    assert lib.m2t_eval_gmsd(1, 1, 1, 2, 64, 64, 1.0, 1, 1, None) == E_ARG and "colors" in _err(lib)
    assert lib.m2t_eval_gmsd(1, 1, 1, 3, 64, 64, 0.0, 1, 1, None) == E_ARG
