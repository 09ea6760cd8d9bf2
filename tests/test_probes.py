"""Hardware probes for the tcgen05 / TMA conventions the tensor-core kernels rely on (B200 only).

Each probe builds raw shared-memory images under an explicit layout hypothesis, runs ONE
tcgen05.mma chain (or one TMA box load) through `m2t_probe_umma` / `m2t_probe_tma`, and compares with
numpy.  Results are also written to gpurun_out/probe_report.json so that the build container can read
which conventions hold.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPORT = {}


def _lib():
    from m2trans_b200 import _lib
    return _lib


def swz128(off):
    return off ^ (((off >> 7) & 7) << 4)


def desc(start, lbo, sbo, layout):
    return ((start >> 4) & 0x3FFF) | (((lbo >> 4) & 0x3FFF) << 16) | (((sbo >> 4) & 0x3FFF) << 32) | (1 << 46) | (layout << 61)


def idesc(M, N, a_mn=0, b_mn=0):
    return (1 << 4) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def img_kmajor_sw128(mat):
    """mat [rows][K] fp16, K multiple of 64 -> bytes; K-block kb at kb*rows*128, row r at r*128, swizzled."""
    rows, K = mat.shape
    out = np.zeros(rows * K * 2, dtype=np.uint8)
    raw = mat.view(np.uint8).reshape(rows, K * 2)
    for kb in range(K // 64):
        for r in range(rows):
            for ch in range(8):
                off = kb * rows * 128 + swz128(r * 128 + ch * 16)
                out[off:off + 16] = raw[r, kb * 128 + ch * 16: kb * 128 + ch * 16 + 16]
    return out


def img_kmajor_interleave(mat):
    """no swizzle: [K/8][rows][8 elems]  (row pitch 16 B, 8-row groups 128 B apart, K chunks rows*16 apart)"""
    rows, K = mat.shape
    return np.ascontiguousarray(mat.reshape(rows, K // 8, 8).transpose(1, 0, 2)).view(np.uint8).reshape(-1).copy()


def img_mnmajor_sw128(mat_kn):
    """mat [K][N] fp16 (N contiguous), N multiple of 64 -> N-block nb at nb*K*128, k row at k*128, swizzled."""
    K, N = mat_kn.shape
    out = np.zeros(K * N * 2, dtype=np.uint8)
    raw = mat_kn.view(np.uint8).reshape(K, N * 2)
    for nb in range(N // 64):
        for k in range(K):
            for ch in range(8):
                off = nb * K * 128 + swz128(k * 128 + ch * 16)
                out[off:off + 16] = raw[k, nb * 128 + ch * 16: nb * 128 + ch * 16 + 16]
    return out


def run_umma(a_img, b_img, a_desc, b_desc, a_step, b_step, k_steps, idsc, n_cols):
    L = _lib()
    lib = L.load()
    pad = lambda x: np.concatenate([x, np.zeros((-len(x)) % 16, np.uint8)])
    a = torch.from_numpy(pad(a_img)).cuda()
    b = torch.from_numpy(pad(b_img)).cuda()
    out = torch.full((128, n_cols & 0xFFFF), float("nan"), dtype=torch.float32, device="cuda")
    L.check(lib.m2t_probe_umma(a.data_ptr(), a.numel(), b.data_ptr(), b.numel(), a_desc, b_desc, a_step, b_step,
                               k_steps, idsc, n_cols, out.data_ptr(), None), "m2t_probe_umma")
    torch.cuda.synchronize()
    return out.cpu().numpy()


def record(name, ok, detail=""):
    REPORT[name] = {"ok": bool(ok), "detail": detail}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "probe_report.json"), "w") as f:
        json.dump(REPORT, f, indent=1)
    print(f"[probe] {name}: {'OK' if ok else 'MISMATCH'} {detail}")


def rnd(shape, seed):
    g = np.random.default_rng(seed)
    return (g.integers(-8, 9, size=shape) / 8.0).astype(np.float16)   # exactly representable products/sums


def close(got, want):
    return got.shape == want.shape and np.allclose(got, want, atol=1e-3, rtol=0)


def test_umma_sw128_kmajor_basic():
    A, B = rnd((128, 64), 1), rnd((64, 64), 2)
    got = run_umma(img_kmajor_sw128(A), img_kmajor_sw128(B), desc(0, 16, 1024, 2), desc(0, 16, 1024, 2), 32, 32, 4,
                   idesc(128, 64), 64)
    want = A.astype(np.float32) @ B.astype(np.float32).T
    ok = close(got, want)
    record("umma_sw128_kmajor_basic", ok, f"maxdiff {np.nanmax(np.abs(got - want)):.4g}")
    assert ok


def test_umma_sw128_n_variants():
    """N = 16..256 (multiples of 16 at M=128) in one instruction."""
    A = rnd((128, 64), 3)
    for N in (16, 32, 48, 112, 192, 256):
        B = rnd((N, 64), 4 + N)
        got = run_umma(img_kmajor_sw128(A), img_kmajor_sw128(B), desc(0, 16, 1024, 2), desc(0, 16, 1024, 2), 32, 32, 4,
                       idesc(128, N), N)
        want = A.astype(np.float32) @ B.astype(np.float32).T
        ok = close(got, want)
        record(f"umma_sw128_N{N}", ok, f"maxdiff {np.nanmax(np.abs(got - want)):.4g}")
        if N % 16 == 0:
            assert ok


def test_umma_sw128_row_shift():
    """Start address advanced by whole 128-byte rows: rows m of D come from A[m + j] iff the swizzle
    phase follows the absolute shared-memory address (what an im2col-free 3x3 conv needs)."""
    A, B = rnd((160, 64), 5), rnd((64, 64), 6)
    for j in (1, 3, 11):
        got = run_umma(img_kmajor_sw128(A), img_kmajor_sw128(B), desc(j * 128, 16, 1024, 2), desc(0, 16, 1024, 2), 32,
                       32, 4, idesc(128, 64), 64)
        want = A[j:j + 128].astype(np.float32) @ B.astype(np.float32).T
        ok = close(got, want)
        record(f"umma_sw128_row_shift_{j}", ok, f"maxdiff {np.nanmax(np.abs(got - want)):.4g}")


def test_umma_sw128_sbo1280():
    """8-row groups 10 rows apart (SBO 1280 B) + row shift: the flattened 10-wide halo tile of the conv."""
    A, B = rnd((200, 64), 7), rnd((64, 64), 8)
    for j in (0, 1, 11, 22):
        rows = np.array([(m // 8) * 10 + m % 8 + j for m in range(128)])
        got = run_umma(img_kmajor_sw128(A), img_kmajor_sw128(B), desc(j * 128, 16, 1280, 2), desc(0, 16, 1024, 2), 32,
                       32, 4, idesc(128, 64), 64)
        want = A[rows].astype(np.float32) @ B.astype(np.float32).T
        ok = close(got, want)
        record(f"umma_sw128_sbo1280_shift_{j}", ok, f"maxdiff {np.nanmax(np.abs(got - want)):.4g}")


def test_umma_interleave():
    """SWIZZLE_NONE K-major: [K/8][rows][8]; LBO = K-chunk stride, SBO = 8-row-group stride."""
    A, B = rnd((200, 64), 9), rnd((64, 64), 10)
    a_img, b_img = img_kmajor_interleave(A), img_kmajor_interleave(B)
    for name, lbo_a, sbo_a in (("lbo_is_kchunk", 200 * 16, 128), ("sbo_is_kchunk", 128, 200 * 16)):
        # per K step (16 elems = 2 chunks) advance 2 * chunk stride
        lbo_b, sbo_b = (64 * 16, 128) if name == "lbo_is_kchunk" else (128, 64 * 16)
        got = run_umma(a_img, b_img, desc(0, lbo_a, sbo_a, 0), desc(0, lbo_b, sbo_b, 0), 2 * 200 * 16, 2 * 64 * 16, 4,
                       idesc(128, 64), 64)
        want = A[:128].astype(np.float32) @ B.astype(np.float32).T
        ok = close(got, want)
        record(f"umma_interleave_{name}", ok, f"maxdiff {np.nanmax(np.abs(got - want)):.4g}")
    # conv-style: groups 10 rows apart and a row shift, 16-byte row pitch
    for j in (0, 1, 11, 22):
        rows = np.array([(m // 8) * 10 + m % 8 + j for m in range(128)])
        got = run_umma(a_img, b_img, desc(j * 16, 200 * 16, 160, 0), desc(0, 64 * 16, 128, 0), 2 * 200 * 16,
                       2 * 64 * 16, 4, idesc(128, 64), 64)
        want = A[rows].astype(np.float32) @ B.astype(np.float32).T
        ok = close(got, want)
        record(f"umma_interleave_sbo160_shift_{j}", ok, f"maxdiff {np.nanmax(np.abs(got - want)):.4g}")


def test_umma_mn_major_b():
    """B operand stored [K][N] (N contiguous), i.e. V of attention: O = P . V."""
    A = rnd((128, 64), 11)
    for N in (64, 256):
        V = rnd((64, N), 12 + N)                       # [K=64 keys][N channels]
        b_img = img_mnmajor_sw128(V)
        for name, lbo, sbo in (("lbo_nblock", 64 * 128, 1024), ("sbo_nblock", 1024, 64 * 128)):
            got = run_umma(img_kmajor_sw128(A), b_img, desc(0, 16, 1024, 2), desc(0, lbo, sbo, 2), 32, 16 * 128, 4,
                           idesc(128, N, 0, 1), N)
            want = A.astype(np.float32) @ V.astype(np.float32)
            ok = close(got, want)
            record(f"umma_mnmajor_b_N{N}_{name}", ok, f"maxdiff {np.nanmax(np.abs(got - want)):.4g}")


def test_umma_m64_layout():
    """Where do the 64 rows of an M=64 accumulator land in TMEM?"""
    A, B = rnd((64, 64), 13), rnd((64, 64), 14)
    got = run_umma(img_kmajor_sw128(A), img_kmajor_sw128(B), desc(0, 16, 1024, 2), desc(0, 16, 1024, 2), 32, 32, 4,
                   idesc(64, 64), 64)
    want = A.astype(np.float32) @ B.astype(np.float32).T
    lanes = []
    for m in range(64):
        hit = [l for l in range(128) if np.allclose(got[l], want[m], atol=1e-3)]
        lanes.append(hit[0] if hit else -1)
    record("umma_m64_lane_of_row", all(l >= 0 for l in lanes), json.dumps(lanes))


def test_umma_m64_two_windows_interleaved():
    """Two M=64 accumulators in the SAME columns: the second with a TMEM lane offset of 16 (rows of window B in
    lanes 16-31 of each 32-lane quadrant).  n_cols bits 16..23 carry the lane offset for the probe."""
    A1, A2, B = rnd((64, 64), 40), rnd((64, 64), 41), rnd((112, 64), 42)
    L = _lib(); lib = L.load()
    want1 = A1.astype(np.float32) @ B.astype(np.float32).T
    want2 = A2.astype(np.float32) @ B.astype(np.float32).T
    got2 = run_umma(img_kmajor_sw128(A2), img_kmajor_sw128(B), desc(0, 16, 1024, 2), desc(0, 16, 1024, 2), 32, 32, 4,
                    idesc(64, 112), 112 | (16 << 16))
    lanes2 = [32 * (r // 16) + 16 + r % 16 for r in range(64)]
    ok2 = np.allclose(got2[lanes2], want2, atol=1e-3)
    got1 = run_umma(img_kmajor_sw128(A1), img_kmajor_sw128(B), desc(0, 16, 1024, 2), desc(0, 16, 1024, 2), 32, 32, 4,
                    idesc(64, 112), 112)
    lanes1 = [32 * (r // 16) + r % 16 for r in range(64)]
    ok1 = np.allclose(got1[lanes1], want1, atol=1e-3)
    record("umma_m64_lane_offset_16", ok1 and ok2, f"lane0 ok {ok1}, lane16 ok {ok2}")


def run_tma(t, dims, strides, box, swizzle, coords):
    L = _lib(); lib = L.load()
    rank = len(dims)
    nbytes = int(np.prod(box)) * 2
    out = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    A64, A32, I32 = C.c_uint64 * rank, C.c_uint32 * rank, C.c_int32 * rank
    L.check(lib.m2t_probe_tma(t.data_ptr(), 2, rank, A64(*dims), A64(*strides), A32(*box), swizzle, I32(*coords),
                              out.data_ptr(), nbytes, None), "m2t_probe_tma")
    torch.cuda.synchronize()
    return out.cpu().numpy()


def test_tma_window_sw128_oob_zero():
    """10x10x64ch fp16 window of an NHWC map at (-1,-1): rows of 128 B, 128B swizzle, zeros outside."""
    H, W = 12, 20
    t = torch.from_numpy(rnd((1, H, W, 64), 20)).cuda()
    tn = t.cpu().numpy()
    for (y0, x0) in ((-1, -1), (3, 11), (4, 12)):
        got = run_tma(t, (64, W, H, 1), (2, 128, W * 128, H * W * 128), (64, 10, 10, 1), 3, (0, x0, y0, 0))
        want = np.zeros(100 * 128, np.uint8)
        for r in range(10):
            for s in range(10):
                y, x = y0 + r, x0 + s
                px = tn[0, y, x] if (0 <= y < H and 0 <= x < W) else np.zeros(64, np.float16)
                raw = px.view(np.uint8)
                p = r * 10 + s
                for ch in range(8):
                    off = swz128(p * 128 + ch * 16)
                    want[off:off + 16] = raw[ch * 16: ch * 16 + 16]
        ok = np.array_equal(got, want)
        record(f"tma_sw128_window_{y0}_{x0}", ok, f"mismatching bytes {int((got != want).sum())}")
        assert ok


def test_tma_interleave_5d():
    """Same window as [C/8][10][10][8] (SWIZZLE_NONE core-matrix layout) through a 5-D map."""
    H, W = 12, 20
    t = torch.from_numpy(rnd((1, H, W, 64), 21)).cuda()
    tn = t.cpu().numpy()
    y0, x0 = -1, 13
    got = run_tma(t, (8, W, H, 8, 1), (2, 128, W * 128, 16, H * W * 128), (8, 10, 10, 8, 1), 0, (0, x0, y0, 0, 0))
    want = np.zeros((8, 10, 10, 8), np.float16)
    for r in range(10):
        for s in range(10):
            y, x = y0 + r, x0 + s
            if 0 <= y < H and 0 <= x < W:
                want[:, r, s, :] = tn[0, y, x].reshape(8, 8)
    ok = np.array_equal(got, want.view(np.uint8).reshape(-1))
    record("tma_interleave_5d", ok, f"mismatching bytes {int((got != want.view(np.uint8).reshape(-1)).sum())}")


def swz32(off):
    return off ^ (((off >> 7) & 1) << 4)


def img_rows32_sw32(mat):
    """mat [rows][16] fp16 (32-byte rows) -> bytes with the 32-byte swizzle (16-byte chunk ^= address bit 7)."""
    rows, K = mat.shape
    assert K == 16
    out = np.zeros(rows * 32, dtype=np.uint8)
    raw = mat.view(np.uint8).reshape(rows, 32)
    for r in range(rows):
        for ch in range(2):
            off = swz32(r * 32 + ch * 16)
            out[off:off + 16] = raw[r, ch * 16: ch * 16 + 16]
    return out


def test_umma_sw32_c16():
    """16-channel tensors (branch 1): 32-byte rows, SWIZZLE_32B, K-major A/B and MN-major B with N = 16."""
    A, B = rnd((128, 16), 30), rnd((224, 16), 31)
    got = run_umma(img_rows32_sw32(A), img_rows32_sw32(B), desc(0, 16, 256, 6), desc(0, 16, 256, 6), 0, 0, 1,
                   idesc(128, 224), 224)
    want = A.astype(np.float32) @ B.astype(np.float32).T
    record("umma_sw32_kmajor_K16", close(got, want), f"maxdiff {np.nanmax(np.abs(got - want)):.4g}")
    # row-shifted / 10-row-pitch start as for SW128
    A2 = rnd((200, 16), 32)
    rows = np.array([(m // 8) * 10 + m % 8 + 11 for m in range(128)])
    got = run_umma(img_rows32_sw32(A2), img_rows32_sw32(B), desc(11 * 32, 16, 320, 6), desc(0, 16, 256, 6), 0, 0, 1,
                   idesc(128, 224), 224)
    want = A2[rows].astype(np.float32) @ B.astype(np.float32).T
    record("umma_sw32_kmajor_shift", close(got, want), f"maxdiff {np.nanmax(np.abs(got - want)):.4g}")
    # P (K-major SW128 over 64 keys) x V (MN-major, 16 channels per key row)
    P, V = rnd((128, 64), 33), rnd((64, 16), 34)
    got = run_umma(img_kmajor_sw128(P), img_rows32_sw32(V), desc(0, 16, 1024, 2), desc(0, 16, 256, 6), 32, 16 * 32, 4,
                   idesc(128, 16, 0, 1), 16)
    want = P.astype(np.float32) @ V.astype(np.float32)
    record("umma_sw32_mnmajor_b_N16", close(got, want), f"maxdiff {np.nanmax(np.abs(got - want)):.4g}")


def test_tma_window_sw32():
    H, W = 12, 20
    t = torch.from_numpy(rnd((1, H, W, 48), 35)).cuda()          # QKV of branch 1: 3C = 48 channels
    tn = t.cpu().numpy()
    y0, x0, c0 = -1, 11, 16                                       # the K slice
    got = run_tma(t, (48, W, H, 1), (2, 96, W * 96, H * W * 96), (16, 10, 10, 1), 1, (c0, x0, y0, 0))
    want = np.zeros(100 * 32, np.uint8)
    for r in range(10):
        for s in range(10):
            y, x = y0 + r, x0 + s
            px = tn[0, y, x, c0:c0 + 16] if (0 <= y < H and 0 <= x < W) else np.zeros(16, np.float16)
            raw = np.ascontiguousarray(px).view(np.uint8)
            p = r * 10 + s
            for ch in range(2):
                off = swz32(p * 32 + ch * 16)
                want[off:off + 16] = raw[ch * 16: ch * 16 + 16]
    ok = np.array_equal(got, want)
    record("tma_sw32_window", ok, f"mismatching bytes {int((got != want).sum())}")
