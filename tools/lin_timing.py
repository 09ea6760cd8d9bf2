"""Timing of the tower's Linear kernel on chosen shapes / epilogues (m2t_clip_stage_linear), CUDA events, 20 launches each."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from m2trans_b200 import _lib  # noqa: E402

lib = _lib.load()
NAMES = {0: "bf16", 1: "gelu->bf16", 2: "fp32 +=", 3: "fp32 ="}
shapes = [(100352, 384, 96), (100352, 384, 384), (100352, 128, 96), (100352, 96, 96), (100352, 96, 384), (100352, 288, 96),
          (25088, 768, 192), (25088, 192, 768), (6272, 1536, 384), (6272, 384, 1536)]
for M, N, K in shapes:
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda")
    line = f"M {M:6d} N {N:4d} K {K:4d}:"
    for epi in (0, 1, 2, 3):
        out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16 if epi < 2 else torch.float32)
        s = torch.cuda.current_stream().cuda_stream
        for _ in range(3):
            _lib.check(lib.m2t_clip_stage_linear(epi, a.data_ptr(), w.data_ptr(), bias.data_ptr(), out.data_ptr(), M, N, K, s), "lin")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            lib.m2t_clip_stage_linear(epi, a.data_ptr(), w.data_ptr(), bias.data_ptr(), out.data_ptr(), M, N, K, s)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        byts = M * K * 2 + M * N * (2 if epi < 2 else (8 if epi == 2 else 4))
        line += f"  {NAMES[epi]} {us:6.1f} us ({byts / us / 1e6:4.2f} TB/s)"
    print(line)
