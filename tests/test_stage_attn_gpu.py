"""tcgen05 halo attention against the CUDA-core variant and a torch restatement of ref
M2Trans_network.py:310-332, on the engine's native tensors (QKV fp16 NHWC, rel tables).  B200 only."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _torch_attention(qkv, relh, relw, C):
    """qkv [B,h,w,3C] fp16 (q pre-scaled), relh/relw [10,C/2] fp32 -> O [B,h,w,C] fp32."""
    B, h, w, _ = qkv.shape
    t = qkv.float().permute(0, 3, 1, 2)
    q, k, v = t[:, :C], t[:, C:2 * C], t[:, 2 * C:]
    nh, nw = h // 8, w // 8
    q = q.reshape(B, C, nh, 8, nw, 8).permute(0, 2, 4, 3, 5, 1).reshape(B * nh * nw, 64, C)

    def neigh(u):
        u = F.unfold(u, kernel_size=10, stride=8, padding=1).reshape(B, C, 100, nh * nw).permute(0, 3, 2, 1)
        return u.reshape(B * nh * nw, 10, 10, C)
    k = neigh(k)
    v = neigh(v).reshape(B * nh * nw, 100, C)
    k = torch.cat((k[..., : C // 2] + relh.view(1, 10, 1, C // 2), k[..., C // 2:] + relw.view(1, 1, 10, C // 2)), -1)
    k = k.reshape(B * nh * nw, 100, C)
    o = torch.softmax(q @ k.transpose(1, 2), -1) @ v
    return o.reshape(B, nh, nw, 8, 8, C).permute(0, 1, 3, 2, 4, 5).reshape(B, h, w, C)


@pytest.mark.parametrize("C", [16, 64, 256])
@pytest.mark.parametrize("B,h,w", [(1, 8, 8), (1, 16, 24), (3, 24, 8), (2, 32, 32)])
def test_stage_attn_tensor_core_vs_cuda_core(C, B, h, w):
    from m2trans_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(C * 100 + B * 10 + h + w)
    qkv = torch.randn(B, h, w, 3 * C, generator=g)
    qkv[..., :C] *= 2.0 * C ** -0.5                      # q pre-scaled (and a bit sharper than init)
    qkv = qkv.half().cuda()
    relh = torch.randn(10, C // 2, generator=g).cuda()
    relw = torch.randn(10, C // 2, generator=g).cuda()
    relf = torch.cat((relh, relw), 0).contiguous()
    relx = torch.zeros(32, C, dtype=torch.float16, device="cuda")
    relx[:10, : C // 2] = relh.half()
    relx[10:20, C // 2:] = relw.half()
    want = _torch_attention(qkv, relh.half().float(), relw.half().float(), C)
    outs = []
    for variant in (0, _lib.VAR_SIMT_ATTN):
        o = torch.full((B, h, w, C), float("nan"), dtype=torch.float16, device="cuda")
        _lib.check(lib.m2t_stage_attn(variant, C, qkv.data_ptr(), relf.data_ptr(), relx.data_ptr(), o.data_ptr(), B, h, w,
                                      None), "m2t_stage_attn")
        torch.cuda.synchronize()
        assert torch.isfinite(o).all(), f"variant {variant}"
        outs.append(o.float())
    tol = 4e-3 * max(1.0, want.abs().max().item())
    err_tc = (outs[0] - want).abs().max().item()
    err_cc = (outs[1] - want).abs().max().item()
    print(f"C={C} {B}x{h}x{w}: tensor-core err {err_tc:.2e}, cuda-core err {err_cc:.2e}, tol {tol:.2e}")
    assert err_cc <= tol
    assert err_tc <= tol
