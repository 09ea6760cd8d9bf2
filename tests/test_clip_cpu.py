"""CPU checks for the MedCLIP image pass: the oracle's Swin-T restatement against the Hugging Face
implementation it cites, and the host-side parameter container."""
import pytest
import torch


def test_oracle_matches_hf_swin():
    """oracle.swin_pooled == transformers.SwinModel(SwinConfig()).pooler_output on randomised weights (the stand-in
    SURVEY.md §8c names: the real medclip package and weights are not available offline)."""
    tr = pytest.importorskip("transformers")
    from oracle import medclip_image_oracle as O
    torch.manual_seed(0)
    m = tr.SwinModel(tr.SwinConfig()).eval()
    with torch.no_grad():
        for n, p in m.named_parameters():
            if p.dim() == 1 or "table" in n:
                p.add_(0.1 * torch.randn_like(p))
    P = dict(m.state_dict())
    assert all(n in P for n in O.param_names() if n != "projection_head.weight")
    x = torch.rand(2, 3, 224, 224)
    with torch.no_grad():
        ref = m(pixel_values=x).pooler_output
        mine = O.swin_pooled(x, P)
    assert mine.shape == (2, 768)
    assert float((ref - mine).abs().max()) <= 1e-5


def test_param_spec_matches_oracle_and_module():
    from oracle import medclip_image_oracle as O
    from m2trans_b200.medclip_image import MedCLIPVisionModelViT, swin_param_spec, synthetic_state_dict
    spec = swin_param_spec()
    assert len(spec) == 220
    assert [n for n, _ in spec] == [n if n == "projection_head.weight" else "model." + n for n in O.param_names()]
    tower = MedCLIPVisionModelViT()
    named = dict(tower.named_parameters())
    assert list(named) == [n for n, _ in spec]                       # registration order is the engine's order
    assert all(tuple(named[n].shape) == tuple(s) for n, s in spec)
    sd = synthetic_state_dict(seed=1)
    missing, unexpected = tower.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.endswith("relative_position_index") for k in missing)
    assert sum(p.numel() for p in tower.parameters()) == 27_519_354 + 512 * 768     # Swin-T + projection head


def test_module_state_dict_keys_are_upstream_names():
    """An upstream vision-tower checkpoint (HF SwinModel keys under `model.`, plus projection_head.weight) loads
    strictly."""
    tr = pytest.importorskip("transformers")
    from m2trans_b200.medclip_image import MedCLIPVisionModelViT
    hf = tr.SwinModel(tr.SwinConfig())
    sd = {"model." + k: v for k, v in hf.state_dict().items()}
    sd["projection_head.weight"] = torch.zeros(512, 768)
    MedCLIPVisionModelViT().load_state_dict(sd, strict=True)


def test_oracle_shift_mask_matches_region_rule():
    """The engine recomputes the shift mask from region ids; same rule here against the oracle's slice construction."""
    from oracle import medclip_image_oracle as O
    for h in (56, 28, 14):
        m = O.shift_mask(h, h, 7, 3)
        ids = torch.empty(h, h)
        for y in range(h):
            for x in range(h):
                iy = 0 if y < h - 7 else (1 if y < h - 3 else 2)
                ix = 0 if x < h - 7 else (1 if x < h - 3 else 2)
                ids[y, x] = iy * 3 + ix
        win = ids.view(h // 7, 7, h // 7, 7).permute(0, 2, 1, 3).reshape(-1, 49)
        mine = torch.where(win[:, :, None] != win[:, None, :], -100.0, 0.0)
        assert torch.equal(mine, m)


def test_cpu_input_is_refused():
    from m2trans_b200._lib import M2TError
    from m2trans_b200.medclip_image import MedCLIPVisionModelViT
    with pytest.raises(M2TError):
        MedCLIPVisionModelViT().encode_image(torch.zeros(1, 3, 224, 224))
