"""CPU-only checks of the boundary: the C-ABI library loads and exports every declared symbol, the
nn.Module mirror has the reference's state_dict surface and load semantics, host-side planning works
without a GPU, and nothing silently falls back to the CPU."""
import ctypes as C
import os
import re
import types

import pytest
import torch

from m2trans_b200 import _lib
from m2trans_b200.M2Trans_network import M2Trans, M2TError, create_model
from m2trans_b200.synthetic import reference_checkpoint, state_dict_spec, synthetic_state_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _args(scale, n_blocks=8):
    return types.SimpleNamespace(scale=scale, rgb_range=1.0, colors=3, n_feats=64, num_heads=4, n_blocks=n_blocks)


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "m2trans_b200.h")).read()
    declared = set(re.findall(r"\b(m2t_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.m2t_version()


def test_host_side_planning_without_gpu():
    lib = _lib.load()
    assert lib.m2t_num_params(4, 8) == 123 and lib.m2t_num_params(2, 8) == 121 and lib.m2t_num_params(3, 8) == 121
    assert lib.m2t_packed_weight_bytes(4, 8) > 3_500_000 * 2
    assert lib.m2t_packed_weight_bytes(5, 8) == 0
    assert lib.m2t_packed_offset(4, 8, b"body.7.attn4.wqkv") < lib.m2t_packed_weight_bytes(4, 8)
    assert lib.m2t_packed_offset(4, 8, b"nonsense") == C.c_size_t(-1).value
    cfg = _lib.m2t_cfg(3, 64, 8, 3, 32, 200, 266, 0, 1.0)          # BASELINE configs[2]
    plan = C.c_void_p()
    assert lib.m2t_plan_create(C.byref(cfg), C.byref(plan)) == 0
    hp, wp = C.c_int(), C.c_int()
    assert lib.m2t_plan_padded(plan, C.byref(hp), C.byref(wp)) == 0
    assert (hp.value, wp.value) == (224, 288)
    assert lib.m2t_workspace_bytes(plan) > 32 * 224 * 288 * 64 * 4 * 2
    assert lib.m2t_plan_num_launches(plan) > 0
    lib.m2t_plan_destroy(plan)
    # the reference raises for reflect pads >= the dimension (F.pad); so does the plan
    bad = _lib.m2t_cfg(4, 64, 8, 3, 1, 8, 40, 0, 1.0)
    assert lib.m2t_plan_create(C.byref(bad), C.byref(plan)) == -4
    assert b"reflect" in lib.m2t_last_error()
    bad = _lib.m2t_cfg(4, 32, 8, 3, 1, 64, 64, 0, 1.0)
    assert lib.m2t_plan_create(C.byref(bad), C.byref(plan)) == -4


@pytest.mark.parametrize("scale", [2, 3, 4])
def test_module_state_dict_surface(golden_dir, scale):
    m = create_model(_args(scale))
    want = []
    for line in open(os.path.join(golden_dir, "state_dict_manifest.txt")):
        s, key, shape, dtype = line.split()
        if s == f"x{scale}":
            want.append((key, tuple(int(t) for t in shape.split("x")), dtype))
    got = [(k, tuple(v.shape), str(v.dtype).replace("torch.", "")) for k, v in m.state_dict().items()]
    assert got == want
    assert [k for k, _ in state_dict_spec(scale)] == [k for k, _, _ in want]
    for attr in ("scale", "window_sizes", "rgb_range", "n_blocks", "head", "body", "tail", "sub_mean", "add_mean"):
        assert hasattr(m, attr)
    assert m.window_sizes == [8, 16, 32]


def test_reference_checkpoint_loads_through_dataparallel():
    ckpt = reference_checkpoint(4, 0)
    assert set(ckpt) == {"epoch", "model_state_dict", "optimizer_state_dict", "scheduler_state_dict", "stat_dict"}
    model = torch.nn.DataParallel(M2Trans(_args(4)))
    model.load_state_dict(ckpt["model_state_dict"], strict=True)    # ref test.py:70
    sd = synthetic_state_dict(4, 0)
    for k, v in model.module.state_dict().items():
        assert torch.equal(v, sd[k]), k


def test_load_state_dict_override_semantics(capsys):
    m = M2Trans(_args(3))
    sd2 = synthetic_state_dict(2, 0)
    m.load_state_dict(sd2)                                           # default strict=False, as the reference
    assert "Replace pre-trained upsampler" in capsys.readouterr().out   # tail.0 shape differs (x2 -> x3)
    assert torch.equal(m.head.weight, sd2["head.weight"])
    with pytest.raises(KeyError):
        m.load_state_dict({k: v for k, v in sd2.items() if k != "head.bias"}, strict=True)
    extra = dict(synthetic_state_dict(3, 0))
    extra["body.0.bogus"] = torch.zeros(1)
    with pytest.raises(KeyError):
        m.load_state_dict(extra, strict=True)
    extra = dict(synthetic_state_dict(3, 0))
    extra["tail.9.weight"] = torch.zeros(1)                          # unexpected tail keys are tolerated
    m.load_state_dict(extra, strict=True)
    bad = dict(synthetic_state_dict(3, 0))
    bad["head.bias"] = torch.zeros(7)
    with pytest.raises(RuntimeError):
        m.load_state_dict(bad)


def test_no_cpu_fallback():
    m = M2Trans(_args(2, n_blocks=1))
    with pytest.raises(M2TError):
        m(torch.rand(1, 3, 32, 32))
    with pytest.raises(M2TError):
        m.body[0].attn1(torch.rand(1, 16, 8, 8))


def test_check_image_size_matches_reference_rule():
    m = M2Trans(_args(4, n_blocks=1))
    x = torch.rand(1, 3, 40, 50)
    xp = m.check_image_size(x)
    assert tuple(xp.shape) == (1, 3, 64, 64)
    assert torch.equal(xp[:, :, :40, :50], x)
    assert torch.equal(xp[:, :, 40, :50], x[:, :, 38, :])            # reflect without repeating the edge


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "m2trans_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), os.path.join(dirpath, f)


def _fake_replica(module):
    """What torch.nn.parallel.replicate() does to every module of a replica, without needing CUDA: shallow __dict__
    copy, EMPTY _parameters, the (here: same) tensors as plain attributes and in _former_parameters."""
    import collections
    mods = list(module.modules())
    copies = [m._replicate_for_data_parallel() for m in mods]
    index = {m: i for i, m in enumerate(mods)}
    for m, r in zip(mods, copies):
        r._former_parameters = collections.OrderedDict()
        for key, child in m._modules.items():
            if child is not None:
                setattr(r, key, copies[index[child]])
        for key, p in m._parameters.items():
            if p is None:
                r._parameters[key] = None
            else:
                t = p.detach()
                setattr(r, key, t)
                r._former_parameters[key] = t
    return copies[0]


def test_parameter_discovery_on_dataparallel_replicas():
    """ref test.py:68 wraps the model in nn.DataParallel over all visible GPUs; a replica's state_dict() is empty, so
    the engine must find the broadcast copies through _former_parameters, in state_dict order."""
    import types
    from m2trans_b200.M2Trans_network import M2Trans
    from m2trans_b200._params import is_replica, state_tensors
    from m2trans_b200.rlutrans import TransBlock
    m = M2Trans(types.SimpleNamespace(scale=4, rgb_range=1.0, colors=3, n_feats=64, n_blocks=8))
    want = [p for _, p in m.state_dict(keep_vars=True).items()]
    assert all(a is b for a, b in zip(m._param_list(), want)) and len(want) == 123
    rep = _fake_replica(m)
    assert len(rep.state_dict()) == 0 and is_replica(rep) and not is_replica(m)
    got = rep._param_list()
    assert len(got) == 123 and all(a.data_ptr() == b.data_ptr() and a.shape == b.shape for a, b in zip(got, want))
    assert m._param_list()[0] is want[0]                       # the original's cached slots are not the replica's
    tb = TransBlock()
    tw = [p for _, p in tb.state_dict(keep_vars=True).items()]
    tr = state_tensors(_fake_replica(tb))
    assert len(tr) == 12 and all(a.data_ptr() == b.data_ptr() for a, b in zip(tr, tw))


def test_library_carries_the_hash_of_the_sources_it_was_built_from():
    """build() rebuilds when the hash compiled into the .so differs from the hash of the tree (not on mtimes), so a
    stale binary cannot travel with changed sources unnoticed."""
    from m2trans_b200 import _lib, build
    tag = build.source_hash()
    assert len(tag) == 16 and build.built_hash() == tag and not build._stale()
    assert _lib.load().m2t_version().decode().endswith("m2t-src-hash:" + tag)


def test_released_checkpoint_identification(tmp_path):
    """git blob SHA-1 (known answers: the empty blob and 'hello\\n'), the container reader and the released-file lookup."""
    from m2trans_b200 import checkpoints as CK
    from m2trans_b200.synthetic import save_reference_checkpoint, synthetic_state_dict
    empty, hello = tmp_path / "empty", tmp_path / "hello"
    empty.write_bytes(b"")
    hello.write_bytes(b"hello\n")
    assert CK.git_blob_sha1(str(empty)) == "e69de29bb2d1d6434b8b29ae775ad8c2e48c5391"
    assert CK.git_blob_sha1(str(hello)) == "ce013625030ba8dba906f756967f9e9ca394464a"
    assert sorted(CK.RELEASED_SHA1) == [2, 3, 4] and all(len(v) == 40 for v in CK.RELEASED_SHA1.values())
    d = tmp_path / "checkpoints"
    d.mkdir()
    save_reference_checkpoint(str(d / "model_x2.pt"), 2, 0)          # right name, synthetic content
    assert CK.identify(str(d / "model_x2.pt")) is None and CK.find_released(str(d)) == {}
    sd = CK.load_model_state_dict(str(d / "model_x2.pt"))
    want = synthetic_state_dict(2, 0)
    assert list(sd) == list(want) and all(torch.equal(sd[k], want[k]) for k in want)
