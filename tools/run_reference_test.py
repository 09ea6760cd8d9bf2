"""Run the REFERENCE's own evaluation script, `test.py`, verbatim (runpy on the file where it lies) against this repo's
drop-in `models/M2Trans_network.py` (ref test.py:23-122, datas/utils.py:26-43, datas/benchmark.py).

    python tools/run_reference_test.py --scale 2 [--ref /root/reference] [--model engine|oracle] [--work DIR]

What the script provides around the unmodified reference file:
  * a synthetic `../SR_datasets` tree in the layout `create_datasets` walks: benchmark/{UI5,US15,US1K_23}/{HR,LR_bicubic/X<s>}
    (a few small frames each) and US1K/US1K_train_{HR,LR_bicubic/X<s>} (test.py builds the TRAINING set too: 1000 tiny frames);
  * `./checkpoints/model_x<s>.pt` in the reference's container format (synthetic.reference_checkpoint);
  * `sys.modules` stand-ins for the third-party packages that are absent offline: imageio (PIL), skimage.color, pytorch_msssim
    (the published SSIM algorithm, oracle/metrics_oracle.py), piq (GMSD and FSIM restated from the published algorithms:
    both are printed by test.py but are not on the engine's path);
  * `models.M2Trans_network` = m2trans_b200.M2Trans_network (--model engine: needs a B200) or, on a CPU-only box, a module
    with the same surface whose forward is the CPU oracle (--model oracle): that run proves the harness -- dataset tree, stubs,
    checkpoint format, DataParallel + strict load -- with the reference script driving everything.
The reference tree is only READ (imports + runpy); nothing is copied.  It does not exist on the GPU box, so the engine run
needs a machine that has both a B200 and the reference checkout; the same loop on the engine alone is
tests/test_loader_gpu.py::test_eval_loop_uint8_to_metrics_matches_oracle.
Prints the script's output and a JSON line with the parsed PSNR / SSIM per eval set next to this repo's own evaluation of the
same tree (oracle forward + metrics oracle)."""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import re
import runpy
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SETS = {"CCA-US": ("UI5", ".jpg"), "US-CASE": ("US15", ".jpg"), "US1K_23": ("US1K_23", ".png")}


def smooth_image(rng, h, w):
    """A smooth, speckled RGB frame (uint8): low-pass noise times multiplicative speckle, like an ultrasound crop."""
    lo = rng.random((h // 8 + 2, w // 8 + 2, 1))
    base = np.kron(lo, np.ones((8, 8, 1)))[:h, :w]
    img = base * (0.35 + 0.65 * rng.random((h, w, 3)))
    return np.clip(img * 255.0, 0, 255).astype(np.uint8)


def downsample(hr, s):
    """Box-filter LR (the real sets use bicubic; any LR of the right size exercises the same code path)."""
    h, w, c = hr.shape
    return hr[: h // s * s, : w // s * s].reshape(h // s, s, w // s, s, c).astype(np.float32).mean((1, 3)).round().astype(np.uint8)


def build_tree(work, scale, n_eval=2, seed=33):
    from PIL import Image
    rng = np.random.default_rng(seed)
    data = os.path.join(work, "SR_datasets")

    def save(path, arr):
        os.makedirs(os.path.dirname(path), exist_ok=True)
        # PNG everywhere (lossless, so the files hold exactly these pixels); the .jpg NAMES are what Benchmark expects
        Image.fromarray(arr).save(path, format="PNG")
    sizes = [(40 * scale, 56 * scale), (33 * scale, 47 * scale)]
    for _set, (folder, ext) in SETS.items():
        for i in range(n_eval):
            h, w = sizes[i % len(sizes)]
            hr = smooth_image(rng, h, w)
            tag = f"img{i:03d}{ext}"
            save(os.path.join(data, "benchmark", folder, "HR", tag), hr)
            save(os.path.join(data, "benchmark", folder, "LR_bicubic", f"X{scale}", tag.replace(ext, f"x{scale}{ext}")), downsample(hr, scale))
    tiny = smooth_image(rng, 16 * scale, 16 * scale)
    tiny_lr = downsample(tiny, scale)
    for i in range(1, 1001):                      # US1K(train=True) reads 0001..1000 at construction (ref datas/us1k.py:71-80)
        idx = str(i).zfill(4)
        save(os.path.join(data, "US1K", "US1K_train_HR", idx + ".png"), tiny)
        save(os.path.join(data, "US1K", "US1K_train_LR_bicubic", f"X{scale}", f"{idx}x{scale}.png"), tiny_lr)
    return data


def install_stubs(model_kind, scale):
    from PIL import Image
    from oracle import metrics_oracle as MO

    imageio = types.ModuleType("imageio")
    imageio.imread = lambda path, pilmode="RGB": np.asarray(Image.open(path).convert(pilmode))
    sys.modules["imageio"] = imageio

    skimage, color = types.ModuleType("skimage"), types.ModuleType("skimage.color")

    def rgb2ycbcr(img):                           # ITU-R BT.601 as skimage defines it, uint8 RGB in -> float YCbCr out
        m = np.array([[65.481, 128.553, 24.966], [-37.797, -74.203, 112.0], [112.0, -93.786, -18.214]]) / 255.0
        return img.astype(np.float64) @ m.T + np.array([16.0, 128.0, 128.0])
    color.rgb2ycbcr = rgb2ycbcr
    skimage.color = color
    sys.modules["skimage"], sys.modules["skimage.color"] = skimage, color

    msssim = types.ModuleType("pytorch_msssim")
    msssim.ssim = lambda x, y, size_average=True, **kw: MO.ssim(x, y)       # pytorch_msssim defaults, data_range 255
    sys.modules["pytorch_msssim"] = msssim

    piq = types.ModuleType("piq")

    def gmsd(x, y, data_range=1.0, reduction="none"):
        return MO.gmsd(x, y, data_range).to(x.dtype)       # the published algorithm, oracle/metrics_oracle.py
    piq.gmsd = gmsd
    piq.fsim = lambda x, y, data_range=1.0, reduction="none": MO.fsim(x, y, data_range).to(x.dtype)     # published algorithm
    sys.modules["piq"] = piq

    # the drop-in: models/M2Trans_network.py of the reference tree is shadowed by this repo's module
    models = types.ModuleType("models")
    models.__path__ = []
    if model_kind == "engine":
        import m2trans_b200.M2Trans_network as net
    else:
        net = oracle_module()
    sys.modules["models"], sys.modules["models.M2Trans_network"] = models, net
    models.M2Trans_network = net


def oracle_module():
    """A module with the surface of models/M2Trans_network.py whose forward is the CPU oracle (CPU-only harness runs)."""
    import torch.nn as nn
    from oracle import m2trans_oracle as O
    from m2trans_b200.synthetic import state_dict_spec
    mod = types.ModuleType("models.M2Trans_network")

    class M2Trans(nn.Module):
        def __init__(self, args):
            super().__init__()
            self.scale = args.scale
            self._keys = []
            for key, shape in state_dict_spec(args.scale, args.n_blocks, args.n_feats, args.colors):
                name = "p_" + key.replace(".", "__")
                self.register_parameter(name, nn.Parameter(torch.zeros(shape), requires_grad=False))
                self._keys.append((key, name))

        def state_dict(self, *a, prefix="", **k):
            return {prefix + key: getattr(self, name).data for key, name in self._keys}

        def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
            for key, name in self._keys:
                if prefix + key in state_dict:
                    getattr(self, name).data.copy_(state_dict[prefix + key])
                elif strict:
                    missing_keys.append(prefix + key)
            known = {prefix + key for key, _ in self._keys}
            unexpected_keys += [k for k in state_dict if k.startswith(prefix) and k not in known]

        def forward(self, x):
            return O.forward({key: getattr(self, name).data for key, name in self._keys}, x, scale=self.scale)

    mod.M2Trans = M2Trans
    mod.create_model = lambda args: M2Trans(args)
    mod.nn, mod.torch = nn, torch
    mod.__all__ = ["M2Trans", "create_model"]
    return mod


def own_evaluation(data, scale):
    """This repo's evaluation of the same tree: oracle forward + metrics oracle, the arithmetic of ref test.py:101-122."""
    from PIL import Image
    from oracle import m2trans_oracle as O
    from oracle import metrics_oracle as MO
    from m2trans_b200.synthetic import synthetic_state_dict
    sd = synthetic_state_dict(scale, 0)
    out = {}
    for name, (folder, ext) in SETS.items():
        hr_dir = os.path.join(data, "benchmark", folder, "HR")
        ps = ss = 0.0
        tags = sorted(os.listdir(hr_dir))
        for tag in tags:
            hr = np.asarray(Image.open(os.path.join(hr_dir, tag)).convert("RGB"))
            lr = np.asarray(Image.open(os.path.join(data, "benchmark", folder, "LR_bicubic", f"X{scale}", tag.replace(ext, f"x{scale}{ext}"))).convert("RGB"))
            hr = hr[: lr.shape[0] * scale, : lr.shape[1] * scale]
            to_t = lambda a: torch.from_numpy(np.ascontiguousarray(a.transpose(2, 0, 1))).float()[None] / 255.0
            with torch.no_grad():
                sr = O.forward(sd, to_t(lr), scale=scale)
            p, s = MO.test_loop_metrics(sr, to_t(hr), scale)
            ps, ss = ps + p, ss + s
        out[name] = {"psnr": round(ps / len(tags) + 5e-3, 2), "ssim": round(ss / len(tags) + 5e-5, 4)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=os.environ.get("M2T_REFERENCE", "/root/reference"))
    ap.add_argument("--scale", type=int, default=2, choices=[2, 3, 4])
    ap.add_argument("--model", default="engine" if torch.cuda.is_available() else "oracle", choices=["engine", "oracle"])
    ap.add_argument("--work", default=None)
    args = ap.parse_args()
    if not os.path.isfile(os.path.join(args.ref, "test.py")):
        raise SystemExit(f"{args.ref}/test.py not found: this runner executes the reference's own file and copies nothing")
    from m2trans_b200.synthetic import save_reference_checkpoint
    work = args.work or tempfile.mkdtemp(prefix="m2t_testpy_")
    cwd = os.path.join(work, "run")
    os.makedirs(os.path.join(cwd, "checkpoints"), exist_ok=True)
    data = build_tree(work, args.scale)
    save_reference_checkpoint(os.path.join(cwd, "checkpoints", f"model_x{args.scale}.pt"), args.scale, 0)
    install_stubs(args.model, args.scale)
    if not torch.cuda.is_available():
        torch.cuda.set_device = lambda *a, **k: None       # test.py:46 calls it unconditionally
    sys.path.insert(0, args.ref)
    old_cwd, old_argv = os.getcwd(), sys.argv
    os.chdir(cwd)                                         # data_path '../SR_datasets/', model_path './checkpoints/...'
    sys.argv = ["test.py", "--config", os.path.join(args.ref, "configs", f"M2Trans_x{args.scale}_test.yml")]
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            runpy.run_path(os.path.join(args.ref, "test.py"), run_name="__main__")
    finally:
        os.chdir(old_cwd)
        sys.argv = old_argv
        torch.set_grad_enabled(True)
    text = buf.getvalue()
    print(text)
    got = [{"psnr": float(p), "ssim": float(s)} for p, s in re.findall(r"PSNR:(-?[0-9.]+),SSIM:(-?[0-9.]+)", text)]
    names = re.findall(r"select (.*) for evaluation", text)
    order = names[0].split() if names else list(SETS)
    report = {"model": args.model, "scale": args.scale, "reference_test_py": dict(zip(order, got)), "own_evaluation": own_evaluation(data, args.scale)}
    print(json.dumps(report))
    return report


if __name__ == "__main__":
    main()
