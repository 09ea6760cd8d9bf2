"""Generate tests/golden/metrics_*.npz from the reference's own functions (run in the build container only).

    python oracle/make_golden_metrics.py

Imports /root/reference/utils.py with its unavailable third-party imports (pytorch_msssim, and whatever else is missing)
replaced by empty stubs -- rgb_to_ycbcr and calc_psnr do not touch them -- and records, for small SR / HR pairs, the Y
tensors the test loop feeds its metrics (ref test.py:103-112) and utils.calc_psnr of them.  SSIM is not recorded: the
package that computes it is absent (oracle/metrics_oracle.py)."""
import importlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def import_reference_utils():
    sys.path.insert(0, REF)
    for _ in range(20):
        try:
            return importlib.import_module("utils")
        except ModuleNotFoundError as e:       # stub the missing package (and its parent packages) and retry
            name = e.name
            mod = types.ModuleType(name)
            mod.__getattr__ = lambda attr, _n=name: (_ for _ in ()).throw(AttributeError(f"{_n}.{attr} is a stub")) \
                if attr.startswith("__") else object()
            sys.modules[name] = mod
    raise RuntimeError("could not import the reference utils")


def main():
    U = import_reference_utils()
    from m2trans_b200.synthetic import synthetic_input
    out = os.path.join(ROOT, "tests", "golden")
    cases = [("x4_48x64", 4, 2, 48, 64), ("x2_37x50", 2, 1, 37, 50), ("x3_33x33", 3, 3, 33, 33)]
    for name, scale, b, h, w in cases:
        hr = synthetic_input(b, h, w, seed=scale)
        sr = (hr + 0.03 * torch.randn(hr.shape, generator=torch.Generator().manual_seed(scale))).clamp(0, 1)
        hy = U.rgb_to_ycbcr(hr)[:, 0:1, :, :][:, :, scale:-scale, scale:-scale] * 255.      # ref test.py:103-112
        sy = U.rgb_to_ycbcr(sr)[:, 0:1, :, :][:, :, scale:-scale, scale:-scale] * 255.
        psnr = U.calc_psnr(sy, hy)
        per_image = [U.calc_psnr(sy[i:i + 1], hy[i:i + 1]) for i in range(b)]
        np.savez_compressed(os.path.join(out, f"metrics_{name}.npz"), sr=sr.numpy(), hr=hr.numpy(), scale=scale,
                            sr_y=sy.numpy(), hr_y=hy.numpy(), psnr=psnr, psnr_per_image=np.array(per_image))
        print(name, psnr, per_image)


if __name__ == "__main__":
    main()
