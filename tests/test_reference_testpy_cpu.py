"""The reference's own evaluation script, test.py, executed VERBATIM (runpy on the file in the reference checkout) through
tools/run_reference_test.py: synthetic SR_datasets tree, reference-format checkpoint, sys.modules stand-ins for the absent
third-party packages, and this repo's drop-in surface as `models.M2Trans_network`.

Here (CPU-only container) the drop-in's forward is the CPU oracle, so the run pins the harness: the tree layout
`create_datasets` walks, the checkpoint container + DataParallel strict load (ref test.py:64-70), the loader arithmetic
and the PSNR / SSIM the script prints, which must equal this repo's own evaluation of the same files.  The reference
checkout is absent on the GPU box, where the same loop runs on the engine in tests/test_loader_gpu.py."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("M2T_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "test.py")), reason="needs the reference checkout (build container only)")
def test_reference_test_py_runs_verbatim_on_the_drop_in_surface(tmp_path):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_reference_test.py"), "--scale", "2", "--model", "oracle",
                        "--work", str(tmp_path), "--ref", REF], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    rep = json.loads(r.stdout.strip().splitlines()[-1])
    assert list(rep["reference_test_py"]) == ["CCA-US", "US-CASE", "US1K_23"]
    for name, got in rep["reference_test_py"].items():
        want = rep["own_evaluation"][name]
        assert abs(got["psnr"] - want["psnr"]) <= 0.011 and abs(got["ssim"] - want["ssim"]) <= 2e-4, (name, got, want)
    assert "GMSD:" in r.stdout and "## use cpu for training! ##" in r.stdout
