"""The reference's released checkpoints `checkpoints/model_x{2,3,4}.pt` (ref test.py:64-70, train.py:343-349).

The files are large blobs that are not part of the reference checkout this repo was built against; only their git blob
SHA-1s are known.  If a user drops them in, `identify` says which scale a file is (by content, not by
name), `load_model_state_dict` reads the container the reference's training loop writes, and `tests/released_checkpoint_check.py`
re-runs the precision table of DESIGN.md section 3 on them."""
from __future__ import annotations

import hashlib
import os
from collections import OrderedDict

import torch

__all__ = ["RELEASED_SHA1", "git_blob_sha1", "identify", "find_released", "load_model_state_dict"]

RELEASED_SHA1 = {
    2: "c5f03dc4edf444b5aa4d46ee25df1f17f601d05a",
    3: "cab6082398efb0cce2ff2e8f034c09a5cb30d852",
    4: "5abde48d5d0e40b4b183fa6aba93dffd4839d2ea",
}


def git_blob_sha1(path: str) -> str:
    """What `git hash-object <path>` prints: SHA-1 over b"blob <size>\\0" followed by the file's bytes."""
    h = hashlib.sha1()
    h.update(b"blob %d\0" % os.path.getsize(path))
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


def identify(path: str):
    """The scale (2, 3 or 4) whose released checkpoint this file is, or None."""
    sha = git_blob_sha1(path)
    for scale, want in RELEASED_SHA1.items():
        if sha == want:
            return scale
    return None


def find_released(directory: str | None = None) -> dict:
    """{scale: path} of the files under `directory` (default: $M2T_CHECKPOINTS, else ./checkpoints) named like the
    reference's and carrying the released content."""
    directory = directory or os.environ.get("M2T_CHECKPOINTS") or "checkpoints"
    found = {}
    for scale in RELEASED_SHA1:
        path = os.path.join(directory, f"model_x{scale}.pt")
        if os.path.isfile(path) and identify(path) == scale:
            found[scale] = path
    return found


def load_model_state_dict(path: str, map_location="cpu") -> "OrderedDict[str, torch.Tensor]":
    """`model_state_dict` of a checkpoint written by ref train.py:343-349, with the `module.` prefix nn.DataParallel adds
    removed, ready for `M2Trans.load_state_dict`.  A bare state_dict is accepted as well."""
    ckpt = torch.load(path, map_location=map_location, weights_only=True)
    sd = ckpt.get("model_state_dict", ckpt) if isinstance(ckpt, dict) else ckpt
    if not isinstance(sd, dict) or not all(isinstance(v, torch.Tensor) for v in sd.values()):
        raise ValueError(f"{path}: no model_state_dict of tensors found")
    return OrderedDict((k[7:] if k.startswith("module.") else k, v) for k, v in sd.items())
