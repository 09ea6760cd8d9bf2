"""Build the engine's shared library in-tree with nvcc (sm_100a only).

    python -m m2trans_b200.build [--force] [--verbose]

The result, `m2trans_b200/libm2trans_b200.so`, is git-ignored but travels with the
source tree.  nvcc cross-compiles without a GPU.  There is no JIT and no other
architecture: the library contains sm_100a SASS only.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libm2trans_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]
if os.environ.get("M2T_CONV_WGS"):          # tuning switch: epilogue warpgroups of the ff conv (1 or 2)
    FLAGS.append("-DM2T_CONV_WGS=" + os.environ["M2T_CONV_WGS"])
for _d in os.environ.get("M2T_DEFS", "").split():   # tuning builds: extra -D definitions (e.g. LG_NWG_OVERRIDE=1)
    FLAGS.append("-D" + _d)
if os.environ.get("M2T_TIMING") == "1":      # development builds: clock64 stamps in the attention kernel
    FLAGS.append("-DM2T_TIMING")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def source_hash() -> str:
    """sha256 over every file the library is built from (sources, headers, the public header) and the compiler flags.
    The hash is compiled into the library (`m2t_version()` ends with it), so whether a .so belongs to this tree is a
    property of its content, not of file modification times (which are arbitrary after a checkout or a copy)."""
    import hashlib
    h = hashlib.sha256()
    deps = sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + \
        sorted(glob.glob(os.path.join(HERE, "..", "include", "*.h")))
    for d in deps:
        h.update(os.path.basename(d).encode() + b"\0")
        with open(d, "rb") as f:
            h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()[:16]


HASH_TAG = b"m2t-src-hash:"


def built_hash() -> str:
    """The source hash a built library carries (searched in the file's bytes: no dlopen needed), or ''."""
    if not os.path.exists(LIB):
        return ""
    with open(LIB, "rb") as f:
        blob = f.read()
    i = blob.find(HASH_TAG)
    return blob[i + len(HASH_TAG): i + len(HASH_TAG) + 16].decode("ascii", "replace") if i >= 0 else ""


def _stale() -> bool:
    return built_hash() != source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ (one object per file, in parallel) and link the .so."""
    if not force and not _stale():
        return LIB
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    tag = source_hash()
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        cmd = [NVCC, *FLAGS, "-c", src, "-o", obj]
        if os.path.basename(src) == "api.cu":
            cmd.insert(1, f'-DM2T_SRC_HASH="{tag}"')
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs, failed = [], []
    for src, obj, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed.append((src, out))
        elif verbose or out.strip():
            print(f"[nvcc] {os.path.basename(src)}\n{out}", flush=True)
        objs.append(obj)
    if failed:
        for src, out in failed:
            sys.stderr.write(f"nvcc failed on {src}:\n{out}\n")
        raise RuntimeError("m2trans_b200: nvcc build failed")
    tmp = LIB + ".tmp"
    link = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp, *objs]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("m2trans_b200: link failed:\n" + r.stdout)
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
