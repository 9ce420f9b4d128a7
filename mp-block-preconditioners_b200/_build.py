"""Builds libmpbp.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

The library embeds a hash of every source it was compiled from (`mpbp_build_id()`); `needs_build()` and
`_cabi.load()` compare it with the hash of the tree, so a stale binary is never used silently."""
from __future__ import annotations

import glob
import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libmpbp.so")
SOURCES = ["plan.cu"]
HEADER = os.path.join(HERE, "..", "include", "mpbp.h")
_MARK = b"MPBP_BUILD_ID="
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-shared"]


def dep_files():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh"))) + [HEADER]


def source_id() -> str:
    """sha256 over the contents of every file the library is compiled from (+ the compile flags)."""
    h = hashlib.sha256(" ".join(FLAGS).encode())
    for f in dep_files():
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:32]


def binary_id(path: str = OUT):
    """The build id embedded in an existing libmpbp.so (None if the file or the marker is missing)."""
    try:
        with open(path, "rb") as fh:
            blob = fh.read()
    except OSError:
        return None
    i = blob.find(_MARK)
    if i < 0:
        return None
    return blob[i + len(_MARK): i + len(_MARK) + 32].decode("ascii", "replace")


def needs_build():
    return binary_id() != source_id()


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + FLAGS + [f'-DMPBP_BUILD_ID_STR="{source_id()}"', "-o", OUT]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ["-lnccl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose=True))
