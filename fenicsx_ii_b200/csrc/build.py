"""Builds libfxii.so (the C-ABI library of include/fxii.h) in-tree with nvcc for sm_100a.

Usage: python -m fenicsx_ii_b200.csrc.build [--force]
The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""

from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
SOURCES = ["mesh_lbvh.cu", "pi_assemble.cu", "csr_ops.cu", "capi.cu"]
HEADERS = ["common.cuh", "../../include/fxii.h"]
LIB = HERE / "libfxii.so"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# -fmad=false on the geometry unit: fp64 results follow the oracle's operation order exactly
EXTRA = {"pi_assemble.cu": ["-fmad=false"], "mesh_lbvh.cu": ["-fmad=false"]}


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: the fxii CUDA library cannot be built")
    return cand


def _stamp() -> str:
    h = hashlib.sha256()
    for f in SOURCES + HEADERS + ["build.py"]:
        h.update((HERE / f).read_bytes())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    stamp_file = HERE / ".build_stamp"
    stamp = _stamp()
    if not force and LIB.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
        return LIB
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in SOURCES:
        obj = HERE / (src[:-3] + ".o")
        cmd = [nvcc, "-O3", "-std=c++17", "-lineinfo", *ARCH, "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
               *EXTRA.get(src, []), *os.environ.get("FXII_NVCC_FLAGS", "").split(), "-c", str(HERE / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc failed on {src}\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(f"--- {src}\n{out}\n")
    if failed:
        raise RuntimeError("nvcc compilation failed")
    link = [nvcc, "-shared", *ARCH, "-o", str(LIB), *objs, "-lcudart"]
    subprocess.run(link, check=True)
    stamp_file.write_text(stamp)
    return LIB


def build_variant(out_path, source: str, defines: list[str]) -> Path:
    """A variant of the library with ONE source recompiled with extra -D flags, linked against the other objects of the
    normal build (tests only: e.g. a tiny traversal stack to trip the overflow check).  Load it with FXII_LIB=<path>."""
    build()
    nvcc = _nvcc()
    out_path = Path(out_path)
    obj = out_path.with_suffix(".o")
    cmd = [nvcc, "-O3", "-std=c++17", "-lineinfo", *ARCH, "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", *EXTRA.get(source, []),
           *defines, "-c", str(HERE / source), "-o", str(obj)]
    subprocess.run(cmd, check=True)
    others = [str(HERE / (src[:-3] + ".o")) for src in SOURCES if src != source]
    subprocess.run([nvcc, "-shared", *ARCH, "-o", str(out_path), str(obj), *others, "-lcudart"], check=True)
    return out_path


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
