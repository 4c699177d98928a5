"""In-tree build of the C-ABI CUDA library (sm_100a only).

`python -m l4p_b200.build` compiles every `csrc/*.cu` with nvcc and links `l4p_b200/libl4p_b200.so`.
The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
# L4P_BUILD_TAG=<tag> builds an experiment variant next to the production library (own object directory, own .so:
# l4p_b200/libl4p_b200_<tag>.so, selected at run time with L4P_LIB=<path>); the untagged build is what ships.
_TAG = os.environ.get("L4P_BUILD_TAG", "")
OBJ = PKG / "csrc" / ("build" if not _TAG else f"build_{_TAG}")
LIB = PKG / ("libl4p_b200.so" if not _TAG else f"libl4p_b200_{_TAG}.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
] + os.environ.get("L4P_NVCC_EXTRA", "").split()  # e.g. -DL4P_GEMM_FINE_PROF=1 for tools/gemm_prof2.py


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _digest(src: Path) -> str:
    h = hashlib.sha256()
    h.update(src.read_bytes())
    for dep in sorted(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "l4p_b200.h"]:
        h.update(dep.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(src: Path, verbose: bool) -> Path:
    obj = OBJ / (src.stem + ".o")
    stamp = OBJ / (src.stem + ".sha")
    dig = _digest(src)
    if obj.exists() and stamp.exists() and stamp.read_text() == dig:
        return obj
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = OBJ / (src.stem + ".log")
    log.write_text(res.stdout + res.stderr)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError(f"nvcc failed for {src.name}")
    if verbose:
        for line in (res.stdout + res.stderr).splitlines():
            if "registers" in line or "spill" in line and "0 bytes spill" not in line:
                print(f"[{src.name}] {line.strip()}")
    stamp.write_text(dig)
    return obj


def build(verbose: bool = False, force: bool = False) -> Path:
    OBJ.mkdir(parents=True, exist_ok=True)
    srcs = sorted(CSRC.glob("*.cu"))
    if force:
        for f in OBJ.glob("*.sha"):
            f.unlink()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    newest = max(o.stat().st_mtime for o in objs)
    if LIB.exists() and LIB.stat().st_mtime >= newest and not force:
        return LIB
    cmd = [_nvcc(), "-shared", "-o", str(LIB), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a",
           "-lcudart"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    p = build(verbose=True, force="--force" in sys.argv)
    print(p)
