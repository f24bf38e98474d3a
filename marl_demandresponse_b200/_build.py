"""In-tree build of ``libdrsim.so`` (nvcc, sm_100a only -- no other arch, no JIT cache)."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.environ.get("DRSIM_LIB") or os.path.join(PKG, "libdrsim.so")
SOURCES = ["drsim_api.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "-Xptxas=-v",
    "-diag-suppress", "177",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libdrsim.so cannot be built (there is no CPU fallback)")


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    # every source / header the translation unit can include: csrc/* and the public header
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    deps.append(os.path.join(PKG, "..", "include", "drsim.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA extension for sm_100a; returns the path of the shared library."""
    if not force and not stale():
        return LIB
    extra = os.environ.get("DRSIM_EXTRA_NVCC_FLAGS", "").split()
    cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-o", LIB, *[os.path.join(CSRC, s) for s in SOURCES]]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(" ".join(cmd))
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libdrsim.so")
    with open(LIB + ".build.log" if os.environ.get("DRSIM_LIB") else os.path.join(PKG, "libdrsim.build.log"), "w") as f:
        # compile times differ from build to build: keep them out of the tracked log
        log = "\n".join(l for l in (res.stdout + res.stderr).splitlines() if "Compile time" not in l)
        f.write(" ".join(cmd) + "\n" + log + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
