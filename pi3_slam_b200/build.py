"""In-tree nvcc build of libpi3b200.so (sm_100a only).

The shared library is built next to this file so that it travels with the repo snapshot to the GPU
box; there is no JIT cache and no fallback: if the library is missing the package refuses to run.
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB_PATH = HERE / "libpi3b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-O2",
]


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _stale() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
    return any(d.stat().st_mtime > t for d in deps)


def build_lib(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ into libpi3b200.so (parallel object builds, then link)."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    objdir = HERE / "build"
    objdir.mkdir(exist_ok=True)
    procs = []
    objs = []
    for src in sources():
        obj = objdir / (src.stem + ".o")
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-I", str(CSRC), "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"[pi3_slam_b200.build] nvcc failed for {src.name}:\n{out}\n")
        elif verbose:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc compilation of pi3_slam_b200/csrc failed")
    link = [nvcc, "-shared", "--cudart", "shared", "-o", str(LIB_PATH), *map(str, objs)]
    subprocess.run(link, check=True)
    return LIB_PATH


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
