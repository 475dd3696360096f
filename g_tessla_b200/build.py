"""Build libgtessla_b200.so (and the g_tessla CLI) in-tree with nvcc for sm_100a."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgtessla_b200.so")
CLI = os.path.join(HERE, "g_tessla")


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libgtessla_b200 is CUDA-only and cannot be built without it")


def _stale(target: str, srcs: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in srcs)


def build(force: bool = False, verbose: bool = False, out: str | None = None) -> str:
    """Compile every CUDA source for sm_100a. -fmad=false: the FP64 predicate filter and the area
    arithmetic must round every operation individually (reference predicates.c:123-133)."""
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h", ".c"))]
    srcs += [os.path.join(HERE, "..", "include", f) for f in sorted(os.listdir(os.path.join(HERE, "..", "include")))]
    lib = out or LIB
    if force or _stale(lib, srcs):
        cu = [s for s in srcs if s.endswith(".cu")]
        cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
               "-fmad=false", "-Xcompiler", "-fPIC,-O2,-fopenmp", "-shared", "-o", lib, *cu,
               "-I" + os.path.join(HERE, "..", "include"), "-lgomp"]
        cmd += os.environ.get("GTA_B200_NVCC_FLAGS", "").split()      # experiments: extra -D switches (scratch/README.md)
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.check_call(cmd, cwd=CSRC)
    if out:
        return lib
    cli_src = os.path.join(HERE, "cli", "g_tessla.c")
    if os.path.exists(cli_src) and (force or _stale(CLI, [cli_src, LIB])):
        subprocess.check_call(["gcc", "-std=c99", "-O2", "-Wall", "-o", CLI, cli_src,
                               "-I" + os.path.join(HERE, "..", "include"), "-L" + HERE, "-lgtessla_b200",
                               "-Wl,-rpath,$ORIGIN", "-Wl,-rpath-link," + os.path.dirname(nvcc_path()) + "/../lib64", "-lm"])
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
