"""Builds libraydar_cuda.so (sm_100a) in-tree with nvcc.  nvcc cross-compiles without a GPU."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libraydar_cuda.so")
SOURCES = ["rdr_kernels.cu", "rdr_api.cpp", "rdr_multi.cpp", "rdr_scene_io.cpp"]


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "raydar_cuda.h"), __file__,
            os.path.join(HERE, "host", "raydar_cuda_main.cpp")]
    if not os.path.exists(os.path.join(HERE, "host", "raydar-cuda")):
        return True
    return any(os.path.getmtime(d) > t for d in deps)


def build_variant(name: str, defines: list[str]) -> str:
    """Kernel experiments: libraydar_cuda_<name>.so with extra -D flags, next to the product library (git-ignored;
    selected with RAYDAR_CUDA_LIB).  Not part of build()."""
    out = os.path.join(HERE, f"libraydar_cuda_{name}.so")
    cmd = [
        nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
        "-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++",
        "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall", "-shared",
        "-I", os.path.join(ROOT, "include"), "-I", CSRC, *[f"-D{d}" for d in defines],
        *[os.path.join(CSRC, s) for s in SOURCES], "-o", out, "-ldl",
    ]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return out


def build(force: bool = False, verbose: bool = False, extra: list[str] | None = None) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [
        nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
        "-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++",
        "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall", "-shared",
        "-I", os.path.join(ROOT, "include"), "-I", CSRC,
        *[os.path.join(CSRC, s) for s in SOURCES], "-o", LIB, "-ldl",
    ]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += extra or []
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    build_cli()
    return LIB


def build_cli() -> str:
    """raydar-cuda: the headless driver (mirror of the reference's `raydar` binary) linked against the library."""
    exe = os.path.join(HERE, "host", "raydar-cuda")
    cmd = ["/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", "-O2", "-std=c++17", "-Wall",
           "-I", os.path.join(ROOT, "include"), os.path.join(HERE, "host", "raydar_cuda_main.cpp"),
           "-L", HERE, "-lraydar_cuda", "-Wl,-rpath,$ORIGIN/..", "-o", exe]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stdout + res.stderr)
    return exe


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
