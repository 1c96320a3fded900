"""In-tree build of liboceanwaves.so for sm_100a (explicit nvcc; no JIT cache, so the .so travels with the repo).

    python fft-ocean-waves_b200/build.py [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "liboceanwaves.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# 128: "loop is not reachable" - the phase functions pick one of two loop forms per plan with a compile-time condition and return from the first
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-strict-aliasing", "-diag-suppress=128"]
# (source, extra flags). The init kernels keep the shader's unfused fp32 operation order.
SOURCES = [
    ("ow_frame_kernels.cu", []),
    ("ow_big_kernels.cu", []),
    ("ow_mega_kernels.cu", []),
    ("ow_init_kernels.cu", ["-fmad=false"]),
    ("ow_pack_kernels.cu", []),
    ("ow_compose_kernels.cu", []),
    ("ow_api.cu", []),
    ("ow_slab.cu", []),
]
HEADERS = ["ow_fft.cuh", "ow_kernels.cuh", "ow_config.cuh", "ow_frame_kernels.cuh", "ow_async.cuh", "ow_internal.h", os.path.join("..", "..", "include", "oceanwaves.h")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the CUDA library cannot be built (there is no CPU fallback)")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False, defines=(), tag: str = "") -> str:
    """defines/tag: a VARIANT build for A/B runs (tools/): extra -D flags, objects and library suffixed with `tag`
    (lib/liboceanwaves_<tag>.so; select it with OCEANWAVES_LIB). The product build has neither."""
    global OBJDIR, LIB
    if tag:
        OBJDIR = os.path.join(HERE, "build", tag)
        LIB = os.path.join(LIBDIR, f"liboceanwaves_{tag}.so")
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = _nvcc()
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    objs, jobs = [], []
    for src, extra in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [nvcc, *ARCH, *COMMON, *extra, *[f"-D{d}" for d in defines], "-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd))
            jobs.append(cmd)
    procs = [subprocess.Popen(cmd) for cmd in jobs]          # the translation units are independent: compile them side by side
    for cmd, pr in zip(jobs, procs):
        if pr.wait() != 0:
            raise subprocess.CalledProcessError(pr.returncode, cmd)
    if force or _stale(LIB, objs):
        cmd = [nvcc, *ARCH, "-shared", "-cudart", "static", "-o", LIB, *objs]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    tags = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--tag=")]
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, defines=defs, tag=tags[0] if tags else ""))
