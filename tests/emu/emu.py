"""ctypes front-end of the CPU emulation of the frame kernels (tests/emu/ow_emu.cu). TEST INFRASTRUCTURE ONLY:
it executes the product's per-thread phase functions as host code to check index algebra and bank conflicts
on the GPU-less build box. The product never loads it."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "libow_emu.so")
_lib = None


def build(force=False):
    src = os.path.join(HERE, "ow_emu.cu")
    csrc = os.path.join(ROOT, "fft-ocean-waves_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in ("ow_fft.cuh", "ow_kernels.cuh", "ow_config.cuh")]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        subprocess.check_call([nvcc, "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-cudart", "static",
                               "-Wno-deprecated-gpu-targets", "-o", LIB, src])
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        fp = C.POINTER(C.c_float)
        L.emu_frame.argtypes = [C.c_int, fp, fp, C.c_float, C.c_float, C.c_float, fp, fp, fp, fp, C.POINTER(C.c_long)]
        L.emu_dft.argtypes = [C.c_int, fp, fp]
        L.emu_frame_fused.argtypes = [C.c_int, fp, fp, C.c_float, C.c_float, C.c_float, C.c_int, fp, fp, fp, C.POINTER(C.c_long)]
        L.emu_big_frame.argtypes = [C.c_int, C.c_int, fp, fp, C.c_float, C.c_float, C.c_float, fp, fp, fp]
        L.emu_slab_frame.argtypes = [C.c_int, C.c_int, fp, fp, C.c_float, C.c_float, C.c_float, fp, fp, fp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


def dft(x):
    x = np.ascontiguousarray(x, np.complex64)
    out = np.empty_like(x)
    assert lib().emu_dft(len(x), _p(x.view(np.float32)), _p(out.view(np.float32))) == 0
    return out


PHASES = ["row0", "row1", "row2", "col0", "col1", "col2"]


def frame(N, h0k, h0minusk, L, t, choppiness=1.0, want_inter=False):
    a = np.ascontiguousarray(h0k, np.float32)
    b = np.ascontiguousarray(h0minusk, np.float32)
    disp = np.empty((3, N, N), np.float32)
    nm = np.empty((N, N, 4), np.float32)
    jac = np.empty((N, N), np.float32)
    inter = np.empty((3, N // 2, N, 2), np.float32) if want_inter else None
    stats = np.zeros(12, np.int64)
    rc = lib().emu_frame(N, _p(a), _p(b), float(L), float(t), float(choppiness), _p(inter), _p(disp), _p(nm), _p(jac),
                         stats.ctypes.data_as(C.POINTER(C.c_long)))
    assert rc == 0
    out = dict(dy=disp[0], dx=disp[1], dz=disp[2], normal=nm, jacobian=jac)
    out["conflicts"] = {n: (int(stats[2 * i]), int(stats[2 * i + 1])) for i, n in enumerate(PHASES)}
    if want_inter:
        out["inter"] = inter
    return out


def slab_frame(N, world, h0k, h0minusk, L, t, choppiness=1.0):
    """The slab-decomposed frame with `world` emulated ranks (row pairs -> peer stores -> column slabs), re-assembled."""
    a = np.ascontiguousarray(h0k, np.float32)
    b = np.ascontiguousarray(h0minusk, np.float32)
    disp = np.empty((3, N, N), np.float32)
    nm = np.empty((N, N, 4), np.float32)
    jac = np.empty((N, N), np.float32)
    rc = lib().emu_slab_frame(N, int(world), _p(a), _p(b), float(L), float(t), float(choppiness), _p(disp), _p(nm), _p(jac))
    assert rc == 0, rc
    return dict(dy=disp[0], dx=disp[1], dz=disp[2], normal=nm, jacobian=jac)


def big_frame(N, A, h0k, h0minusk, L, t, choppiness=1.0):
    """The N = A*B line decomposition used for N > 4096, emulated on a small grid (B = N/A = 256)."""
    a = np.ascontiguousarray(h0k, np.float32)
    b = np.ascontiguousarray(h0minusk, np.float32)
    disp = np.empty((3, N, N), np.float32)
    nm = np.empty((N, N, 4), np.float32)
    jac = np.empty((N, N), np.float32)
    rc = lib().emu_big_frame(N, int(A), _p(a), _p(b), float(L), float(t), float(choppiness), _p(disp), _p(nm), _p(jac))
    assert rc == 0, rc
    return dict(dy=disp[0], dx=disp[1], dz=disp[2], normal=nm, jacobian=jac)


def frame_fused(N, h0k, h0minusk, L, t, choppiness=1.0, staged=False):
    """The frame through ow_col2_kernel's phases: normal map as the epilogue of the dy column tiles (interior quads out of shared
    memory, seam quads from the stored heights), Jacobian from ow_jac_kernel's walk; staged=True feeds stage 0 from an emulated TMA
    staging buffer."""
    a = np.ascontiguousarray(h0k, np.float32)
    b = np.ascontiguousarray(h0minusk, np.float32)
    disp = np.empty((3, N, N), np.float32)
    nm = np.full((N, N, 4), np.nan, np.float32)
    jac = np.full((N, N), np.nan, np.float32)
    stats = np.zeros(14, np.int64)
    rc = lib().emu_frame_fused(N, _p(a), _p(b), float(L), float(t), float(choppiness), int(bool(staged)), _p(disp), _p(nm), _p(jac),
                               stats.ctypes.data_as(C.POINTER(C.c_long)))
    assert rc == 0, rc
    out = dict(dy=disp[0], dx=disp[1], dz=disp[2], normal=nm, jacobian=jac)
    out["conflicts"] = {n: (int(stats[2 * i]), int(stats[2 * i + 1])) for i, n in enumerate(PHASES + ["col_normals"])}
    return out
