"""ctypes front-end of oracle/_ref/libow_ref.so — the reference's OWN compute shaders compiled for the CPU (oracle/make_ref.py).
TEST INFRASTRUCTURE, not the product. Available where the reference tree is (this container: built on demand) or where a prebuilt
library travelled with the repository snapshot (the GPU box); elsewhere `available()` is False and the tests that need it skip."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import make_ref

_lib = None


def available(reference: str = "/root/reference") -> bool:
    try:
        make_ref.build(reference)
        return True
    except (FileNotFoundError, OSError, Exception):  # noqa: BLE001 - no reference tree, no compiler: the pin is simply not available
        return False


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(make_ref.build())
        fp = C.POINTER(C.c_float)
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.POINTER(C.c_uint8), C.c_int, C.c_int]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_get_h0.argtypes = [C.c_void_p, fp, fp]
        L.ref_get_twiddle.argtypes = [C.c_void_p, fp, C.POINTER(C.c_int32)]
        L.ref_frame.argtypes = [C.c_void_p, C.c_float, fp, fp, fp, fp]
        L.ref_set_threads.argtypes = [C.c_int]
        L.ref_max_threads.restype = C.c_int
        _lib = L
    return _lib


def max_threads() -> int:
    return int(lib().ref_max_threads())


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class RefSim:
    """The reference's sim state (src/main.cpp:1640-1646; u_L is an int uniform there) driven through its own shaders."""

    def __init__(self, N, L, wind_speed, wind_dir, amplitude, suppression, noise, threads=0):
        if threads:
            lib().ref_set_threads(int(threads))
        noise = np.ascontiguousarray(noise, dtype=np.uint8)
        assert noise.ndim == 3 and noise.shape[0] == 4 and float(L) == int(L)
        self.N = int(N)
        self._noise = noise
        self._h = lib().ref_create(self.N, int(L), float(wind_speed), float(wind_dir[0]), float(wind_dir[1]), float(amplitude),
                                   float(suppression), noise.ctypes.data_as(C.POINTER(C.c_uint8)), noise.shape[2], noise.shape[1])
        if not self._h:
            raise ValueError("ref_create failed (N must be a power of two)")

    def close(self):
        if self._h:
            lib().ref_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def h0(self):
        n = self.N
        a, b = np.empty((n, n, 2), np.float32), np.empty((n, n, 2), np.float32)
        lib().ref_get_h0(self._h, _fp(a), _fp(b))
        return a, b

    def twiddle(self):
        n = self.N
        l2 = n.bit_length() - 1
        tw = np.empty((n, l2, 4), np.float32)
        br = np.empty(n, np.int32)
        lib().ref_get_twiddle(self._h, _fp(tw), br.ctypes.data_as(C.POINTER(C.c_int32)))
        return tw, br

    def frame(self, t):
        n = self.N
        dy, dx, dz = (np.empty((n, n), np.float32) for _ in range(3))
        nm = np.empty((n, n, 4), np.float32)
        lib().ref_frame(self._h, float(t), _fp(dy), _fp(dx), _fp(dz), _fp(nm))
        return dict(dy=dy, dx=dx, dz=dz, normal=nm)
