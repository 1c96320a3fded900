"""ctypes front-end of the CPU ORACLE (oracle/ow_oracle.cpp) — test infrastructure, NOT the product.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this.
PARITY PIN: the reference has no golden vectors for this path, but its six compute shaders compile for the
CPU (oracle/_ref: their GLSL text through oracle/glsl_emu.hpp, dispatched in the reference's order), and this oracle reproduces
them — tests/test_ref_pin.py; see the header of ow_oracle.cpp.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libow_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (g++ only; no GPU, no reference tree needed)."""
    src = os.path.join(_HERE, "ow_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libow_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        fp = C.POINTER(C.c_float)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                    C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_max_threads.restype = C.c_int
        L.oracle_set_h0.argtypes = [C.c_void_p, fp, fp]
        L.oracle_get_h0.argtypes = [C.c_void_p, fp, fp]
        L.oracle_get_twiddle.argtypes = [C.c_void_p, fp, C.POINTER(C.c_int32)]
        L.oracle_spectrum.argtypes = [C.c_void_p, C.c_float, fp]
        L.oracle_frame.argtypes = [C.c_void_p, C.c_float, C.c_float, fp, fp, fp, fp, fp]
        L.oracle_normal_of.argtypes = [C.c_void_p, fp, fp]
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def max_threads() -> int:
    return int(lib().oracle_max_threads())


class OracleSim:
    """The reference's FFTOceanWaves sim state (src/main.cpp:1640-1646) driven on the CPU."""

    def __init__(self, N, L, wind_speed, wind_dir, amplitude, suppression, noise, threads=1):
        noise = np.ascontiguousarray(noise, dtype=np.uint8)
        assert noise.ndim == 3 and noise.shape[0] == 4
        self.N = int(N)
        self.L = float(L)
        self._noise = noise
        self._h = lib().oracle_create(self.N, float(L), float(wind_speed), float(wind_dir[0]), float(wind_dir[1]),
                                      float(amplitude), float(suppression),
                                      noise.ctypes.data_as(C.POINTER(C.c_uint8)), noise.shape[2], noise.shape[1],
                                      int(threads))
        if not self._h:
            raise ValueError("oracle_create failed (N must be a power of two)")

    def close(self):
        if self._h:
            lib().oracle_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def h0(self):
        n = self.N
        a = np.empty((n, n, 2), np.float32)
        b = np.empty((n, n, 2), np.float32)
        lib().oracle_get_h0(self._h, _fp(a), _fp(b))
        return a, b

    def set_h0(self, h0k, h0minusk):
        a = np.ascontiguousarray(h0k, np.float32).reshape(self.N, self.N, 2)
        b = np.ascontiguousarray(h0minusk, np.float32).reshape(self.N, self.N, 2)
        lib().oracle_set_h0(self._h, _fp(a), _fp(b))

    def twiddle(self):
        n = self.N
        l2 = n.bit_length() - 1
        tw = np.empty((n, l2, 4), np.float32)
        br = np.empty(n, np.int32)
        lib().oracle_get_twiddle(self._h, _fp(tw), br.ctypes.data_as(C.POINTER(C.c_int32)))
        return tw, br

    def spectrum(self, t):
        n = self.N
        out = np.empty((3, n, n, 2), np.float32)
        lib().oracle_spectrum(self._h, float(t), _fp(out))
        return out

    def frame(self, t, choppiness=None):
        """Returns dict(dy,dx,dz,normal[,jacobian]) for one update() (src/main.cpp:240-244)."""
        n = self.N
        dy = np.empty((n, n), np.float32)
        dx = np.empty((n, n), np.float32)
        dz = np.empty((n, n), np.float32)
        nm = np.empty((n, n, 4), np.float32)
        jac = np.empty((n, n), np.float32) if choppiness is not None else None
        lib().oracle_frame(self._h, float(t), float(choppiness) if choppiness is not None else -1.0,
                           _fp(dy), _fp(dx), _fp(dz), _fp(nm), _fp(jac) if jac is not None else None)
        out = dict(dy=dy, dx=dx, dz=dz, normal=nm)
        if jac is not None:
            out["jacobian"] = jac
        return out

    def normal_of(self, height):
        n = self.N
        h = np.ascontiguousarray(height, np.float32).reshape(n, n)
        nm = np.empty((n, n, 4), np.float32)
        lib().oracle_normal_of(self._h, _fp(h), _fp(nm))
        return nm
