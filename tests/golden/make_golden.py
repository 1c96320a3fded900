"""Regenerates tests/golden/*.npz from the CPU oracle (oracle/ow_oracle.cpp).

The reference has no golden vectors for this path and its executable cannot run here (OpenGL), so these fixtures
are generated from the ORACLE, which tests/test_ref_pin.py pins to the reference's own shaders compiled for the CPU. Each file holds a
strided sample of every output image plus whole-image statistics, for BASELINE.md configs C1 and C2.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import OracleSim  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
noise = np.fromfile(os.path.join(ROOT, "fft-ocean-waves_b200", "data", "noise_LDR_LLL1_R.u8"), dtype=np.uint8).reshape(4, 256, 256)


def dump(name, N, times, stride):
    sim = OracleSim(N, 1000.0, 40.0, (1.0, 1.0), 2.0, 0.1, noise, threads=8)
    a, b = sim.h0()
    out = dict(N=N, stride=stride, times=np.asarray(times, np.float32), h0k=a[::stride, ::stride], h0minusk=b[::stride, ::stride],
               h0k_center=a[N // 2 - 2:N // 2 + 3, N // 2 - 2:N // 2 + 3], h0minusk_center=b[N // 2 - 2:N // 2 + 3, N // 2 - 2:N // 2 + 3])
    for i, t in enumerate(times):
        f = sim.frame(np.float32(t), choppiness=1.0)
        for k, v in f.items():
            out[f"{k}_{i}"] = v[::stride, ::stride].copy()
            v64 = v.astype(np.float64)
            out[f"{k}_{i}_stats"] = np.array([np.abs(v64).max(), np.sqrt((v64 ** 2).mean()), v64.sum()])
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, {k: getattr(v, "shape", v) for k, v in out.items() if not k.endswith("stats")})


if __name__ == "__main__":
    dump("c1_n256.npz", 256, [0.0, 1.0, 10.0], 8)
    dump("c2_n512.npz", 512, [0.0, 1.0 / 60.0, 299.0 / 60.0, 599.0 / 60.0], 16)
