"""fft-ocean-waves_b200 — B200-native Tessendorf hot path behind a C ABI (include/oceanwaves.h).

Python here is only the test/bench harness over the C ABI; the product is `lib/liboceanwaves.so`
(hand-written sm_100a kernels, csrc/). There is NO CPU fallback: importing works without a GPU (so the
CPU-only checks can inspect symbols), but creating a simulation without the library or without a CUDA
device raises.

The directory name has a hyphen; import it through the alias module at the repo root:

    import fft_ocean_waves_b200 as fow
    sim = fow.FFTOceanWaves(N=512); sim.init(); sim.update(t=1.0); dy = sim.download("dy")

One grid over several GPUs (torchrun, one process per GPU): `fow.SlabOcean` (slab.py).
"""
from __future__ import annotations

from .sim import (  # noqa: F401
    EXPORTED_SYMBOLS,
    FFTOceanWaves,
    OceanParams,
    OceanWavesError,
    default_noise,
    lib_path,
    load_library,
)
from .slab import SlabOcean, slab_plan  # noqa: F401

__all__ = ["FFTOceanWaves", "OceanParams", "OceanWavesError", "default_noise", "lib_path", "load_library",
           "EXPORTED_SYMBOLS", "SlabOcean", "slab_plan"]
