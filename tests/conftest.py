import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_gpu() -> bool:
    try:
        import torch
        return bool(torch.cuda.is_available())
    except Exception:  # noqa: BLE001
        return False


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without a CUDA device: skip the gpu-marked tests instead of failing them one by one in ow_create.
    (The product itself still fails loudly there: tests/test_slab_host.py and tests/test_bench_contract.py check that.)"""
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (run with -m gpu on the B200 box)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def noise():
    import fft_ocean_waves_b200 as fow
    return fow.default_noise()


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.build()
    return oracle


# L=1000, wind 40 m/s along (1,1), A=2, suppression 0.1
C1 = dict(L=1000.0, wind_speed=40.0, wind_dir=(1.0, 1.0), amplitude=2.0, suppression=0.1)


@pytest.fixture(scope="session")
def c1_params():
    return dict(C1)


def rng_noise(seed: int, n: int) -> np.ndarray:
    """BASELINE.md C3/C4 noise: default_rng(seed) uniform bytes, one per texel (1:1 lookup)."""
    return np.random.default_rng(seed).integers(0, 256, (4, n, n), dtype=np.uint8)
