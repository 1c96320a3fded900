"""GPU tests of the slab-decomposed frame (BASELINE config C5 at the sizes this round supports, SURVEY.md §8 e2/d2:
"parity is checked on a down-scaled N=4096 run of the identical slab code against the single-GPU path")."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import fft_ocean_waves_b200 as fow
from oracle import numpy_ref as R
from oracle.oracle import OracleSim

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = fow.OceanParams(L=1000.0, wind_speed=40.0, wind_dir=(1.0, 1.0), amplitude=2.0, suppression=0.1, choppiness=1.0)
NAMES = ("dy", "dx", "dz", "normal", "jacobian")


def test_seeded_noise_is_philox_and_matches_the_oracle():
    """ow_set_noise_seed: the device-generated planes equal the NumPy Philox restatement (checked through h0 and a frame)."""
    N, seed = 256, 32768
    nz = R.philox_noise(seed, N)
    orc = OracleSim(N, P.L, P.wind_speed, P.wind_dir, P.amplitude, P.suppression, nz, threads=8)
    with fow.FFTOceanWaves(N=N, cascades=[P], jacobian=True) as sim:
        sim.set_noise_seed(seed)
        sim.tilde_h0_k()
        a, b = sim.download("h0k"), sim.download("h0minusk")
        got = sim.frame(1.0)
    ra, rb = orc.h0()
    assert np.abs(a - ra).max() <= 2e-6 * np.abs(ra).max() and np.abs(b - rb).max() <= 2e-6 * np.abs(rb).max()
    ref = orc.frame(1.0, choppiness=1.0)
    for k in ("dy", "dx", "dz"):
        assert np.abs(got[k] - ref[k]).max() <= 1e-4 * np.abs(ref[k]).max()


@pytest.mark.parametrize("N", [256, 1024, 4096, 8192])
@pytest.mark.parametrize("transport", ["peer", "alltoall"])
def test_single_rank_slab_equals_single_gpu_path(N, transport):
    """world = 1: the slab kernels (permuted h0 rows, transposing sink with wrap-around halo columns, strided column
    pass, stencil without x wrap) must reproduce ow_step. Same phase functions, but the single-GPU row kernel is the
    persistent pipelined variant, so the compiler may contract a*b+c differently: agreement to fp32 round-off
    (1e-6 of peak, 100x tighter than the parity tolerance), not bit for bit. The CPU emulation, where both paths are
    one compilation, IS bit-exact (tests/test_emu.py)."""
    seed = 4096
    with fow.FFTOceanWaves(N=N, cascades=[P], jacobian=True) as one:
        one.set_noise_seed(seed)
        one.tilde_h0_k()
        refs = [one.frame(t) for t in (0.0, 2.5)]
    with fow.SlabOcean(N=N, params=P, jacobian=True, transport=transport) as sim:
        sim.init(seed)
        assert sim.world == 1 and sim.transport == transport
        for t, ref in zip((0.0, 2.5), refs):
            sim.update(t)
            sim.sync()
            for k in NAMES:
                got = sim.download(k)
                tol = 1e-6 * float(np.abs(ref[k]).max()) + (2e-6 if k in ("normal", "jacobian") else 0.0)   # fp32 round-off of peak (+2 ulp-ish for O(1) images)
                assert np.abs(got - ref[k]).max() <= tol, (k, t)


def test_c5_full_size_slab_equals_single_gpu_path():
    """BASELINE config C5 at its quoted size, N = 32768, on one rank: the slab kernels (line decomposition A = 16, transposing
    sink, halo columns, column slab stencils) against ow_step on the same Philox seed. One image at a time: a context is ~90 GB."""
    import psutil
    import torch
    N, seed, t = 32768, 32768, 1.0
    if psutil.virtual_memory().available < 40e9 or torch.cuda.mem_get_info()[0] < 120e9:
        pytest.skip("not enough host or device memory for N = 32768")
    ref = {}
    with fow.FFTOceanWaves(N=N, cascades=[P], jacobian=True) as one:
        one.set_noise_seed(seed)
        one.tilde_h0_k()
        one.update(t)
        one.sync()
        for k in ("dy", "dz", "jacobian"):
            ref[k] = one.download(k)
    with fow.SlabOcean(N=N, params=P, jacobian=True) as sim:
        sim.init(seed)
        sim.update(t)
        sim.sync()
        for k in ("dy", "dz", "jacobian"):
            got = sim.download(k)
            tol = 1e-6 * float(np.abs(ref[k]).max()) + (2e-6 if k == "jacobian" else 0.0)
            assert np.abs(got - ref[k]).max() <= tol, k
            del got
    assert float(np.abs(ref["dy"]).max()) > 0 and np.isfinite(ref["jacobian"]).all()


def test_slab_api_errors():
    lib = fow.load_library()
    with pytest.raises(fow.OceanWavesError):
        fow.SlabOcean(N=300, params=P)                       # unsupported N
    with fow.SlabOcean(N=256, params=P) as sim:
        with pytest.raises(fow.OceanWavesError):
            sim.update(0.0)                                  # rows before init -> OW_ERR_STATE
        sim.init(1)
        with pytest.raises(fow.OceanWavesError):
            sim.download("jacobian")                         # created without the Jacobian flag
        assert lib.ow_slab_rows(sim.backend._h, 0.0, 7, None) == 1    # unknown transport -> OW_ERR_INVALID


def _torchrun(world, *worker_args, timeout=900):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "slab_worker.py"), *[str(a) for a in worker_args]]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


@pytest.mark.parametrize("N,world", [(1024, 2), (4096, 4), (8192, 2)])
def test_multi_rank_slab_on_one_gpu(N, world):
    """Runs on ANY GPU box, also one with a single device: `world` processes, each with its own ow_slab context on GPU 0, row results
    stored into each other's receive buffers through CUDA IPC peer mappings (the same ow_slab_rows(OW_SLAB_PEER_STORES) code that
    crosses NVLink between devices), ordered by host-side barriers over gloo; rank 0 compares the gathered column slabs with the
    single-GPU path. N = 8192 exercises the N = A*B line decomposition under the slab geometry."""
    r = _torchrun(world, N, "shared")
    assert r.returncode == 0 and "SLAB OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("N", [1024, 4096, 8192])
def test_multi_rank_slab_over_nvlink(N):
    """2 (or 4/8 when present) ranks under torchrun: peer-store and all-to-all transports vs the single-GPU path."""
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 8 if ngpu >= 8 else 4 if ngpu >= 4 else 2
    r = _torchrun(world, N)
    assert r.returncode == 0 and "SLAB OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
