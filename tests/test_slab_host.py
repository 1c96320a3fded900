"""CPU tests of the multi-GPU (slab) path's HOST logic: the partition arithmetic, the Philox noise
both paths share, and SlabOcean's orchestration over a real torch.distributed group (gloo, world_size 2) with a NumPy
stand-in for the device kernels (tests/slab_double.py). The kernels' own slab index algebra is covered bit for bit by
tests/test_emu.py::test_emulated_slab_frame_equals_full_frame; the GPU run is tests/test_gpu_slab.py."""
import os
import socket

import numpy as np
import pytest

import fft_ocean_waves_b200 as fow
from oracle import numpy_ref as R


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    z = R.philox4x32_10(np.zeros((1, 4), np.uint32), (0, 0))[0]
    assert [int(v) for v in z] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    o = R.philox4x32_10(np.full((1, 4), 0xFFFFFFFF, np.uint32), (0xFFFFFFFF, 0xFFFFFFFF))[0]
    assert [int(v) for v in o] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    p = R.philox4x32_10(np.array([[0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344]], np.uint32), (0xA4093822, 0x299F31D0))[0]
    assert [int(v) for v in p] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_philox_noise_planes_are_uniform_bytes():
    n = R.philox_noise(32768, 256)
    assert n.shape == (4, 256, 256) and n.dtype == np.uint8
    counts = np.bincount(n.ravel(), minlength=256)
    assert counts.min() > 800 and counts.max() < 1250          # 1024 expected per value
    assert not np.array_equal(n[0], n[1]) and not np.array_equal(R.philox_noise(1, 256), n)


@pytest.mark.parametrize("N,world", [(256, 1), (256, 2), (1024, 8), (4096, 8), (4096, 2)])
def test_slab_plan_partitions_rows_and_columns(N, world):
    seen_rows, seen_cols = [], []
    for r in range(world):
        p = fow.slab_plan(N, world, r)
        assert p["pairs_per_rank"] * world == N // 2 and p["cols_per_rank"] * world == N
        assert p["padded_cols"] == p["cols_per_rank"] + 2 * p["halo"] and p["padded_cols"] % 16 == 0
        assert p["block_bytes"] == p["pairs_per_rank"] * 3 * p["padded_cols"] * 8
        rows = p["h0_rows"]
        assert len(rows) == 2 * p["pairs_per_rank"]
        # the mirror of every owned row is owned too (rows 0 and N/2 mirror onto themselves)
        assert {(N - v) % N for v in rows} == set(rows)
        seen_rows += rows
        seen_cols += list(range(p["first_col"], p["first_col"] + p["cols_per_rank"]))
    assert sorted(seen_rows) == list(range(N)) and sorted(seen_cols) == list(range(N))
    with pytest.raises(ValueError):
        fow.slab_plan(N, 3, 0)


def test_slab_needs_the_cuda_library_or_a_gpu():
    """No CPU fallback: without a CUDA device the product path must raise, not compute."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(fow.OceanWavesError):
        fow.SlabOcean(N=256, params=fow.OceanParams())


def _reference_frame(N, L, t, lam, seed=7):
    nz = R.philox_noise(seed, N)
    a, b = R.h0_fields(N, L, 40.0, (1, 1), 2.0, 0.1, nz)
    return a, b, R.frame_from_h0(a, b, N, L, t, lam)


def _check(full, ref):
    for k in ("dy", "dx", "dz"):
        assert np.abs(full[k] - ref[k]).max() <= 2e-6 * np.abs(ref[k]).max(), k
    assert np.abs(full["normal"] - ref["normal"]).max() < 1e-5
    assert np.abs(full["jacobian"] - ref["jacobian"]).max() < 1e-5


def test_single_process_slab_matches_closed_form():
    from tests.slab_double import NumpySlabBackend
    N, L, t = 256, 1000.0, 1.0
    a, b, ref = _reference_frame(N, L, t, 1.0)
    sim = fow.SlabOcean(N=N, params=fow.OceanParams(L=L, wind_speed=40.0, choppiness=1.0), transport="alltoall",
                        backend=NumpySlabBackend(N, 1, 0, a, b, L))
    sim.init(7)
    sim.update(t)
    _check({k: sim.gather(k) for k in ("dy", "dx", "dz", "normal", "jacobian")}, ref)


def _worker(rank, world, port, N, L, t, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests.slab_double import NumpySlabBackend
        a, b, ref = _reference_frame(N, L, t, 1.0)
        sim = fow.SlabOcean(N=N, params=fow.OceanParams(L=L, wind_speed=40.0, choppiness=1.0), transport="alltoall",
                            backend=NumpySlabBackend(N, world, rank, a, b, L))
        sim.init(7)
        for tt in (0.25, t):                       # two frames: the second must not see stale blocks
            sim.update(tt)
        full = {k: sim.gather(k) for k in ("dy", "dx", "dz", "normal", "jacobian")}
        _check(full, ref)
        assert sim.exchange_bytes_per_frame() == sim.plan["block_bytes"] * (world - 1)
        mine = sim.download("dy")
        assert mine.shape == (N, N // world)
        assert np.array_equal(mine, full["dy"][:, rank * (N // world):(rank + 1) * (N // world)])
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_two_rank_slab_over_gloo():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 256, 1000.0, 1.0, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
