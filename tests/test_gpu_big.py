"""GPU tests of the N = A*B line decomposition that grids above N = 4096 use (BASELINE config C5 is quoted on N = 32768):
forced onto N = 1024/2048 against the direct kernels, N = 8192 against the dense fp64 closed form, and analytic
known answers at N = 16384 and N = 32768 (a dense reference does not fit in reasonable host time there)."""
import numpy as np
import pytest

import fft_ocean_waves_b200 as fow
from oracle import numpy_ref as R

pytestmark = pytest.mark.gpu
P = fow.OceanParams(L=1000.0, wind_speed=40.0, wind_dir=(1.0, 1.0), amplitude=2.0, suppression=0.1, choppiness=1.0)


@pytest.mark.parametrize("N", [1024, 2048])
def test_forced_line_decomposition_equals_direct_kernels(N):
    frames = {}
    for four in (False, True):
        with fow.FFTOceanWaves(N=N, cascades=[P], jacobian=True, four_step=four) as sim:
            sim.set_noise_seed(N)
            sim.tilde_h0_k()
            frames[four] = sim.frame(1.25)
    for k in ("dy", "dx", "dz"):
        peak = np.abs(frames[False][k]).max()
        assert np.abs(frames[True][k] - frames[False][k]).max() <= 2e-6 * peak, k
    assert np.abs(frames[True]["normal"] - frames[False]["normal"]).max() < 1e-5
    assert np.abs(frames[True]["jacobian"] - frames[False]["jacobian"]).max() < 1e-5
    with pytest.raises(fow.OceanWavesError):
        fow.FFTOceanWaves(N=512, cascades=[P], four_step=True)


def test_n8192_vs_fp64_closed_form():
    """N = 8192 = A * B with A = 4, B = 2048 (the production line decomposition, OW_BIG_B) against Re(ifft2(ifftshift(H))) in double precision."""
    N = 8192
    with fow.FFTOceanWaves(N=N, cascades=[P]) as sim:
        sim.set_noise_seed(8192)
        sim.tilde_h0_k()
        got = sim.frame(1.0)
        a, b = sim.download("h0k"), sim.download("h0minusk")
    hk = a[..., 0].astype(np.complex128) + 1j * a[..., 1]
    hm = b[..., 0].astype(np.complex128) + 1j * b[..., 1]
    del a, b
    hdy, hdx, hdz = R.spectra(hk, hm, N, P.L, 1.0)
    del hk, hm
    for k, H in (("dy", hdy), ("dx", hdx), ("dz", hdz)):
        ref = R.displacement(H)
        peak = np.abs(ref).max()
        err = np.abs(got[k] - ref)
        assert err.max() <= 1e-4 * peak, (k, err.max(), peak)
        assert np.sqrt((err ** 2).mean()) <= 0.25e-4 * peak, k
        if k == "dy":
            assert np.abs(got["normal"] - R.normal_map(ref)).max() < 1e-4
        del ref, err


@pytest.mark.parametrize("N", [16384, 32768])
def test_large_grid_known_answers(N):
    """A = 4 and A = 8. Impulse at the DC texel -> 1/N^2 everywhere; one imaginary mode at (kx, ky) = (+mx, +my) ->
    dy = -sin(2 pi (mx x + my y)/N)/N^2 (inversion_cs.glsl's (-1)^(x+y) undoes the centred index), checked on whole rows
    and columns that cross every sub-line boundary."""
    import psutil
    import torch
    need_host, need_dev = 11.0 * N * N * 4, 22.0 * N * N * 4          # bytes: host arrays of this test / device buffers of the context
    if psutil.virtual_memory().available < 1.3 * need_host or torch.cuda.mem_get_info()[0] < 1.1 * need_dev:
        pytest.skip("not enough host or device memory for this grid")
    with fow.FFTOceanWaves(N=N, cascades=[P]) as sim:
        a = np.zeros((N, N, 2), np.float32)
        z = np.zeros((N, N, 2), np.float32)
        a[N // 2, N // 2, 0] = 1.0
        sim.set_h0(a, z)
        sim.update(0.0)
        sim.sync()
        dy = sim.download("dy")
        assert np.allclose(dy, 1.0 / (float(N) * N), rtol=1e-6, atol=0)
        assert np.abs(sim.download("dx")).max() < 1e-14
        a[N // 2, N // 2, 0] = 0.0
        mx, my = 37, 4099                      # my > 4096: exercises every residue ka of the column decomposition
        a[N // 2 + my, N // 2 + mx, 1] = 1.0
        sim.set_h0(a, z)
        del a, z
        sim.update(0.0)
        sim.sync()
        dy = sim.download("dy")
        n2 = float(N) * N
        idx = np.arange(N, dtype=np.float64)
        for y in (0, 1, 4095, 4096, N // 2 + 3, N - 1):
            expect = -np.sin(2 * np.pi * ((mx * idx + my * y) % N) / N) / n2
            assert np.abs(dy[y] - expect).max() < 2e-5 / n2, ("row", y)
        for x in (0, 5, 4097, N - 2):
            expect = -np.sin(2 * np.pi * ((mx * x + my * idx) % N) / N) / n2
            assert np.abs(dy[:, x] - expect).max() < 2e-5 / n2, ("col", x)
        nm = sim.download("normal")
        assert np.isfinite(nm).all() and np.abs(np.linalg.norm(nm[::257, ::263, :3], axis=-1) - 1).max() < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("N,four", [(1024, True), (2048, True), (8192, False)])
def test_cluster_line_decomposition_equals_scratch_path(N, four):
    """The N = A*B decomposition as thread-block clusters (sub-lines combined through distributed shared memory, no global scratch;
    16-column and 8-column tiles) must give the SAME images as the two-kernel scratch path: identical arithmetic, different plumbing."""
    with fow.FFTOceanWaves(N=N, cascades=[P], jacobian=True, four_step=four) as sim:
        sim.set_noise_seed(N)
        sim.tilde_h0_k()
        assert sim.line_clusters() == 0          # default: the scratch path
        sim.set_line_clusters(-1)
        if sim.line_clusters() == 0:
            pytest.skip("this device cannot co-schedule the clusters")
        sim.set_line_clusters(0)
        ref = sim.frame(1.5)
        assert sim.last_launch_count() == 5
        for mode in (1, 3, 7):
            sim.set_line_clusters(mode)
            got = sim.frame(1.5)
            used = sim.line_clusters()
            assert sim.last_launch_count() == 5 - (used & 1) - ((used >> 1) & 1)
            for k in ("dy", "dx", "dz", "normal", "jacobian"):
                assert np.array_equal(got[k], ref[k]), (mode, used, k)
