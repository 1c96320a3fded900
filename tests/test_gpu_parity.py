"""GPU parity tests (run on the B200 box): the CUDA path through the C ABI vs the CPU oracle on the same
seeded inputs, vs the committed golden fixtures, plus size-independent properties at full sizes.

Tolerance (BASELINE.json north_star): max abs error <= 1e-4 x peak displacement amplitude per channel; the
per-channel RMS error is also checked. Normals/Jacobian: 1e-4 absolute (unit vectors / O(1) values).
"""
import os

import numpy as np
import pytest

import fft_ocean_waves_b200 as fow
from oracle import numpy_ref as R
from oracle.oracle import OracleSim
from tests.conftest import rng_noise

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
REL_TOL = 1e-4
C1 = dict(L=1000.0, wind_speed=40.0, wind_dir=(1.0, 1.0), amplitude=2.0, suppression=0.1, choppiness=1.0)


def params(**kw):
    d = dict(C1)
    d.update(kw)
    return fow.OceanParams(**d)


def oracle_for(N, noise, p=None, threads=8):
    p = p or params()
    return OracleSim(N, p.L, p.wind_speed, p.wind_dir, p.amplitude, p.suppression, noise, threads=threads)


def check_frame(got, ref, tag=""):
    for k in ("dy", "dx", "dz"):
        peak = float(np.abs(ref[k]).max())
        err = np.abs(got[k].astype(np.float64) - ref[k])
        assert err.max() <= REL_TOL * peak, f"{tag}{k}: max err {err.max():.3e} vs peak {peak:.3e}"
        assert np.sqrt((err ** 2).mean()) <= 0.25 * REL_TOL * peak, f"{tag}{k}: rms err"
    assert np.abs(got["normal"] - ref["normal"]).max() < 1e-4, tag + "normal"
    if "jacobian" in got and "jacobian" in ref:
        assert np.abs(got["jacobian"] - ref["jacobian"]).max() < 1e-4, tag + "jacobian"


def test_h0_matches_oracle(noise):
    """tilde_h0_k: same Phillips/Box-Muller values as the shader transcription, incl. the DC texel and clamp."""
    for N in (256, 512):
        with fow.FFTOceanWaves(N=N, cascades=[params()]) as sim:
            sim.init(noise)
            a, b = sim.download("h0k"), sim.download("h0minusk")
        ra, rb = oracle_for(N, noise).h0()
        scale = np.abs(ra).max()
        assert np.abs(a - ra).max() <= 2e-6 * scale and np.abs(b - rb).max() <= 2e-6 * scale
        assert a[N // 2, N // 2, 0] == pytest.approx(ra[N // 2, N // 2, 0], rel=1e-6)     # -4000 * gaussian
        assert np.isfinite(a).all() and np.isfinite(b).all()


@pytest.mark.parametrize("t", [0.0, 1.0, 10.0])
def test_c1_n256_vs_oracle(noise, t):
    """BASELINE config C1: N=256, L=1000, wind 40, A=2, reference PNG noise."""
    with fow.FFTOceanWaves(N=256, cascades=[params()], jacobian=True) as sim:
        sim.init(noise)
        got = sim.frame(t)
    check_frame(got, oracle_for(256, noise).frame(t, choppiness=1.0), f"t={t} ")


def test_c2_n512_sweep_spot_frames(noise):
    """BASELINE config C2: N=512 time sweep t_f = f/60; parity at f in {0, 1, 299, 599}."""
    orc = oracle_for(512, noise)
    with fow.FFTOceanWaves(N=512, cascades=[params()], jacobian=True) as sim:
        sim.init(noise)
        for f in (0, 1, 299, 599):
            t = np.float32(f / 60.0)
            check_frame(sim.frame(float(t)), orc.frame(t, choppiness=1.0), f"f={f} ")


@pytest.mark.parametrize("name", ["c1_n256.npz", "c2_n512.npz"])
def test_golden_fixtures(noise, name):
    g = np.load(os.path.join(GOLD, name))
    N, st = int(g["N"]), int(g["stride"])
    with fow.FFTOceanWaves(N=N, cascades=[params()], jacobian=True) as sim:
        sim.init(noise)
        a = sim.download("h0k")
        assert np.abs(a[::st, ::st] - g["h0k"]).max() <= 2e-6 * 6000
        for i, t in enumerate(g["times"]):
            got = sim.frame(float(t))
            for k in ("dy", "dx", "dz"):
                peak = g[f"{k}_{i}_stats"][0]
                assert np.abs(got[k][::st, ::st] - g[f"{k}_{i}"]).max() <= REL_TOL * peak, (k, i)
                v = got[k].astype(np.float64)
                assert np.abs(v).max() == pytest.approx(peak, rel=1e-4)
                assert np.sqrt((v ** 2).mean()) == pytest.approx(g[f"{k}_{i}_stats"][1], rel=1e-4)
            assert np.abs(got["normal"][::st, ::st] - g[f"normal_{i}"]).max() < 1e-4
            assert np.abs(got["jacobian"][::st, ::st] - g[f"jacobian_{i}"]).max() < 1e-4


def test_c3_n2048_with_jacobian():
    """BASELINE config C3: N=2048, default_rng(2048) noise (1:1 lookup), Jacobian on, t=1."""
    N = 2048
    nz = rng_noise(2048, N)
    with fow.FFTOceanWaves(N=N, cascades=[params()], jacobian=True) as sim:
        sim.init(nz)
        got = sim.frame(1.0)
        a, b = sim.download("h0k"), sim.download("h0minusk")
    ref = oracle_for(N, nz).frame(1.0, choppiness=1.0)
    check_frame(got, ref, "C3 ")
    # independent fp64 closed form on the GPU's own h0
    r64 = R.frame_from_h0(a[..., 0] + 1j * a[..., 1].astype(np.float64), b[..., 0] + 1j * b[..., 1].astype(np.float64),
                          N, 1000.0, 1.0, 1.0)
    check_frame(got, r64, "C3/fp64 ")


def test_c3_n2048_png_noise(noise):
    with fow.FFTOceanWaves(N=2048, cascades=[params()]) as sim:
        sim.init(noise)
        got = sim.frame(1.0)
    check_frame(got, oracle_for(2048, noise).frame(1.0), "C3/png ")


def c4_cascade(c):
    """BASELINE config C4 cascade c: L = 100*1.08^c, wind 10+0.5c, dir angle 2 pi c/64, A=2."""
    ang = 2 * np.pi * c / 64
    return params(L=float(100.0 * 1.08 ** c), wind_speed=float(10 + 0.5 * c), wind_dir=(float(np.cos(ang)), float(np.sin(ang))))


def test_c4_cascades_n1024_batch():
    """Config C4 (subset): independent cascades in ONE context must equal each cascade's own oracle."""
    N, ids = 1024, [0, 17, 40, 63]
    ps = [c4_cascade(c) for c in ids]
    with fow.FFTOceanWaves(N=N, cascades=ps) as sim:
        for i, c in enumerate(ids):
            sim.set_noise(rng_noise(1024 + c, N), cascade=i)
        sim.tilde_h0_k()
        sim.update(1.0)
        sim.sync()
        for i, c in enumerate(ids):
            got = {k: sim.download(k, i) for k in ("dy", "dx", "dz", "normal")}
            ref = oracle_for(N, rng_noise(1024 + c, N), ps[i]).frame(1.0)
            check_frame(got, ref, f"cascade {c} ")


def test_n4096_vs_fp64():
    N = 4096
    nz = rng_noise(4096, N)
    with fow.FFTOceanWaves(N=N, cascades=[params()], jacobian=True) as sim:
        sim.init(nz)
        got = sim.frame(2.5)
        a, b = sim.download("h0k"), sim.download("h0minusk")
    r64 = R.frame_from_h0(a[..., 0] + 1j * a[..., 1].astype(np.float64), b[..., 0] + 1j * b[..., 1].astype(np.float64),
                          N, 1000.0, 2.5, 1.0)
    check_frame(got, r64, "N4096 ")


def test_known_answers_impulse_and_mode():
    """KATs through the ABI: impulse at the DC texel -> 1/N^2 everywhere; single mode -> cosine rows."""
    N = 256
    with fow.FFTOceanWaves(N=N, cascades=[params()]) as sim:
        a = np.zeros((N, N, 2), np.float32)
        a[N // 2, N // 2, 0] = 1.0
        sim.set_h0(a, np.zeros_like(a))
        f = sim.frame(0.0)
        assert np.allclose(f["dy"], 1.0 / (N * N), rtol=1e-6, atol=0)
        assert np.abs(f["dx"]).max() < 1e-12 and np.abs(f["dz"]).max() < 1e-12
        assert np.allclose(f["normal"], (0, 1, 0, 1), atol=1e-7)
        m = 5
        a[:] = 0
        a[N // 2 + m, N // 2, 1] = 1.0          # imaginary unit amplitude at ky = +m: dy = -sin(2 pi m y/N)/N^2
        sim.set_h0(a, np.zeros_like(a))
        f = sim.frame(0.0)
        y = np.arange(N)
        expect = (-np.sin(2 * np.pi * m * y / N))[:, None].repeat(N, 1) / (N * N)
        assert np.abs(f["dy"] - expect).max() < 1e-5 / (N * N)


def test_nyquist_row_and_column_are_literal(noise):
    """Energy only on the Nyquist row/column (texel index 0), where the mirror texel is the texel itself."""
    N = 256
    rng = np.random.default_rng(7)
    a = np.zeros((N, N, 2), np.float32)
    b = np.zeros((N, N, 2), np.float32)
    a[0, :, :] = rng.standard_normal((N, 2))
    a[:, 0, :] = rng.standard_normal((N, 2))
    b[0, :, :] = rng.standard_normal((N, 2))
    b[:, 0, :] = rng.standard_normal((N, 2))
    orc = oracle_for(N, noise)
    orc.set_h0(a, b)
    with fow.FFTOceanWaves(N=N, cascades=[params()]) as sim:
        sim.set_h0(a, b)
        check_frame(sim.frame(3.0), orc.frame(3.0), "nyquist ")


def test_linearity_and_determinism(noise):
    """Size-independent properties at N=2048: D(h0_1 + h0_2) = D(h0_1) + D(h0_2); bitwise repeatability."""
    N = 2048
    rng = np.random.default_rng(1)
    h = [rng.standard_normal((N, N, 2)).astype(np.float32) for _ in range(4)]
    with fow.FFTOceanWaves(N=N, cascades=[params()]) as sim:
        sim.set_h0(h[0], h[1])
        f1 = sim.frame(1.5)
        f1b = sim.frame(1.5)
        sim.set_h0(h[2], h[3])
        f2 = sim.frame(1.5)
        sim.set_h0(h[0] + h[2], h[1] + h[3])
        f12 = sim.frame(1.5)
    for k in ("dy", "dx", "dz"):
        assert np.array_equal(f1[k], f1b[k])
        peak = np.abs(f12[k]).max()
        assert np.abs(f12[k] - (f1[k] + f2[k])).max() <= 2e-5 * peak
    # sum identities
    assert abs(f12["dx"].astype(np.float64).sum()) < 1e-3 and abs(f12["dz"].astype(np.float64).sum()) < 1e-3


def test_multi_slot_time_sweep_equals_single_steps(noise):
    """ow_step_multi: one cascade evaluated at several times in one launch == separate ow_step calls."""
    N, times = 512, [0.0, 0.5, 2.0, 9.98]
    with fow.FFTOceanWaves(N=N, cascades=[params()], n_slots=len(times)) as sim:
        sim.init(noise)
        singles = [sim.frame(t)["dy"].copy() for t in times]
        sim.update_multi([0] * len(times), times)
        sim.sync()
        for i in range(len(times)):
            assert np.array_equal(sim.download("dy", i), singles[i])
        sim.set_group_size(1)            # different launch grouping, same numbers
        sim.update_multi([0] * len(times), times)
        sim.sync()
        assert sim.last_group_count() == len(times)
        assert sim.last_launch_count() == 3 * len(times)
        for i in range(len(times)):
            assert np.array_equal(sim.download("dy", i), singles[i])


def test_exact_and_fast_phase_paths(noise):
    """e^{iwt}: SFU sin/cos after an exact 2*pi reduction (default for |w t| < 2e4) vs full-range sincosf
    (OW_FLAG_EXACT_SINCOS, and automatically for huge t). Both must sit inside the parity tolerance."""
    N = 512
    orc = oracle_for(N, noise)
    for t in (9.98, 300.0):
        ref = orc.frame(np.float32(t))
        for exact in (False, True):
            with fow.FFTOceanWaves(N=N, cascades=[params()], exact_sincos=exact) as sim:
                sim.init(noise)
                check_frame(sim.frame(t), ref, f"t={t} exact={exact} ")
    t = 40000.0      # max |w t| ~ 2.7e5 > 2e4: the launcher must fall back to sincosf on its own
    with fow.FFTOceanWaves(N=N, cascades=[params()]) as sim:
        sim.init(noise)
        check_frame(sim.frame(t), orc.frame(np.float32(t)), "t=4e4 ")


def test_api_errors(noise):
    with fow.FFTOceanWaves(N=256, cascades=[params()]) as sim:
        with pytest.raises(fow.OceanWavesError):
            sim.update(0.0)                      # ow_step before ow_init_spectrum -> OW_ERR_STATE
        sim.init(noise)
        with pytest.raises(fow.OceanWavesError):
            sim.download("jacobian")             # context created without OW_FLAG_JACOBIAN
        with pytest.raises(fow.OceanWavesError):
            sim.download("dy", 5)                # slot out of range
        with pytest.raises(fow.OceanWavesError):
            sim.update_multi([0, 0], [0.0, 1.0])  # more entries than slots
        lib = fow.load_library()
        assert lib.ow_gl_step(sim._h, 0.0) == 4   # OW_ERR_NO_GL: nothing registered
        assert lib.ow_gl_register(sim._h, 1, 2, 3, 4) == 4   # no GL context on this box


def test_set_params_requires_reinit(noise):
    with fow.FFTOceanWaves(N=256, cascades=[params()]) as sim:
        sim.init(noise)
        d40 = sim.frame(1.0)["dy"]
        sim.set_params(0, params(wind_speed=80.0))
        with pytest.raises(fow.OceanWavesError):
            sim.update(1.0)
        sim.tilde_h0_k()
        d80 = sim.frame(1.0)["dy"]
    ref = OracleSim(256, 1000.0, 80.0, (1.0, 1.0), 2.0, 0.1, noise, threads=8).frame(1.0)["dy"]
    assert np.abs(d80 - ref).max() <= REL_TOL * np.abs(ref).max() and not np.allclose(d40, d80)


@pytest.mark.parametrize("N", [256, 512, 1024, 2048])
def test_fused_normal_epilogue_equals_separate_normal_kernel(noise, N):
    """OW_FLAG_FUSED_NORMALS: the normal map out of the column kernel's dy tiles (ow_col2_kernel: interior quads from shared memory,
    seam quads by the last-arriving neighbour tile) instead of the stand-alone normal kernel. Same column kernel, same stencil on the
    same heights: identical displacement and normal images, every texel written."""
    with fow.FFTOceanWaves(N=N, cascades=[params()]) as sim:
        sim.init(noise)
        sim.set_column_kernel(3, 0)          # the same column kernel, normal map by the separate kernel
        sep = sim.frame(2.0)
        assert sim.last_launch_count() == 3
    with fow.FFTOceanWaves(N=N, cascades=[params()], fused_normals=True) as sim:
        sim.init(noise)
        sim.update(0.0)                      # fill the normal buffer with another frame first: stale texels would show
        fused = sim.frame(2.0)
        assert sim.last_launch_count() == 3          # row, column + interior normals, seam quads
    for k in ("dy", "dx", "dz", "normal"):
        assert np.array_equal(fused[k], sep[k]), k
    check_frame(fused, oracle_for(N, noise).frame(2.0), f"fused N={N} ")


@pytest.mark.parametrize("N,t", [(128, 1.5), (256, 1.0), (512, 9.98), (1024, 2.0)])
def test_cuda_path_vs_the_reference_shaders_themselves(noise, N, t):
    """The CUDA path against oracle/_ref — the reference's OWN compute shaders (GLSL text compiled for the CPU, dispatched in the
    reference's order) — without the hand-written oracle in between. Same tolerance as everywhere: 1e-4 of peak, plus the RMS bound."""
    from oracle import ref as refmod
    if not refmod.available():
        pytest.skip("oracle/_ref not available (needs the reference tree at build time or the prebuilt library)")
    ref = refmod.RefSim(N, 1000, 40.0, (1.0, 1.0), 2.0, 0.1, noise, threads=8).frame(np.float32(t))
    with fow.FFTOceanWaves(N=N, cascades=[params()]) as sim:
        sim.init(noise)
        a = sim.download("h0k")
        got = sim.frame(float(np.float32(t)))
    ra, _ = refmod.RefSim(N, 1000, 40.0, (1.0, 1.0), 2.0, 0.1, noise).h0()
    assert np.abs(a - ra).max() <= 2e-6 * np.abs(ra).max()
    check_frame(got, ref, f"vs reference shaders N={N} ")


def test_smallest_grid_n128(noise):
    """N = 128 (the reference's DISPLACEMENT_MAP_SIZE is a free #define, src/main.cpp:17): plain kernels, checked against the oracle, with the
    Jacobian, through the graph and through a multi-slot launch; N = 64 is rejected (the normal kernel's warp tiles are 128 columns wide)."""
    ref = OracleSim(128, 1000.0, 40.0, (1.0, 1.0), 2.0, 0.1, noise, threads=4).frame(np.float32(2.5), choppiness=1.0)
    with fow.FFTOceanWaves(N=128, cascades=[params()], jacobian=True, n_slots=3) as sim:
        sim.init(noise)
        got = sim.frame(2.5)
        check_frame(got, ref, "N=128 ")
        assert np.abs(got["jacobian"] - ref["jacobian"]).max() < 1e-4
        sim.update_multi([0, 0, 0], [0.0, 2.5, 1.0])
        sim.sync()
        assert np.array_equal(sim.download("dy", 1), got["dy"]) and np.array_equal(sim.download("normal", 1), got["normal"])
    with pytest.raises(fow.OceanWavesError):
        fow.FFTOceanWaves(N=64, cascades=[params()])
