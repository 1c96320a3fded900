"""Pins the oracle against the reference ITSELF: oracle/_ref is the reference's own six compute shaders (their GLSL text, read from
the reference tree at build time and compiled as C++ with a small GLSL emulation header) driven in the reference's dispatch order.
oracle/ow_oracle.cpp — the hand restatement every parity test uses — must reproduce it. Skips where neither the reference tree
nor a prebuilt oracle/_ref/libow_ref.so exists."""
import numpy as np
import pytest

from oracle import ref as R
from oracle.oracle import OracleSim

pytestmark = pytest.mark.skipif(not R.available(), reason="reference tree / oracle/_ref not available")
C1 = dict(L=1000, wind_speed=40.0, wind_dir=(1.0, 1.0), amplitude=2.0, suppression=0.1)


def sims(N, noise, **kw):
    p = dict(C1)
    p.update(kw)
    ref = R.RefSim(N, p["L"], p["wind_speed"], p["wind_dir"], p["amplitude"], p["suppression"], noise)
    orc = OracleSim(N, float(p["L"]), p["wind_speed"], p["wind_dir"], p["amplitude"], p["suppression"], noise, threads=8)
    return ref, orc


@pytest.mark.parametrize("N", [256, 512])
def test_initial_spectrum_and_butterfly_table_match_the_reference_shaders(noise, N):
    ref, orc = sims(N, noise)
    ra, rb = ref.h0()
    oa, ob = orc.h0()
    assert np.isfinite(ra).all() and np.isfinite(rb).all()
    # tilde_h0_k_cs.glsl, incl. the DC texel (-4000 * gaussian) and the clamp: equal to the last bit on all but a handful of texels
    # (one-ulp differences where the compiler groups a product differently)
    for r_, o_ in ((ra, oa), (rb, ob)):
        assert np.abs(r_ - o_).max() <= 2.5e-7 * np.abs(o_).max() and (r_ != o_).mean() < 1e-3
    assert ra[N // 2, N // 2, 0] == oa[N // 2, N // 2, 0]                # DC texel
    rtw, rbr = ref.twiddle()
    otw, obr = orc.twiddle()
    assert np.array_equal(rbr, obr)                                     # reverse_bits table
    assert np.array_equal(rtw[..., 2:], otw[..., 2:])                   # butterfly index pairs, every stage
    assert np.array_equal(rtw[..., :2], otw[..., :2])                   # twiddles


@pytest.mark.parametrize("N,t", [(256, 0.0), (256, 1.0), (256, 10.0), (512, 599.0 / 60.0)])
def test_frame_matches_the_reference_shaders(noise, N, t):
    ref, orc = sims(N, noise)
    r = ref.frame(np.float32(t))
    o = orc.frame(np.float32(t))
    for k in ("dy", "dx", "dz"):
        peak = np.abs(r[k]).max()
        assert np.abs(o[k] - r[k]).max() <= 2e-6 * peak, k               # same arithmetic, different summation grouping at most
    assert np.abs(o["normal"] - r["normal"]).max() <= 2e-6
    # SURVEY.md App. A.6 spot values (N=256, t=1) straight from the reference's shaders
    if N == 256 and t == 1.0:
        assert r["dy"][0, 0] == pytest.approx(0.127508, abs=2e-5) and r["dy"][128, 128] == pytest.approx(0.518812, abs=2e-5)
        assert r["dx"][17, 201] == pytest.approx(0.026620, abs=2e-5) and r["dz"][255, 3] == pytest.approx(0.174160, abs=2e-5)


def test_reference_shaders_with_other_parameters(noise):
    """Another wind, amplitude and (integer) patch size: the pin is not an accident of the default parameters."""
    ref, orc = sims(256, noise, L=700, wind_speed=80.0, wind_dir=(0.3, -1.0), amplitude=3.5, suppression=0.07)
    ra, _ = ref.h0()
    oa, _ = orc.h0()
    assert np.abs(ra - oa).max() <= 2.5e-7 * np.abs(oa).max()
    r, o = ref.frame(2.5), orc.frame(2.5)
    for k in ("dy", "dx", "dz"):
        assert np.abs(o[k] - r[k]).max() <= 2e-6 * np.abs(r[k]).max(), k


@pytest.mark.parametrize("name", ["c1_n256.npz", "c2_n512.npz"])
def test_committed_golden_fixtures_agree_with_the_reference_shaders(noise, name):
    """tests/golden/*.npz were generated from the oracle; here they are held against the reference's own shaders, so the fixtures the
    GPU parity tests use are pinned to the reference too (they travel to the GPU box; the reference tree does not)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name))
    N, st = int(g["N"]), int(g["stride"])
    ref = R.RefSim(N, 1000, 40.0, (1.0, 1.0), 2.0, 0.1, noise)
    ra, _ = ref.h0()
    assert np.abs(ra[::st, ::st] - g["h0k"]).max() <= 2.5e-7 * np.abs(ra).max()
    for i, t in enumerate(g["times"]):
        r = ref.frame(np.float32(t))
        for k in ("dy", "dx", "dz"):
            peak = g[f"{k}_{i}_stats"][0]
            assert np.abs(r[k][::st, ::st] - g[f"{k}_{i}"]).max() <= 2e-6 * peak, (k, i)
            assert np.abs(r[k]).max() == pytest.approx(peak, rel=1e-5)
        assert np.abs(r["normal"][::st, ::st] - g[f"normal_{i}"]).max() <= 2e-6
