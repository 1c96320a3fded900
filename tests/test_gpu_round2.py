"""GPU tests of the round-2 additions, all through the C ABI: the CUDA-graph single-frame path, per-cascade readiness,
the packed output set, all 64 cascades of BASELINE config C4, and a random-spectrum check of the
N = A*B line decomposition at N = 16384 against an independent fp64 DFT of sampled output rows and columns."""
import numpy as np
import pytest

import fft_ocean_waves_b200 as fow
from oracle import numpy_ref as R
from oracle.oracle import OracleSim
from tests.conftest import rng_noise

pytestmark = pytest.mark.gpu
C1 = dict(L=1000.0, wind_speed=40.0, wind_dir=(1.0, 1.0), amplitude=2.0, suppression=0.1, choppiness=1.0)


def params(**kw):
    d = dict(C1)
    d.update(kw)
    return fow.OceanParams(**d)


@pytest.mark.parametrize("N", [256, 1024])
def test_graph_path_equals_plain_launches(noise, N):
    """ow_step as ONE cudaGraphLaunch (default) vs the same kernels launched one by one: bit-identical images, and the time
    patched into the graph's row-kernel node really changes from frame to frame."""
    times = [0.0, 1.0, 1.0, 7.25, 0.5]
    with fow.FFTOceanWaves(N=N, cascades=[params()], jacobian=True) as sim:
        sim.init(noise)
        sim.set_graph(False)
        plain = [sim.frame(t) for t in times]
        sim.set_graph(True)
        graph = [sim.frame(t) for t in times]
        assert sim.last_launch_count() == 3 and sim.last_group_count() == 1
    for a, b, t in zip(plain, graph, times):
        for k in ("dy", "dx", "dz", "normal", "jacobian"):
            assert np.array_equal(a[k], b[k]), (k, t)
    assert np.array_equal(graph[1]["dy"], graph[2]["dy"]) and not np.array_equal(graph[0]["dy"], graph[1]["dy"])
    ref = OracleSim(N, 1000.0, 40.0, (1.0, 1.0), 2.0, 0.1, noise, threads=8).frame(np.float32(7.25), choppiness=1.0)
    for k in ("dy", "dx", "dz"):
        assert np.abs(graph[3][k] - ref[k]).max() <= 1e-4 * np.abs(ref[k]).max()     # parity tolerance: 1e-4 of peak


def test_graph_with_several_cascades_and_large_time(noise):
    """Several cascades in one graph (one chain of nodes per launch group), and the automatic switch to the full-range sincos
    variant (its own graph) when |w t| leaves the fast path's range."""
    ps = [params(), params(L=400.0, wind_speed=25.0, wind_dir=(0.3, -1.0))]
    with fow.FFTOceanWaves(N=512, cascades=ps) as sim:
        sim.init(noise)
        for t in (2.0, 40000.0, 3.0):
            sim.set_graph(True)
            sim.update(t)
            sim.sync()
            g = [{k: sim.download(k, i) for k in ("dy", "dz", "normal")} for i in range(2)]
            sim.set_graph(False)
            sim.update(t)
            sim.sync()
            for i in range(2):
                for k in ("dy", "dz", "normal"):
                    assert np.array_equal(g[i][k], sim.download(k, i)), (t, i, k)


def test_readiness_is_per_cascade(noise):
    """ow_set_params / ow_set_h0 touch ONE cascade: the others stay usable, and a stale cascade is refused (OW_ERR_STATE)."""
    ps = [params(), params(wind_speed=60.0)]
    with fow.FFTOceanWaves(N=256, cascades=ps, n_slots=2) as sim:
        a = np.zeros((256, 256, 2), np.float32)
        a[128, 128, 0] = 1.0
        sim.set_h0(a, np.zeros_like(a), cascade=0)           # only cascade 0 has a spectrum now
        sim.update_multi([0], [0.0])
        sim.sync()
        assert np.allclose(sim.download("dy", 0), 1.0 / 65536, rtol=1e-6)
        with pytest.raises(fow.OceanWavesError):
            sim.update_multi([0, 1], [0.0, 0.0])             # cascade 1 was never initialised
        with pytest.raises(fow.OceanWavesError):
            sim.update(0.0)
        sim.init(noise)
        d1 = sim.frame(1.0, slot=1)["dy"]
        sim.set_params(1, params(wind_speed=20.0))
        sim.update_multi([0], [1.0])                          # cascade 0 is untouched by cascade 1's edit
        with pytest.raises(fow.OceanWavesError):
            sim.update_multi([1], [1.0])
        sim.tilde_h0_k_cascade(1)                             # re-generate just that cascade
        d1b = sim.frame(1.0, slot=1)["dy"]
    ref = OracleSim(256, 1000.0, 20.0, (1.0, 1.0), 2.0, 0.1, noise, threads=8).frame(1.0)["dy"]
    assert np.abs(d1b - ref).max() <= 1e-4 * np.abs(ref).max() and not np.allclose(d1, d1b)


@pytest.mark.parametrize("mode,N,jac", [("f32", 512, True), ("f16", 512, True), ("f16", 2048, False)])
def test_packed_outputs_decode_to_the_reference_formats(noise, mode, N, jac):
    """OW_FLAG_PACKED_*: (dx,dy,dz,J) RGBA32F/RGBA16F + RG16_SNORM normal.xz, decoded the way the consumer's shader does
    (INTEGRATION.md), against the default R32F / RGBA32F outputs of the same frame. Stated tolerances:
      f32 displacement: exact.  f16 displacement: 1e-3 of the channel's peak (half precision: 2^-11 relative).
      normal: 1e-4 absolute per component (SNORM16 step 3.1e-5; y rebuilt from x, z)."""
    with fow.FFTOceanWaves(N=N, cascades=[params()], jacobian=jac, packed=mode) as sim:
        sim.init(noise)
        full = sim.frame(1.0)                                  # the reference formats are still produced
        dec = sim.decode_packed(sim.download_packed(0))
        assert sim.last_launch_count() == 4                    # row, column, normal, pack
        assert sim.packed_bytes() == N * N * ((16 if mode == "f32" else 8) + 4)
    for k in ("dx", "dy", "dz"):
        if mode == "f32":
            assert np.array_equal(dec[k], full[k]), k
        else:
            assert np.abs(dec[k] - full[k]).max() <= 1e-3 * np.abs(full[k]).max(), k
    if jac:
        assert np.abs(dec["jacobian"] - full["jacobian"]).max() <= (0.0 if mode == "f32" else 1e-3 * np.abs(full["jacobian"]).max())
    else:
        assert np.all(dec["jacobian"] == 1.0)
    assert np.abs(dec["normal"] - full["normal"]).max() <= 1e-4


def test_packed_api_errors(noise):
    with fow.FFTOceanWaves(N=256, cascades=[params()]) as sim:
        sim.init(noise)
        assert sim.packed_bytes() == 0
        with pytest.raises(fow.OceanWavesError):
            sim.download_packed_async(0, 0x1000, 16)
        lib = fow.load_library()
        assert lib.ow_gl_register_packed(sim._h, 1, 2) == 3      # OW_ERR_STATE: no packed set in this context
    with pytest.raises(fow.OceanWavesError):
        import ctypes as C
        h = C.c_void_p()
        p = params().to_c()
        rc = fow.load_library().ow_create(256, 1, 1, C.byref(p), 0, 0x10 | 0x20, C.byref(h))
        if rc != 0:
            raise fow.OceanWavesError("both packed flags")


def c4_cascade(c):
    ang = 2 * np.pi * c / 64
    return params(L=float(100.0 * 1.08 ** c), wind_speed=float(10 + 0.5 * c), wind_dir=(float(np.cos(ang)), float(np.sin(ang))))


def test_c4_all_64_cascades():
    """BASELINE config C4 complete: 64 cascades N=1024 in ONE context and one ow_step, every cascade against its own oracle."""
    N = 1024
    ps = [c4_cascade(c) for c in range(64)]
    with fow.FFTOceanWaves(N=N, cascades=ps) as sim:
        for c in range(64):
            sim.set_noise(rng_noise(1024 + c, N), cascade=c)
        sim.tilde_h0_k()
        sim.update(1.0)
        sim.sync()
        for c in range(64):
            ref = OracleSim(N, ps[c].L, ps[c].wind_speed, ps[c].wind_dir, ps[c].amplitude, ps[c].suppression, rng_noise(1024 + c, N), threads=16).frame(1.0)
            for k in ("dy", "dx", "dz"):
                got = sim.download(k, c)
                peak = float(np.abs(ref[k]).max())
                assert np.abs(got - ref[k]).max() <= 1e-4 * peak, (c, k)
            assert np.abs(sim.download("normal", c) - ref["normal"]).max() < 1e-4, c


def test_n16384_random_spectrum_vs_fp64_dft_of_sampled_lines():
    """N = 16384 (line decomposition A = 8): sampled output rows AND columns of dy, dx, dz against a direct fp64 evaluation of
    D = Re(ifft2(ifftshift(H))) on the GPU's own h0 (Philox noise) — independent of every FFT in this repository.
    Tolerance: 1e-4 of the channel's peak, as everywhere."""
    import torch
    N, t, L = 16384, 1.0, 1000.0
    if torch.cuda.mem_get_info()[0] < 40e9:
        pytest.skip("not enough device memory")
    ys = np.array([0, 1, 4097, N // 2, N - 3])
    xs = np.array([0, 2, 5000, N // 2 + 1, N - 1])
    with fow.FFTOceanWaves(N=N, cascades=[params()]) as sim:
        sim.set_noise_seed(16384)
        sim.tilde_h0_k()
        sim.update(t)
        sim.sync()
        got = {k: sim.download(k) for k in ("dy", "dx", "dz")}
        h0k, h0m = sim.download("h0k"), sim.download("h0minusk")
    idx = np.arange(N) - N / 2.0
    k1 = ((2.0 * np.float32(np.pi) * idx.astype(np.float32)) / np.float32(L)).astype(np.float64)      # the shader's fp32 k, then fp64
    sh = (np.arange(N) + N // 2) % N                           # ifftshift: spectrum index a' = (a + N/2) mod N sits at DFT index a
    pos = np.empty(N, np.int64); pos[sh] = np.arange(N)        # DFT index of texel index i
    Trow = {k: np.zeros((len(ys), N), np.complex128) for k in got}   # T_y[b] = sum_a H'[a][b] e^{2 pi i a y/N}
    col = {k: np.zeros((N, len(xs)), np.complex128) for k in got}    # T_x[a] = sum_b H'[a][b] e^{2 pi i b x/N}, per texel row
    ex = np.exp(2j * np.pi * np.outer(pos, xs) / N)             # [texel column u][x]
    CH = 1024
    for r0 in range(0, N, CH):
        v = slice(r0, r0 + CH)
        A = h0k[v, :, 0].astype(np.float64) + 1j * h0k[v, :, 1]
        B = h0m[v, :, 0].astype(np.float64) + 1j * h0m[v, :, 1]
        kx, ky = k1[None, :], k1[v, None]
        km = np.maximum(np.sqrt(kx * kx + ky * ky), 1e-5)
        ph = np.sqrt(9.81 * km) * t
        e = np.cos(ph) + 1j * np.sin(ph)
        H = {"dy": A * e + B * np.conj(e)}
        H["dx"] = -1j * (kx / km) * H["dy"]
        H["dz"] = -1j * (ky / km) * H["dy"]
        ey = np.exp(2j * np.pi * np.outer(ys, pos[v]) / N)      # [y][texel row]
        for k in got:
            Trow[k] += ey @ H[k]
            col[k][v] = H[k] @ ex
    for k in got:
        peak = float(np.abs(got[k]).max())
        # rows: D[y][x] = Re( (1/N^2) sum_b T_y[b] e^{2 pi i b x/N} ), b = DFT index of texel column u
        Ty = np.zeros_like(Trow[k]); Ty[:, pos] = Trow[k]
        rows = np.real(np.fft.ifft(Ty, axis=1)) / N
        assert np.abs(got[k][ys, :] - rows).max() <= 1e-4 * peak, (k, "rows")
        Tx = np.zeros_like(col[k]); Tx[pos, :] = col[k]
        cols = np.real(np.fft.ifft(Tx, axis=0)) / N
        assert np.abs(got[k][:, xs] - cols).max() <= 1e-4 * peak, (k, "columns")


def test_jacobian_against_spectral_derivatives():
    """SURVEY.md §8 f1's validation target: the CUDA Jacobian (central differences riding the normal walk) against the Jacobian from
    EXACT spectral derivatives in fp64 (oracle/numpy_ref.jacobian_spectral), on the same band-limited physical waves at N = 256,
    512 and 1024: within the discretisation error (|k| L/N)^2/6 of the derivative terms, and that error falls 4x per doubling of N."""
    L, lam, t, err = 1000.0, 1.0, 1.0, {}
    for N in (256, 512, 1024):
        a, b = R.band_limited_h0(N, 5, seed=7, amplitude=2.0)
        a32 = np.stack([a.real, a.imag], -1).astype(np.float32)
        b32 = np.stack([b.real, b.imag], -1).astype(np.float32)
        with fow.FFTOceanWaves(N=N, cascades=[params(L=L, choppiness=lam)], jacobian=True) as sim:
            sim.set_h0(a32, b32)
            got = sim.frame(t)
        a64, b64 = a32[..., 0] + 1j * a32[..., 1].astype(np.float64), b32[..., 0] + 1j * b32[..., 1].astype(np.float64)
        js = R.jacobian_spectral(a64, b64, N, L, t, lam)
        ref = R.frame_from_h0(a64, b64, N, L, t, lam)
        assert js.min() < 0.8 and js.max() > 1.2
        assert np.abs(got["jacobian"] - ref["jacobian"]).max() < 1e-4           # the same finite differences in fp64
        err[N] = np.abs(got["jacobian"] - js).max()
    assert err[256] < 1e-3 and err[512] < 2.5e-4 and err[1024] < 1.5e-4
    assert 3.0 < err[256] / err[512] < 5.0, err


def test_cascade_blend_matches_the_restatement(noise):
    """ow_sample_points / ow_compose_grid (the consumer's grid_tes.glsl:60-64 displacement summed over cascades with
    blending weights, LINEAR/REPEAT taps) against oracle/numpy_ref.blend_cascades on the frames the context itself produced."""
    N, Ls, lams = 256, (250.0, 1000.0, 4000.0), (0.5, 1.0, 0.75)
    casc = [params(L=L, choppiness=lam, wind_speed=10.0 + 10.0 * i) for i, (L, lam) in enumerate(zip(Ls, lams))]
    with fow.FFTOceanWaves(N=N, cascades=casc, n_slots=4) as sim:
        with pytest.raises(fow.OceanWavesError):
            sim.sample_points(np.zeros((1, 2)), [0], [1.0])                     # nothing stepped yet
        sim.init(noise)
        sim.update_multi([0, 1, 2, 1], [1.0, 1.0, 1.0, 5.0])                     # slot 3: cascade 1 at another time
        sim.sync()
        frames = [{k: sim.download(k, s) for k in ("dy", "dx", "dz", "normal")} for s in range(4)]
        rng = np.random.default_rng(11)
        pts = rng.uniform(-2.0e4, 2.0e4, (4000, 2)).astype(np.float32)
        pts[:64, 0] = (np.arange(64) + 0.5) * Ls[1] / N                          # texel centres of cascade 1, row 7
        pts[:64, 1] = 7.5 * Ls[1] / N
        slots, w = [0, 1, 2, 3], [0.3, 1.0, 1.5, -0.25]
        got = sim.sample_points(pts, slots, w, displacement_scale=0.5)
        off, nrm = R.blend_cascades([frames[s] for s in slots], [Ls[0], Ls[1], Ls[2], Ls[1]], [lams[0], lams[1], lams[2], lams[1]], w, 0.5,
                                    pts[:, 0].astype(np.float64), pts[:, 1].astype(np.float64))
        peak = np.abs(off[:, :3]).max()
        assert np.abs(got["offset"] - off).max() < 2e-5 * peak
        assert np.abs(got["normal"] - nrm).max() < 1e-4      # slopes n.x/n.y of steep normals (n.y ~ 0.2) in fp32
        one = sim.sample_points(pts[:64], [1], [1.0], displacement_scale=1.0)      # one cascade at its texel centres = the texels
        assert np.abs(one["offset"][:, 1] - frames[1]["dy"][7, :64]).max() < 1e-5 * np.abs(frames[1]["dy"]).max()
        M, origin, extent = 192, (-333.0, 125.0), 1500.0
        grid = sim.compose_grid(M, origin, extent, slots, w, displacement_scale=0.5)
        gi, gj = np.meshgrid(np.arange(M), np.arange(M))
        gx = np.float32(origin[0]) + (gi.astype(np.float32) + np.float32(0.5)) * np.float32(extent / M)
        gz = np.float32(origin[1]) + (gj.astype(np.float32) + np.float32(0.5)) * np.float32(extent / M)
        goff, gnrm = R.blend_cascades([frames[s] for s in slots], [Ls[0], Ls[1], Ls[2], Ls[1]], [lams[0], lams[1], lams[2], lams[1]], w, 0.5,
                                      gx.astype(np.float64), gz.astype(np.float64))
        assert np.abs(grid["offset"] - goff).max() < 2e-5 * peak and np.abs(grid["normal"] - gnrm).max() < 1e-4
        for bad_slots, bad_w in (([4], [1.0]), ([0] * 17, [1.0] * 17), ([], [])):
            with pytest.raises(fow.OceanWavesError):
                sim.sample_points(pts[:4], bad_slots, bad_w)


def test_wall_clock_time_scale(noise):
    """The demo's clock (src/main.cpp:599): ow_step_wall_clock(w) is ow_step(offset + scale * w), bit for bit; scale 0 pauses."""
    with fow.FFTOceanWaves(N=256, cascades=[params()]) as sim:
        sim.init(noise)
        sim.set_time_scale(0.5, 2.0)
        sim.update_wall_clock(6.0)
        sim.sync()
        a = sim.download("dy")
        ref = sim.frame(np.float32(2.0) + np.float32(0.5) * np.float32(6.0))
        assert np.array_equal(a, ref["dy"])
        sim.set_time_scale(0.0, 5.0)
        sim.update_wall_clock(123.0)
        sim.sync()
        assert np.array_equal(sim.download("dy"), sim.frame(5.0)["dy"])
        with pytest.raises(fow.OceanWavesError):
            sim.set_time_scale(float("nan"), 0.0)


@pytest.mark.parametrize("N", [256, 512, 1024, 2048])
def test_row_and_column_kernel_variants_agree(noise, N):
    """Every row kernel (1 = CTA per row-pair group, 2 = persistent + register prefetch, 3 = persistent + cp.async.bulk/mbarrier staging)
    with every column kernel (1 = ow_col_kernel, 2 = ow_col2_kernel direct loads, 3 = ow_col2_kernel TMA-staged, 4 = ow_col_pipe_kernel), normal map fused
    into the column kernel or not. All of them must sit inside the parity tolerance of the oracle; variants differ only by fp32 round-off
    (2e-6 of peak); and for a FIXED (row, column) pair the fused normal map (interior quads out of shared memory + seams by the last
    arriving tile) must be bit-identical to the separate normal kernel's, the Jacobian (ow_jac_kernel) within round-off."""
    t = 2.0
    ref = OracleSim(N, 1000.0, 40.0, (1.0, 1.0), 2.0, 0.1, noise, threads=8).frame(np.float32(t), choppiness=1.0)
    with fow.FFTOceanWaves(N=N, cascades=[params()], jacobian=True, n_slots=5) as sim:
        sim.init(noise)
        base = None
        for rm in (1, 2, 3):
            sim.set_row_kernel(rm)
            for cm in (1, 2, 3, 4):
                sim.set_column_kernel(cm, 0)
                sep = sim.frame(t)
                assert sim.last_launch_count() == 3
                for k in ("dy", "dx", "dz"):
                    peak = float(np.abs(ref[k]).max())
                    assert np.abs(sep[k] - ref[k]).max() <= 1e-4 * peak, (rm, cm, k)
                    if base is not None:
                        assert np.abs(sep[k] - base[k]).max() <= 2e-6 * peak, (rm, cm, k)
                assert np.abs(sep["normal"] - ref["normal"]).max() < 1e-4 and np.abs(sep["jacobian"] - ref["jacobian"]).max() < 1e-4
                base = base or sep
                if cm in (1, 4):
                    continue
                sim.set_column_kernel(cm, 1)
                sim.update(0.0)                      # another frame first: stale texels / seam counters would show
                fused = sim.frame(t)
                assert sim.last_launch_count() == 4  # row, column(+ interior normals), seam quads, Jacobian
                for k in ("dy", "dx", "dz", "normal"):
                    assert np.array_equal(fused[k], sep[k]), (rm, cm, k)
                assert np.abs(fused["jacobian"] - sep["jacobian"]).max() <= 2e-6, (rm, cm)
        # several slots per launch, twice in a row (persistent loops over many tiles; seam counters re-armed between frames)
        sim.set_row_kernel(3)
        sim.set_column_kernel(3, 1)
        times = [0.0, 0.5, 2.0, 9.98, 2.0]
        for _ in range(2):
            sim.update_multi([0] * 5, times)
            sim.sync()
            assert np.array_equal(sim.download("dy", 2), sim.download("dy", 4))
            assert np.array_equal(sim.download("normal", 2), sim.download("normal", 4))
            assert np.abs(sim.download("dy", 2) - ref["dy"]).max() <= 1e-4 * np.abs(ref["dy"]).max()
            assert np.abs(sim.download("normal", 4) - ref["normal"]).max() < 1e-4


def test_fused_normals_without_jacobian(noise):
    with fow.FFTOceanWaves(N=512, cascades=[params()]) as sim:
        sim.init(noise)
        sim.set_column_kernel(1, 0)
        sep = sim.frame(1.0)
        sim.set_column_kernel(3, 1)
        fused = sim.frame(1.0)
        assert sim.last_launch_count() == 3 and sim.kernel_modes() == {"row": sim.kernel_modes()["row"], "column": 3, "fused": True}
        sim.set_graph(False)
        fused2 = sim.frame(1.0)
    for k in ("dy", "dx", "dz"):
        assert np.abs(fused[k] - sep[k]).max() <= 2e-6 * np.abs(sep[k]).max()
        assert np.array_equal(fused[k], fused2[k])
    assert np.abs(fused["normal"] - sep["normal"]).max() < 2e-6 and np.array_equal(fused["normal"], fused2["normal"])


@pytest.mark.parametrize("N", [256, 512, 1024])
@pytest.mark.parametrize("jac", [False, True])
def test_single_frame_latency_shapes_agree(noise, N, jac):
    """ow_step of ONE cascade runs latency-oriented kernel shapes (more threads per row pair, one pair per CTA, shorter normal-map walks;
    Cfg<N>::LAT): the same butterflies dealt out differently. The images must agree with the throughput shapes' to fp32 round-off (2e-6 of
    peak: the compiler may contract multiply-adds differently in another instantiation) - through the CUDA graph and through plain
    launches - and sit inside the parity tolerance of the oracle."""
    t = 3.25
    ref_o = OracleSim(N, 1000.0, 40.0, (1.0, 1.0), 2.0, 0.1, noise, threads=8).frame(np.float32(t), choppiness=1.0)
    with fow.FFTOceanWaves(N=N, cascades=[params()], jacobian=jac, n_slots=2) as sim:
        sim.init(noise)
        names = ["dy", "dx", "dz", "normal"] + (["jacobian"] if jac else [])
        sim.set_latency_shapes(False)
        ref = sim.frame(t)
        sim.set_latency_shapes(True)
        for graph in (True, False):
            sim.set_graph(graph)
            sim.update(0.5)                  # another frame first
            got = sim.frame(t)
            assert sim.last_launch_count() == 3
            for k in names:
                peak = float(np.abs(ref[k]).max())
                assert np.abs(got[k] - ref[k]).max() <= 2e-6 * max(peak, 1.0), (graph, k)
            for k in ("dy", "dx", "dz"):
                assert np.abs(got[k] - ref_o[k]).max() <= 1e-4 * np.abs(ref_o[k]).max(), (graph, k)


@pytest.mark.parametrize("N", [256, 512, 1024])
@pytest.mark.parametrize("jac", [False, True])
def test_one_kernel_frame_agrees_with_the_three_kernel_frame(noise, N, jac):
    """ow_set_frame_kernel(1): a launch group as ONE persistent kernel (ow_mega_kernel) that walks row, column and normal-map work items of
    consecutive frames behind per-frame dependency counters. Same phase functions as the separate kernels, so every slot of a multi-frame
    call must agree with the three-kernel path to fp32 round-off (2e-6 of peak), twice in a row (the counters are re-armed per launch), with
    one frame per call as well as many, and sit inside the parity tolerance of the oracle."""
    times = [0.0, 0.5, 2.0, 9.98, 2.0, 1.25, 7.5]
    n = len(times)
    ref_o = OracleSim(N, 1000.0, 40.0, (1.0, 1.0), 2.0, 0.1, noise, threads=8).frame(np.float32(2.0), choppiness=1.0)
    names = ["dy", "dx", "dz", "normal"] + (["jacobian"] if jac else [])
    with fow.FFTOceanWaves(N=N, cascades=[params()], jacobian=jac, n_slots=n) as sim:
        sim.init(noise)
        sim.update_multi([0] * n, times)
        sim.sync()
        ref = [{k: sim.download(k, s) for k in names} for s in range(n)]
        sim.set_frame_kernel(1)
        for rep in range(2):
            sim.update_multi([0] * n, [t + 1.0 for t in times])     # other frames first: stale outputs would show
            sim.update_multi([0] * n, times)
            sim.sync()
            for s in range(n):
                for k in names:
                    got = sim.download(k, s)
                    peak = float(np.abs(ref[s][k]).max())
                    assert np.abs(got - ref[s][k]).max() <= 2e-6 * max(peak, 1.0), (rep, s, k)
        for k in ("dy", "dx", "dz"):
            assert np.abs(sim.download(k, 2) - ref_o[k]).max() <= 1e-4 * np.abs(ref_o[k]).max(), k
        sim.update_multi([0], [2.0])                                # a single frame through the same kernel
        sim.sync()
        for k in names:
            assert np.abs(sim.download(k, 0) - ref[2][k]).max() <= 2e-6 * max(float(np.abs(ref[2][k]).max()), 1.0), ("single", k)
        assert sim.last_launch_count() == 1
