"""CPU check of the CUDA kernels' index algebra: the product's per-thread phase functions
(fft-ocean-waves_b200/csrc/ow_kernels.cuh) compiled as host code and run thread by thread (tests/emu),
compared with the oracle. Also asserts that every shared-memory access pattern is bank-conflict free.
This does not replace the GPU parity tests (tests/test_gpu_parity.py); it keeps index bugs off the GPU box."""
import numpy as np
import pytest

from oracle import numpy_ref as R
from oracle.oracle import OracleSim
from tests.emu import emu

REL_TOL = 1e-4   # north star: max abs error <= 1e-4 x peak displacement; the emulator lands near 1e-6


@pytest.mark.parametrize("R_", [2, 4, 8, 16])
def test_register_dft(R_):
    rng = np.random.default_rng(R_)
    x = (rng.standard_normal(R_) + 1j * rng.standard_normal(R_)).astype(np.complex64)
    ref = np.fft.ifft(x.astype(np.complex128)) * R_     # e^{+2 pi i nk/R}: the reference's inverse sign
    assert np.abs(emu.dft(x) - ref).max() < 2e-6 * np.abs(ref).max()


@pytest.mark.parametrize("N,t", [(128, 1.0), (256, 0.0), (256, 1.0), (256, 10.0), (512, 599.0 / 60.0), (1024, 1.0)])
def test_emulated_frame_matches_oracle(noise, N, t):
    s = OracleSim(N, 1000.0, 40.0, (1.0, 1.0), 2.0, 0.1, noise, threads=8)
    a, b = s.h0()
    ref = s.frame(t, choppiness=1.0)
    got = emu.frame(N, a, b, 1000.0, t, 1.0)
    for k in ("dy", "dx", "dz"):
        peak = np.abs(ref[k]).max()
        assert np.abs(got[k] - ref[k]).max() <= REL_TOL * peak, k
        assert np.abs(got[k] - ref[k]).max() <= 5e-6 * peak, k          # what fp32 actually delivers
    assert np.abs(got["normal"] - ref["normal"]).max() < 1e-4
    assert np.abs(got["jacobian"] - ref["jacobian"]).max() < 1e-4
    for phase, (req, wf) in got["conflicts"].items():
        assert req > 0 and wf == req, f"{phase}: {wf} wavefronts for {req} requests (bank conflicts)"


def test_emulated_2048_matches_fp64(noise):
    """Full-size C3 grid: compare with the fp64 closed form (the scalar oracle would also do, but slower)."""
    N = 2048
    rngn = np.random.default_rng(2048).integers(0, 256, (4, N, N), dtype=np.uint8)
    ak, bk = R.h0_fields(N, 1000.0, 40.0, (1, 1), 2.0, 0.1, rngn)
    a = np.stack([ak.real, ak.imag], -1).astype(np.float32)
    b = np.stack([bk.real, bk.imag], -1).astype(np.float32)
    ref = R.frame_from_h0(a[..., 0] + 1j * a[..., 1].astype(np.float64), b[..., 0] + 1j * b[..., 1].astype(np.float64),
                          N, 1000.0, 1.0, 1.0)
    got = emu.frame(N, a, b, 1000.0, 1.0, 1.0)
    for k in ("dy", "dx", "dz"):
        assert np.abs(got[k] - ref[k]).max() <= 5e-6 * np.abs(ref[k]).max(), k
    assert np.abs(got["normal"] - ref["normal"]).max() < 1e-5
    for phase, (req, wf) in got["conflicts"].items():
        assert wf == req, phase


def test_intermediate_is_row_transform_of_hermitian_part(noise):
    """The row kernel's output is the row IFFT of S = H + conj(H(-k)) for rows 0..N/2-1 (row 0 packs rows 0, N/2)."""
    N, t = 256, 1.0
    s = OracleSim(N, 1000.0, 40.0, (1.0, 1.0), 2.0, 0.1, noise, threads=4)
    a, b = s.h0()
    got = emu.frame(N, a, b, 1000.0, t, 1.0, want_inter=True)["inter"]
    H = R.spectra(a[..., 0] + 1j * a[..., 1].astype(np.float64), b[..., 0] + 1j * b[..., 1].astype(np.float64), N, 1000.0, t)
    idx = (-np.arange(N)) % N
    for c in range(3):
        S = H[c] + np.conj(H[c][idx][:, idx])
        rows = np.fft.ifft(S, axis=1) * N
        ref = rows[:N // 2].copy()
        ref[0] = rows[0].real + 1j * rows[N // 2].real
        g = got[c][..., 0] + 1j * got[c][..., 1]
        assert np.abs(g - ref).max() < 5e-6 * np.abs(ref).max()


@pytest.mark.parametrize("N,world", [(512, 1), (512, 2), (512, 4), (1024, 8), (1024, 2)])
def test_emulated_slab_frame_equals_full_frame(noise, N, world):
    """the slab decomposition (row pairs per rank, transposing sink with halo columns, column slabs
    without x wrap) must reproduce the single-GPU kernels bit for bit — same arithmetic, different addresses."""
    s = OracleSim(N, 1000.0, 40.0, (1.0, 1.0), 2.0, 0.1, noise, threads=4)
    a, b = s.h0()
    full = emu.frame(N, a, b, 1000.0, 1.0, 1.0)
    slab = emu.slab_frame(N, world, a, b, 1000.0, 1.0, 1.0)
    for k in ("dy", "dx", "dz", "normal", "jacobian"):
        assert np.array_equal(full[k], slab[k]), k


@pytest.mark.parametrize("N,A", [(512, 2), (1024, 4), (2048, 8)])
def test_emulated_line_decomposition_matches_oracle_and_direct_kernels(noise, N, A):
    """Grids above N=4096 split every line as N = A*B (Cooley-Tukey, sub-line B in one CTA). Same phase functions with
    B = 256 on a grid the oracle and the direct kernels can also do: all three must agree (covers A = 2, 4, 8)."""
    s = OracleSim(N, 1000.0, 40.0, (1.0, 1.0), 2.0, 0.1, noise, threads=8)
    a, b = s.h0()
    ref = s.frame(1.0, choppiness=1.0)
    big = emu.big_frame(N, A, a, b, 1000.0, 1.0, 1.0)
    direct = emu.frame(N, a, b, 1000.0, 1.0, 1.0)
    for k in ("dy", "dx", "dz"):
        peak = np.abs(ref[k]).max()
        assert np.abs(big[k] - ref[k]).max() <= 5e-6 * peak, k
        assert np.abs(big[k] - direct[k]).max() <= 5e-6 * peak, k
    assert np.abs(big["normal"] - ref["normal"]).max() < 1e-4
    assert np.abs(big["jacobian"] - ref["jacobian"]).max() < 1e-4


@pytest.mark.parametrize("N,staged", [(256, False), (256, True), (512, True), (1024, False), (1024, True), (2048, True)])
def test_emulated_fused_column_normal_kernel_equals_separate_kernels(noise, N, staged):
    """ow_col2_kernel: the normal map as the epilogue of the dy column tiles (three interior quads per 16-column tile out of the tile's
    shared-memory lines, the seam quad between two tiles from the stored heights), the Jacobian from its own walk, and - staged - stage 0
    fed from the TMA staging buffer (box layout of ColStage, incl. the extra row of pair id 0 and the zero-filled out-of-bounds row).
    Same stencil code on the same heights: bit-identical to the column kernel + stand-alone normal kernel, every texel written, and
    every shared-memory phase conflict-free."""
    if N == 2048:
        rngn = np.random.default_rng(2048).integers(0, 256, (4, N, N), dtype=np.uint8)
        ak, bk = R.h0_fields(N, 1000.0, 40.0, (1, 1), 2.0, 0.1, rngn)
        a = np.stack([ak.real, ak.imag], -1).astype(np.float32)
        b = np.stack([bk.real, bk.imag], -1).astype(np.float32)
    else:
        s = OracleSim(N, 1000.0, 40.0, (1.0, 1.0), 2.0, 0.1, noise, threads=4)
        a, b = s.h0()
    sep = emu.frame(N, a, b, 1000.0, 1.0, 1.0)
    fused = emu.frame_fused(N, a, b, 1000.0, 1.0, 1.0, staged=staged)
    assert not np.isnan(fused["normal"]).any() and not np.isnan(fused["jacobian"]).any()
    for k in ("dy", "dx", "dz", "normal", "jacobian"):
        assert np.array_equal(fused[k], sep[k]), k
    for phase, (req, wf) in fused["conflicts"].items():
        assert req > 0 and wf == req, f"{phase}: {wf} wavefronts for {req} requests (bank conflicts)"


def test_emulated_frame_matches_the_reference_shaders(noise):
    """The kernels' phase functions (emulated on the CPU) against oracle/_ref, the reference's own shaders compiled for the CPU."""
    from oracle import ref as refmod
    if not refmod.available():
        pytest.skip("oracle/_ref not available")
    N, t = 256, 1.0
    rs = refmod.RefSim(N, 1000, 40.0, (1.0, 1.0), 2.0, 0.1, noise)
    a, b = rs.h0()
    ref = rs.frame(t)
    got = emu.frame(N, a, b, 1000.0, t, 1.0)
    for k in ("dy", "dx", "dz"):
        assert np.abs(got[k] - ref[k]).max() <= 5e-6 * np.abs(ref[k]).max(), k
    assert np.abs(got["normal"] - ref["normal"]).max() < 1e-4
