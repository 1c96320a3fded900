"""Independent fp64 closed-form statement of the hot path (test infrastructure, NOT the product).

Where ow_oracle.cpp replays the reference's literal dispatch chain in fp32, this file states WHAT that
chain computes, in closed form and in double precision:

    D_c = Re( ifft2( ifftshift( H_c(k, t) ) ) ),   c in {dy, dx, dz}

It is used only by tests/ to pin the oracle (the two must agree to fp32 round-off) and to give
size-independent properties at sizes where the scalar oracle is too slow.
"""
from __future__ import annotations

import numpy as np

G = 9.81
PI32 = np.float32(3.1415926535897932384626433832795)


def wave_vectors(N: int, L: float):
    """k for texel (iy, ix): tilde_h0_k_cs.glsl:76-77 (centred index). Returned in fp64 from fp32 inputs."""
    idx = np.arange(N, dtype=np.float32) - np.float32(N) / np.float32(2.0)
    k1 = ((np.float32(2.0) * PI32 * idx) / np.float32(L)).astype(np.float64)
    kx = np.broadcast_to(k1[None, :], (N, N))
    ky = np.broadcast_to(k1[:, None], (N, N))
    return kx, ky


def h0_amplitude(N, L, wind_speed, wind_dir, amplitude, suppression):
    """sqrt(Phillips)/sqrt(2) clamped to +-4000 (tilde_h0_k_cs.glsl:42-45, 84-88); DC texel -> -4000."""
    kx, ky = wave_vectors(N, L)
    wd = np.asarray(wind_dir, np.float64)
    wd = wd / np.sqrt((wd ** 2).sum())
    km = np.sqrt(kx ** 2 + ky ** 2)
    kmc = np.maximum(km, 1e-5)
    k2 = kmc ** 2
    Lp = wind_speed ** 2 / G
    with np.errstate(divide="ignore", invalid="ignore", over="ignore", under="ignore"):
        d = (kx * wd[0] + ky * wd[1]) / km
        P = amplitude * np.exp(-1.0 / (k2 * Lp * Lp)) * d * d * np.exp(-k2 * suppression ** 2) / (k2 * k2)
        a = np.sqrt(P) / np.sqrt(2.0)
    a = np.where(np.isnan(a), -4000.0, np.clip(a, -4000.0, 4000.0))  # fminf(fmaxf(NaN,-4000),4000) == -4000
    return a


def gauss_rnd(noise_u8, N):
    """Box-Muller on the nearest-sampled noise planes (tilde_h0_k_cs.glsl:51-68). noise_u8: (4,H,W) bytes."""
    _, H, W = noise_u8.shape
    iy = np.floor(np.arange(N) / N * H).astype(np.int64).clip(0, H - 1)
    ix = np.floor(np.arange(N) / N * W).astype(np.int64).clip(0, W - 1)
    n = noise_u8[:, iy[:, None], ix[None, :]].astype(np.float64) / 255.0
    n = np.clip(n, 0.001, 1.0)
    u0 = 2 * np.pi * n[0]
    v0 = np.sqrt(-2 * np.log(n[1]))
    u1 = 2 * np.pi * n[2]
    v1 = np.sqrt(-2 * np.log(n[3]))
    return v0 * np.cos(u0), v0 * np.sin(u0), v1 * np.cos(u1), v1 * np.sin(u1)


def h0_fields(N, L, wind_speed, wind_dir, amplitude, suppression, noise_u8):
    a = h0_amplitude(N, L, wind_speed, wind_dir, amplitude, suppression)
    r0, r1, r2, r3 = gauss_rnd(noise_u8, N)
    return (r0 + 1j * r1) * a, (r2 + 1j * r3) * a


def spectra(h0k, h0minusk, N, L, t):
    """H_dy, H_dx, H_dz at time t (tilde_h0_t_cs.glsl:70-131): no conjugate on h0minusk (quirk 1)."""
    kx, ky = wave_vectors(N, L)
    km = np.maximum(np.sqrt(kx ** 2 + ky ** 2), 1e-5)
    # w = sqrt(g*|k|) and the product w*t are fp32 in the shader; their fp32 rounding is part of the spec
    # because it moves the phase by up to ulp(w*t), far above fp64 noise.
    km32 = np.maximum(np.sqrt((kx.astype(np.float32) ** 2 + ky.astype(np.float32) ** 2).astype(np.float32)),
                      np.float32(1e-5)).astype(np.float32)
    w32 = np.sqrt((np.float32(G) * km32).astype(np.float32)).astype(np.float32)
    ph = (w32 * np.float32(t)).astype(np.float32).astype(np.float64)
    e = np.cos(ph) + 1j * np.sin(ph)
    hdy = h0k * e + h0minusk * np.conj(e)
    hdx = (-1j * kx / km) * hdy
    hdz = (-1j * ky / km) * hdy
    return hdy, hdx, hdz


def displacement(H):
    """What the 2*log2(N) butterfly passes + inversion compute (butterfly_cs.glsl, inversion_cs.glsl:29-36)."""
    return np.real(np.fft.ifft2(np.fft.ifftshift(H)))


def normal_map(h):
    """normal_map_cs.glsl:24-54 with LINEAR+REPEAT sampling at texel corners == 4-texel box means."""
    box = 0.25 * (h + np.roll(h, 1, 0) + np.roll(h, 1, 1) + np.roll(np.roll(h, 1, 0), 1, 1))

    def z(dx, dy):
        return np.roll(np.roll(box, -dy, 0), -dx, 1)

    z0, z1, z2 = z(-1, -1), z(0, -1), z(1, -1)
    z3, z4 = z(-1, 0), z(1, 0)
    z5, z6, z7 = z(-1, 1), z(0, 1), z(1, 1)
    nz = z0 + 2 * z1 + z2 - z5 - 2 * z6 - z7
    nx = z0 + 2 * z3 + z5 - z2 - 2 * z4 - z7
    ny = np.ones_like(nx)
    r = 1.0 / np.sqrt(nx * nx + ny * ny + nz * nz)
    return np.stack([nx * r, ny * r, nz * r, np.ones_like(nx)], axis=-1)


def jacobian(dx, dz, L, lam):
    """Extension: central differences with wrap, spacing L/N, consumer sign convention."""
    N = dx.shape[0]
    inv2h = N / (2.0 * L)
    dxdx = (np.roll(dx, -1, 1) - np.roll(dx, 1, 1)) * inv2h
    dxdz = (np.roll(dx, -1, 0) - np.roll(dx, 1, 0)) * inv2h
    dzdx = (np.roll(dz, -1, 1) - np.roll(dz, 1, 1)) * inv2h
    dzdz = (np.roll(dz, -1, 0) - np.roll(dz, 1, 0)) * inv2h
    return (1 - lam * dxdx) * (1 - lam * dzdz) - (lam * dxdz) * (lam * dzdx)


def jacobian_spectral(h0k, h0minusk, N, L, t, lam):
    """The Jacobian of the horizontal map X = x - lam*Dx, Z = z - lam*Dz (consumer convention, grid_tes.glsl:61-62) with EXACT
    spectral derivatives (SURVEY.md §8 f1's validation target): D(x) = Re sum_k H(k) e^{+i k.x} on x_n = n L / N, so
    d/dx D = Re(ifft2(ifftshift(i kx H))). Meaningful for spectra with no energy on the Nyquist row/column (index 0), whose
    derivative is not defined on the grid; the central-difference Jacobian converges to this one as (|k| L/N)^2 / 6."""
    kx, ky = wave_vectors(N, L)
    _, hdx, hdz = spectra(h0k, h0minusk, N, L, t)
    dxdx, dxdz = displacement(1j * kx * hdx), displacement(1j * ky * hdx)
    dzdx, dzdz = displacement(1j * kx * hdz), displacement(1j * ky * hdz)
    return (1 - lam * dxdx) * (1 - lam * dzdz) - (lam * dxdz) * (lam * dzdx)


def band_limited_h0(N, m, seed, amplitude=1.0):
    """A random initial spectrum with energy only on the (2m+1)^2 - 1 lowest non-zero wave vectors (|index - N/2| <= m): the same
    PHYSICAL waves at every N for a fixed L, which is what a grid-convergence check of the finite-difference Jacobian needs."""
    rng = np.random.default_rng(seed)
    a = np.zeros((N, N), np.complex128)
    b = np.zeros((N, N), np.complex128)
    c = N // 2
    # the chain divides by N^2 (inversion_cs.glsl:35): amplitudes in metres per mode need a factor N^2 in the spectrum
    blk = (rng.standard_normal((2, 2 * m + 1, 2 * m + 1)) + 1j * rng.standard_normal((2, 2 * m + 1, 2 * m + 1))) * (amplitude * N * N / (2 * m + 1))
    blk[:, m, m] = 0.0
    a[c - m:c + m + 1, c - m:c + m + 1] = blk[0]
    b[c - m:c + m + 1, c - m:c + m + 1] = blk[1]
    return a, b


def bilinear_repeat(img, u, v):
    """GL_LINEAR + GL_REPEAT fetch of img[row][col(, channels)] at normalised coordinates (u along columns, v along rows): texel
    centres at (i + 0.5)/N, taps floor(u N - 0.5) and the next, wrapped (what texture() does for the reference's sim textures,
    src/main.cpp:1142-1144 and the texture class defaults, SURVEY.md §8 b1, grid_tes.glsl:60-64)."""
    img = np.asarray(img, np.float64)
    n_r, n_c = img.shape[:2]
    tu = (np.asarray(u, np.float64) % 1.0) * n_c - 0.5
    tv = (np.asarray(v, np.float64) % 1.0) * n_r - 0.5
    fu, fv = np.floor(tu), np.floor(tv)
    a, b = tu - fu, tv - fv
    i0, j0 = fu.astype(np.int64) % n_c, fv.astype(np.int64) % n_r
    i1, j1 = (i0 + 1) % n_c, (j0 + 1) % n_r
    if img.ndim == 3:
        a, b = a[..., None], b[..., None]
    return (img[j0, i0] * (1 - a) + img[j0, i1] * a) * (1 - b) + (img[j1, i0] * (1 - a) + img[j1, i1] * a) * b


def blend_cascades(frames, Ls, choppiness, weights, displacement_scale, x, z):
    """the consumer's vertex displacement (grid_tes.glsl:60-64) summed over cascades with blending weights.
    frames[c] = dict(dy, dx, dz, normal) of cascade c (patch size Ls[c], sampled at uv = (x, z)/L). Returns (offset[...,4], normal[...,4])."""
    x, z = np.asarray(x, np.float64), np.asarray(z, np.float64)
    ox, oy, oz, ws, sx, sz = (np.zeros(x.shape) for _ in range(6))
    for f, L, lam, w in zip(frames, Ls, choppiness, weights):
        u, v = x / L, z / L
        oy += w * bilinear_repeat(f["dy"], u, v)
        ox -= w * lam * bilinear_repeat(f["dx"], u, v)
        oz -= w * lam * bilinear_repeat(f["dz"], u, v)
        n = bilinear_repeat(f["normal"], u, v)
        sx += w * n[..., 0] / n[..., 1]
        sz += w * n[..., 2] / n[..., 1]
        ws += w
    r = 1.0 / np.sqrt(sx * sx + 1.0 + sz * sz)
    return np.stack([ox, displacement_scale * oy, oz, ws], -1), np.stack([sx * r, r, sz * r, np.ones_like(r)], -1)


def frame_from_h0(h0k, h0minusk, N, L, t, choppiness=None):
    hdy, hdx, hdz = spectra(h0k, h0minusk, N, L, t)
    dy, dx, dz = displacement(hdy), displacement(hdx), displacement(hdz)
    out = dict(dy=dy, dx=dx, dz=dz, normal=normal_map(dy))
    if choppiness is not None:
        out["jacobian"] = jacobian(dx, dz, L, choppiness)
    return out


def philox4x32_10(ctr, key):
    """Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11), vectorised.
    ctr: (..., 4) uint32, key: (2,) uint32 -> (..., 4) uint32. Restates ow_init_kernels.cu:philox4x32_10 (config C5 noise)."""
    c = [np.asarray(ctr[..., i], np.uint64) for i in range(4)]
    k0, k1 = np.uint64(key[0]), np.uint64(key[1])
    M0, M1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(0x9E3779B9)) & MASK
        k1 = (k1 + np.uint64(0xBB67AE85)) & MASK
    return np.stack(c, -1).astype(np.uint32)


def philox_noise(seed: int, N: int) -> np.ndarray:
    """The four N x N noise byte planes ow_set_noise_seed / ow_slab_init_spectrum_seeded generate on the device:
    counter (ix, iy, 0, 0), key (seed lo, seed hi), plane j = low byte of output word j."""
    ctr = np.zeros((N, N, 4), np.uint32)
    ctr[..., 0] = np.arange(N, dtype=np.uint32)[None, :]
    ctr[..., 1] = np.arange(N, dtype=np.uint32)[:, None]
    out = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
    return np.ascontiguousarray(np.moveaxis(out & np.uint32(0xFF), -1, 0).astype(np.uint8))


def fold_pairs(h0k, h0minusk):
    """fp64 restatement of the init-time fold (csrc/ow_kernels.cuh: fold_pair / fold_pair_nyq; ow_init_kernels.cu: ow_fold_kernel).
    A = (h0k, h0minusk) at texel (u, v), B = the same at the mirror texel ((N-u) mod N, (N-v) mod N). Returns f[N][N][4] with
    S_y = H + conj(H_mirror) = (f0 c + f1 s, f2 s + f3 c), and g[N][N][4] with D = H - conj(H_mirror) = (g0 c + g1 s, g2 s + g3 c),
    c = cos(wt), s = sin(wt). The kernels keep f for rows 1..N/2-1 and g for their column 0 only."""
    N = h0k.shape[0]
    idx = (-np.arange(N)) % N
    Ax, Ay, Az, Aw = h0k.real, h0k.imag, h0minusk.real, h0minusk.imag
    Bk, Bm = h0k[idx][:, idx], h0minusk[idx][:, idx]
    Bx, By, Bz, Bw = Bk.real, Bk.imag, Bm.real, Bm.imag
    f = np.stack([Ax + Az + Bx + Bz, Aw + Bw - Ay - By, Ax - Az - Bx + Bz, Ay + Aw - By - Bw], -1)
    g = np.stack([Ax + Az - Bx - Bz, Aw - Ay - Bw + By, Ax - Az + Bx - Bz, Ay + Aw + By + Bw], -1)
    return f, g


def hermitian_parts_from_fold(f, g, N, L, t):
    """S_y, S_x, S_z (the spectra whose plain inverse DFT is TWICE the real displacement) from the folded coefficients,
    the way the row kernel evaluates them: S_x = -i kx/|k| S_y and S_z = -i ky/|k| S_y, except S_x on the Nyquist column
    (u = 0: the shader's k is not negated by mirroring there) which uses D, and likewise S_z on the Nyquist row (v = 0)."""
    kx, ky = wave_vectors(N, L)
    km = np.maximum(np.sqrt(kx ** 2 + ky ** 2), 1e-5)
    km32 = np.maximum(np.sqrt((kx.astype(np.float32) ** 2 + ky.astype(np.float32) ** 2).astype(np.float32)),
                      np.float32(1e-5)).astype(np.float32)
    w32 = np.sqrt((np.float32(G) * km32).astype(np.float32)).astype(np.float32)
    ph = (w32 * np.float32(t)).astype(np.float32).astype(np.float64)
    c, s = np.cos(ph), np.sin(ph)
    Sy = (f[..., 0] * c + f[..., 1] * s) + 1j * (f[..., 2] * s + f[..., 3] * c)
    D = (g[..., 0] * c + g[..., 1] * s) + 1j * (g[..., 2] * s + g[..., 3] * c)
    Sx = -1j * (kx / km) * Sy
    Sz = -1j * (ky / km) * Sy
    Sx[:, 0] = (-1j * (kx / km) * D)[:, 0]
    Sz[0, :] = (-1j * (ky / km) * D)[0, :]
    return Sy, Sx, Sz
