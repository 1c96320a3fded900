// ow_kernels.cuh — bodies of the per-frame kernels, written as per-thread "phase" functions.
//
// A kernel is  phase0 ; barrier ; phase1 ; barrier ; phase2  where every phase is a __host__ __device__
// function of (thread id, block coordinates, shared-memory accessor). ow_frame_kernels.cu wraps them in
// __global__ functions; tests/emu runs the same phases thread-by-thread on the CPU to check the index
// algebra and count shared-memory bank conflicts (the emulator is test infrastructure, never a fallback).
//
// What the frame computes (reference: src/main.cpp:240-244 and the shaders it dispatches):
//   tilde_h0_t_cs.glsl:70-131   h(k,t) for dy/dx/dz from h0k, h0minusk     -> fused into row phase 0
//   butterfly_cs.glsl:54-132    log2N horizontal + log2N vertical passes   -> row kernel + column kernel
//   inversion_cs.glsl:25-43     (-1)^(x+y) * Re(h) / N^2                    -> column kernel phase 2
//   normal_map_cs.glsl:24-54    Sobel of the box-filtered height           -> normal kernel
//
// Only Re(.) of each inverse transform is kept by the reference, so instead of three full complex 2-D
// transforms we transform the Hermitian part of each spectrum: S(k) = H(k) + conj(H(-k)) (the factor 1/2 is
// folded into the final scale). Row p of S and row N-p are conjugates after the row transform, so only
// rows p in [0, N/2) are produced (rows 0 and N/2 are both real and share row 0 as re/im), and the column
// pass rebuilds two real output columns from one complex transform.  1.5 complex FFTs per line instead of 3,
// and a 12 B/texel intermediate instead of 24.
#pragma once
#include "ow_fft.cuh"

namespace ow {

#ifdef __CUDA_ARCH__
// IEEE, never-fused fp32 ops where bit-parity of the phase w*t with the reference matters.
#define OW_MUL(a, b) __fmul_rn((a), (b))
#define OW_ADD(a, b) __fadd_rn((a), (b))
#define OW_SQRT(a) __fsqrt_rn((a))
#define OW_RCP(a) __fdividef(1.0f, (a))
#define OW_LDG(p) __ldg(p)
#else
#define OW_MUL(a, b) ((a) * (b))
#define OW_ADD(a, b) ((a) + (b))
#define OW_SQRT(a) sqrtf((a))
#define OW_RCP(a) (1.0f / (a))
#define OW_LDG(p) (*(p))
#endif

constexpr float kGravity = 9.81f;   // tilde_h0_t_cs.glsl:58

// ---------------------------------------------------------------------------------------------------
// Spectrum at one texel pair: (u,v) and its mirror (mu,mv) = (-k).  Returns the Hermitian parts
//   Sy = Hdy(u,v) + conj(Hdy(mu,mv)),  Sx, Sz likewise with the texels' OWN choppy multipliers
// (literal at the Nyquist row/column where the mirror texel is the texel itself).
// h0 is stored interleaved: float4(h0k.re, h0k.im, h0minusk.re, h0minusk.im).
// ---------------------------------------------------------------------------------------------------
struct Sym3 {
    float2 y, x, z;
};

// sin/cos of the phase w*t. FAST: two-constant Cody-Waite reduction by 2*pi + the SFU's sin/cos (absolute
// error ~5e-7 for |x| < 2e4; the launcher only selects it when max |w*t| of the cascade is below that);
// otherwise libdevice sincosf (full-range Payne-Hanek).
template <bool FAST>
OW_HD void phase_sincos(float x, float* s, float* c) {
#ifdef __CUDA_ARCH__
    if (FAST) {
        const float n = rintf(x * 0.15915494309189535f);
        float r = fmaf(n, -6.2831854820251465f, x);
        r = fmaf(n, 1.7484555314695172e-7f, r);
        *s = __sinf(r);
        *c = __cosf(r);
        return;
    }
#endif
    sincosf(x, s, c);
}

// Raw inputs of one texel pair; loading is split from the arithmetic so a thread can put all the loads of a
// butterfly in flight before the first use (the sincos code otherwise serialises them: one DRAM round trip each).
struct TexelPair {
    float4 A, B;   // h0 at (u, v) and at the mirror texel (N-u, N-v)
    float kx;      // k_x at u
};

template <int N>
OW_HD TexelPair load_pair(const float4* __restrict__ h0, const float* __restrict__ ktab, int u, int v, int mv) {
    const int mu = (N - u) & (N - 1);
    TexelPair tp;
    tp.A = OW_LDG(h0 + (size_t)v * N + u);
    tp.B = OW_LDG(h0 + (size_t)mv * N + mu);
    tp.kx = OW_LDG(ktab + u);
    return tp;
}

template <bool FAST>
OW_HD Sym3 spectrum_sym(const TexelPair& tp, int u, float ky, bool self_row, float t) {
    const float4 A = tp.A, B = tp.B;
    const float kx = tp.kx;
    // k at the mirror texel is -k, except on the Nyquist row/column (index 0) whose mirror is itself
    const float kxm = (u == 0) ? kx : -kx, kym = self_row ? ky : -ky;
    // tilde_h0_t_cs.glsl:74-79 — same operation order as the shader so w*t matches to the bit.
    float km = OW_SQRT(OW_ADD(OW_MUL(kx, kx), OW_MUL(ky, ky)));
    if (km < 0.00001f) km = 0.00001f;
    const float w = OW_SQRT(OW_MUL(kGravity, km));
    float s, c;
    phase_sincos<FAST>(OW_MUL(w, t), &s, &c);                       // :96-97
    // :110  h = h0k * e^{iwt} + h0minusk * e^{-iwt}   (conjugate() is a no-op in the shader, :42-48)
    const float2 H  = make_float2((A.x * c - A.y * s) + (A.z * c + A.w * s), (A.x * s + A.y * c) + (A.w * c - A.z * s));
    const float2 Hm = make_float2((B.x * c - B.y * s) + (B.z * c + B.w * s), (B.x * s + B.y * c) + (B.w * c - B.z * s));
    const float ik = OW_RCP(km);
    const float rx = kx * ik, rz = ky * ik, rxm = kxm * ik, rzm = kym * ik;   // k/|k| at the texel and at its mirror
    Sym3 o;
    o.y = make_float2(H.x + Hm.x, H.y - Hm.y);
    // :113-126  (0, -kx/|k|) * h = (kx/|k| * h.im, -kx/|k| * h.re); add the conjugate of the mirror's.
    o.x = make_float2(rx * H.y + rxm * Hm.y, rxm * Hm.x - rx * H.x);
    o.z = make_float2(rz * H.y + rzm * Hm.y, rzm * Hm.x - rz * H.x);
    return o;
}

// ---------------------------------------------------------------------------------------------------
// ROW KERNEL.  One group of P::T threads transforms row pair p (rows p and N-p; p==0: rows 0 and N/2) for
// the three channels. Shared memory per group: 3 lines of P::LINE float2 (dy, dx, dz).
// Output: inter[c][p][x] (float2), c in {dy,dx,dz}, p < N/2, x < N — the row transform of S_c(., p).
// ---------------------------------------------------------------------------------------------------
template <class P, bool FAST, class Smem>
OW_HD void row_phase0(const Smem& sm, int ft, int p, const float4* __restrict__ h0, const float* __restrict__ ktab,
                      float t) {
    constexpr int N = P::N, R0 = P::R0;
#pragma unroll 1
    for (int c = 0; c < P::C0; ++c) {
        const int b = ft + P::T * c;
        if (b >= P::M) break;
        float2 vy[R0], vx[R0], vz[R0];
        TexelPair tp[R0];
        if (p != 0) {
            const float ky = OW_LDG(ktab + p);
#pragma unroll
            for (int d0 = 0; d0 < R0; ++d0) tp[d0] = load_pair<N>(h0, ktab, d0 * P::M + b, p, N - p);
#pragma unroll
            for (int d0 = 0; d0 < R0; ++d0) {
                const Sym3 s = spectrum_sym<FAST>(tp[d0], d0 * P::M + b, ky, false, t);
                vy[d0] = s.y; vx[d0] = s.x; vz[d0] = s.z;
            }
        } else {
            // rows 0 (Nyquist) and N/2 (DC) mirror onto themselves; both row transforms are real, so they
            // travel as one complex line: Z = S(.,0) + i*S(.,N/2).
            const float ky0 = OW_LDG(ktab), kyh = OW_LDG(ktab + N / 2);
#pragma unroll
            for (int d0 = 0; d0 < R0; ++d0) tp[d0] = load_pair<N>(h0, ktab, d0 * P::M + b, 0, 0);
#pragma unroll
            for (int d0 = 0; d0 < R0; ++d0) {
                const Sym3 a = spectrum_sym<FAST>(tp[d0], d0 * P::M + b, ky0, true, t);
                vy[d0] = a.y; vx[d0] = a.x; vz[d0] = a.z;
            }
#pragma unroll
            for (int d0 = 0; d0 < R0; ++d0) tp[d0] = load_pair<N>(h0, ktab, d0 * P::M + b, N / 2, N / 2);
#pragma unroll
            for (int d0 = 0; d0 < R0; ++d0) {
                const Sym3 q = spectrum_sym<FAST>(tp[d0], d0 * P::M + b, kyh, true, t);
                vy[d0] = make_float2(vy[d0].x - q.y.y, vy[d0].y + q.y.x);
                vx[d0] = make_float2(vx[d0].x - q.x.y, vx[d0].y + q.x.x);
                vz[d0] = make_float2(vz[d0].x - q.z.y, vz[d0].y + q.z.x);
            }
        }
        float2 tw[R0];
        twiddle_powers<R0>(unit_root(b, N), tw);
        stage0_finish<P>(sm, 0 * P::LINE, b, vy, tw);
        stage0_finish<P>(sm, 1 * P::LINE, b, vx, tw);
        stage0_finish<P>(sm, 2 * P::LINE, b, vz, tw);
    }
}

template <class P, class Smem>
OW_HD void row_phase1(const Smem& sm, int ft) {
#pragma unroll 1
    for (int c = 0; c < P::C1; ++c) {
        const int q = ft + P::T * c;
        if (q >= P::B1) break;
        float2 tw[P::R1];
        stage1_twiddles<P>(q, tw);
#pragma unroll 1
        for (int f = 0; f < 3; ++f) stage1<P>(sm, f * P::LINE, q, tw);
    }
}

template <class P, class Smem>
OW_HD void row_phase2(const Smem& sm, int ft, int p, float2* __restrict__ inter) {
    constexpr int N = P::N;
#pragma unroll 1
    for (int c = 0; c < P::C2; ++c) {
        const int bp = ft + P::T * c;
        if (bp >= P::B2) break;
#pragma unroll 1
        for (int f = 0; f < 3; ++f) {
            float2 v[P::R2];
            stage2<P>(sm, f * P::LINE, bp, v);
            float2* dst = inter + ((size_t)f * (N / 2) + p) * N + bp;
#pragma unroll
            for (int k2 = 0; k2 < P::R2; ++k2) dst[k2 * P::B2] = v[k2];
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// COLUMN KERNEL.  A CTA owns G "jobs" of one channel; job g = columns (x, x+1), x = 2*(tile*G + g).
// One complex length-N transform per job yields both real columns:
//   Q_v = I[v][x] + i*I[v][x+1]                  0 < v < N/2
//   Q_v = conj(I[N-v][x]) + i*conj(I[N-v][x+1])  N/2 < v < N      (row N-v of S is the conjugate of row v)
//   Q_0 = Re I[0][x] + i*Re I[0][x+1],  Q_{N/2} = Im I[0][x] + i*Im I[0][x+1]   (packed real rows)
// then D[y][x] = Re X_y, D[y][x+1] = Im X_y, times (-1)^(x+y) * 0.5/N^2 (inversion_cs.glsl:29-36).
// Threads: tid = g + G*ft (job fastest, so a warp's global accesses cover G adjacent column pairs).
// Stage 0 pairs butterflies j and M-j in one thread: each loaded row feeds Q_v of one and Q_{N-v} of the
// other, so every intermediate element is loaded exactly once (pair ids j in [0, M/2)).
// ---------------------------------------------------------------------------------------------------
OW_HD float2 pack_fwd(float4 r) { return make_float2(r.x - r.w, r.y + r.z); }   // P1 + i*P2
OW_HD float2 pack_cnj(float4 r) { return make_float2(r.x + r.w, r.z - r.y); }   // conj(P1) + i*conj(P2)

template <class P, class Smem>
OW_HD void col_phase0(const Smem& sm, int base, int j /* pair id in [0, M/2) */, const float2* __restrict__ src /* inter[c] + x */) {
    constexpr int N = P::N, R0 = P::R0, M = P::M, H = R0 / 2;
    const int bA = j, bB = (j == 0) ? M / 2 : M - j;
    float2 qa[R0], qb[R0];
    float4 la[H], lb[H];
#pragma unroll
    for (int i = 0; i < H; ++i) {
        la[i] = OW_LDG(reinterpret_cast<const float4*>(src + (size_t)(i * M + bA) * N));
        lb[i] = OW_LDG(reinterpret_cast<const float4*>(src + (size_t)(i * M + bB) * N));
    }
    if (j != 0) {
#pragma unroll
        for (int i = 0; i < H; ++i) {
            qa[i] = pack_fwd(la[i]);          qb[R0 - 1 - i] = pack_cnj(la[i]);   // v = i*M+j ; N-v = (R0-1-i)*M + (M-j)
            qb[i] = pack_fwd(lb[i]);          qa[R0 - 1 - i] = pack_cnj(lb[i]);
        }
    } else {
        qa[0] = make_float2(la[0].x, la[0].z);    // v = 0   : real parts of packed row 0
        qa[H] = make_float2(la[0].y, la[0].w);    // v = N/2 : imaginary parts of packed row 0
#pragma unroll
        for (int i = 1; i < H; ++i) { qa[i] = pack_fwd(la[i]); qa[R0 - i] = pack_cnj(la[i]); }       // v = i*M ; N-v = (R0-i)*M
#pragma unroll
        for (int i = 0; i < H; ++i) { qb[i] = pack_fwd(lb[i]); qb[R0 - 1 - i] = pack_cnj(lb[i]); }   // v = i*M+M/2
    }
    float2 tw[R0];
    twiddle_powers<R0>(unit_root(bA, N), tw);
    stage0_finish<P>(sm, base, bA, qa, tw);
    twiddle_powers<R0>(unit_root(bB, N), tw);
    stage0_finish<P>(sm, base, bB, qb, tw);
}

template <class P, class Smem>
OW_HD void col_phase1(const Smem& sm, int base, int ft) {
#pragma unroll 1
    for (int c = 0; c < P::C1; ++c) {
        const int q = ft + P::T * c;
        if (q >= P::B1) break;
        float2 tw[P::R1];
        stage1_twiddles<P>(q, tw);
        stage1<P>(sm, base, q, tw);
    }
}

template <class P, class Smem>
OW_HD void col_phase2(const Smem& sm, int base, int ft, float* __restrict__ dst /* out[c] + x */, float scale) {
    constexpr int N = P::N;
#pragma unroll 1
    for (int c = 0; c < P::C2; ++c) {
        const int bp = ft + P::T * c;
        if (bp >= P::B2) break;
        float2 v[P::R2];
        stage2<P>(sm, base, bp, v);
        // y = bp + B2*k2 has the parity of bp (B2 is even); x is even: sign(x,y) = (-1)^y, sign(x+1,y) = -(-1)^y.
        const float sg = (bp & 1) ? -scale : scale;
#pragma unroll
        for (int k2 = 0; k2 < P::R2; ++k2) {
            const int y = bp + P::B2 * k2;
            *reinterpret_cast<float2*>(dst + (size_t)y * N) = make_float2(sg * v[k2].x, -sg * v[k2].y);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// NORMAL (+ JACOBIAN).  normal_map_cs.glsl:24-54: the eight texture() taps sit on texel corners, so with
// LINEAR+REPEAT each tap is the mean of a 2x2 block ("box"); the stencil covers columns x-2..x+1 and rows
// y-2..y+1 with wrap-around. One thread owns column x and walks down RY output rows keeping a sliding window:
//   hs(r)[c]  = h[r][x+c-2] + h[r][x+c-1]          c = 0,1,2   (horizontal pair sums, 4 loads per row)
//   box(r)[c] = (hs(r-1)[c] + hs(r)[c]) / 4        = taps at (x-1, x, x+1) of tap-row r
//   sx(r)  = box[0] + 2 box[1] + box[2],   dxb(r) = box[0] - box[2]
//   n.z(j) = sx(j-1) - sx(j+1)                     (:49)     n.x(j) = dxb(j-1) + 2 dxb(j) + dxb(j+1)   (:50)
// so each texel costs 4 loads instead of 16. Lanes run along x: every load and the float4 store are coalesced.
// The Jacobian (extension) rides the same walk:
//   J = (1 - l*dDx/dx)(1 - l*dDz/dz) - l^2 (dDx/dz)(dDz/dx), central differences with wrap, spacing L/N.
// ---------------------------------------------------------------------------------------------------
template <int N, int RY, bool JAC>
OW_HD void normal_column_walk(const float* __restrict__ disp /* dy,dx,dz planes */, float4* __restrict__ normal,
                              float* __restrict__ jac, int x, int y0, float lambda, float inv2h) {
    constexpr int MSK = N - 1;
    const int c0 = (x - 2) & MSK, c1 = (x - 1) & MSK, c3 = (x + 1) & MSK;
    const float* hy = disp;
    const float* hx = disp + (size_t)N * N;
    const float* hz = disp + (size_t)2 * N * N;
    float hs_prev[3], sx_m1 = 0.f, sx_0 = 0.f, dxb_m1 = 0.f, dxb_0 = 0.f;   // window state
    float xc_m1 = 0.f, xc_0 = 0.f, zc_m1 = 0.f, zc_0 = 0.f, ddx_0 = 0.f, ddz_0 = 0.f;
    {
        const float* r = hy + (size_t)((y0 - 2) & MSK) * N;
        const float a = OW_LDG(r + c0), b = OW_LDG(r + c1), c = OW_LDG(r + x), d = OW_LDG(r + c3);
        hs_prev[0] = a + b; hs_prev[1] = b + c; hs_prev[2] = c + d;
    }
#pragma unroll
    for (int i = -1; i <= RY; ++i) {              // tap-row r = y0 + i ; emits output row r - 1 once i >= 1
        const int rr = (y0 + i) & MSK;
        const float* r = hy + (size_t)rr * N;
        const float a = OW_LDG(r + c0), b = OW_LDG(r + c1), c = OW_LDG(r + x), d = OW_LDG(r + c3);
        const float h0 = a + b, h1 = b + c, h2 = c + d;
        const float b0 = (hs_prev[0] + h0) * 0.25f, b1 = (hs_prev[1] + h1) * 0.25f, b2 = (hs_prev[2] + h2) * 0.25f;
        hs_prev[0] = h0; hs_prev[1] = h1; hs_prev[2] = h2;
        const float sx = b0 + 2.0f * b1 + b2, dxb = b0 - b2;
        float xc = 0.f, zc = 0.f, ddx = 0.f, ddz = 0.f;
        if (JAC) {
            const float* rx = hx + (size_t)rr * N;
            const float* rz = hz + (size_t)rr * N;
            xc = OW_LDG(rx + x); zc = OW_LDG(rz + x);
            if (i >= 0 && i < RY) {
                ddx = OW_LDG(rx + c3) - OW_LDG(rx + c1);      // dDx/dx * 2h at row r
                ddz = OW_LDG(rz + c3) - OW_LDG(rz + c1);      // dDz/dx * 2h at row r
            }
        }
        if (i >= 1) {
            const int j = y0 + i - 1;             // output row: taps j-1 (.._m1), j (.._0), j+1 (current)
            const float nz = sx_m1 - sx, nx = dxb_m1 + 2.0f * dxb_0 + dxb;
            const float rinv = rsqrtf(nx * nx + 1.0f + nz * nz);
            normal[(size_t)j * N + x] = make_float4(nx * rinv, rinv, nz * rinv, 1.0f);   // :53
            if (JAC) {
                const float dxdx = ddx_0 * inv2h, dzdx = ddz_0 * inv2h;
                const float dxdz = (xc - xc_m1) * inv2h, dzdz = (zc - zc_m1) * inv2h;
                jac[(size_t)j * N + x] = (1.0f - lambda * dxdx) * (1.0f - lambda * dzdz) - (lambda * dxdz) * (lambda * dzdx);
            }
        }
        sx_m1 = sx_0; sx_0 = sx; dxb_m1 = dxb_0; dxb_0 = dxb;
        xc_m1 = xc_0; xc_0 = xc; zc_m1 = zc_0; zc_0 = zc; ddx_0 = ddx; ddz_0 = ddz;
    }
}

}  // namespace ow
