// ow_fft.cuh — register-resident inverse DFT butterflies and the 3-stage in-CTA FFT plan.
//
// Replaces the reference's twiddle/index table (twiddle_factors_cs.glsl:34-69, main.cpp:711-744) and its
// one-radix-2-stage-per-dispatch butterfly (butterfly_cs.glsl:54-132): here a whole length-N line is
// transformed inside one CTA in three radix-R stages (R in {2,4,8,16}) that live in registers, with two
// shared-memory exchanges in between. Sign convention is the reference's: X[k] = sum_n x[n] e^{+2 pi i nk/N}
// (unnormalised inverse transform, butterfly_cs.glsl:71 with the +sin twiddle of twiddle_factors_cs.glsl:38).
//
// Everything here is __host__ __device__ so tests/emu can run the identical index logic on the CPU.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define OW_HD __host__ __device__ __forceinline__

namespace ow {

OW_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
OW_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
OW_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
OW_HD float2 csqr(float2 a) { return make_float2(a.x * a.x - a.y * a.y, 2.0f * a.x * a.y); }
OW_HD float2 cmul_i(float2 a) { return make_float2(-a.y, a.x); }    // * (+i)
OW_HD float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

// e^{2 pi i * num / den}; sincospif keeps full accuracy without range reduction.
OW_HD float2 unit_root(int num, int den) {
    float s, c;
    sincospif(2.0f * (float)num / (float)den, &s, &c);
    return make_float2(c, s);
}

// ---------------------------------------------------------------------------------------------------
// In-register inverse DFTs. dft<R>(v): v[k] <- sum_n v[n] e^{+2 pi i nk/R}, natural order in and out.
// With full unrolling all index shuffles are register renames.
// ---------------------------------------------------------------------------------------------------
OW_HD void dft2(float2& a, float2& b) {
    const float2 t = a;
    a = cadd(t, b);
    b = csub(t, b);
}

OW_HD void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
    const float2 t0 = cadd(a0, a2), t1 = csub(a0, a2);
    const float2 t2 = cadd(a1, a3), t3 = cmul_i(csub(a1, a3));
    a0 = cadd(t0, t2);
    a1 = cadd(t1, t3);
    a2 = csub(t0, t2);
    a3 = csub(t1, t3);
}

template <int R>
struct Dft;

template <>
struct Dft<2> {
    static OW_HD void run(float2 (&v)[2]) { dft2(v[0], v[1]); }
};

template <>
struct Dft<4> {
    static OW_HD void run(float2 (&v)[4]) { dft4(v[0], v[1], v[2], v[3]); }
};

template <>
struct Dft<8> {
    // n = 2*n1 + n2 (n1<4, n2<2), k = k1 + 4*k2.
    static OW_HD void run(float2 (&v)[8]) {
        const float h = 0.70710678118654752440f;
        dft4(v[0], v[2], v[4], v[6]);   // y[k1][0] at v[2*k1]
        dft4(v[1], v[3], v[5], v[7]);   // y[k1][1] at v[2*k1+1]
        // y[k1][1] *= w8^{k1}
        v[3] = make_float2(h * (v[3].x - v[3].y), h * (v[3].x + v[3].y));      // * (1+i)/sqrt2
        v[5] = cmul_i(v[5]);                                                   // * i
        v[7] = make_float2(-h * (v[7].x + v[7].y), h * (v[7].x - v[7].y));     // * (-1+i)/sqrt2
        dft2(v[0], v[1]);   // X[0], X[4]
        dft2(v[2], v[3]);   // X[1], X[5]
        dft2(v[4], v[5]);   // X[2], X[6]
        dft2(v[6], v[7]);   // X[3], X[7]
        // currently X[k1 + 4*k2] sits at v[2*k1 + k2]; restore natural order.
        const float2 x0 = v[0], x4 = v[1], x1 = v[2], x5 = v[3], x2 = v[4], x6 = v[5], x3 = v[6], x7 = v[7];
        v[0] = x0; v[1] = x1; v[2] = x2; v[3] = x3; v[4] = x4; v[5] = x5; v[6] = x6; v[7] = x7;
    }
};

template <>
struct Dft<16> {
    // n = 4*n1 + n2, k = k1 + 4*k2 (all digits < 4).
    static OW_HD void run(float2 (&v)[16]) {
        const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
        dft4(v[0], v[4], v[8], v[12]);    // y[k1][n2] at v[4*k1 + n2]
        dft4(v[1], v[5], v[9], v[13]);
        dft4(v[2], v[6], v[10], v[14]);
        dft4(v[3], v[7], v[11], v[15]);
        // y[k1][n2] *= w16^{n2*k1}
        v[5]  = cmul(v[5], make_float2(c1, s1));                                   // e=1
        v[6]  = make_float2(h * (v[6].x - v[6].y), h * (v[6].x + v[6].y));          // e=2
        v[7]  = cmul(v[7], make_float2(s1, c1));                                   // e=3
        v[9]  = make_float2(h * (v[9].x - v[9].y), h * (v[9].x + v[9].y));          // e=2
        v[10] = cmul_i(v[10]);                                                     // e=4
        v[11] = make_float2(-h * (v[11].x + v[11].y), h * (v[11].x - v[11].y));     // e=6
        v[13] = cmul(v[13], make_float2(s1, c1));                                  // e=3
        v[14] = make_float2(-h * (v[14].x + v[14].y), h * (v[14].x - v[14].y));     // e=6
        v[15] = cmul(v[15], make_float2(-c1, -s1));                                // e=9
        dft4(v[0], v[1], v[2], v[3]);     // X[0 + 4*k2] at v[k2]
        dft4(v[4], v[5], v[6], v[7]);     // X[1 + 4*k2] at v[4 + k2]
        dft4(v[8], v[9], v[10], v[11]);
        dft4(v[12], v[13], v[14], v[15]);
        // X[k1 + 4*k2] sits at v[4*k1 + k2]: transpose the 4x4 register tile.
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = a + 1; b < 4; ++b) {
                const float2 t = v[4 * a + b];
                v[4 * a + b] = v[4 * b + a];
                v[4 * b + a] = t;
            }
    }
};

// e^{2 pi i k/R} as compile-time constants (R in {2,4,8,16}); get(k) folds to immediates when k is unrolled.
template <int R>
struct RootsOfUnity {
    static OW_HD float2 get(int k) {
        // quarter-wave table in sixteenths of a turn: cos(2 pi m/16), m = 0..4
        const float c16[5] = {1.0f, 0.92387953251128675613f, 0.70710678118654752440f, 0.38268343236508977173f, 0.0f};
        const int m = (k * (16 / R)) & 15;           // angle in sixteenths
        const int q = m >> 2, r = m & 3;             // quadrant, remainder
        const float c = c16[r], s = c16[4 - r];      // cos, sin of the remainder angle
        switch (q) {
            case 0: return make_float2(c, s);
            case 1: return make_float2(-s, c);
            case 2: return make_float2(-c, -s);
            default: return make_float2(s, -c);
        }
    }
};

// tw[k] = w^k for k in [0,R) with multiplication depth <= 4 (keeps the error at a few ulp).
template <int R>
OW_HD void twiddle_powers(float2 w, float2 (&tw)[R]) {
    tw[0] = make_float2(1.0f, 0.0f);
    if (R > 1) tw[1] = w;
    if (R > 2) tw[2] = csqr(w);
    if (R > 3) tw[3] = cmul(tw[2], w);
    if (R > 4) {
        tw[4] = csqr(tw[2]);
#pragma unroll
        for (int k = 5; k < 8 && k < R; ++k) tw[k] = cmul(tw[4], tw[k - 4]);
    }
    if (R > 8) {
        tw[8] = csqr(tw[4]);
#pragma unroll
        for (int k = 9; k < 16 && k < R; ++k) tw[k] = cmul(tw[8], tw[k - 8]);
    }
}

// ---------------------------------------------------------------------------------------------------
// 3-stage plan for one length-N line.  N = R0*R1*R2, digits d0<R0, d1<R1, d2<R2.
//   input index   n = d0*(R1*R2) + d1*R2 + d2      (d0 transformed first)
//   output index  k = k0 + R0*k1 + R0*R1*k2
// The line lives in shared memory IN PLACE: digit slot s holds n_s before stage s and k_s after it, so a
// butterfly reads and writes the same R_s addresses and one barrier per exchange is enough.
//   address(d0,d1,d2) = d0*S0 + d1*S1 + d2      (float2 elements; S1,S0 padded against bank conflicts)
// Stage butterflies and their ids (the id is what consecutive threads enumerate):
//   stage 0: id b  = d1*R2 + d2            loops d0   twiddle after: e^{2 pi i k0*b/N}
//   stage 1: id q  = d2 + R2*k0            loops d1   twiddle after: e^{2 pi i k1*d2/(R1*R2)}
//   stage 2: id b' = k0 + R0*k1            loops d2   output k = b' + R0*R1*k2
// T threads serve one line; a stage with fewer butterflies than threads leaves the upper threads idle.
// ---------------------------------------------------------------------------------------------------
template <int N_, int R0_, int R1_, int R2_, int T_, int P1_, int P0_>
struct Plan {
    static constexpr int N = N_, R0 = R0_, R1 = R1_, R2 = R2_, T = T_;
    static constexpr int M = R1 * R2;                 // stage-0 butterflies
    static constexpr int B1 = R0 * R2;                // stage-1 butterflies
    static constexpr int B2 = R0 * R1;                // stage-2 butterflies
    static constexpr int C0 = (M + T - 1) / T, C1 = (B1 + T - 1) / T, C2 = (B2 + T - 1) / T;
    static constexpr int S1 = R2 + P1_;
    static constexpr int S0 = R1 * S1 + P0_;
    static constexpr int LINE = R0 * S0;              // float2 elements per line
    static_assert(R0 * R1 * R2 == N, "radices must multiply to N");
    static OW_HD int addr(int d0, int d1, int d2) { return d0 * S0 + d1 * S1 + d2; }
};

// Shared-memory access shim: on the device a plain pointer; the CPU emulator swaps in a recorder that
// also tallies bank conflicts.
struct SmemDirect {
    float2* p;
    OW_HD float2 ld(int i) const { return p[i]; }
    OW_HD void st(int i, float2 v) const { p[i] = v; }
};

// Stage 1 for butterfly id q of one line (base = line offset inside the accessor). tw = powers of
// e^{2 pi i d2/(R1*R2)} (stage1_twiddles), shared by every line the thread transforms at this q.
template <class P>
OW_HD void stage1_twiddles(int q, float2 (&tw)[P::R1]) {
    twiddle_powers<P::R1>(unit_root(q % P::R2, P::R1 * P::R2), tw);
}

template <class P, class Smem>
OW_HD void stage1(const Smem& sm, int base, int q, const float2 (&tw)[P::R1]) {
    const int d2 = q % P::R2, k0 = q / P::R2;
    float2 v[P::R1];
#pragma unroll
    for (int d1 = 0; d1 < P::R1; ++d1) v[d1] = sm.ld(base + P::addr(k0, d1, d2));
    Dft<P::R1>::run(v);
#pragma unroll
    for (int k1 = 1; k1 < P::R1; ++k1) v[k1] = cmul(v[k1], tw[k1]);
#pragma unroll
    for (int k1 = 0; k1 < P::R1; ++k1) sm.st(base + P::addr(k0, k1, d2), v[k1]);
}

// Stage 2 load + transform for butterfly id b'; results stay in v[k2] for output index k = b' + R0*R1*k2.
template <class P, class Smem>
OW_HD void stage2(const Smem& sm, int base, int bp, float2 (&v)[P::R2]) {
    const int k0 = bp % P::R0, k1 = bp / P::R0;
#pragma unroll
    for (int d2 = 0; d2 < P::R2; ++d2) v[d2] = sm.ld(base + P::addr(k0, k1, d2));
    Dft<P::R2>::run(v);
}

// Finish stage 0 for butterfly id b: transform, twiddle by e^{2 pi i k0*b/N}, store to the line.
template <class P, class Smem>
OW_HD void stage0_finish(const Smem& sm, int base, int b, float2 (&v)[P::R0], const float2 (&tw)[P::R0]) {
    Dft<P::R0>::run(v);
    const int d1 = b / P::R2, d2 = b % P::R2;
#pragma unroll
    for (int k0 = 0; k0 < P::R0; ++k0) {
        const float2 o = (k0 == 0) ? v[0] : cmul(v[k0], tw[k0]);
        sm.st(base + P::addr(k0, d1, d2), o);
    }
}

}  // namespace ow
