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
// Loads of data PRODUCED on the device (the row->column intermediate, the displacement planes). Between separate kernels the
// read-only path is fine; the frame-pipelined kernel (ow_mega_kernels.cu) reads what other CTAs of the SAME launch wrote, so there
// they must come from L2 (ld.global.cg), never from the non-coherent L1 path.
#ifdef OW_COHERENT_LOADS
#define OW_LDP(p) __ldcg(p)
#else
#define OW_LDP(p) __ldg(p)
#endif
// Drop a 128-byte line from L2 WITHOUT writing it back (its contents become undefined). Used on the row->column
// intermediate once the column kernel has consumed it: nobody reads it again before the next frame's row kernel
// rewrites the whole line, so the dirty data never has to travel to DRAM.
#define OW_DISCARD_L2(p) asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory")
// A data dependency on two loaded values, ordered (volatile) before the volatile asm statements that follow it.
#define OW_USE_BEFORE(a, b) asm volatile("" ::"f"(a), "f"(b) : "memory")
#else
#define OW_DISCARD_L2(p) ((void)(p))
#define OW_USE_BEFORE(a, b) ((void)(a), (void)(b))
#define OW_MUL(a, b) ((a) * (b))
#define OW_ADD(a, b) ((a) + (b))
#define OW_SQRT(a) sqrtf((a))
#define OW_RCP(a) (1.0f / (a))
#define OW_LDG(p) (*(p))
#define OW_LDP(p) (*(p))
#endif

constexpr float kGravity = 9.81f;   // tilde_h0_t_cs.glsl:58

// Ablation hooks for tools/tune only (never defined in the library build): bit 0 = no h0/ktab loads, bit 1 = no
// dispersion/sincos math, bit 2 = skip FFT stages 1 and 2 of the row kernel, bit 3 = row kernel stores suppressed.
#ifndef OW_ABLATE
#define OW_ABLATE 0
#endif

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

// rowA = base of h0 row v, rowB = base of the mirror row (N - v) mod N (see the row maps below).
template <int N>
OW_HD TexelPair load_pair(const float4* __restrict__ rowA, const float4* __restrict__ rowB, const float* __restrict__ ktab, int u) {
    const int mu = (N - u) & (N - 1);
    TexelPair tp;
#if OW_ABLATE & 1
    tp.A = make_float4(u * 1e-3f, 1e-3f, mu * 1e-3f, 1.0f); tp.B = make_float4(2e-3f, 0.5f, u * 2e-3f, 0.25f); tp.kx = (u - N / 2) * 6.28e-3f;
    return tp;
#endif
    tp.A = OW_LDG(rowA + u);
    tp.B = OW_LDG(rowB + mu);
    tp.kx = OW_LDG(ktab + u);
    return tp;
}

// ---------------------------------------------------------------------------------------------------
// Where the rows of h0 live and where the row-transformed lines go.
//   FullRows/FullSink : one GPU owns the whole grid: h0[v][u], inter[c][p][x]  (FrameBuffers layout).
//   SlabRows/SlabSink : slab decomposition of ONE grid over `world` GPUs. Rank r owns the row
//     pairs p in [p0, p0 + PL), PL = N/2/world, i.e. rows p and N-p (pair 0: rows 0 and N/2), stored locally as
//     h0_loc[2*PL][N]: row v < N/2 at index v - p0, row v >= N/2 at index PL + ((N-v) mod N/2) - p0. The sink is the
//     transpose: column x belongs to rank h = x / XL (XL = N/world); rank h's receive buffer is laid out
//     [p][c][XH] with XH = XL + 2*kHalo columns, column x at index kHalo + (x mod XL); the kHalo columns either
//     side are copies of the neighbours' edge columns (wrap-around), so the normal/Jacobian stencil of a column
//     slab needs no second exchange. base[h] points at THIS rank's [PL][3][XH] block inside rank h's buffer (a peer
//     mapping for direct NVLink stores, or a slice of the local send buffer for the NCCL all-to-all).
// ---------------------------------------------------------------------------------------------------
constexpr int kHalo = 8;          // halo columns either side of a column slab (stencil needs 2 left, 1 right; 8 keeps
                                  // 16-column CTA tiles and float4 alignment)
constexpr int kMaxWorld = 8;

// A block of folded pair rows is [npairs][N] float4 (fold_pair), followed - for N <= 512 - by [npairs][N] float2 (w, 1/|k|): the
// dispersion w = sqrt(g |k|) and 1/|k| of the pair's texel depend on the grid only, so they are tabulated at init (ow_fold_kernel, with
// the shader's operation order) instead of being re-derived through two IEEE square roots per texel per frame. Measured on B200: +4 %
// frames/s at N = 512 (a time sweep keeps the cascade's block L2-resident); at N = 1024 neutral in a 64-cascade step and -4 % in an
// 8-cascade one (every block is read once per step, from DRAM), at N = 2048 slower: the table's extra 4 B/texel come from DRAM every
// frame and cost more than they save (the frame is DRAM-bound there). So grids above 512 have no table and their kernels
// re-derive (w, 1/|k|) in registers: w with the same operations (bit-identical), 1/|k| through the fast reciprocal (1 ulp).
OW_HD bool use_wk(int N) { return N <= 512; }
template <int N>
constexpr bool kUseWk = (N <= 512);
OW_HD size_t hp_block_f4(int npairs, int N) { return (size_t)npairs * N * (use_wk(N) ? 3 : 2) / 2; }    // float4 elements of a block

template <int N>
struct FullRows {
    const float4* h0;     // [N][N]    full initial spectrum (pair 0 and downloads)
    const float4* hp;     // [N/2][N]  folded pairs, row p (row 0 unused), then [N/2][N] float2 (w, 1/|k|)
    const float4* nyq;    // [N/2]     Nyquist-column extras
    OW_HD const float4* row(int v) const { return h0 + (size_t)v * N; }
    OW_HD const float4* pair_row(int p) const { return hp + (size_t)p * N; }
    OW_HD const float2* wk_row(int p) const { return reinterpret_cast<const float2*>(hp + (size_t)(N / 2) * N) + (size_t)p * N; }
    OW_HD const float4* nyq_of(int p) const { return nyq + p; }
};

template <int N>
struct FullSink {
    float2* inter;
    OW_HD void put(int c, int p, int x, float2 v) const { inter[((size_t)c * (N / 2) + p) * N + x] = v; }
};

template <int N>
struct SlabRows {
    const float4* h0;   // [2*PL][N]
    const float4* hp;   // [PL][N]   folded pairs of this rank (local pair index p - p0), then [PL][N] float2 (w, 1/|k|)
    const float4* nyq;  // [PL]
    int p0, PL;
    OW_HD const float2* wk_row(int p) const { return reinterpret_cast<const float2*>(hp + (size_t)PL * N) + (size_t)(p - p0) * N; }
    OW_HD int local(int v) const { return (v < N / 2 ? v : PL + ((N - v) & (N / 2 - 1))) - p0; }
    OW_HD const float4* row(int v) const { return h0 + (size_t)local(v) * N; }
    OW_HD const float4* pair_row(int p) const { return hp + (size_t)(p - p0) * N; }
    OW_HD const float4* nyq_of(int p) const { return nyq + (p - p0); }
};

template <int N>
struct SlabSink {
    float2* base[kMaxWorld];
    int world, p0, XL, XH, xl_shift;      // XL = 1 << xl_shift
    OW_HD void put(int c, int p, int x, float2 v) const {
        const int h = x >> xl_shift, xl = x & (XL - 1);
        const size_t line = ((size_t)(p - p0) * 3 + c) * XH;
        base[h][line + kHalo + xl] = v;
        if (xl < kHalo) base[h == 0 ? world - 1 : h - 1][line + kHalo + XL + xl] = v;          // right halo of the left neighbour
        if (xl >= XL - kHalo) base[h == world - 1 ? 0 : h + 1][line + xl - (XL - kHalo)] = v;   // left halo of the right neighbour
    }
};

// Row strides of the column kernel's source (float2 elements between consecutive pair rows p of one channel) and
// destination (floats between output rows y).
template <int N>
struct FullColGeom {
    OW_HD size_t src_stride() const { return N; }
    OW_HD size_t dst_stride() const { return N; }
};
struct SlabColGeom {
    size_t ss, ds;
    OW_HD size_t src_stride() const { return ss; }
    OW_HD size_t dst_stride() const { return ds; }
};

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
#if OW_ABLATE & 2
    s = w * t; c = 1.0f - s; km = 1.0f + kx;
#else
    phase_sincos<FAST>(OW_MUL(w, t), &s, &c);                       // :96-97
#endif
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
// Folded initial spectrum. With e = e^{iwt}, A = h0 at (u,v), B = h0 at the mirror texel, the Hermitian part of dy is
//   S_y = H + conj(Hm) = (f0 c + f1 s,  f2 s + f3 c),   f = fold_pair(A, B)              (c = cos wt, s = sin wt)
// which depends on the EIGHT h0 floats of the pair only through FOUR time-independent sums. They are formed once at
// init (ow_fold_kernel), so the row kernel reads 16 B per texel PAIR instead of 32 and forms S_y in 4 flops. Away from
// the Nyquist column/row the mirror's wave vector is -k, hence S_x = -i (kx/|k|) S_y and S_z = -i (ky/|k|) S_y. On the
// Nyquist column (u = 0, where the shader's k is NOT negated by mirroring) S_x needs D = H - conj(Hm) instead:
//   D = (g0 c + g1 s, g2 s + g3 c),  g = fold_pair_nyq(A, B),   S_x = -i (kx/|k|) D       (one float4 per row pair)
// Pair 0 (rows 0 and N/2, their own mirrors) keeps the unfolded path (spectrum_sym) on the full h0 rows.
// ---------------------------------------------------------------------------------------------------
OW_HD float4 fold_pair(float4 A, float4 B) {
    return make_float4(((A.x + A.z) + B.x) + B.z, ((A.w + B.w) - A.y) - B.y, ((A.x - A.z) - B.x) + B.z, ((A.y + A.w) - B.y) - B.w);
}
OW_HD float4 fold_pair_nyq(float4 A, float4 B) {
    return make_float4(((A.x + A.z) - B.x) - B.z, ((A.w - A.y) - B.w) + B.y, ((A.x - A.z) + B.x) - B.z, ((A.y + A.w) + B.y) + B.w);
}

// Folded texel pair as loaded: f and (w, 1/|k|) at (u, p), k_x at u. g (Nyquist column only) is loaded by the one thread that owns u = 0.
struct FoldedPair {
    float4 f;
    float2 wk;
    float kx;
};

template <bool WK = true>
OW_HD FoldedPair load_folded(const float4* __restrict__ prow, const float2* __restrict__ wrow, const float* __restrict__ ktab, int u) {
    FoldedPair fp;
#if OW_ABLATE & 1
    fp.f = make_float4(u * 1e-3f, 1e-3f, u * 2e-3f, 1.0f); fp.kx = (u - 512) * 6.28e-3f; fp.wk = make_float2(1.0f + u * 1e-3f, 0.5f);
    return fp;
#endif
    fp.f = OW_LDG(prow + u);
    fp.wk = WK ? OW_LDG(wrow + u) : make_float2(0.f, 0.f);
    fp.kx = OW_LDG(ktab + u);
    return fp;
}

// (w, 1/|k|) of the texel (kx, ky): tilde_h0_t_cs.glsl:74-79 with the shader's operation order (the init kernels are compiled
// without FMA contraction), so w*t in the row kernel matches the oracle's to the bit.
OW_HD float2 dispersion_of(float kx, float ky) {
    float km = OW_SQRT(OW_ADD(OW_MUL(kx, kx), OW_MUL(ky, ky)));
    if (km < 0.00001f) km = 0.00001f;
    return make_float2(OW_SQRT(OW_MUL(kGravity, km)), 1.0f / km);
}
OW_HD float2 dispersion_fast(float kx, float ky) {      // in-kernel variant: same w, 1/|k| through the fast reciprocal
    float km = OW_SQRT(OW_ADD(OW_MUL(kx, kx), OW_MUL(ky, ky)));
    if (km < 0.00001f) km = 0.00001f;
    return make_float2(OW_SQRT(OW_MUL(kGravity, km)), OW_RCP(km));
}

template <bool FAST, bool WK = true>
OW_HD Sym3 spectrum_folded(const FoldedPair& fp, float ky, float t, const float4* __restrict__ nyq_g /* non-null iff u == 0 */) {
    const float kx = fp.kx;
    // tilde_h0_t_cs.glsl:74-79: tabulated at init (WK), or re-derived here with the same operation order - bit-identical either way
    const float2 wk = WK ? fp.wk : dispersion_fast(kx, ky);
    const float w = wk.x, ik = wk.y;
    float s, c;
#if OW_ABLATE & 2
    s = w * t; c = 1.0f - s;
#else
    phase_sincos<FAST>(OW_MUL(w, t), &s, &c);                       // :96-97
#endif
    const float4 f = fp.f;
    const float rx = kx * ik, rz = ky * ik;
    Sym3 o;
    o.y = make_float2(f.x * c + f.y * s, f.z * s + f.w * c);
    o.x = make_float2(rx * o.y.y, -rx * o.y.x);                     // :113-126 folded: -i (kx/|k|) S_y
    o.z = make_float2(rz * o.y.y, -rz * o.y.x);
    if (nyq_g) {
        const float4 g = OW_LDG(nyq_g);
        const float2 D = make_float2(g.x * c + g.y * s, g.z * s + g.w * c);
        o.x = make_float2(rx * D.y, -rx * D.x);
    }
    return o;
}

// ---------------------------------------------------------------------------------------------------
// ROW KERNEL.  One group of P::T threads transforms row pair p (rows p and N-p; p==0: rows 0 and N/2) for
// the three channels. Shared memory per group: 3 lines of P::LINE float2 (dy, dx, dz).
// Output: inter[c][p][x] (float2), c in {dy,dx,dz}, p < N/2, x < N — the row transform of S_c(., p).
// ---------------------------------------------------------------------------------------------------
// Pair 0: rows 0 (Nyquist) and N/2 (DC) mirror onto themselves; both row transforms are real, so they travel as one
// complex line Z = S(.,0) + i*S(.,N/2). It is ONE row pair in N/2, so this path is written for a small register
// footprint, not speed (rolled loops, the butterfly inputs parked in local memory, out of line): the hot path's
// register budget — and with it the occupancy of the whole kernel — must not be set by it.
template <class P, bool FAST, class Smem, class Rows>
__host__ __device__ __noinline__ void row_phase0_pair0(const Smem& sm, int ft, const Rows& rows, const float* __restrict__ ktab, float t) {
    constexpr int N = P::N, R0 = P::R0;
    const float ky0 = OW_LDG(ktab), kyh = OW_LDG(ktab + N / 2);
    const float4* row0 = rows.row(0);
    const float4* rowh = rows.row(N / 2);
#pragma unroll 1
    for (int c = 0; c < P::C0; ++c) {
        const int b = ft + P::T * c;
        if (b >= P::M) break;
        float2 v[3][R0];
#pragma unroll 1
        for (int d0 = 0; d0 < R0; ++d0) {
            const int u = d0 * P::M + b;
            const Sym3 a = spectrum_sym<FAST>(load_pair<N>(row0, row0, ktab, u), u, ky0, true, t);
            const Sym3 q = spectrum_sym<FAST>(load_pair<N>(rowh, rowh, ktab, u), u, kyh, true, t);
            v[0][d0] = make_float2(a.y.x - q.y.y, a.y.y + q.y.x);
            v[1][d0] = make_float2(a.x.x - q.x.y, a.x.y + q.x.x);
            v[2][d0] = make_float2(a.z.x - q.z.y, a.z.y + q.z.x);
        }
        float2 tw[R0];
        twiddle_powers<R0>(unit_root(b, N), tw);
#pragma unroll 1
        for (int f = 0; f < 3; ++f) {
            float2 w[R0];
#pragma unroll
            for (int d0 = 0; d0 < R0; ++d0) w[d0] = v[f][d0];
            stage0_finish<P>(sm, f * P::LINE, b, w, tw);
        }
    }
}

template <class P, bool FAST, class Smem, class Rows>
OW_HD void row_phase0(const Smem& sm, int ft, int p, const Rows& rows, const float* __restrict__ ktab,
                      float t) {
    constexpr int N = P::N, R0 = P::R0;
    if (p == 0) {
        row_phase0_pair0<P, FAST>(sm, ft, rows, ktab, t);
        return;
    }
    const float ky = OW_LDG(ktab + p);
    const float4* prow = rows.pair_row(p);
    const float2* wrow = rows.wk_row(p);
#pragma unroll 1
    for (int c = 0; c < P::C0; ++c) {
        const int b = ft + P::T * c;
        if (b >= P::M) break;
        float2 vy[R0], vx[R0], vz[R0];
        FoldedPair fp[R0];
#pragma unroll
        for (int d0 = 0; d0 < R0; ++d0) fp[d0] = load_folded<kUseWk<N>>(prow, wrow, ktab, d0 * P::M + b);
#pragma unroll
        for (int d0 = 0; d0 < R0; ++d0) {
            const Sym3 s = spectrum_folded<FAST, kUseWk<N>>(fp[d0], ky, t, (d0 == 0 && b == 0) ? rows.nyq_of(p) : nullptr);
            vy[d0] = s.y; vx[d0] = s.x; vz[d0] = s.z;
        }
        float2 tw[R0];
        twiddle_powers<R0>(unit_root(b, N), tw);
        stage0_finish<P>(sm, 0 * P::LINE, b, vy, tw);
        stage0_finish<P>(sm, 1 * P::LINE, b, vx, tw);
        stage0_finish<P>(sm, 2 * P::LINE, b, vz, tw);
    }
}

// Stages 1 and 2 of the three lines of a row pair. When a line has fewer stage butterflies than the group has threads (N = 2048:
// 128 butterflies, 256 threads) the 3 x B butterflies of the pair are dealt out as ONE list over all threads (task = line * B + id),
// so nobody idles through two thirds of the kernel; otherwise every thread does its butterflies of all three lines.
template <class P, class Smem>
OW_HD void row_phase1(const Smem& sm, int ft) {
#if OW_ABLATE & 4
    return;
#endif
    if (P::B1 < P::T) {
#pragma unroll 1
        for (int task = ft; task < 3 * P::B1; task += P::T) {
            const int f = task / P::B1, q = task - f * P::B1;
            float2 tw[P::R1];
            stage1_twiddles<P>(q, tw);
            stage1<P>(sm, f * P::LINE, q, tw);
        }
        return;
    }
    // d2 = q % R2 is the same for every butterfly q = ft + T*c of this thread when R2 divides T
    constexpr bool kSameTw = (P::T % P::R2 == 0);
    float2 tw[P::R1];
    if (kSameTw) stage1_twiddles<P>(ft, tw);
#pragma unroll 1
    for (int c = 0; c < P::C1; ++c) {
        const int q = ft + P::T * c;
        if (q >= P::B1) break;
        if (!kSameTw) stage1_twiddles<P>(q, tw);
#pragma unroll 1
        for (int f = 0; f < 3; ++f) stage1<P>(sm, f * P::LINE, q, tw);
    }
}

template <class P, class Smem, class Sink>
OW_HD void row_phase2_line(const Smem& sm, int f, int bp, int p, const Sink& sink) {
    float2 v[P::R2];
#if OW_ABLATE & 4
#pragma unroll
    for (int d2 = 0; d2 < P::R2; ++d2) v[d2] = sm.ld(f * P::LINE + P::addr(bp % P::R0, bp / P::R0, d2));
#else
    stage2<P>(sm, f * P::LINE, bp, v);
#endif
#if OW_ABLATE & 8
#pragma unroll
    for (int k2 = 0; k2 < P::R2; ++k2) if (v[k2].x == 12345.678f) sink.put(f, p, bp + k2 * P::B2, v[k2]);
#else
#pragma unroll
    for (int k2 = 0; k2 < P::R2; ++k2) sink.put(f, p, bp + k2 * P::B2, v[k2]);
#endif
}

template <class P, class Smem, class Sink>
OW_HD void row_phase2(const Smem& sm, int ft, int p, const Sink& sink) {
    if (P::B2 < P::T) {           // see row_phase1
#pragma unroll 1
        for (int task = ft; task < 3 * P::B2; task += P::T) {
            const int f = task / P::B2;
            row_phase2_line<P>(sm, f, task - f * P::B2, p, sink);
        }
        return;
    }
#pragma unroll 1
    for (int c = 0; c < P::C2; ++c) {
        const int bp = ft + P::T * c;
        if (bp >= P::B2) break;
#pragma unroll 1
        for (int f = 0; f < 3; ++f) row_phase2_line<P>(sm, f, bp, p, sink);
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

// Stage 0 of pair id j given its 2*H loaded rows: la[i] = row i*M + bA, lb[i] = row i*M + bB (bA = j, bB = M - j; j == 0: bB = M/2).
template <class P, class Smem>
OW_HD void col_phase0_math(const Smem& sm, int base, int j, const float4 (&la)[P::R0 / 2], const float4 (&lb)[P::R0 / 2]) {
    constexpr int N = P::N, R0 = P::R0, M = P::M, H = R0 / 2;
    const int bA = j, bB = (j == 0) ? M / 2 : M - j;
    float2 qa[R0], qb[R0];
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
    if (j != 0) {
        // bB = M - bA:  e^{2 pi i k0 (M - bA)/N} = e^{2 pi i k0/R0} * conj(e^{2 pi i k0 bA/N})  (R0-th roots are constants)
#pragma unroll
        for (int k0 = 1; k0 < R0; ++k0) tw[k0] = cmul(RootsOfUnity<R0>::get(k0), cconj(tw[k0]));
    } else {
        twiddle_powers<R0>(unit_root(bB, N), tw);
    }
    stage0_finish<P>(sm, base, bB, qb, tw);
}

template <class P, class Smem, class Geom>
OW_HD void col_phase0(const Smem& sm, int base, int j /* pair id in [0, M/2) */, const float2* __restrict__ src /* inter[c] + x */,
                      const Geom& geom, bool discard_line = false /* this thread drops the 128-B lines it read (job 0 of a G=8 tile) */) {
    constexpr int R0 = P::R0, M = P::M, H = R0 / 2;
    const size_t ss = geom.src_stride();
    const int bA = j, bB = (j == 0) ? M / 2 : M - j;
    float4 la[H], lb[H];
#pragma unroll
    for (int i = 0; i < H; ++i) {
        la[i] = OW_LDP(reinterpret_cast<const float4*>(src + (size_t)(i * M + bA) * ss));
        lb[i] = OW_LDP(reinterpret_cast<const float4*>(src + (size_t)(i * M + bB) * ss));
    }
    if (discard_line) {     // only AFTER the loaded values exist in registers: then every lane's piece of these lines has arrived
#pragma unroll
        for (int i = 0; i < H; ++i) {
            OW_USE_BEFORE(la[i].x, lb[i].x);
            OW_DISCARD_L2(src + (size_t)(i * M + bA) * ss);
            OW_DISCARD_L2(src + (size_t)(i * M + bB) * ss);
        }
    }
    col_phase0_math<P>(sm, base, j, la, lb);
}

// ---------------------------------------------------------------------------------------------------
// TMA staging of a column tile (ow_col2_kernel). One STEP = the rows the T pair ids j = s*T + ft (ft < T) of every job need:
// for each i < H a "forward" box of T consecutive rows starting at i*M + s*T (row of bA = j) and a "mirror" box of T consecutive
// rows starting at i*M + M - s*T - T + 1 (row of bB = M - j sits at box row T-1-ft). Every box row is the tile's G column pairs =
// G*16 contiguous bytes of one intermediate row, so a box is one 2-D TMA copy (T rows x G*16 B) and lands densely:
//   staging[(2*i + which) * T + r][job]  (float4).       Lane order is job-fastest, so a warp's LDS.128 covers whole box rows.
// The one row no box covers is i*M + M/2 (bB of j == 0, step 0): H extra rows behind the boxes, fetched by 1-D bulk copies.
// (The mirror box of step 0 ends at row i*M + M, which nobody reads: for i = H-1 it belongs to the next channel or is out of bounds
// and zero-filled.)
// ---------------------------------------------------------------------------------------------------
template <class P, int G>
struct ColStage {
    static constexpr int H = P::R0 / 2, T = P::T;
    static constexpr int STEPS = (P::M / 2) / T;
    static constexpr int BOX_F4 = T * G;                    // float4 elements per box
    static constexpr int EXTRA_F4 = 2 * H * BOX_F4;         // offset of the H extra rows
    static constexpr int TOTAL_F4 = EXTRA_F4 + H * G;
    static constexpr size_t BYTES = (size_t)TOTAL_F4 * sizeof(float4);
    static constexpr unsigned STEP_TX_BYTES = 2u * H * BOX_F4 * 16u;       // bytes the boxes of one step deliver
    static constexpr unsigned EXTRA_TX_BYTES = (unsigned)H * G * 16u;      // + the extra rows (step 0 only)
    static_assert((P::M / 2) % T == 0, "whole steps");
    static OW_HD int fwd_row0(int i, int s) { return i * P::M + s * T; }
    static OW_HD int mir_row0(int i, int s) { return i * P::M + P::M - s * T - T + 1; }
    static OW_HD int extra_row(int i) { return i * P::M + P::M / 2; }
};

// The 2*H rows pair id j = s*T + ft of `job` needs, out of the staging buffer (then: col_phase0_math).
template <class P, int G>
OW_HD void col_stage_read(const float4* __restrict__ staging, int job, int ft, int s, float4 (&la)[P::R0 / 2], float4 (&lb)[P::R0 / 2]) {
    using CS = ColStage<P, G>;
    constexpr int H = CS::H, T = CS::T;
    const int j = s * T + ft;
#pragma unroll
    for (int i = 0; i < H; ++i) {
        la[i] = staging[((2 * i + 0) * T + ft) * G + job];
        lb[i] = (j == 0) ? staging[CS::EXTRA_F4 + i * G + job] : staging[((2 * i + 1) * T + (T - 1 - ft)) * G + job];
    }
}

template <class P, class Smem>
OW_HD void col_phase1(const Smem& sm, int base, int ft) {
    constexpr bool kSameTw = (P::T % P::R2 == 0);
    float2 tw[P::R1];
    if (kSameTw) stage1_twiddles<P>(ft, tw);
#pragma unroll 1
    for (int c = 0; c < P::C1; ++c) {
        const int q = ft + P::T * c;
        if (q >= P::B1) break;
        if (!kSameTw) stage1_twiddles<P>(q, tw);
        stage1<P>(sm, base, q, tw);
    }
}

template <class P, class Smem, class Geom>
OW_HD void col_phase2(const Smem& sm, int base, int ft, float* __restrict__ dst /* out[c] + x */, float scale, const Geom& geom) {
    const size_t ds = geom.dst_stride();
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
            *reinterpret_cast<float2*>(dst + (size_t)y * ds) = make_float2(sg * v[k2].x, -sg * v[k2].y);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// LINES LONGER THAN ONE CTA'S SHARED MEMORY:  N = A * B with B = P::N <= 4096 and A in {2, 4, 8, 16}  (N = 8192 .. 32768,
// BASELINE config C5). Cooley-Tukey split n = A*m + a, k = kb + B*ka of the same unnormalised inverse DFT:
//     X[kb + B*ka] = sum_a W_A^{a ka} * ( W_N^{a kb} * z_a[kb] ),     z_a[kb] = sum_m x[A*m + a] W_B^{m kb},     W_n = e^{+2 pi i / n}
// Two kernels per direction ("four-step" FFT with the transposes folded into the index maps):
//   lines  one CTA (row kernel: one thread group) per sub-line a: the stage-0 loader gathers the decimated input x[A*m + a]
//          (rows: the spectrum evaluated at those texels; columns: the Hermitian-unpacked intermediate at those rows), the usual
//          three in-CTA stages transform the length-B sub-line, and z_a goes to a scratch array [a][kb] — 12 B/texel written once
//          and read once;
//   post   one thread per kb: the A values z_a[kb], the twiddles W_N^{a kb}, a radix-A DFT in registers, and A stores to
//          k = kb + B*ka — for each ka a run that is CONTIGUOUS in kb across the warp, so the final stores (rows: through the same
//          sinks as the direct kernel, i.e. the slab transpose over NVLink; columns: output rows with the inversion sign/scale)
//          are as coalesced as the direct kernels'.
// ---------------------------------------------------------------------------------------------------
// Row lines, stage 0, pair 0 only: rows 0 and N/2 are their own mirrors; Z = S(., 0) + i*S(., N/2) at the decimated texels
// u = A*m + a. One pair in N/2: small register footprint, not speed (see row_phase0_pair0).
template <class P, int A, bool FAST, class Smem, class Rows>
__host__ __device__ __noinline__ void bigrow_phase0_pair0(const Smem& sm, int ft, int a, const Rows& rows, const float* __restrict__ ktab, float t) {
    constexpr int B = P::N, N = A * B, R0 = P::R0;
    const float4* row0 = rows.row(0);
    const float4* rowh = rows.row(N / 2);
    const float ky0 = OW_LDG(ktab), kyh = OW_LDG(ktab + N / 2);
#pragma unroll 1
    for (int c = 0; c < P::C0; ++c) {
        const int b = ft + P::T * c;
        if (b >= P::M) break;
        float2 v[3][R0];
#pragma unroll 1
        for (int d0 = 0; d0 < R0; ++d0) {
            const int u = A * (d0 * P::M + b) + a;
            const Sym3 x0 = spectrum_sym<FAST>(load_pair<N>(row0, row0, ktab, u), u, ky0, true, t);
            const Sym3 xh = spectrum_sym<FAST>(load_pair<N>(rowh, rowh, ktab, u), u, kyh, true, t);
            v[0][d0] = make_float2(x0.y.x - xh.y.y, x0.y.y + xh.y.x);
            v[1][d0] = make_float2(x0.x.x - xh.x.y, x0.x.y + xh.x.x);
            v[2][d0] = make_float2(x0.z.x - xh.z.y, x0.z.y + xh.z.x);
        }
        float2 tw[R0];
        twiddle_powers<R0>(unit_root(b, B), tw);
#pragma unroll 1
        for (int f = 0; f < 3; ++f) {
            float2 w[R0];
#pragma unroll
            for (int d0 = 0; d0 < R0; ++d0) w[d0] = v[f][d0];
            stage0_finish<P>(sm, f * P::LINE, b, w, tw);
        }
    }
}

// Contexts that run the line decomposition keep every folded pair row (and a copy of the k table) SUB-LINE-MAJOR: texel u = A*m + a
// sits at index a*B + m, so the B texels of sub-line a are contiguous. The lines kernel's stage-0 loads are then coalesced 512-byte runs
// per warp instruction; in the natural order they were 32 separate 16-byte pieces 16*A bytes apart, i.e. 32 L1 wavefronts per load
// instruction - two thirds of the kernel's L1/shared-memory pipe time at N = 32768.
OW_HD int subline_index(int u, int A, int N) { return (u % A) * (N / A) + u / A; }

// Row lines, stage 0: spectrum at the decimated texels u = A*m + a of pair p — row_phase0 on the sub-line-major folded row
// (all loads of a butterfly in flight before the first use). ktab_sub = the k table in the same order (kx of the texels);
// ktab = the natural one (ky of the pair, pair 0's literal path).
template <class P, int A, bool FAST, class Smem, class Rows>
OW_HD void bigrow_phase0(const Smem& sm, int ft, int p, int a, const Rows& rows, const float* __restrict__ ktab, const float* __restrict__ ktab_sub,
                         float t) {
    constexpr int B = P::N, R0 = P::R0;
    if (p == 0) {
        bigrow_phase0_pair0<P, A, FAST>(sm, ft, a, rows, ktab, t);
        return;
    }
    const float ky = OW_LDG(ktab + p);
    const float4* prow = rows.pair_row(p) + (size_t)a * B;      // sub-line a of the pair's folded row
    const float2* wrow = rows.wk_row(p);
    const float* ksub = ktab_sub + (size_t)a * B;
#pragma unroll 1
    for (int c = 0; c < P::C0; ++c) {
        const int b = ft + P::T * c;
        if (b >= P::M) break;
        float2 vy[R0], vx[R0], vz[R0];
        FoldedPair fp[R0];
#pragma unroll
        for (int d0 = 0; d0 < R0; ++d0) fp[d0] = load_folded<false>(prow, wrow, ksub, d0 * P::M + b);
#pragma unroll
        for (int d0 = 0; d0 < R0; ++d0) {
            const Sym3 s = spectrum_folded<FAST, false>(fp[d0], ky, t, (d0 == 0 && b == 0 && a == 0) ? rows.nyq_of(p) : nullptr);
            vy[d0] = s.y; vx[d0] = s.x; vz[d0] = s.z;
        }
        float2 tw[R0];
        twiddle_powers<R0>(unit_root(b, B), tw);
        stage0_finish<P>(sm, 0 * P::LINE, b, vy, tw);
        stage0_finish<P>(sm, 1 * P::LINE, b, vx, tw);
        stage0_finish<P>(sm, 2 * P::LINE, b, vz, tw);
    }
}

// Row lines, stage 2: z_a[kb] of the three channels -> scratch row of pair p, layout [c][a][kb].
template <class P, int A, class Smem>
OW_HD void bigrow_phase2(const Smem& sm, int ft, int a, float2* __restrict__ zrow) {
    constexpr int B = P::N;
#pragma unroll 1
    for (int c = 0; c < P::C2; ++c) {
        const int bp = ft + P::T * c;
        if (bp >= P::B2) break;
#pragma unroll 1
        for (int f = 0; f < 3; ++f) {
            float2 v[P::R2];
            stage2<P>(sm, f * P::LINE, bp, v);
            float2* dst = zrow + (size_t)(f * A + a) * B + bp;
#pragma unroll
            for (int k2 = 0; k2 < P::R2; ++k2) dst[k2 * P::B2] = v[k2];
        }
    }
}

// Row post for (pair p, channel c, kb): X[kb + B*ka] for all ka, through the sink.
template <int B, int A, class Sink>
OW_HD void bigrow_post(const float2* __restrict__ zrow /* scratch row of pair p: [3][A][B] */, int c, int p, int kb, const Sink& sink) {
    constexpr int N = A * B;
    float2 v[A], tw[A];
    twiddle_powers<A>(unit_root(kb, N), tw);                      // W_N^{a kb}
#pragma unroll
    for (int a = 0; a < A; ++a) {
        const float2 z = OW_LDG(zrow + (size_t)(c * A + a) * B + kb);
        v[a] = a ? cmul(z, tw[a]) : z;
    }
    Dft<A>::run(v);
#pragma unroll
    for (int ka = 0; ka < A; ++ka) sink.put(c, p, kb + B * ka, v[ka]);
}

// Column lines, stage 0: the decimated source rows v = A*m + a of this job's column pair. Loading is split from the arithmetic so that
// the pipelined kernel can keep the next batch's R0 row loads in flight while this batch is transformed.
// Row of v: v itself below N/2, N - v above (conjugated on unpacking), row 0 for the two packed real rows v = 0 and v = N/2.
template <class P, int A>
OW_HD void bigcol_issue(int b, int a, const float2* __restrict__ src /* inter[c] + x */, size_t ss, float4 (&r)[P::R0]) {
    constexpr int B = P::N, N = A * B, R0 = P::R0;
#pragma unroll
    for (int d0 = 0; d0 < R0; ++d0) {
        const int vv = A * (d0 * P::M + b) + a;
        const int row = (vv < N / 2 ? vv : N - vv) & (N / 2 - 1);
        r[d0] = OW_LDG(reinterpret_cast<const float4*>(src + (size_t)row * ss));
    }
}

template <class P, int A, class Smem>
OW_HD void bigcol_phase0_math(const Smem& sm, int base, int b, int a, const float4 (&r)[P::R0]) {
    constexpr int B = P::N, N = A * B, R0 = P::R0;
    float2 v[R0], tw[R0];
#pragma unroll
    for (int d0 = 0; d0 < R0; ++d0) {
        const int vv = A * (d0 * P::M + b) + a;
        v[d0] = vv == 0 ? make_float2(r[d0].x, r[d0].z) : vv == N / 2 ? make_float2(r[d0].y, r[d0].w) : vv < N / 2 ? pack_fwd(r[d0]) : pack_cnj(r[d0]);
    }
    twiddle_powers<R0>(unit_root(b, B), tw);
    stage0_finish<P>(sm, base, b, v, tw);
}

// Two batches of row loads (2*R0 float4 per thread) are put in flight together when the sub-line has an even number of batches: a CTA
// of this kernel is alone on its SM (N = 32768: 141 KB of shared memory), so every load round trip it waits for is exposed.
#ifndef OW_BIGCOL_BATCH2
#define OW_BIGCOL_BATCH2 1
#endif
template <class P, int A, class Smem, class Geom>
OW_HD void bigcol_phase0(const Smem& sm, int base, int ft, int a, const float2* __restrict__ src /* inter[c] + x */, const Geom& geom) {
    const size_t ss = geom.src_stride();
    if (OW_BIGCOL_BATCH2 && (P::M / P::T) % 2 == 0 && P::M % P::T == 0) {
#pragma unroll 1
        for (int b = ft; b < P::M; b += 2 * P::T) {
            float4 r0[P::R0], r1[P::R0];
            bigcol_issue<P, A>(b, a, src, ss, r0);
            bigcol_issue<P, A>(b + P::T, a, src, ss, r1);
            bigcol_phase0_math<P, A>(sm, base, b, a, r0);
            bigcol_phase0_math<P, A>(sm, base, b + P::T, a, r1);
        }
        return;
    }
#pragma unroll 1
    for (int b = ft; b < P::M; b += P::T) {
        float4 r[P::R0];
        bigcol_issue<P, A>(b, a, src, ss, r);
        bigcol_phase0_math<P, A>(sm, base, b, a, r);
    }
}

// Column lines, stage 2: z_a[kb] of this job's column pair -> scratch [a][kb][pair] (zsub = scratch[c][a] + pair).
template <class P, class Smem>
OW_HD void bigcol_phase2(const Smem& sm, int base, int ft, float2* __restrict__ zsub, size_t zs) {
#pragma unroll 1
    for (int c = 0; c < P::C2; ++c) {
        const int bp = ft + P::T * c;
        if (bp >= P::B2) break;
        float2 v[P::R2];
        stage2<P>(sm, base, bp, v);
#pragma unroll
        for (int k2 = 0; k2 < P::R2; ++k2) zsub[(size_t)(bp + P::B2 * k2) * zs] = v[k2];
    }
}

// Column post for (column pair, kb): output rows y = kb + B*ka with the inversion sign/scale (inversion_cs.glsl:29-36).
template <int B, int A>
OW_HD void bigcol_post(const float2* __restrict__ z /* scratch[c] + pair */, size_t zs, int kb, float* __restrict__ dst /* out[c] + x */, size_t ds,
                       float scale) {
    constexpr int N = A * B;
    float2 v[A], tw[A];
    twiddle_powers<A>(unit_root(kb, N), tw);
#pragma unroll
    for (int a = 0; a < A; ++a) {
        const float2 q = OW_LDG(z + ((size_t)a * B + kb) * zs);
        v[a] = a ? cmul(q, tw[a]) : q;
    }
    Dft<A>::run(v);
    const float sg = (kb & 1) ? -scale : scale;            // y = kb + B*ka has the parity of kb (B is even); x is even
#pragma unroll
    for (int ka = 0; ka < A; ++ka)
        *reinterpret_cast<float2*>(dst + (size_t)(kb + B * ka) * ds) = make_float2(sg * v[ka].x, -sg * v[ka].y);
}

// ---------------------------------------------------------------------------------------------------
// THE SAME DECOMPOSITION INSIDE ONE THREAD-BLOCK CLUSTER (sm_90+; ow_bigrow_cluster_kernel / ow_bigcol_cluster_kernel).
// The A sub-lines of a line are transformed by the A CTAs of a cluster; instead of leaving through a global scratch array, z_a stays
// in its CTA's shared memory (stage 2 writes it back IN PLACE: slot (k0,k1,k2) of the line then holds z_a[k0 + R0 k1 + R0 R1 k2]),
// the cluster synchronises, and the radix-A stage reads the A partial results of each kb straight out of the peers' shared memory
// (distributed shared memory). CTA a finishes the outputs of kb in [a B/A, (a+1) B/A). No scratch traffic: 24 B/texel less per direction
// and one kernel instead of two. `Peers::ld(a, i)` loads element i of CTA a's lines (mapa + ld.shared::cluster on the device).
// ---------------------------------------------------------------------------------------------------
template <class P, class Smem>
OW_HD void stage2_inplace(const Smem& sm, int base, int bp) {
    const int k0 = bp % P::R0, k1 = bp / P::R0;
    float2 v[P::R2];
    stage2<P>(sm, base, bp, v);
#pragma unroll
    for (int k2 = 0; k2 < P::R2; ++k2) sm.st(base + P::addr(k0, k1, k2), v[k2]);
}

template <class P>
OW_HD int slot_of(int kb) { return P::addr(kb % P::R0, (kb / P::R0) % P::R1, kb / (P::R0 * P::R1)); }

// Stage 2 of the three lines of a row pair, in place (task list over the lines, as row_phase2).
template <class P, class Smem>
OW_HD void bigrow_phase2_inplace(const Smem& sm, int ft) {
#pragma unroll 1
    for (int task = ft; task < 3 * P::B2; task += P::T) {
        const int f = task / P::B2;
        stage2_inplace<P>(sm, f * P::LINE, task - f * P::B2);
    }
}

// Radix-A stage of (channel c, kb) of row pair p out of the cluster's shared memory, through the sink.
template <class P, int A, class Peers, class Sink>
OW_HD void bigrow_post_dsm(const Peers& peers, int c, int p, int kb, const Sink& sink) {
    constexpr int B = P::N, N = A * B;
    float2 v[A], tw[A];
    twiddle_powers<A>(unit_root(kb, N), tw);                      // W_N^{a kb}
    const int i = c * P::LINE + slot_of<P>(kb);
#pragma unroll
    for (int a = 0; a < A; ++a) {
        const float2 z = peers.ld(a, i);
        v[a] = a ? cmul(z, tw[a]) : z;
    }
    Dft<A>::run(v);
#pragma unroll
    for (int ka = 0; ka < A; ++ka) sink.put(c, p, kb + B * ka, v[ka]);
}

// Column direction: stage 2 of this job's line in place ...
template <class P, class Smem>
OW_HD void bigcol_phase2_inplace(const Smem& sm, int base, int ft) {
#pragma unroll 1
    for (int bp = ft; bp < P::B2; bp += P::T) stage2_inplace<P>(sm, base, bp);
}

// ... and the radix-A stage of (job, kb): output rows y = kb + B*ka of the job's column pair, inversion sign/scale as bigcol_post.
template <class P, int A, class Peers>
OW_HD void bigcol_post_dsm(const Peers& peers, int base /* job's line offset */, int kb, float* __restrict__ dst /* out[c] + x */, size_t ds, float scale) {
    constexpr int B = P::N, N = A * B;
    float2 v[A], tw[A];
    twiddle_powers<A>(unit_root(kb, N), tw);
    const int i = base + slot_of<P>(kb);
#pragma unroll
    for (int a = 0; a < A; ++a) {
        const float2 q = peers.ld(a, i);
        v[a] = a ? cmul(q, tw[a]) : q;
    }
    Dft<A>::run(v);
    const float sg = (kb & 1) ? -scale : scale;
#pragma unroll
    for (int ka = 0; ka < A; ++ka)
        *reinterpret_cast<float2*>(dst + (size_t)(kb + B * ka) * ds) = make_float2(sg * v[ka].x, -sg * v[ka].y);
}

// ---------------------------------------------------------------------------------------------------
// NORMAL (+ JACOBIAN).  normal_map_cs.glsl:24-54: the eight texture() taps sit on texel corners, so with
// LINEAR+REPEAT each tap is the mean of a 2x2 block ("box"); the stencil covers columns x-2..x+1 and rows
// y-2..y+1 with wrap-around. One thread owns FOUR adjacent columns x0..x0+3 and walks down RY output rows with
// a sliding window (all sums are kept unscaled, V = 4*box; the 1/4 is applied once to nx, nz):
//   hs(r)[c] = h[r][x0+c-2] + h[r][x0+c-1]          c = 0..5   (7 loaded values: one float4, one float2, one float)
//   V(r)[c]  = hs(r-1)[c] + hs(r)[c]                = 4 * box at tap-row r, tap-columns x0-1 .. x0+4
//   sx(r)[j] = V[j] + 2 V[j+1] + V[j+2],   dxb(r)[j] = V[j] - V[j+2]          j = 0..3 (column x0+j)
//   n.z(y) = (sx(y-1) - sx(y+1)) / 4     (:49)      n.x(y) = (dxb(y-1) + 2 dxb(y) + dxb(y+1)) / 4     (:50)
// which costs ~20 instructions per texel instead of 16 texture taps. Lanes run along x (a warp covers 128
// columns), every global access is a coalesced 16/8/4-byte-per-lane request; the normals of a row are handed to
// `emit`, which on the device transposes them through shared memory so each store instruction writes 512
// contiguous bytes. The Jacobian (extension) rides the same walk:
//   J = (1 - l*dDx/dx)(1 - l*dDz/dz) - l^2 (dDx/dz)(dDz/dx), central differences with wrap, spacing L/N;
//   s = l * N / (2 L) is folded into each difference.
// ---------------------------------------------------------------------------------------------------
// Everything one walk iteration needs from global memory (h row r): loaded one iteration AHEAD of its use, so
// the L2 round trip overlaps the previous row's arithmetic and staged stores (which contain warp barriers the
// compiler will not move loads across).
struct NormalRowIn {
    float2 l;            // dy[x0-2], dy[x0-1]
    float4 m;            // dy[x0 .. x0+3]
    float e;             // dy[x0+4]
    float4 a, b;         // Dx[x0 .. x0+3], Dz[x0 .. x0+3]
    float al, ar, bl, br;   // Dx[x0-1], Dx[x0+4], Dz[x0-1], Dz[x0+4]
};

// Geometry of the displacement planes the stencil reads. Full grid: [3][N][N], x wraps with mask N-1. Column slab:
// [3][N][XH] with halo columns present, so x never wraps (mask = -1) and x0 is an index into the padded row.
template <int N>
struct FullNrmGeom {
    OW_HD size_t row_stride() const { return N; }
    OW_HD size_t plane_stride() const { return (size_t)N * N; }
    OW_HD int xmask() const { return N - 1; }
};
struct SlabNrmGeom {
    size_t rs, ps;
    OW_HD size_t row_stride() const { return rs; }
    OW_HD size_t plane_stride() const { return ps; }
    OW_HD int xmask() const { return -1; }
};

template <int N, bool JAC, class Geom>
OW_HD NormalRowIn normal_row_load(const float* __restrict__ disp, const Geom& geom, int x0, int rr, bool want_xz, bool want_dd) {
    const int MSK = geom.xmask();
    const int xl2 = (x0 - 2) & MSK, xl1 = (x0 - 1) & MSK, xr = (x0 + 4) & MSK;
    const float* r = disp + (size_t)rr * geom.row_stride();
    NormalRowIn in;
    in.l = OW_LDP(reinterpret_cast<const float2*>(r + xl2));
    in.m = OW_LDP(reinterpret_cast<const float4*>(r + x0));
    in.e = OW_LDP(r + xr);
    in.a = in.b = make_float4(0.f, 0.f, 0.f, 0.f);
    in.al = in.ar = in.bl = in.br = 0.f;
    if (JAC && want_xz) {
        const float* rx = r + geom.plane_stride();
        const float* rz = r + 2 * geom.plane_stride();
        in.a = OW_LDP(reinterpret_cast<const float4*>(rx + x0));
        in.b = OW_LDP(reinterpret_cast<const float4*>(rz + x0));
        if (want_dd) {
            in.al = OW_LDP(rx + xl1); in.ar = OW_LDP(rx + xr);
            in.bl = OW_LDP(rz + xl1); in.br = OW_LDP(rz + xr);
        }
    }
    return in;
}

// Row source of the walk: where the seven heights (and, for the Jacobian, Dx/Dz) of h row rr come from.
template <int N, bool JAC, class Geom>
struct GlobalRowSrc {           // the displacement planes in global memory (the stand-alone normal kernel)
    const float* disp;
    Geom geom;
    int x0;
    OW_HD NormalRowIn load(int rr, bool want_xz, bool want_dd) const { return normal_row_load<N, JAC>(disp, geom, x0, rr, want_xz, want_dd); }
};

template <int N, int RY, bool JAC, class Src, class Emit>
OW_HD void normal_quad_walk_src(const Src& src, int y0, float s, const Emit& emit);

template <int N, int RY, bool JAC, class Geom, class Emit>
OW_HD void normal_quad_walk(const float* __restrict__ disp /* dy,dx,dz planes */, const Geom& geom, int x0, int y0, float s, const Emit& emit) {
    normal_quad_walk_src<N, RY, JAC>(GlobalRowSrc<N, JAC, Geom>{disp, geom, x0}, y0, s, emit);
}

template <int N, int RY, bool JAC, class Src, class Emit>
OW_HD void normal_quad_walk_src(const Src& src, int y0, float s, const Emit& emit) {
    constexpr int MSK = N - 1;
    float hs_prev[6], sx_m1[4], sx_0[4], dxb_m1[4], dxb_0[4];
    float xc_m1[4], xc_0[4], zc_m1[4], zc_0[4], ddx_0[4], ddz_0[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        sx_m1[j] = sx_0[j] = dxb_m1[j] = dxb_0[j] = 0.f;
        xc_m1[j] = xc_0[j] = zc_m1[j] = zc_0[j] = ddx_0[j] = ddz_0[j] = 0.f;
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) hs_prev[c] = 0.f;
    NormalRowIn nxt = src.load((y0 - 2) & MSK, false, false);
#pragma unroll
    for (int i = -2; i <= RY; ++i) {              // h row r = y0 + i ; emits output row r - 1 once i >= 1
        const NormalRowIn in = nxt;
        if (i < RY) nxt = src.load((y0 + i + 1) & MSK, i + 1 >= -1, i + 1 >= 0 && i + 1 < RY);
        const float4 m = in.m;
        const float hs[6] = {in.l.x + in.l.y, in.l.y + m.x, m.x + m.y, m.y + m.z, m.z + m.w, m.w + in.e};
        float sx[4] = {0.f, 0.f, 0.f, 0.f}, dxb[4] = {0.f, 0.f, 0.f, 0.f};
        if (i >= -1) {
            float V[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) V[c] = hs_prev[c] + hs[c];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                sx[j] = fmaf(2.0f, V[j + 1], V[j] + V[j + 2]);
                dxb[j] = V[j] - V[j + 2];
            }
        }
#pragma unroll
        for (int c = 0; c < 6; ++c) hs_prev[c] = hs[c];
        const float4 a = in.a, b = in.b;
        const float xc[4] = {a.x, a.y, a.z, a.w}, zc[4] = {b.x, b.y, b.z, b.w};
        // x differences of Dx and Dz at row r (only meaningful for 0 <= i < RY; otherwise unused)
        const float ddx[4] = {a.y - in.al, a.z - a.x, a.w - a.y, in.ar - a.z};
        const float ddz[4] = {b.y - in.bl, b.z - b.x, b.w - b.y, in.br - b.z};
        if (i >= 1) {
            float4 n[4];
            float J[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 4; ++j) {         // output row y = y0+i-1: taps y-1 (.._m1), y (.._0), y+1 (current)
                const float nz = 0.25f * (sx_m1[j] - sx[j]);
                const float nx = 0.25f * (dxb_m1[j] + 2.0f * dxb_0[j] + dxb[j]);
                const float rinv = rsqrtf(fmaf(nx, nx, fmaf(nz, nz, 1.0f)));
                n[j] = make_float4(nx * rinv, rinv, nz * rinv, 1.0f);   // :53
                if (JAC)
                    J[j] = fmaf(-s, ddx_0[j], 1.0f) * fmaf(-s, zc[j] - zc_m1[j], 1.0f) - (s * (xc[j] - xc_m1[j])) * (s * ddz_0[j]);
            }
            emit(y0 + i - 1, n, make_float4(J[0], J[1], J[2], J[3]));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            sx_m1[j] = sx_0[j]; sx_0[j] = sx[j]; dxb_m1[j] = dxb_0[j]; dxb_0[j] = dxb[j];
            xc_m1[j] = xc_0[j]; xc_0[j] = xc[j]; zc_m1[j] = zc_0[j]; zc_0[j] = zc[j]; ddx_0[j] = ddx[j]; ddz_0[j] = ddz[j];
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// COLUMN KERNEL WITH THE NORMAL MAP AS ITS EPILOGUE (ow_col2_kernel; tiles of G = 8 column pairs = 16 columns).
// After stage 2 a dy tile writes its final heights back IN PLACE into its lines (row y = k0 + R0 k1 + R0 R1 k2 at addr(k0,k1,k2);
// .x = even column, .y = odd column), the tile syncs, and the same sliding-window stencil as the stand-alone normal kernel runs
// out of shared memory for the three column quads whose 4x4 neighbourhoods lie inside the tile: columns c0+2 .. c0+13 (c0 = first
// column of the tile; a quad at x0 needs columns x0-2 .. x0+4). One thread per (column quad, RY rows), lanes spread over the
// unit-stride digit k2 so that every LDS.64 phase is conflict-free. Those heights never make the round trip through L2/HBM.
// The fourth quad of every 16 columns, c0+14 .. c0+17, straddles two tiles: a SEAM. It is computed from global memory (L2) by
// whichever of the two neighbouring tiles finishes LAST (an atomic counter per seam: no waiting, no ordering assumption between
// CTAs), with the same walk on the same stored heights - so the image is bit-identical to the separate normal kernel's.
// No column is transformed twice (the round-1 variant re-transformed 2 halo pairs per 6 output pairs).
// ---------------------------------------------------------------------------------------------------
template <class P, class Smem, class Geom>
OW_HD void col_phase2_keep(const Smem& sm, int base, int ft, float* __restrict__ dst /* out[c] + x */, float scale, const Geom& geom, bool store) {
    const size_t ds = geom.dst_stride();
#pragma unroll 1
    for (int c = 0; c < P::C2; ++c) {
        const int bp = ft + P::T * c;
        if (bp >= P::B2) break;
        const int k0 = bp % P::R0, k1 = bp / P::R0;
        float2 v[P::R2];
        stage2<P>(sm, base, bp, v);
        const float sg = (bp & 1) ? -scale : scale;
#pragma unroll
        for (int k2 = 0; k2 < P::R2; ++k2) {
            const float2 h = make_float2(sg * v[k2].x, -sg * v[k2].y);
            if (store) *reinterpret_cast<float2*>(dst + (size_t)(bp + P::B2 * k2) * ds) = h;
            sm.st(base + P::addr(k0, k1, k2), h);
        }
    }
}

template <class P, class Smem>
struct SmemRowSrc {             // final heights of four adjacent jobs (pairs q-1, q, q+1, q+2 of a column quad) in the tile's lines
    Smem sm;
    int base0, SJ;              // line base of the leftmost job; float2 elements between jobs
    OW_HD NormalRowIn load(int rr, bool, bool) const {
        const int a = base0 + P::addr(rr % P::R0, (rr / P::R0) % P::R1, rr / (P::R0 * P::R1));
        const float2 q0 = sm.ld(a), q1 = sm.ld(a + SJ), q2 = sm.ld(a + 2 * SJ), q3 = sm.ld(a + 3 * SJ);
        NormalRowIn in;
        in.l = q0; in.m = make_float4(q1.x, q1.y, q2.x, q2.y); in.e = q3.x;
        in.a = in.b = make_float4(0.f, 0.f, 0.f, 0.f);
        in.al = in.ar = in.bl = in.br = 0.f;
        return in;
    }
};

// Emit of the fused epilogue. The four lanes of a group (c = lane % 4) hold the normals of the SAME column quad on four
// different rows (y_c = y + (r - c) * row_step for the lane r). A 4x4 transpose through two xor-shuffle steps turns
// "lane c: four columns of row y_c" into "lane c: column c of rows y_0..y_3", so that every store instruction covers whole
// rows: 64 contiguous bytes per group, 192 per half-warp (the three quads of a tile are neighbours in the lane index)
// instead of 32 scattered 16-byte pieces per warp. On the host (emulator) the plain per-thread stores give the same image.
struct EmitQuad {
    float4* normal;             // slot base, [N][N]
    size_t ostride;
    int xout;                   // first column of the quad
    int c;                      // lane % 4
    int row_step;               // rows between the lanes of a group
    OW_HD void operator()(int y, const float4 (&n)[4], float4) const {
#ifdef __CUDA_ARCH__
        float4 a[4] = {n[0], n[1], n[2], n[3]};
        const unsigned mask = 0xFu << ((threadIdx.x & 31) & ~3);
#pragma unroll
        for (int bit = 0; bit < 2; ++bit) {
            const bool up = (c >> bit) & 1;
#pragma unroll
            for (int lo = 0; lo < 4; ++lo) {
                if ((lo >> bit) & 1) continue;
                const int hi = lo | (1 << bit);
                const float4 send = up ? a[lo] : a[hi];
                float4 recv;
                recv.x = __shfl_xor_sync(mask, send.x, 1 << bit); recv.y = __shfl_xor_sync(mask, send.y, 1 << bit);
                recv.z = __shfl_xor_sync(mask, send.z, 1 << bit); recv.w = __shfl_xor_sync(mask, send.w, 1 << bit);
                if (up) a[lo] = recv; else a[hi] = recv;
            }
        }
        // a[r] = column c of the row held by lane r of the group
#pragma unroll
        for (int r = 0; r < 4; ++r) __stcs(normal + (size_t)(y + (r - c) * row_step) * ostride + xout + c, a[r]);
#else
        float4* d = normal + (size_t)y * ostride + xout;
        for (int j = 0; j < 4; ++j) d[j] = n[j];
#endif
    }
};

// Normal-map epilogue of one dy tile. Lane layout inside a half-warp: c = lane % 4 -> low bits of the unit-stride digit k2,
// j = (lane / 4) % 4 -> column quad (3 quads per tile; j == 3 idles). The 16 LDS.64 of a phase then hit 16 different bank pairs
// (jobs are SJ = 2 (mod 16) apart: bank pair = const + c + 4j), and the stores of the 12 active lanes are 192 contiguous bytes.
template <class P, int RY, class Smem>
OW_HD void col_normals_phase(const Smem& sm, int tid, int nthreads, int SJ, int x_first /* first output column of the tile */,
                             float4* __restrict__ normal /* slot base */) {
    constexpr int N = P::N, LOW = P::R0 * P::R1, WPER = LOW / RY, NITEMS = 4 * P::R2 * WPER;
    static_assert(LOW % RY == 0 && P::R2 % 4 == 0, "row chunks must tile a block of R0*R1 rows; k2 is split 4 x R2/4");
#pragma unroll 1
    for (int item = tid; item < NITEMS; item += nthreads) {
        const int c = item % 4, j = (item / 4) % 4, h = item / 16;
        const int k2 = c + 4 * (h % (P::R2 / 4)), w = h / (P::R2 / 4);
        const int x0 = x_first + 4 * j;
        if (j == 3 || x0 >= N) continue;                        // idle quad slot / wrapped duplicate pairs of the last tile
        const SmemRowSrc<P, Smem> src{sm, 2 * j * SJ, SJ};
        normal_quad_walk_src<N, RY, false>(src, LOW * k2 + RY * w, 0.f, EmitQuad{normal, (size_t)N, x0, c, LOW});
    }
}

// Seam quad k of a slot: output columns cA+2, cA+3 (the last two of tile k) and cB, cB+1 (the first two of tile k+1, wrapping),
// cA = 16k+12, cB = 16(k+1) mod N. The seven heights of a row are two aligned float4 loads from L2 (never L1: the neighbour tile
// was written by another SM during this kernel).
struct SeamRowSrc {
    const float* dy;        // slot's dy plane [N][N]
    size_t stride;
    int cA, cB;
    OW_HD NormalRowIn load(int rr, bool, bool) const {
        const float* r = dy + (size_t)rr * stride;
#ifdef __CUDA_ARCH__
        const float4 A = __ldcg(reinterpret_cast<const float4*>(r + cA)), B = __ldcg(reinterpret_cast<const float4*>(r + cB));
#else
        const float4 A = *reinterpret_cast<const float4*>(r + cA), B = *reinterpret_cast<const float4*>(r + cB);
#endif
        NormalRowIn in;
        in.l = make_float2(A.x, A.y); in.m = make_float4(A.z, A.w, B.x, B.y); in.e = B.z;
        in.a = in.b = make_float4(0.f, 0.f, 0.f, 0.f);
        in.al = in.ar = in.bl = in.br = 0.f;
        return in;
    }
};

struct EmitSeam {
    float4* normal;         // slot base, [N][N]
    size_t ostride;
    int cA, cB;
    OW_HD void operator()(int y, const float4 (&n)[4], float4) const {
        float4* d = normal + (size_t)y * ostride;
#ifdef __CUDA_ARCH__
        __stcs(d + cA + 2, n[0]); __stcs(d + cA + 3, n[1]); __stcs(d + cB, n[2]); __stcs(d + cB + 1, n[3]);
#else
        d[cA + 2] = n[0]; d[cA + 3] = n[1]; d[cB] = n[2]; d[cB + 1] = n[3];
#endif
    }
};

// All N rows of seam k, RY rows per thread.
template <int N, int RY>
OW_HD void col_seam_phase(const float* __restrict__ dy, float4* __restrict__ normal, int k, int tid, int nthreads) {
    const int cA = 16 * k + 12, cB = (16 * (k + 1)) & (N - 1);
    const SeamRowSrc src{dy, (size_t)N, cA, cB};
    const EmitSeam emit{normal, (size_t)N, cA, cB};
#pragma unroll 1
    for (int w = tid; w < N / RY; w += nthreads) normal_quad_walk_src<N, RY, false>(src, w * RY, 0.f, emit);
}

// ---------------------------------------------------------------------------------------------------
// JACOBIAN ALONE (extension) for frames whose normal map came out of the column kernel: the same expression, in the
// same operation order, as the Jacobian that rides normal_quad_walk_src, so both paths give identical images. One thread owns four
// adjacent columns and walks down RY rows; row r of Dx and Dz is loaded one iteration ahead of its use.
//   J = (1 - s (Dx[y][x+1] - Dx[y][x-1])) (1 - s (Dz[y+1][x] - Dz[y-1][x])) - (s (Dx[y+1][x] - Dx[y-1][x])) (s (Dz[y][x+1] - Dz[y][x-1]))
// ---------------------------------------------------------------------------------------------------
struct JacRowIn {
    float4 a, b;            // Dx[x0 .. x0+3], Dz[x0 .. x0+3]
    float al, ar, bl, br;   // Dx[x0-1], Dx[x0+4], Dz[x0-1], Dz[x0+4]
};

template <class Geom>
OW_HD JacRowIn jac_row_load(const float* __restrict__ disp, const Geom& geom, int x0, int rr, bool want_dd) {
    const int MSK = geom.xmask();
    const int xl1 = (x0 - 1) & MSK, xr = (x0 + 4) & MSK;
    const float* rx = disp + (size_t)rr * geom.row_stride() + geom.plane_stride();
    const float* rz = rx + geom.plane_stride();
    JacRowIn in;
    in.a = OW_LDP(reinterpret_cast<const float4*>(rx + x0));
    in.b = OW_LDP(reinterpret_cast<const float4*>(rz + x0));
    in.al = in.ar = in.bl = in.br = 0.f;
    if (want_dd) {
        in.al = OW_LDP(rx + xl1); in.ar = OW_LDP(rx + xr);
        in.bl = OW_LDP(rz + xl1); in.br = OW_LDP(rz + xr);
    }
    return in;
}

template <int N, int RY, class Geom, class Emit>
OW_HD void jac_quad_walk(const float* __restrict__ disp /* dy,dx,dz planes */, const Geom& geom, int x0, int y0, float s, const Emit& emit) {
    constexpr int MSK = N - 1;
    float xc_m1[4], xc_0[4], zc_m1[4], zc_0[4], ddx_0[4], ddz_0[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) xc_m1[j] = xc_0[j] = zc_m1[j] = zc_0[j] = ddx_0[j] = ddz_0[j] = 0.f;
    JacRowIn nxt = jac_row_load(disp, geom, x0, (y0 - 1) & MSK, false);
#pragma unroll
    for (int i = -1; i <= RY; ++i) {              // row r = y0 + i ; emits output row r - 1 once i >= 1
        const JacRowIn in = nxt;
        if (i < RY) nxt = jac_row_load(disp, geom, x0, (y0 + i + 1) & MSK, i + 1 < RY);
        const float4 a = in.a, b = in.b;
        const float xc[4] = {a.x, a.y, a.z, a.w}, zc[4] = {b.x, b.y, b.z, b.w};
        const float ddx[4] = {a.y - in.al, a.z - a.x, a.w - a.y, in.ar - a.z};
        const float ddz[4] = {b.y - in.bl, b.z - b.x, b.w - b.y, in.br - b.z};
        if (i >= 1) {
            float J[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                J[j] = fmaf(-s, ddx_0[j], 1.0f) * fmaf(-s, zc[j] - zc_m1[j], 1.0f) - (s * (xc[j] - xc_m1[j])) * (s * ddz_0[j]);
            emit(y0 + i - 1, make_float4(J[0], J[1], J[2], J[3]));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            xc_m1[j] = xc_0[j]; xc_0[j] = xc[j]; zc_m1[j] = zc_0[j]; zc_0[j] = zc[j]; ddx_0[j] = ddx[j]; ddz_0[j] = ddz[j];
        }
    }
}

// Plain (un-staged) emit: each thread stores its own four normals; used by the CPU emulator.
template <bool JAC>
struct EmitDirect {
    float4* normal;
    float* jac;
    size_t ostride;   // elements between output rows
    int xout;         // first of the four output columns
    OW_HD void operator()(int y, const float4 (&n)[4], float4 J) const {
        float4* d = normal + (size_t)y * ostride + xout;
#pragma unroll
        for (int j = 0; j < 4; ++j) d[j] = n[j];
        if (JAC) *reinterpret_cast<float4*>(jac + (size_t)y * ostride + xout) = J;
    }
};

}  // namespace ow
