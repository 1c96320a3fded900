// ow_emu.cu — CPU emulation of the frame kernels' index logic (TEST INFRASTRUCTURE, never a product path).
//
// Compiles the very same per-thread phase functions the CUDA kernels run (csrc/ow_kernels.cuh) as host code
// and executes them thread by thread, phase by phase, CTA by CTA. It exists so that index-algebra mistakes
// (digit maps, Hermitian mirroring, paddings) are caught on the CPU-only build box, and to count
// shared-memory bank conflicts of each access pattern. Nothing in the product loads this library.
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

#include "../../fft-ocean-waves_b200/csrc/ow_config.cuh"
#include "../../fft-ocean-waves_b200/csrc/ow_kernels.cuh"

using namespace ow;

namespace {

// Records every shared-memory access of one phase: seq[tid] = list of element addresses in program order.
struct Recorder {
    std::vector<std::vector<int>> seq;
    long requests = 0, wavefronts = 0;
    void begin(int nthreads) { seq.assign(nthreads, {}); }
    // 8-byte accesses are served per half-warp; a bank pair is (element address mod 16).
    void end() {
        const int nt = (int)seq.size();
        for (int h = 0; h < nt; h += 16) {
            size_t len = 0;
            for (int t = h; t < h + 16 && t < nt; ++t) len = std::max(len, seq[t].size());
            for (size_t i = 0; i < len; ++i) {
                std::map<int, std::vector<int>> banks;
                for (int t = h; t < h + 16 && t < nt; ++t) {
                    if (i >= seq[t].size()) continue;
                    auto& v = banks[seq[t][i] & 15];
                    bool dup = false;
                    for (int a : v) dup |= (a == seq[t][i]);
                    if (!dup) v.push_back(seq[t][i]);
                }
                size_t worst = 0;
                for (auto& kv : banks) worst = std::max(worst, kv.second.size());
                if (worst) { requests += 1; wavefronts += (long)worst; }
            }
        }
    }
};

struct SmemEmu {
    float2* p;
    std::vector<int>* log;
    OW_HD float2 ld(int i) const {
#ifndef __CUDA_ARCH__
        log->push_back(i);
#endif
        return p[i];
    }
    OW_HD void st(int i, float2 v) const {
#ifndef __CUDA_ARCH__
        log->push_back(i);
#endif
        p[i] = v;
    }
};

// Host restatement of ow_fold_kernel (ow_init_kernels.cu): hp[p][u] = fold_pair(h0[p][u], h0[N-p][(N-u) mod N]), followed by the
// (w, 1/|k|) table of the same pair texels (second half of the block, hp_block_f4).
// sub_A > 0: rows stored sub-line-major (subline_index), as the contexts that run the N = sub_A * B line decomposition keep them.
template <int N>
void fold_full(const std::vector<float4>& h0, std::vector<float4>& hp, std::vector<float4>& nyq, float L, int sub_A = 0) {
    hp.assign(hp_block_f4(N / 2, N), make_float4(0, 0, 0, 0));
    nyq.assign(N / 2, make_float4(0, 0, 0, 0));
    float2* wk = reinterpret_cast<float2*>(hp.data() + (size_t)(N / 2) * N);
    const float pi = 3.1415926535897932384626433832795f;
    auto kof = [&](int i) { return (2.0f * pi * ((float)i - (float)N / 2.0f)) / L; };
    for (int p = 1; p < N / 2; ++p)
        for (int u = 0; u < N; ++u) {
            const float4 A = h0[(size_t)p * N + u], B = h0[(size_t)(N - p) * N + ((N - u) & (N - 1))];
            hp[(size_t)p * N + (sub_A > 0 ? subline_index(u, sub_A, N) : u)] = fold_pair(A, B);
            if (u == 0) nyq[p] = fold_pair_nyq(A, B);
            if (use_wk(N)) wk[(size_t)p * N + u] = dispersion_of(kof(u), kof(p));
        }
}

struct Stats {
    long req[6] = {0, 0, 0, 0, 0, 0}, wf[6] = {0, 0, 0, 0, 0, 0};   // row phase 0..2, col phase 0..2
};

// Row kernel of the CTAs [blk0, blk0 + nblk): pairs p = pbase + blk * PAIRS + g (pbase = first pair of a slab, 0 for the full grid).
template <int N, class Rows, class Sink>
void emu_rows(const Rows& rows, const float* ktab, float t, const Sink& sink, int pbase, int npairs, Stats& st) {
    using C = Cfg<N>;
    using P = typename C::Row;
    constexpr int PAIRS = C::ROW_PAIRS, NT = P::T * PAIRS;
    std::vector<float2> smem((size_t)PAIRS * 3 * P::LINE);
    Recorder rec;
    for (int blk = 0; blk < npairs / PAIRS; ++blk) {
        for (int phase = 0; phase < 3; ++phase) {
            rec.begin(NT);
            for (int tid = 0; tid < NT; ++tid) {
                const int ft = tid % P::T, g = tid / P::T, p = pbase + blk * PAIRS + g;
                const SmemEmu sm{smem.data() + (size_t)g * 3 * P::LINE, &rec.seq[tid]};
                if (phase == 0) row_phase0<P, false>(sm, ft, p, rows, ktab, t);
                if (phase == 1) row_phase1<P>(sm, ft);
                if (phase == 2) row_phase2<P>(sm, ft, p, sink);
            }
            if (blk < 2) { rec.requests = rec.wavefronts = 0; rec.end(); st.req[phase] += rec.requests; st.wf[phase] += rec.wavefronts; }
        }
    }
}

// Column kernel over `ncols` columns: src0/dst0 = channel-0 bases, channel f at src0 + f*src_chan, dst0 + f*dst_chan.
template <int N, class Geom>
void emu_cols(const float2* src0, size_t src_chan, float* dst0, size_t dst_chan, int ncols, const Geom& geom, Stats& st) {
    using C = Cfg<N>;
    using P = typename C::Col;
    constexpr int G = C::COL_G, NT = P::T * G;
    using LY = ColLayout<P, G>;
    std::vector<float2> smem((size_t)G * LY::SJ);
    const float scale = 0.5f / ((float)N * (float)N);
    Recorder rec;
    for (int f = 0; f < 3; ++f)
        for (int blk = 0; blk < ncols / (2 * G); ++blk) {
            for (int phase = 0; phase < 3; ++phase) {
                rec.begin(NT);
                for (int tid = 0; tid < NT; ++tid) {
                    const int job = tid % G, ft = tid / G, x = 2 * (blk * G + job);
                    const SmemEmu sm{smem.data(), &rec.seq[tid]};
                    const int base = job * LY::SJ;
                    const float2* src = src0 + (size_t)f * src_chan + x;
                    float* dst = dst0 + (size_t)f * dst_chan + x;
                    if (phase == 0) for (int j = ft; j < P::M / 2; j += P::T) col_phase0<P>(sm, base, j, src, geom);
                    if (phase == 1) col_phase1<P>(sm, base, ft);
                    if (phase == 2) col_phase2<P>(sm, base, ft, dst, scale, geom);
                }
                if (f == 0 && blk < 2) { rec.requests = rec.wavefronts = 0; rec.end(); st.req[3 + phase] += rec.requests; st.wf[3 + phase] += rec.wavefronts; }
            }
        }
}

// Normal/Jacobian walk over output columns [0, ncols); xin0 = index of output column 0 inside a row of `disp`.
template <int N, class Geom>
void emu_normals(const float* disp, const Geom& geom, int xin0, int ncols, float4* normal, float* jac, size_t ostride, float lambda, float L) {
    constexpr int RY = 8;
    const float s = lambda * ((float)N / (2.0f * L));
    for (int y0 = 0; y0 < N; y0 += RY)
        for (int x0 = 0; x0 < ncols; x0 += 4) {
            if (jac) normal_quad_walk<N, RY, true>(disp, geom, xin0 + x0, y0, s, EmitDirect<true>{normal, jac, ostride, x0});
            else normal_quad_walk<N, RY, false>(disp, geom, xin0 + x0, y0, 0.f, EmitDirect<false>{normal, nullptr, ostride, x0});
        }
}

template <int N>
int emu_frame_n(const float* h0k, const float* h0minusk, float L, float t, float lambda, float* inter_out, float* disp,
                float* normal, float* jac, long* stats) {
    std::vector<float4> h0((size_t)N * N);
    for (size_t i = 0; i < (size_t)N * N; ++i) h0[i] = make_float4(h0k[2 * i], h0k[2 * i + 1], h0minusk[2 * i], h0minusk[2 * i + 1]);
    std::vector<float> ktab(N);
    const float pi = 3.1415926535897932384626433832795f;
    for (int i = 0; i < N; ++i) ktab[i] = (2.0f * pi * ((float)i - (float)N / 2.0f)) / L;
    std::vector<float2> inter((size_t)3 * (N / 2) * N);
    Stats st;
    std::vector<float4> hp, nyq;
    fold_full<N>(h0, hp, nyq, L);
    emu_rows<N>(FullRows<N>{h0.data(), hp.data(), nyq.data()}, ktab.data(), t, FullSink<N>{inter.data()}, 0, N / 2, st);
    emu_cols<N>(inter.data(), (size_t)(N / 2) * N, disp, (size_t)N * N, N, FullColGeom<N>{}, st);
    emu_normals<N>(disp, FullNrmGeom<N>{}, 0, N, reinterpret_cast<float4*>(normal), jac, N, lambda, L);
    if (inter_out) std::memcpy(inter_out, inter.data(), inter.size() * sizeof(float2));
    if (stats) for (int i = 0; i < 6; ++i) { stats[2 * i] = st.req[i]; stats[2 * i + 1] = st.wf[i]; }
    return 0;
}

// ow_col2_kernel thread by thread: 16-column tiles; dy tiles keep their heights in the lines and run the normal-map epilogue for
// their three interior quads; the seam quads are walked afterwards from the stored heights (on the GPU: by the last of the two
// neighbouring tiles to finish). staged = true routes stage 0 through a host copy of the TMA staging buffer, filled box by box the
// way ow_col2_kernel's issue() asks the copy engine to (ColStage), so the box/row/extra-row algebra is checked here too.
// st.req/wf[3..5] = the three FFT phases of the dy tiles, stats6 = the stencil phase out of shared memory, stats6[2..3] = the
// staging reads of stage 0.
template <int N>
void emu_cols2(const float2* inter, float* disp, float4* normal, bool staged, bool fuse, Stats& st, long* stats6) {
    using C = Cfg<N>;
    using P = typename C::Col;
    constexpr int G = C::COL_G, NT_ = P::T * G, HP = N / 2, RY = C::NRM_RY, NTILES = N / (2 * G), H = P::R0 / 2;
    using LY = ColLayout<P, G>;
    using CS = ColStage<P, G>;
    std::vector<float2> smem((size_t)G * LY::SJ);
    std::vector<float4> staging(CS::TOTAL_F4);
    const float scale = 0.5f / ((float)N * (float)N);
    const FullColGeom<N> geom{};
    Recorder rec;
    auto inter_row = [&](int f, long row, int tile, float4* dstrow) {     // G float4 = the tile's 16 columns of one intermediate row
        const long total_rows = 3L * HP;
        const long gr = (long)f * HP + row;
        for (int g = 0; g < G; ++g) {
            if (gr < 0 || gr >= total_rows) { dstrow[g] = make_float4(0, 0, 0, 0); continue; }           // TMA zero-fills out-of-bounds rows
            const float2* p = inter + (size_t)gr * N + 2 * (tile * G + g);
            dstrow[g] = make_float4(p[0].x, p[0].y, p[1].x, p[1].y);
        }
    };
    for (int blk = 0; blk < 3 * NTILES; ++blk) {
        const int f = blk / NTILES, tile = blk % NTILES;
        const bool dy_tile = f == 0 && fuse;
        // phase 0
        if (staged) {
            for (int s = 0; s < CS::STEPS; ++s) {
                for (int i = 0; i < H; ++i)
                    for (int r = 0; r < P::T; ++r) {
                        inter_row(f, CS::fwd_row0(i, s) + r, tile, staging.data() + ((2 * i + 0) * P::T + r) * G);
                        inter_row(f, CS::mir_row0(i, s) + r, tile, staging.data() + ((2 * i + 1) * P::T + r) * G);
                    }
                if (s == 0) for (int i = 0; i < H; ++i) inter_row(f, CS::extra_row(i), tile, staging.data() + CS::EXTRA_F4 + i * G);
                rec.begin(NT_);
                for (int tid = 0; tid < NT_; ++tid) {
                    const int job = tid % G, ft = tid / G;
                    const SmemEmu sm{smem.data(), &rec.seq[tid]};
                    float4 la[H], lb[H];
                    col_stage_read<P, G>(staging.data(), job, ft, s, la, lb);
                    col_phase0_math<P>(sm, job * LY::SJ, s * P::T + ft, la, lb);
                }
                if (blk < 2) { rec.requests = rec.wavefronts = 0; rec.end(); st.req[3] += rec.requests; st.wf[3] += rec.wavefronts; }
            }
        }
        for (int phase = staged ? 1 : 0; phase < 4; ++phase) {
            if (phase == 3 && !dy_tile) break;
            rec.begin(NT_);
            for (int tid = 0; tid < NT_; ++tid) {
                const int job = tid % G, ft = tid / G, base = job * LY::SJ;
                const int pair = tile * G + job;
                const SmemEmu sm{smem.data(), &rec.seq[tid]};
                const float2* src = inter + (size_t)f * HP * N + 2 * pair;
                float* dst = disp + (size_t)f * N * N + 2 * pair;
                if (phase == 0) for (int j = ft; j < P::M / 2; j += P::T) col_phase0<P>(sm, base, j, src, geom);
                if (phase == 1) col_phase1<P>(sm, base, ft);
                if (phase == 2) { if (dy_tile) col_phase2_keep<P>(sm, base, ft, dst, scale, geom, true); else col_phase2<P>(sm, base, ft, dst, scale, geom); }
                if (phase == 3) col_normals_phase<P, RY>(sm, tid, NT_, LY::SJ, 16 * tile + 2, normal);
            }
            if (blk < 2) {
                rec.requests = rec.wavefronts = 0; rec.end();
                if (phase < 3) { st.req[3 + phase] += rec.requests; st.wf[3 + phase] += rec.wavefronts; }
                else { stats6[0] += rec.requests; stats6[1] += rec.wavefronts; }
            }
        }
    }
    if (fuse)
        for (int k = 0; k < NTILES; ++k)
            for (int tid = 0; tid < NT_; ++tid) col_seam_phase<N, 4>(disp, normal, k, tid, NT_);
}

template <int N>
int emu_frame_fused_n(const float* h0k, const float* h0minusk, float L, float t, float lambda, int staged, float* disp, float* normal, float* jac,
                      long* stats) {
    std::vector<float4> h0((size_t)N * N), hp, nyq;
    for (size_t i = 0; i < (size_t)N * N; ++i) h0[i] = make_float4(h0k[2 * i], h0k[2 * i + 1], h0minusk[2 * i], h0minusk[2 * i + 1]);
    fold_full<N>(h0, hp, nyq, L);
    std::vector<float> ktab(N);
    const float pi = 3.1415926535897932384626433832795f;
    for (int i = 0; i < N; ++i) ktab[i] = (2.0f * pi * ((float)i - (float)N / 2.0f)) / L;
    std::vector<float2> inter((size_t)3 * (N / 2) * N);
    Stats st;
    long s6[2] = {0, 0};
    emu_rows<N>(FullRows<N>{h0.data(), hp.data(), nyq.data()}, ktab.data(), t, FullSink<N>{inter.data()}, 0, N / 2, st);
    emu_cols2<N>(inter.data(), disp, reinterpret_cast<float4*>(normal), staged != 0, true, st, s6);
    if (jac) {       // ow_jac_kernel: the Jacobian alone, from the stored dx/dz planes
        const float s = lambda * ((float)N / (2.0f * L));
        for (int y0 = 0; y0 < N; y0 += 8)
            for (int x0 = 0; x0 < N; x0 += 4)
                jac_quad_walk<N, 8>(disp, FullNrmGeom<N>{}, x0, y0, s,
                                    [&](int y, float4 J) { *reinterpret_cast<float4*>(jac + (size_t)y * N + x0) = J; });
    }
    if (stats) { for (int i = 0; i < 6; ++i) { stats[2 * i] = st.req[i]; stats[2 * i + 1] = st.wf[i]; } stats[12] = s6[0]; stats[13] = s6[1]; }
    return 0;
}

// The slab-decomposed frame with `world` emulated ranks run one after the other: every rank's row
// kernel stores through SlabSink straight into the owners' receive buffers (what the peer-store mode does over
// NVLink), then every rank runs the column + normal kernels on its padded column slab. Outputs are re-assembled into
// full [N][N] images so the test can compare them bit for bit with emu_frame.
template <int N>
int emu_slab_frame_n(int world, const float* h0k, const float* h0minusk, float L, float t, float lambda, float* disp,
                     float* normal, float* jac) {
    if (world < 1 || world > kMaxWorld || (N / 2) % world) return -2;
    const int PL = N / 2 / world, XL = N / world, XH = XL + 2 * kHalo;
    if (PL % Cfg<N>::ROW_PAIRS || XH % (2 * Cfg<N>::COL_G) || XL % 128) return -3;
    int shift = 0;
    while ((1 << shift) < XL) ++shift;
    std::vector<float> ktab(N);
    const float pi = 3.1415926535897932384626433832795f;
    for (int i = 0; i < N; ++i) ktab[i] = (2.0f * pi * ((float)i - (float)N / 2.0f)) / L;
    std::vector<std::vector<float2>> recv(world, std::vector<float2>((size_t)(N / 2) * 3 * XH));
    Stats st;
    for (int r = 0; r < world; ++r) {
        SlabRows<N> rows{nullptr, nullptr, nullptr, r * PL, PL};
        std::vector<float4> h0((size_t)2 * PL * N);
        for (int v = 0; v < N; ++v) {
            const int pair = (v < N / 2) ? v : ((N - v) & (N / 2 - 1));
            if (pair < r * PL || pair >= (r + 1) * PL) continue;
            float4* dst = h0.data() + (size_t)rows.local(v) * N;
            for (int u = 0; u < N; ++u) {
                const size_t i = (size_t)v * N + u;
                dst[u] = make_float4(h0k[2 * i], h0k[2 * i + 1], h0minusk[2 * i], h0minusk[2 * i + 1]);
            }
        }
        rows.h0 = h0.data();
        // ow_fold_kernel on the slab-local layout: primary row pl, mirror row PL + pl
        std::vector<float4> hp(hp_block_f4(PL, N), make_float4(0, 0, 0, 0)), nyq(PL, make_float4(0, 0, 0, 0));
        float2* wk = reinterpret_cast<float2*>(hp.data() + (size_t)PL * N);
        for (int pl = 0; pl < PL; ++pl) {
            if (r * PL + pl == 0) continue;
            for (int u = 0; u < N; ++u) {
                const float4 A = h0[(size_t)pl * N + u], B = h0[(size_t)(PL + pl) * N + ((N - u) & (N - 1))];
                hp[(size_t)pl * N + u] = fold_pair(A, B);
                if (u == 0) nyq[pl] = fold_pair_nyq(A, B);
                if (use_wk(N)) wk[(size_t)pl * N + u] = dispersion_of(ktab[u], ktab[r * PL + pl]);
            }
        }
        rows.hp = hp.data(); rows.nyq = nyq.data();
        SlabSink<N> sink{};
        for (int h = 0; h < world; ++h) sink.base[h] = recv[h].data() + (size_t)r * PL * 3 * XH;
        sink.world = world; sink.p0 = r * PL; sink.XL = XL; sink.XH = XH; sink.xl_shift = shift;
        emu_rows<N>(rows, ktab.data(), t, sink, r * PL, PL, st);
    }
    for (int r = 0; r < world; ++r) {
        std::vector<float> dloc((size_t)3 * N * XH);
        emu_cols<N>(recv[r].data(), (size_t)XH, dloc.data(), (size_t)N * XH, XH, SlabColGeom{(size_t)3 * XH, (size_t)XH}, st);
        std::vector<float4> nloc((size_t)N * XL);
        std::vector<float> jloc(jac ? (size_t)N * XL : 0);
        emu_normals<N>(dloc.data(), SlabNrmGeom{(size_t)XH, (size_t)N * XH}, kHalo, XL, nloc.data(), jac ? jloc.data() : nullptr, XL, lambda, L);
        for (int f = 0; f < 3; ++f)
            for (int y = 0; y < N; ++y)
                std::memcpy(disp + ((size_t)f * N + y) * N + (size_t)r * XL, dloc.data() + ((size_t)f * N + y) * XH + kHalo, XL * sizeof(float));
        for (int y = 0; y < N; ++y) {
            std::memcpy(normal + ((size_t)y * N + (size_t)r * XL) * 4, nloc.data() + (size_t)y * XL, XL * sizeof(float4));
            if (jac) std::memcpy(jac + (size_t)y * N + (size_t)r * XL, jloc.data() + (size_t)y * XL, XL * sizeof(float));
        }
    }
    return 0;
}

}  // namespace

extern "C" int emu_frame(int N, const float* h0k, const float* h0minusk, float L, float t, float lambda, float* inter_out,
                         float* disp, float* normal, float* jac, long* stats) {
    switch (N) {
        case 128: return emu_frame_n<128>(h0k, h0minusk, L, t, lambda, inter_out, disp, normal, jac, stats);
        case 256: return emu_frame_n<256>(h0k, h0minusk, L, t, lambda, inter_out, disp, normal, jac, stats);
        case 512: return emu_frame_n<512>(h0k, h0minusk, L, t, lambda, inter_out, disp, normal, jac, stats);
        case 1024: return emu_frame_n<1024>(h0k, h0minusk, L, t, lambda, inter_out, disp, normal, jac, stats);
        case 2048: return emu_frame_n<2048>(h0k, h0minusk, L, t, lambda, inter_out, disp, normal, jac, stats);
        case 4096: return emu_frame_n<4096>(h0k, h0minusk, L, t, lambda, inter_out, disp, normal, jac, stats);
    }
    return -1;
}

// The N = A*B line decomposition (ow_big_kernels.cu) emulated for a small grid: sub-line plan Cfg<B>, A sub-lines per line.
template <int B, int A>
int emu_big_frame_n(const float* h0k, const float* h0minusk, float L, float t, float lambda, float* disp, float* normal, float* jac) {
    constexpr int N = A * B;
    using C = Cfg<B>;
    using PR = typename C::Row;
    using PK = typename C::Col;
    constexpr int G = C::COL_G;
    using LY = ColLayout<PK, G>;
    std::vector<float4> h0((size_t)N * N), hp, nyq;
    for (size_t i = 0; i < (size_t)N * N; ++i) h0[i] = make_float4(h0k[2 * i], h0k[2 * i + 1], h0minusk[2 * i], h0minusk[2 * i + 1]);
    fold_full<N>(h0, hp, nyq, L, A);
    std::vector<float> ktab(N), ktab_sub(N);
    const float pi = 3.1415926535897932384626433832795f;
    for (int i = 0; i < N; ++i) ktab[i] = (2.0f * pi * ((float)i - (float)N / 2.0f)) / L;
    for (int i = 0; i < N; ++i) ktab_sub[subline_index(i, A, N)] = ktab[i];
    std::vector<float2> inter((size_t)3 * (N / 2) * N), scratch((size_t)3 * (N / 2) * N);
    const FullRows<N> rows{h0.data(), hp.data(), nyq.data()};
    const FullSink<N> sink{inter.data()};
    std::vector<int> dummy;
    // rows: one group per (pair, sub-line a) -> scratch, then the post kernel (one thread per (pair, channel, kb))
    {
        std::vector<float2> smem((size_t)3 * PR::LINE);
        const SmemEmu sm{smem.data(), &dummy};
        for (int p = 0; p < N / 2; ++p)
            for (int a = 0; a < A; ++a) {
                for (int ft = 0; ft < PR::T; ++ft) { dummy.clear(); bigrow_phase0<PR, A, false>(sm, ft, p, a, rows, ktab.data(), ktab_sub.data(), t); }
                for (int ft = 0; ft < PR::T; ++ft) { dummy.clear(); row_phase1<PR>(sm, ft); }
                for (int ft = 0; ft < PR::T; ++ft) { dummy.clear(); bigrow_phase2<PR, A>(sm, ft, a, scratch.data() + (size_t)p * 3 * N); }
            }
    }
    for (int p = 0; p < N / 2; ++p)
        for (int c = 0; c < 3; ++c)
            for (int kb = 0; kb < B; ++kb) bigrow_post<B, A>(scratch.data() + (size_t)p * 3 * N, c, p, kb, sink);
    // columns: one CTA per (channel, tile, sub-line a) -> scratch, then the post kernel (one thread per (channel, pair, kb))
    const int npairs = N / 2;
    const float scale = 0.5f / ((float)N * (float)N);
    {
        std::vector<float2> smem((size_t)G * LY::SJ);
        const SmemEmu sm{smem.data(), &dummy};
        const FullColGeom<N> geom{};
        for (int f = 0; f < 3; ++f)
            for (int tile = 0; tile < N / (2 * G); ++tile)
                for (int a = 0; a < A; ++a)
                    for (int phase = 0; phase < 3; ++phase)
                        for (int tid = 0; tid < PK::T * G; ++tid) {
                            dummy.clear();
                            const int job = tid % G, ft = tid / G, pair = tile * G + job, base = job * LY::SJ;
                            if (phase == 0) bigcol_phase0<PK, A>(sm, base, ft, a, inter.data() + (size_t)f * (N / 2) * N + 2 * pair, geom);
                            if (phase == 1) col_phase1<PK>(sm, base, ft);
                            if (phase == 2) bigcol_phase2<PK>(sm, base, ft, scratch.data() + (size_t)f * N * npairs + (size_t)a * B * npairs + pair, (size_t)npairs);
                        }
    }
    for (int c = 0; c < 3; ++c)
        for (int pair = 0; pair < npairs; ++pair)
            for (int kb = 0; kb < B; ++kb)
                bigcol_post<B, A>(scratch.data() + (size_t)c * N * npairs + pair, (size_t)npairs, kb, disp + (size_t)c * N * N + 2 * pair, (size_t)N, scale);
    emu_normals<N>(disp, FullNrmGeom<N>{}, 0, N, reinterpret_cast<float4*>(normal), jac, N, lambda, L);
    return 0;
}

extern "C" int emu_frame_fused(int N, const float* h0k, const float* h0minusk, float L, float t, float lambda, int staged, float* disp,
                               float* normal, float* jac, long* stats) {
    switch (N) {
        case 256: return emu_frame_fused_n<256>(h0k, h0minusk, L, t, lambda, staged, disp, normal, jac, stats);
        case 512: return emu_frame_fused_n<512>(h0k, h0minusk, L, t, lambda, staged, disp, normal, jac, stats);
        case 1024: return emu_frame_fused_n<1024>(h0k, h0minusk, L, t, lambda, staged, disp, normal, jac, stats);
        case 2048: return emu_frame_fused_n<2048>(h0k, h0minusk, L, t, lambda, staged, disp, normal, jac, stats);
    }
    return -1;
}

extern "C" int emu_big_frame(int N, int A, const float* h0k, const float* h0minusk, float L, float t, float lambda, float* disp,
                             float* normal, float* jac) {
    if (N == 1024 && A == 4) return emu_big_frame_n<256, 4>(h0k, h0minusk, L, t, lambda, disp, normal, jac);
    if (N == 512 && A == 2) return emu_big_frame_n<256, 2>(h0k, h0minusk, L, t, lambda, disp, normal, jac);
    if (N == 2048 && A == 8) return emu_big_frame_n<256, 8>(h0k, h0minusk, L, t, lambda, disp, normal, jac);
    return -1;
}

extern "C" int emu_slab_frame(int N, int world, const float* h0k, const float* h0minusk, float L, float t, float lambda,
                              float* disp, float* normal, float* jac) {
    switch (N) {
        case 256: return emu_slab_frame_n<256>(world, h0k, h0minusk, L, t, lambda, disp, normal, jac);
        case 512: return emu_slab_frame_n<512>(world, h0k, h0minusk, L, t, lambda, disp, normal, jac);
        case 1024: return emu_slab_frame_n<1024>(world, h0k, h0minusk, L, t, lambda, disp, normal, jac);
        case 2048: return emu_slab_frame_n<2048>(world, h0k, h0minusk, L, t, lambda, disp, normal, jac);
        case 4096: return emu_slab_frame_n<4096>(world, h0k, h0minusk, L, t, lambda, disp, normal, jac);
    }
    return -1;
}

// In-register DFT self-test: out[k] = dft<R>(in), for R in {2,4,8,16}.
extern "C" int emu_dft(int R, const float* in, float* out) {
    float2 v2[2], v4[4], v8[8], v16[16];
    auto load = [&](float2* v) { for (int i = 0; i < R; ++i) v[i] = make_float2(in[2 * i], in[2 * i + 1]); };
    auto store = [&](float2* v) { for (int i = 0; i < R; ++i) { out[2 * i] = v[i].x; out[2 * i + 1] = v[i].y; } };
    switch (R) {
        case 2: load(v2); Dft<2>::run(v2); store(v2); return 0;
        case 4: load(v4); Dft<4>::run(v4); store(v4); return 0;
        case 8: load(v8); Dft<8>::run(v8); store(v8); return 0;
        case 16: load(v16); Dft<16>::run(v16); store(v16); return 0;
    }
    return -1;
}
