// ow_frame_kernels.cuh — __global__ wrappers of the per-frame kernels (sm_100a). Bodies live in ow_kernels.cuh
// (shared with the CPU emulator used by the tests); the per-N choices of the template parameters are in
// ow_config.cuh. Included by ow_frame_kernels.cu (the library) and by tools/tune (the variant sweeper).
#pragma once
#include "ow_internal.h"
#include "ow_kernels.cuh"
#include "ow_config.cuh"
#include "ow_async.cuh"
#include <cooperative_groups.h>

namespace ow {

// Barrier between the phases of one line group. A row-pair group of 32 threads is exactly one warp (N <= 512): its lines are
// private to the warp, so a warp barrier orders the shared-memory traffic and the CTA's other groups are not held up.
template <int T>
__device__ __forceinline__ void group_sync() {
    if (T == 32) __syncwarp();
    else __syncthreads();
}

// Per-CTA timeline for tools/tune (never defined in the library build): thread 0 stamps clock64 + globaltimer + smid at
// the phase boundaries into g_ow_trace[cta][8].
#ifdef OW_TRACE
__device__ unsigned long long* g_ow_trace;
__device__ __forceinline__ void ow_stamp(int cta, int i) {
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        unsigned long long gt; unsigned sm;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
        asm volatile("mov.u32 %0, %smid;" : "=r"(sm));
        g_ow_trace[(size_t)cta * 8 + i] = (i == 0) ? ((gt << 8) | sm) : (unsigned long long)clock64();
        if (i == 1) g_ow_trace[(size_t)cta * 8 + 7] = gt;
    }
}
#define OW_STAMP(cta, i) ow_stamp(cta, i)
#else
#define OW_STAMP(cta, i)
#endif

// ---------------------------------------------------------------------------------------------------
template <class P, int PAIRS, int MINB, bool FAST>
__global__ void __launch_bounds__(P::T* PAIRS, MINB) ow_row_kernel(FrameBuffers fb, SlotTable tab) {
    extern __shared__ __align__(16) float2 smem[];
    constexpr int N = P::N;
    const int ft = threadIdx.x % P::T, g = threadIdx.x / P::T;
    const int p = blockIdx.x * PAIRS + g;
    const int e = blockIdx.y;
    const int cascade = tab.cascade[e];
    const float t = tab.time[e];
    const int slot = tab.slot[e];
    const SmemDirect sm{smem + (size_t)g * 3 * P::LINE};
    const FullRows<N> rows{fb.h0 + (size_t)cascade * N * N, fb.hp + (size_t)cascade * hp_block_f4(N / 2, N), fb.nyq + (size_t)cascade * (N / 2)};
    const float* ktab = fb.ktab + (size_t)cascade * N;
    float2* inter = fb.inter + (size_t)slot * 3 * (N / 2) * N;
    OW_STAMP(blockIdx.x, 0); OW_STAMP(blockIdx.x, 1);
    row_phase0<P, FAST>(sm, ft, p, rows, ktab, t);
    OW_STAMP(blockIdx.x, 2);
    group_sync<P::T>();
    OW_STAMP(blockIdx.x, 3);
    row_phase1<P>(sm, ft);
    OW_STAMP(blockIdx.x, 4);
    group_sync<P::T>();
    OW_STAMP(blockIdx.x, 5);
    row_phase2<P>(sm, ft, p, FullSink<N>{inter});
    OW_STAMP(blockIdx.x, 6);
}

// ---------------------------------------------------------------------------------------------------
// Persistent, software-pipelined row kernel. The per-CTA timeline of ow_row_kernel (tools/tune/trace.cu) shows a CTA
// spending two thirds of its life in phase 0, and a third of that waiting on the h0 loads of each stage-0 batch
// (issue 16 LDG.128 -> wait a DRAM round trip -> compute, C0 times, with nothing in flight in between). Here a CTA
// walks over many row pairs ("items", (slot entry, pair) flattened) and always has the NEXT batch's loads in flight:
// they are issued right after the spectrum of the current batch has consumed its registers, so the DRAM round trip
// overlaps the stage-0 butterflies of this batch or, across items, all of stages 1 and 2 and their stores.
// Pair 0 (rows 0 and N/2, two batches per butterfly) takes the plain path.
// ---------------------------------------------------------------------------------------------------
template <int N>
struct RowItem {
    FullRows<N> rows;
    const float4* prow;   // folded pair row of this item
    const float2* wrow;   // its (w, 1/|k|) row
    const float* ktab;
    float2* inter;
    float t;
    int p;
};

template <class P, int PAIRS, int MINB, bool FAST>
__global__ void __launch_bounds__(P::T* PAIRS, MINB) ow_row_pipe_kernel(FrameBuffers fb, SlotTable tab, int n_cta_items) {
    extern __shared__ __align__(16) float2 smem[];
    constexpr int N = P::N, R0 = P::R0, HP = N / 2, C0 = P::C0;
    static_assert(P::M % P::T == 0, "pipelined row kernel needs whole stage-0 batches");
    const int ft = threadIdx.x % P::T, g = threadIdx.x / P::T;
    const SmemDirect sm{smem + (size_t)g * 3 * P::LINE};

    auto item_of = [&](int ci) {
        const int item = ci * PAIRS + g, e = item / HP;
        RowItem<N> it;
        it.p = item - e * HP;
        const int cascade = tab.cascade[e];
        it.rows = FullRows<N>{fb.h0 + (size_t)cascade * N * N, fb.hp + (size_t)cascade * hp_block_f4(HP, N), fb.nyq + (size_t)cascade * HP};
        it.prow = it.rows.pair_row(it.p);
        it.wrow = it.rows.wk_row(it.p);
        it.ktab = fb.ktab + (size_t)cascade * N;
        it.inter = fb.inter + (size_t)tab.slot[e] * 3 * HP * N;
        it.t = tab.time[e];
        return it;
    };
    auto issue = [&](const RowItem<N>& it, int c, FoldedPair (&fp)[R0]) {
        const int b = ft + P::T * c;
#pragma unroll
        for (int d0 = 0; d0 < R0; ++d0) fp[d0] = load_folded<kUseWk<N>>(it.prow, it.wrow, it.ktab, d0 * P::M + b);
    };

    FoldedPair nxt[R0];
    int ci = blockIdx.x;
    if (ci >= n_cta_items) return;
    RowItem<N> cur = item_of(ci);
    if (cur.p != 0) issue(cur, 0, nxt);
    for (; ci < n_cta_items; ci += gridDim.x) {
        const int cn = ci + gridDim.x;
        const bool has_next = cn < n_cta_items;
        RowItem<N> nx = cur;
        if (has_next) nx = item_of(cn);
        if (cur.p == 0) {
            row_phase0_pair0<P, FAST>(sm, ft, cur.rows, cur.ktab, cur.t);
            if (has_next && nx.p != 0) issue(nx, 0, nxt);
        } else {
            const float ky = OW_LDG(cur.ktab + cur.p);
#pragma unroll
            for (int c = 0; c < C0; ++c) {
                const int b = ft + P::T * c;
                float2 vy[R0], vx[R0], vz[R0];
#pragma unroll
                for (int d0 = 0; d0 < R0; ++d0) {
                    const Sym3 s = spectrum_folded<FAST, kUseWk<N>>(nxt[d0], ky, cur.t, (d0 == 0 && b == 0) ? cur.rows.nyq_of(cur.p) : nullptr);
                    vy[d0] = s.y; vx[d0] = s.x; vz[d0] = s.z;
                }
                if (c + 1 < C0) issue(cur, c + 1, nxt);                       // next batch of this pair
                else if (has_next && nx.p != 0) issue(nx, 0, nxt);            // first batch of the next pair
                float2 tw[R0];
                twiddle_powers<R0>(unit_root(b, N), tw);
                stage0_finish<P>(sm, 0 * P::LINE, b, vy, tw);
                stage0_finish<P>(sm, 1 * P::LINE, b, vx, tw);
                stage0_finish<P>(sm, 2 * P::LINE, b, vz, tw);
            }
        }
        __syncthreads();
        row_phase1<P>(sm, ft);
        __syncthreads();
        row_phase2<P>(sm, ft, cur.p, FullSink<N>{cur.inter});
        __syncthreads();          // the next pair's stage-0 stores reuse the lines
        cur = nx;
    }
}

// ---------------------------------------------------------------------------------------------------
// Persistent row kernel with bulk-async staging. Same walk over (slot entry, row pair) items as ow_row_pipe_kernel, but the
// folded spectrum row of the NEXT item (N float4, contiguous) is fetched by ONE cp.async.bulk per item into a per-group
// shared-memory buffer, signalled through the group's mbarrier: no registers hold data in flight, the copy is issued the moment
// stage 0 has consumed the previous row and lands during stages 1 and 2 and their stores. Groups (T threads, one row pair each)
// never wait for each other: warp barriers for T = 32, named barriers otherwise. Pair 0 (rows 0 and N/2) takes the literal
// out-of-line path from global memory.
// ---------------------------------------------------------------------------------------------------
template <class P>
struct RowBulkSmem {
    static constexpr size_t GROUP = (((size_t)P::N * 16 + (size_t)3 * P::LINE * 8) + 127) / 128 * 128;   // staged row + the three lines
};
template <class P, int PAIRS>
constexpr size_t row_bulk_smem() { return PAIRS * RowBulkSmem<P>::GROUP; }

template <int T, int PAIRS>
__device__ __forceinline__ void row_group_sync(int g) {
    if (T == 32) __syncwarp();
    else if (PAIRS == 1) __syncthreads();
    else named_barrier(1 + g, T);
}

template <class P, int PAIRS, int MINB, bool FAST>
__global__ void __launch_bounds__(P::T* PAIRS, MINB) ow_row_bulk_kernel(FrameBuffers fb, SlotTable tab, int n_cta_items) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bars[PAIRS];
    constexpr int N = P::N, R0 = P::R0, HP = N / 2, C0 = P::C0, T = P::T;
    static_assert(P::M % P::T == 0, "whole stage-0 batches");
    const int ft = threadIdx.x % T, g = threadIdx.x / T;
    unsigned char* mine = smem_raw + (size_t)g * RowBulkSmem<P>::GROUP;
    const float4* stg = reinterpret_cast<const float4*>(mine);
    const SmemDirect sm{reinterpret_cast<float2*>(mine + (size_t)N * 16)};
    uint64_t* bar = &bars[g];

    auto item_of = [&](int ci) {
        const int item = ci * PAIRS + g, e = item / HP;
        RowItem<N> it;
        it.p = item - e * HP;
        const int cascade = tab.cascade[e];
        it.rows = FullRows<N>{fb.h0 + (size_t)cascade * N * N, fb.hp + (size_t)cascade * hp_block_f4(HP, N), fb.nyq + (size_t)cascade * HP};
        it.prow = it.rows.pair_row(it.p);
        it.wrow = it.rows.wk_row(it.p);
        it.ktab = fb.ktab + (size_t)cascade * N;
        it.inter = fb.inter + (size_t)tab.slot[e] * 3 * HP * N;
        it.t = tab.time[e];
        return it;
    };
    auto issue = [&](const RowItem<N>& it) {      // one thread of the group
        mbar_arrive_expect_tx(bar, (unsigned)N * 16u);
        bulk_load(mine, it.prow, (unsigned)N * 16u, bar, l2_policy_evict_first());
    };

    int ci = blockIdx.x;
    if (ci >= n_cta_items) return;
    if (ft == 0) {
        mbar_init(bar, 1);
        mbar_init_fence();
    }
    __syncthreads();
    RowItem<N> cur = item_of(ci);
    if (cur.p != 0 && ft == 0) issue(cur);
    unsigned phase = 0;
    for (; ci < n_cta_items; ci += gridDim.x) {
        const int cn = ci + gridDim.x;
        const bool has_next = cn < n_cta_items;
        RowItem<N> nx = cur;
        if (has_next) nx = item_of(cn);
        if (cur.p == 0) {
            row_phase0_pair0<P, FAST>(sm, ft, cur.rows, cur.ktab, cur.t);
        } else {
            const float ky = OW_LDG(cur.ktab + cur.p);
            mbar_wait(bar, phase);
            phase ^= 1u;
#pragma unroll 1
            for (int c = 0; c < C0; ++c) {
                const int b = ft + T * c;
                float2 vy[R0], vx[R0], vz[R0];
                FoldedPair fp[R0];
#pragma unroll
                for (int d0 = 0; d0 < R0; ++d0) {
                    fp[d0].f = stg[d0 * P::M + b];
                    fp[d0].wk = kUseWk<N> ? OW_LDG(cur.wrow + d0 * P::M + b) : make_float2(0.f, 0.f);
                    fp[d0].kx = OW_LDG(cur.ktab + d0 * P::M + b);
                }
#pragma unroll
                for (int d0 = 0; d0 < R0; ++d0) {
                    const Sym3 s = spectrum_folded<FAST, kUseWk<N>>(fp[d0], ky, cur.t, (d0 == 0 && b == 0) ? cur.rows.nyq_of(cur.p) : nullptr);
                    vy[d0] = s.y; vx[d0] = s.x; vz[d0] = s.z;
                }
                float2 tw[R0];
                twiddle_powers<R0>(unit_root(b, N), tw);
                stage0_finish<P>(sm, 0 * P::LINE, b, vy, tw);
                stage0_finish<P>(sm, 1 * P::LINE, b, vx, tw);
                stage0_finish<P>(sm, 2 * P::LINE, b, vz, tw);
            }
        }
        row_group_sync<T, PAIRS>(g);      // stage-0 results visible; the staged row is consumed
        if (has_next && nx.p != 0 && ft == 0) issue(nx);
        row_phase1<P>(sm, ft);
        row_group_sync<T, PAIRS>(g);
        row_phase2<P>(sm, ft, cur.p, FullSink<N>{cur.inter});
        row_group_sync<T, PAIRS>(g);      // the next pair's stage-0 stores reuse the lines
        cur = nx;
    }
}

// Slab variant (one grid over several GPUs): this rank's row pairs [p0, p0 + PL), results stored straight into the
// column owners' receive buffers (peer mappings over NVLink) or into the local send buffer (see SlabSink).
template <class P, int PAIRS, int MINB, bool FAST>
__global__ void __launch_bounds__(P::T* PAIRS, MINB) ow_row_slab_kernel(SlabRows<P::N> rows, const float* __restrict__ ktab,
                                                                        SlabSink<P::N> sink, float t) {
    extern __shared__ __align__(16) float2 smem[];
    const int ft = threadIdx.x % P::T, g = threadIdx.x / P::T;
    const int p = rows.p0 + blockIdx.x * PAIRS + g;
    const SmemDirect sm{smem + (size_t)g * 3 * P::LINE};
    row_phase0<P, FAST>(sm, ft, p, rows, ktab, t);
    __syncthreads();
    row_phase1<P>(sm, ft);
    __syncthreads();
    row_phase2<P>(sm, ft, p, sink);
}

template <class P, int G, int MINB>
__global__ void __launch_bounds__(P::T* G, MINB) ow_col_kernel(FrameBuffers fb, SlotTable tab, float scale) {
    extern __shared__ __align__(16) float2 smem[];
    constexpr int N = P::N;
    using LY = ColLayout<P, G>;
    const int job = threadIdx.x % G, ft = threadIdx.x / G;
    const int x = 2 * (blockIdx.x * G + job);
    const int f = blockIdx.y;
    const int slot = tab.slot[blockIdx.z];
    const SmemDirect sm{smem};
    const int base = job * LY::SJ;
    const float2* src = fb.inter + ((size_t)slot * 3 + f) * (N / 2) * N + x;
    float* dst = fb.disp + ((size_t)slot * 3 + f) * N * N + x;
    const FullColGeom<N> geom{};
    // a G=8 tile reads whole 128-byte lines of the intermediate (8 jobs x 16 B): job 0 drops them from L2 after use
    const bool discard = (G == 8) && job == 0 && fb.discard_inter;
#ifdef OW_TRACE
    const int cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
#endif
    OW_STAMP(cta, 0); OW_STAMP(cta, 1);
#pragma unroll 1
    for (int j = ft; j < P::M / 2; j += P::T) col_phase0<P>(sm, base, j, src, geom, discard);
    OW_STAMP(cta, 2);
    __syncthreads();
    OW_STAMP(cta, 3);
    col_phase1<P>(sm, base, ft);
    OW_STAMP(cta, 4);
    __syncthreads();
    OW_STAMP(cta, 5);
    col_phase2<P>(sm, base, ft, dst, scale, geom);
    OW_STAMP(cta, 6);
}

// ---------------------------------------------------------------------------------------------------
// Persistent, register-pipelined column kernel. The per-CTA timeline of ow_col_kernel (tools/tune/trace.cu, N = 2048:
// profiles/r02_col_cta_timeline.txt) shows a tile spending 45 % of its life in phase 0, most of it waiting for its global loads
// with nothing else in flight, and another 10 % between CTAs: a CTA cannot retire (and its successor cannot start) until its
// stores have drained. Here a CTA walks over many (slot entry, channel, 16-column tile) items and always has the FIRST load batch
// of its NEXT tile in flight in registers: it is issued as soon as stage 0 of the current tile has consumed the registers and
// lands during stages 1 and 2; the remaining batches of a tile (N = 2048: one more) are issued at the top of phase 0, before
// the first batch is transformed. Stores drain while the next tile's stage 0 runs.
// ---------------------------------------------------------------------------------------------------
template <class P, int G, int MINB>
__global__ void __launch_bounds__(P::T* G, MINB) ow_col_pipe_kernel(FrameBuffers fb, SlotTable tab, float scale, int total) {
    extern __shared__ __align__(16) float2 smem[];
    constexpr int N = P::N, HP = N / 2, NT = N / (2 * G), TPF = 3 * NT, H = P::R0 / 2, M = P::M;
    constexpr int C0 = (M / 2) / P::T;
    static_assert((M / 2) % P::T == 0 && C0 >= 1, "whole stage-0 batches");
    using LY = ColLayout<P, G>;
    const int job = threadIdx.x % G, ft = threadIdx.x / G;
    const SmemDirect sm{smem};
    const int base = job * LY::SJ;
    const FullColGeom<N> geom{};
    int w = blockIdx.x;
    if (w >= total) return;

    auto src_of = [&](int wi) {
        const int e = wi / TPF, r = wi - e * TPF, f = r / NT, tile = r - f * NT;
        return fb.inter + ((size_t)tab.slot[e] * 3 + f) * HP * N + 2 * (tile * G + job);
    };
    auto issue = [&](const float2* src, int j, float4 (&la)[H], float4 (&lb)[H]) {
        const int bA = j, bB = (j == 0) ? M / 2 : M - j;
#pragma unroll
        for (int i = 0; i < H; ++i) {
            la[i] = OW_LDP(reinterpret_cast<const float4*>(src + (size_t)(i * M + bA) * N));
            lb[i] = OW_LDP(reinterpret_cast<const float4*>(src + (size_t)(i * M + bB) * N));
        }
    };

    float4 na[H], nb[H];                       // batch 0 of the tile about to be transformed
    const float2* src = src_of(w);
    issue(src, ft, na, nb);
    for (; w < total; w += gridDim.x) {
        const int e = w / TPF, r = w - e * TPF, f = r / NT, tile = r - f * NT;
        float* dst = fb.disp + ((size_t)tab.slot[e] * 3 + f) * N * N + 2 * (tile * G + job);
        const int wn = w + gridDim.x;
        if (C0 == 1) {
            col_phase0_math<P>(sm, base, ft, na, nb);
        } else {
            float4 la[H], lb[H];
#pragma unroll
            for (int i = 0; i < H; ++i) { la[i] = na[i]; lb[i] = nb[i]; }
#pragma unroll
            for (int c = 1; c < C0; ++c) {
                issue(src, ft + c * P::T, na, nb);                 // batch c in flight while batch c-1 is transformed
                col_phase0_math<P>(sm, base, ft + (c - 1) * P::T, la, lb);
#pragma unroll
                for (int i = 0; i < H; ++i) { la[i] = na[i]; lb[i] = nb[i]; }
            }
            col_phase0_math<P>(sm, base, ft + (C0 - 1) * P::T, la, lb);
        }
        if (wn < total) {
            src = src_of(wn);
            issue(src, ft, na, nb);                                // lands during stages 1 and 2
        }
        __syncthreads();
        col_phase1<P>(sm, base, ft);
        __syncthreads();
        col_phase2<P>(sm, base, ft, dst, scale, geom);
        __syncthreads();                                           // the next tile's stage-0 stores reuse the lines
    }
}

// ---------------------------------------------------------------------------------------------------
// Column kernel, second generation: persistent CTAs walking (slot entry, channel, 16-column tile) work items, optionally
//   STAGED  the tile's input rows arrive through 2-D TMA copies (ColStage in ow_kernels.cuh) into a shared-memory staging buffer
//           behind an mbarrier; the copy of the NEXT step (or of the next tile's first step) is issued as soon as every thread has
//           taken the current step into registers, so it is in flight during stage-0 math, stages 1-2, the stores and the epilogue;
//   fuse    dy tiles keep their final heights in shared memory and produce the normal map of their three interior column quads as
//           their epilogue (see "COLUMN KERNEL WITH THE NORMAL MAP AS ITS EPILOGUE"); the seam quads between tiles are left to
//           ow_seam_kernel, a light pass over the stored heights.
// Work order inside a frame: dy tiles first (they are the long ones), then dx, then dz.
// ---------------------------------------------------------------------------------------------------
template <class P, int G, bool STAGED>
struct Col2Smem {
    using LY = ColLayout<P, G>;
    using CS = ColStage<P, G>;
    static constexpr size_t STAGE = STAGED ? ((CS::BYTES + 127) / 128) * 128 : 0;
    static constexpr size_t BYTES = STAGE + LY::SMEM;
};

template <class P, int G, int MINB, int RY, bool STAGED>
__global__ void __launch_bounds__(P::T* G, MINB) ow_col2_kernel(const __grid_constant__ CUtensorMap tmap, FrameBuffers fb, SlotTable tab, float scale,
                                                                int total, int fuse) {
    static_assert(G == 8, "tiles are 16 columns: three interior quads + one seam quad");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    constexpr int N = P::N, HP = N / 2, NT = N / (2 * G), TPF = 3 * NT, H = P::R0 / 2, NTHREADS = P::T * G;
    using LY = ColLayout<P, G>;
    using CS = ColStage<P, G>;
    using SM_ = Col2Smem<P, G, STAGED>;
    float4* staging = reinterpret_cast<float4*>(smem_raw);
    const SmemDirect sm{reinterpret_cast<float2*>(smem_raw + SM_::STAGE)};
    const int tid = threadIdx.x, job = tid % G, ft = tid / G;
    const int base = job * LY::SJ;
    const FullColGeom<N> geom{};
    int w = blockIdx.x;
    if (w >= total) return;

    // one elected thread: arm the barrier with the byte count of step s of work item wi and issue its copies
    auto issue = [&](int wi, int s) {
        const int e = wi / TPF, r = wi - e * TPF, f = r / NT, tile = r - f * NT;
        const int row_base = (tab.slot[e] * 3 + f) * HP;
        const uint64_t pol = l2_policy_evict_first();        // the intermediate is read exactly once
        mbar_arrive_expect_tx(&bar, CS::STEP_TX_BYTES + (s == 0 ? CS::EXTRA_TX_BYTES : 0u));
#pragma unroll
        for (int i = 0; i < H; ++i) {
            tma_load_2d(staging + (2 * i + 0) * CS::BOX_F4, &tmap, 4 * tile * G, row_base + CS::fwd_row0(i, s), &bar, pol);
            tma_load_2d(staging + (2 * i + 1) * CS::BOX_F4, &tmap, 4 * tile * G, row_base + CS::mir_row0(i, s), &bar, pol);
        }
        if (s == 0) {
#pragma unroll
            for (int i = 0; i < H; ++i)
                bulk_load(staging + CS::EXTRA_F4 + i * G, fb.inter + ((size_t)(row_base + CS::extra_row(i)) * N + 2 * tile * G), G * 16, &bar, pol);
        }
    };

    unsigned phase = 0;
    if (STAGED) {
        if (tid == 0) {
            mbar_init(&bar, 1);
            mbar_init_fence();
            tma_prefetch_descriptor(&tmap);
        }
        __syncthreads();
        if (tid == 0) issue(w, 0);
    }
    for (; w < total; w += gridDim.x) {
        const int e = w / TPF, r = w - e * TPF, f = r / NT, tile = r - f * NT;
        const int slot = tab.slot[e];
        const int x = 2 * (tile * G + job);
        float* dst = fb.disp + ((size_t)slot * 3 + f) * N * N + x;
        if (STAGED) {
#pragma unroll 1
            for (int s = 0; s < CS::STEPS; ++s) {
                float4 la[H], lb[H];
                mbar_wait(&bar, phase);
                phase ^= 1u;
                col_stage_read<P, G>(staging, job, ft, s, la, lb);
                __syncthreads();          // the step is in registers everywhere (and the previous tile's epilogue is done with the lines)
                if (tid == 0) {
                    if (s + 1 < CS::STEPS) issue(w, s + 1);
                    else if (w + (int)gridDim.x < total) issue(w + gridDim.x, 0);
                }
                col_phase0_math<P>(sm, base, s * P::T + ft, la, lb);
            }
        } else {
            const float2* src = fb.inter + ((size_t)slot * 3 + f) * HP * N + x;
            const bool discard = job == 0 && fb.discard_inter;
#pragma unroll 1
            for (int j = ft; j < P::M / 2; j += P::T) col_phase0<P>(sm, base, j, src, geom, discard);
        }
        __syncthreads();
        col_phase1<P>(sm, base, ft);
        __syncthreads();
        if (f == 0 && fuse) {
            col_phase2_keep<P>(sm, base, ft, dst, scale, geom, true);
            __syncthreads();
            float4* normal = fb.normal + (size_t)slot * N * N;
            col_normals_phase<P, RY>(sm, tid, NTHREADS, LY::SJ, 16 * tile + 2, normal);
        } else {
            col_phase2<P>(sm, base, ft, dst, scale, geom);
        }
        if (!STAGED) __syncthreads();     // the next tile's stage-0 stores reuse the lines
    }
}

// The seam quads of the fused normal map: columns 16k+14 .. 16k+17 of every slot entry, from the stored heights (L2). One thread walks
// RY rows of one seam; consecutive lanes take consecutive row chunks. A quarter of the normal map's texels, none of its FFT work.
template <int N, int RY>
__global__ void __launch_bounds__(128) ow_seam_kernel(FrameBuffers fb, SlotTable tab) {
    const int k = blockIdx.y, slot = tab.slot[blockIdx.z];
    const float* dy = fb.disp + (size_t)slot * 3 * N * N;
    float4* normal = fb.normal + (size_t)slot * N * N;
    const int cA = 16 * k + 12, cB = (16 * (k + 1)) & (N - 1);
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w < N / RY) normal_quad_walk_src<N, RY, false>(SeamRowSrc{dy, (size_t)N, cA, cB}, w * RY, 0.f, EmitSeam{normal, (size_t)N, cA, cB});
}

// Jacobian/foam map alone (frames whose normal map came out of ow_col2_kernel). Lanes run along x: 128 columns per warp.
template <int N, int RY, int WARPS, int MINB>
__global__ void __launch_bounds__(32 * WARPS, MINB) ow_jac_kernel(FrameBuffers fb, SlotTable tab) {
    const int lane = threadIdx.x, w = threadIdx.y;
    const int x0 = blockIdx.x * 128 + 4 * lane, y0 = (blockIdx.y * WARPS + w) * RY;
    const int e = blockIdx.z, slot = tab.slot[e];
    const CascadeDev c = fb.casc[tab.cascade[e]];
    const float s = c.choppiness * ((float)N / (2.0f * c.L));
    float* jac = fb.jacobian + (size_t)slot * N * N;
    jac_quad_walk<N, RY>(fb.disp + (size_t)slot * 3 * N * N, FullNrmGeom<N>{}, x0, y0, s,
                         [=](int y, float4 J) { __stcs(reinterpret_cast<float4*>(jac + (size_t)y * N + x0), J); });
}

// Slab variant: the column slab's receive buffer [p][c][XH] (row stride 3*XH) -> disp_loc[c][y][XH].
template <class P, int G, int MINB>
__global__ void __launch_bounds__(P::T* G, MINB) ow_col_slab_kernel(const float2* __restrict__ recv, float* __restrict__ disp, int XH,
                                                                    float scale) {
    extern __shared__ __align__(16) float2 smem[];
    constexpr int N = P::N;
    using LY = ColLayout<P, G>;
    const int job = threadIdx.x % G, ft = threadIdx.x / G;
    const int x = 2 * (blockIdx.x * G + job);
    const int f = blockIdx.y;
    const SmemDirect sm{smem};
    const int base = job * LY::SJ;
    const float2* src = recv + (size_t)f * XH + x;
    float* dst = disp + (size_t)f * N * XH + x;
    const SlabColGeom geom{(size_t)3 * XH, (size_t)XH};
#pragma unroll 1
    for (int j = ft; j < P::M / 2; j += P::T) col_phase0<P>(sm, base, j, src, geom);
    __syncthreads();
    col_phase1<P>(sm, base, ft);
    __syncthreads();
    col_phase2<P>(sm, base, ft, dst, scale, geom);
}

// ---------------------------------------------------------------------------------------------------
// N = A * B > 4096 (see "LINES LONGER THAN ONE CTA'S SHARED MEMORY" in ow_kernels.cuh): a lines kernel (one CTA per sub-line
// of decimated input) and a post kernel (twiddles + radix-A + coalesced final stores) per direction. P is the sub-line
// plan (P::N = B). Rows is FullRows<N> or SlabRows<N>, Sink FullSink<N> or SlabSink<N>.
//   folded rows   [pair][A][B] float4  sub-line-major (subline_index), and the k table in the same order (ktab_sub)
//   row scratch   [pair][3][A][B]      (pair index local to the launch: p - p_first)
//   col scratch   [3][A][B][npairs]    (npairs = column pairs of the launch: N/2, or XH/2 for a slab)
// ---------------------------------------------------------------------------------------------------
template <class P, int A, int MINB, bool FAST, class Rows>
__global__ void __launch_bounds__(P::T, MINB) ow_bigrow_lines_kernel(Rows rows, const float* __restrict__ ktab, const float* __restrict__ ktab_sub,
                                                                     int p_first, float t, float2* __restrict__ scratch) {
    extern __shared__ __align__(16) float2 smem[];
    constexpr int N = A * P::N;
    const int ft = threadIdx.x;
    const int pl = blockIdx.x / A, a = blockIdx.x % A;
    const SmemDirect sm{smem};
    bigrow_phase0<P, A, FAST>(sm, ft, p_first + pl, a, rows, ktab, ktab_sub, t);
    __syncthreads();
    row_phase1<P>(sm, ft);
    __syncthreads();
    bigrow_phase2<P, A>(sm, ft, a, scratch + (size_t)pl * 3 * N);
}

template <int B, int A, class Sink>
__global__ void __launch_bounds__(256) ow_bigrow_post_kernel(const float2* __restrict__ scratch, int p_first, Sink sink) {
    constexpr int N = A * B;
    const int kb = blockIdx.x * 256 + threadIdx.x, pl = blockIdx.y, c = blockIdx.z;
    if (kb < B) bigrow_post<B, A>(scratch + (size_t)pl * 3 * N, c, p_first + pl, kb, sink);
}

// The same work from a SMALL persistent grid (a few CTAs per SM, grid-stride over the (channel, pair, kb block) items). With peer stores the
// post kernel is bound by NVLink, not by the SMs: a full grid parks 8 stalled CTAs on every SM and keeps the column kernels of the
// previous frame (which run concurrently on another stream in the frame-pipelined slab path) from being scheduled; a slim one leaves them room.
template <int B, int A, class Sink>
__global__ void __launch_bounds__(256) ow_bigrow_post_slim_kernel(const float2* __restrict__ scratch, int p_first, int npairs, Sink sink) {
    constexpr int N = A * B, KBB = (B + 255) / 256;
    const int total = KBB * npairs * 3;
#pragma unroll 1
    for (int item = blockIdx.x; item < total; item += gridDim.x) {
        const int kbb = item % KBB, r = item / KBB, pl = r % npairs, c = r / npairs;
        const int kb = kbb * 256 + threadIdx.x;
        if (kb < B) bigrow_post<B, A>(scratch + (size_t)pl * 3 * N, c, p_first + pl, kb, sink);
    }
}

// Order of the column-lines items of one channel: CTA index -> (tile, sub-line), sub-line fastest, so the CTAs in flight cover all A
// sub-lines of ~9 adjacent tiles and the rows that sub-lines a and A - a share are read close together (served once from DRAM).
// (Measured alternative, N = 32768: sub-line PAIRS {a, A-a} x 64 adjacent tiles in flight, for longer contiguous runs of every row:
// column pass 20.1 ms instead of 18.4 ms - profiles/r02_experiments.md.)
__host__ __device__ __forceinline__ void bigcol_item(int idx, int A, int* tile, int* a) {
    *tile = idx / A;
    *a = idx - (idx / A) * A;
}

// src: channel-0 base of the Hermitian-packed intermediate; channel c at src + c*src_chan.
template <class P, int A, int G, int MINB, class Geom>
__global__ void __launch_bounds__(P::T* G, MINB) ow_bigcol_lines_kernel(const float2* __restrict__ src, size_t src_chan, int npairs,
                                                                        float2* __restrict__ scratch, Geom geom) {
    extern __shared__ __align__(16) float2 smem[];
    constexpr int N = A * P::N;
    using LY = ColLayout<P, G>;
    const int job = threadIdx.x % G, ft = threadIdx.x / G;
    int tile, a;
    bigcol_item(blockIdx.x, A, &tile, &a);
    const int pair = tile * G + job;
    const int f = blockIdx.y;
    const SmemDirect sm{smem};
    const int base = job * LY::SJ;
    bigcol_phase0<P, A>(sm, base, ft, a, src + (size_t)f * src_chan + 2 * pair, geom);
    __syncthreads();
    col_phase1<P>(sm, base, ft);
    __syncthreads();
    bigcol_phase2<P>(sm, base, ft, scratch + (size_t)f * N * npairs + (size_t)a * P::N * npairs + pair, (size_t)npairs);
}

// The same kernel, persistent and register-pipelined (the big-grid counterpart of ow_col_pipe_kernel). ow_bigcol_lines_kernel holds ONE
// 512-thread CTA per SM (a 16-column tile of 2048-point sub-lines is 141 KB of shared memory), and that CTA runs load batch -> wait a DRAM
// round trip -> transform, four times per tile, then stages 1-2 and the stores with nothing in flight: 3.3 TB/s at N = 32768. Here a CTA walks
// over the (channel, tile, sub-line) items - sub-line fastest, so the rows sub-lines a and A-a share are read close together - and always has
// the NEXT batch's R0 row loads in flight in registers: the next batch of this tile during a batch's arithmetic, the first batch of the next
// tile during stages 1 and 2 and their stores.
template <class P, int A, int G, int MINB, class Geom>
__global__ void __launch_bounds__(P::T* G, MINB) ow_bigcol_lines_pipe_kernel(const float2* __restrict__ src, size_t src_chan, int npairs,
                                                                             float2* __restrict__ scratch, Geom geom, int total) {
    extern __shared__ __align__(16) float2 smem[];
    constexpr int N = A * P::N, R0 = P::R0, C0 = P::M / P::T;
    static_assert(P::M % P::T == 0 && C0 >= 1, "whole stage-0 batches");
    using LY = ColLayout<P, G>;
    const int job = threadIdx.x % G, ft = threadIdx.x / G;
    const SmemDirect sm{smem};
    const int base = job * LY::SJ;
    const size_t ss = geom.src_stride();
    const int per_f = npairs / G * A;                 // items per channel
    int w = blockIdx.x;
    if (w >= total) return;
    auto src_of = [&](int wi, int* a) {
        const int f = wi / per_f;
        int tile;
        bigcol_item(wi - f * per_f, A, &tile, a);
        return src + (size_t)f * src_chan + 2 * (tile * G + job);
    };
    float4 nxt[R0];
    int a = 0;
    const float2* s = src_of(w, &a);
    bigcol_issue<P, A>(ft, a, s, ss, nxt);
    for (; w < total; w += gridDim.x) {
        const int f = w / per_f;
        int tile_w, a_w;
        bigcol_item(w - f * per_f, A, &tile_w, &a_w);
        const int pair = tile_w * G + job;
        const int wn = w + gridDim.x;
        int an = 0;
        const float2* sn = wn < total ? src_of(wn, &an) : s;
#pragma unroll
        for (int c = 0; c < C0; ++c) {
            float4 cur[R0];
#pragma unroll
            for (int d0 = 0; d0 < R0; ++d0) cur[d0] = nxt[d0];
            if (c + 1 < C0) bigcol_issue<P, A>(ft + (c + 1) * P::T, a, s, ss, nxt);          // next batch of this tile
            else if (wn < total) bigcol_issue<P, A>(ft, an, sn, ss, nxt);                      // first batch of the next tile: lands during stages 1-2
            bigcol_phase0_math<P, A>(sm, base, ft + c * P::T, a, cur);
        }
        __syncthreads();
        col_phase1<P>(sm, base, ft);
        __syncthreads();
        bigcol_phase2<P>(sm, base, ft, scratch + (size_t)f * N * npairs + (size_t)a * P::N * npairs + pair, (size_t)npairs);
        __syncthreads();                               // the next tile's stage-0 stores reuse the lines
        s = sn; a = an;
    }
}

// dst: channel-0 base of the displacement planes; channel c at dst + c*dst_chan, rows ds floats apart.
template <int B, int A>
__global__ void __launch_bounds__(256) ow_bigcol_post_kernel(const float2* __restrict__ scratch, int npairs, float* __restrict__ dst, size_t dst_chan,
                                                             size_t ds, float scale) {
    constexpr int N = A * B;
    const int pair = blockIdx.x * 32 + threadIdx.x, kb = blockIdx.y * 8 + threadIdx.y, c = blockIdx.z;
    if (pair < npairs && kb < B)
        bigcol_post<B, A>(scratch + (size_t)c * N * npairs + pair, (size_t)npairs, kb, dst + (size_t)c * dst_chan + 2 * pair, ds, scale);
}

// ---------------------------------------------------------------------------------------------------
// Cluster versions (see "THE SAME DECOMPOSITION INSIDE ONE THREAD-BLOCK CLUSTER" in ow_kernels.cuh): launched with cluster dimension
// (A, 1, 1), so the A CTAs blockIdx.x = A*line + a hold the A sub-lines of one line (row pair / column tile) and cluster rank == a.
// ---------------------------------------------------------------------------------------------------
struct ClusterPeers {               // element i of CTA a's dynamic shared memory
    float2* mine;
    __device__ __forceinline__ float2 ld(int a, int i) const {
        return *cooperative_groups::this_cluster().map_shared_rank(mine + i, a);
    }
};

template <class P, int A, int MINB, bool FAST, class Rows, class Sink>
__global__ void __launch_bounds__(P::T, MINB) ow_bigrow_cluster_kernel(Rows rows, const float* __restrict__ ktab, const float* __restrict__ ktab_sub,
                                                                       int p_first, float t, Sink sink) {
    extern __shared__ __align__(16) float2 smem[];
    constexpr int B = P::N, SLICE = B / A;
    static_assert(B % A == 0, "kb slices");
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int ft = threadIdx.x;
    const int pl = blockIdx.x / A, a = blockIdx.x % A;
    const SmemDirect sm{smem};
    bigrow_phase0<P, A, FAST>(sm, ft, p_first + pl, a, rows, ktab, ktab_sub, t);
    __syncthreads();
    row_phase1<P>(sm, ft);
    __syncthreads();
    bigrow_phase2_inplace<P>(sm, ft);
    cluster.sync();                                  // every sub-line's z is in its CTA's shared memory
    const ClusterPeers peers{smem};
#pragma unroll 1
    for (int task = ft; task < 3 * SLICE; task += P::T) {
        const int c = task / SLICE, kb = a * SLICE + (task - c * SLICE);
        bigrow_post_dsm<P, A>(peers, c, p_first + pl, kb, sink);
    }
    cluster.sync();                                  // nobody retires while a peer may still read its lines
}

template <class P, int A, int G, int MINB, class Geom>
__global__ void __launch_bounds__(P::T* G, MINB) ow_bigcol_cluster_kernel(const float2* __restrict__ src, size_t src_chan, float* __restrict__ dst,
                                                                         size_t dst_chan, float scale, Geom geom) {
    extern __shared__ __align__(16) float2 smem[];
    constexpr int B = P::N, SLICE = B / A;
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    using LY = ColLayout<P, G>;
    const int job = threadIdx.x % G, ft = threadIdx.x / G;
    const int tile = blockIdx.x / A, a = blockIdx.x % A;
    const int pair = tile * G + job;
    const int f = blockIdx.y;
    const SmemDirect sm{smem};
    const int base = job * LY::SJ;
    bigcol_phase0<P, A>(sm, base, ft, a, src + (size_t)f * src_chan + 2 * pair, geom);
    __syncthreads();
    col_phase1<P>(sm, base, ft);
    __syncthreads();
    bigcol_phase2_inplace<P>(sm, base, ft);
    cluster.sync();
    const ClusterPeers peers{smem};
    float* out = dst + (size_t)f * dst_chan + 2 * pair;
#pragma unroll 1
    for (int kl = ft; kl < SLICE; kl += P::T) bigcol_post_dsm<P, A>(peers, base, a * SLICE + kl, out, geom.dst_stride(), scale);
    cluster.sync();
}

constexpr int kNormalRows = 8;      // output rows per thread of the normal kernel's walk
constexpr int kNormalWarps = 4;     // warps per CTA; each warp owns a 128-column x kNormalRows-row tile

// Device emit of normal_quad_walk: the four normals a thread produced for columns x0..x0+3 go through a
// per-warp 2 KB shared-memory tile (XOR-swizzled, conflict-free both ways) so that every STG.128 of the warp
// covers 512 contiguous bytes instead of 32 separate 16-byte pieces.
template <bool JAC>
struct EmitStaged {
    float4* normal;   // slot base
    float* jac;
    float4* tile;     // this warp's 128 float4
    size_t ostride;   // elements between output rows
    int xw, lane;     // first OUTPUT column of the warp tile
    static __device__ __forceinline__ int swz(int t) { return t ^ ((t >> 3) & 7); }
    __device__ __forceinline__ void operator()(int y, const float4 (&n)[4], float4 J) const {
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) tile[swz(4 * lane + j)] = n[j];
        __syncwarp();
        float4* d = normal + (size_t)y * ostride + xw + lane;
#pragma unroll
        for (int k = 0; k < 4; ++k) __stcs(&d[32 * k], tile[swz(32 * k + lane)]);      // streaming: never re-read on the device
        if (JAC) __stcs(reinterpret_cast<float4*>(jac + (size_t)y * ostride + xw + 4 * lane), J);
    }
};

// RY = output rows per thread walk, WARPS = warps per CTA (each owns a 128-column x RY-row tile).
template <int N, bool JAC, int RY, int WARPS, int MINB>
__global__ void __launch_bounds__(32 * WARPS, MINB) ow_normal_kernel(FrameBuffers fb, SlotTable tab) {
    __shared__ float4 tiles[WARPS][128];
    const int lane = threadIdx.x, w = threadIdx.y;
    const int xw = blockIdx.x * 128, y0 = (blockIdx.y * WARPS + w) * RY;
    const int e = blockIdx.z;
    const int slot = tab.slot[e];
    const float* disp = fb.disp + (size_t)slot * 3 * N * N;
    float s = 0.f;
    if (JAC) {
        const CascadeDev c = fb.casc[tab.cascade[e]];
        s = c.choppiness * ((float)N / (2.0f * c.L));
    }
    const EmitStaged<JAC> emit{fb.normal + (size_t)slot * N * N, JAC ? fb.jacobian + (size_t)slot * N * N : nullptr, tiles[w], (size_t)N, xw, lane};
    normal_quad_walk<N, RY, JAC>(disp, FullNrmGeom<N>{}, xw + 4 * lane, y0, s, emit);
}

// Slab variant: stencil over the padded column slab disp_loc[3][N][XH] (halo columns present, no x wrap);
// outputs normal_loc[N][XL], jac_loc[N][XL] for the XL interior columns.
template <int N, bool JAC, int RY, int WARPS, int MINB>
__global__ void __launch_bounds__(32 * WARPS, MINB) ow_normal_slab_kernel(const float* __restrict__ disp, float4* __restrict__ normal,
                                                                          float* __restrict__ jac, int XL, int XH, float s) {
    __shared__ float4 tiles[WARPS][128];
    const int lane = threadIdx.x, w = threadIdx.y;
    const int xw = blockIdx.x * 128, y0 = (blockIdx.y * WARPS + w) * RY;
    const EmitStaged<JAC> emit{normal, jac, tiles[w], (size_t)XL, xw, lane};
    normal_quad_walk<N, RY, JAC>(disp, SlabNrmGeom{(size_t)XH, (size_t)N * XH}, kHalo + xw + 4 * lane, y0, s, emit);
}

}  // namespace ow
