// ow_mega_kernels.cu — the whole frame as ONE persistent kernel (sm_100a): a dataflow walk over row, column and normal-map work
// items of consecutive frames behind per-frame dependency counters.
//
// Why: with separate kernels every configuration is DRAM-bound on the separate-kernel traffic (64-76 B/texel at 85-95 % of
// the copy peak) — the row->column intermediate and the displacement planes make a round trip through DRAM because a launch group must be
// large (tens of frames) to amortise the ramp-up and tail of three dependent launches, and by then nothing is left in L2. Here a grid of
// resident CTAs claims items from one ordered queue
//     block b = [ rows(frame b) | columns(frame b - LAG_COL) | normals(frame b - LAG_COL - LAG_NRM) ]
// so that at any time only a few frames are in flight: the intermediate (12 B/texel) and the displacement planes (12 B/texel) are
// consumed out of L2 a few microseconds after they were written, every SM runs compute-heavy (row, column) and memory-heavy (normal) items side
// by side, and there are no launch boundaries. A column item waits until all row items of its frame have signalled (release/acquire on a
// counter in global memory), a normal item until all column items have; an item only ever waits for items EARLIER in the queue and every
// claimed item is held by a resident CTA, so the walk cannot deadlock. Data produced inside the launch is read through L2 (ld.global.cg:
// OW_COHERENT_LOADS), never through the non-coherent L1 path.
//
// Replaces, like the three kernels it is made of, the reference's tilde_h0_t + 2*log2(N) butterfly dispatches per channel + inversion +
// normal map (src/main.cpp:240-244, 587-707). The per-thread phase functions are the ones of ow_kernels.cuh (CPU-emulated in tests/emu).
#define OW_COHERENT_LOADS 1
#include "ow_frame_kernels.cuh"

namespace ow {

namespace {

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// One thread of the CTA: block until *p >= want. A dependency that never arrives (a bug) must not wedge the GPU: after ~1 s the kernel traps.
__device__ __forceinline__ void wait_count(const int* p, int want) {
    unsigned spins = 0;
    while (ld_acquire_gpu(p) < want) {
        __nanosleep(40);
        if (++spins > (1u << 24)) __trap();
    }
}
// After a __syncthreads that follows the item's last store: publish the CTA's writes at gpu scope, then count the item.
__device__ __forceinline__ void signal_done(int* p) {
    __threadfence();
    atomicAdd(p, 1);
}

// Queue lags (in frames): block b = [ rows(b) | columns(b - LAG_COL) | normals(b - LAG_COL - LAG_NRM) ]. A lag of 1 makes a column item wait for
// row items claimed only a fraction of a grid round earlier; larger lags trade frames in flight (L2 footprint) for fewer stalls.
template <int N>
struct MegaCfg;
template <>
struct MegaCfg<256> {
    using Row = Plan<256, 4, 4, 16, 64, 1, 0>;
    static constexpr int PAIRS = 4, MINB = 2, LAG_COL = 3, LAG_NRM = 2;
};
template <>
struct MegaCfg<512> {
    using Row = Plan<512, 8, 4, 16, 64, 1, 14>;
    static constexpr int PAIRS = 4, MINB = 2, LAG_COL = 3, LAG_NRM = 2;      // measured: lags (1,1) 228 k, (2,2) 281 k, (3,2) 307 k, (4,3) 308 k frames/s at C2
};
template <>
struct MegaCfg<1024> {
    using Row = Plan<1024, 8, 8, 16, 128, 1, 10>;
    static constexpr int PAIRS = 4, MINB = 1, LAG_COL = 1, LAG_NRM = 1;      // one 512-thread CTA per SM: longer lags only add frames in flight (59.8 k vs 55.9 k at (2,2))
};

template <int N, bool JAC>
struct MegaShape {
    using C = Cfg<N>;
    using M = MegaCfg<N>;
    using PR = typename M::Row;
    using PK = typename C::Col;
    static constexpr int G = C::COL_G, PAIRS = M::PAIRS, NT = PK::T * G, RY = C::NRM_RY;
    static_assert(PR::T * PAIRS == NT, "row and column items use the same CTA");
    static constexpr int RI = N / 2 / PAIRS;                 // row items per frame
    static constexpr int NTILE = N / (2 * G), CI = 3 * NTILE;   // column items per frame
    static constexpr int WARPS = NT / 32, NXT = N / 128, NYC = N / RY;
    static_assert((NXT * NYC) % WARPS == 0, "whole normal items");
    static constexpr int NI = NXT * NYC / WARPS;             // normal items per frame
    static constexpr int S = RI + CI + NI;                   // items per queue block
    static constexpr int LAG_COL = M::LAG_COL, LAG_NRM = M::LAG_NRM;
    static constexpr size_t SMEM_ROW = row_smem<PR, PAIRS>(), SMEM_COL = ColLayout<PK, G>::SMEM, SMEM_NRM = (size_t)WARPS * 128 * sizeof(float4);
    static constexpr size_t SMEM = SMEM_ROW > SMEM_COL ? (SMEM_ROW > SMEM_NRM ? SMEM_ROW : SMEM_NRM) : (SMEM_COL > SMEM_NRM ? SMEM_COL : SMEM_NRM);
};

template <int N, bool JAC, bool FAST>
__global__ void __launch_bounds__(MegaShape<N, JAC>::NT, MegaCfg<N>::MINB) ow_mega_kernel(FrameBuffers fb, SlotTable tab, int count, float scale, int* sched) {
    using SH = MegaShape<N, JAC>;
    using PR = typename SH::PR;
    using PK = typename SH::PK;
    using LY = ColLayout<PK, SH::G>;
    constexpr int G = SH::G, RI = SH::RI, CI = SH::CI, S = SH::S, NTILE = SH::NTILE;
    extern __shared__ __align__(16) float2 smem[];
    __shared__ int s_item;
    int* head = sched;
    int* rows_done = sched + 1;
    int* cols_done = sched + 1 + kMaxGroup;
    const int tid = threadIdx.x;
    const int total = (count + SH::LAG_COL + SH::LAG_NRM) * S;
    for (;;) {
        if (tid == 0) s_item = atomicAdd(head, 1);
        __syncthreads();
        const int item = s_item;
        __syncthreads();                         // everyone has read s_item before the next claim overwrites it; also fences the shared-memory reuse
        if (item >= total) break;
        const int blk = item / S, r = item - blk * S;
        if (r < RI) {
            // ---- rows of frame entry e: spectrum at time t + row IFFT of PAIRS row pairs ------------------------------------
            const int e = blk;
            if (e >= count) continue;
            const int ft = tid % PR::T, g = tid / PR::T;
            const int p = r * SH::PAIRS + g;
            const int cascade = tab.cascade[e], slot = tab.slot[e];
            const float t = tab.time[e];
            const SmemDirect sm{smem + (size_t)g * 3 * PR::LINE};
            const FullRows<N> rows{fb.h0 + (size_t)cascade * N * N, fb.hp + (size_t)cascade * hp_block_f4(N / 2, N), fb.nyq + (size_t)cascade * (N / 2)};
            const float* ktab = fb.ktab + (size_t)cascade * N;
            float2* inter = fb.inter + (size_t)slot * 3 * (N / 2) * N;
            row_phase0<PR, FAST>(sm, ft, p, rows, ktab, t);
            __syncthreads();
            row_phase1<PR>(sm, ft);
            __syncthreads();
            row_phase2<PR>(sm, ft, p, FullSink<N>{inter});
            __syncthreads();
            if (tid == 0) signal_done(&rows_done[e]);
        } else if (r < RI + CI) {
            // ---- columns of frame entry e: one channel's 16-column tile, inversion as the epilogue --------------------------
            const int e = blk - SH::LAG_COL;
            if (e < 0 || e >= count) continue;
            const int idx = r - RI, f = idx / NTILE, tile = idx - f * NTILE;
            const int job = tid % G, ft = tid / G;
            const int slot = tab.slot[e];
            const int x = 2 * (tile * G + job);
            const SmemDirect sm{smem};
            const int base = job * LY::SJ;
            const float2* src = fb.inter + ((size_t)slot * 3 + f) * (N / 2) * N + x;
            float* dst = fb.disp + ((size_t)slot * 3 + f) * N * N + x;
            const FullColGeom<N> geom{};
            if (tid == 0) wait_count(&rows_done[e], RI);
            __syncthreads();
#pragma unroll 1
            for (int j = ft; j < PK::M / 2; j += PK::T) col_phase0<PK>(sm, base, j, src, geom, false);
            __syncthreads();
            col_phase1<PK>(sm, base, ft);
            __syncthreads();
            col_phase2<PK>(sm, base, ft, dst, scale, geom);
            __syncthreads();
            if (tid == 0) signal_done(&cols_done[e]);
        } else {
            // ---- normal map (+ Jacobian) of frame entry e: one 128-column x RY-row tile per warp ----------------------------
            const int e = blk - SH::LAG_COL - SH::LAG_NRM;
            if (e < 0 || e >= count) continue;
            const int idx = r - RI - CI;
            const int lane = tid & 31, w = tid >> 5;
            const int task = idx * SH::WARPS + w, xt = task % SH::NXT, yc = task / SH::NXT;
            const int xw = xt * 128, y0 = yc * SH::RY;
            const int slot = tab.slot[e];
            const float* disp = fb.disp + (size_t)slot * 3 * N * N;
            float s = 0.f;
            if (JAC) {
                const CascadeDev c = fb.casc[tab.cascade[e]];
                s = c.choppiness * ((float)N / (2.0f * c.L));
            }
            float4* tile = reinterpret_cast<float4*>(smem) + (size_t)w * 128;
            const EmitStaged<JAC> emit{fb.normal + (size_t)slot * N * N, JAC ? fb.jacobian + (size_t)slot * N * N : nullptr, tile, (size_t)N, xw, lane};
            if (tid == 0) wait_count(&cols_done[e], CI);
            __syncthreads();
            normal_quad_walk<N, SH::RY, JAC>(disp, FullNrmGeom<N>{}, xw + 4 * lane, y0, s, emit);
        }
    }
}

template <int N, bool JAC, bool FAST>
cudaError_t configure_one(int* ctas_per_sm) {
    using SH = MegaShape<N, JAC>;
    auto k = ow_mega_kernel<N, JAC, FAST>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SH::SMEM);
    if (e != cudaSuccess) return e;
    int n = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, SH::NT, SH::SMEM);
    if (e != cudaSuccess) return e;
    if (n < *ctas_per_sm || *ctas_per_sm == 0) *ctas_per_sm = n;
    return cudaSuccess;
}

template <int N>
cudaError_t configure_mega_n(KernelConfig* cfg) {
    int n = 0;
    cudaError_t e;
    if ((e = configure_one<N, false, false>(&n)) != cudaSuccess) return e;
    if ((e = configure_one<N, false, true>(&n)) != cudaSuccess) return e;
    if ((e = configure_one<N, true, false>(&n)) != cudaSuccess) return e;
    if ((e = configure_one<N, true, true>(&n)) != cudaSuccess) return e;
    cfg->mega_ctas = n;
    return cudaSuccess;
}

template <int N>
int launch_mega_n(const FrameBuffers& fb, const SlotTable& tab, int count, bool with_jac, bool fast, cudaStream_t st) {
    const float scale = 0.5f / ((float)N * (float)N);
    cudaError_t e = cudaMemsetAsync(fb.mega_sched, 0, kMegaSchedInts * sizeof(int), st);
    if (e != cudaSuccess) { stash_launch_error(e); return -1; }
    auto go = [&](auto jac, auto fa) {
        using SH = MegaShape<N, decltype(jac)::value>;
        const int total = (count + SH::LAG_COL + SH::LAG_NRM) * SH::S, resident = fb.sm_count * fb.mega_ctas;
        ow_mega_kernel<N, decltype(jac)::value, decltype(fa)::value><<<total < resident ? total : resident, SH::NT, SH::SMEM, st>>>(fb, tab, count, scale, fb.mega_sched);
    };
    if (with_jac) { if (fast) go(std::true_type{}, std::true_type{}); else go(std::true_type{}, std::false_type{}); }
    else { if (fast) go(std::false_type{}, std::true_type{}); else go(std::false_type{}, std::false_type{}); }
    return launches_ok() ? 1 : -1;
}

}  // namespace

bool mega_supported(int N) { return N == 256 || N == 512 || N == 1024; }

cudaError_t configure_mega(int N, KernelConfig* cfg) {
    switch (N) {
        case 256: return configure_mega_n<256>(cfg);
        case 512: return configure_mega_n<512>(cfg);
        case 1024: return configure_mega_n<1024>(cfg);
    }
    return cudaSuccess;
}

int launch_mega_frame(const FrameBuffers& fb, const SlotTable& tab, int count, bool with_jac, bool fast, cudaStream_t st) {
    if (!fb.mega_sched || fb.mega_ctas < 1) return -1;
    switch (fb.N) {
        case 256: return launch_mega_n<256>(fb, tab, count, with_jac, fast, st);
        case 512: return launch_mega_n<512>(fb, tab, count, with_jac, fast, st);
        case 1024: return launch_mega_n<1024>(fb, tab, count, with_jac, fast, st);
    }
    return -1;
}

}  // namespace ow
