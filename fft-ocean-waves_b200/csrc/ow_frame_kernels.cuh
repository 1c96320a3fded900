// ow_frame_kernels.cuh — __global__ wrappers of the per-frame kernels (sm_100a). Bodies live in ow_kernels.cuh
// (shared with the CPU emulator used by the tests); the per-N choices of the template parameters are in
// ow_config.cuh. Included by ow_frame_kernels.cu (the library) and by tools/tune (the variant sweeper).
#pragma once
#include "ow_internal.h"
#include "ow_kernels.cuh"
#include "ow_config.cuh"

namespace ow {

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
    const float4* h0 = fb.h0 + (size_t)cascade * N * N;
    const float* ktab = fb.ktab + (size_t)cascade * N;
    float2* inter = fb.inter + (size_t)slot * 3 * (N / 2) * N;
    row_phase0<P, FAST>(sm, ft, p, FullRows<N>{h0}, ktab, t);
    __syncthreads();
    row_phase1<P>(sm, ft);
    __syncthreads();
    row_phase2<P>(sm, ft, p, FullSink<N>{inter});
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
#pragma unroll 1
    for (int j = ft; j < P::M / 2; j += P::T) col_phase0<P>(sm, base, j, src, geom);
    __syncthreads();
    col_phase1<P>(sm, base, ft);
    __syncthreads();
    col_phase2<P>(sm, base, ft, dst, scale, geom);
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
        for (int k = 0; k < 4; ++k) d[32 * k] = tile[swz(32 * k + lane)];
        if (JAC) *reinterpret_cast<float4*>(jac + (size_t)y * ostride + xw + 4 * lane) = J;
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
