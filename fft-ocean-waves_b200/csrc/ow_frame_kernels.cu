// ow_frame_kernels.cu — __global__ wrappers and launchers of the per-frame kernels (sm_100a).
// Kernel bodies live in ow_kernels.cuh (shared with the CPU emulator used by the tests).
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
    row_phase0<P, FAST>(sm, ft, p, h0, ktab, t);
    __syncthreads();
    row_phase1<P>(sm, ft);
    __syncthreads();
    row_phase2<P>(sm, ft, p, inter);
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
#pragma unroll 1
    for (int j = ft; j < P::M / 2; j += P::T) col_phase0<P>(sm, base, j, src);
    __syncthreads();
    col_phase1<P>(sm, base, ft);
    __syncthreads();
    col_phase2<P>(sm, base, ft, dst, scale);
}

constexpr int kNormalRows = 8;   // output rows per thread of the normal kernel's column walk

template <int N, bool JAC>
__global__ void __launch_bounds__(256) ow_normal_kernel(FrameBuffers fb, SlotTable tab) {
    const int x = blockIdx.x * 32 + threadIdx.x, y0 = (blockIdx.y * 8 + threadIdx.y) * kNormalRows;
    const int e = blockIdx.z;
    const int slot = tab.slot[e];
    const float* disp = fb.disp + (size_t)slot * 3 * N * N;
    float lambda = 0.f, inv2h = 0.f;
    if (JAC) {
        const CascadeDev c = fb.casc[tab.cascade[e]];
        lambda = c.choppiness;
        inv2h = (float)N / (2.0f * c.L);
    }
    normal_column_walk<N, kNormalRows, JAC>(disp, fb.normal + (size_t)slot * N * N,
                                            JAC ? fb.jacobian + (size_t)slot * N * N : nullptr, x, y0, lambda, inv2h);
}

// ---------------------------------------------------------------------------------------------------
template <int N>
cudaError_t configure_n() {
    using C = Cfg<N>;
    cudaError_t e = cudaFuncSetAttribute(ow_row_kernel<typename C::Row, C::ROW_PAIRS, C::ROW_MINB, false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)row_smem<typename C::Row, C::ROW_PAIRS>());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(ow_row_kernel<typename C::Row, C::ROW_PAIRS, C::ROW_MINB, true>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_smem<typename C::Row, C::ROW_PAIRS>());
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(ow_col_kernel<typename C::Col, C::COL_G, C::COL_MINB>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)ColLayout<typename C::Col, C::COL_G>::SMEM);
}

template <int N>
int launch_n(const FrameBuffers& fb, const SlotTable& tab, int count, bool with_jac, bool fast_phase, cudaStream_t st, cudaEvent_t* ev) {
    using C = Cfg<N>;
    using R = typename C::Row;
    using K = typename C::Col;
    if (ev) cudaEventRecord(ev[0], st);
    const dim3 rgrid(N / 2 / C::ROW_PAIRS, count);
    if (fast_phase)
        ow_row_kernel<R, C::ROW_PAIRS, C::ROW_MINB, true><<<rgrid, R::T * C::ROW_PAIRS, row_smem<R, C::ROW_PAIRS>(), st>>>(fb, tab);
    else
        ow_row_kernel<R, C::ROW_PAIRS, C::ROW_MINB, false><<<rgrid, R::T * C::ROW_PAIRS, row_smem<R, C::ROW_PAIRS>(), st>>>(fb, tab);
    if (ev) cudaEventRecord(ev[1], st);
    const float scale = 0.5f / ((float)N * (float)N);   // 1/2 from the Hermitian split, 1/N^2 from inversion_cs.glsl:36
    ow_col_kernel<K, C::COL_G, C::COL_MINB>
        <<<dim3(N / (2 * C::COL_G), 3, count), K::T * C::COL_G, ColLayout<K, C::COL_G>::SMEM, st>>>(fb, tab, scale);
    if (ev) cudaEventRecord(ev[2], st);
    const dim3 ngrid(N / 32, N / (8 * kNormalRows), count);
    if (with_jac) ow_normal_kernel<N, true><<<ngrid, dim3(32, 8), 0, st>>>(fb, tab);
    else ow_normal_kernel<N, false><<<ngrid, dim3(32, 8), 0, st>>>(fb, tab);
    if (ev) cudaEventRecord(ev[3], st);
    return cudaGetLastError() == cudaSuccess ? 3 : -1;
}

bool frame_supported(int N) { return N == 256 || N == 512 || N == 1024 || N == 2048 || N == 4096; }

cudaError_t configure_frame_kernels(int N) {
    switch (N) {
        case 256: return configure_n<256>();
        case 512: return configure_n<512>();
        case 1024: return configure_n<1024>();
        case 2048: return configure_n<2048>();
        case 4096: return configure_n<4096>();
    }
    return cudaErrorInvalidValue;
}

int launch_frame(const FrameBuffers& fb, const SlotTable& tab, int count, bool with_jac, bool fast_phase, cudaStream_t st, cudaEvent_t* ev) {
    switch (fb.N) {
        case 256: return launch_n<256>(fb, tab, count, with_jac, fast_phase, st, ev);
        case 512: return launch_n<512>(fb, tab, count, with_jac, fast_phase, st, ev);
        case 1024: return launch_n<1024>(fb, tab, count, with_jac, fast_phase, st, ev);
        case 2048: return launch_n<2048>(fb, tab, count, with_jac, fast_phase, st, ev);
        case 4096: return launch_n<4096>(fb, tab, count, with_jac, fast_phase, st, ev);
    }
    return -1;
}

}  // namespace ow
