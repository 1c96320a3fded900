// Micro-benchmark: can the 126 MB L2 of a B200 carry a write->read hand-over of X MB between two kernels (the row->column intermediate
// and the column->normal displacement planes of one frame), and do per-access L2 eviction hints change the answer?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2handover l2handover.cu && ./l2handover
// One "frame" = three kernels on one stream, shaped like the frame kernels at N = 2048 scaled to X:
//   W : read S (0.67 X, the folded spectrum, re-read every frame)    -> write I (X)
//   R : read I (X) in 128-byte pieces, one per row of 16 KB (the column-tile order) or linearly -> write D (X)
//   Nn: read D (X)                                                    -> write O (1.67 X, never re-read)
// Hint modes: 0 none | 1 streams only (S loads evict_first + no L1 allocate, O stores evict_first) | 2 = 1 + I,D stores evict_last,
// I,D loads evict_first | 3 = 1 + I,D stores evict_last, loads unhinted.
// Reported: microseconds per kernel (CUDA events) and the DRAM-equivalent bandwidth on the bytes each kernel touches; a kernel whose reads
// hit L2 shows up as a "bandwidth" above the copy peak.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

enum Pol { PLAIN = 0, FIRST = 1, LAST = 2 };

template <int P>
__device__ __forceinline__ float4 ld(const float4* p, uint64_t pf, uint64_t pl) {
    float4 v;
    if (P == PLAIN) return __ldg(p);
    const uint64_t pol = P == FIRST ? pf : pl;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}
template <int P>
__device__ __forceinline__ void st(float4* p, float4 v, uint64_t pf, uint64_t pl) {
    if (P == PLAIN) { *p = v; return; }
    const uint64_t pol = P == FIRST ? pf : pl;
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void policies(uint64_t& pf, uint64_t& pl) {
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pf));
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pl));
}

// Linear streaming: n_in float4 read, n_out float4 written (grid-stride, 4 loads in flight per thread).
template <int PL_, int PS_>
__global__ void __launch_bounds__(256) k_lin(const float4* __restrict__ in, size_t n_in, float4* __restrict__ out, size_t n_out) {
    uint64_t pf, pl;
    policies(pf, pl);
    const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nt = gridDim.x * (size_t)blockDim.x;
    float4 acc = make_float4(0, 0, 0, 0);
    size_t i = tid;
    for (; i + 3 * nt < n_in; i += 4 * nt) {
        const float4 a = ld<PL_>(in + i, pf, pl), b = ld<PL_>(in + i + nt, pf, pl), c = ld<PL_>(in + i + 2 * nt, pf, pl), d = ld<PL_>(in + i + 3 * nt, pf, pl);
        acc.x += a.x + b.x + c.x + d.x; acc.y += a.y + b.y + c.y + d.y; acc.z += a.z + b.z + c.z + d.z; acc.w += a.w + b.w + c.w + d.w;
    }
    for (; i < n_in; i += nt) { const float4 a = ld<PL_>(in + i, pf, pl); acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w; }
    for (size_t j = tid; j < n_out; j += nt) st<PS_>(out + j, acc, pf, pl);
}

// Column-tile order: the source is [rows][1024] float4 (16 KB rows); tile t = float4 columns [8t, 8t+8) of every row (128 bytes per row).
// One CTA per tile at a time (persistent over tiles), 8 lanes per row piece, 32 rows per CTA pass; output written linearly per tile.
template <int PL_, int PS_>
__global__ void __launch_bounds__(256) k_tiles(const float4* __restrict__ in, int rows, float4* __restrict__ out) {
    uint64_t pf, pl;
    policies(pf, pl);
    const int lane8 = threadIdx.x & 7, r0 = threadIdx.x >> 3;
    const int q = blockIdx.y, rq = rows / 4;                      // four CTAs per tile, a quarter of the rows each
    for (int t = blockIdx.x; t < 128; t += gridDim.x) {
        float4 acc = make_float4(0, 0, 0, 0);
        for (int r = q * rq + r0; r < (q + 1) * rq; r += 128) {
            const float4 a = ld<PL_>(in + (size_t)r * 1024 + 8 * t + lane8, pf, pl), b = ld<PL_>(in + (size_t)(r + 32) * 1024 + 8 * t + lane8, pf, pl),
                         c = ld<PL_>(in + (size_t)(r + 64) * 1024 + 8 * t + lane8, pf, pl), d = ld<PL_>(in + (size_t)(r + 96) * 1024 + 8 * t + lane8, pf, pl);
            acc.x += a.x + b.x + c.x + d.x; acc.y += a.y + b.y + c.y + d.y; acc.z += a.z + b.z + c.z + d.z; acc.w += a.w + b.w + c.w + d.w;
        }
        float4* o = out + ((size_t)t * 4 + q) * rq * 8;
        for (int j = threadIdx.x; j < rq * 8; j += 256) st<PS_>(o + j, acc, pf, pl);
    }
}

template <int MODE>
static void frame(const float4* S, size_t nS, float4* I, float4* D, float4* O, size_t nX, int rows, bool tiles, int grid, cudaEvent_t* e) {
    constexpr int sL = MODE >= 1 ? FIRST : PLAIN;                 // S loads
    constexpr int xS = MODE >= 2 ? LAST : PLAIN;                  // I, D stores
    constexpr int xL = MODE == 2 ? FIRST : PLAIN;                 // I, D loads
    constexpr int oS = MODE >= 1 ? FIRST : PLAIN;                 // O stores
    cudaEventRecord(e[0]);
    k_lin<sL, xS><<<grid, 256>>>(S, nS, I, nX);
    cudaEventRecord(e[1]);
    if (tiles) k_tiles<xL, xS><<<dim3(128, 4), 256>>>(I, rows, D);
    else k_lin<xL, xS><<<grid, 256>>>(I, nX, D, nX);
    cudaEventRecord(e[2]);
    k_lin<xL, oS><<<grid, 256>>>(D, nX, O, nX * 5 / 3);
    cudaEventRecord(e[3]);
}

int main(int argc, char** argv) {
    const size_t MB = 1 << 20;
    float4 *S, *I, *D, *O;
    cudaMalloc(&S, 64 * MB); cudaMalloc(&I, 96 * MB); cudaMalloc(&D, 96 * MB); cudaMalloc(&O, 160 * MB);
    cudaMemset(S, 0, 64 * MB);
    cudaEvent_t e[4];
    for (auto& x : e) cudaEventCreate(&x);
    const int grid = 148 * 8;
    const double xs[] = {12, 24, 36, 48, 64, 80};
    for (int tiles = 0; tiles < 2; ++tiles)
        for (double xmb : xs) {
            const int rows = (int)(xmb * MB / 16384) / 512 * 512;            // 16 KB rows, a multiple of 512
            const size_t nX = (size_t)rows * 1024, nS = nX * 2 / 3;
            for (int mode = 0; mode < 4; ++mode) {
                float us[3] = {0, 0, 0};
                const int reps = 20;
                for (int r = -3; r < reps; ++r) {
                    switch (mode) {
                        case 0: frame<0>(S, nS, I, D, O, nX, rows, tiles, grid, e); break;
                        case 1: frame<1>(S, nS, I, D, O, nX, rows, tiles, grid, e); break;
                        case 2: frame<2>(S, nS, I, D, O, nX, rows, tiles, grid, e); break;
                        default: frame<3>(S, nS, I, D, O, nX, rows, tiles, grid, e); break;
                    }
                    cudaEventSynchronize(e[3]);
                    if (r < 0) continue;
                    for (int i = 0; i < 3; ++i) { float ms; cudaEventElapsedTime(&ms, e[i], e[i + 1]); us[i] += ms * 1e3f / reps; }
                }
                const double x = nX * 16.0 / 1e6, mb[3] = {x * 5 / 3, 2 * x, x * 8 / 3};
                printf("%s X=%5.1f MB mode %d: W %6.2f us (%5.0f GB/s)  R %6.2f us (%5.0f GB/s)  N %6.2f us (%5.0f GB/s)  frame %6.2f us (%5.0f GB/s on %5.1f MB)\n",
                       tiles ? "tiles " : "linear", x, mode, us[0], mb[0] / us[0] * 1e3, us[1], mb[1] / us[1] * 1e3, us[2], mb[2] / us[2] * 1e3,
                       us[0] + us[1] + us[2], (mb[0] + mb[1] + mb[2]) / (us[0] + us[1] + us[2]) * 1e3, mb[0] + mb[1] + mb[2]);
            }
        }
    const cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(err)); return 1; }
    return 0;
}
