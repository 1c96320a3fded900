// ow_init_kernels.cu — one-time precompute kernels (compiled with -fmad=false so that the fp32 operation
// order below is exactly the shader's; these run once per parameter change, not per frame).
//
//   ow_h0_kernel    = tilde_h0_k_cs.glsl:35-94 (Phillips spectrum x Box-Muller Gaussians), reference call
//                     site src/main.cpp:553-583.
//   ow_ktab_kernel  = the k-vector component of tilde_h0_t_cs.glsl:72-73, tabulated per index.
// The reference's other two init steps (bit-reversal table main.cpp:733-744 and twiddle texture
// twiddle_factors_cs.glsl) have no equivalent: the in-CTA FFT derives twiddles in registers.
#include "ow_internal.h"
#include "ow_kernels.cuh"

namespace ow {

namespace {
constexpr float kPi = 3.1415926535897932384626433832795f;   // "#define M_PI" of every *_cs.glsl
constexpr float kG = 9.81f;                                 // tilde_h0_k_cs.glsl:27

__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

__global__ void ow_ktab_kernel(float* __restrict__ ktab, int N, float L) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) {
        const float x = (float)i - (float)N / 2.0f;    // :72
        ktab[i] = (2.0f * kPi * x) / L;                // :73
    }
}

// tilde_h0_k_cs.glsl:70-94 for texel (ix, iy) given its four uniform noise bytes.
__device__ __forceinline__ float4 h0_texel(int ix, int iy, int N, const CascadeDev& c, uint8_t b0, uint8_t b1, uint8_t b2, uint8_t b3) {
    const float xx = (float)ix - (float)N / 2.0f, xy = (float)iy - (float)N / 2.0f;   // :76
    const float kx = (2.0f * kPi * xx) / c.L, ky = (2.0f * kPi * xy) / c.L;           // :77
    const float L_philips = (c.wind_speed * c.wind_speed) / kG;                       // :78
    float k_mag = sqrtf(kx * kx + ky * ky);                                           // :79
    if (k_mag < 0.00001f) k_mag = 0.00001f;                                           // :81-82
    const float k_mag_sqr = k_mag * k_mag;                                            // :84
    const float sup = expf(-k_mag_sqr * c.suppression * c.suppression);               // :35-38
    // philips_power_spectrum(), :42-45; normalize() acts on the unclamped k: 0 * inf = NaN at k = 0, and
    // fminf(fmaxf(NaN, -4000), 4000) = -4000 (IEEE maxNum), which is what NVIDIA's GL compiler produces too.
    const float rs = 1.0f / sqrtf(kx * kx + ky * ky);
    const float base = c.amplitude * expf(-1.0f / (k_mag_sqr * L_philips * L_philips));
    const float dp = (kx * rs) * c.wdx + (ky * rs) * c.wdy;
    const float dm = (-kx * rs) * c.wdx + (-ky * rs) * c.wdy;
    const float Pp = (base * (dp * dp) * sup) / (k_mag_sqr * k_mag_sqr);
    const float Pm = (base * (dm * dm) * sup) / (k_mag_sqr * k_mag_sqr);
    const float h0k = clampf(sqrtf(Pp) / sqrtf(2.0f), -4000.0f, 4000.0f);             // :87
    const float h0m = clampf(sqrtf(Pm) / sqrtf(2.0f), -4000.0f, 4000.0f);             // :88
    const float n0 = clampf((float)b0 / 255.0f, 0.001f, 1.0f);                        // :53-58
    const float n1 = clampf((float)b1 / 255.0f, 0.001f, 1.0f);
    const float n2 = clampf((float)b2 / 255.0f, 0.001f, 1.0f);
    const float n3 = clampf((float)b3 / 255.0f, 0.001f, 1.0f);
    const float u0 = 2.0f * kPi * n0, v0 = sqrtf(-2.0f * logf(n1));                   // :60-63
    const float u1 = 2.0f * kPi * n2, v1 = sqrtf(-2.0f * logf(n3));
    float4 out;
    out.x = (v0 * cosf(u0)) * h0k;   // :92
    out.y = (v0 * sinf(u0)) * h0k;
    out.z = (v1 * cosf(u1)) * h0m;   // :93
    out.w = (v1 * sinf(u1)) * h0m;
    return out;
}

__global__ void ow_h0_kernel(float4* __restrict__ h0, const uint8_t* __restrict__ noise, int nw, int nh, int N,
                             CascadeDev c) {
    const int ix = blockIdx.x * blockDim.x + threadIdx.x, iy = blockIdx.y * blockDim.y + threadIdx.y;
    if (ix >= N || iy >= N) return;
    // gauss_rnd(), :51-68: texture(noiseJ, gid/N).r with NEAREST + CLAMP_TO_EDGE on an nw x nh RGBA8 image.
    int tx = (int)floorf(((float)ix / (float)N) * (float)nw), ty = (int)floorf(((float)iy / (float)N) * (float)nh);
    tx = min(tx, nw - 1);
    ty = min(ty, nh - 1);
    const size_t plane = (size_t)nw * nh, o = (size_t)ty * nw + tx;
    h0[(size_t)iy * N + ix] = h0_texel(ix, iy, N, c, noise[o], noise[plane + o], noise[2 * plane + o], noise[3 * plane + o]);
}

// Counter-based uniform noise for grids too large to ship noise images for (BASELINE config C5): Philox4x32-10
// (Salmon et al., SC'11) with counter (ix, iy, 0, 0) and key (seed lo, seed hi); plane j's byte for texel
// (ix, iy) is the low byte of output word j. One N x N "noise image" per plane, looked up 1:1.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}

__global__ void ow_noise_seed_kernel(uint8_t* __restrict__ noise /* [4][N][N] */, int N, uint64_t seed) {
    const int ix = blockIdx.x * blockDim.x + threadIdx.x, iy = blockIdx.y * blockDim.y + threadIdx.y;
    if (ix >= N || iy >= N) return;
    const uint4 r = philox4x32_10(make_uint4((uint32_t)ix, (uint32_t)iy, 0u, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const size_t plane = (size_t)N * N, o = (size_t)iy * N + ix;
    noise[o] = (uint8_t)r.x; noise[plane + o] = (uint8_t)r.y; noise[2 * plane + o] = (uint8_t)r.z; noise[3 * plane + o] = (uint8_t)r.w;
}

// Slab-local initial spectrum: local row lr of h0_loc[2*PL][N] holds global row v (see SlabRows in ow_kernels.cuh):
// lr < PL -> v = p0 + lr; otherwise pair j = p0 + lr - PL -> v = N - j (pair 0: N/2). Noise straight from Philox.
__global__ void ow_h0_slab_kernel(float4* __restrict__ h0, int N, int p0, int PL, uint64_t seed, CascadeDev c) {
    const int ix = blockIdx.x * blockDim.x + threadIdx.x, lr = blockIdx.y * blockDim.y + threadIdx.y;
    if (ix >= N || lr >= 2 * PL) return;
    const int j = p0 + lr - PL;
    const int iy = lr < PL ? p0 + lr : (j == 0 ? N / 2 : N - j);
    const uint4 r = philox4x32_10(make_uint4((uint32_t)ix, (uint32_t)iy, 0u, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    h0[(size_t)lr * N + ix] = h0_texel(ix, iy, N, c, (uint8_t)r.x, (uint8_t)r.y, (uint8_t)r.z, (uint8_t)r.w);
}

// hp[pl][u] = fold_pair(h0 at (u, p), h0 at the mirror texel); rowsA/rowsB = first "primary"/"mirror" row of the block,
// mirror rows advance by mirror_step (-N for the full grid where row N-p follows row N-p+1 downwards, +N in a slab).
// sub_A > 0: the row is stored sub-line-major (subline_index) for the line decomposition N = sub_A * B.
__global__ void ow_fold_kernel(const float4* __restrict__ rowsA, const float4* __restrict__ rowsB, long long mirror_step,
                               float4* __restrict__ hp, float4* __restrict__ nyq, const float* __restrict__ ktab, int N, int npairs, int first_pair,
                               int sub_A) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x, pl = blockIdx.y;
    if (u >= N || pl >= npairs) return;
    if (first_pair + pl == 0) return;      // pair 0 (rows 0 and N/2) keeps the unfolded path
    const float4 A = rowsA[(size_t)pl * N + u];
    const float4 B = (rowsB + (long long)pl * mirror_step)[(N - u) & (N - 1)];
    hp[(size_t)pl * N + (sub_A > 0 ? subline_index(u, sub_A, N) : u)] = fold_pair(A, B);
    if (u == 0) nyq[pl] = fold_pair_nyq(A, B);
    if (use_wk(N)) {
        float2* wk = reinterpret_cast<float2*>(hp + (size_t)npairs * N);     // second half of the block (hp_block_f4)
        wk[(size_t)pl * N + u] = dispersion_of(ktab[u], ktab[first_pair + pl]);
    }
}

__global__ void ow_ktab_sub_kernel(const float* __restrict__ ktab, float* __restrict__ ktab_sub, int N, int A) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u < N) ktab_sub[subline_index(u, A, N)] = ktab[u];
}

__global__ void ow_split_h0_kernel(const float4* __restrict__ h0, float2* __restrict__ a, float2* __restrict__ b, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float4 v = h0[i];
        a[i] = make_float2(v.x, v.y);
        b[i] = make_float2(v.z, v.w);
    }
}

__global__ void ow_merge_h0_kernel(float4* __restrict__ h0, const float2* __restrict__ a, const float2* __restrict__ b, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) h0[i] = make_float4(a[i].x, a[i].y, b[i].x, b[i].y);
}
}  // namespace

cudaError_t launch_ktab(float* ktab, int N, float L, cudaStream_t st) {
    ow_ktab_kernel<<<(N + 255) / 256, 256, 0, st>>>(ktab, N, L);
    return cudaGetLastError();
}

cudaError_t launch_ktab_sub(const float* ktab, float* ktab_sub, int N, int A, cudaStream_t st) {
    ow_ktab_sub_kernel<<<(N + 255) / 256, 256, 0, st>>>(ktab, ktab_sub, N, A);
    return cudaGetLastError();
}

cudaError_t launch_h0(float4* h0, const uint8_t* noise, int nw, int nh, int N, const CascadeDev& c, cudaStream_t st) {
    ow_h0_kernel<<<dim3(N / 32, N / 8), dim3(32, 8), 0, st>>>(h0, noise, nw, nh, N, c);
    return cudaGetLastError();
}

cudaError_t launch_noise_seed(uint8_t* noise, int N, uint64_t seed, cudaStream_t st) {
    ow_noise_seed_kernel<<<dim3(N / 32, N / 8), dim3(32, 8), 0, st>>>(noise, N, seed);
    return cudaGetLastError();
}

cudaError_t launch_h0_slab(float4* h0, int N, int p0, int PL, uint64_t seed, const CascadeDev& c, cudaStream_t st) {
    ow_h0_slab_kernel<<<dim3(N / 32, (2 * PL + 7) / 8), dim3(32, 8), 0, st>>>(h0, N, p0, PL, seed, c);
    return cudaGetLastError();
}

cudaError_t launch_fold(const float4* h0, float4* hp, float4* nyq, const float* ktab, int N, int sub_A, cudaStream_t st) {
    // pair p: rows p and N-p; block starts at pair 0 (skipped), mirror of pair pl is row N - pl = rowsB - pl*N with rowsB = row N
    ow_fold_kernel<<<dim3((N + 255) / 256, N / 2), 256, 0, st>>>(h0, h0 + (size_t)N * N, -(long long)N, hp, nyq, ktab, N, N / 2, 0, sub_A);
    return cudaGetLastError();
}

cudaError_t launch_fold_slab(const float4* h0_loc, float4* hp_loc, float4* nyq_loc, const float* ktab, int N, int first_pair, int PL, int sub_A,
                             cudaStream_t st) {
    ow_fold_kernel<<<dim3((N + 255) / 256, PL), 256, 0, st>>>(h0_loc, h0_loc + (size_t)PL * N, (long long)N, hp_loc, nyq_loc, ktab, N, PL, first_pair, sub_A);
    return cudaGetLastError();
}

cudaError_t launch_split_h0(const float4* h0, float* h0k, float* h0minusk, int n, cudaStream_t st) {
    ow_split_h0_kernel<<<(n + 255) / 256, 256, 0, st>>>(h0, reinterpret_cast<float2*>(h0k),
                                                        reinterpret_cast<float2*>(h0minusk), n);
    return cudaGetLastError();
}

cudaError_t launch_merge_h0(float4* h0, const float* h0k, const float* h0minusk, int n, cudaStream_t st) {
    ow_merge_h0_kernel<<<(n + 255) / 256, 256, 0, st>>>(h0, reinterpret_cast<const float2*>(h0k),
                                                        reinterpret_cast<const float2*>(h0minusk), n);
    return cudaGetLastError();
}

}  // namespace ow
