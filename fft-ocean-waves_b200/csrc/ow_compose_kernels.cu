// ow_compose_kernels.cu — multi-cascade composition on the device (cascade blending weights).
//
// The reference renders ONE cascade: its tessellation shader samples the four sim textures with LINEAR/REPEAT filtering
// (src/main.cpp:1142-1144; the texture class defaults for m_dy/m_dx/m_dz) and displaces the vertex (grid_tes.glsl:60-64):
//     pos.y += texture(s_Dy, uv).r * u_DisplacementScale;   pos.x -= texture(s_Dx, uv).r * u_Choppiness;
//     pos.z -= texture(s_Dz, uv).r * u_Choppiness;          normal = texture(s_NormalMap, uv).xyz;
// With several cascades (BASELINE config C4) the consumer sums those terms over the cascades, each sampled at uv = (x, z) / L_c
// (a patch of L_c metres tiles the plane) and scaled by a caller-supplied blending weight w_c (LOD / distance fades). This file is
// that sum as a kernel, for arbitrary world positions (height and normal queries: buoyancy, picking) and for a regular grid of
// them (one combined displacement + normal image per clip-map level):
//     offset = ( -sum w_c l_c Dx_c ,  scale * sum w_c Dy_c ,  -sum w_c l_c Dz_c ),      l_c = the cascade's choppiness
//     normal = normalize( sum w_c n_c.x / n_c.y ,  1 ,  sum w_c n_c.z / n_c.y )         (slopes add, normal_map_cs.glsl:49-53 has n.y > 0)
// Bilinear taps follow the GL rule: texel centres at (i + 0.5)/N, coordinate u*N - 0.5, weights from its fraction, indices wrapped.
// The wrap of the world coordinate is done in fp64 (a handful of operations per point and cascade) so that positions kilometres from
// the origin keep the full fp32 resolution inside their texel.
#include "ow_internal.h"

namespace ow {

namespace {

struct Taps {
    int i0, i1, j0, j1;
    float a, b;          // fractions along x (columns) and z (rows)
};

__device__ __forceinline__ Taps taps_of(float x, float z, double inv_L, int N) {
    double u = (double)x * inv_L, v = (double)z * inv_L;
    u -= floor(u); v -= floor(v);                               // [0, 1)
    const float tu = (float)(u * N) - 0.5f, tv = (float)(v * N) - 0.5f;      // [-0.5, N - 0.5)
    const float fu = floorf(tu), fv = floorf(tv);
    Taps t;
    t.a = tu - fu; t.b = tv - fv;
    const int iu = (int)fu, iv = (int)fv;                       // -1 .. N-1
    t.i0 = iu < 0 ? N - 1 : iu; t.i1 = iu + 1 >= N ? 0 : iu + 1;
    t.j0 = iv < 0 ? N - 1 : iv; t.j1 = iv + 1 >= N ? 0 : iv + 1;
    return t;
}

__device__ __forceinline__ float bilerp(const float* __restrict__ img, int N, const Taps& t) {
    const float* r0 = img + (size_t)t.j0 * N;
    const float* r1 = img + (size_t)t.j1 * N;
    const float top = __ldg(r0 + t.i0) * (1.0f - t.a) + __ldg(r0 + t.i1) * t.a;
    const float bot = __ldg(r1 + t.i0) * (1.0f - t.a) + __ldg(r1 + t.i1) * t.a;
    return top * (1.0f - t.b) + bot * t.b;
}

__device__ __forceinline__ float4 bilerp4(const float4* __restrict__ img, int N, const Taps& t) {
    const float4* r0 = img + (size_t)t.j0 * N;
    const float4* r1 = img + (size_t)t.j1 * N;
    const float4 p = __ldg(r0 + t.i0), q = __ldg(r0 + t.i1), r = __ldg(r1 + t.i0), s = __ldg(r1 + t.i1);
    const float w00 = (1.0f - t.a) * (1.0f - t.b), w10 = t.a * (1.0f - t.b), w01 = (1.0f - t.a) * t.b, w11 = t.a * t.b;
    return make_float4(p.x * w00 + q.x * w10 + r.x * w01 + s.x * w11, p.y * w00 + q.y * w10 + r.y * w01 + s.y * w11,
                       p.z * w00 + q.z * w10 + r.z * w01 + s.z * w11, p.w * w00 + q.w * w10 + r.w * w01 + s.w * w11);
}

__device__ __forceinline__ void blend_at(const ComposeArgs& A, float x, float z, float4* offset, float4* normal) {
    const size_t nn = (size_t)A.N * A.N;
    float ox = 0.f, oy = 0.f, oz = 0.f, wsum = 0.f, sx = 0.f, sz = 0.f;
#pragma unroll 1
    for (int i = 0; i < A.n_terms; ++i) {
        const ComposeTerm tm = A.term[i];
        const Taps t = taps_of(x, z, tm.inv_L, A.N);
        const float* disp = A.disp + (size_t)tm.slot * 3 * nn;
        const float dy = bilerp(disp, A.N, t), dx = bilerp(disp + nn, A.N, t), dz = bilerp(disp + 2 * nn, A.N, t);
        const float4 n = bilerp4(A.normal + (size_t)tm.slot * nn, A.N, t);
        oy += tm.weight * dy;
        ox -= tm.weight * tm.choppiness * dx;
        oz -= tm.weight * tm.choppiness * dz;
        const float iy = 1.0f / n.y;                      // the filtered normal keeps y > 0 (every tap has)
        sx += tm.weight * n.x * iy;
        sz += tm.weight * n.z * iy;
        wsum += tm.weight;
    }
    const float r = rsqrtf(sx * sx + 1.0f + sz * sz);
    *offset = make_float4(ox, A.displacement_scale * oy, oz, wsum);
    *normal = make_float4(sx * r, r, sz * r, 1.0f);
}

__global__ void __launch_bounds__(256) ow_sample_points_kernel(ComposeArgs A, int n_points, const float2* __restrict__ xz, float4* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_points) return;
    const float2 p = __ldg(xz + i);
    float4 o, n;
    blend_at(A, p.x, p.y, &o, &n);
    out[2 * (size_t)i] = o;
    out[2 * (size_t)i + 1] = n;
}

// M x M world positions origin + (i + 0.5) * extent / M: lanes run along x, so a warp's taps of one cascade are runs of adjacent texels.
__global__ void __launch_bounds__(256) ow_compose_grid_kernel(ComposeArgs A, int M, float ox, float oz, float step, float4* __restrict__ out_offset,
                                                              float4* __restrict__ out_normal) {
    const int i = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 8 + threadIdx.y;
    if (i >= M || j >= M) return;
    float4 o, n;
    blend_at(A, ox + ((float)i + 0.5f) * step, oz + ((float)j + 0.5f) * step, &o, &n);
    __stcs(out_offset + (size_t)j * M + i, o);
    __stcs(out_normal + (size_t)j * M + i, n);
}

}  // namespace

cudaError_t launch_sample_points(const ComposeArgs& A, int n_points, const float2* xz, float4* out, cudaStream_t st) {
    if (n_points < 1) return cudaSuccess;
    ow_sample_points_kernel<<<(n_points + 255) / 256, 256, 0, st>>>(A, n_points, xz, out);
    return cudaGetLastError();
}

cudaError_t launch_compose_grid(const ComposeArgs& A, int M, float ox, float oz, float extent, float4* out_offset, float4* out_normal, cudaStream_t st) {
    ow_compose_grid_kernel<<<dim3((M + 31) / 32, (M + 7) / 8), dim3(32, 8), 0, st>>>(A, M, ox, oz, extent / (float)M, out_offset, out_normal);
    return cudaGetLastError();
}

}  // namespace ow
