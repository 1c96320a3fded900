// ow_pack_kernels.cu — optional packed output formats for the consumer.
//
// The reference's renderer samples four textures per vertex (grid_tes.glsl:60-64: s_Dy, s_Dx, s_Dz R32F and s_NormalMap
// RGBA32F, allocated at src/main.cpp:1096-1099): 28 B/texel, of which the normal map's y (recoverable from x and z: the
// normal is a unit vector with y > 0, normal_map_cs.glsl:53) and w (constant 1) are redundant. The packed set is two textures:
//   displacement  RGBA32F or RGBA16F   (dx, dy, dz, J)     J = Jacobian when the context computes it, else 1
//   normal_xz     RG16_SNORM           (n.x, n.z)          n.y = sqrt(1 - n.x^2 - n.z^2) in the consumer
// 20 or 12 B/texel. The default formats stay the reference's; this pass runs after the normal kernel only when the context
// was created with OW_FLAG_PACKED_F32 / OW_FLAG_PACKED_F16 (INTEGRATION.md shows the matching grid_tes.glsl decode).
#include <cuda_fp16.h>

#include "ow_internal.h"

namespace ow {

namespace {

__device__ __forceinline__ int snorm16(float v) {                    // GL 4.5 spec 2.3.5.2: round(clamp(v,-1,1) * 32767)
    return __float2int_rn(fminf(fmaxf(v, -1.0f), 1.0f) * 32767.0f);
}
__device__ __forceinline__ uint32_t pack_snorm2(float x, float z) {
    return ((uint32_t)snorm16(x) & 0xffffu) | ((uint32_t)snorm16(z) << 16);
}
__device__ __forceinline__ uint2 pack_half4(float a, float b, float c, float d) {
    const __half2 lo = __floats2half2_rn(a, b), hi = __floats2half2_rn(c, d);
    return make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
}

// One thread packs four adjacent texels of one slot. grid.y = slot entries of the launch group.
template <bool HALF>
__global__ void __launch_bounds__(256) ow_pack_kernel(FrameBuffers fb, SlotTable tab, PackedBuffers pk, int quads) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= quads) return;
    const int slot = tab.slot[blockIdx.y];
    const size_t nn = (size_t)fb.N * fb.N, i = (size_t)q * 4;
    const float* disp = fb.disp + (size_t)slot * 3 * nn;
    const float4 dy = __ldcs(reinterpret_cast<const float4*>(disp + i));
    const float4 dx = __ldcs(reinterpret_cast<const float4*>(disp + nn + i));
    const float4 dz = __ldcs(reinterpret_cast<const float4*>(disp + 2 * nn + i));
    float4 J = make_float4(1.f, 1.f, 1.f, 1.f);
    if (fb.jacobian) J = __ldcs(reinterpret_cast<const float4*>(fb.jacobian + (size_t)slot * nn + i));
    const float4* nrm = fb.normal + (size_t)slot * nn + i;
    const float4 n0 = __ldcs(nrm), n1 = __ldcs(nrm + 1), n2 = __ldcs(nrm + 2), n3 = __ldcs(nrm + 3);
    char* base = pk.base + (size_t)slot * pk.slot_bytes;
    if (HALF) {
        uint4* d = reinterpret_cast<uint4*>(base) + (size_t)q * 2;          // 4 texels x 8 B
        const uint2 a = pack_half4(dx.x, dy.x, dz.x, J.x), b = pack_half4(dx.y, dy.y, dz.y, J.y);
        const uint2 c = pack_half4(dx.z, dy.z, dz.z, J.z), e = pack_half4(dx.w, dy.w, dz.w, J.w);
        __stcs(d, make_uint4(a.x, a.y, b.x, b.y));
        __stcs(d + 1, make_uint4(c.x, c.y, e.x, e.y));
    } else {
        float4* d = reinterpret_cast<float4*>(base) + i;                     // 4 texels x 16 B
        __stcs(d, make_float4(dx.x, dy.x, dz.x, J.x));
        __stcs(d + 1, make_float4(dx.y, dy.y, dz.y, J.y));
        __stcs(d + 2, make_float4(dx.z, dy.z, dz.z, J.z));
        __stcs(d + 3, make_float4(dx.w, dy.w, dz.w, J.w));
    }
    uint4* nx = reinterpret_cast<uint4*>(base + pk.normal_offset) + q;       // 4 texels x 4 B
    __stcs(nx, make_uint4(pack_snorm2(n0.x, n0.z), pack_snorm2(n1.x, n1.z), pack_snorm2(n2.x, n2.z), pack_snorm2(n3.x, n3.z)));
}

}  // namespace

int launch_pack(const FrameBuffers& fb, const SlotTable& tab, int count, const PackedBuffers& pk, Launcher& L) {
    const int quads = (int)((size_t)fb.N * fb.N / 4);
    const dim3 grid((quads + 255) / 256, count);
    if (pk.half) L(ow_pack_kernel<true>, grid, 256, 0, fb, tab, pk, quads);
    else L(ow_pack_kernel<false>, grid, 256, 0, fb, tab, pk, quads);
    if (L.err != cudaSuccess) stash_launch_error(L.err);
    return launches_ok() && L.err == cudaSuccess ? 1 : -1;
}

}  // namespace ow
