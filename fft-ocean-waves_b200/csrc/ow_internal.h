// ow_internal.h — interface between the C-ABI layer (ow_api.cu) and the kernel launchers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ow {

// Per-cascade constants as the kernels want them (wind direction already normalised).
struct CascadeDev {
    float L, wind_speed, wdx, wdy, amplitude, suppression, choppiness, pad;
};

constexpr int kMaxGroup = 128;  // slots per launch group (slot table travels by value in the kernel params: 1.5 KB of the 4 KB)

struct SlotTable {
    int32_t cascade[kMaxGroup];
    float time[kMaxGroup];
    int32_t slot[kMaxGroup];    // which output set each entry writes
};

// All device buffers of a context. Strides are in elements of the pointed-to type.
struct FrameBuffers {
    int N;
    const float4* h0;      // [cascade][N][N]   (h0k.re, h0k.im, h0minusk.re, h0minusk.im)
    const float4* hp;      // [cascade][N/2][N] folded texel pairs (fold_pair in ow_kernels.cuh); what the row kernel streams
    const float4* nyq;     // [cascade][N/2]    Nyquist-column extras (fold_pair_nyq)
    const float* ktab;     // [cascade][N]      k(i) = 2*pi*(i - N/2)/L, computed with the shader's operation order
    const CascadeDev* casc;
    float2* inter;         // [slot][3][N/2][N] row-transformed Hermitian half spectra (dy, dx, dz)
    float* disp;           // [slot][3][N][N]   dy, dx, dz
    float4* normal;        // [slot][N][N]
    float* jacobian;       // [slot][N][N] or nullptr
    int discard_inter;     // column kernel drops the intermediate's lines from L2 after reading them (no DRAM write-back)
    float2* scratch;       // N = A*B decomposition only: one frame of radix-A sums, 12 B/texel (ow_big_kernels.cu)
    int fuse_normals;      // OW_FLAG_FUSED_NORMALS: normal map as the epilogue of the dy column tiles (ow_col_fused_kernel)
    int four_step;         // force the N = A*B line decomposition (ow_big_kernels.cu) on a grid the direct kernels could do
};

bool frame_supported(int N);            // direct kernels (N <= 4096) or the N = A*B decomposition (8192 .. 32768)
// N = A*B line decomposition (ow_big_kernels.cu). forced: the test-only mapping of N = 1024 / 2048 onto it.
bool big_supported(int N, bool forced);
cudaError_t configure_big(int N, bool forced);
int launch_big_frame(const FrameBuffers& fb, const SlotTable& tab, int count, bool with_jacobian, bool fast_phase, cudaStream_t st,
                     cudaEvent_t* ev, bool forced);
// Launch the three frame kernels for `count` table entries. Returns number of kernels launched (<0: error).
// ev (optional): 4 events recorded before the row kernel and after each of the three kernels.
// fast_phase: every |w*t| of this launch is below kFastPhaseLimit, so the SFU sin/cos path is accurate enough.
int launch_frame(const FrameBuffers& fb, const SlotTable& tab, int count, bool with_jacobian, bool fast_phase,
                 cudaStream_t st, cudaEvent_t* ev = nullptr);
constexpr float kFastPhaseLimit = 2.0e4f;
cudaError_t configure_frame_kernels(int N);   // opt-in shared memory sizes; call once per device

// ---- slab decomposition of one grid over `world` GPUs (layouts: SlabRows/SlabSink in ow_kernels.cuh) ----
constexpr int kSlabHalo = 8;       // == ow::kHalo
constexpr int kSlabMaxWorld = 8;   // == ow::kMaxWorld
struct SlabGeom {
    int N, world, rank;
    int PL;   // row pairs per rank   = N / 2 / world
    int XL;   // columns per rank     = N / world
    int XH;   // padded columns       = XL + 2 * kSlabHalo
};
bool slab_supported(int N, int world);
// Row kernel for this rank's pairs; block h of the result goes to sink_base[h] ([PL][3][XH] float2 each).
int launch_slab_rows(const SlabGeom& g, const float4* h0_loc, const float4* hp_loc, const float4* nyq_loc, const float* ktab,
                     float2* const sink_base[kSlabMaxWorld], float t, bool fast_phase, float2* scratch, cudaStream_t st);
// Column kernel on recv[N/2][3][XH] -> disp_loc[3][N][XH], then normals (+ Jacobian when jac != nullptr) for the XL
// interior columns -> normal_loc[N][XL], jac_loc[N][XL]. jac_scale = choppiness * N / (2 L).
int launch_slab_cols(const SlabGeom& g, const float2* recv, float* disp_loc, float4* normal_loc, float* jac_loc, float jac_scale,
                     float2* scratch, cudaStream_t st);

bool big_slab_supported(int N, int world, bool forced);
int launch_big_slab_rows(const SlabGeom& g, const float4* h0_loc, const float4* hp_loc, const float4* nyq_loc, const float* ktab,
                         float2* const sink_base[kSlabMaxWorld], float t, bool fast_phase, float2* scratch, cudaStream_t st, bool forced);
int launch_big_slab_cols(const SlabGeom& g, const float2* recv, float* disp_loc, float4* normal_loc, float* jac_loc, float jac_scale,
                         float2* scratch, cudaStream_t st, bool forced);
// float2 elements of scratch a slab rank needs for the N = A*B decomposition (0 when the direct kernels apply).
size_t slab_scratch_elems(const SlabGeom& g);

// Init-time kernels (ow_init_kernels.cu)
cudaError_t launch_noise_seed(uint8_t* noise /* [4][N][N] */, int N, uint64_t seed, cudaStream_t st);
cudaError_t launch_h0_slab(float4* h0_loc, int N, int p0, int PL, uint64_t seed, const CascadeDev& c, cudaStream_t st);
cudaError_t launch_ktab(float* ktab, int N, float L, cudaStream_t st);
cudaError_t launch_h0(float4* h0, const uint8_t* noise, int noise_w, int noise_h, int N, const CascadeDev& c,
                      cudaStream_t st);
// Fold h0 into the per-pair coefficients the row kernel streams. Full grid: pair p uses rows p and N-p of h0[N][N];
// slab: local rows pl and PL+pl of h0_loc[2*PL][N] (first_pair = rank*PL; pair 0 is skipped in both).
cudaError_t launch_fold(const float4* h0, float4* hp, float4* nyq, int N, cudaStream_t st);
cudaError_t launch_fold_slab(const float4* h0_loc, float4* hp_loc, float4* nyq_loc, int N, int first_pair, int PL, cudaStream_t st);
cudaError_t launch_split_h0(const float4* h0, float* h0k, float* h0minusk, int n, cudaStream_t st);
cudaError_t launch_merge_h0(float4* h0, const float* h0k, const float* h0minusk, int n, cudaStream_t st);

}  // namespace ow
