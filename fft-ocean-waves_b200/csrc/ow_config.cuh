// ow_config.cuh — per-N tuning table shared by the kernels (ow_frame_kernels.cu) and the CPU emulator (tests/emu).
#pragma once
#include "ow_fft.cuh"

namespace ow {

// ---------------------------------------------------------------------------------------------------
// Per-N configuration. Row plans have R2 = 16 so that 16 consecutive butterfly ids of stages 0 and 1 walk
// the unit-stride digit; pads (P1,P0) make stage 2 conflict-free for 8-byte accesses (16-lane phases).
// ROW_PIPE selects the persistent software-pipelined row kernel (ow_row_pipe_kernel) instead of one CTA per ROW_PAIRS
// row pairs; chosen per N from whole-frame throughput in multi-stream sweeps (tools/tune/tune.cu -DTUNE_SWEEP), where a
// variant that wins in isolation does not always win (profiles/r01d_tune_sweep_*.txt).
// COL_FUSE: ow_col_fused_kernel (normal map as the epilogue of the dy column tiles) is available for this N (needs COL_G == 8);
// used only when the context asks for it (OW_FLAG_FUSED_NORMALS): on B200 it is slower than the two separate kernels.
// Normal kernel: NRM_RY output rows per thread walk, NRM_WARPS warps per CTA, NRM_MINB resident CTAs per SM.
// Column plans interleave G jobs in the lane index, need S0 odd and a job stride == 16/G (mod 16).
// ---------------------------------------------------------------------------------------------------
template <int N>
struct Cfg;

template <>
struct Cfg<256> {
    using Row = Plan<256, 4, 4, 16, 32, 1, 0>;
    static constexpr int ROW_PAIRS = 4, ROW_MINB = 4;
    static constexpr bool ROW_PIPE = false;
    using Col = Plan<256, 4, 4, 16, 32, 0, 1>;
    static constexpr int COL_G = 8, COL_MINB = 2;
    static constexpr bool COL_FUSE = true;
    static constexpr int NRM_RY = 8, NRM_WARPS = 4, NRM_MINB = 4;
};
template <>
struct Cfg<512> {
    using Row = Plan<512, 8, 4, 16, 32, 1, 14>;
    static constexpr int ROW_PAIRS = 4, ROW_MINB = 3;
    static constexpr bool ROW_PIPE = false;
    using Col = Plan<512, 8, 4, 16, 32, 0, 1>;
    static constexpr int COL_G = 8, COL_MINB = 2;
    static constexpr bool COL_FUSE = true;
    static constexpr int NRM_RY = 8, NRM_WARPS = 4, NRM_MINB = 4;
};
template <>
struct Cfg<1024> {
    using Row = Plan<1024, 8, 8, 16, 64, 1, 10>;
    static constexpr int ROW_PAIRS = 2, ROW_MINB = 3;
    static constexpr bool ROW_PIPE = true;
    using Col = Plan<1024, 8, 8, 16, 64, 0, 1>;
    static constexpr int COL_G = 8, COL_MINB = 2;
    static constexpr bool COL_FUSE = true;
    static constexpr int NRM_RY = 8, NRM_WARPS = 4, NRM_MINB = 4;
};
template <>
struct Cfg<2048> {
    using Row = Plan<2048, 8, 16, 16, 256, 1, 2>;      // 256 threads per row pair: one stage-0 butterfly per thread
    static constexpr int ROW_PAIRS = 1, ROW_MINB = 3;
    static constexpr bool ROW_PIPE = false;
    using Col = Plan<2048, 8, 16, 16, 64, 0, 1>;
    static constexpr int COL_G = 8, COL_MINB = 1;
    static constexpr bool COL_FUSE = true;
    static constexpr int NRM_RY = 8, NRM_WARPS = 4, NRM_MINB = 4;
};
template <>
struct Cfg<4096> {
    using Row = Plan<4096, 16, 16, 16, 256, 0, 1>;
    static constexpr int ROW_PAIRS = 1, ROW_MINB = 1;
    static constexpr bool ROW_PIPE = false;
    using Col = Plan<4096, 16, 16, 16, 128, 0, 1>;
    static constexpr int COL_G = 4, COL_MINB = 1;
    static constexpr bool COL_FUSE = false;
    static constexpr int NRM_RY = 8, NRM_WARPS = 4, NRM_MINB = 4;
};

template <class P, int G>
struct ColLayout {
    static constexpr int PJ = ((16 / G) - (P::LINE % 16) + 16) % 16;
    static constexpr int SJ = P::LINE + PJ;          // float2 elements between jobs
    static constexpr size_t SMEM = (size_t)G * SJ * sizeof(float2);
};

template <class P, int PAIRS>
constexpr size_t row_smem() { return (size_t)PAIRS * 3 * P::LINE * sizeof(float2); }

}  // namespace ow
