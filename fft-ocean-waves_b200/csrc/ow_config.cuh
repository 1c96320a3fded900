// ow_config.cuh — per-N tuning table shared by the kernels (ow_frame_kernels.cu) and the CPU emulator (tests/emu).
#pragma once
#include "ow_fft.cuh"

namespace ow {

// ---------------------------------------------------------------------------------------------------
// Per-N configuration. Row plans have R2 = 16 so that 16 consecutive butterfly ids of stages 0 and 1 walk
// the unit-stride digit; pads (P1,P0) make stage 2 conflict-free for 8-byte accesses (16-lane phases).
// ROW_MODE: 1 = one CTA per ROW_PAIRS row pairs (ow_row_kernel), 2 = persistent register-pipelined kernel (ow_row_pipe_kernel),
// 3 = persistent kernel with bulk-async staging (ow_row_bulk_kernel); chosen per N from whole-frame throughput (bench.py
// --row-kernel; profiles/r02*), where a variant that wins in isolation does not always win.
// COL_FUSE: ow_col2_kernel (16-column tiles, optional TMA staging, normal map as the epilogue of the dy tiles) is available for
// this N (needs COL_G == 8). COL_MODE: 1 = ow_col_kernel, 2 = ow_col2_kernel with direct loads, 3 = ow_col2_kernel TMA-staged,
// 4 = ow_col_pipe_kernel (persistent, register-pipelined loads).
// COLP_MINB: resident CTAs per SM the register allocation of ow_col_pipe_kernel (COL_MODE 4) aims at.
// COL_FUSED: the normal map comes out of the column kernel (no separate normal kernel) by default.
// Normal kernel: NRM_RY output rows per thread walk, NRM_WARPS warps per CTA, NRM_MINB resident CTAs per SM.
// LAT: a second, LATENCY-oriented shape for launches of ONE frame (the drop-in ow_step of a single cascade): the throughput shapes above
// leave most SMs idle there (N = 512: 64 row CTAs on 148 SMs, 5-7 us per kernel), so a single frame runs RowL (the same radices and
// paddings with LAT_T threads per row pair instead of T, one pair per CTA: identical butterflies, hence identical bits, dealt out over
// more threads) and a normal-map walk of NRML_RY rows per thread.
// Column plans interleave G jobs in the lane index, need S0 odd and a job stride == 16/G (mod 16).
// ---------------------------------------------------------------------------------------------------
#ifndef OW_C2MB_SMALL
#define OW_C2MB_SMALL 3     // resident ow_col2_kernel CTAs per SM the register allocation aims at, N <= 512 (80 registers, 4 B of spills)
#endif
#ifndef OW_C2MB_1024
#define OW_C2MB_1024 1
#endif
// A/B knobs of tools/ variant builds (build.py -D... --tag=...): N = 512 row-pair groups per CTA / padding, column and normal kernel residency
#ifndef OW_COLP_1024
#define OW_COLP_1024 1    // resident ow_col_pipe_kernel CTAs per SM at N = 1024 (2 forces 64 registers: spills)
#endif
#ifndef OW_COLMB_512
#define OW_COLMB_512 2
#endif
#ifndef OW_NRM_MINB
#define OW_NRM_MINB 4
#endif
#ifndef OW_ROW512_T
#define OW_ROW512_T 32
#endif
#ifndef OW_ROW512_PAIRS
#define OW_ROW512_PAIRS 4
#endif
#ifndef OW_ROW512_MINB
#define OW_ROW512_MINB 3
#endif
template <int N>
struct Cfg;

// N = 128: the reference's DISPLACEMENT_MAP_SIZE is a free #define (src/main.cpp:17); the smallest grid whose rows still fill the normal kernel's
// 128-column warp tiles. Not tuned: one warp per row pair, the plain kernels only (no TMA-staged / fused column variant).
template <>
struct Cfg<128> {
    using Row = Plan<128, 2, 4, 16, 32, 1, 0>;
    static constexpr int ROW_PAIRS = 4, ROW_MINB = 4;
    static constexpr int ROW_MODE = 1;
    using Col = Plan<128, 2, 4, 16, 32, 0, 1>;
    static constexpr int COL_G = 8, COL_MINB = 2;
    static constexpr int COL2_MINB = 2;
    static constexpr int COLP_MINB = 2;
    static constexpr bool COL_FUSE = false;
    static constexpr int COL_MODE = 1;
    static constexpr bool COL_FUSED = false;
    static constexpr int NRM_RY = 8, NRM_WARPS = 4, NRM_MINB = 4;
    static constexpr bool LAT = false;
    using RowL = Row;
    static constexpr int NRML_RY = NRM_RY;
};
template <>
struct Cfg<256> {
    using Row = Plan<256, 4, 4, 16, 32, 1, 0>;
    static constexpr int ROW_PAIRS = 4, ROW_MINB = 4;
    static constexpr int ROW_MODE = 1;
    using Col = Plan<256, 4, 4, 16, 32, 0, 1>;
    static constexpr int COL_G = 8, COL_MINB = 2;
    static constexpr int COL2_MINB = OW_C2MB_SMALL;
    static constexpr int COLP_MINB = 2;
    static constexpr bool COL_FUSE = true;
    static constexpr int COL_MODE = 1;
    static constexpr bool COL_FUSED = false;
    static constexpr int NRM_RY = 8, NRM_WARPS = 4, NRM_MINB = 4;
    static constexpr bool LAT = true;
    using RowL = Plan<256, 4, 4, 16, 64, 1, 0>;
    static constexpr int NRML_RY = 2;
};
template <>
struct Cfg<512> {
    using Row = Plan<512, 8, 4, 16, OW_ROW512_T, 1, 14>;
    static constexpr int ROW_PAIRS = OW_ROW512_PAIRS, ROW_MINB = OW_ROW512_MINB;
    static constexpr int ROW_MODE = 1;
    using Col = Plan<512, 8, 4, 16, 32, 0, 1>;
    static constexpr int COL_G = 8, COL_MINB = OW_COLMB_512;
    static constexpr int COL2_MINB = OW_C2MB_SMALL;
    static constexpr int COLP_MINB = 2;
    static constexpr bool COL_FUSE = true;
    static constexpr int COL_MODE = 1;
    static constexpr bool COL_FUSED = false;
    static constexpr int NRM_RY = 8, NRM_WARPS = 4, NRM_MINB = OW_NRM_MINB;
    static constexpr bool LAT = true;
    using RowL = Plan<512, 8, 4, 16, 128, 1, 14>;
    static constexpr int NRML_RY = 2;
};
template <>
struct Cfg<1024> {
    using Row = Plan<1024, 8, 8, 16, 64, 1, 10>;
    static constexpr int ROW_PAIRS = 2, ROW_MINB = 3;
    static constexpr int ROW_MODE = 2;
    using Col = Plan<1024, 8, 8, 16, 64, 0, 1>;
    static constexpr int COL_G = 8, COL_MINB = 2;
    static constexpr int COL2_MINB = OW_C2MB_1024;
    static constexpr int COLP_MINB = OW_COLP_1024;
    static constexpr bool COL_FUSE = true;
    static constexpr int COL_MODE = 1;
    static constexpr bool COL_FUSED = false;
    static constexpr int NRM_RY = 8, NRM_WARPS = 4, NRM_MINB = 4;
    static constexpr bool LAT = true;
    using RowL = Plan<1024, 8, 8, 16, 128, 1, 10>;
    static constexpr int NRML_RY = 2;
};
template <>
struct Cfg<2048> {
    using Row = Plan<2048, 8, 16, 16, 256, 1, 2>;      // 256 threads per row pair: one stage-0 butterfly per thread
    static constexpr int ROW_PAIRS = 1, ROW_MINB = 3;
    static constexpr int ROW_MODE = 1;
    using Col = Plan<2048, 8, 16, 16, 64, 0, 1>;
    static constexpr int COL_G = 8, COL_MINB = 1;
    static constexpr int COL2_MINB = 1;
    static constexpr int COLP_MINB = 1;
    static constexpr bool COL_FUSE = true;
    static constexpr int COL_MODE = 4;
    static constexpr bool COL_FUSED = false;
    static constexpr int NRM_RY = 8, NRM_WARPS = 4, NRM_MINB = 4;
    static constexpr bool LAT = false;
    using RowL = Row;
    static constexpr int NRML_RY = NRM_RY;
};
template <>
struct Cfg<4096> {
    using Row = Plan<4096, 16, 16, 16, 256, 0, 1>;
    static constexpr int ROW_PAIRS = 1, ROW_MINB = 1;
    static constexpr int ROW_MODE = 1;
    using Col = Plan<4096, 16, 16, 16, 128, 0, 1>;
    static constexpr int COL_G = 4, COL_MINB = 1;
    static constexpr int COL2_MINB = 1;
    static constexpr int COLP_MINB = 1;
    static constexpr bool COL_FUSE = false;
    static constexpr int COL_MODE = 1;
    static constexpr bool COL_FUSED = false;
    static constexpr int NRM_RY = 8, NRM_WARPS = 4, NRM_MINB = 4;
    static constexpr bool LAT = false;
    using RowL = Row;
    static constexpr int NRML_RY = NRM_RY;
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
