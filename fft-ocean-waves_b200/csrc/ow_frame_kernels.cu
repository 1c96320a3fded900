// ow_frame_kernels.cu — __global__ wrappers and launchers of the per-frame kernels (sm_100a).
// Kernel bodies live in ow_kernels.cuh (shared with the CPU emulator used by the tests).
#include "ow_frame_kernels.cuh"

namespace ow {

// ---------------------------------------------------------------------------------------------------
static thread_local cudaError_t t_launch_error = cudaSuccess;
void stash_launch_error(cudaError_t e) { t_launch_error = e; }
cudaError_t take_launch_error() {
    const cudaError_t e = t_launch_error;
    t_launch_error = cudaSuccess;
    return e;
}

template <class K>
static cudaError_t opt_in_smem(K kernel, size_t bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

template <int N>
cudaError_t configure_n(KernelConfig* cfg) {
    using C = Cfg<N>;
    using R = typename C::Row;
    using K = typename C::Col;
    constexpr size_t rs = row_smem<R, C::ROW_PAIRS>(), cs = ColLayout<K, C::COL_G>::SMEM;
    cudaError_t e;
    if ((e = opt_in_smem(ow_row_pipe_kernel<R, C::ROW_PAIRS, C::ROW_MINB, false>, rs)) != cudaSuccess) return e;
    if ((e = opt_in_smem(ow_row_pipe_kernel<R, C::ROW_PAIRS, C::ROW_MINB, true>, rs)) != cudaSuccess) return e;
    if ((e = opt_in_smem(ow_row_kernel<R, C::ROW_PAIRS, C::ROW_MINB, false>, rs)) != cudaSuccess) return e;
    if ((e = opt_in_smem(ow_row_kernel<R, C::ROW_PAIRS, C::ROW_MINB, true>, rs)) != cudaSuccess) return e;
    if ((e = opt_in_smem(ow_row_slab_kernel<R, C::ROW_PAIRS, C::ROW_MINB, false>, rs)) != cudaSuccess) return e;
    if ((e = opt_in_smem(ow_row_slab_kernel<R, C::ROW_PAIRS, C::ROW_MINB, true>, rs)) != cudaSuccess) return e;
    if ((e = opt_in_smem(ow_col_kernel<K, C::COL_G, C::COL_MINB>, cs)) != cudaSuccess) return e;
    if ((e = opt_in_smem(ow_col_slab_kernel<K, C::COL_G, C::COL_MINB>, cs)) != cudaSuccess) return e;
    if ((e = opt_in_smem(ow_col_pipe_kernel<K, C::COL_G, C::COLP_MINB>, cs)) != cudaSuccess) return e;
    {
        int n = 0;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, ow_col_pipe_kernel<K, C::COL_G, C::COLP_MINB>, K::T * C::COL_G, cs)) != cudaSuccess) return e;
        cfg->col_pipe_ctas = n > 0 ? n : 1;
    }
    if constexpr (C::LAT) {
        using RL = typename C::RowL;
        if ((e = opt_in_smem(ow_row_kernel<RL, 1, 2, false>, row_smem<RL, 1>())) != cudaSuccess) return e;
        if ((e = opt_in_smem(ow_row_kernel<RL, 1, 2, true>, row_smem<RL, 1>())) != cudaSuccess) return e;
    }
    constexpr size_t rbs = row_bulk_smem<R, C::ROW_PAIRS>();
    if ((e = opt_in_smem(ow_row_bulk_kernel<R, C::ROW_PAIRS, C::ROW_MINB, false>, rbs)) != cudaSuccess) return e;
    if ((e = opt_in_smem(ow_row_bulk_kernel<R, C::ROW_PAIRS, C::ROW_MINB, true>, rbs)) != cudaSuccess) return e;
    // resident CTAs per SM of the persistent kernels ON THIS DEVICE (their grids)
    for (int fast = 0; fast < 2; ++fast) {
        int n = 0;
        e = fast ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, ow_row_pipe_kernel<R, C::ROW_PAIRS, C::ROW_MINB, true>, R::T * C::ROW_PAIRS, rs)
                 : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, ow_row_pipe_kernel<R, C::ROW_PAIRS, C::ROW_MINB, false>, R::T * C::ROW_PAIRS, rs);
        if (e != cudaSuccess) return e;
        cfg->row_pipe_ctas[fast] = n > 0 ? n : 1;
        e = fast ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, ow_row_bulk_kernel<R, C::ROW_PAIRS, C::ROW_MINB, true>, R::T * C::ROW_PAIRS, rbs)
                 : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, ow_row_bulk_kernel<R, C::ROW_PAIRS, C::ROW_MINB, false>, R::T * C::ROW_PAIRS, rbs);
        if (e != cudaSuccess) return e;
        cfg->row_bulk_ctas[fast] = n > 0 ? n : 1;
    }
    if constexpr (C::COL_FUSE) {
        constexpr size_t c2d = Col2Smem<K, C::COL_G, false>::BYTES, c2s = Col2Smem<K, C::COL_G, true>::BYTES;
        if ((e = opt_in_smem(ow_col2_kernel<K, C::COL_G, C::COL2_MINB, C::NRM_RY, false>, c2d)) != cudaSuccess) return e;
        if ((e = opt_in_smem(ow_col2_kernel<K, C::COL_G, C::COL2_MINB, C::NRM_RY, true>, c2s)) != cudaSuccess) return e;
        int n = 0;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, ow_col2_kernel<K, C::COL_G, C::COL2_MINB, C::NRM_RY, false>, K::T * C::COL_G, c2d)) != cudaSuccess) return e;
        cfg->col2_ctas[0] = n > 0 ? n : 1;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, ow_col2_kernel<K, C::COL_G, C::COL2_MINB, C::NRM_RY, true>, K::T * C::COL_G, c2s)) != cudaSuccess) return e;
        cfg->col2_ctas[1] = n > 0 ? n : 1;
    }
    return cudaSuccess;
}

template <int N>
bool slab_ok(int world) {
    using C = Cfg<N>;
    if (world < 1 || world > kMaxWorld || (N / 2) % world) return false;
    const int PL = N / 2 / world, XL = N / world, XH = XL + 2 * kHalo;
    return PL % C::ROW_PAIRS == 0 && XH % (2 * C::COL_G) == 0 && XL % 128 == 0 && (XL & (XL - 1)) == 0;
}

template <int N>
int slab_rows_n(const SlabGeom& g, const float4* h0_loc, const float4* hp_loc, const float4* nyq_loc, const float* ktab,
                float2* const sink_base[kSlabMaxWorld], float t, bool fast_phase, cudaStream_t st) {
    using C = Cfg<N>;
    using R = typename C::Row;
    SlabRows<N> rows{h0_loc, hp_loc, nyq_loc, g.rank * g.PL, g.PL};
    SlabSink<N> sink{};
    for (int h = 0; h < g.world; ++h) sink.base[h] = sink_base[h];
    sink.world = g.world; sink.p0 = g.rank * g.PL; sink.XL = g.XL; sink.XH = g.XH;
    sink.xl_shift = 0;
    while ((1 << sink.xl_shift) < g.XL) ++sink.xl_shift;
    const dim3 grid(g.PL / C::ROW_PAIRS);
    if (fast_phase)
        ow_row_slab_kernel<R, C::ROW_PAIRS, C::ROW_MINB, true><<<grid, R::T * C::ROW_PAIRS, row_smem<R, C::ROW_PAIRS>(), st>>>(rows, ktab, sink, t);
    else
        ow_row_slab_kernel<R, C::ROW_PAIRS, C::ROW_MINB, false><<<grid, R::T * C::ROW_PAIRS, row_smem<R, C::ROW_PAIRS>(), st>>>(rows, ktab, sink, t);
    return launches_ok() ? 1 : -1;
}

template <int N>
int slab_cols_n(const SlabGeom& g, const float2* recv, float* disp_loc, float4* normal_loc, float* jac_loc, float jac_scale,
                cudaStream_t st) {
    using C = Cfg<N>;
    using K = typename C::Col;
    const float scale = 0.5f / ((float)N * (float)N);
    ow_col_slab_kernel<K, C::COL_G, C::COL_MINB>
        <<<dim3(g.XH / (2 * C::COL_G), 3), K::T * C::COL_G, ColLayout<K, C::COL_G>::SMEM, st>>>(recv, disp_loc, g.XH, scale);
    const dim3 ngrid(g.XL / 128, N / (C::NRM_WARPS * C::NRM_RY));
    if (jac_loc)
        ow_normal_slab_kernel<N, true, C::NRM_RY, C::NRM_WARPS, C::NRM_MINB><<<ngrid, dim3(32, C::NRM_WARPS), 0, st>>>(disp_loc, normal_loc, jac_loc, g.XL, g.XH, jac_scale);
    else
        ow_normal_slab_kernel<N, false, C::NRM_RY, C::NRM_WARPS, C::NRM_MINB><<<ngrid, dim3(32, C::NRM_WARPS), 0, st>>>(disp_loc, normal_loc, nullptr, g.XL, g.XH, 0.f);
    return launches_ok() ? 2 : -1;
}

// Grid of a persistent kernel over `items` work items with at most `resident` CTAs on the device. (Sizing it to whole rounds - 384 CTAs x 2 items
// instead of 444 CTAs for 768 items - was measured and is 3.5 % SLOWER on an 8-cascade C4 step: the extra CTAs' loads in flight are worth more
// than the slots they take from the neighbouring launch groups; profiles/r02_experiments.md.)
static int persistent_grid(int items, int resident) { return items < resident ? items : resident; }

template <int N>
int launch_n(const FrameBuffers& fb, const SlotTable& tab, int count, bool with_jac, bool fast_phase, Launcher& L, cudaEvent_t* ev) {
    using C = Cfg<N>;
    using R = typename C::Row;
    using K = typename C::Col;
    constexpr size_t rs = row_smem<R, C::ROW_PAIRS>(), cs = ColLayout<K, C::COL_G>::SMEM;
    const int fi = fast_phase ? 1 : 0;
    if (ev) cudaEventRecord(ev[0], L.st);
    // ---- spectrum + row IFFT --------------------------------------------------------------------------------------------
    const int row_mode = fb.row_mode ? fb.row_mode : C::ROW_MODE;
    const int n_cta_items = count * (N / 2 / C::ROW_PAIRS);
    L.reads_time = true;
    // one frame, nothing forced: the latency shapes (Cfg<N>::LAT)
    const bool lat = C::LAT && count == 1 && fb.row_mode == 0 && fb.col_mode == 0 && fb.fuse_mode < 0 && fb.latency_shapes;
    const bool wide_rows = C::LAT && fb.latency_shapes == 2 && fb.row_mode == 0;      // tuning: the wide row shape for every launch
    if (lat || wide_rows) {
        using RL = typename C::RowL;
        const dim3 rgrid(N / 2, count);
        if (fast_phase) L(ow_row_kernel<RL, 1, 2, true>, rgrid, RL::T, row_smem<RL, 1>(), fb, tab);
        else L(ow_row_kernel<RL, 1, 2, false>, rgrid, RL::T, row_smem<RL, 1>(), fb, tab);
    } else if (row_mode == 1) {
        const dim3 rgrid(N / 2 / C::ROW_PAIRS, count);
        if (fast_phase) L(ow_row_kernel<R, C::ROW_PAIRS, C::ROW_MINB, true>, rgrid, R::T * C::ROW_PAIRS, rs, fb, tab);
        else L(ow_row_kernel<R, C::ROW_PAIRS, C::ROW_MINB, false>, rgrid, R::T * C::ROW_PAIRS, rs, fb, tab);
    } else if (row_mode == 2) {
        const int resident = fb.sm_count * fb.row_pipe_ctas[fi];
        const int grid = persistent_grid(n_cta_items, resident);
        if (fast_phase) L(ow_row_pipe_kernel<R, C::ROW_PAIRS, C::ROW_MINB, true>, grid, R::T * C::ROW_PAIRS, rs, fb, tab, n_cta_items);
        else L(ow_row_pipe_kernel<R, C::ROW_PAIRS, C::ROW_MINB, false>, grid, R::T * C::ROW_PAIRS, rs, fb, tab, n_cta_items);
    } else {
        constexpr size_t rbs = row_bulk_smem<R, C::ROW_PAIRS>();
        const int resident = fb.sm_count * fb.row_bulk_ctas[fi];
        const int grid = persistent_grid(n_cta_items, resident);
        if (fast_phase) L(ow_row_bulk_kernel<R, C::ROW_PAIRS, C::ROW_MINB, true>, grid, R::T * C::ROW_PAIRS, rbs, fb, tab, n_cta_items);
        else L(ow_row_bulk_kernel<R, C::ROW_PAIRS, C::ROW_MINB, false>, grid, R::T * C::ROW_PAIRS, rbs, fb, tab, n_cta_items);
    }
    if (ev) cudaEventRecord(ev[1], L.st);
    // ---- column IFFT + inversion (+ normal map as its epilogue) ----------------------------------------------------------
    const float scale = 0.5f / ((float)N * (float)N);   // 1/2 from the Hermitian split, 1/N^2 from inversion_cs.glsl:36
    int col_mode = fb.col_mode ? fb.col_mode : C::COL_MODE;
    if (!C::COL_FUSE && col_mode != 4) col_mode = 1;
    if (col_mode == 3 && !fb.inter_tmap) col_mode = 2;
    bool fused = false;
    int launches = 1;
    if (col_mode == 4) {
        const int total = 3 * (N / (2 * C::COL_G)) * count, resident = fb.sm_count * fb.col_pipe_ctas;
        L(ow_col_pipe_kernel<K, C::COL_G, C::COLP_MINB>, persistent_grid(total, resident), K::T * C::COL_G, cs, fb, tab, scale, total);
    }
    if constexpr (C::COL_FUSE) {
        if (col_mode == 2 || col_mode == 3) {
            fused = fb.fuse_mode < 0 ? C::COL_FUSED : fb.fuse_mode != 0;
            const int total = 3 * (N / (2 * C::COL_G)) * count;
            if (col_mode == 3) {
                const int resident = fb.sm_count * fb.col2_ctas[1];
                L(ow_col2_kernel<K, C::COL_G, C::COL2_MINB, C::NRM_RY, true>, persistent_grid(total, resident), K::T * C::COL_G,
                  Col2Smem<K, C::COL_G, true>::BYTES, *static_cast<const CUtensorMap*>(fb.inter_tmap), fb, tab, scale, total, fused ? 1 : 0);
            } else {
                CUtensorMap none{};
                L(ow_col2_kernel<K, C::COL_G, C::COL2_MINB, C::NRM_RY, false>, total, K::T * C::COL_G, Col2Smem<K, C::COL_G, false>::BYTES, none, fb, tab,
                  scale, total, fused ? 1 : 0);
            }
        }
    }
    if (col_mode == 1) L(ow_col_kernel<K, C::COL_G, C::COL_MINB>, dim3(N / (2 * C::COL_G), 3, count), K::T * C::COL_G, cs, fb, tab, scale);
    if (ev) cudaEventRecord(ev[2], L.st);
    // ---- normal map (+ Jacobian) ------------------------------------------------------------------------------------------
    const dim3 ngrid(N / 128, N / (C::NRM_WARPS * C::NRM_RY), count);
    if (!fused && lat && C::NRML_RY != C::NRM_RY) {
        const dim3 lgrid(N / 128, N / (C::NRM_WARPS * C::NRML_RY), count);
        if (with_jac) L(ow_normal_kernel<N, true, C::NRML_RY, C::NRM_WARPS, C::NRM_MINB>, lgrid, dim3(32, C::NRM_WARPS), 0, fb, tab);
        else L(ow_normal_kernel<N, false, C::NRML_RY, C::NRM_WARPS, C::NRM_MINB>, lgrid, dim3(32, C::NRM_WARPS), 0, fb, tab);
        ++launches;
    } else if (!fused) {
        if (with_jac) L(ow_normal_kernel<N, true, C::NRM_RY, C::NRM_WARPS, C::NRM_MINB>, ngrid, dim3(32, C::NRM_WARPS), 0, fb, tab);
        else L(ow_normal_kernel<N, false, C::NRM_RY, C::NRM_WARPS, C::NRM_MINB>, ngrid, dim3(32, C::NRM_WARPS), 0, fb, tab);
        ++launches;
    } else {
        constexpr int SRY = 4;        // rows per seam thread
        L(ow_seam_kernel<N, SRY>, dim3((N / SRY + 127) / 128, N / 16, count), 128, 0, fb, tab);
        ++launches;
        if (with_jac) {
            L(ow_jac_kernel<N, C::NRM_RY, C::NRM_WARPS, C::NRM_MINB>, ngrid, dim3(32, C::NRM_WARPS), 0, fb, tab);
            ++launches;
        }
    }
    if (ev) cudaEventRecord(ev[3], L.st);
    if (L.err != cudaSuccess) stash_launch_error(L.err);
    return launches_ok() && L.err == cudaSuccess ? 1 + launches : -1;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda).
bool make_inter_tensor_map(void* out, const float2* inter, int N, int n_slots) {
    int T = 0;
    switch (N) {
        case 256: T = Cfg<256>::Col::T; break;
        case 512: T = Cfg<512>::Col::T; break;
        case 1024: T = Cfg<1024>::Col::T; break;
        case 2048: T = Cfg<2048>::Col::T; break;
        default: return false;
    }
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
        cudaGetLastError();
        return false;
    }
    const cuuint64_t dims[2] = {(cuuint64_t)2 * N, (cuuint64_t)n_slots * 3 * (N / 2)};      // innermost first: floats per row, rows
    const cuuint64_t strides[1] = {(cuuint64_t)N * sizeof(float2)};                          // bytes between rows
    const cuuint32_t box[2] = {32u, (cuuint32_t)T};                                          // 16 columns of float2 x T rows
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = reinterpret_cast<EncodeFn>(fn)(static_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float2*>(inter), dims,
                                                      strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

size_t hp_block_elems(int npairs, int N) { return hp_block_f4(npairs, N); }

bool frame_supported(int N) { return N == 128 || N == 256 || N == 512 || N == 1024 || N == 2048 || N == 4096 || big_supported(N, false); }

cudaError_t configure_frame_kernels(int N, KernelConfig* cfg) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&cfg->sm_count, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    if (big_supported(N, false)) return configure_big(N, false, cfg);
    if (big_supported(N, true)) {
        if ((e = configure_big(N, true, cfg)) != cudaSuccess) return e;
    }
    if (mega_supported(N)) {
        if ((e = configure_mega(N, cfg)) != cudaSuccess) return e;
    }
    switch (N) {
        case 128: return configure_n<128>(cfg);
        case 256: return configure_n<256>(cfg);
        case 512: return configure_n<512>(cfg);
        case 1024: return configure_n<1024>(cfg);
        case 2048: return configure_n<2048>(cfg);
        case 4096: return configure_n<4096>(cfg);
    }
    return cudaErrorInvalidValue;
}

bool slab_supported(int N, int world) {
    if (big_supported(N, false)) return big_slab_supported(N, world, false);
    switch (N) {
        case 256: return slab_ok<256>(world);
        case 512: return slab_ok<512>(world);
        case 1024: return slab_ok<1024>(world);
        case 2048: return slab_ok<2048>(world);
        case 4096: return slab_ok<4096>(world);
    }
    return false;
}

int launch_slab_rows(const SlabGeom& g, const float4* h0_loc, const float4* hp_loc, const float4* nyq_loc, const float* ktab, const float* ktab_sub,
                     float2* const sink_base[kSlabMaxWorld], float t, bool fast_phase, float2* scratch, cudaStream_t st) {
    if (big_supported(g.N, false)) return launch_big_slab_rows(g, h0_loc, hp_loc, nyq_loc, ktab, ktab_sub, sink_base, t, fast_phase, scratch, st, false);
    switch (g.N) {
        case 256: return slab_rows_n<256>(g, h0_loc, hp_loc, nyq_loc, ktab, sink_base, t, fast_phase, st);
        case 512: return slab_rows_n<512>(g, h0_loc, hp_loc, nyq_loc, ktab, sink_base, t, fast_phase, st);
        case 1024: return slab_rows_n<1024>(g, h0_loc, hp_loc, nyq_loc, ktab, sink_base, t, fast_phase, st);
        case 2048: return slab_rows_n<2048>(g, h0_loc, hp_loc, nyq_loc, ktab, sink_base, t, fast_phase, st);
        case 4096: return slab_rows_n<4096>(g, h0_loc, hp_loc, nyq_loc, ktab, sink_base, t, fast_phase, st);
    }
    return -1;
}

size_t slab_scratch_elems(const SlabGeom& g) {
    if (!big_supported(g.N, false)) return 0;
    const size_t rows = (size_t)g.PL * 3 * g.N, cols = (size_t)3 * g.N * (g.XH / 2);
    return rows > cols ? rows : cols;
}

int launch_slab_cols(const SlabGeom& g, const float2* recv, float* disp_loc, float4* normal_loc, float* jac_loc, float jac_scale,
                     float2* scratch, cudaStream_t st) {
    if (big_supported(g.N, false)) return launch_big_slab_cols(g, recv, disp_loc, normal_loc, jac_loc, jac_scale, scratch, st, false);
    switch (g.N) {
        case 256: return slab_cols_n<256>(g, recv, disp_loc, normal_loc, jac_loc, jac_scale, st);
        case 512: return slab_cols_n<512>(g, recv, disp_loc, normal_loc, jac_loc, jac_scale, st);
        case 1024: return slab_cols_n<1024>(g, recv, disp_loc, normal_loc, jac_loc, jac_scale, st);
        case 2048: return slab_cols_n<2048>(g, recv, disp_loc, normal_loc, jac_loc, jac_scale, st);
        case 4096: return slab_cols_n<4096>(g, recv, disp_loc, normal_loc, jac_loc, jac_scale, st);
    }
    return -1;
}

template <int N>
void modes_n(const FrameBuffers& fb, int* row, int* col, int* fused) {
    using C = Cfg<N>;
    *row = fb.row_mode ? fb.row_mode : C::ROW_MODE;
    int cm = fb.col_mode ? fb.col_mode : C::COL_MODE;
    if (!C::COL_FUSE && cm != 4) cm = 1;
    if (cm == 3 && !fb.inter_tmap) cm = 2;
    *col = cm;
    *fused = (cm == 1 || cm == 4) ? 0 : (fb.fuse_mode < 0 ? (C::COL_FUSED ? 1 : 0) : (fb.fuse_mode != 0 ? 1 : 0));
}

// The kernels launch_frame will actually use for this context (per-N defaults resolved; the line decomposition has its own).
void effective_modes(const FrameBuffers& fb, int* row, int* col, int* fused) {
    *row = *col = 1;
    *fused = 0;
    if (!frame_graphable(fb)) return;
    switch (fb.N) {
        case 128: return modes_n<128>(fb, row, col, fused);
        case 256: return modes_n<256>(fb, row, col, fused);
        case 512: return modes_n<512>(fb, row, col, fused);
        case 1024: return modes_n<1024>(fb, row, col, fused);
        case 2048: return modes_n<2048>(fb, row, col, fused);
        case 4096: return modes_n<4096>(fb, row, col, fused);
    }
}

bool frame_graphable(const FrameBuffers& fb) { return !big_supported(fb.N, false) && !(fb.four_step && big_supported(fb.N, true)); }

int launch_frame(const FrameBuffers& fb, const SlotTable& tab, int count, bool with_jac, bool fast_phase, Launcher& L, cudaEvent_t* ev) {
    if (!frame_graphable(fb)) {
        if (L.plan) return -1;                  // the line decomposition launches on a stream only
        return launch_big_frame(fb, tab, count, with_jac, fast_phase, L.st, ev, !big_supported(fb.N, false));
    }
    if (fb.frame_mode == 1 && mega_supported(fb.N) && !L.plan && fb.mega_sched && fb.mega_ctas > 0) {
        if (ev) cudaEventRecord(ev[0], L.st);
        const int k = launch_mega_frame(fb, tab, count, with_jac, fast_phase, L.st);
        if (ev) { cudaEventRecord(ev[1], L.st); cudaEventRecord(ev[2], L.st); cudaEventRecord(ev[3], L.st); }   // one kernel: all of it under the first slot
        return k;
    }
    switch (fb.N) {
        case 128: return launch_n<128>(fb, tab, count, with_jac, fast_phase, L, ev);
        case 256: return launch_n<256>(fb, tab, count, with_jac, fast_phase, L, ev);
        case 512: return launch_n<512>(fb, tab, count, with_jac, fast_phase, L, ev);
        case 1024: return launch_n<1024>(fb, tab, count, with_jac, fast_phase, L, ev);
        case 2048: return launch_n<2048>(fb, tab, count, with_jac, fast_phase, L, ev);
        case 4096: return launch_n<4096>(fb, tab, count, with_jac, fast_phase, L, ev);
    }
    return -1;
}

}  // namespace ow
