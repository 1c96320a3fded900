// ow_big_kernels.cu — launchers for grids whose lines do not fit one CTA's shared memory: N = A * B, B <= 4096
// (N = 8192, 16384, 32768: BASELINE config C5 at its quoted size), plus the same code path forced onto small grids
// (OW_FLAG_FOUR_STEP: N = 1024 as 4 x 256, N = 2048 as 4 x 512) so that the tests can compare it with the direct kernels.
// Kernel bodies: "LINES LONGER THAN ONE CTA'S SHARED MEMORY" in ow_kernels.cuh.
#include <algorithm>

#include "ow_frame_kernels.cuh"

// Sub-line length of the production instantiations (N = 8192, 16384, 32768 -> A = N / OW_BIG_B).
#ifndef OW_BIG_B
#define OW_BIG_B 2048
#endif

namespace ow {

namespace {

// true when no launch error is stashed for this thread (a failed cluster launch stashes its error and must fail the call)
bool take_and_restash() {
    const cudaError_t e = take_launch_error();
    if (e == cudaSuccess) return true;
    stash_launch_error(e);
    return false;
}

template <int B, int A>
struct Big {
    static constexpr int N = A * B;
    using C = Cfg<B>;
    using R = typename C::Row;
    using K = typename C::Col;
    static constexpr int G = C::COL_G;
    static constexpr int RMB = C::ROW_MINB, KMB = C::COL_MINB;
    static constexpr int RY = 8, WARPS = 4, NMINB = 4;

    static constexpr int G4 = 4;                      // narrow column tiles of the cluster kernel: 3 CTAs per SM instead of 1
    static constexpr int KMB4 = 3;

    // How many clusters of A CTAs of `kernel` the device can hold at once (0: this cluster shape cannot be scheduled here).
    template <class Kern>
    static int max_clusters(Kern kernel, int threads, size_t smem) {
        if (A > 8 && cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); return 0; }
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(A * 64); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = A; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
        return n;
    }

    static cudaError_t configure(KernelConfig* cfg) {
        cudaError_t e = cudaFuncSetAttribute(ow_bigrow_lines_kernel<R, A, RMB, false, FullRows<N>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_smem<R, 1>());
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(ow_bigrow_lines_kernel<R, A, RMB, true, FullRows<N>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_smem<R, 1>());
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(ow_bigrow_lines_kernel<R, A, RMB, false, SlabRows<N>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_smem<R, 1>());
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(ow_bigrow_lines_kernel<R, A, RMB, true, SlabRows<N>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_smem<R, 1>());
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(ow_bigcol_lines_kernel<K, A, G, KMB, FullColGeom<N>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ColLayout<K, G>::SMEM);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(ow_bigcol_lines_kernel<K, A, G, KMB, SlabColGeom>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ColLayout<K, G>::SMEM);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(ow_bigcol_lines_kernel<K, A, G4, KMB4, FullColGeom<N>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ColLayout<K, G4>::SMEM);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(ow_bigcol_lines_kernel<K, A, G4, KMB4, SlabColGeom>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ColLayout<K, G4>::SMEM);
        if (e != cudaSuccess) return e;
        {   // persistent pipelined column lines kernel: how many CTAs of it an SM holds
            int n = 0, m = 0;
            e = cudaFuncSetAttribute(ow_bigcol_lines_pipe_kernel<K, A, G, KMB, FullColGeom<N>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ColLayout<K, G>::SMEM);
            if (e != cudaSuccess) return e;
            e = cudaFuncSetAttribute(ow_bigcol_lines_pipe_kernel<K, A, G, KMB, SlabColGeom>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ColLayout<K, G>::SMEM);
            if (e != cudaSuccess) return e;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, ow_bigcol_lines_pipe_kernel<K, A, G, KMB, FullColGeom<N>>, K::T * G, ColLayout<K, G>::SMEM);
            if (e != cudaSuccess) return e;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&m, ow_bigcol_lines_pipe_kernel<K, A, G, KMB, SlabColGeom>, K::T * G, ColLayout<K, G>::SMEM);
            if (e != cudaSuccess) return e;
            cfg->bigcol_pipe_ctas = n < m ? n : m;
        }
        // cluster versions: opt in to their shared memory, then ask the device how many clusters it can co-schedule
        const size_t rs = row_smem<R, 1>(), cs8 = ColLayout<K, G>::SMEM, cs4 = ColLayout<K, G4>::SMEM;
#define OW_OPT(kern, bytes) if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))) != cudaSuccess) return e
        OW_OPT((ow_bigrow_cluster_kernel<R, A, RMB, false, FullRows<N>, FullSink<N>>), rs);
        OW_OPT((ow_bigrow_cluster_kernel<R, A, RMB, true, FullRows<N>, FullSink<N>>), rs);
        OW_OPT((ow_bigrow_cluster_kernel<R, A, RMB, false, SlabRows<N>, SlabSink<N>>), rs);
        OW_OPT((ow_bigrow_cluster_kernel<R, A, RMB, true, SlabRows<N>, SlabSink<N>>), rs);
        OW_OPT((ow_bigcol_cluster_kernel<K, A, G, KMB, FullColGeom<N>>), cs8);
        OW_OPT((ow_bigcol_cluster_kernel<K, A, G, KMB, SlabColGeom>), cs8);
        OW_OPT((ow_bigcol_cluster_kernel<K, A, G4, KMB4, FullColGeom<N>>), cs4);
        OW_OPT((ow_bigcol_cluster_kernel<K, A, G4, KMB4, SlabColGeom>), cs4);
#undef OW_OPT
        int rows_ok = max_clusters(ow_bigrow_cluster_kernel<R, A, RMB, true, FullRows<N>, FullSink<N>>, R::T, rs);
        rows_ok = std::min(rows_ok, max_clusters(ow_bigrow_cluster_kernel<R, A, RMB, false, FullRows<N>, FullSink<N>>, R::T, rs));
        rows_ok = std::min(rows_ok, max_clusters(ow_bigrow_cluster_kernel<R, A, RMB, true, SlabRows<N>, SlabSink<N>>, R::T, rs));
        rows_ok = std::min(rows_ok, max_clusters(ow_bigrow_cluster_kernel<R, A, RMB, false, SlabRows<N>, SlabSink<N>>, R::T, rs));
        int c8 = std::min(max_clusters(ow_bigcol_cluster_kernel<K, A, G, KMB, FullColGeom<N>>, K::T * G, cs8),
                          max_clusters(ow_bigcol_cluster_kernel<K, A, G, KMB, SlabColGeom>, K::T * G, cs8));
        int c4 = std::min(max_clusters(ow_bigcol_cluster_kernel<K, A, G4, KMB4, FullColGeom<N>>, K::T * G4, cs4),
                          max_clusters(ow_bigcol_cluster_kernel<K, A, G4, KMB4, SlabColGeom>, K::T * G4, cs4));
        cfg->big_clusters_rows = rows_ok; cfg->big_clusters_cols8 = c8; cfg->big_clusters_cols4 = c4;
        // a cluster shape is used when its CTAs can cover (nearly) every SM: rows hold several CTAs per SM; 16-column tiles hold one per SM
        const int sms = cfg->sm_count;
        int bits = 0;
        if (rows_ok * A >= sms) bits |= 1;
        if (c8 * A * 10 >= sms * 9) bits |= 2;
        else if (c4 * A >= sms) bits |= 2 | 4;
        cfg->big_cluster = bits;
        return cudaSuccess;
    }

    template <class Rows, class Sink>
    static void rows_pass(const Rows& rows, const float* ktab, const float* ktab_sub, int p_first, int npairs_rows, float t, bool fast, float2* scratch,
                          const Sink& sink, cudaStream_t st, int cluster_bits, int post_ctas = 0) {
        if (cluster_bits & 1) {
            cudaError_t e = fast ? launch_cluster(ow_bigrow_cluster_kernel<R, A, RMB, true, Rows, Sink>, dim3(npairs_rows * A), dim3(R::T), row_smem<R, 1>(), st, A,
                                                  rows, ktab, ktab_sub, p_first, t, sink)
                                 : launch_cluster(ow_bigrow_cluster_kernel<R, A, RMB, false, Rows, Sink>, dim3(npairs_rows * A), dim3(R::T), row_smem<R, 1>(), st, A,
                                                  rows, ktab, ktab_sub, p_first, t, sink);
            if (e != cudaSuccess) stash_launch_error(e);
            return;
        }
        if (fast) ow_bigrow_lines_kernel<R, A, RMB, true, Rows><<<npairs_rows * A, R::T, row_smem<R, 1>(), st>>>(rows, ktab, ktab_sub, p_first, t, scratch);
        else ow_bigrow_lines_kernel<R, A, RMB, false, Rows><<<npairs_rows * A, R::T, row_smem<R, 1>(), st>>>(rows, ktab, ktab_sub, p_first, t, scratch);
        if (post_ctas > 0) ow_bigrow_post_slim_kernel<B, A, Sink><<<post_ctas, 256, 0, st>>>(scratch, p_first, npairs_rows, sink);
        else ow_bigrow_post_kernel<B, A, Sink><<<dim3((B + 255) / 256, npairs_rows, 3), 256, 0, st>>>(scratch, p_first, sink);
    }

    template <class Geom>
    static void cols_pass(const float2* src, size_t src_chan, int npairs, float2* scratch, float* dst, size_t dst_chan, const Geom& geom,
                          cudaStream_t st, int cluster_bits, int pipe_grid) {
        const float scale = 0.5f / ((float)N * (float)N);
        if ((cluster_bits & 2) && npairs % ((cluster_bits & 4) ? G4 : G) == 0) {
            cudaError_t e = (cluster_bits & 4)
                ? launch_cluster(ow_bigcol_cluster_kernel<K, A, G4, KMB4, Geom>, dim3(npairs / G4 * A, 3), dim3(K::T * G4), ColLayout<K, G4>::SMEM, st, A,
                                 src, src_chan, dst, dst_chan, scale, geom)
                : launch_cluster(ow_bigcol_cluster_kernel<K, A, G, KMB, Geom>, dim3(npairs / G * A, 3), dim3(K::T * G), ColLayout<K, G>::SMEM, st, A,
                                 src, src_chan, dst, dst_chan, scale, geom);
            if (e != cudaSuccess) stash_launch_error(e);
            return;
        }
        const int total = npairs / G * A * 3;
        if (pipe_grid < 0 && npairs % G4 == 0)          // 8-column tiles: three 256-thread CTAs per SM instead of one 512-thread CTA
            ow_bigcol_lines_kernel<K, A, G4, KMB4, Geom><<<dim3(npairs / G4 * A, 3), K::T * G4, ColLayout<K, G4>::SMEM, st>>>(src, src_chan, npairs, scratch, geom);
        else if (pipe_grid > 0)
            ow_bigcol_lines_pipe_kernel<K, A, G, KMB, Geom><<<total < pipe_grid ? total : pipe_grid, K::T * G, ColLayout<K, G>::SMEM, st>>>(src, src_chan, npairs, scratch,
                                                                                                                                 geom, total);
        else
            ow_bigcol_lines_kernel<K, A, G, KMB, Geom><<<dim3(npairs / G * A, 3), K::T * G, ColLayout<K, G>::SMEM, st>>>(src, src_chan, npairs, scratch, geom);
        ow_bigcol_post_kernel<B, A><<<dim3((npairs + 31) / 32, B / 8, 3), dim3(32, 8), 0, st>>>(scratch, npairs, dst, dst_chan, geom.dst_stride(), scale);
    }

    // One slot entry per call (the scratch holds one frame).
    static int frame(const FrameBuffers& fb, const SlotTable& tab, int count, bool with_jac, bool fast, cudaStream_t st, cudaEvent_t* ev) {
        if (count != 1 || !fb.scratch || !fb.ktab_sub) return -1;
        const int cascade = tab.cascade[0], slot = tab.slot[0];
        const size_t nn = (size_t)N * N;
        if (ev) cudaEventRecord(ev[0], st);
        const FullRows<N> rows{fb.h0 + (size_t)cascade * nn, fb.hp + (size_t)cascade * hp_block_f4(N / 2, N), fb.nyq + (size_t)cascade * (N / 2)};
        float2* inter = fb.inter + (size_t)slot * 3 * (nn / 2);
        rows_pass(rows, fb.ktab + (size_t)cascade * N, fb.ktab_sub + (size_t)cascade * N, 0, N / 2, tab.time[0], fast, fb.scratch, FullSink<N>{inter}, st,
                  fb.big_cluster);
        if (ev) cudaEventRecord(ev[1], st);
        cols_pass(inter, nn / 2, N / 2, fb.scratch, fb.disp + (size_t)slot * 3 * nn, nn, FullColGeom<N>{}, st, fb.big_cluster, fb.bigcol_pipe_grid);
        if (ev) cudaEventRecord(ev[2], st);
        const dim3 ngrid(N / 128, N / (WARPS * RY), 1);
        if (with_jac) ow_normal_kernel<N, true, RY, WARPS, NMINB><<<ngrid, dim3(32, WARPS), 0, st>>>(fb, tab);
        else ow_normal_kernel<N, false, RY, WARPS, NMINB><<<ngrid, dim3(32, WARPS), 0, st>>>(fb, tab);
        if (ev) cudaEventRecord(ev[3], st);
        const int launches = 3 + ((fb.big_cluster & 1) ? 0 : 1) + ((fb.big_cluster & 2) ? 0 : 1);
        return launches_ok() && take_and_restash() ? launches : -1;
    }

    static bool slab_ok(int world) {
        if (world < 1 || world > kMaxWorld || (N / 2) % world) return false;
        const int XL = N / world, XH = XL + 2 * kHalo;
        return XH % (2 * G) == 0 && XL % 128 == 0 && (XL & (XL - 1)) == 0;
    }

    static int slab_rows(const SlabGeom& g, const float4* h0_loc, const float4* hp_loc, const float4* nyq_loc, const float* ktab, const float* ktab_sub,
                         float2* const sink_base[kSlabMaxWorld], float t, bool fast, float2* scratch, cudaStream_t st) {
        if (!ktab_sub) return -1;
        SlabRows<N> rows{h0_loc, hp_loc, nyq_loc, g.rank * g.PL, g.PL};
        SlabSink<N> sink{};
        for (int h = 0; h < g.world; ++h) sink.base[h] = sink_base[h];
        sink.world = g.world; sink.p0 = g.rank * g.PL; sink.XL = g.XL; sink.XH = g.XH;
        sink.xl_shift = 0;
        while ((1 << sink.xl_shift) < g.XL) ++sink.xl_shift;
        rows_pass(rows, ktab, ktab_sub, g.rank * g.PL, g.PL, t, fast, scratch, sink, st, g.big_cluster, g.post_ctas);
        return launches_ok() && take_and_restash() ? ((g.big_cluster & 1) ? 1 : 2) : -1;
    }

    static int slab_cols(const SlabGeom& g, const float2* recv, float* disp_loc, float4* normal_loc, float* jac_loc, float jac_scale,
                         float2* scratch, cudaStream_t st) {
        cols_pass(recv, (size_t)g.XH, g.XH / 2, scratch, disp_loc, (size_t)N * g.XH, SlabColGeom{(size_t)3 * g.XH, (size_t)g.XH}, st, g.big_cluster,
                  g.bigcol_pipe_grid);
        const dim3 ngrid(g.XL / 128, N / (WARPS * RY));
        if (jac_loc) ow_normal_slab_kernel<N, true, RY, WARPS, NMINB><<<ngrid, dim3(32, WARPS), 0, st>>>(disp_loc, normal_loc, jac_loc, g.XL, g.XH, jac_scale);
        else ow_normal_slab_kernel<N, false, RY, WARPS, NMINB><<<ngrid, dim3(32, WARPS), 0, st>>>(disp_loc, normal_loc, nullptr, g.XL, g.XH, 0.f);
        return launches_ok() && take_and_restash() ? ((g.big_cluster & 2) ? 2 : 3) : -1;
    }
};

// (N, forced) -> instantiation. Forced four-step exists for N = 1024 (4 x 256) and N = 2048 (4 x 512).
#define OW_BIG_DISPATCH(N_, forced_, CALL)                                   \
    do {                                                                     \
        if (!(forced_)) {                                                    \
            if ((N_) == 8192) return Big<OW_BIG_B, 8192 / OW_BIG_B>::CALL;   \
            if ((N_) == 16384) return Big<OW_BIG_B, 16384 / OW_BIG_B>::CALL; \
            if ((N_) == 32768) return Big<OW_BIG_B, 32768 / OW_BIG_B>::CALL; \
        } else {                                                             \
            if ((N_) == 1024) return Big<256, 4>::CALL;                      \
            if ((N_) == 2048) return Big<512, 4>::CALL;                      \
        }                                                                    \
    } while (0)

}  // namespace

// Radix A of the line decomposition N = A * B a context of this (N, forced) runs; 0 when it runs the direct kernels.
int big_radix(int N, bool forced) {
    if (forced) return (N == 1024 || N == 2048) ? 4 : 0;
    return (N == 8192 || N == 16384 || N == 32768) ? N / OW_BIG_B : 0;
}

bool big_supported(int N, bool forced) {
    return forced ? (N == 1024 || N == 2048) : (N == 8192 || N == 16384 || N == 32768);
}

cudaError_t configure_big(int N, bool forced, KernelConfig* cfg) {
    OW_BIG_DISPATCH(N, forced, configure(cfg));
    return cudaErrorInvalidValue;
}

int launch_big_frame(const FrameBuffers& fb, const SlotTable& tab, int count, bool with_jac, bool fast, cudaStream_t st, cudaEvent_t* ev,
                     bool forced) {
    OW_BIG_DISPATCH(fb.N, forced, frame(fb, tab, count, with_jac, fast, st, ev));
    return -1;
}

bool big_slab_supported(int N, int world, bool forced) {
    OW_BIG_DISPATCH(N, forced, slab_ok(world));
    return false;
}

int launch_big_slab_rows(const SlabGeom& g, const float4* h0_loc, const float4* hp_loc, const float4* nyq_loc, const float* ktab, const float* ktab_sub,
                         float2* const sink_base[kSlabMaxWorld], float t, bool fast, float2* scratch, cudaStream_t st, bool forced) {
    OW_BIG_DISPATCH(g.N, forced, slab_rows(g, h0_loc, hp_loc, nyq_loc, ktab, ktab_sub, sink_base, t, fast, scratch, st));
    return -1;
}

int launch_big_slab_cols(const SlabGeom& g, const float2* recv, float* disp_loc, float4* normal_loc, float* jac_loc, float jac_scale,
                         float2* scratch, cudaStream_t st, bool forced) {
    OW_BIG_DISPATCH(g.N, forced, slab_cols(g, recv, disp_loc, normal_loc, jac_loc, jac_scale, scratch, st));
    return -1;
}

}  // namespace ow
