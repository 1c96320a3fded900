// ow_internal.h — interface between the C-ABI layer (ow_api.cu) and the kernel launchers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <memory>
#include <tuple>
#include <type_traits>
#include <vector>

namespace ow {

// Per-cascade constants as the kernels want them (wind direction already normalised).
struct CascadeDev {
    float L, wind_speed, wdx, wdy, amplitude, suppression, choppiness, pad;
};

constexpr int kMaxGroup = 128;  // slots per launch group (slot table travels by value in the kernel params: 1.5 KB of the 4 KB)

constexpr int kMegaSchedInts = 1 + 2 * kMaxGroup;   // ow_mega_kernel's queue head + per-entry rows-done and columns-done counters

struct SlotTable {
    int32_t cascade[kMaxGroup];
    float time[kMaxGroup];
    int32_t slot[kMaxGroup];    // which output set each entry writes
};

// All device buffers of a context. Strides are in elements of the pointed-to type.
struct FrameBuffers {
    int N;
    const float4* h0;      // [cascade][N][N]   (h0k.re, h0k.im, h0minusk.re, h0minusk.im)
    const float4* hp;      // [cascade] blocks of [N/2][N] float4 folded texel pairs (fold_pair) + [N/2][N] float2 (w, 1/|k|): what the row kernel streams
    const float4* nyq;     // [cascade][N/2]    Nyquist-column extras (fold_pair_nyq)
    const float* ktab;     // [cascade][N]      k(i) = 2*pi*(i - N/2)/L, computed with the shader's operation order
    const float* ktab_sub; // [cascade][N]      the same table sub-line-major (subline_index) - contexts that run the N = A*B line decomposition only,
                           //                   whose folded rows `hp` are stored in that order too; nullptr otherwise
    const CascadeDev* casc;
    float2* inter;         // [slot][3][N/2][N] row-transformed Hermitian half spectra (dy, dx, dz)
    float* disp;           // [slot][3][N][N]   dy, dx, dz
    float4* normal;        // [slot][N][N]
    float* jacobian;       // [slot][N][N] or nullptr
    int discard_inter;     // column kernel drops the intermediate's lines from L2 after reading them (no DRAM write-back)
    float2* scratch;       // N = A*B decomposition only: one frame of radix-A sums, 12 B/texel (ow_big_kernels.cu)
    int four_step;         // force the N = A*B line decomposition (ow_big_kernels.cu) on a grid the direct kernels could do
    int row_mode;          // 0 = the per-N choice Cfg<N>::ROW_MODE, 1 = one CTA per ROW_PAIRS row pairs, 2 = persistent register-pipelined kernel,
                           // 3 = persistent kernel with bulk-async (cp.async.bulk + mbarrier) staging of the spectrum rows (ow_set_row_kernel)
    int col_mode;          // 0 = the per-N choice Cfg<N>::COL_MODE, 1 = ow_col_kernel, 2 = ow_col2_kernel with direct loads, 3 = ow_col2_kernel with TMA staging,
                           // 4 = ow_col_pipe_kernel (persistent, next tile's first load batch in flight in registers)
    int fuse_mode;         // -1 = the per-N choice Cfg<N>::COL_FUSED, 0 = separate normal kernel, 1 = normal map as the column kernel's epilogue (col_mode 2/3)
    int* seam;             // [slot][N/16] arrival counters of the seams between neighbouring dy tiles (fused normal map), all zero between launches
    const void* inter_tmap;   // host pointer to the CUtensorMap over `inter` ([n_slots*3*N/2 rows][2N floats], box = 32 floats x T rows), or nullptr
    int sm_count;          // of the context's device (grid size of the persistent kernels)
    int row_pipe_ctas[2];  // resident CTAs per SM of the persistent row kernels on that device: [exact sincos, fast sincos]
    int row_bulk_ctas[2];
    int col2_ctas[2];      // resident CTAs per SM of ow_col2_kernel: [direct loads, TMA staged]
    int col_pipe_ctas;     // resident CTAs per SM of ow_col_pipe_kernel
    int frame_mode;        // 0 = separate row / column / normal kernels, 1 = ow_mega_kernel (one persistent dataflow kernel per launch group; N <= 1024)
    int* mega_sched;       // kMegaSchedInts counters of the launch group being submitted (one area per stream of the context)
    int mega_ctas;         // resident CTAs per SM of ow_mega_kernel
    int latency_shapes;    // launches of ONE frame use the latency-oriented kernel shapes (Cfg<N>::LAT); ow_set_latency_shapes
    int big_cluster;       // N = A*B decomposition: bit 0 = rows, bit 1 = columns run as thread-block clusters (DSMEM radix-A stage, no scratch),
                           // bit 2 = the column clusters use 8-column tiles (3 CTAs per SM) instead of 16-column ones (1 CTA per SM)
    int bigcol_pipe_grid;  // N = A*B decomposition: grid of the persistent column lines kernel (resident CTAs of the device), 0 = one CTA per item
};

// What configure_frame_kernels found out about the device the calling context lives on (kept per context: no process-global state).
struct KernelConfig {
    int sm_count = 148;
    int row_pipe_ctas[2] = {1, 1};
    int row_bulk_ctas[2] = {1, 1};     // ow_row_bulk_kernel, [exact, fast]
    int col2_ctas[2] = {1, 1};         // ow_col2_kernel, [direct loads, TMA staged]
    int col_pipe_ctas = 1;             // ow_col_pipe_kernel
    int mega_ctas = 0;                 // ow_mega_kernel (0: not available for this N)
    int big_cluster = 0;               // what the device can co-schedule (FrameBuffers::big_cluster bits)
    int big_clusters_rows = 0, big_clusters_cols8 = 0, big_clusters_cols4 = 0;   // cudaOccupancyMaxActiveClusters of the three cluster shapes
    int bigcol_pipe_ctas = 0;          // resident CTAs per SM of ow_bigcol_lines_pipe_kernel (0: not a line-decomposition grid)
};

// ---------------------------------------------------------------------------------------------------
// Where a launcher's kernels go: straight onto a stream, or into a CUDA graph under construction (the single-frame path
// of ow_step: the three kernels of a frame become one cudaGraphLaunch, and only the time inside the row kernel's slot table
// is patched per frame). In graph mode consecutive launches form a dependency chain; begin_chain() starts a new,
// independent chain (one per launch group, so groups still overlap the way they do on the auxiliary streams).
// ---------------------------------------------------------------------------------------------------
struct GraphNodeRec {
    cudaGraphNode_t node = nullptr;
    cudaKernelNodeParams params{};
    std::shared_ptr<void> blob;        // owns the argument tuple
    std::vector<void*> ptrs;           // params.kernelParams
    SlotTable* tab = nullptr;          // the SlotTable argument inside the blob (kernels that read the frame time), else nullptr
};

struct GraphPlan {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    std::vector<GraphNodeRec> nodes;
    int launches = 0, groups = 0;
    ~GraphPlan() {
        if (exec) cudaGraphExecDestroy(exec);
        if (graph) cudaGraphDestroy(graph);
    }
};

struct Launcher {
    cudaStream_t st = nullptr;
    GraphPlan* plan = nullptr;         // non-null: append kernel nodes instead of launching
    cudaGraphNode_t prev = nullptr;
    cudaError_t err = cudaSuccess;
    bool reads_time = false;           // set before launching a kernel that reads SlotTable::time: its node is patched per frame

    explicit Launcher(cudaStream_t s) : st(s) {}
    explicit Launcher(GraphPlan* p) : plan(p) {}
    void begin_chain() { prev = nullptr; }

    template <class... KArgs, class... Args>
    void operator()(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args) {
        static_assert(sizeof...(KArgs) == sizeof...(Args), "argument count");
        const bool timed = reads_time;
        reads_time = false;
        if (!plan) {
            kernel<<<grid, block, smem, st>>>(static_cast<KArgs>(args)...);
            return;
        }
        using Tup = std::tuple<std::decay_t<KArgs>...>;
        auto blob = std::make_shared<Tup>(static_cast<KArgs>(args)...);
        GraphNodeRec rec;
        rec.blob = blob;
        std::apply([&](auto&... a) {
            (rec.ptrs.push_back(static_cast<void*>(&a)), ...);
            ((rec.tab = pick_tab(&a, rec.tab)), ...);
        }, *blob);
        if (!timed) rec.tab = nullptr;
        rec.params.func = reinterpret_cast<void*>(kernel);
        rec.params.gridDim = grid;
        rec.params.blockDim = block;
        rec.params.sharedMemBytes = (unsigned)smem;
        rec.params.kernelParams = rec.ptrs.data();
        rec.params.extra = nullptr;
        const cudaError_t e = cudaGraphAddKernelNode(&rec.node, plan->graph, prev ? &prev : nullptr, prev ? 1 : 0, &rec.params);
        if (e != cudaSuccess && err == cudaSuccess) err = e;
        prev = rec.node;
        plan->nodes.push_back(std::move(rec));
        plan->nodes.back().params.kernelParams = plan->nodes.back().ptrs.data();
    }

private:
    static SlotTable* pick_tab(SlotTable* a, SlotTable* cur) { return cur ? cur : a; }
    template <class T>
    static SlotTable* pick_tab(T*, SlotTable* cur) { return cur; }
};

// float4 elements of one block of folded pair rows (npairs rows of N texels): hp_block_f4 in ow_kernels.cuh
size_t hp_block_elems(int npairs, int N);
bool frame_supported(int N);            // direct kernels (N <= 4096) or the N = A*B decomposition (8192 .. 32768)
// N = A*B line decomposition (ow_big_kernels.cu). forced: the test-only mapping of N = 1024 / 2048 onto it.
bool big_supported(int N, bool forced);
int big_radix(int N, bool forced);      // A of N = A*B for such a grid, else 0
cudaError_t configure_big(int N, bool forced, KernelConfig* cfg);
int launch_big_frame(const FrameBuffers& fb, const SlotTable& tab, int count, bool with_jacobian, bool fast_phase, cudaStream_t st,
                     cudaEvent_t* ev, bool forced);
// Launch with cluster dimension (csize, 1, 1) (cudaLaunchKernelEx); the error also lands in the thread's launch-error stash.
template <class... KArgs, class... Args>
cudaError_t launch_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int csize, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// Launch the three frame kernels for `count` table entries. Returns number of kernels launched (<0: error).
// ev (optional): 4 events recorded before the row kernel and after each of the three kernels.
// fast_phase: every |w*t| of this launch is below kFastPhaseLimit, so the SFU sin/cos path is accurate enough.
int launch_frame(const FrameBuffers& fb, const SlotTable& tab, int count, bool with_jacobian, bool fast_phase,
                 Launcher& L, cudaEvent_t* ev = nullptr);
constexpr float kFastPhaseLimit = 2.0e4f;
bool frame_graphable(const FrameBuffers& fb);
// The whole frame as one persistent kernel (ow_mega_kernels.cu).
bool mega_supported(int N);
cudaError_t configure_mega(int N, KernelConfig* cfg);
int launch_mega_frame(const FrameBuffers& fb, const SlotTable& tab, int count, bool with_jacobian, bool fast_phase, cudaStream_t st);
void effective_modes(const FrameBuffers& fb, int* row, int* col, int* fused);   // launch_frame can build graph nodes for this context (direct kernels, N <= 4096)
// Tensor map over the row->column intermediate for the TMA-staged column kernel (64 bytes, 64-byte aligned, at `out`).
// Returns false when the driver entry point is missing or N has no staged kernel; the context then runs without staging.
bool make_inter_tensor_map(void* out, const float2* inter, int N, int n_slots);
cudaError_t configure_frame_kernels(int N, KernelConfig* cfg);   // opt-in shared memory sizes + occupancy queries; call once per context
// The launchers below return a launch count (< 0: error) and have then already consumed the CUDA error; it is kept here
// (per thread) so the API layer can report what actually went wrong. cudaSuccess when the failure was not a CUDA error.
cudaError_t take_launch_error();
void stash_launch_error(cudaError_t e);
// cudaGetLastError() wrapper for the launchers: true when the launches since the last check went through.
inline bool launches_ok() {
    const cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return true;
    stash_launch_error(e);
    return false;
}

// ---- slab decomposition of one grid over `world` GPUs (layouts: SlabRows/SlabSink in ow_kernels.cuh) ----
constexpr int kSlabHalo = 8;       // == ow::kHalo
constexpr int kSlabMaxWorld = 8;   // == ow::kMaxWorld
struct SlabGeom {
    int N, world, rank;
    int PL;   // row pairs per rank   = N / 2 / world
    int XL;   // columns per rank     = N / world
    int XH;   // padded columns       = XL + 2 * kSlabHalo
    int big_cluster = 0;   // FrameBuffers::big_cluster bits (KernelConfig::big_cluster of the rank's device, or 0 to force the scratch path)
    int post_ctas = 0;     // N > 4096: grid of the row post kernel (0 = one CTA per item; > 0 = slim persistent grid, see ow_bigrow_post_slim_kernel)
    int bigcol_pipe_grid = 0;   // N > 4096: grid of the persistent column lines kernel (0 = one CTA per item, ow_bigcol_lines_kernel)
};
bool slab_supported(int N, int world);
// Row kernel for this rank's pairs; block h of the result goes to sink_base[h] ([PL][3][XH] float2 each).
// ktab_sub: the sub-line-major k table (N > 4096, where hp_loc is sub-line-major too), else unused.
int launch_slab_rows(const SlabGeom& g, const float4* h0_loc, const float4* hp_loc, const float4* nyq_loc, const float* ktab, const float* ktab_sub,
                     float2* const sink_base[kSlabMaxWorld], float t, bool fast_phase, float2* scratch, cudaStream_t st);
// Column kernel on recv[N/2][3][XH] -> disp_loc[3][N][XH], then normals (+ Jacobian when jac != nullptr) for the XL
// interior columns -> normal_loc[N][XL], jac_loc[N][XL]. jac_scale = choppiness * N / (2 L).
int launch_slab_cols(const SlabGeom& g, const float2* recv, float* disp_loc, float4* normal_loc, float* jac_loc, float jac_scale,
                     float2* scratch, cudaStream_t st);

bool big_slab_supported(int N, int world, bool forced);
int launch_big_slab_rows(const SlabGeom& g, const float4* h0_loc, const float4* hp_loc, const float4* nyq_loc, const float* ktab, const float* ktab_sub,
                         float2* const sink_base[kSlabMaxWorld], float t, bool fast_phase, float2* scratch, cudaStream_t st, bool forced);
int launch_big_slab_cols(const SlabGeom& g, const float2* recv, float* disp_loc, float4* normal_loc, float* jac_loc, float jac_scale,
                         float2* scratch, cudaStream_t st, bool forced);
// float2 elements of scratch a slab rank needs for the N = A*B decomposition (0 when the direct kernels apply).
size_t slab_scratch_elems(const SlabGeom& g);

// ---- packed output formats (ow_pack_kernels.cu) ----
struct PackedBuffers {
    char* base;            // [slot][slot_bytes]: displacement image first, normal_xz image at normal_offset
    size_t slot_bytes, normal_offset;
    int half;              // displacement texels are RGBA16F (8 B) instead of RGBA32F (16 B)
};
int launch_pack(const FrameBuffers& fb, const SlotTable& tab, int count, const PackedBuffers& pk, Launcher& L);

// Multi-cascade composition (ow_compose_kernels.cu): term i samples output slot `slot` (whose cascade has patch
// size 1/inv_L and the given choppiness) and contributes with `weight`.
constexpr int kMaxComposeTerms = 16;
struct ComposeTerm {
    double inv_L;
    int slot;
    float weight, choppiness;
};
struct ComposeArgs {
    const float* disp;        // [slot][3][N][N]
    const float4* normal;     // [slot][N][N]
    int N, n_terms;
    float displacement_scale;
    ComposeTerm term[kMaxComposeTerms];
};
cudaError_t launch_sample_points(const ComposeArgs& A, int n_points, const float2* xz, float4* out /* [n_points][2] */, cudaStream_t st);
cudaError_t launch_compose_grid(const ComposeArgs& A, int M, float ox, float oz, float extent, float4* out_offset, float4* out_normal, cudaStream_t st);

// Init-time kernels (ow_init_kernels.cu)
cudaError_t launch_noise_seed(uint8_t* noise /* [4][N][N] */, int N, uint64_t seed, cudaStream_t st);
cudaError_t launch_h0_slab(float4* h0_loc, int N, int p0, int PL, uint64_t seed, const CascadeDev& c, cudaStream_t st);
cudaError_t launch_ktab(float* ktab, int N, float L, cudaStream_t st);
cudaError_t launch_ktab_sub(const float* ktab, float* ktab_sub, int N, int A, cudaStream_t st);    // ktab_sub[subline_index(u, A, N)] = ktab[u]
cudaError_t launch_h0(float4* h0, const uint8_t* noise, int noise_w, int noise_h, int N, const CascadeDev& c,
                      cudaStream_t st);
// Fold h0 into the per-pair coefficients the row kernel streams. Full grid: pair p uses rows p and N-p of h0[N][N];
// slab: local rows pl and PL+pl of h0_loc[2*PL][N] (first_pair = rank*PL; pair 0 is skipped in both).
// hp / hp_loc are blocks of hp_block_f4(npairs, N) float4: the fold coefficients, then (w, 1/|k|) per pair texel (ktab: this cascade's k table).
// sub_A > 0: every folded row is written sub-line-major for the N = sub_A * B line decomposition (texel u at subline_index(u, sub_A, N)).
cudaError_t launch_fold(const float4* h0, float4* hp, float4* nyq, const float* ktab, int N, int sub_A, cudaStream_t st);
cudaError_t launch_fold_slab(const float4* h0_loc, float4* hp_loc, float4* nyq_loc, const float* ktab, int N, int first_pair, int PL, int sub_A,
                             cudaStream_t st);
cudaError_t launch_split_h0(const float4* h0, float* h0k, float* h0minusk, int n, cudaStream_t st);
cudaError_t launch_merge_h0(float4* h0, const float* h0k, const float* h0minusk, int n, cudaStream_t st);

}  // namespace ow
