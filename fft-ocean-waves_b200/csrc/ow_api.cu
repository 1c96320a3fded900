// ow_api.cu — the C ABI declared in include/oceanwaves.h: context, buffers, launch sequencing.
// Host side of the drop-in: what FFTOceanWaves::init()/update() do for the sim (reference
// src/main.cpp:199-255, 553-744, 1083-1145) minus windowing, rendering and GL plumbing.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <vector>

#include "../../include/oceanwaves.h"
#include "ow_internal.h"

using namespace ow;

struct ow_ctx {
    int N = 0, n_cascades = 0, n_slots = 0, device = 0;
    uint32_t flags = 0;
    std::vector<ow_params> params;
    std::vector<CascadeDev> casc_host;
    // device buffers
    uint8_t* d_noise = nullptr;   // [cascade][4][nh][nw]
    int noise_w = 0, noise_h = 0;
    std::vector<char> noise_set;
    float4* d_h0 = nullptr;
    float4* d_hp = nullptr;       // [cascade][N/2][N] folded texel pairs (what the row kernel streams)
    float4* d_nyq = nullptr;      // [cascade][N/2]
    float* d_ktab = nullptr;
    float* d_ktab_sub = nullptr;  // contexts that run the N = A*B line decomposition: the k table sub-line-major (their folded rows d_hp are too)
    int sub_A = 0;                // A of that decomposition, 0 for the direct kernels
    CascadeDev* d_casc = nullptr;
    float2* d_inter = nullptr;
    float* d_disp = nullptr;
    float4* d_normal = nullptr;
    float* d_jac = nullptr;
    float* d_tmp = nullptr;       // 2 * N*N*2 floats staging for h0 split/merge
    float2* d_scratch = nullptr;  // N = A*B line decomposition only (N > 4096 or OW_FLAG_FOUR_STEP): one frame of radix-A sums
    cudaStream_t stream = nullptr;
    // Launch groups of one ow_step_multi call are independent frames: they are spread round-robin over these
    // auxiliary streams (fork/join around the caller's stream) so one group's tail overlaps the next group's head.
    static constexpr int kMaxAux = 4;
    cudaStream_t aux[kMaxAux] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[kMaxAux] = {nullptr, nullptr, nullptr, nullptr};
    int n_streams = 3;
    std::vector<char> h0_ready;   // per cascade: h0/hp/nyq/ktab hold the spectrum of the CURRENT parameters
    bool casc_uploaded = false;   // d_casc matches params
    int group_size = 0;
    int last_launches = 0, last_groups = 0;
    std::string err;
    KernelConfig kcfg;            // per-context (device) launch facts: SM count, persistent-kernel occupancy
    int row_mode = 0;             // ow_set_row_kernel
    int col_mode = 0, fuse_mode = -1;   // ow_set_column_kernel
    int* d_seam = nullptr;        // [n_slots][N/16] seam counters of the fused normal map
    alignas(64) unsigned char inter_tmap[128];   // CUtensorMap over d_inter (64 bytes used) for the TMA-staged column kernel
    bool have_tmap = false;
    int discard_inter = 0;        // ow_set_discard_intermediate
    // L2 residency of the folded spectrum: a time sweep re-reads the same cascade's block every frame, so when all blocks together fit the
    // device's persisting-L2 carve-out the launch streams carry an access-policy window over d_hp (ow_set_l2_persist). OFF by default: measured on B200 at N = 2048 the carve-out costs the streaming data more than the spectrum's hits save (profiles/r02_l2_persist_ab.txt)
    int l2_persist = 0;
    bool l2_window_on = false;
    cudaStream_t l2_user_stream = nullptr;   // last caller stream the window was applied to
    int frame_mode = 0;           // ow_set_frame_kernel
    int* d_mega_sched = nullptr;  // (kMaxAux + 1) areas of kMegaSchedInts counters: one per auxiliary stream + one for the caller's stream
    int latency_shapes = 1;       // ow_set_latency_shapes
    int line_clusters = 0;        // ow_set_line_clusters: 0 = global scratch (default: the DSMEM exchange measured 3.5x slower on B200), -1 = whatever
                                  // cluster shapes the device can co-schedule (kcfg.big_cluster), else a bit mask
    int cap_row = 0, cap_col = 0; // ow_set_resident_ctas: CTAs per SM of the persistent row / column kernels (0 = what fits)
    // ow_step (slot i <- cascade i at ONE time t) as a CUDA graph: [exact sincos, fast sincos]; rebuilt when a tuning knob changes
    bool graph_enabled = true;
    std::unique_ptr<GraphPlan> plan[2];
    std::vector<int32_t> ident;   // 0..n_cascades-1 (ow_step's slot map; no per-call allocation)
    std::vector<float> tbuf;
    // packed outputs (OW_FLAG_PACKED_F32 / _F16)
    char* d_packed = nullptr;
    PackedBuffers pk{};
    // GL interop
    cudaGraphicsResource* gl_res[4] = {nullptr, nullptr, nullptr, nullptr};
    int gl_count = 0;             // 4 = dy,dx,dz,normal; 2 = packed displacement + normal_xz
    bool gl_registered = false;
    // multi-cascade composition / the demo's clock
    struct SlotPatch { float L = 0.0f, choppiness = 0.0f; };     // patch size / choppiness the frame in a slot was computed with (L = 0: never stepped);
    std::vector<SlotPatch> slot_patch;                           // NOT the cascade index: ow_set_params may change the cascade before its next step
    float time_scale = 1.0f, time_offset = 0.0f;
    float* d_query = nullptr;         // staging of ow_sample_points_host: [cap][2] positions + [cap][8] results
    size_t query_cap = 0;
};

static thread_local std::string g_create_error;

namespace {

int fail(ow_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}

int cuda_fail(ow_ctx* c, cudaError_t e, const char* what) {
    return fail(c, OW_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define OW_CUDA(ctx, call)                                             \
    do {                                                               \
        cudaError_t e__ = (call);                                      \
        if (e__ != cudaSuccess) return cuda_fail((ctx), e__, #call);   \
    } while (0)

CascadeDev to_dev(const ow_params& p) {
    CascadeDev c{};
    c.L = p.L; c.wind_speed = p.wind_speed; c.amplitude = p.amplitude; c.suppression = p.suppression;
    c.choppiness = p.choppiness;
    // glm::normalize(vec2) = v * inversesqrt(dot(v,v))  (reference src/main.cpp:555)
    const float inv = 1.0f / sqrtf(p.wind_dir[0] * p.wind_dir[0] + p.wind_dir[1] * p.wind_dir[1]);
    c.wdx = p.wind_dir[0] * inv; c.wdy = p.wind_dir[1] * inv;
    return c;
}

bool valid_params(const ow_params& p) {
    return p.L > 0.0f && (p.wind_dir[0] != 0.0f || p.wind_dir[1] != 0.0f) && p.wind_speed > 0.0f;
}

cudaStream_t pick(ow_ctx* c, void* s) { return s ? static_cast<cudaStream_t>(s) : c->stream; }

FrameBuffers buffers(const ow_ctx* c) {
    FrameBuffers fb{};
    fb.N = c->N; fb.h0 = c->d_h0; fb.hp = c->d_hp; fb.nyq = c->d_nyq; fb.ktab = c->d_ktab; fb.ktab_sub = c->d_ktab_sub; fb.casc = c->d_casc; fb.inter = c->d_inter;
    fb.disp = c->d_disp; fb.normal = c->d_normal; fb.jacobian = c->d_jac;
    // Measured on B200 (profiles/r01d_discard_ab.txt): dropping the consumed intermediate from L2 changed nothing in the
    // multi-stream sweep, so it is off unless ow_set_discard_intermediate turns it on.
    fb.discard_inter = c->discard_inter;
    fb.four_step = (c->flags & OW_FLAG_FOUR_STEP) ? 1 : 0;
    fb.scratch = c->d_scratch;
    fb.row_mode = c->row_mode;
    fb.col_mode = c->col_mode;
    fb.fuse_mode = c->fuse_mode;
    if (c->flags & OW_FLAG_FUSED_NORMALS) {          // force the fused epilogue whatever the per-N default is
        fb.fuse_mode = 1;
        if (fb.col_mode <= 1) fb.col_mode = c->have_tmap ? 3 : 2;
    }
    fb.seam = c->d_seam;
    fb.inter_tmap = c->have_tmap ? c->inter_tmap : nullptr;
    fb.row_bulk_ctas[0] = c->kcfg.row_bulk_ctas[0]; fb.row_bulk_ctas[1] = c->kcfg.row_bulk_ctas[1];
    fb.col2_ctas[0] = c->kcfg.col2_ctas[0]; fb.col2_ctas[1] = c->kcfg.col2_ctas[1];
    fb.sm_count = c->kcfg.sm_count;
    fb.row_pipe_ctas[0] = c->kcfg.row_pipe_ctas[0]; fb.row_pipe_ctas[1] = c->kcfg.row_pipe_ctas[1];
    for (int i = 0; i < 2; ++i) {
        if (c->cap_row > 0) { fb.row_pipe_ctas[i] = std::min(fb.row_pipe_ctas[i], c->cap_row); fb.row_bulk_ctas[i] = std::min(fb.row_bulk_ctas[i], c->cap_row); }
        if (c->cap_col > 0) fb.col2_ctas[i] = std::min(fb.col2_ctas[i], c->cap_col);
    }
    fb.latency_shapes = c->latency_shapes;
    fb.frame_mode = c->frame_mode;
    fb.mega_ctas = c->kcfg.mega_ctas;
    fb.mega_sched = c->d_mega_sched ? c->d_mega_sched + (size_t)ow_ctx::kMaxAux * kMegaSchedInts : nullptr;   // the caller's stream; step_impl re-points it per group
    fb.big_cluster = c->line_clusters < 0 ? c->kcfg.big_cluster
                                          : (c->line_clusters & c->kcfg.big_cluster & 3) | ((c->line_clusters & 2) ? (c->line_clusters & 4) : 0);
    fb.col_pipe_ctas = c->cap_col > 0 ? std::min(c->kcfg.col_pipe_ctas, c->cap_col) : c->kcfg.col_pipe_ctas;
    // line decomposition: the persistent pipelined column lines kernel unless ow_set_column_kernel(1) asks for one CTA per item
    fb.bigcol_pipe_grid = c->col_mode == 4 ? c->kcfg.sm_count * c->kcfg.bigcol_pipe_ctas : c->col_mode == 2 ? -1 : 0;
    return fb;
}

void drop_plans(ow_ctx* c) { c->plan[0].reset(); c->plan[1].reset(); }

size_t hp_bytes(const ow_ctx* c) { return hp_block_elems(c->N / 2, c->N) * c->n_cascades * sizeof(float4); }

// Access-policy window over the folded spectrum on one stream (hits persist in the L2 carve-out, everything else streams).
void apply_l2_window(ow_ctx* c, cudaStream_t st, bool on) {
    cudaStreamAttrValue v{};
    v.accessPolicyWindow.base_ptr = on ? static_cast<void*>(c->d_hp) : nullptr;
    v.accessPolicyWindow.num_bytes = on ? hp_bytes(c) : 0;
    v.accessPolicyWindow.hitRatio = on ? 1.0f : 0.0f;
    v.accessPolicyWindow.hitProp = on ? cudaAccessPropertyPersisting : cudaAccessPropertyNormal;
    v.accessPolicyWindow.missProp = on ? cudaAccessPropertyStreaming : cudaAccessPropertyNormal;
    if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) cudaGetLastError();   // a hint: never fatal
}

// (Re)decides whether the window is used and applies it to the context's own streams.
void configure_l2(ow_ctx* c) {
    int max_persist = 0, max_window = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, c->device);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, c->device);
    const size_t bytes = hp_bytes(c);
    // automatic: only when every cascade's block fits at once and leaves at least a third of the carve-out's budget to the streaming data
    bool on = c->l2_persist > 0 || (c->l2_persist < 0 && bytes <= (size_t)max_persist * 2 / 3);
    if (bytes > (size_t)max_persist || bytes > (size_t)max_window) on = false;
    if (on && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, bytes) != cudaSuccess) { cudaGetLastError(); on = false; }
    if (!on && c->l2_window_on) { cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0); cudaGetLastError(); }    // give the carve-out back
    c->l2_window_on = on;
    c->l2_user_stream = nullptr;
    apply_l2_window(c, c->stream, on);
    for (auto& s : c->aux) apply_l2_window(c, s, on);
}

bool all_ready(const ow_ctx* c) {
    for (char r : c->h0_ready) if (!r) return false;
    return true;
}

// Destroys the timing events of ow_step_multi_timed on every exit path.
struct EventGuard {
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    ~EventGuard() { for (auto& e : ev) if (e) cudaEventDestroy(e); }
};

// Slots per launch group. Upper bound: a group's 12 B/texel intermediate should not exceed ~100 MB (measured on
// B200: launch count and tail efficiency matter more than keeping the intermediate strictly inside the 126 MB L2).
// Then the slots are split evenly (32 -> 11+11+10, not 15+15+2) into at least n_streams groups, so that the
// independent groups can overlap on the auxiliary streams.
int pick_group(const ow_ctx* c, int count) {
    if (c->d_scratch) return 1;       // the line decomposition's scratch holds one frame
    int gmax;
    if (c->group_size > 0) {
        gmax = c->group_size;
    } else {
        // small grids: a launch has to be long enough to amortise its ramp-up and tail (measured at N=512: 11 frames per
        // group 301 k fps, 21 per group 326 k fps), so the budget is larger there
        // ... and at N = 2048 three frames per launch beat one (17.8 k vs 17.2 k frames/s: the row kernel's 2.3 waves of CTAs per frame)
        gmax = (int)((c->N <= 512 ? 400.0e6 : c->N >= 2048 ? 160.0e6 : 100.0e6) / (12.0 * c->N * (double)c->N));
        if (gmax < 1) gmax = 1;
    }
    if (gmax > kMaxGroup) gmax = kMaxGroup;
    int ngroups = (count + gmax - 1) / gmax;
    if (c->group_size <= 0) {
        const int want = c->n_streams < count ? c->n_streams : count;
        if (ngroups < want) ngroups = want;
    }
    return (count + ngroups - 1) / ngroups;
}

void release(ow_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->gl_registered) for (auto& r : c->gl_res) if (r) cudaGraphicsUnregisterResource(r);
    c->gl_registered = false;
    cudaFree(c->d_noise); cudaFree(c->d_h0); cudaFree(c->d_hp); cudaFree(c->d_nyq); cudaFree(c->d_ktab); cudaFree(c->d_ktab_sub); cudaFree(c->d_casc); cudaFree(c->d_inter);
    drop_plans(c);
    cudaFree(c->d_mega_sched);
    cudaFree(c->d_query);
    cudaFree(c->d_disp); cudaFree(c->d_normal); cudaFree(c->d_jac); cudaFree(c->d_tmp); cudaFree(c->d_scratch); cudaFree(c->d_packed); cudaFree(c->d_seam);
    if (c->stream) cudaStreamDestroy(c->stream);
    for (auto& s : c->aux) if (s) cudaStreamDestroy(s);
    for (auto& e : c->ev_join) if (e) cudaEventDestroy(e);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    delete c;
}

}  // namespace

extern "C" {

int ow_create(int32_t N, int32_t n_cascades, int32_t n_slots, const ow_params* cascades, int32_t device,
              uint32_t flags, ow_ctx** out) {
    if (!out) return fail(nullptr, OW_ERR_INVALID, "ow_create: out is NULL");
    *out = nullptr;
    if (!frame_supported(N)) return fail(nullptr, OW_ERR_INVALID, "ow_create: N must be a power of two in [128, 32768]");
    if ((flags & OW_FLAG_FOUR_STEP) && !big_supported(N, true))
        return fail(nullptr, OW_ERR_INVALID, "ow_create: OW_FLAG_FOUR_STEP is a test mode for N = 1024 or 2048");
    if ((flags & OW_FLAG_PACKED_F32) && (flags & OW_FLAG_PACKED_F16))
        return fail(nullptr, OW_ERR_INVALID, "ow_create: OW_FLAG_PACKED_F32 and OW_FLAG_PACKED_F16 are mutually exclusive");
    if (n_cascades < 1 || !cascades) return fail(nullptr, OW_ERR_INVALID, "ow_create: need >= 1 cascade");
    if (n_slots < n_cascades) return fail(nullptr, OW_ERR_INVALID, "ow_create: n_slots must be >= n_cascades");
    for (int i = 0; i < n_cascades; ++i)
        if (!valid_params(cascades[i])) return fail(nullptr, OW_ERR_INVALID, "ow_create: invalid cascade parameters");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceCount (no usable CUDA device; there is no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(nullptr, OW_ERR_INVALID, "ow_create: device ordinal out of range");
    ow_ctx* c = new (std::nothrow) ow_ctx();
    if (!c) return fail(nullptr, OW_ERR_NOMEM, "ow_create: out of host memory");
    c->N = N; c->n_cascades = n_cascades; c->n_slots = n_slots; c->device = device; c->flags = flags;
    c->params.assign(cascades, cascades + n_cascades);
    c->noise_set.assign(n_cascades, 0);
    c->h0_ready.assign(n_cascades, 0);
    c->ident.resize(n_cascades);
    for (int i = 0; i < n_cascades; ++i) c->ident[i] = i;
    c->tbuf.assign(n_cascades, 0.0f);
    c->casc_host.resize(n_cascades);
    c->slot_patch.assign(n_slots, ow_ctx::SlotPatch{});
    const size_t nn = (size_t)N * N;
#define OW_TRY(call)                                                                 \
    do {                                                                             \
        cudaError_t e__ = (call);                                                    \
        if (e__ != cudaSuccess) { int r__ = cuda_fail(nullptr, e__, #call); release(c); return r__; } \
    } while (0)
    OW_TRY(cudaSetDevice(device));
    OW_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (int i = 0; i < ow_ctx::kMaxAux; ++i) {
        // (graded priorities on these streams, so that the chains of one call run staggered instead of in phase, were measured and are
        // slower: C2 336 k vs 362 k frames/s, C4 75 k vs 80 k - profiles/r02_experiments.md)
        OW_TRY(cudaStreamCreateWithFlags(&c->aux[i], cudaStreamNonBlocking));
        OW_TRY(cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming));
    }
    OW_TRY(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    OW_TRY(cudaMalloc(&c->d_h0, nn * n_cascades * sizeof(float4)));
    OW_TRY(cudaMalloc(&c->d_hp, hp_block_elems(N / 2, N) * n_cascades * sizeof(float4)));   // per cascade: [N/2][N] float4 (+ [N/2][N] float2, N <= 512)
    OW_TRY(cudaMalloc(&c->d_nyq, (size_t)(N / 2) * n_cascades * sizeof(float4)));
    OW_TRY(cudaMemsetAsync(c->d_hp, 0, hp_block_elems(N / 2, N) * n_cascades * sizeof(float4), c->stream));
    OW_TRY(cudaMemsetAsync(c->d_nyq, 0, (size_t)(N / 2) * n_cascades * sizeof(float4), c->stream));
    OW_TRY(cudaMalloc(&c->d_ktab, (size_t)N * n_cascades * sizeof(float)));
    OW_TRY(cudaMalloc(&c->d_casc, n_cascades * sizeof(CascadeDev)));
    OW_TRY(cudaMalloc(&c->d_inter, nn / 2 * 3 * n_slots * sizeof(float2)));
    OW_TRY(cudaMalloc(&c->d_disp, nn * 3 * n_slots * sizeof(float)));
    OW_TRY(cudaMalloc(&c->d_normal, nn * n_slots * sizeof(float4)));
    if (flags & OW_FLAG_JACOBIAN) OW_TRY(cudaMalloc(&c->d_jac, nn * n_slots * sizeof(float)));
    OW_TRY(cudaMalloc(&c->d_tmp, nn * 4 * sizeof(float)));
    if (big_supported(N, false) || (flags & OW_FLAG_FOUR_STEP)) {
        OW_TRY(cudaMalloc(&c->d_scratch, nn / 2 * 3 * sizeof(float2)));
        c->sub_A = big_radix(N, (flags & OW_FLAG_FOUR_STEP) != 0);
        OW_TRY(cudaMalloc(&c->d_ktab_sub, (size_t)N * n_cascades * sizeof(float)));
    }
    if (flags & (OW_FLAG_PACKED_F32 | OW_FLAG_PACKED_F16)) {
        c->pk.half = (flags & OW_FLAG_PACKED_F16) ? 1 : 0;
        c->pk.normal_offset = nn * (c->pk.half ? 8 : 16);
        c->pk.slot_bytes = c->pk.normal_offset + nn * 4;
        OW_TRY(cudaMalloc(&c->d_packed, c->pk.slot_bytes * n_slots));
        c->pk.base = c->d_packed;
    }
    OW_TRY(cudaMalloc(&c->d_seam, (size_t)n_slots * (N / 16) * sizeof(int)));
    OW_TRY(cudaMemsetAsync(c->d_seam, 0, (size_t)n_slots * (N / 16) * sizeof(int), c->stream));
    OW_TRY(cudaMalloc(&c->d_mega_sched, (size_t)(ow_ctx::kMaxAux + 1) * kMegaSchedInts * sizeof(int)));
    OW_TRY(configure_frame_kernels(N, &c->kcfg));
    configure_l2(c);
    c->have_tmap = make_inter_tensor_map(c->inter_tmap, c->d_inter, N, n_slots);
#undef OW_TRY
    *out = c;
    return OW_OK;
}

void ow_destroy(ow_ctx* ctx) { release(ctx); }

const char* ow_last_error(const ow_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int ow_set_params(ow_ctx* c, int32_t cascade, const ow_params* p) {
    if (!c || !p) return OW_ERR_INVALID;
    if (cascade < 0 || cascade >= c->n_cascades) return fail(c, OW_ERR_INVALID, "ow_set_params: cascade out of range");
    if (!valid_params(*p)) return fail(c, OW_ERR_INVALID, "ow_set_params: invalid parameters");
    c->params[cascade] = *p;
    c->h0_ready[cascade] = 0;          // only this cascade needs a new ow_init_spectrum / ow_set_h0
    c->casc_uploaded = false;
    return OW_OK;
}

int ow_set_noise(ow_ctx* c, int32_t cascade, const uint8_t* const planes[4], int32_t w, int32_t h) {
    if (!c || !planes) return OW_ERR_INVALID;
    if (w < 1 || h < 1) return fail(c, OW_ERR_INVALID, "ow_set_noise: empty noise image");
    if (cascade < -1 || cascade >= c->n_cascades) return fail(c, OW_ERR_INVALID, "ow_set_noise: cascade out of range");
    for (int j = 0; j < 4; ++j) if (!planes[j]) return fail(c, OW_ERR_INVALID, "ow_set_noise: NULL plane");
    OW_CUDA(c, cudaSetDevice(c->device));
    const size_t plane = (size_t)w * h;
    if (c->d_noise && (w != c->noise_w || h != c->noise_h)) {
        OW_CUDA(c, cudaStreamSynchronize(c->stream));
        cudaFree(c->d_noise); c->d_noise = nullptr;
        std::fill(c->noise_set.begin(), c->noise_set.end(), 0);
        std::fill(c->h0_ready.begin(), c->h0_ready.end(), 0);
    }
    if (!c->d_noise) {
        OW_CUDA(c, cudaMalloc(&c->d_noise, plane * 4 * c->n_cascades));
        c->noise_w = w; c->noise_h = h;
    }
    const int lo = cascade < 0 ? 0 : cascade, hi = cascade < 0 ? c->n_cascades : cascade + 1;
    for (int i = lo; i < hi; ++i) {
        for (int j = 0; j < 4; ++j)
            OW_CUDA(c, cudaMemcpyAsync(c->d_noise + ((size_t)i * 4 + j) * plane, planes[j], plane, cudaMemcpyHostToDevice, c->stream));
        c->noise_set[i] = 1;
        c->h0_ready[i] = 0;
    }
    OW_CUDA(c, cudaStreamSynchronize(c->stream));   // host planes may be freed by the caller on return
    return OW_OK;
}

int ow_set_noise_seed(ow_ctx* c, int32_t cascade, uint64_t seed) {
    if (!c) return OW_ERR_INVALID;
    if (cascade < -1 || cascade >= c->n_cascades) return fail(c, OW_ERR_INVALID, "ow_set_noise_seed: cascade out of range");
    OW_CUDA(c, cudaSetDevice(c->device));
    const int N = c->N;
    const size_t plane = (size_t)N * N;
    if (c->d_noise && (c->noise_w != N || c->noise_h != N)) {
        OW_CUDA(c, cudaStreamSynchronize(c->stream));
        cudaFree(c->d_noise); c->d_noise = nullptr;
        std::fill(c->noise_set.begin(), c->noise_set.end(), 0);
        std::fill(c->h0_ready.begin(), c->h0_ready.end(), 0);
    }
    if (!c->d_noise) {
        OW_CUDA(c, cudaMalloc(&c->d_noise, plane * 4 * c->n_cascades));
        c->noise_w = N; c->noise_h = N;
    }
    const int lo = cascade < 0 ? 0 : cascade, hi = cascade < 0 ? c->n_cascades : cascade + 1;
    for (int i = lo; i < hi; ++i) {
        OW_CUDA(c, launch_noise_seed(c->d_noise + (size_t)i * 4 * plane, N, seed, c->stream));
        c->noise_set[i] = 1;
        c->h0_ready[i] = 0;
    }
    OW_CUDA(c, cudaStreamSynchronize(c->stream));
    return OW_OK;
}

// tilde_h0_k_cs.glsl for the cascades in [lo, hi) + the constants every cascade's kernels read.
static int init_range(ow_ctx* c, int lo, int hi, const char* who) {
    for (int i = lo; i < hi; ++i)
        if (!c->noise_set[i]) return fail(c, OW_ERR_STATE, std::string(who) + ": ow_set_noise has not been called for every cascade involved");
    OW_CUDA(c, cudaSetDevice(c->device));
    const size_t nn = (size_t)c->N * c->N, plane = (size_t)c->noise_w * c->noise_h;
    for (int i = lo; i < hi; ++i) {
        c->casc_host[i] = to_dev(c->params[i]);
        OW_CUDA(c, launch_ktab(c->d_ktab + (size_t)i * c->N, c->N, c->params[i].L, c->stream));
        OW_CUDA(c, launch_h0(c->d_h0 + (size_t)i * nn, c->d_noise + (size_t)i * 4 * plane, c->noise_w, c->noise_h, c->N,
                             c->casc_host[i], c->stream));
        if (c->sub_A) OW_CUDA(c, launch_ktab_sub(c->d_ktab + (size_t)i * c->N, c->d_ktab_sub + (size_t)i * c->N, c->N, c->sub_A, c->stream));
        OW_CUDA(c, launch_fold(c->d_h0 + (size_t)i * nn, c->d_hp + (size_t)i * hp_block_elems(c->N / 2, c->N), c->d_nyq + (size_t)i * (c->N / 2), c->d_ktab + (size_t)i * c->N, c->N,
                               c->sub_A, c->stream));
    }
    for (int i = 0; i < c->n_cascades; ++i) c->casc_host[i] = to_dev(c->params[i]);
    OW_CUDA(c, cudaMemcpyAsync(c->d_casc, c->casc_host.data(), c->n_cascades * sizeof(CascadeDev), cudaMemcpyHostToDevice, c->stream));
    OW_CUDA(c, cudaStreamSynchronize(c->stream));   // the reference ends tilde_h0_k() with glFinish (main.cpp:582)
    c->casc_uploaded = true;
    for (int i = lo; i < hi; ++i) c->h0_ready[i] = 1;
    return OW_OK;
}

int ow_init_spectrum(ow_ctx* c) {
    if (!c) return OW_ERR_INVALID;
    return init_range(c, 0, c->n_cascades, "ow_init_spectrum");
}

int ow_init_spectrum_cascade(ow_ctx* c, int32_t cascade) {
    if (!c) return OW_ERR_INVALID;
    if (cascade < 0 || cascade >= c->n_cascades) return fail(c, OW_ERR_INVALID, "ow_init_spectrum_cascade: cascade out of range");
    return init_range(c, cascade, cascade + 1, "ow_init_spectrum_cascade");
}

int ow_set_h0(ow_ctx* c, int32_t cascade, const float* h0k, const float* h0minusk) {
    if (!c || !h0k || !h0minusk) return OW_ERR_INVALID;
    if (cascade < 0 || cascade >= c->n_cascades) return fail(c, OW_ERR_INVALID, "ow_set_h0: cascade out of range");
    OW_CUDA(c, cudaSetDevice(c->device));
    const size_t nn = (size_t)c->N * c->N;
    OW_CUDA(c, cudaMemcpyAsync(c->d_tmp, h0k, nn * 2 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    OW_CUDA(c, cudaMemcpyAsync(c->d_tmp + nn * 2, h0minusk, nn * 2 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    OW_CUDA(c, launch_merge_h0(c->d_h0 + (size_t)cascade * nn, c->d_tmp, c->d_tmp + nn * 2, (int)nn, c->stream));
    // this cascade's k table and the cascade constants follow the CURRENT parameters (h0 itself is the caller's)
    OW_CUDA(c, launch_ktab(c->d_ktab + (size_t)cascade * c->N, c->N, c->params[cascade].L, c->stream));
    if (c->sub_A) OW_CUDA(c, launch_ktab_sub(c->d_ktab + (size_t)cascade * c->N, c->d_ktab_sub + (size_t)cascade * c->N, c->N, c->sub_A, c->stream));
    OW_CUDA(c, launch_fold(c->d_h0 + (size_t)cascade * nn, c->d_hp + (size_t)cascade * hp_block_elems(c->N / 2, c->N), c->d_nyq + (size_t)cascade * (c->N / 2),
                           c->d_ktab + (size_t)cascade * c->N, c->N, c->sub_A, c->stream));
    for (int i = 0; i < c->n_cascades; ++i) c->casc_host[i] = to_dev(c->params[i]);
    OW_CUDA(c, cudaMemcpyAsync(c->d_casc, c->casc_host.data(), c->n_cascades * sizeof(CascadeDev), cudaMemcpyHostToDevice, c->stream));
    OW_CUDA(c, cudaStreamSynchronize(c->stream));
    c->casc_uploaded = true;
    c->h0_ready[cascade] = 1;         // ONLY this cascade: the others keep their own state
    return OW_OK;
}

// Fast sincos is allowed when every |w*t| of the entries stays below kFastPhaseLimit: w_max = sqrt(g*|k|max), |k|max = sqrt(2)*pi*N/L.
static bool fast_phase_ok(const ow_ctx* c, int cascade, float t) {
    const float kmax = 1.41421356f * 3.14159265f * (float)c->N / c->params[cascade].L;
    return sqrtf(9.81f * kmax) * fabsf(t) < kFastPhaseLimit;
}

static int launch_failed(ow_ctx* c, const char* what) {
    const cudaError_t e = take_launch_error();
    if (e != cudaSuccess) return cuda_fail(c, e, what);
    return fail(c, OW_ERR_STATE, std::string(what) + ": this launch shape is not supported by the context (e.g. several slots per group with the N = A*B line decomposition)");
}

static int step_impl(ow_ctx* c, int32_t count, const int32_t* cascade_of_slot, const float* time_of_slot, void* stream,
                     float* kernel_ms) {
    if (!c || !cascade_of_slot || !time_of_slot) return OW_ERR_INVALID;
    if (count < 1 || count > c->n_slots) return fail(c, OW_ERR_INVALID, "ow_step_multi: count out of range");
    for (int i = 0; i < count; ++i) {
        if (cascade_of_slot[i] < 0 || cascade_of_slot[i] >= c->n_cascades)
            return fail(c, OW_ERR_INVALID, "ow_step_multi: cascade index out of range");
        if (!c->h0_ready[cascade_of_slot[i]])
            return fail(c, OW_ERR_STATE, "ow_step: a referenced cascade has no current spectrum: call ow_init_spectrum (or ow_set_h0) after "
                                         "ow_create / ow_set_params / ow_set_noise");
    }
    OW_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = pick(c, stream);
    if (c->l2_window_on && st != c->stream && st != c->l2_user_stream) {      // a caller-owned stream: give it the window once
        apply_l2_window(c, st, true);
        c->l2_user_stream = st;
    }
    const FrameBuffers fb = buffers(c);
    const int group = pick_group(c, count);
    int launches = 0;
    EventGuard guard;
    if (kernel_ms) {
        kernel_ms[0] = kernel_ms[1] = kernel_ms[2] = 0.0f;
        for (auto& e : guard.ev) OW_CUDA(c, cudaEventCreate(&e));
    }
    const int ngroups = (count + group - 1) / group;
    const int nfan = (kernel_ms || ngroups < 2 || c->n_streams < 2 || c->d_scratch) ? 0 : (ngroups < c->n_streams ? ngroups : c->n_streams);
    if (nfan) {
        OW_CUDA(c, cudaEventRecord(c->ev_fork, st));
        for (int i = 0; i < nfan; ++i) OW_CUDA(c, cudaStreamWaitEvent(c->aux[i], c->ev_fork, 0));
    }
    int gi = 0;
    for (int base = 0; base < count; base += group, ++gi) {
        const int n = count - base < group ? count - base : group;
        Launcher L(nfan ? c->aux[gi % nfan] : st);
        SlotTable tab{};
        bool fast = (c->flags & OW_FLAG_EXACT_SINCOS) == 0;
        for (int i = 0; i < n; ++i) {
            tab.cascade[i] = cascade_of_slot[base + i];
            tab.time[i] = time_of_slot[base + i];
            tab.slot[i] = base + i;
            if (!fast_phase_ok(c, tab.cascade[i], tab.time[i])) fast = false;
        }
        FrameBuffers fbg = fb;
        if (fbg.mega_sched && nfan) fbg.mega_sched = c->d_mega_sched + (size_t)(gi % nfan) * kMegaSchedInts;     // this group's stream's counters
        int k = launch_frame(fbg, tab, n, (c->flags & OW_FLAG_JACOBIAN) != 0, fast, L, kernel_ms ? guard.ev : nullptr);
        if (k < 0) return launch_failed(c, "launch_frame");
        launches += k;
        if (c->d_packed) {
            k = launch_pack(fb, tab, n, c->pk, L);
            if (k < 0) return launch_failed(c, "launch_pack");
            launches += k;
        }
        if (kernel_ms) {
            OW_CUDA(c, cudaEventSynchronize(guard.ev[3]));
            for (int i = 0; i < 3; ++i) {
                float ms = 0.0f;
                OW_CUDA(c, cudaEventElapsedTime(&ms, guard.ev[i], guard.ev[i + 1]));
                kernel_ms[i] += ms;
            }
        }
    }
    for (int i = 0; i < nfan; ++i) {
        OW_CUDA(c, cudaEventRecord(c->ev_join[i], c->aux[i]));
        OW_CUDA(c, cudaStreamWaitEvent(st, c->ev_join[i], 0));
    }
    c->last_launches = launches;
    c->last_groups = ngroups;
    for (int i = 0; i < count; ++i) c->slot_patch[i] = {c->casc_host[cascade_of_slot[i]].L, c->casc_host[cascade_of_slot[i]].choppiness};
    return OW_OK;
}

int ow_step_multi(ow_ctx* c, int32_t count, const int32_t* cascade_of_slot, const float* time_of_slot, void* stream) {
    return step_impl(c, count, cascade_of_slot, time_of_slot, stream, nullptr);
}

int ow_step_multi_timed(ow_ctx* c, int32_t count, const int32_t* cascade_of_slot, const float* time_of_slot, void* stream,
                        float* kernel_ms) {
    if (!kernel_ms) return OW_ERR_INVALID;
    return step_impl(c, count, cascade_of_slot, time_of_slot, stream, kernel_ms);
}

// The frame of ow_step as a CUDA graph: one chain of kernel nodes per launch group (groups stay independent, as on the
// auxiliary streams). Built once per (context, sincos variant); per frame only the time inside the row kernels' slot
// tables changes, patched with cudaGraphExecKernelNodeSetParams. This is the reference's update() (src/main.cpp:240-244)
// - 53 dispatches there - as ONE submission.
static int build_plan(ow_ctx* c, bool fast, std::unique_ptr<GraphPlan>* out) {
    auto plan = std::make_unique<GraphPlan>();
    OW_CUDA(c, cudaGraphCreate(&plan->graph, 0));
    const FrameBuffers fb = buffers(c);
    const int count = c->n_cascades, group = pick_group(c, count);
    for (int base = 0; base < count; base += group) {
        const int n = count - base < group ? count - base : group;
        SlotTable tab{};
        for (int i = 0; i < n; ++i) { tab.cascade[i] = base + i; tab.time[i] = 0.0f; tab.slot[i] = base + i; }
        Launcher L(plan.get());
        int k = launch_frame(fb, tab, n, (c->flags & OW_FLAG_JACOBIAN) != 0, fast, L, nullptr);
        if (k < 0) return launch_failed(c, "launch_frame (graph)");
        plan->launches += k;
        if (c->d_packed) {
            k = launch_pack(fb, tab, n, c->pk, L);
            if (k < 0) return launch_failed(c, "launch_pack (graph)");
            plan->launches += k;
        }
        ++plan->groups;
    }
    OW_CUDA(c, cudaGraphInstantiate(&plan->exec, plan->graph, 0));
    *out = std::move(plan);
    return OW_OK;
}

int ow_step(ow_ctx* c, float t, void* stream) {
    if (!c) return OW_ERR_INVALID;
    const FrameBuffers fb = buffers(c);
    if (!c->graph_enabled || !frame_graphable(fb)) {
        std::fill(c->tbuf.begin(), c->tbuf.end(), t);
        return step_impl(c, c->n_cascades, c->ident.data(), c->tbuf.data(), stream, nullptr);
    }
    if (!all_ready(c))
        return fail(c, OW_ERR_STATE, "ow_step: a cascade has no current spectrum: call ow_init_spectrum (or ow_set_h0) after ow_create / "
                                     "ow_set_params / ow_set_noise");
    OW_CUDA(c, cudaSetDevice(c->device));
    bool fast = (c->flags & OW_FLAG_EXACT_SINCOS) == 0;
    for (int i = 0; i < c->n_cascades && fast; ++i) fast = fast_phase_ok(c, i, t);
    std::unique_ptr<GraphPlan>& plan = c->plan[fast ? 1 : 0];
    if (!plan) {
        const int rc = build_plan(c, fast, &plan);
        if (rc != OW_OK) return rc;
    }
    for (GraphNodeRec& n : plan->nodes) {
        if (!n.tab) continue;
        for (int i = 0; i < kMaxGroup; ++i) n.tab->time[i] = t;
        OW_CUDA(c, cudaGraphExecKernelNodeSetParams(plan->exec, n.node, &n.params));
    }
    OW_CUDA(c, cudaGraphLaunch(plan->exec, pick(c, stream)));
    c->last_launches = plan->launches;
    c->last_groups = plan->groups;
    for (int i = 0; i < c->n_cascades; ++i) c->slot_patch[i] = {c->casc_host[i].L, c->casc_host[i].choppiness};
    return OW_OK;
}

int ow_set_graph(ow_ctx* c, int32_t enabled) {
    if (!c) return OW_ERR_INVALID;
    c->graph_enabled = enabled != 0;
    if (!c->graph_enabled) drop_plans(c);
    return OW_OK;
}

int ow_set_row_kernel(ow_ctx* c, int32_t mode) {
    if (!c) return OW_ERR_INVALID;
    if (mode < 0 || mode > 3)
        return fail(c, OW_ERR_INVALID, "ow_set_row_kernel: mode must be 0 (per-N default), 1 (one CTA per row-pair group), 2 (persistent, register-pipelined) or 3 (persistent, bulk-async staged)");
    c->row_mode = mode;
    drop_plans(c);
    return OW_OK;
}

int ow_set_column_kernel(ow_ctx* c, int32_t mode, int32_t fused) {
    if (!c) return OW_ERR_INVALID;
    if (mode < 0 || mode > 4 || fused < -1 || fused > 1)
        return fail(c, OW_ERR_INVALID, "ow_set_column_kernel: mode in 0..4 (0 = per-N default, 1 = ow_col_kernel, 2 = ow_col2_kernel, 3 = ow_col2_kernel TMA-staged, 4 = ow_col_pipe_kernel), fused in -1..1");
    c->col_mode = mode;
    c->fuse_mode = fused;
    drop_plans(c);
    return OW_OK;
}

int ow_get_kernel_modes(ow_ctx* c, int32_t* row, int32_t* column, int32_t* fused) {
    if (!c || !row || !column || !fused) return OW_ERR_INVALID;
    int r, k, f;
    effective_modes(buffers(c), &r, &k, &f);
    *row = r; *column = k; *fused = f;
    return OW_OK;
}

int ow_set_resident_ctas(ow_ctx* c, int32_t row_per_sm, int32_t col_per_sm) {
    if (!c || row_per_sm < 0 || col_per_sm < 0) return OW_ERR_INVALID;
    c->cap_row = row_per_sm;
    c->cap_col = col_per_sm;
    drop_plans(c);
    return OW_OK;
}

int ow_set_frame_kernel(ow_ctx* c, int32_t mode) {
    if (!c || mode < 0 || mode > 1) return OW_ERR_INVALID;
    if (mode == 1 && (!mega_supported(c->N) || c->kcfg.mega_ctas < 1))
        return fail(c, OW_ERR_INVALID, "ow_set_frame_kernel: the one-kernel frame exists for N = 256, 512 and 1024");
    c->frame_mode = mode;
    drop_plans(c);
    return OW_OK;
}

int ow_set_latency_shapes(ow_ctx* c, int32_t on) {
    if (!c) return OW_ERR_INVALID;
    c->latency_shapes = on < 0 ? 0 : on > 2 ? 2 : on;     // 2 (tuning): the wide row shape for every launch, not only single frames
    drop_plans(c);
    return OW_OK;
}

int ow_set_line_clusters(ow_ctx* c, int32_t mode) {
    if (!c || mode < -1 || mode > 7) return OW_ERR_INVALID;
    c->line_clusters = mode;
    return OW_OK;
}

int ow_get_line_clusters(ow_ctx* c) { return c ? buffers(c).big_cluster : 0; }

int ow_set_l2_persist(ow_ctx* c, int32_t mode) {
    if (!c || mode < -1 || mode > 1) return OW_ERR_INVALID;
    OW_CUDA(c, cudaSetDevice(c->device));
    if (c->l2_user_stream) apply_l2_window(c, c->l2_user_stream, false);
    c->l2_persist = mode;
    configure_l2(c);
    return OW_OK;
}

int ow_set_discard_intermediate(ow_ctx* c, int32_t on) {
    if (!c) return OW_ERR_INVALID;
    c->discard_inter = on ? 1 : 0;
    drop_plans(c);
    return OW_OK;
}

int ow_sync(ow_ctx* c, void* stream) {
    if (!c) return OW_ERR_INVALID;
    OW_CUDA(c, cudaSetDevice(c->device));
    OW_CUDA(c, cudaStreamSynchronize(pick(c, stream)));
    return OW_OK;
}

int ow_get_outputs(ow_ctx* c, int32_t slot, ow_outputs* out) {
    if (!c || !out) return OW_ERR_INVALID;
    if (slot < 0 || slot >= c->n_slots) return fail(c, OW_ERR_INVALID, "ow_get_outputs: slot out of range");
    const size_t nn = (size_t)c->N * c->N;
    out->N = c->N;
    out->dy = c->d_disp + (size_t)slot * 3 * nn;
    out->dx = out->dy + nn;
    out->dz = out->dx + nn;
    out->normal = reinterpret_cast<float*>(c->d_normal + (size_t)slot * nn);
    out->jacobian = c->d_jac ? c->d_jac + (size_t)slot * nn : nullptr;
    return OW_OK;
}

size_t ow_frame_bytes(const ow_ctx* c) {
    if (!c) return 0;
    const size_t nn = (size_t)c->N * c->N;
    return nn * sizeof(float) * (3 + 4 + (c->d_jac ? 1 : 0));
}

int ow_download(ow_ctx* c, int32_t index, int32_t which, void* host, size_t bytes, void* stream) {
    if (!c || !host) return OW_ERR_INVALID;
    OW_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = pick(c, stream);
    const size_t nn = (size_t)c->N * c->N;
    const void* src = nullptr;
    size_t need = 0;
    if (which == OW_IMG_H0K || which == OW_IMG_H0MINUSK) {
        if (index < 0 || index >= c->n_cascades) return fail(c, OW_ERR_INVALID, "ow_download: cascade out of range");
        need = nn * 2 * sizeof(float);
        if (bytes != need) return fail(c, OW_ERR_INVALID, "ow_download: size mismatch");
        OW_CUDA(c, launch_split_h0(c->d_h0 + (size_t)index * nn, c->d_tmp, c->d_tmp + nn * 2, (int)nn, st));
        src = which == OW_IMG_H0K ? c->d_tmp : c->d_tmp + nn * 2;
    } else {
        if (index < 0 || index >= c->n_slots) return fail(c, OW_ERR_INVALID, "ow_download: slot out of range");
        switch (which) {
            case OW_IMG_DY: src = c->d_disp + ((size_t)index * 3 + 0) * nn; need = nn * sizeof(float); break;
            case OW_IMG_DX: src = c->d_disp + ((size_t)index * 3 + 1) * nn; need = nn * sizeof(float); break;
            case OW_IMG_DZ: src = c->d_disp + ((size_t)index * 3 + 2) * nn; need = nn * sizeof(float); break;
            case OW_IMG_NORMAL: src = c->d_normal + (size_t)index * nn; need = nn * sizeof(float4); break;
            case OW_IMG_JACOBIAN:
                if (!c->d_jac) return fail(c, OW_ERR_STATE, "ow_download: context created without OW_FLAG_JACOBIAN");
                src = c->d_jac + (size_t)index * nn; need = nn * sizeof(float); break;
            default: return fail(c, OW_ERR_INVALID, "ow_download: unknown image");
        }
        if (bytes != need) return fail(c, OW_ERR_INVALID, "ow_download: size mismatch");
    }
    OW_CUDA(c, cudaMemcpyAsync(host, src, need, cudaMemcpyDeviceToHost, st));
    OW_CUDA(c, cudaStreamSynchronize(st));
    return OW_OK;
}

int ow_download_frame_async(ow_ctx* c, int32_t slot, void* host, size_t bytes, void* stream) {
    if (!c || !host) return OW_ERR_INVALID;
    if (slot < 0 || slot >= c->n_slots) return fail(c, OW_ERR_INVALID, "ow_download_frame_async: slot out of range");
    if (bytes != ow_frame_bytes(c)) return fail(c, OW_ERR_INVALID, "ow_download_frame_async: size mismatch");
    OW_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = pick(c, stream);
    const size_t nn = (size_t)c->N * c->N;
    char* h = static_cast<char*>(host);
    OW_CUDA(c, cudaMemcpyAsync(h, c->d_disp + (size_t)slot * 3 * nn, nn * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
    OW_CUDA(c, cudaMemcpyAsync(h + nn * 3 * sizeof(float), c->d_normal + (size_t)slot * nn, nn * sizeof(float4), cudaMemcpyDeviceToHost, st));
    if (c->d_jac)
        OW_CUDA(c, cudaMemcpyAsync(h + nn * 7 * sizeof(float), c->d_jac + (size_t)slot * nn, nn * sizeof(float), cudaMemcpyDeviceToHost, st));
    return OW_OK;
}

int ow_get_packed(ow_ctx* c, int32_t slot, ow_packed* out) {
    if (!c || !out) return OW_ERR_INVALID;
    if (!c->d_packed) return fail(c, OW_ERR_STATE, "ow_get_packed: context created without OW_FLAG_PACKED_F32 / OW_FLAG_PACKED_F16");
    if (slot < 0 || slot >= c->n_slots) return fail(c, OW_ERR_INVALID, "ow_get_packed: slot out of range");
    out->N = c->N;
    out->displacement_texel_bytes = c->pk.half ? 8 : 16;
    out->displacement = c->d_packed + (size_t)slot * c->pk.slot_bytes;
    out->normal_xz = c->d_packed + (size_t)slot * c->pk.slot_bytes + c->pk.normal_offset;
    return OW_OK;
}

size_t ow_packed_bytes(const ow_ctx* c) { return (c && c->d_packed) ? c->pk.slot_bytes : 0; }

int ow_download_packed_async(ow_ctx* c, int32_t slot, void* host, size_t bytes, void* stream) {
    if (!c || !host) return OW_ERR_INVALID;
    if (!c->d_packed) return fail(c, OW_ERR_STATE, "ow_download_packed_async: context created without OW_FLAG_PACKED_F32 / OW_FLAG_PACKED_F16");
    if (slot < 0 || slot >= c->n_slots) return fail(c, OW_ERR_INVALID, "ow_download_packed_async: slot out of range");
    if (bytes != c->pk.slot_bytes) return fail(c, OW_ERR_INVALID, "ow_download_packed_async: size mismatch");
    OW_CUDA(c, cudaSetDevice(c->device));
    OW_CUDA(c, cudaMemcpyAsync(host, c->d_packed + (size_t)slot * c->pk.slot_bytes, bytes, cudaMemcpyDeviceToHost, pick(c, stream)));
    return OW_OK;
}

int ow_set_group_size(ow_ctx* c, int32_t g) {
    if (!c || g < 0) return OW_ERR_INVALID;
    c->group_size = g;
    drop_plans(c);
    return OW_OK;
}

int ow_set_streams(ow_ctx* c, int32_t n) {
    if (!c || n < 1 || n > ow_ctx::kMaxAux) return c ? fail(c, OW_ERR_INVALID, "ow_set_streams: n must be in [1, 4]") : OW_ERR_INVALID;
    c->n_streams = n;
    drop_plans(c);
    return OW_OK;
}

int ow_last_launch_count(const ow_ctx* c) { return c ? c->last_launches : 0; }
int ow_last_group_count(const ow_ctx* c) { return c ? c->last_groups : 0; }

// ---- multi-cascade composition / the demo's clock ----------------------------------
static int compose_args(ow_ctx* c, int32_t n_terms, const ow_blend_term* terms, float displacement_scale, const char* who, ComposeArgs* A) {
    if (!terms || n_terms < 1 || n_terms > kMaxComposeTerms)
        return fail(c, OW_ERR_INVALID, std::string(who) + ": need 1..16 blend terms");
    A->disp = c->d_disp; A->normal = c->d_normal; A->N = c->N; A->n_terms = n_terms; A->displacement_scale = displacement_scale;
    for (int i = 0; i < n_terms; ++i) {
        const int slot = terms[i].slot;
        if (slot < 0 || slot >= c->n_slots) return fail(c, OW_ERR_INVALID, std::string(who) + ": slot out of range");
        const ow_ctx::SlotPatch& sp = c->slot_patch[slot];
        if (!(sp.L > 0.0f)) return fail(c, OW_ERR_STATE, std::string(who) + ": a referenced slot has not been stepped yet");
        A->term[i].inv_L = 1.0 / (double)sp.L;
        A->term[i].slot = slot;
        A->term[i].weight = terms[i].weight;
        A->term[i].choppiness = sp.choppiness;
    }
    return OW_OK;
}

int ow_sample_points(ow_ctx* c, int32_t n_terms, const ow_blend_term* terms, float displacement_scale, int32_t n_points, const float* xz,
                     float* out, void* stream) {
    if (!c) return OW_ERR_INVALID;
    if (n_points < 0 || (n_points > 0 && (!xz || !out))) return fail(c, OW_ERR_INVALID, "ow_sample_points: bad point arguments");
    ComposeArgs A{};
    const int rc = compose_args(c, n_terms, terms, displacement_scale, "ow_sample_points", &A);
    if (rc != OW_OK) return rc;
    OW_CUDA(c, cudaSetDevice(c->device));
    OW_CUDA(c, launch_sample_points(A, n_points, reinterpret_cast<const float2*>(xz), reinterpret_cast<float4*>(out), pick(c, stream)));
    return OW_OK;
}

int ow_sample_points_host(ow_ctx* c, int32_t n_terms, const ow_blend_term* terms, float displacement_scale, int32_t n_points, const float* xz,
                          float* out, void* stream) {
    if (!c) return OW_ERR_INVALID;
    if (n_points < 0 || (n_points > 0 && (!xz || !out))) return fail(c, OW_ERR_INVALID, "ow_sample_points_host: bad point arguments");
    ComposeArgs A{};
    const int rc = compose_args(c, n_terms, terms, displacement_scale, "ow_sample_points_host", &A);
    if (rc != OW_OK) return rc;
    if (n_points == 0) return OW_OK;
    OW_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = pick(c, stream);
    if ((size_t)n_points > c->query_cap) {
        OW_CUDA(c, cudaStreamSynchronize(st));                 // an earlier query on this stream may still use the old buffer
        cudaFree(c->d_query);
        c->d_query = nullptr; c->query_cap = 0;
        OW_CUDA(c, cudaMalloc(&c->d_query, (size_t)n_points * 10 * sizeof(float)));
        c->query_cap = (size_t)n_points;
    }
    float* d_xz = c->d_query + (size_t)c->query_cap * 8;       // results first: they are float4-aligned
    OW_CUDA(c, cudaMemcpyAsync(d_xz, xz, (size_t)n_points * 2 * sizeof(float), cudaMemcpyHostToDevice, st));
    OW_CUDA(c, launch_sample_points(A, n_points, reinterpret_cast<const float2*>(d_xz), reinterpret_cast<float4*>(c->d_query), st));
    OW_CUDA(c, cudaMemcpyAsync(out, c->d_query, (size_t)n_points * 8 * sizeof(float), cudaMemcpyDeviceToHost, st));
    OW_CUDA(c, cudaStreamSynchronize(st));
    return OW_OK;
}

int ow_compose_grid(ow_ctx* c, int32_t n_terms, const ow_blend_term* terms, float displacement_scale, int32_t M, float origin_x, float origin_z,
                    float extent, float* out_offset, float* out_normal, void* stream) {
    if (!c) return OW_ERR_INVALID;
    if (M < 1 || M > 32768 || !(extent > 0.0f) || !out_offset || !out_normal) return fail(c, OW_ERR_INVALID, "ow_compose_grid: bad grid arguments");
    ComposeArgs A{};
    const int rc = compose_args(c, n_terms, terms, displacement_scale, "ow_compose_grid", &A);
    if (rc != OW_OK) return rc;
    OW_CUDA(c, cudaSetDevice(c->device));
    OW_CUDA(c, launch_compose_grid(A, M, origin_x, origin_z, extent, reinterpret_cast<float4*>(out_offset), reinterpret_cast<float4*>(out_normal),
                                   pick(c, stream)));
    return OW_OK;
}

int ow_set_time_scale(ow_ctx* c, float scale, float offset) {
    if (!c) return OW_ERR_INVALID;
    if (!std::isfinite(scale) || !std::isfinite(offset)) return fail(c, OW_ERR_INVALID, "ow_set_time_scale: scale and offset must be finite");
    c->time_scale = scale;
    c->time_offset = offset;
    return OW_OK;
}

int ow_step_wall_clock(ow_ctx* c, double wall_seconds, void* stream) {
    if (!c) return OW_ERR_INVALID;
    return ow_step(c, c->time_offset + c->time_scale * (float)wall_seconds, stream);
}

// ---- CUDA-GL interop -------------------------------------------------------------------------------
// The image has no GL headers, so the three cudart entry points are declared here with GL's own scalar
// typedefs (GLuint/GLenum are 32-bit unsigned by the GL spec). They are exported by libcudart.
typedef unsigned int ow_GLuint;
typedef unsigned int ow_GLenum;
extern cudaError_t cudaGraphicsGLRegisterImage(struct cudaGraphicsResource** resource, ow_GLuint image, ow_GLenum target, unsigned int flags);
#define OW_GL_TEXTURE_2D 0x0DE1

static int gl_register_n(ow_ctx* c, const uint32_t* tex, int n) {
    if (c->gl_registered) ow_gl_unregister(c);
    cudaSetDevice(c->device);
    for (int i = 0; i < n; ++i) {
        cudaError_t e = cudaGraphicsGLRegisterImage(&c->gl_res[i], tex[i], OW_GL_TEXTURE_2D, cudaGraphicsRegisterFlagsWriteDiscard);
        if (e != cudaSuccess) {
            for (int j = 0; j < i; ++j) { cudaGraphicsUnregisterResource(c->gl_res[j]); c->gl_res[j] = nullptr; }
            c->gl_res[i] = nullptr;
            cudaGetLastError();
            return fail(c, OW_ERR_NO_GL, std::string("cudaGraphicsGLRegisterImage: ") + cudaGetErrorString(e) +
                                             " (is the caller's GL context current on this thread?)");
        }
    }
    c->gl_count = n;
    c->gl_registered = true;
    return OW_OK;
}

int ow_gl_register(ow_ctx* c, uint32_t tex_dy, uint32_t tex_dx, uint32_t tex_dz, uint32_t tex_normal) {
    if (!c) return OW_ERR_INVALID;
    const uint32_t tex[4] = {tex_dy, tex_dx, tex_dz, tex_normal};
    return gl_register_n(c, tex, 4);
}

int ow_gl_register_packed(ow_ctx* c, uint32_t tex_displacement, uint32_t tex_normal_xz) {
    if (!c) return OW_ERR_INVALID;
    if (!c->d_packed) return fail(c, OW_ERR_STATE, "ow_gl_register_packed: context created without OW_FLAG_PACKED_F32 / OW_FLAG_PACKED_F16");
    const uint32_t tex[2] = {tex_displacement, tex_normal_xz};
    return gl_register_n(c, tex, 2);
}

int ow_gl_unregister(ow_ctx* c) {
    if (!c) return OW_ERR_INVALID;
    if (c->gl_registered) {
        cudaSetDevice(c->device);
        for (auto& r : c->gl_res) { if (r) cudaGraphicsUnregisterResource(r); r = nullptr; }
        c->gl_registered = false;
        c->gl_count = 0;
    }
    return OW_OK;
}

int ow_gl_step(ow_ctx* c, float t) {
    if (!c) return OW_ERR_INVALID;
    if (!c->gl_registered) return fail(c, OW_ERR_NO_GL, "ow_gl_step: ow_gl_register has not succeeded");
    int r = ow_step(c, t, nullptr);
    if (r != OW_OK) return r;
    OW_CUDA(c, cudaGraphicsMapResources(c->gl_count, c->gl_res, c->stream));
    const size_t nn = (size_t)c->N * c->N;
    for (int i = 0; i < c->gl_count; ++i) {
        cudaArray_t arr = nullptr;
        OW_CUDA(c, cudaGraphicsSubResourceGetMappedArray(&arr, c->gl_res[i], 0, 0));
        const void* src;
        size_t texel;
        if (c->gl_count == 2) {       // packed set of slot 0: displacement (RGBA32F / RGBA16F), normal_xz (RG16_SNORM)
            src = i == 0 ? (const void*)c->d_packed : (const void*)(c->d_packed + c->pk.normal_offset);
            texel = i == 0 ? (c->pk.half ? 8 : 16) : 4;
        } else {
            src = i < 3 ? (const void*)(c->d_disp + (size_t)i * nn) : (const void*)c->d_normal;
            texel = i < 3 ? sizeof(float) : sizeof(float4);
        }
        const size_t pitch = (size_t)c->N * texel;
        OW_CUDA(c, cudaMemcpy2DToArrayAsync(arr, 0, 0, src, pitch, pitch, c->N, cudaMemcpyDeviceToDevice, c->stream));
    }
    OW_CUDA(c, cudaGraphicsUnmapResources(c->gl_count, c->gl_res, c->stream));   // unmap orders the copies before GL's next use
    return OW_OK;
}

}  // extern "C"
