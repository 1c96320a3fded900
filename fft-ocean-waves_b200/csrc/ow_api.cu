// ow_api.cu — the C ABI declared in include/oceanwaves.h: context, buffers, launch sequencing.
// Host side of the drop-in: what FFTOceanWaves::init()/update() do for the sim (reference
// src/main.cpp:199-255, 553-744, 1083-1145) minus windowing, rendering and GL plumbing.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
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
    bool spectrum_ready = false;
    int group_size = 0;
    int last_launches = 0, last_groups = 0;
    std::string err;
    // GL interop
    cudaGraphicsResource* gl_res[4] = {nullptr, nullptr, nullptr, nullptr};
    bool gl_registered = false;
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
    fb.N = c->N; fb.h0 = c->d_h0; fb.hp = c->d_hp; fb.nyq = c->d_nyq; fb.ktab = c->d_ktab; fb.casc = c->d_casc; fb.inter = c->d_inter;
    fb.disp = c->d_disp; fb.normal = c->d_normal; fb.jacobian = c->d_jac;
    // Measured on B200 (profiles/r01d_discard_ab.txt): dropping the consumed intermediate from L2 changes nothing — the
    // frame is not bound by DRAM write-back — so it stays off; OW_DISCARD=1 turns it on for experiments.
    static const int discard = getenv("OW_DISCARD") ? 1 : 0;
    fb.discard_inter = discard;
    fb.four_step = (c->flags & OW_FLAG_FOUR_STEP) ? 1 : 0;
    fb.scratch = c->d_scratch;
    fb.fuse_normals = (c->flags & OW_FLAG_FUSED_NORMALS) && !(c->flags & OW_FLAG_JACOBIAN) && c->N <= 2048 ? 1 : 0;
    return fb;
}

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
        gmax = (int)((c->N <= 512 ? 400.0e6 : 100.0e6) / (12.0 * c->N * (double)c->N));
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
    cudaFree(c->d_noise); cudaFree(c->d_h0); cudaFree(c->d_hp); cudaFree(c->d_nyq); cudaFree(c->d_ktab); cudaFree(c->d_casc); cudaFree(c->d_inter);
    cudaFree(c->d_disp); cudaFree(c->d_normal); cudaFree(c->d_jac); cudaFree(c->d_tmp); cudaFree(c->d_scratch);
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
    if (!frame_supported(N)) return fail(nullptr, OW_ERR_INVALID, "ow_create: N must be a power of two in [256, 32768]");
    if ((flags & OW_FLAG_FOUR_STEP) && !big_supported(N, true))
        return fail(nullptr, OW_ERR_INVALID, "ow_create: OW_FLAG_FOUR_STEP is a test mode for N = 1024 or 2048");
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
    const size_t nn = (size_t)N * N;
#define OW_TRY(call)                                                                 \
    do {                                                                             \
        cudaError_t e__ = (call);                                                    \
        if (e__ != cudaSuccess) { int r__ = cuda_fail(nullptr, e__, #call); release(c); return r__; } \
    } while (0)
    OW_TRY(cudaSetDevice(device));
    OW_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (int i = 0; i < ow_ctx::kMaxAux; ++i) {
        OW_TRY(cudaStreamCreateWithFlags(&c->aux[i], cudaStreamNonBlocking));
        OW_TRY(cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming));
    }
    OW_TRY(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    OW_TRY(cudaMalloc(&c->d_h0, nn * n_cascades * sizeof(float4)));
    OW_TRY(cudaMalloc(&c->d_hp, nn / 2 * n_cascades * sizeof(float4)));
    OW_TRY(cudaMalloc(&c->d_nyq, (size_t)(N / 2) * n_cascades * sizeof(float4)));
    OW_TRY(cudaMemsetAsync(c->d_hp, 0, nn / 2 * n_cascades * sizeof(float4), c->stream));
    OW_TRY(cudaMemsetAsync(c->d_nyq, 0, (size_t)(N / 2) * n_cascades * sizeof(float4), c->stream));
    OW_TRY(cudaMalloc(&c->d_ktab, (size_t)N * n_cascades * sizeof(float)));
    OW_TRY(cudaMalloc(&c->d_casc, n_cascades * sizeof(CascadeDev)));
    OW_TRY(cudaMalloc(&c->d_inter, nn / 2 * 3 * n_slots * sizeof(float2)));
    OW_TRY(cudaMalloc(&c->d_disp, nn * 3 * n_slots * sizeof(float)));
    OW_TRY(cudaMalloc(&c->d_normal, nn * n_slots * sizeof(float4)));
    if (flags & OW_FLAG_JACOBIAN) OW_TRY(cudaMalloc(&c->d_jac, nn * n_slots * sizeof(float)));
    OW_TRY(cudaMalloc(&c->d_tmp, nn * 4 * sizeof(float)));
    if (big_supported(N, false) || (flags & OW_FLAG_FOUR_STEP)) OW_TRY(cudaMalloc(&c->d_scratch, nn / 2 * 3 * sizeof(float2)));
    OW_TRY(configure_frame_kernels(N));
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
    c->spectrum_ready = false;
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
    }
    OW_CUDA(c, cudaStreamSynchronize(c->stream));   // host planes may be freed by the caller on return
    c->spectrum_ready = false;
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
    }
    if (!c->d_noise) {
        OW_CUDA(c, cudaMalloc(&c->d_noise, plane * 4 * c->n_cascades));
        c->noise_w = N; c->noise_h = N;
    }
    const int lo = cascade < 0 ? 0 : cascade, hi = cascade < 0 ? c->n_cascades : cascade + 1;
    for (int i = lo; i < hi; ++i) {
        OW_CUDA(c, launch_noise_seed(c->d_noise + (size_t)i * 4 * plane, N, seed, c->stream));
        c->noise_set[i] = 1;
    }
    OW_CUDA(c, cudaStreamSynchronize(c->stream));
    c->spectrum_ready = false;
    return OW_OK;
}

int ow_init_spectrum(ow_ctx* c) {
    if (!c) return OW_ERR_INVALID;
    for (int i = 0; i < c->n_cascades; ++i)
        if (!c->noise_set[i]) return fail(c, OW_ERR_STATE, "ow_init_spectrum: ow_set_noise has not been called for every cascade");
    OW_CUDA(c, cudaSetDevice(c->device));
    c->casc_host.resize(c->n_cascades);
    const size_t nn = (size_t)c->N * c->N, plane = (size_t)c->noise_w * c->noise_h;
    for (int i = 0; i < c->n_cascades; ++i) {
        c->casc_host[i] = to_dev(c->params[i]);
        OW_CUDA(c, launch_ktab(c->d_ktab + (size_t)i * c->N, c->N, c->params[i].L, c->stream));
        OW_CUDA(c, launch_h0(c->d_h0 + (size_t)i * nn, c->d_noise + (size_t)i * 4 * plane, c->noise_w, c->noise_h, c->N,
                             c->casc_host[i], c->stream));
        OW_CUDA(c, launch_fold(c->d_h0 + (size_t)i * nn, c->d_hp + (size_t)i * (nn / 2), c->d_nyq + (size_t)i * (c->N / 2), c->N, c->stream));
    }
    OW_CUDA(c, cudaMemcpyAsync(c->d_casc, c->casc_host.data(), c->n_cascades * sizeof(CascadeDev), cudaMemcpyHostToDevice, c->stream));
    OW_CUDA(c, cudaStreamSynchronize(c->stream));   // the reference ends tilde_h0_k() with glFinish (main.cpp:582)
    c->spectrum_ready = true;
    return OW_OK;
}

int ow_set_h0(ow_ctx* c, int32_t cascade, const float* h0k, const float* h0minusk) {
    if (!c || !h0k || !h0minusk) return OW_ERR_INVALID;
    if (cascade < 0 || cascade >= c->n_cascades) return fail(c, OW_ERR_INVALID, "ow_set_h0: cascade out of range");
    OW_CUDA(c, cudaSetDevice(c->device));
    const size_t nn = (size_t)c->N * c->N;
    OW_CUDA(c, cudaMemcpyAsync(c->d_tmp, h0k, nn * 2 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    OW_CUDA(c, cudaMemcpyAsync(c->d_tmp + nn * 2, h0minusk, nn * 2 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    OW_CUDA(c, launch_merge_h0(c->d_h0 + (size_t)cascade * nn, c->d_tmp, c->d_tmp + nn * 2, (int)nn, c->stream));
    OW_CUDA(c, launch_fold(c->d_h0 + (size_t)cascade * nn, c->d_hp + (size_t)cascade * (nn / 2), c->d_nyq + (size_t)cascade * (c->N / 2), c->N, c->stream));
    if (!c->spectrum_ready) {
        // k table and cascade constants are needed even when h0 is supplied directly
        c->casc_host.resize(c->n_cascades);
        for (int i = 0; i < c->n_cascades; ++i) {
            c->casc_host[i] = to_dev(c->params[i]);
            OW_CUDA(c, launch_ktab(c->d_ktab + (size_t)i * c->N, c->N, c->params[i].L, c->stream));
        }
        OW_CUDA(c, cudaMemcpyAsync(c->d_casc, c->casc_host.data(), c->n_cascades * sizeof(CascadeDev), cudaMemcpyHostToDevice, c->stream));
    }
    OW_CUDA(c, cudaStreamSynchronize(c->stream));
    c->spectrum_ready = true;
    return OW_OK;
}

static int step_impl(ow_ctx* c, int32_t count, const int32_t* cascade_of_slot, const float* time_of_slot, void* stream,
                     float* kernel_ms) {
    if (!c || !cascade_of_slot || !time_of_slot) return OW_ERR_INVALID;
    if (!c->spectrum_ready) return fail(c, OW_ERR_STATE, "ow_step: call ow_init_spectrum (or ow_set_h0) first");
    if (count < 1 || count > c->n_slots) return fail(c, OW_ERR_INVALID, "ow_step_multi: count out of range");
    for (int i = 0; i < count; ++i)
        if (cascade_of_slot[i] < 0 || cascade_of_slot[i] >= c->n_cascades)
            return fail(c, OW_ERR_INVALID, "ow_step_multi: cascade index out of range");
    OW_CUDA(c, cudaSetDevice(c->device));
    cudaStream_t st = pick(c, stream);
    const FrameBuffers fb = buffers(c);
    const int group = pick_group(c, count);
    int launches = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    if (kernel_ms) {
        kernel_ms[0] = kernel_ms[1] = kernel_ms[2] = 0.0f;
        for (auto& e : ev) OW_CUDA(c, cudaEventCreate(&e));
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
        cudaStream_t gst = nfan ? c->aux[gi % nfan] : st;
        SlotTable tab{};
        bool fast = (c->flags & OW_FLAG_EXACT_SINCOS) == 0;
        for (int i = 0; i < n; ++i) {
            tab.cascade[i] = cascade_of_slot[base + i];
            tab.time[i] = time_of_slot[base + i];
            tab.slot[i] = base + i;
            // largest phase of this cascade: w_max = sqrt(g*|k|max), |k|max = sqrt(2)*pi*N/L
            const float kmax = 1.41421356f * 3.14159265f * (float)c->N / c->params[tab.cascade[i]].L;
            if (!(sqrtf(9.81f * kmax) * fabsf(tab.time[i]) < kFastPhaseLimit)) fast = false;
        }
        const int k = launch_frame(fb, tab, n, (c->flags & OW_FLAG_JACOBIAN) != 0, fast, gst, kernel_ms ? ev : nullptr);
        if (k < 0) return cuda_fail(c, cudaGetLastError(), "launch_frame");
        launches += k;
        if (kernel_ms) {
            OW_CUDA(c, cudaEventSynchronize(ev[3]));
            for (int i = 0; i < 3; ++i) {
                float ms = 0.0f;
                OW_CUDA(c, cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
                kernel_ms[i] += ms;
            }
        }
    }
    for (int i = 0; i < nfan; ++i) {
        OW_CUDA(c, cudaEventRecord(c->ev_join[i], c->aux[i]));
        OW_CUDA(c, cudaStreamWaitEvent(st, c->ev_join[i], 0));
    }
    if (kernel_ms) for (auto& e : ev) cudaEventDestroy(e);
    c->last_launches = launches;
    c->last_groups = ngroups;
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

int ow_step(ow_ctx* c, float t, void* stream) {
    if (!c) return OW_ERR_INVALID;
    std::vector<int32_t> cs(c->n_cascades);
    std::vector<float> ts(c->n_cascades, t);
    for (int i = 0; i < c->n_cascades; ++i) cs[i] = i;
    return ow_step_multi(c, c->n_cascades, cs.data(), ts.data(), stream);
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

int ow_set_group_size(ow_ctx* c, int32_t g) {
    if (!c || g < 0) return OW_ERR_INVALID;
    c->group_size = g;
    return OW_OK;
}

int ow_set_streams(ow_ctx* c, int32_t n) {
    if (!c || n < 1 || n > ow_ctx::kMaxAux) return c ? fail(c, OW_ERR_INVALID, "ow_set_streams: n must be in [1, 4]") : OW_ERR_INVALID;
    c->n_streams = n;
    return OW_OK;
}

int ow_last_launch_count(const ow_ctx* c) { return c ? c->last_launches : 0; }
int ow_last_group_count(const ow_ctx* c) { return c ? c->last_groups : 0; }

// ---- CUDA-GL interop -------------------------------------------------------------------------------
// The image has no GL headers, so the three cudart entry points are declared here with GL's own scalar
// typedefs (GLuint/GLenum are 32-bit unsigned by the GL spec). They are exported by libcudart.
typedef unsigned int ow_GLuint;
typedef unsigned int ow_GLenum;
extern cudaError_t cudaGraphicsGLRegisterImage(struct cudaGraphicsResource** resource, ow_GLuint image, ow_GLenum target, unsigned int flags);
#define OW_GL_TEXTURE_2D 0x0DE1

int ow_gl_register(ow_ctx* c, uint32_t tex_dy, uint32_t tex_dx, uint32_t tex_dz, uint32_t tex_normal) {
    if (!c) return OW_ERR_INVALID;
    if (c->gl_registered) ow_gl_unregister(c);
    cudaSetDevice(c->device);
    const uint32_t tex[4] = {tex_dy, tex_dx, tex_dz, tex_normal};
    for (int i = 0; i < 4; ++i) {
        cudaError_t e = cudaGraphicsGLRegisterImage(&c->gl_res[i], tex[i], OW_GL_TEXTURE_2D, cudaGraphicsRegisterFlagsWriteDiscard);
        if (e != cudaSuccess) {
            for (int j = 0; j < i; ++j) { cudaGraphicsUnregisterResource(c->gl_res[j]); c->gl_res[j] = nullptr; }
            c->gl_res[i] = nullptr;
            cudaGetLastError();
            return fail(c, OW_ERR_NO_GL, std::string("cudaGraphicsGLRegisterImage: ") + cudaGetErrorString(e) +
                                             " (is the caller's GL context current on this thread?)");
        }
    }
    c->gl_registered = true;
    return OW_OK;
}

int ow_gl_unregister(ow_ctx* c) {
    if (!c) return OW_ERR_INVALID;
    if (c->gl_registered) {
        cudaSetDevice(c->device);
        for (auto& r : c->gl_res) { if (r) cudaGraphicsUnregisterResource(r); r = nullptr; }
        c->gl_registered = false;
    }
    return OW_OK;
}

int ow_gl_step(ow_ctx* c, float t) {
    if (!c) return OW_ERR_INVALID;
    if (!c->gl_registered) return fail(c, OW_ERR_NO_GL, "ow_gl_step: ow_gl_register has not succeeded");
    int r = ow_step(c, t, nullptr);
    if (r != OW_OK) return r;
    OW_CUDA(c, cudaGraphicsMapResources(4, c->gl_res, c->stream));
    const size_t nn = (size_t)c->N * c->N;
    for (int i = 0; i < 4; ++i) {
        cudaArray_t arr = nullptr;
        OW_CUDA(c, cudaGraphicsSubResourceGetMappedArray(&arr, c->gl_res[i], 0, 0));
        const void* src = i < 3 ? (const void*)(c->d_disp + (size_t)i * nn) : (const void*)c->d_normal;
        const size_t pitch = (size_t)c->N * (i < 3 ? sizeof(float) : sizeof(float4));
        OW_CUDA(c, cudaMemcpy2DToArrayAsync(arr, 0, 0, src, pitch, pitch, c->N, cudaMemcpyDeviceToDevice, c->stream));
    }
    OW_CUDA(c, cudaGraphicsUnmapResources(4, c->gl_res, c->stream));   // unmap orders the copies before GL's next use
    return OW_OK;
}

}  // extern "C"
