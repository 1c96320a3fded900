// ow_slab.cu — C ABI of the slab-decomposed frame: ONE N x N grid spread over the GPUs of a box
// (BASELINE config C5). One process per GPU; every process creates an ow_slab for its rank.
//
//   rank r owns   row pairs  p in [r*PL, (r+1)*PL), PL = N/2/world   (rows p and N-p: the Hermitian mirror is local)
//          and    columns    x in [r*XL, (r+1)*XL), XL = N/world
//   ow_slab_rows : h0 -> h(k,t) -> row IFFT of this rank's pairs (the reference's tilde_h0_t + horizontal butterfly passes,
//                  src/main.cpp:587-643), each result stored where its COLUMN owner wants it: the transpose of the
//                  distributed 2-D IFFT is the row kernel's store pattern. Two transports:
//                    OW_SLAB_PEER_STORES  straight into the owners' receive buffers over NVLink (peer mappings opened
//                                         from CUDA IPC handles the host exchanged once); the host only has to order
//                                         "all rows done" before "columns start" (a barrier).
//                    OW_SLAB_SEND_BUFFER  into a local send buffer [dest][PL][3][XH]; the host moves it with ONE equal-split
//                                         all-to-all (NCCL over NVLink) into the owners' receive buffers.
//   ow_slab_cols : receive buffer [N/2][3][XH] -> column IFFT + inversion (src/main.cpp:645-682) -> normals (+Jacobian)
//                  (:687-707) for this rank's columns. Outputs stay column-slabbed: dy/dx/dz [N][XL], normal [N][XL][4].
// Each block carries kSlabHalo wrap-around halo columns either side, so the stencils need no second exchange.
#include <cmath>
#include <cstring>
#include <new>
#include <string>

#include "../../include/oceanwaves.h"
#include "ow_internal.h"

using namespace ow;

struct ow_slab {
    SlabGeom g{};
    int device = 0;
    uint32_t flags = 0;
    ow_params params{};
    CascadeDev casc{};
    float4* d_h0 = nullptr;       // [2*PL][N]
    float4* d_hp = nullptr;       // [PL][N] folded pairs
    float4* d_nyq = nullptr;      // [PL]
    int cluster_caps = 0;         // KernelConfig::big_cluster of this rank's device (N = A*B decomposition as thread-block clusters)
    float* d_ktab = nullptr;      // [N]
    float* d_ktab_sub = nullptr;  // [N] sub-line-major copy (N > 4096: the folded rows d_hp are stored in that order too)
    int sub_A = 0;                // A of the line decomposition N = A*B, 0 for the direct kernels
    float2* d_send = nullptr;     // [world][PL][3][XH]
    float2* d_recv = nullptr;     // [N/2][3][XH] = [world][PL][3][XH]   (receive buffer 0)
    int post_ctas_per_sm = 2;     // ow_slab_set_post_ctas
    int bigcol_pipe_full = 0;     // grid of the persistent pipelined column lines kernel on this device (ow_slab_set_column_lines(4))
    float2* d_recv1 = nullptr;    // receive buffer 1 (ow_slab_enable_double_buffer): frame f+1's rows land here while frame f's columns read buffer 0
    float2* d_scratch_cols = nullptr;   // N > 4096 with two buffers: the column pass gets its own scratch (rows and columns then run concurrently)
    float2* peer_recv1[kSlabMaxWorld] = {};
    bool peers_open1 = false;
    float* d_disp = nullptr;      // [3][N][XH]
    float4* d_normal = nullptr;   // [N][XL]
    float* d_jac = nullptr;       // [N][XL]
    float2* d_scratch = nullptr;  // N > 4096 only: radix-A sums of the line decomposition
    float2* peer_recv[kSlabMaxWorld] = {};   // peer mappings of every rank's d_recv (own entry = d_recv)
    bool peers_open = false;
    bool spectrum_ready = false;
    cudaStream_t stream = nullptr;
    std::string err;
};

static thread_local std::string g_slab_create_error;

namespace {

int sfail(ow_slab* s, int code, const std::string& msg) {
    if (s) s->err = msg; else g_slab_create_error = msg;
    return code;
}
int scuda(ow_slab* s, cudaError_t e, const char* what) { return sfail(s, OW_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e)); }

#define OWS_CUDA(s, call)                                            \
    do {                                                             \
        cudaError_t e__ = (call);                                    \
        if (e__ != cudaSuccess) return scuda((s), e__, #call);       \
    } while (0)

size_t block_elems(const SlabGeom& g) { return (size_t)g.PL * 3 * g.XH; }   // float2 elements one rank sends to one rank

void srelease(ow_slab* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->peers_open)
        for (int h = 0; h < s->g.world; ++h)
            if (h != s->g.rank && s->peer_recv[h]) cudaIpcCloseMemHandle(s->peer_recv[h]);
    if (s->peers_open1)
        for (int h = 0; h < s->g.world; ++h)
            if (h != s->g.rank && s->peer_recv1[h]) cudaIpcCloseMemHandle(s->peer_recv1[h]);
    cudaFree(s->d_recv1); cudaFree(s->d_scratch_cols);
    cudaFree(s->d_h0); cudaFree(s->d_hp); cudaFree(s->d_nyq); cudaFree(s->d_ktab); cudaFree(s->d_ktab_sub); cudaFree(s->d_send); cudaFree(s->d_recv); cudaFree(s->d_disp);
    cudaFree(s->d_normal); cudaFree(s->d_jac); cudaFree(s->d_scratch);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

cudaStream_t spick(ow_slab* s, void* st) { return st ? static_cast<cudaStream_t>(st) : s->stream; }

}  // namespace

extern "C" {

int ow_slab_create(int32_t N, int32_t world, int32_t rank, const ow_params* p, int32_t device, uint32_t flags, ow_slab** out) {
    if (!out) return sfail(nullptr, OW_ERR_INVALID, "ow_slab_create: out is NULL");
    *out = nullptr;
    if (!p) return sfail(nullptr, OW_ERR_INVALID, "ow_slab_create: params is NULL");
    if (!frame_supported(N) || !slab_supported(N, world))
        return sfail(nullptr, OW_ERR_INVALID, "ow_slab_create: unsupported (N, world): N a power of two in [256, 32768], world in {1,2,4,8} with N/world >= 128");
    if (rank < 0 || rank >= world) return sfail(nullptr, OW_ERR_INVALID, "ow_slab_create: rank out of range");
    if (!(p->L > 0.0f) || !(p->wind_speed > 0.0f) || (p->wind_dir[0] == 0.0f && p->wind_dir[1] == 0.0f))
        return sfail(nullptr, OW_ERR_INVALID, "ow_slab_create: invalid parameters");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess) return scuda(nullptr, e, "cudaGetDeviceCount (no usable CUDA device; there is no CPU fallback)");
    if (device < 0 || device >= ndev) return sfail(nullptr, OW_ERR_INVALID, "ow_slab_create: device ordinal out of range");
    ow_slab* s = new (std::nothrow) ow_slab();
    if (!s) return sfail(nullptr, OW_ERR_NOMEM, "ow_slab_create: out of host memory");
    s->g.N = N; s->g.world = world; s->g.rank = rank;
    s->g.PL = N / 2 / world; s->g.XL = N / world; s->g.XH = s->g.XL + 2 * kSlabHalo;
    s->device = device; s->flags = flags; s->params = *p;
    s->casc.L = p->L; s->casc.wind_speed = p->wind_speed; s->casc.amplitude = p->amplitude; s->casc.suppression = p->suppression;
    s->casc.choppiness = p->choppiness;
    const float inv = 1.0f / sqrtf(p->wind_dir[0] * p->wind_dir[0] + p->wind_dir[1] * p->wind_dir[1]);   // glm::normalize, src/main.cpp:555
    s->casc.wdx = p->wind_dir[0] * inv; s->casc.wdy = p->wind_dir[1] * inv;
    const SlabGeom& g = s->g;
#define OWS_TRY(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess) { int r__ = scuda(nullptr, e__, #call); srelease(s); return r__; }      \
    } while (0)
    OWS_TRY(cudaSetDevice(device));
    OWS_TRY(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    OWS_TRY(cudaMalloc(&s->d_h0, (size_t)2 * g.PL * N * sizeof(float4)));
    OWS_TRY(cudaMalloc(&s->d_hp, hp_block_elems(g.PL, N) * sizeof(float4)));     // [PL][N] float4 + [PL][N] float2 (w, 1/|k|)
    OWS_TRY(cudaMalloc(&s->d_nyq, (size_t)g.PL * sizeof(float4)));
    OWS_TRY(cudaMemsetAsync(s->d_hp, 0, hp_block_elems(g.PL, N) * sizeof(float4), s->stream));
    OWS_TRY(cudaMemsetAsync(s->d_nyq, 0, (size_t)g.PL * sizeof(float4), s->stream));
    OWS_TRY(cudaMalloc(&s->d_ktab, (size_t)N * sizeof(float)));
    s->sub_A = big_radix(N, false);
    if (s->sub_A) OWS_TRY(cudaMalloc(&s->d_ktab_sub, (size_t)N * sizeof(float)));
    OWS_TRY(cudaMalloc(&s->d_send, block_elems(g) * world * sizeof(float2)));
    OWS_TRY(cudaMalloc(&s->d_recv, block_elems(g) * world * sizeof(float2)));
    OWS_TRY(cudaMalloc(&s->d_disp, (size_t)3 * N * g.XH * sizeof(float)));
    OWS_TRY(cudaMalloc(&s->d_normal, (size_t)N * g.XL * sizeof(float4)));
    if (flags & OW_FLAG_JACOBIAN) OWS_TRY(cudaMalloc(&s->d_jac, (size_t)N * g.XL * sizeof(float)));
    if (slab_scratch_elems(g)) OWS_TRY(cudaMalloc(&s->d_scratch, slab_scratch_elems(g) * sizeof(float2)));
    KernelConfig kcfg;
    OWS_TRY(configure_frame_kernels(N, &kcfg));
    s->cluster_caps = kcfg.big_cluster;
    s->bigcol_pipe_full = kcfg.sm_count * kcfg.bigcol_pipe_ctas;
    s->g.bigcol_pipe_grid = 0;      // N > 4096: one CTA per column-lines item (ow_slab_set_column_lines; the persistent pipelined kernel measured slower)
    s->g.big_cluster = 0;          // scratch path by default (ow_slab_set_line_clusters)
#undef OWS_TRY
    s->peer_recv[rank] = s->d_recv;
    *out = s;
    return OW_OK;
}

void ow_slab_destroy(ow_slab* s) { srelease(s); }

const char* ow_slab_last_error(const ow_slab* s) { return s ? s->err.c_str() : g_slab_create_error.c_str(); }

int ow_slab_get_info(const ow_slab* s, ow_slab_info* info) {
    if (!s || !info) return OW_ERR_INVALID;
    const SlabGeom& g = s->g;
    info->N = g.N; info->world = g.world; info->rank = g.rank;
    info->pairs_per_rank = g.PL; info->cols_per_rank = g.XL; info->padded_cols = g.XH; info->halo = kSlabHalo;
    info->block_bytes = block_elems(g) * sizeof(float2);
    info->send = s->d_send; info->recv = s->d_recv;
    info->dy = s->d_disp; info->dx = s->d_disp + (size_t)g.N * g.XH; info->dz = s->d_disp + (size_t)2 * g.N * g.XH;
    info->normal = reinterpret_cast<float*>(s->d_normal); info->jacobian = s->d_jac;
    return OW_OK;
}

int ow_slab_init_spectrum_seeded(ow_slab* s, uint64_t seed) {
    if (!s) return OW_ERR_INVALID;
    OWS_CUDA(s, cudaSetDevice(s->device));
    const SlabGeom& g = s->g;
    OWS_CUDA(s, launch_ktab(s->d_ktab, g.N, s->params.L, s->stream));
    OWS_CUDA(s, launch_h0_slab(s->d_h0, g.N, g.rank * g.PL, g.PL, seed, s->casc, s->stream));
    if (s->sub_A) OWS_CUDA(s, launch_ktab_sub(s->d_ktab, s->d_ktab_sub, g.N, s->sub_A, s->stream));
    OWS_CUDA(s, launch_fold_slab(s->d_h0, s->d_hp, s->d_nyq, s->d_ktab, g.N, g.rank * g.PL, g.PL, s->sub_A, s->stream));
    OWS_CUDA(s, cudaStreamSynchronize(s->stream));   // like the reference's glFinish after tilde_h0_k (src/main.cpp:582)
    s->spectrum_ready = true;
    return OW_OK;
}

int ow_slab_enable_double_buffer(ow_slab* s) {
    if (!s) return OW_ERR_INVALID;
    if (s->d_recv1) return OW_OK;
    OWS_CUDA(s, cudaSetDevice(s->device));
    const SlabGeom& g = s->g;
    OWS_CUDA(s, cudaMalloc(&s->d_recv1, block_elems(g) * g.world * sizeof(float2)));
    if (slab_scratch_elems(g)) OWS_CUDA(s, cudaMalloc(&s->d_scratch_cols, slab_scratch_elems(g) * sizeof(float2)));
    s->peer_recv1[g.rank] = s->d_recv1;
    // the row post kernel then shares the SMs with the previous frame's column pass: slim persistent grid (2 CTAs per SM)
    int sms = 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device) != cudaSuccess) { cudaGetLastError(); sms = 148; }
    s->g.post_ctas = (g.world > 1 ? s->post_ctas_per_sm : 0) * sms;
    return OW_OK;
}

static int ipc_handle_of(ow_slab* s, float2* buf, void* handle, size_t bytes) {
    if (bytes != sizeof(cudaIpcMemHandle_t)) return sfail(s, OW_ERR_INVALID, "ow_slab_ipc_handle: handle buffer must be OW_SLAB_IPC_HANDLE_BYTES");
    OWS_CUDA(s, cudaSetDevice(s->device));
    cudaIpcMemHandle_t h;
    OWS_CUDA(s, cudaIpcGetMemHandle(&h, buf));
    std::memcpy(handle, &h, sizeof(h));
    return OW_OK;
}

static int open_peers_of(ow_slab* s, float2** peer, bool* open, const void* handles, size_t bytes) {
    const SlabGeom& g = s->g;
    if (bytes != sizeof(cudaIpcMemHandle_t) * (size_t)g.world) return sfail(s, OW_ERR_INVALID, "ow_slab_open_peers: need world handles");
    if (*open) return OW_OK;
    OWS_CUDA(s, cudaSetDevice(s->device));
    const char* hb = static_cast<const char*>(handles);
    for (int h = 0; h < g.world; ++h) {
        if (h == g.rank) continue;
        cudaIpcMemHandle_t mh;
        std::memcpy(&mh, hb + (size_t)h * sizeof(mh), sizeof(mh));
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, mh, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            for (int k = 0; k < h; ++k) if (k != g.rank && peer[k]) { cudaIpcCloseMemHandle(peer[k]); peer[k] = nullptr; }
            cudaGetLastError();
            return scuda(s, e, "cudaIpcOpenMemHandle (peer access between the GPUs of this box is required for OW_SLAB_PEER_STORES)");
        }
        peer[h] = static_cast<float2*>(ptr);
    }
    *open = true;
    return OW_OK;
}

int ow_slab_ipc_handle(ow_slab* s, void* handle, size_t bytes) {
    if (!s || !handle) return OW_ERR_INVALID;
    return ipc_handle_of(s, s->d_recv, handle, bytes);
}

int ow_slab_ipc_handle_buf(ow_slab* s, int32_t buf, void* handle, size_t bytes) {
    if (!s || !handle || buf < 0 || buf > 1) return OW_ERR_INVALID;
    if (buf == 1 && !s->d_recv1) return sfail(s, OW_ERR_STATE, "ow_slab_ipc_handle_buf: call ow_slab_enable_double_buffer first");
    return ipc_handle_of(s, buf ? s->d_recv1 : s->d_recv, handle, bytes);
}

int ow_slab_open_peers(ow_slab* s, const void* handles, size_t bytes) {
    if (!s || !handles) return OW_ERR_INVALID;
    return open_peers_of(s, s->peer_recv, &s->peers_open, handles, bytes);
}

int ow_slab_open_peers_buf(ow_slab* s, int32_t buf, const void* handles, size_t bytes) {
    if (!s || !handles || buf < 0 || buf > 1) return OW_ERR_INVALID;
    if (buf == 1 && !s->d_recv1) return sfail(s, OW_ERR_STATE, "ow_slab_open_peers_buf: call ow_slab_enable_double_buffer first");
    return buf ? open_peers_of(s, s->peer_recv1, &s->peers_open1, handles, bytes) : open_peers_of(s, s->peer_recv, &s->peers_open, handles, bytes);
}

int ow_slab_rows_buf(ow_slab* s, float t, int32_t transport, int32_t buf, void* stream) {
    if (!s || buf < 0 || buf > 1) return OW_ERR_INVALID;
    if (!s->spectrum_ready) return sfail(s, OW_ERR_STATE, "ow_slab_rows: call ow_slab_init_spectrum_seeded first");
    if (buf == 1 && !s->d_recv1) return sfail(s, OW_ERR_STATE, "ow_slab_rows_buf: call ow_slab_enable_double_buffer first");
    const SlabGeom& g = s->g;
    float2* base[kSlabMaxWorld] = {};
    if (transport == OW_SLAB_PEER_STORES) {
        float2** peer = buf ? s->peer_recv1 : s->peer_recv;
        if (g.world > 1 && !(buf ? s->peers_open1 : s->peers_open)) return sfail(s, OW_ERR_STATE, "ow_slab_rows: OW_SLAB_PEER_STORES needs ow_slab_open_peers");
        for (int h = 0; h < g.world; ++h) base[h] = peer[h] + (size_t)g.rank * block_elems(g);
    } else if (transport == OW_SLAB_SEND_BUFFER) {
        for (int h = 0; h < g.world; ++h) base[h] = s->d_send + (size_t)h * block_elems(g);
    } else {
        return sfail(s, OW_ERR_INVALID, "ow_slab_rows: unknown transport");
    }
    OWS_CUDA(s, cudaSetDevice(s->device));
    bool fast = (s->flags & OW_FLAG_EXACT_SINCOS) == 0;
    const float kmax = 1.41421356f * 3.14159265f * (float)g.N / s->params.L;
    if (!(sqrtf(9.81f * kmax) * fabsf(t) < kFastPhaseLimit)) fast = false;
    if (launch_slab_rows(g, s->d_h0, s->d_hp, s->d_nyq, s->d_ktab, s->d_ktab_sub, base, t, fast, s->d_scratch, spick(s, stream)) < 0) return scuda(s, cudaGetLastError(), "launch_slab_rows");
    return OW_OK;
}

int ow_slab_rows(ow_slab* s, float t, int32_t transport, void* stream) { return ow_slab_rows_buf(s, t, transport, 0, stream); }

int ow_slab_cols_buf(ow_slab* s, int32_t buf, void* stream) {
    if (!s || buf < 0 || buf > 1) return OW_ERR_INVALID;
    if (!s->spectrum_ready) return sfail(s, OW_ERR_STATE, "ow_slab_cols: call ow_slab_init_spectrum_seeded first");
    if (buf == 1 && !s->d_recv1) return sfail(s, OW_ERR_STATE, "ow_slab_cols_buf: call ow_slab_enable_double_buffer first");
    OWS_CUDA(s, cudaSetDevice(s->device));
    const SlabGeom& g = s->g;
    const float js = s->casc.choppiness * ((float)g.N / (2.0f * s->casc.L));
    // with two receive buffers the column pass may overlap the next frame's row pass: it then has its own scratch
    float2* scratch = s->d_scratch_cols ? s->d_scratch_cols : s->d_scratch;
    if (launch_slab_cols(g, buf ? s->d_recv1 : s->d_recv, s->d_disp, s->d_normal, s->d_jac, js, scratch, spick(s, stream)) < 0)
        return scuda(s, cudaGetLastError(), "launch_slab_cols");
    return OW_OK;
}

int ow_slab_cols(ow_slab* s, void* stream) { return ow_slab_cols_buf(s, 0, stream); }

int ow_slab_recv_buffer(ow_slab* s, int32_t buf, void** ptr) {
    if (!s || !ptr || buf < 0 || buf > 1) return OW_ERR_INVALID;
    if (buf == 1 && !s->d_recv1) return sfail(s, OW_ERR_STATE, "ow_slab_recv_buffer: call ow_slab_enable_double_buffer first");
    *ptr = buf ? s->d_recv1 : s->d_recv;
    return OW_OK;
}

int ow_slab_set_column_lines(ow_slab* s, int32_t mode) {
    if (!s || (mode != 0 && mode != 2 && mode != 4)) return OW_ERR_INVALID;
    s->g.bigcol_pipe_grid = mode == 4 ? s->bigcol_pipe_full : mode == 2 ? -1 : 0;
    return OW_OK;
}

int ow_slab_set_post_ctas(ow_slab* s, int32_t per_sm) {
    if (!s || per_sm < 0 || per_sm > 8) return OW_ERR_INVALID;
    int sms = 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device) != cudaSuccess) { cudaGetLastError(); sms = 148; }
    s->post_ctas_per_sm = per_sm;
    s->g.post_ctas = per_sm * sms;
    return OW_OK;
}

int ow_slab_set_line_clusters(ow_slab* s, int32_t mode) {
    if (!s || mode < -1 || mode > 7) return OW_ERR_INVALID;
    s->g.big_cluster = mode < 0 ? s->cluster_caps : (mode & s->cluster_caps & 3) | ((mode & 2) ? (mode & 4) : 0);
    return OW_OK;
}

int ow_slab_get_line_clusters(const ow_slab* s) { return s ? s->g.big_cluster : 0; }

int ow_slab_sync(ow_slab* s, void* stream) {
    if (!s) return OW_ERR_INVALID;
    OWS_CUDA(s, cudaSetDevice(s->device));
    OWS_CUDA(s, cudaStreamSynchronize(spick(s, stream)));
    return OW_OK;
}

int ow_slab_download(ow_slab* s, int32_t which, void* host, size_t bytes, void* stream) {
    if (!s || !host) return OW_ERR_INVALID;
    OWS_CUDA(s, cudaSetDevice(s->device));
    const SlabGeom& g = s->g;
    cudaStream_t st = spick(s, stream);
    if (which == OW_IMG_DY || which == OW_IMG_DX || which == OW_IMG_DZ) {
        if (bytes != (size_t)g.N * g.XL * sizeof(float)) return sfail(s, OW_ERR_INVALID, "ow_slab_download: size mismatch");
        const float* src = s->d_disp + (size_t)which * g.N * g.XH + kSlabHalo;   // strip the halo columns
        OWS_CUDA(s, cudaMemcpy2DAsync(host, (size_t)g.XL * sizeof(float), src, (size_t)g.XH * sizeof(float), (size_t)g.XL * sizeof(float), g.N,
                                      cudaMemcpyDeviceToHost, st));
    } else if (which == OW_IMG_NORMAL) {
        if (bytes != (size_t)g.N * g.XL * sizeof(float4)) return sfail(s, OW_ERR_INVALID, "ow_slab_download: size mismatch");
        OWS_CUDA(s, cudaMemcpyAsync(host, s->d_normal, bytes, cudaMemcpyDeviceToHost, st));
    } else if (which == OW_IMG_JACOBIAN) {
        if (!s->d_jac) return sfail(s, OW_ERR_STATE, "ow_slab_download: created without OW_FLAG_JACOBIAN");
        if (bytes != (size_t)g.N * g.XL * sizeof(float)) return sfail(s, OW_ERR_INVALID, "ow_slab_download: size mismatch");
        OWS_CUDA(s, cudaMemcpyAsync(host, s->d_jac, bytes, cudaMemcpyDeviceToHost, st));
    } else {
        return sfail(s, OW_ERR_INVALID, "ow_slab_download: unknown image");
    }
    OWS_CUDA(s, cudaStreamSynchronize(st));
    return OW_OK;
}

// Single-process transport for world == 1 (and for tests): what the all-to-all does when there is nobody else.
int ow_slab_local_exchange(ow_slab* s, void* stream) {
    if (!s) return OW_ERR_INVALID;
    if (s->g.world != 1) return sfail(s, OW_ERR_STATE, "ow_slab_local_exchange: only for world == 1 (use an all-to-all otherwise)");
    OWS_CUDA(s, cudaSetDevice(s->device));
    OWS_CUDA(s, cudaMemcpyAsync(s->d_recv, s->d_send, block_elems(s->g) * sizeof(float2), cudaMemcpyDeviceToDevice, spick(s, stream)));
    return OW_OK;
}

}  // extern "C"
