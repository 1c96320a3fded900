/* oceanwaves.h — C ABI of the B200-native Tessendorf hot path (liboceanwaves.so).
 *
 * Drop-in boundary for the per-frame ocean simulation of diharaw/fft-ocean-waves. The reference has no
 * library/FFI surface: the seam is a group of private member functions of `FFTOceanWaves`
 * (reference src/main.cpp) that communicate through member textures. Each entry point below names the
 * reference code it replaces. All functions return an ow_status (0 = OK), never throw, never abort, keep no
 * global state; one context is single-threaded (externally synchronised); contexts on different devices are
 * independent. No torch / C++ types cross this boundary: plain pointers and sizes only.
 *
 * Texel (x, y) <-> gl_GlobalInvocationID.xy; all images are row-major [y][x], first row first in memory,
 * exactly the layout of the reference's GL textures (src/main.cpp:1089-1099).
 */
#ifndef OCEANWAVES_H
#define OCEANWAVES_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OW_VERSION 200

typedef enum ow_status {
    OW_OK = 0,
    OW_ERR_INVALID = 1,   /* bad argument (N not a supported power of two, null pointer, index out of range) */
    OW_ERR_CUDA = 2,      /* a CUDA runtime call failed; see ow_last_error() */
    OW_ERR_STATE = 3,     /* call order violated (e.g. ow_step before ow_init_spectrum) */
    OW_ERR_NO_GL = 4,     /* GL interop requested but no current GL context / registration failed */
    OW_ERR_NOMEM = 5
} ow_status;

/* Simulation parameters of ONE cascade (patch). Reference: FFTOceanWaves members, src/main.cpp:1640-1646
 * (m_wind_speed 80, m_amplitude 2, m_suppression_factor 0.1, m_wind_direction (1,1), m_N 256, m_L 1000) and
 * :1632-1633 (m_choppiness). wind_dir need not be normalised (the library normalises it like
 * glm::normalize at src/main.cpp:555). L is float here; the reference's u_L is an int uniform. */
typedef struct ow_params {
    float L;            /* patch size in metres (u_L) */
    float wind_speed;   /* u_WindSpeed */
    float wind_dir[2];  /* u_WindDirection before normalisation */
    float amplitude;    /* u_Amplitude */
    float suppression;  /* u_SuppressFactor */
    float choppiness;   /* lambda; only used by the Jacobian (the consumer applies it to dx/dz, grid_tes.glsl:61-62) */
} ow_params;

#define OW_FLAG_JACOBIAN 0x1u   /* also produce the Jacobian/foam map (extension; not in the reference) */
#define OW_FLAG_EXACT_SINCOS 0x2u /* always use full-range sincosf for e^{iwt} (default: SFU sin/cos after an exact
                                    2*pi reduction whenever max|w*t| < 2e4, absolute error ~5e-7) */

#define OW_FLAG_FOUR_STEP 0x4u   /* N = 1024 / 2048 only: run the N = A*B line decomposition that N > 4096 uses (A = 4), for
                                    testing that code path against the direct kernels; slower, same results to round-off */

#define OW_FLAG_FUSED_NORMALS 0x8u /* N <= 2048: force the normal map to be produced as the epilogue of the dy column tiles (ow_col2_kernel)
                                      instead of by a separate kernel, whatever the per-N default is. Identical images */

/* Packed output set for the consumer, in ADDITION to the reference formats: per slot one `displacement`
 * image RGBA32F (OW_FLAG_PACKED_F32) or RGBA16F (OW_FLAG_PACKED_F16) = (dx, dy, dz, J), J = 1 without OW_FLAG_JACOBIAN, and one
 * `normal_xz` image RG16_SNORM = (n.x, n.z); the consumer rebuilds n.y = sqrt(1 - n.x^2 - n.z^2) (the normal is a unit vector with
 * y > 0, normal_map_cs.glsl:53). 20 / 12 B per texel instead of 28 (+4 with the Jacobian). Replaces the four sampler binds of
 * src/main.cpp:477-487 / grid_tes.glsl:60-64 by two; INTEGRATION.md shows the decode. */
#define OW_FLAG_PACKED_F32 0x10u
#define OW_FLAG_PACKED_F16 0x20u

typedef struct ow_ctx ow_ctx;

/* Device pointers to one output set ("slot"); library-owned, valid until ow_destroy. Written by ow_step*.
 * Replaces the reference's output textures m_dy/m_dx/m_dz (R32F) and m_normal_map (RGBA32F),
 * src/main.cpp:1096-1099. */
typedef struct ow_outputs {
    int32_t N;
    float* dy;         /* [N*N]   height displacement            */
    float* dx;         /* [N*N]   choppy displacement along x    */
    float* dz;         /* [N*N]   choppy displacement along z    */
    float* normal;     /* [N*N*4] unit normal xyz, w = 1         */
    float* jacobian;   /* [N*N]   NULL unless OW_FLAG_JACOBIAN   */
} ow_outputs;

/* Device pointers to one slot's packed images (OW_FLAG_PACKED_*); library-owned, written by ow_step*. */
typedef struct ow_packed {
    int32_t N;
    int32_t displacement_texel_bytes;   /* 16 (RGBA32F) or 8 (RGBA16F) */
    void* displacement;                 /* [N*N] texels (dx, dy, dz, J) */
    void* normal_xz;                    /* [N*N] texels RG16_SNORM (n.x, n.z) */
} ow_packed;

typedef enum ow_image {
    OW_IMG_DY = 0, OW_IMG_DX = 1, OW_IMG_DZ = 2, OW_IMG_NORMAL = 3, OW_IMG_JACOBIAN = 4,
    OW_IMG_H0K = 5,        /* [N*N*2] tilde_h0k        (index = cascade, not slot) */
    OW_IMG_H0MINUSK = 6    /* [N*N*2] tilde_h0minusk   (index = cascade, not slot) */
} ow_image;

/* ---- lifecycle: replaces create_textures() (src/main.cpp:1083-1145) for the sim resources ------------- */

/* n_cascades independent patches of size N x N (N a power of two in [128, 32768]; N > 4096 uses the N = A*B line decomposition); n_slots >= n_cascades
 * output sets (extra slots let one cascade be evaluated at several times per launch, see ow_step_multi).
 * device = CUDA ordinal. */
int ow_create(int32_t N, int32_t n_cascades, int32_t n_slots, const ow_params* cascades, int32_t device,
              uint32_t flags, ow_ctx** out);
void ow_destroy(ow_ctx* ctx);
/* Last error text of this context (or of the calling thread's last failed ow_create when ctx == NULL). */
const char* ow_last_error(const ow_ctx* ctx);

/* ---- init: replaces tilde_h0_k(), generate_bit_reversed_indices(), generate_twiddle_factors()
 *      (src/main.cpp:218-220, 553-583, 711-744) ------------------------------------------------------- */

/* Change one cascade's parameters; takes effect at the next ow_init_spectrum (the reference never
 * re-runs tilde_h0_k after its GUI edits, SURVEY.md quirk 8 — this API makes it explicit). */
int ow_set_params(ow_ctx* ctx, int32_t cascade, const ow_params* p);
/* Uniform noise for the Box-Muller draw: 4 planes of w*h bytes (the R channel of the reference's
 * data/noise/LDR_LLL1_{0..3}.png, w = h = 256), HOST pointers. cascade = -1 sets every cascade.
 * Lookup rule is the shader's NEAREST/CLAMP fetch at gid/N (tilde_h0_k_cs.glsl:53-58). */
int ow_set_noise(ow_ctx* ctx, int32_t cascade, const uint8_t* const planes[4], int32_t w, int32_t h);
/* Counter-based noise for grids with no noise image to ship (BASELINE config C5): the four N x N byte planes are
 * generated ON THE DEVICE with Philox4x32-10 (counter (ix, iy, 0, 0), key = seed; plane j = low byte of output word j)
 * and looked up 1:1; everything downstream (Box-Muller, Phillips) is the same tilde_h0_k_cs.glsl path. The slab
 * context (ow_slab_init_spectrum_seeded) draws from the same function, so both paths see identical h0. */
int ow_set_noise_seed(ow_ctx* ctx, int32_t cascade, uint64_t seed);
/* = tilde_h0_k_cs.glsl for every cascade. Re-callable. Synchronous (like the reference's glFinish). */
int ow_init_spectrum(ow_ctx* ctx);
/* The same for ONE cascade: what a GUI edit of one patch's wind/amplitude needs (ow_set_params marks only that cascade
 * stale; ow_step* refuses to evaluate a stale cascade with OW_ERR_STATE). */
int ow_init_spectrum_cascade(ow_ctx* ctx, int32_t cascade);
/* Overwrite a cascade's initial spectrum (HOST pointers, N*N*2 floats each); makes THAT cascade current, no other. */
int ow_set_h0(ow_ctx* ctx, int32_t cascade, const float* h0k, const float* h0minusk);

/* ---- per frame: replaces tilde_h0_t(); butterfly_fft() x3; generate_normal_map()
 *      (src/main.cpp:240-244, 587-707) ---------------------------------------------------------------- */

/* Slot i <- cascade i at time t, for every cascade. Asynchronous on `stream` (a cudaStream_t passed as
 * void*; NULL = the context's own stream). t replaces float(glfwGetTime()) (src/main.cpp:599).
 * For N <= 4096 the whole frame is ONE cudaGraphLaunch (built on first use; only the time is patched per call) instead of
 * the reference's 53 dispatches; ow_set_graph(ctx, 0) falls back to plain launches. */
int ow_step(ow_ctx* ctx, float t, void* stream);
int ow_set_graph(ow_ctx* ctx, int32_t enabled);
/* Slot i <- cascade cascade_of_slot[i] at time time_of_slot[i], i < count <= n_slots (HOST arrays). */
int ow_step_multi(ow_ctx* ctx, int32_t count, const int32_t* cascade_of_slot, const float* time_of_slot, void* stream);
/* Same as ow_step_multi but synchronous, with CUDA events around each kernel on `stream`: kernel_ms[0..2] receive
 * the summed durations (ms) of the row-IFFT, column-IFFT and normal/Jacobian kernels. For bench.py's roofline. */
int ow_step_multi_timed(ow_ctx* ctx, int32_t count, const int32_t* cascade_of_slot, const float* time_of_slot, void* stream,
                        float* kernel_ms);
/* Block until everything queued by this context on `stream` has finished. */
int ow_sync(ow_ctx* ctx, void* stream);

int ow_get_outputs(ow_ctx* ctx, int32_t slot, ow_outputs* out);
/* Copy one image to HOST memory (synchronous w.r.t. `stream`). bytes must equal the image size. */
int ow_download(ow_ctx* ctx, int32_t index, int32_t which /* ow_image */, void* host, size_t bytes, void* stream);
/* Asynchronous variant of ow_download of a slot's dy,dx,dz,normal[,jacobian] into ONE pinned host block
 * laid out in that order; used by the headless end-to-end path. */
int ow_download_frame_async(ow_ctx* ctx, int32_t slot, void* pinned_host, size_t bytes, void* stream);
size_t ow_frame_bytes(const ow_ctx* ctx);
/* Packed set (OW_FLAG_PACKED_*): device pointers, bytes per slot (displacement image followed by normal_xz), and the
 * asynchronous copy of one slot's pair of images into ONE pinned host block of that size. */
int ow_get_packed(ow_ctx* ctx, int32_t slot, ow_packed* out);
size_t ow_packed_bytes(const ow_ctx* ctx);
int ow_download_packed_async(ow_ctx* ctx, int32_t slot, void* pinned_host, size_t bytes, void* stream);

/* Tuning: upper bound on the slots that share one row/column/normal launch group. 0 = automatic (the group's
 * 12 B/texel intermediate <= ~100 MB, slots split evenly into at least as many groups as there are streams). */
int ow_set_group_size(ow_ctx* ctx, int32_t slots_per_group);
/* Tuning: independent launch groups of one ow_step_multi call are spread over n internal streams (forked from and
 * joined back into the caller's stream), so one group's tail overlaps the next group's head. 1 = strictly serial.
 * Default 3; n in [1, 4]. Results do not depend on it. */
int ow_set_streams(ow_ctx* ctx, int32_t n);
/* Number of kernels the last ow_step/ow_step_multi launched (for launch accounting in bench.py), and the number of launch
 * groups they formed (a group = the row, column[, normal] kernels of the slots that share launches). */
int ow_last_launch_count(const ow_ctx* ctx);
int ow_last_group_count(const ow_ctx* ctx);
/* Tuning / A-B runs (per context; results agree to fp32 round-off): which row kernel runs (0 = the per-N default, 1 = one CTA
 * per row-pair group, 2 = persistent, next row prefetched into registers, 3 = persistent, next row staged by cp.async.bulk behind an
 * mbarrier); which column kernel (0 = per-N default, 1 = ow_col_kernel, 2 = ow_col2_kernel with direct loads, 3 = ow_col2_kernel with
 * TMA-staged tiles, 4 = ow_col_pipe_kernel: persistent, next tile's first load batch in flight in registers; on a line-decomposition
 * grid, N > 4096, modes 2 and 4 select the column LINES kernel's 8-column-tile and persistent shapes instead) and whether the normal map is its epilogue (fused: -1 = per-N default, 0 = separate normal kernel, 1 = fused;
 * needs mode 2 or 3 and N <= 2048); and whether the column kernel drops the consumed intermediate from L2 without writing it back
 * (discard.global.L2). */
int ow_set_row_kernel(ow_ctx* ctx, int32_t mode);
int ow_set_column_kernel(ow_ctx* ctx, int32_t mode, int32_t fused);
/* The kernels this context will actually run (per-N defaults resolved): row 1..3, column 1..3, fused 0/1. */
int ow_get_kernel_modes(ow_ctx* ctx, int32_t* row, int32_t* column, int32_t* fused);
/* Tuning: resident CTAs per SM of the PERSISTENT row (modes 2, 3) and column (modes 2, 3) kernels; 0 = as many as fit. Smaller grids
 * leave room on every SM for the other kernels of frames in flight on the context's other streams. */
int ow_set_resident_ctas(ow_ctx* ctx, int32_t row_per_sm, int32_t col_per_sm);
/* L2 residency of the folded initial spectrum (what every frame of a cascade re-reads): 1 = the context's launch streams carry an
 * access-policy window over it (hits persist in the device's L2 carve-out, cudaLimitPersistingL2CacheSize - a per-device limit this
 * call sets), 0 = off (default: measured slower on B200), -1 = on when all cascades' blocks fit two thirds of the carve-out. A hint only. */
int ow_set_l2_persist(ow_ctx* ctx, int32_t mode);
/* Launches of ONE frame (ow_step of a single cascade, the reference's update()) use latency-oriented kernel shapes for N <= 1024: the same
 * butterflies dealt out over more threads and CTAs, bit-identical images. 1 = on (default), 0 = the throughput shapes everywhere. */
int ow_set_latency_shapes(ow_ctx* ctx, int32_t on);
/* How ow_step_multi runs a launch group: 0 = three kernels (row, column, normal; default), 1 = ONE persistent kernel that walks row, column and
 * normal-map work items of consecutive frames behind per-frame dependency counters, so that only two or three frames are in flight and the
 * intermediate and the displacement planes are consumed out of L2 (N = 256, 512, 1024; ow_step's CUDA graph keeps the three kernels). */
int ow_set_frame_kernel(ow_ctx* ctx, int32_t mode);
/* Lines longer than one CTA's shared memory (N > 4096, or OW_FLAG_FOUR_STEP): N = A*B, the A sub-lines of a line are transformed by the
 * A CTAs of a thread-block cluster and combined through distributed shared memory (no global scratch). mode: -1 = wherever the device can
 * co-schedule the cluster, 0 = never (default: two kernels per direction through a global scratch array - measured 3.5x FASTER on B200
 * at N = 32768, where the exchange through distributed shared memory is bound by its ~20 B/clk per SM), else a bit mask:
 * 1 = rows, 2 = columns, 4 = columns on 8-column tiles (3 CTAs per SM) instead of 16-column ones. ow_get_line_clusters returns the mask in use. */
int ow_set_line_clusters(ow_ctx* ctx, int32_t mode);
int ow_get_line_clusters(ow_ctx* ctx);
int ow_set_discard_intermediate(ow_ctx* ctx, int32_t on);

/* ---- multi-cascade composition and the demo's clock ----------------------------------
 * The reference renders ONE cascade: grid_tes.glsl:60-64 samples the sim textures (LINEAR/REPEAT: src/main.cpp:1142-1144 for the normal map, the texture class defaults for m_dy/m_dx/m_dz) and
 * displaces the vertex: pos.y += dy*u_DisplacementScale, pos.x -= dx*u_Choppiness, pos.z -= dz*u_Choppiness, normal = texture(s_NormalMap).
 * With several cascades the consumer sums these terms over the cascades, cascade c sampled at uv = (x, z)/L_c and scaled by a
 * blending weight w_c (LOD / distance fades). ow_sample_points / ow_compose_grid evaluate that sum on the device:
 *   offset = (-sum w_c*choppiness_c*dx_c, displacement_scale * sum w_c*dy_c, -sum w_c*choppiness_c*dz_c, sum w_c)
 *   normal = normalize(sum w_c*n_c.x/n_c.y, 1, sum w_c*n_c.z/n_c.y), w = 1          (slopes add)
 * A term names an output SLOT; its cascade (L, choppiness) is the one the last ow_step* put there. Up to 16 terms. Asynchronous on
 * `stream`, ordered after the steps submitted to the same stream. */
typedef struct ow_blend_term {
    int32_t slot;
    float weight;
} ow_blend_term;
/* xz: DEVICE pointer, n_points x (x, z) world positions in metres; out: DEVICE pointer, n_points x 8 floats (offset.xyzw, normal.xyzw). */
int ow_sample_points(ow_ctx* ctx, int32_t n_terms, const ow_blend_term* terms, float displacement_scale, int32_t n_points,
                     const float* xz, float* out, void* stream);
/* The same with HOST pointers (copies in and out on `stream`, then synchronises it): buoyancy / picking queries of a few points. */
int ow_sample_points_host(ow_ctx* ctx, int32_t n_terms, const ow_blend_term* terms, float displacement_scale, int32_t n_points,
                          const float* xz, float* out, void* stream);
/* M x M world positions (origin_x + (i+0.5)*extent/M, origin_z + (j+0.5)*extent/M), row j / column i of two DEVICE images of M*M float4:
 * the combined offset and normal of one clip-map level. */
int ow_compose_grid(ow_ctx* ctx, int32_t n_terms, const ow_blend_term* terms, float displacement_scale, int32_t M, float origin_x,
                    float origin_z, float extent, float* out_offset, float* out_normal, void* stream);
/* The demo's clock: t = float(glfwGetTime()) (src/main.cpp:599). ow_step_wall_clock(ctx, wall) = ow_step(ctx, offset + scale*wall):
 * scale 0 pauses, negative runs backwards; offset re-bases (e.g. after a pause). Defaults 1, 0. */
int ow_set_time_scale(ow_ctx* ctx, float scale, float offset);
int ow_step_wall_clock(ow_ctx* ctx, double wall_seconds, void* stream);

/* ---- CUDA-GL interop: replaces the renderer's texture binds (src/main.cpp:477-487) ------------------- */

/* Register the caller's GL textures (R32F dy,dx,dz; RGBA32F normal) via cudaGraphicsGLRegisterImage. The
 * caller's GL context must be current on the calling thread. Returns OW_ERR_NO_GL when that fails. */
int ow_gl_register(ow_ctx* ctx, uint32_t tex_dy, uint32_t tex_dx, uint32_t tex_dz, uint32_t tex_normal);
/* ow_step for slot 0 and copy the results into the registered textures (map -> copy -> unmap). */
int ow_gl_step(ow_ctx* ctx, float t);
int ow_gl_unregister(ow_ctx* ctx);
/* The same for the packed set: the caller's RGBA32F/RGBA16F displacement texture and RG16_SNORM normal texture. */
int ow_gl_register_packed(ow_ctx* ctx, uint32_t tex_displacement, uint32_t tex_normal_xz);

/* ---- one grid over several GPUs (BASELINE config C5): slab-decomposed 2-D IFFT --------------------------
 * Replaces the same reference functions as ow_step (tilde_h0_t; butterfly_fft x3; generate_normal_map,
 * src/main.cpp:240-244) for a grid too large or too slow for one GPU. One process per GPU, one ow_slab per process.
 * Rank r owns row pairs [r*PL, (r+1)*PL) (rows p and N-p, PL = N/2/world) for the row pass and columns
 * [r*XL, (r+1)*XL) (XL = N/world) for the column pass; the transpose between them is the row kernel's store pattern
 * (see ow_slab_rows). Outputs stay column-slabbed. N is a power of two in [256, 32768] with N/world >= 128; above 4096 the
 * lines use the N = A*B decomposition. */
typedef struct ow_slab ow_slab;

#define OW_SLAB_IPC_HANDLE_BYTES 64
enum { OW_SLAB_SEND_BUFFER = 0, OW_SLAB_PEER_STORES = 1 };

typedef struct ow_slab_info {
    int32_t N, world, rank;
    int32_t pairs_per_rank;   /* PL */
    int32_t cols_per_rank;    /* XL */
    int32_t padded_cols;      /* XH = XL + 2*halo: row length of recv / dy / dx / dz */
    int32_t halo;
    size_t block_bytes;       /* bytes one rank sends to one rank per frame: PL*3*XH*8; send and recv hold `world` blocks */
    void* send;               /* device: [world][PL][3][XH] float2, filled by ow_slab_rows(OW_SLAB_SEND_BUFFER) */
    void* recv;               /* device: [world][PL][3][XH] float2 = [N/2][3][XH], read by ow_slab_cols */
    float *dy, *dx, *dz;      /* device: [N][XH] each; this rank's columns start at index `halo` of every row */
    float* normal;            /* device: [N][XL][4] */
    float* jacobian;          /* device: [N][XL] or NULL */
} ow_slab_info;

int ow_slab_create(int32_t N, int32_t world, int32_t rank, const ow_params* p, int32_t device, uint32_t flags, ow_slab** out);
void ow_slab_destroy(ow_slab* s);
const char* ow_slab_last_error(const ow_slab* s);
int ow_slab_get_info(const ow_slab* s, ow_slab_info* info);
/* tilde_h0_k for the rows this rank owns, noise from Philox (see ow_set_noise_seed). */
int ow_slab_init_spectrum_seeded(ow_slab* s, uint64_t seed);
/* CUDA IPC handle of this rank's receive buffer (OW_SLAB_IPC_HANDLE_BYTES bytes); the host all-gathers them once. */
int ow_slab_ipc_handle(ow_slab* s, void* handle, size_t bytes);
/* handles: world x OW_SLAB_IPC_HANDLE_BYTES, rank order. Maps every peer's receive buffer into this process. */
int ow_slab_open_peers(ow_slab* s, const void* handles, size_t bytes);
/* Spectrum at time t + row IFFT of this rank's row pairs. transport OW_SLAB_PEER_STORES: results are stored straight
 * into the column owners' receive buffers over NVLink (the caller must order every rank's ow_slab_rows before any
 * rank's ow_slab_cols, and the previous frame's ow_slab_cols before the next ow_slab_rows: two barriers per frame).
 * OW_SLAB_SEND_BUFFER: results go to info.send; the caller then runs one equal-split all-to-all send -> recv
 * (block_bytes per pair of ranks), e.g. ncclSend/ncclRecv grouped or torch.distributed.all_to_all_single. */
int ow_slab_rows(ow_slab* s, float t, int32_t transport, void* stream);
/* Column IFFT + inversion + normals (+ Jacobian) on the receive buffer. */
int ow_slab_cols(ow_slab* s, void* stream);
/* world == 1 stand-in for the all-to-all (send -> recv device copy). */
int ow_slab_local_exchange(ow_slab* s, void* stream);
int ow_slab_sync(ow_slab* s, void* stream);
/* Frame pipelining over NVLink: a second receive buffer (and, above N = 4096, a second scratch), so that frame f+1's row pass - whose
 * stores ARE the exchange - runs on one stream while frame f's column pass reads the other buffer on another. The *_buf entry points
 * take the buffer index (0 or 1); ow_slab_rows / ow_slab_cols / ow_slab_ipc_handle / ow_slab_open_peers are buffer 0. The caller orders
 * every rank's rows(f) before any rank's cols(f), and every rank's cols(f) before any rank's rows(f+2). */
int ow_slab_enable_double_buffer(ow_slab* s);
int ow_slab_ipc_handle_buf(ow_slab* s, int32_t buf, void* handle, size_t bytes);
int ow_slab_open_peers_buf(ow_slab* s, int32_t buf, const void* handles, size_t bytes);
int ow_slab_rows_buf(ow_slab* s, float t, int32_t transport, int32_t buf, void* stream);
int ow_slab_cols_buf(ow_slab* s, int32_t buf, void* stream);
/* Device pointer of receive buffer `buf` (the all-to-all's destination; info.recv is buffer 0). */
int ow_slab_recv_buffer(ow_slab* s, int32_t buf, void** ptr);
/* N > 4096: CTAs per SM of the row pass's store kernel (0 = one CTA per work item; ow_slab_enable_double_buffer sets 2 when world > 1, so that the
 * NVLink-bound stores leave the SMs to the column pass running beside them). */
int ow_slab_set_post_ctas(ow_slab* s, int32_t per_sm);
/* Tuning (N > 4096): shape of the column pass's lines kernel. 0 = one 512-thread CTA per (channel, 16-column tile, sub-line) item (default),
 * 2 = 8-column tiles (three 256-thread CTAs per SM), 4 = persistent CTAs with the next batch's row loads in flight in registers. Same results.
 * For a single-GPU context the same shapes are ow_set_column_kernel modes 0/1, 2 and 4. */
int ow_slab_set_column_lines(ow_slab* s, int32_t mode);
/* As ow_set_line_clusters / ow_get_line_clusters, for a slab rank. */
int ow_slab_set_line_clusters(ow_slab* s, int32_t mode);
int ow_slab_get_line_clusters(const ow_slab* s);
/* which = OW_IMG_DY/DX/DZ ([N][XL] floats, halo stripped), OW_IMG_NORMAL ([N][XL][4]), OW_IMG_JACOBIAN ([N][XL]). */
int ow_slab_download(ow_slab* s, int32_t which, void* host, size_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OCEANWAVES_H */
