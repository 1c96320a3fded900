// oceanwaves.hpp — header-only C++ host wrapper over the C ABI (oceanwaves.h).
//
// The reference is compiled C++ (src/main.cpp, class FFTOceanWaves), so the host side above the C ABI is C++
// too. `ow::OceanSim` keeps the names and meanings of the reference's sim members and private methods, so
// that replacing the GL dispatch chain inside FFTOceanWaves is a mechanical edit:
//
//   reference (src/main.cpp)                                   here
//   m_N, m_L, m_wind_speed, m_wind_direction, m_amplitude,     OceanSim::Params (same defaults, :1632-1646)
//     m_suppression_factor, m_choppiness
//   create_textures()            :1083-1145                    OceanSim(N, params)        -> ow_create
//   tilde_h0_k()                 :553-583                      tilde_h0_k(noise planes)   -> ow_set_noise + ow_init_spectrum
//   generate_bit_reversed_indices(), generate_twiddle_factors() :711-744   no-ops (twiddles live in registers)
//   tilde_h0_t(); butterfly_fft() x3; generate_normal_map()    :240-244   update(t)       -> ow_step
//   m_dy / m_dx / m_dz / m_normal_map                          dy()/dx()/dz()/normal_map() device pointers,
//                                                              or gl_register()+gl_update(t) into the app's textures
//
// Error behaviour mirrors the reference: nothing throws; methods return false and `last_error()` holds the
// text the reference would have sent to DW_LOG_ERROR (init() returns false on failure, src/main.cpp:199-212).
#pragma once
#include <cstdint>
#include <string>

#include "oceanwaves.h"

namespace ow {

class OceanSim {
public:
    struct Params {
        float L = 1000.0f;                  // m_L            (src/main.cpp:1646)
        float wind_speed = 80.0f;           // m_wind_speed   (:1641)
        float wind_direction[2] = {1, 1};   // m_wind_direction (:1644), normalised by the library like :555
        float amplitude = 2.0f;             // m_amplitude    (:1642)
        float suppression_factor = 0.1f;    // m_suppression_factor (:1643)
        float choppiness = 0.75f;           // m_choppiness   (:1633)
    };

    OceanSim() = default;
    OceanSim(const OceanSim&) = delete;
    OceanSim& operator=(const OceanSim&) = delete;
    ~OceanSim() { destroy(); }

    // = create_textures() for the sim resources. with_jacobian adds the foam map (extension); extra_flags e.g. OW_FLAG_PACKED_F16.
    bool create(int N, const Params& p, int device = 0, bool with_jacobian = false, uint32_t extra_flags = 0u) {
        destroy();
        n_ = N;
        const ow_params c = to_c(p);
        const int rc = ow_create(N, 1, 1, &c, device, (with_jacobian ? OW_FLAG_JACOBIAN : 0u) | extra_flags, &ctx_);
        if (rc != OW_OK) { err_ = ow_last_error(nullptr); ctx_ = nullptr; return false; }
        return true;
    }
    void destroy() {
        if (ctx_) ow_destroy(ctx_);
        ctx_ = nullptr;
    }
    bool valid() const { return ctx_ != nullptr; }
    int N() const { return n_; }
    const std::string& last_error() const { return err_; }

    // = tilde_h0_k(): noise = R channel of noise/LDR_LLL1_{0..3}.png as decoded by stb_image (w*h bytes each).
    bool tilde_h0_k(const uint8_t* const noise[4], int w = 256, int h = 256) {
        return ok(ow_set_noise(ctx_, -1, noise, w, h)) && ok(ow_init_spectrum(ctx_));
    }
    // Same, with device-generated Philox noise instead of the PNGs (grids with no noise image to ship: config C5).
    bool tilde_h0_k_seeded(uint64_t seed) { return ok(ow_set_noise_seed(ctx_, -1, seed)) && ok(ow_init_spectrum(ctx_)); }
    // The reference never re-runs tilde_h0_k() after a GUI edit; this makes the re-generation explicit.
    bool set_params(const Params& p) {
        const ow_params c = to_c(p);
        return ok(ow_set_params(ctx_, 0, &c)) && ok(ow_init_spectrum(ctx_));
    }
    // Kept so a patched init() reads like the original; both tables are gone (ow_fft.cuh derives them in registers).
    void generate_bit_reversed_indices() {}
    void generate_twiddle_factors() {}

    // = tilde_h0_t(); butterfly_fft(dy); butterfly_fft(dx); butterfly_fft(dz); generate_normal_map();
    // t replaces float(glfwGetTime()) (:599). Asynchronous on `stream` (cudaStream_t as void*, NULL = own stream).
    bool update(float t, void* stream = nullptr) { return ok(ow_step(ctx_, t, stream)); }
    // The demo's clock: t = float(glfwGetTime()) (:599). time_scale / time_offset let the app slow, speed up, pause (scale 0) or
    // scrub the sea without touching the sim (the reference has no such control; a GUI slider next to the ones at :348-358 would set them).
    float time_scale = 1.0f, time_offset = 0.0f;
    bool update_wall_clock(double wall_seconds, void* stream = nullptr) {
        return ok(ow_set_time_scale(ctx_, time_scale, time_offset)) && ok(ow_step_wall_clock(ctx_, wall_seconds, stream));
    }
    // What grid_tes.glsl:60-64 does to a vertex, evaluated on the device at world positions (x, z) in metres: out[i] = 8 floats
    // (offset.xyz, weight sum, normal.xyz, 1). HOST pointers; buoyancy / picking queries. displacement_scale = m_displacement_scale (:1644).
    bool sample_points(int n, const float* xz, float* out, float displacement_scale = 0.5f, void* stream = nullptr) {
        const ow_blend_term term{0, 1.0f};
        return ok(ow_sample_points_host(ctx_, 1, &term, displacement_scale, n, xz, out, stream));
    }
    bool sync(void* stream = nullptr) { return ok(ow_sync(ctx_, stream)); }

    // Device pointers (row-major [y][x], the layout of the reference's R32F / RGBA32F textures).
    const float* dy() { return outs().dy; }
    const float* dx() { return outs().dx; }
    const float* dz() { return outs().dz; }
    const float* normal_map() { return outs().normal; }
    const float* jacobian() { return outs().jacobian; }

    // CUDA-GL interop: ids of the app's own m_dy/m_dx/m_dz (R32F) and m_normal_map (RGBA32F) textures.
    bool gl_register(uint32_t dy, uint32_t dx, uint32_t dz, uint32_t normal) { return ok(ow_gl_register(ctx_, dy, dx, dz, normal)); }
    bool gl_update(float t) { return ok(ow_gl_step(ctx_, t)); }
    // Packed set (context created with OW_FLAG_PACKED_F16 / _F32): the app's RGBA16F/RGBA32F displacement and RG16_SNORM normal textures.
    bool gl_register_packed(uint32_t displacement, uint32_t normal_xz) { return ok(ow_gl_register_packed(ctx_, displacement, normal_xz)); }
    ow_packed packed() {
        ow_packed p{};
        if (ctx_) ow_get_packed(ctx_, 0, &p);
        return p;
    }

    // Headless dump (tests, tools): which = OW_IMG_*.
    bool download(int which, void* host, size_t bytes) { return ok(ow_download(ctx_, 0, which, host, bytes, nullptr)); }

    ow_ctx* handle() { return ctx_; }

private:
    static ow_params to_c(const Params& p) {
        ow_params c;
        c.L = p.L; c.wind_speed = p.wind_speed; c.wind_dir[0] = p.wind_direction[0]; c.wind_dir[1] = p.wind_direction[1];
        c.amplitude = p.amplitude; c.suppression = p.suppression_factor; c.choppiness = p.choppiness;
        return c;
    }
    bool ok(int rc) {
        if (rc == OW_OK) return true;
        err_ = ctx_ ? ow_last_error(ctx_) : "OceanSim: not created";
        return false;
    }
    ow_outputs outs() {
        ow_outputs o{};
        if (ctx_) ow_get_outputs(ctx_, 0, &o);
        return o;
    }
    ow_ctx* ctx_ = nullptr;
    int n_ = 0;
    std::string err_;
};

}  // namespace ow
