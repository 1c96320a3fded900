// ref_driver.cpp — host side of oracle/_ref: runs the reference's OWN compute shaders (compiled from their GLSL source by
// oracle/make_ref.py + glsl_emu.hpp) in the order the reference dispatches them. TEST INFRASTRUCTURE: it pins the oracle
// (oracle/ow_oracle.cpp is a restatement; this executes the original shader text) and can serve as the CPU baseline.
//
// What is restated here is only the dispatch logic of src/main.cpp, each piece citing the lines it follows:
//   init():   generate_bit_reversed_indices :733-744 (reverse_bits :33-64), generate_twiddle_factors :711-729, tilde_h0_k :553-583
//   update(): tilde_h0_t :587-608, butterfly_fft x3 :612-683 (dy, dx, dz), generate_normal_map :687-707       (:240-244)
// plus the "GL driver" built-ins declared in glsl_emu.hpp.
#include <omp.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "glsl_emu.hpp"

// ---- built-ins ----------------------------------------------------------------------------------------------------------
namespace glsl {
thread_local uvec3 gl_GlobalInvocationID;
float sqrt(float x) { return ::sqrtf(x); }
float exp(float x) { return ::expf(x); }
float log(float x) { return ::logf(x); }
float sin(float x) { return ::sinf(x); }
float cos(float x) { return ::cosf(x); }
float pow(float x, float y) { return ::powf(x, y); }
float mod(float x, float y) { return x - y * ::floorf(x / y); }
float clamp(float x, float lo, float hi) { return ::fminf(::fmaxf(x, lo), hi); }
float dot(const vec2& a, const vec2& b) { return a.x * b.x + a.y * b.y; }
float length(const vec2& a) { return ::sqrtf(a.x * a.x + a.y * a.y); }
vec2 normalize(const vec2& a) {
    const float r = 1.0f / ::sqrtf(a.x * a.x + a.y * a.y);       // v * inversesqrt(dot(v, v)): 0 * inf = NaN at k = 0, as on a GPU
    return vec2(a.x * r, a.y * r);
}
vec3 normalize(const vec3& a) {
    const float r = 1.0f / ::sqrtf(a.x * a.x + a.y * a.y + a.z * a.z);
    return vec3(a.x * r, a.y * r, a.z * r);
}
vec4 texture(const sampler2D& s, const vec2& uv) {
    auto wrap = [&](int i, int n) { return s.repeat ? ((i % n) + n) % n : (i < 0 ? 0 : (i >= n ? n - 1 : i)); };
    if (!s.linear) {
        const int x = wrap((int)::floorf(uv.x * (float)s.w), s.w), y = wrap((int)::floorf(uv.y * (float)s.h), s.h);
        return vec4(s.data[(long)y * s.w + x], 0.0f, 0.0f, 1.0f);
    }
    const float fx = uv.x * (float)s.w - 0.5f, fy = uv.y * (float)s.h - 0.5f;
    const float x0f = ::floorf(fx), y0f = ::floorf(fy);
    const float ax = fx - x0f, ay = fy - y0f;
    const int x0 = wrap((int)x0f, s.w), x1 = wrap((int)x0f + 1, s.w), y0 = wrap((int)y0f, s.h), y1 = wrap((int)y0f + 1, s.h);
    const float t00 = s.data[(long)y0 * s.w + x0], t10 = s.data[(long)y0 * s.w + x1];
    const float t01 = s.data[(long)y1 * s.w + x0], t11 = s.data[(long)y1 * s.w + x1];
    const float v = (1.0f - ay) * ((1.0f - ax) * t00 + ax * t10) + ay * ((1.0f - ax) * t01 + ax * t11);
    return vec4(v, 0.0f, 0.0f, 1.0f);
}
}  // namespace glsl

// ---- the six shaders: globals (uniforms, images) and main() of each, in their own namespaces (generated) ---------------------
#include "_ref/shaders_decl.hpp"

using namespace glsl;

namespace {

struct Image {
    std::vector<float> px;
    image2D view;
    Image(int w, int h) : px((size_t)4 * w * h, 0.0f) { view.data = px.data(); view.w = w; view.h = h; }
};

template <class F>
void dispatch(int nx, int ny, F&& shader_main) {     // glDispatchCompute over an nx x ny grid of invocations
#pragma omp parallel for schedule(static)
    for (int y = 0; y < ny; ++y)
        for (int x = 0; x < nx; ++x) {
            gl_GlobalInvocationID.x = (unsigned)x; gl_GlobalInvocationID.y = (unsigned)y; gl_GlobalInvocationID.z = 0;
            shader_main();
        }
}

int ilog2(int n) { int l = 0; while ((1 << l) < n) ++l; return l; }      // int(log(m_N)/log(2)) for powers of two (:620, :715, :738)

struct Ref {
    int N, L, log2n;
    float wind_speed, amplitude, suppression;
    vec2 wind_dir;
    std::vector<float> noise[4];
    int nw, nh;
    std::vector<int> bitrev;
    Image twiddle, h0k, h0minusk, hkt_dx, hkt_dy, hkt_dz, pingpong, dy, dx, dz, normal;
    std::vector<float> height;     // dy as a single-channel texture for s_HeightMap
    Ref(int n, int l)
        : N(n), L(l), log2n(ilog2(n)), twiddle(ilog2(n), n), h0k(n, n), h0minusk(n, n), hkt_dx(n, n), hkt_dy(n, n), hkt_dz(n, n),
          pingpong(n, n), dy(n, n), dx(n, n), dz(n, n), normal(n, n), height((size_t)n * n) {}

    void init() {
        // generate_bit_reversed_indices, src/main.cpp:733-744: indices[i] = reverse_bits(i, log2 N)
        bitrev.resize(N);
        for (int i = 0; i < N; ++i) {
            int r = 0;
            for (int b = 0; b < log2n; ++b) r |= ((i >> b) & 1) << (log2n - 1 - b);
            bitrev[i] = r;
        }
        // generate_twiddle_factors, :711-729: dispatch (log2 N) x N
        shader_twiddle_factors::twiddle_factors = twiddle.view;
        shader_twiddle_factors::bit_reversed.j = bitrev.data();
        shader_twiddle_factors::u_N = N;
        dispatch(log2n, N, [] { shader_twiddle_factors::main(); });
        // tilde_h0_k, :553-583: wind direction normalised on the host (glm::normalize), noise images bound as samplers
        sampler2D* ns[4] = {&shader_tilde_h0_k::noise0, &shader_tilde_h0_k::noise1, &shader_tilde_h0_k::noise2, &shader_tilde_h0_k::noise3};
        for (int j = 0; j < 4; ++j) { ns[j]->data = noise[j].data(); ns[j]->w = nw; ns[j]->h = nh; ns[j]->linear = false; ns[j]->repeat = false; }
        shader_tilde_h0_k::u_Amplitude = amplitude;
        shader_tilde_h0_k::u_WindSpeed = wind_speed;
        shader_tilde_h0_k::u_WindDirection = wind_dir;
        shader_tilde_h0_k::u_SuppressFactor = suppression;
        shader_tilde_h0_k::u_N = N;
        shader_tilde_h0_k::u_L = L;
        shader_tilde_h0_k::tilde_h0k = h0k.view;
        shader_tilde_h0_k::tilde_h0minusk = h0minusk.view;
        dispatch(N, N, [] { shader_tilde_h0_k::main(); });
    }

    // butterfly_fft(src, dst), :612-683
    void butterfly_fft(Image& src, Image& dst) {
        shader_butterfly::twiddle_factors = twiddle.view;
        shader_butterfly::pingpong0 = src.view;
        shader_butterfly::pingpong1 = pingpong.view;
        int pp = 0;
        for (int dir = 0; dir < 2; ++dir)                       // horizontal (:626-641), then vertical (:648-661)
            for (int i = 0; i < log2n; ++i) {
                shader_butterfly::u_PingPong = pp;
                shader_butterfly::u_Direction = dir;
                shader_butterfly::u_Stage = i;
                dispatch(N, N, [] { shader_butterfly::main(); });
                pp = (pp + 1) % 2;
            }
        shader_inversion::u_PingPong = pp;                      // :667-680
        shader_inversion::u_N = N;
        shader_inversion::displacement = dst.view;
        shader_inversion::pingpong0 = src.view;
        shader_inversion::pingpong1 = pingpong.view;
        dispatch(N, N, [] { shader_inversion::main(); });
    }

    void frame(float t) {
        // tilde_h0_t, :587-608 (u_Time = the caller's t instead of glfwGetTime())
        shader_tilde_h0_t::u_Time = t;
        shader_tilde_h0_t::u_N = N;
        shader_tilde_h0_t::u_L = L;
        shader_tilde_h0_t::tilde_h0k = h0k.view;
        shader_tilde_h0_t::tilde_h0minusk = h0minusk.view;
        shader_tilde_h0_t::tilde_hkt_dx = hkt_dx.view;
        shader_tilde_h0_t::tilde_hkt_dy = hkt_dy.view;
        shader_tilde_h0_t::tilde_hkt_dz = hkt_dz.view;
        dispatch(N, N, [] { shader_tilde_h0_t::main(); });
        butterfly_fft(hkt_dy, dy);                              // update(), :241-243
        butterfly_fft(hkt_dx, dx);
        butterfly_fft(hkt_dz, dz);
        // generate_normal_map, :687-707: m_dy bound as s_HeightMap (single mip, LINEAR, REPEAT)
        for (size_t i = 0; i < (size_t)N * N; ++i) height[i] = dy.px[4 * i];
        shader_normal_map::s_HeightMap.data = height.data();
        shader_normal_map::s_HeightMap.w = shader_normal_map::s_HeightMap.h = N;
        shader_normal_map::s_HeightMap.linear = true;
        shader_normal_map::s_HeightMap.repeat = true;
        shader_normal_map::u_N = N;
        shader_normal_map::normal_map = normal.view;
        dispatch(N, N, [] { shader_normal_map::main(); });
    }
};

}  // namespace

extern "C" {

void* ref_create(int N, int L, float wind_speed, float wdx, float wdy, float amplitude, float suppression, const uint8_t* noise, int nw, int nh) {
    if (N < 2 || (N & (N - 1))) return nullptr;
    Ref* r = new Ref(N, L);
    r->wind_speed = wind_speed; r->amplitude = amplitude; r->suppression = suppression;
    const float inv = 1.0f / sqrtf(wdx * wdx + wdy * wdy);      // glm::normalize, src/main.cpp:555
    r->wind_dir = vec2(wdx * inv, wdy * inv);
    r->nw = nw; r->nh = nh;
    for (int j = 0; j < 4; ++j) {                               // RGBA8 .r / 255 (fw/src/ogl.cpp:255-329: no flip, no sRGB)
        r->noise[j].resize((size_t)nw * nh);
        for (size_t i = 0; i < (size_t)nw * nh; ++i) r->noise[j][i] = (float)noise[(size_t)j * nw * nh + i] / 255.0f;
    }
    r->init();
    return r;
}

void ref_destroy(void* h) { delete static_cast<Ref*>(h); }

// Threads the dispatch loops use (torchrun exports OMP_NUM_THREADS=1; the CPU baseline wants every host core).
void ref_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int ref_max_threads() { return omp_get_num_procs(); }

void ref_get_h0(void* h, float* a, float* b) {
    Ref* r = static_cast<Ref*>(h);
    for (size_t i = 0; i < (size_t)r->N * r->N; ++i) {
        a[2 * i] = r->h0k.px[4 * i]; a[2 * i + 1] = r->h0k.px[4 * i + 1];
        b[2 * i] = r->h0minusk.px[4 * i]; b[2 * i + 1] = r->h0minusk.px[4 * i + 1];
    }
}

void ref_get_twiddle(void* h, float* tw /* [N][log2N][4] */, int32_t* bitrev) {
    Ref* r = static_cast<Ref*>(h);
    std::memcpy(tw, r->twiddle.px.data(), r->twiddle.px.size() * sizeof(float));
    for (int i = 0; i < r->N; ++i) bitrev[i] = r->bitrev[i];
}

void ref_frame(void* h, float t, float* dy, float* dx, float* dz, float* normal) {
    Ref* r = static_cast<Ref*>(h);
    r->frame(t);
    const size_t nn = (size_t)r->N * r->N;
    for (size_t i = 0; i < nn; ++i) { dy[i] = r->dy.px[4 * i]; dx[i] = r->dx.px[4 * i]; dz[i] = r->dz.px[4 * i]; }
    std::memcpy(normal, r->normal.px.data(), nn * 4 * sizeof(float));
}

}  // extern "C"
