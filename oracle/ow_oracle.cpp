// ow_oracle.cpp — CPU ORACLE (test infrastructure, NOT the product).
//
// A scalar C++ restatement of the reference's per-frame Tessendorf pipeline: the six GLSL
// compute shaders under /root/reference/src/shader and the host loop in
// /root/reference/src/main.cpp that drives them. It executes the reference's LITERAL dispatch
// chain (one full-grid pass per butterfly stage, RGBA32F ping-pong planes, the same
// twiddle/index table) so it doubles as the "reference CPU implementation" timed by bench.py.
//
// Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may
// load this file's library. The product path (fft-ocean-waves_b200/csrc) never does.
//
// PARITY STATUS: pinned against the reference's own shaders. The reference ships no tests, golden vectors or KATs for
// this path and its executable needs an OpenGL 4.5 driver (none in the image), but its six compute
// shaders compile for the CPU: oracle/make_ref.py reads their GLSL text from the reference tree and builds oracle/_ref/
// (glsl_emu.hpp = the GLSL types/built-ins, ref_driver.cpp = the dispatch order of src/main.cpp). This restatement reproduces
// that build to the last bit for the butterfly/index table and the DC texel, to 1 ulp on the initial spectrum and to 1e-7 of
// peak on whole frames (tests/test_ref_pin.py). What stays a restatement is the host dispatch loop (src/main.cpp cannot be
// compiled without GL) and the precision of GL built-ins, which the GL spec leaves to the driver. Further pins: (a) an
// independent fp64 numpy formulation (oracle/numpy_ref.py), (b) analytic known-answer tests derived from the shaders alone,
// (c) the survey-time spot values.
//
// Build: see oracle/Makefile  (g++ -O2 -ffp-contract=off -fopenmp -shared -fPIC).
// GL built-ins are restated as IEEE fp32 libm calls; pow(x,2.0) is x*x; clamp() is
// fminf(fmaxf(x,lo),hi) (IEEE maxNum semantics, which turns the k=0 NaN into -4000 exactly
// like NVIDIA's GL compiler does).

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// "#define M_PI 3.1415926535897932384626433832795" in every *_cs.glsl, line 2; GLSL literals are fp32.
constexpr float kPi = 3.1415926535897932384626433832795f;
constexpr float kG  = 9.81f;  // tilde_h0_k_cs.glsl:27, tilde_h0_t_cs.glsl:58

struct vec4 { float x, y, z, w; };
struct cpx { float real, im; };

// tilde_h0_t_cs.glsl:22-28 / butterfly_cs.glsl:34-40
inline cpx mul(cpx c0, cpx c1) {
    cpx c;
    c.real = c0.real * c1.real - c0.im * c1.im;
    c.im   = c0.real * c1.im + c0.im * c1.real;
    return c;
}
// tilde_h0_t_cs.glsl:32-38 / butterfly_cs.glsl:44-50
inline cpx add(cpx c0, cpx c1) { return cpx{c0.real + c1.real, c0.im + c1.im}; }
// tilde_h0_t_cs.glsl:42-48 — the shader builds the conjugate and then returns its ARGUMENT.
inline cpx conjugate_as_shipped(cpx c) { return c; }

inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

inline int ilog2(int n) { int l = 0; while ((1 << l) < n) ++l; return l; }  // main.cpp:620 int(log(N)/log(2))

// main.cpp:29-64 reverse_bits() restated as a plain bit loop (same permutation for 1..24 bits).
inline uint32_t reverse_bits(uint32_t v, int bits) {
    uint32_t r = 0;
    for (int i = 0; i < bits; ++i) { r = (r << 1) | (v & 1u); v >>= 1; }
    return r;
}

struct Sim {
    int N = 0;
    int log2N = 0;
    float L = 0, wind_speed = 0, wind_dir[2] = {0, 0}, amplitude = 0, suppression = 0;
    int noise_w = 0, noise_h = 0;
    std::vector<uint8_t> noise[4];
    std::vector<int32_t> bit_reversed;          // main.cpp:733-744
    std::vector<vec4> twiddle;                  // [y*log2N + stage]  (image is log2N wide, N tall; main.cpp:1094)
    std::vector<float> h0k, h0minusk;           // RG32F, 2 floats per texel (main.cpp:1089-1090)
    std::vector<vec4> hkt[3];                   // dy, dx, dz RGBA32F (main.cpp:1091-1093), order matches ow outputs
    std::vector<vec4> pingpong;                 // main.cpp:1095
    std::vector<float> disp[3];                 // dy, dx, dz R32F (main.cpp:1096-1098)
    std::vector<vec4> normal;                   // RGBA32F (main.cpp:1099)
    std::vector<float> jacobian;                // extension (not in the reference)
    int threads = 1;
};

// main.cpp:733-744
void generate_bit_reversed_indices(Sim& s) {
    s.bit_reversed.resize(s.N);
    for (int i = 0; i < s.N; ++i) s.bit_reversed[i] = (int32_t)reverse_bits((uint32_t)i, s.log2N);
}

// main.cpp:711-729 + twiddle_factors_cs.glsl:34-69.  Dispatch is (log2N, N/32, 1) x (1,32,1): x = stage, y = row.
void generate_twiddle_factors(Sim& s) {
    const int N = s.N;
    s.twiddle.resize((size_t)N * s.log2N);
    for (int stage = 0; stage < s.log2N; ++stage) {
        for (int yy = 0; yy < N; ++yy) {
            const float xx = (float)stage, xy = (float)yy;
            const float p2s1 = powf(2.0f, xx + 1.0f);  // pow(2, x.x + 1)
            const float p2s  = powf(2.0f, xx);
            const float k    = fmodf(xy * ((float)N / p2s1), (float)N);                  // :37 (operands >= 0: mod == fmod)
            const cpx tw{cosf(2.0f * kPi * k / (float)N), sinf(2.0f * kPi * k / (float)N)};  // :38
            const int span = (int)p2s;                                                  // :40
            const bool top = fmodf(xy, p2s1) < p2s;                                     // :44-47
            vec4 out;
            out.x = tw.real; out.y = tw.im;
            if (stage == 0) {                                                           // :50-58
                if (top) { out.z = (float)s.bit_reversed[yy];     out.w = (float)s.bit_reversed[yy + 1]; }
                else     { out.z = (float)s.bit_reversed[yy - 1]; out.w = (float)s.bit_reversed[yy]; }
            } else {                                                                    // :60-68
                if (top) { out.z = xy;               out.w = xy + (float)span; }
                else     { out.z = xy - (float)span; out.w = xy; }
            }
            s.twiddle[(size_t)yy * s.log2N + stage] = out;
        }
    }
}

// texture(noiseJ, texCoord).r with NEAREST + CLAMP_TO_EDGE on a W x H RGBA8 image
// (main.cpp:1102-1116; tilde_h0_k_cs.glsl:53-58). texCoord = vec2(gid)/float(N).
inline float noise_fetch(const Sim& s, int j, int ix, int iy) {
    const float u = (float)ix / (float)s.N, v = (float)iy / (float)s.N;
    int tx = (int)floorf(u * (float)s.noise_w), ty = (int)floorf(v * (float)s.noise_h);
    if (tx > s.noise_w - 1) tx = s.noise_w - 1;
    if (ty > s.noise_h - 1) ty = s.noise_h - 1;
    return (float)s.noise[j][(size_t)ty * s.noise_w + tx] / 255.0f;  // UNORM8 -> float
}

// main.cpp:553-583 + tilde_h0_k_cs.glsl:35-94
void tilde_h0_k(Sim& s) {
    const int N = s.N;
    // main.cpp:555  m_wind_direction = glm::normalize(m_wind_direction)  == v * inversesqrt(dot(v,v))
    const float inv = 1.0f / sqrtf(s.wind_dir[0] * s.wind_dir[0] + s.wind_dir[1] * s.wind_dir[1]);
    const float wdx = s.wind_dir[0] * inv, wdy = s.wind_dir[1] * inv;
    s.h0k.assign((size_t)N * N * 2, 0.f);
    s.h0minusk.assign((size_t)N * N * 2, 0.f);
#pragma omp parallel for num_threads(s.threads) schedule(static)
    for (int iy = 0; iy < N; ++iy) {
        for (int ix = 0; ix < N; ++ix) {
            const float xx = (float)ix - (float)N / 2.0f, xy = (float)iy - (float)N / 2.0f;   // :76
            const float kx = (2.0f * kPi * xx) / s.L, ky = (2.0f * kPi * xy) / s.L;           // :77
            const float L_philips = (s.wind_speed * s.wind_speed) / kG;                       // :78
            float k_mag = sqrtf(kx * kx + ky * ky);                                           // :79 length(k)
            if (k_mag < 0.00001f) k_mag = 0.00001f;                                           // :81-82
            const float k_mag_sqr = k_mag * k_mag;                                            // :84
            const float sup = expf(-k_mag_sqr * s.suppression * s.suppression);               // :35-38
            // philips_power_spectrum(), :42-45. normalize(k) = k * inversesqrt(dot(k,k)) on the UNCLAMPED k.
            auto philips = [&](float px, float py) {
                const float rs = 1.0f / sqrtf(px * px + py * py);   // inf at k=0 -> 0*inf = NaN
                const float nx = px * rs, ny = py * rs;
                const float d  = nx * wdx + ny * wdy;
                return (s.amplitude * expf(-1.0f / (k_mag_sqr * L_philips * L_philips)) * (d * d) * sup) /
                       (k_mag_sqr * k_mag_sqr);
            };
            const float h0k      = clampf(sqrtf(philips(kx, ky)) / sqrtf(2.0f), -4000.0f, 4000.0f);   // :87
            const float h0minusk = clampf(sqrtf(philips(-kx, -ky)) / sqrtf(2.0f), -4000.0f, 4000.0f); // :88
            // gauss_rnd(), :51-68
            const float n0 = clampf(noise_fetch(s, 0, ix, iy), 0.001f, 1.0f);
            const float n1 = clampf(noise_fetch(s, 1, ix, iy), 0.001f, 1.0f);
            const float n2 = clampf(noise_fetch(s, 2, ix, iy), 0.001f, 1.0f);
            const float n3 = clampf(noise_fetch(s, 3, ix, iy), 0.001f, 1.0f);
            const float u0 = 2.0f * kPi * n0, v0 = sqrtf(-2.0f * logf(n1));
            const float u1 = 2.0f * kPi * n2, v1 = sqrtf(-2.0f * logf(n3));
            const float rx = v0 * cosf(u0), ry = v0 * sinf(u0), rz = v1 * cosf(u1), rw = v1 * sinf(u1);
            const size_t o = ((size_t)iy * N + ix) * 2;
            s.h0k[o] = rx * h0k;           s.h0k[o + 1] = ry * h0k;             // :92
            s.h0minusk[o] = rz * h0minusk; s.h0minusk[o + 1] = rw * h0minusk;   // :93
        }
    }
}

// main.cpp:587-608 + tilde_h0_t_cs.glsl:70-131
void tilde_h0_t(Sim& s, float t) {
    const int N = s.N;
#pragma omp parallel for num_threads(s.threads) schedule(static)
    for (int iy = 0; iy < N; ++iy) {
        for (int ix = 0; ix < N; ++ix) {
            const float xx = (float)ix - (float)N / 2.0f, xy = (float)iy - (float)N / 2.0f;  // :72
            const float kx = (2.0f * kPi * xx) / s.L, ky = (2.0f * kPi * xy) / s.L;          // :73
            float k_mag = sqrtf(kx * kx + ky * ky);                                          // :74
            if (k_mag < 0.00001f) k_mag = 0.00001f;                                          // :76-77
            const float w = sqrtf(kG * k_mag);                                               // :79
            const size_t o = ((size_t)iy * N + ix) * 2;
            const cpx fourier_amp{s.h0k[o], s.h0k[o + 1]};                                   // :81-87
            cpx fourier_amp_conj{s.h0minusk[o], s.h0minusk[o + 1]};                          // :89-92
            fourier_amp_conj = conjugate_as_shipped(fourier_amp_conj);                       // :94 (no-op)
            const float cosinus = cosf(w * t), sinus = sinf(w * t);                          // :96-97
            const cpx e_p{cosinus, sinus}, e_m{cosinus, -sinus};                             // :99-107
            const cpx hdy = add(mul(fourier_amp, e_p), mul(fourier_amp_conj, e_m));          // :110
            const cpx mx{0.0f, -kx / k_mag};                                                 // :113-116
            const cpx hdx = mul(mx, hdy);                                                    // :118
            const cpx mz{0.0f, -ky / k_mag};                                                 // :121-124
            const cpx hdz = mul(mz, hdy);                                                    // :126
            const size_t i = (size_t)iy * N + ix;
            s.hkt[1][i] = vec4{hdx.real, hdx.im, 0.0f, 1.0f};                                // :128
            s.hkt[0][i] = vec4{hdy.real, hdy.im, 0.0f, 1.0f};                                // :129
            s.hkt[2][i] = vec4{hdz.real, hdz.im, 0.0f, 1.0f};                                // :130
        }
    }
}

// One dispatch of butterfly_cs.glsl (:54-144): every texel does one radix-2 butterfly.
void butterfly_pass(const Sim& s, const vec4* src, vec4* dst, int direction, int stage) {
    const int N = s.N;
#pragma omp parallel for num_threads(s.threads) schedule(static)
    for (int y = 0; y < N; ++y) {
        for (int x = 0; x < N; ++x) {
            const int line = direction == 0 ? x : y;                              // :61 / :102
            const vec4 data = s.twiddle[(size_t)line * s.log2N + stage];
            vec4 p_, q_;
            if (direction == 0) {                                                 // horizontal, :62-63
                p_ = src[(size_t)y * N + (int)data.z];
                q_ = src[(size_t)y * N + (int)data.w];
            } else {                                                              // vertical, :103-104
                p_ = src[(size_t)((int)data.z) * N + x];
                q_ = src[(size_t)((int)data.w) * N + x];
            }
            const cpx p{p_.x, p_.y}, q{q_.x, q_.y}, w{data.x, data.y};
            const cpx H = add(p, mul(w, q));                                      // :71
            dst[(size_t)y * N + x] = vec4{H.real, H.im, 0.0f, 1.0f};              // :73
        }
    }
}

// main.cpp:612-683 butterfly_fft(tilde_h0_t, dst) + inversion_cs.glsl:25-43
void butterfly_fft(Sim& s, std::vector<vec4>& spectrum, std::vector<float>& dst) {
    const int N = s.N;
    vec4* pp[2] = {spectrum.data(), s.pingpong.data()};   // binding 1 = tilde_h0_t, binding 2 = m_ping_pong
    int pingpong = 0;
    for (int i = 0; i < s.log2N; ++i) {                   // :626-640 horizontal
        butterfly_pass(s, pp[pingpong], pp[pingpong ^ 1], 0, i);
        pingpong = (pingpong + 1) % 2;
    }
    for (int i = 0; i < s.log2N; ++i) {                   // :647-661 vertical
        butterfly_pass(s, pp[pingpong], pp[pingpong ^ 1], 1, i);
        pingpong = (pingpong + 1) % 2;
    }
    const vec4* fin = pp[pingpong];                       // :669 u_PingPong selects the plane read
    const float perms[2] = {1.0f, -1.0f};
#pragma omp parallel for num_threads(s.threads) schedule(static)
    for (int y = 0; y < N; ++y)
        for (int x = 0; x < N; ++x) {
            const float perm = perms[(x + y) % 2];                                       // inversion_cs.glsl:29-31
            const float h = fin[(size_t)y * N + x].x;
            dst[(size_t)y * N + x] = perm * (h / (float)(N * N));                        // :36
        }
}

// texture(s_HeightMap, uv).r with LINEAR + REPEAT on the single-mip R32F N x N height map
// (m_dy keeps the GL defaults: fw/src/ogl.cpp:449-456). Generic bilinear; at the shader's
// coordinates (texel corners) every weight is exactly 0.5.
inline float sample_linear_repeat(const float* h, int N, float u, float v) {
    const float fx = u * (float)N - 0.5f, fy = v * (float)N - 0.5f;
    const float flx = floorf(fx), fly = floorf(fy);
    const float ax = fx - flx, ay = fy - fly;
    const int m = N - 1;  // N is a power of two
    const int x0 = ((int)flx) & m, x1 = ((int)flx + 1) & m;
    const int y0 = ((int)fly) & m, y1 = ((int)fly + 1) & m;
    const float t00 = h[(size_t)y0 * N + x0], t10 = h[(size_t)y0 * N + x1];
    const float t01 = h[(size_t)y1 * N + x0], t11 = h[(size_t)y1 * N + x1];
    const float top = t00 * (1.0f - ax) + t10 * ax;
    const float bot = t01 * (1.0f - ax) + t11 * ax;
    return top * (1.0f - ay) + bot * ay;
}

// main.cpp:687-707 + normal_map_cs.glsl:24-54
void generate_normal_map(Sim& s) {
    const int N = s.N;
    const float* h = s.disp[0].data();
#pragma omp parallel for num_threads(s.threads) schedule(static)
    for (int y = 0; y < N; ++y)
        for (int x = 0; x < N; ++x) {
            const float tu = (float)x / (float)N, tv = (float)y / (float)N;   // :33
            const float ts = 1.0f / (float)N;                                 // :35
            const float z0 = sample_linear_repeat(h, N, tu - ts, tv - ts);    // :37-44
            const float z1 = sample_linear_repeat(h, N, tu, tv - ts);
            const float z2 = sample_linear_repeat(h, N, tu + ts, tv - ts);
            const float z3 = sample_linear_repeat(h, N, tu - ts, tv);
            const float z4 = sample_linear_repeat(h, N, tu + ts, tv);
            const float z5 = sample_linear_repeat(h, N, tu - ts, tv + ts);
            const float z6 = sample_linear_repeat(h, N, tu, tv + ts);
            const float z7 = sample_linear_repeat(h, N, tu + ts, tv + ts);
            const float nz = z0 + 2.0f * z1 + z2 - z5 - 2.0f * z6 - z7;       // :49
            const float nx = z0 + 2.0f * z3 + z5 - z2 - 2.0f * z4 - z7;       // :50
            const float ny = 1.0f;                                            // :51
            const float r = 1.0f / sqrtf(nx * nx + ny * ny + nz * nz);        // normalize()
            s.normal[(size_t)y * N + x] = vec4{nx * r, ny * r, nz * r, 1.0f}; // :53
        }
}

// EXTENSION (not in the reference): Jacobian of the horizontal displacement with the
// consumer's sign convention (grid_tes.glsl:61-62: X = x - lambda*Dx, Z = z - lambda*Dz), central
// differences with wrap on the output grid, spacing L/N.
void generate_jacobian(Sim& s, float lambda) {
    const int N = s.N, m = N - 1;
    const float* dx = s.disp[1].data();
    const float* dz = s.disp[2].data();
    const float inv2h = (float)N / (2.0f * s.L);
    s.jacobian.resize((size_t)N * N);
#pragma omp parallel for num_threads(s.threads) schedule(static)
    for (int y = 0; y < N; ++y)
        for (int x = 0; x < N; ++x) {
            const int xm = (x - 1) & m, xp = (x + 1) & m, ym = (y - 1) & m, yp = (y + 1) & m;
            const float dxdx = (dx[(size_t)y * N + xp] - dx[(size_t)y * N + xm]) * inv2h;
            const float dxdz = (dx[(size_t)yp * N + x] - dx[(size_t)ym * N + x]) * inv2h;
            const float dzdx = (dz[(size_t)y * N + xp] - dz[(size_t)y * N + xm]) * inv2h;
            const float dzdz = (dz[(size_t)yp * N + x] - dz[(size_t)ym * N + x]) * inv2h;
            s.jacobian[(size_t)y * N + x] =
                (1.0f - lambda * dxdx) * (1.0f - lambda * dzdz) - (lambda * dxdz) * (lambda * dzdx);
        }
}

}  // namespace

extern "C" {

// Create an oracle simulation. noise: 4 planes of noise_w*noise_h bytes (R channel of the RGBA8 images).
void* oracle_create(int N, float L, float wind_speed, float wind_dir_x, float wind_dir_y, float amplitude,
                    float suppression, const uint8_t* noise, int noise_w, int noise_h, int threads) {
    if (N < 2 || (N & (N - 1)) != 0) return nullptr;
    Sim* s = new Sim();
    s->N = N; s->log2N = ilog2(N); s->L = L; s->wind_speed = wind_speed;
    s->wind_dir[0] = wind_dir_x; s->wind_dir[1] = wind_dir_y;
    s->amplitude = amplitude; s->suppression = suppression;
    s->noise_w = noise_w; s->noise_h = noise_h;
    for (int j = 0; j < 4; ++j)
        s->noise[j].assign(noise + (size_t)j * noise_w * noise_h, noise + (size_t)(j + 1) * noise_w * noise_h);
    s->threads = threads > 0 ? threads : 1;
    for (auto& p : s->hkt) p.assign((size_t)N * N, vec4{0, 0, 0, 0});
    s->pingpong.assign((size_t)N * N, vec4{0, 0, 0, 0});
    for (auto& d : s->disp) d.assign((size_t)N * N, 0.f);
    s->normal.assign((size_t)N * N, vec4{0, 0, 0, 0});
    // init(): main.cpp:218-220
    tilde_h0_k(*s);
    generate_bit_reversed_indices(*s);
    generate_twiddle_factors(*s);
    return s;
}

void oracle_destroy(void* h) { delete static_cast<Sim*>(h); }

int oracle_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// Overwrite the initial spectrum (for known-answer tests that drive the FFT chain with chosen spectra).
void oracle_set_h0(void* h, const float* h0k, const float* h0minusk) {
    Sim& s = *static_cast<Sim*>(h);
    std::memcpy(s.h0k.data(), h0k, s.h0k.size() * sizeof(float));
    std::memcpy(s.h0minusk.data(), h0minusk, s.h0minusk.size() * sizeof(float));
}

void oracle_get_h0(void* h, float* h0k, float* h0minusk) {
    Sim& s = *static_cast<Sim*>(h);
    std::memcpy(h0k, s.h0k.data(), s.h0k.size() * sizeof(float));
    std::memcpy(h0minusk, s.h0minusk.data(), s.h0minusk.size() * sizeof(float));
}

void oracle_get_twiddle(void* h, float* out /* N*log2N*4 */, int32_t* bitrev /* N */) {
    Sim& s = *static_cast<Sim*>(h);
    std::memcpy(out, s.twiddle.data(), s.twiddle.size() * sizeof(vec4));
    std::memcpy(bitrev, s.bit_reversed.data(), s.bit_reversed.size() * sizeof(int32_t));
}

// Spectrum planes after tilde_h0_t (before the FFT destroys them): 3 planes (dy,dx,dz) of N*N*2 floats.
void oracle_spectrum(void* h, float t, float* out) {
    Sim& s = *static_cast<Sim*>(h);
    tilde_h0_t(s, t);
    const size_t n = (size_t)s.N * s.N;
    for (int c = 0; c < 3; ++c)
        for (size_t i = 0; i < n; ++i) { out[(c * n + i) * 2] = s.hkt[c][i].x; out[(c * n + i) * 2 + 1] = s.hkt[c][i].y; }
}

// One frame: update() lines main.cpp:240-244. Outputs may be null. lambda < 0 skips the Jacobian.
void oracle_frame(void* h, float t, float lambda, float* dy, float* dx, float* dz, float* normal, float* jac) {
    Sim& s = *static_cast<Sim*>(h);
    tilde_h0_t(s, t);
    butterfly_fft(s, s.hkt[0], s.disp[0]);   // main.cpp:241  dy
    butterfly_fft(s, s.hkt[1], s.disp[1]);   // main.cpp:242  dx
    butterfly_fft(s, s.hkt[2], s.disp[2]);   // main.cpp:243  dz
    generate_normal_map(s);                  // main.cpp:244
    const size_t n = (size_t)s.N * s.N;
    if (dy) std::memcpy(dy, s.disp[0].data(), n * sizeof(float));
    if (dx) std::memcpy(dx, s.disp[1].data(), n * sizeof(float));
    if (dz) std::memcpy(dz, s.disp[2].data(), n * sizeof(float));
    if (normal) std::memcpy(normal, s.normal.data(), n * sizeof(vec4));
    if (lambda >= 0.0f) {
        generate_jacobian(s, lambda);
        if (jac) std::memcpy(jac, s.jacobian.data(), n * sizeof(float));
    }
}

// Normal map of an arbitrary height field (known-answer tests of normal_map_cs.glsl).
void oracle_normal_of(void* h, const float* height, float* normal) {
    Sim& s = *static_cast<Sim*>(h);
    std::memcpy(s.disp[0].data(), height, (size_t)s.N * s.N * sizeof(float));
    generate_normal_map(s);
    std::memcpy(normal, s.normal.data(), (size_t)s.N * s.N * sizeof(vec4));
}

}  // extern "C"
