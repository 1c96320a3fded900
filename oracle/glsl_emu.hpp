// glsl_emu.hpp — just enough of GLSL 4.50 compute-shader semantics, as C++20, to compile the reference's OWN compute
// shaders (read at build time from <reference>/src/shader/*_cs.glsl by oracle/make_ref.py) into oracle/_ref/libow_ref.so.
// TEST INFRASTRUCTURE (pins the oracle; see oracle/make_ref.py). Nothing of the reference is copied here: this file is the
// "GL driver" side — vector types with the swizzles those shaders use, images, samplers, built-ins.
//
// Arithmetic conventions (the GL spec leaves built-in precision to the driver; same choices as oracle/ow_oracle.cpp so that the
// two can be compared to the last bit where they compute the same expression): fp32 everywhere, libm-accurate sqrtf/expf/logf/
// sinf/cosf/powf, normalize(v) = v * (1/sqrt(dot(v,v))), clamp = fminf(fmaxf()), mod(x,y) = x - y*floor(x/y), texture() with the
// sampler's own filter/wrap state (NEAREST+CLAMP_TO_EDGE for the noise images, LINEAR+REPEAT for the height map: the states the
// reference's textures have, fw/src/ogl.cpp:449-456 and src/main.cpp:1085-1145).
//
// No <cmath> here on purpose: the generated translation unit must not see ::sqrt(double) & co., or `sqrt(2.0f)` in shader code
// would be ambiguous or silently double. The built-ins are defined in oracle/ref_driver.cpp.
#pragma once

namespace glsl {

struct vec2;
struct ivec2;
struct vec4;

template <class V, class S, int A, int B>
struct swz2 {                         // two-component swizzle living inside its parent's storage
    S d[4];
    operator V() const;
};

struct uvec2 {
    unsigned x, y;
};
using uswz_xy = swz2<uvec2, unsigned, 0, 1>;          // gl_GlobalInvocationID.xy

struct swz4 {                                          // .rgba of a vec4
    float d[4];
    operator vec4() const;
};

struct vec2 {
    float x, y;
    vec2() : x(0), y(0) {}
    template <class T, class U> vec2(T a, U b) : x((float)a), y((float)b) {}
    vec2(const uvec2& u) : x((float)u.x), y((float)u.y) {}      // GLSL's implicit uvec2 -> vec2
    vec2(const uswz_xy& u) : x((float)u.d[0]), y((float)u.d[1]) {}
};

struct ivec2 {
    int x, y;
    ivec2() : x(0), y(0) {}
    template <class T, class U> ivec2(T a, U b) : x((int)a), y((int)b) {}
    explicit ivec2(const vec2& v) : x((int)v.x), y((int)v.y) {}
    explicit ivec2(const uvec2& v) : x((int)v.x), y((int)v.y) {}
    explicit ivec2(const uswz_xy& v) : x((int)v.d[0]), y((int)v.d[1]) {}
};

struct vec3 {
    float x, y, z;
    vec3() : x(0), y(0), z(0) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
};

struct vec4 {
    union {
        struct { float x, y, z, w; };
        struct { float r, g, b, a; };
        swz2<vec2, float, 0, 1> xy, rg;
        swz2<vec2, float, 2, 3> zw;
        swz4 rgba;
        float d[4];
    };
    vec4() : x(0), y(0), z(0), w(0) {}
    template <class A, class B, class C, class D> vec4(A a_, B b_, C c_, D d_) : x((float)a_), y((float)b_), z((float)c_), w((float)d_) {}
    template <class C, class D> vec4(const vec2& v, C c_, D d_) : x(v.x), y(v.y), z((float)c_), w((float)d_) {}
    template <class D> vec4(const vec3& v, D d_) : x(v.x), y(v.y), z(v.z), w((float)d_) {}
};
inline swz4::operator vec4() const { return vec4(d[0], d[1], d[2], d[3]); }

template <class V, class S, int A, int B>
inline swz2<V, S, A, B>::operator V() const { return V(d[A], d[B]); }

struct uvec3 {
    union {
        struct { unsigned x, y, z; };
        swz2<uvec2, unsigned, 0, 1> xy;
        unsigned d[4];
    };
    uvec3() : x(0), y(0), z(0) {}
};
template <>
inline swz2<uvec2, unsigned, 0, 1>::operator uvec2() const { return uvec2{d[0], d[1]}; }

// ---- operators the shaders use ----
inline vec2 operator-(const vec2& a) { return vec2(-a.x, -a.y); }
inline vec2 operator-(const vec2& a, float s) { return vec2(a.x - s, a.y - s); }
inline vec2 operator+(const vec2& a, const vec2& b) { return vec2(a.x + b.x, a.y + b.y); }
inline vec2 operator*(const vec2& a, float s) { return vec2(a.x * s, a.y * s); }
inline vec2 operator/(const vec2& a, float s) { return vec2(a.x / s, a.y / s); }
inline vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
template <int A, int B> inline vec2 operator*(const swz2<vec2, float, A, B>& a, float s) { return vec2(a.d[A] * s, a.d[B] * s); }
inline vec2 operator/(const swz2<uvec2, unsigned, 0, 1>& a, float s) { return vec2((float)a.d[0] / s, (float)a.d[1] / s); }

// ---- built-ins (defined in ref_driver.cpp with libm) ----
float sqrt(float);
float exp(float);
float log(float);
float sin(float);
float cos(float);
float pow(float, float);
float mod(float, float);
float clamp(float, float, float);
float dot(const vec2&, const vec2&);
float length(const vec2&);
vec2 normalize(const vec2&);
vec3 normalize(const vec3&);

// ---- images and samplers ----
struct image2D {                      // every format is held as 4 floats per texel; rg32f / r32f images just ignore the rest
    float* data = nullptr;
    int w = 0, h = 0;
};
inline vec4 imageLoad(const image2D& im, const ivec2& p) {
    const float* t = im.data + 4 * ((long)p.y * im.w + p.x);
    return vec4(t[0], t[1], t[2], t[3]);
}
inline void imageStore(const image2D& im, const ivec2& p, const vec4& v) {
    float* t = im.data + 4 * ((long)p.y * im.w + p.x);
    t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
}

struct sampler2D {
    const float* data = nullptr;      // single channel (the .r the shaders read)
    int w = 0, h = 0;
    bool linear = false, repeat = false;
};
vec4 texture(const sampler2D& s, const vec2& uv);

extern thread_local uvec3 gl_GlobalInvocationID;

}  // namespace glsl
