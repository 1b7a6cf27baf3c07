// TEST INFRASTRUCTURE ONLY (oracle/): a GLSL 4.30 subset as C++, so that the reference's
// UNMODIFIED compute-shader sources (interp_shader.cs, advect_tracer_points.cs, shift_tracers.cs) compile and
// run on the CPU, one call of main() per invocation.  This image has no GL implementation
// (no libGL / libEGL / OSMesa), so this is how the C restatement of the shaders
// (ubgl_oracle_next.c: orc_colocate, orc_tracers_advect) and the CUDA kernels are pinned to
// the shader SOURCE: oracle/Makefile pipes each .cs file where it lies in /root/reference
// through `sed` (drops `#version`; rewrites the four `buffer NAME { T name[]; };` interface
// blocks, which have no C++ spelling, to `T *name;`) straight into g++ -include glsl_shim.hpp.
// Functions, main() and every expression of the shaders are compiled as written.
//
// What the shim must get right, with the GLSL / GL 4.5 rule it follows:
//  * float literals are single precision        -> the Makefile passes -fsingle-precision-constant
//  * constructor arguments are evaluated left to right (GLSL 4.30 section 6.1.1) -> the
//    constructors are function-like macros expanding to braced initialisation
//  * uint arithmetic wraps mod 2^32, uint -> float conversions as C's
//  * texture(sampler2D, vec2) from a compute shader: no derivatives, level 0, magnification;
//    the textures are created with glTexStorage2D and never get a glTexParameter before the
//    shaders run (velocity_textures.cpp:31-60), so MAG_FILTER = GL_LINEAR and WRAP_S/T =
//    GL_REPEAT (the defaults); GL 4.5 section 8.14.2: u' = s w, i0 = wrap(floor(u' - 1/2)),
//    i1 = wrap(i0 + 1), alpha = frac(u' - 1/2), tau = (1-a)(1-b) t00 + a(1-b) t10 + (1-a) b t01
//    + a b t11.  The specification fixes neither the precision of the weights nor the order of
//    the sum; fp32 weights and the order ((t00 + t10) + t01) + t11 are this project's choice for
//    the restatement, the kernels and this shim alike (real GPUs use ~8-bit weights).
//  * imageStore writes the texel's components for the image format (rg32f: x, y; r32f: x)
#pragma once
#include <cmath>

typedef unsigned int uint;

struct vec2_t {
  float x, y;
  vec2_t() : x(0), y(0) {}
  vec2_t(float a, float b) : x(a), y(b) {}
};
struct ivec2_t;
struct uvec2_t {
  uint x, y;
  uvec2_t() : x(0), y(0) {}
  uvec2_t(uint a, uint b) : x(a), y(b) {}
};
struct ivec2_t {
  int x, y;
  ivec2_t() : x(0), y(0) {}
  ivec2_t(int a, int b) : x(a), y(b) {}
  ivec2_t(const uvec2_t &u) : x((int)u.x), y((int)u.y) {}
};
struct bvec2_t {
  bool x, y;
  bvec2_t(bool a, bool b) : x(a), y(b) {}
};
struct bvec4_t {
  bool x, y, z, w;
  bvec4_t(const bvec2_t &a, const bvec2_t &b) : x(a.x), y(a.y), z(b.x), w(b.y) {}
};
template <int A, int B> struct glsl_swz2f { // .xy of a vec4: shares its storage
  float v[4];
  operator vec2_t() const { return vec2_t{v[A], v[B]}; }
};
template <int A, int B> struct glsl_swz2u {
  uint v[3];
  operator uvec2_t() const { return uvec2_t{v[A], v[B]}; }
};
struct vec4_t {
  union {
    struct { float x, y, z, w; };
    struct { float r, g, b, a; };
    glsl_swz2f<0, 1> xy;
  };
  vec4_t() : x(0), y(0), z(0), w(0) {}
  vec4_t(float a_, float b_, float c_, float d_) : x(a_), y(b_), z(c_), w(d_) {}
};
struct uvec3_t {
  union {
    struct { uint x, y, z; };
    glsl_swz2u<0, 1> xy;
  };
  uvec3_t() : x(0), y(0), z(0) {}
};

// constructors: braced initialisation = left-to-right argument evaluation, as GLSL requires
#define vec2(...) vec2_t{__VA_ARGS__}
#define vec4(...) vec4_t{__VA_ARGS__}
#define ivec2(...) ivec2_t{__VA_ARGS__}
#define bvec4(...) bvec4_t{__VA_ARGS__}
typedef vec2_t vec2;
typedef vec4_t vec4;
typedef ivec2_t ivec2;
typedef bvec4_t bvec4;

inline vec2_t operator+(const vec2_t &a, const vec2_t &b) { return vec2_t{a.x + b.x, a.y + b.y}; }
inline vec2_t &operator+=(vec2_t &a, const vec2_t &b) { a.x += b.x; a.y += b.y; return a; }
inline vec2_t operator*(const vec2_t &a, float s) { return vec2_t{a.x * s, a.y * s}; }
inline vec2_t operator*(const vec2_t &a, const vec2_t &b) { return vec2_t{a.x * b.x, a.y * b.y}; }
inline vec2_t operator/(const vec2_t &a, const vec2_t &b) { return vec2_t{a.x / b.x, a.y / b.y}; }
inline bvec2_t lessThan(const vec2_t &a, const vec2_t &b) { return bvec2_t{a.x < b.x, a.y < b.y}; }
inline bvec2_t greaterThan(const vec2_t &a, const vec2_t &b) { return bvec2_t{a.x > b.x, a.y > b.y}; }
inline bool any(const bvec4_t &b) { return b.x || b.y || b.z || b.w; }
inline float length(const vec2_t &a) { return sqrtf(a.x * a.x + a.y * a.y); }

// a bound texture / image: level 0 of a w x h texture with nc interleaved fp32 components
struct sampler2D {
  const float *texels = nullptr;
  int w = 0, h = 0, nc = 1;
};
struct image2D {
  float *texels = nullptr;
  int w = 0, h = 0, nc = 1;
};
inline int glsl_repeat(int i, int n) {
  const int m = i % n;
  return m < 0 ? m + n : m;
}
inline vec4_t texture(const sampler2D &t, const vec2_t &st) { // GL_LINEAR, GL_REPEAT, level 0
  const float u = st.x * (float)t.w - 0.5f, v = st.y * (float)t.h - 0.5f;
  const float fu = floorf(u), fv = floorf(v);
  const float a = u - fu, b = v - fv;
  const int i0 = glsl_repeat((int)fu, t.w), i1 = glsl_repeat((int)fu + 1, t.w);
  const int j0 = glsl_repeat((int)fv, t.h), j1 = glsl_repeat((int)fv + 1, t.h);
  const float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
  float out[4] = {0.0f, 0.0f, 0.0f, 1.0f}; // missing components read (0, 0, 1)
  for (int c = 0; c < t.nc; c++) {
    const float t00 = t.texels[((size_t)j0 * t.w + i0) * t.nc + c], t10 = t.texels[((size_t)j0 * t.w + i1) * t.nc + c];
    const float t01 = t.texels[((size_t)j1 * t.w + i0) * t.nc + c], t11 = t.texels[((size_t)j1 * t.w + i1) * t.nc + c];
    out[c] = ((w00 * t00 + w10 * t10) + w01 * t01) + w11 * t11;
  }
  return vec4_t{out[0], out[1], out[2], out[3]};
}
inline void imageStore(const image2D &img, const ivec2_t &p, const vec4_t &v) {
  if (p.x < 0 || p.y < 0 || p.x >= img.w || p.y >= img.h) return; // out-of-bounds stores are discarded
  const float c[4] = {v.x, v.y, v.z, v.w};
  for (int k = 0; k < img.nc; k++) img.texels[((size_t)p.y * img.w + p.x) * img.nc + k] = c[k];
}

extern uvec3_t gl_GlobalInvocationID;

// storage / layout qualifiers with no meaning on the CPU
#define uniform
#define writeonly
#define restrict
#define layout(...)
#define in
