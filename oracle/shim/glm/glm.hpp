// TEST INFRASTRUCTURE ONLY (oracle build shim) -- not part of the product path.
//
// Minimal stand-in for the parts of the GLM maths library that the reference
// fluid-solver translation units touch when they are compiled, unmodified, from
// /root/reference into oracle/_ref (see oracle/Makefile).  GLM itself is an
// un-vendored system dependency of the reference (README.md:35, Makefile:9)
// and is absent from this image.  Only trivial vector arithmetic is used on the
// hot path (simulation.cpp:174 sink position, db2dgrid.hpp:32-33 ivec2 index).
#pragma once
#include <cmath>
#include <initializer_list>

namespace glm {

template <typename T> struct tvec3;

template <typename T> struct tvec2 {
  T x, y;
  tvec2() : x(0), y(0) {}
  tvec2(T a, T b) : x(a), y(b) {}
  explicit tvec2(T a) : x(a), y(a) {}
  template <typename U> tvec2(const tvec2<U> &o) : x((T)o.x), y((T)o.y) {}
  template <typename U> explicit tvec2(const tvec3<U> &o);
  T &operator[](int i) { return i == 0 ? x : y; }
  T operator[](int i) const { return i == 0 ? x : y; }
  tvec2 &operator+=(const tvec2 &o) { x += o.x; y += o.y; return *this; }
  tvec2 &operator-=(const tvec2 &o) { x -= o.x; y -= o.y; return *this; }
  tvec2 &operator*=(T s) { x *= s; y *= s; return *this; }
  tvec2 &operator/=(T s) { x /= s; y /= s; return *this; }
  tvec2 operator-() const { return tvec2(-x, -y); }
};

template <typename T> struct tvec3 {
  T x, y, z;
  tvec3() : x(0), y(0), z(0) {}
  tvec3(T a, T b, T c) : x(a), y(b), z(c) {}
  template <typename A, typename B, typename C>
  tvec3(A a, B b, C c) : x((T)a), y((T)b), z((T)c) {}
  tvec3(const tvec2<T> &v, T c) : x(v.x), y(v.y), z(c) {}
};

template <typename T>
template <typename U>
tvec2<T>::tvec2(const tvec3<U> &o) : x((T)o.x), y((T)o.y) {}

template <typename T> struct tvec4 { T x, y, z, w; };

typedef tvec2<float> vec2;
typedef tvec2<int> ivec2;
typedef tvec3<float> vec3;
typedef tvec4<float> vec4;

struct mat4 {
  float m[16];
  mat4() : m{} {}
  explicit mat4(float) : m{} {}
};

#define UBGL_SHIM_BINOP(OP)                                                    \
  template <typename T> tvec2<T> operator OP(const tvec2<T> &a,                \
                                             const tvec2<T> &b) {              \
    return tvec2<T>(a.x OP b.x, a.y OP b.y);                                   \
  }                                                                            \
  template <typename T> tvec2<T> operator OP(const tvec2<T> &a, T s) {         \
    return tvec2<T>(a.x OP s, a.y OP s);                                       \
  }                                                                            \
  template <typename T> tvec2<T> operator OP(T s, const tvec2<T> &a) {         \
    return tvec2<T>(s OP a.x, s OP a.y);                                       \
  }
UBGL_SHIM_BINOP(+)
UBGL_SHIM_BINOP(-)
UBGL_SHIM_BINOP(*)
UBGL_SHIM_BINOP(/)
#undef UBGL_SHIM_BINOP

inline float min(float a, float b) { return b < a ? b : a; }
inline float max(float a, float b) { return a < b ? b : a; }
inline int min(int a, int b) { return b < a ? b : a; }
inline int max(int a, int b) { return a < b ? b : a; }
inline ivec2 min(ivec2 a, ivec2 b) { return ivec2(min(a.x, b.x), min(a.y, b.y)); }
inline ivec2 max(ivec2 a, ivec2 b) { return ivec2(max(a.x, b.x), max(a.y, b.y)); }
inline vec2 min(vec2 a, vec2 b) { return vec2(min(a.x, b.x), min(a.y, b.y)); }
inline vec2 max(vec2 a, vec2 b) { return vec2(max(a.x, b.x), max(a.y, b.y)); }
inline vec2 clamp(vec2 v, vec2 lo, vec2 hi) { return min(max(v, lo), hi); }
inline float clamp(float v, float lo, float hi) { return min(max(v, lo), hi); }
inline float fract(float v) { return v - std::floor(v); }
inline vec2 fract(vec2 v) { return vec2(fract(v.x), fract(v.y)); }
inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline float dot(vec2 a, vec2 b) { return a.x * b.x + a.y * b.y; }
inline float length(vec2 a) { return std::sqrt(dot(a, a)); }
inline vec2 normalize(vec2 a) { return a / length(a); }
inline float sign(float v) { return (float)((0.0f < v) - (v < 0.0f)); }
inline float abs(float v) { return std::fabs(v); }
inline vec2 abs(vec2 v) { return vec2(std::fabs(v.x), std::fabs(v.y)); }
// glm::reflect (detail/func_geometric.inl): I - N * dot(N, I) * 2
inline vec2 reflect(vec2 I, vec2 N) { return I - N * dot(N, I) * 2.0f; }

} // namespace glm
