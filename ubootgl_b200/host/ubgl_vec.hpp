// ubgl_vec.hpp -- the few glm vector types the Simulation API mentions
// (simulation.hpp:82,100-103,140).  When the real glm is on the include path it
// is used; otherwise this header supplies source-compatible minimal types so
// that the drop-in classes build on a box without glm (the B200 image has none).
#pragma once
#if __has_include(<glm/glm.hpp>) && !defined(UBGL_NO_GLM)
#include <glm/glm.hpp>
#else
#include <cmath>
namespace glm {
template <typename T> struct tvec2 {
  T x{}, y{};
  tvec2() = default;
  tvec2(T x_, T y_) : x(x_), y(y_) {}
  template <typename U> tvec2(const tvec2<U> &o) : x((T)o.x), y((T)o.y) {}
  tvec2 operator+(const tvec2 &o) const { return {T(x + o.x), T(y + o.y)}; }
  tvec2 operator-(const tvec2 &o) const { return {T(x - o.x), T(y - o.y)}; }
  tvec2 operator*(T s) const { return {T(x * s), T(y * s)}; }
  tvec2 operator/(T s) const { return {T(x / s), T(y / s)}; }
  tvec2 operator-(T s) const { return {T(x - s), T(y - s)}; }
  tvec2 operator+(T s) const { return {T(x + s), T(y + s)}; }
};
using vec2 = tvec2<float>;
using ivec2 = tvec2<int>;
struct vec3 {
  float x{}, y{}, z{};
  vec3() = default;
  vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
  vec3(const vec2 &v, float z_) : x(v.x), y(v.y), z(z_) {}
};
inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline vec2 fract(const vec2 &v) { return {v.x - std::floor(v.x), v.y - std::floor(v.y)}; }
} // namespace glm
#endif
