// TEST INFRASTRUCTURE ONLY (oracle build shim): glm/gtx/vector_angle.hpp
// (included by advect_floating_items.cpp:6, where nothing of it is used, and by swarm.cpp:4,
// which calls glm::orientedAngle and -- as the real header pulls it in -- glm::rotate).
// glm::orientedAngle(vec2, vec2) as GLM defines it (gtx/vector_angle.inl).
#pragma once
#include "../glm.hpp"
#include "rotate_vector.hpp"
namespace glm {
inline float orientedAngle(vec2 x, vec2 y) {
  const float d = x.x * y.x + x.y * y.y;
  const float a = std::acos(d < -1.0f ? -1.0f : (d > 1.0f ? 1.0f : d));
  const vec2 r = rotate(x, a);
  const bool same = std::fabs(y.x - r.x) < 0.0001f && std::fabs(y.y - r.y) < 0.0001f;
  return same ? a : -a;
}
} // namespace glm
