// TEST INFRASTRUCTURE ONLY (oracle build shim): glm/gtx/rotate_vector.hpp.
// glm::rotate(vec2, angle) as GLM defines it (gtx/rotate_vector.inl): the
// counter-clockwise 2-D rotation.
#pragma once
#include "../glm.hpp"
namespace glm {
inline vec2 rotate(vec2 v, float angle) {
  const float c = std::cos(angle), s = std::sin(angle);
  return vec2(v.x * c - v.y * s, v.x * s + v.y * c);
}
} // namespace glm
