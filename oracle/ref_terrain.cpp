// TEST INFRASTRUCTURE ONLY -- fixture generation, never on the product path.
// Forwards to the reference's own Terrain class (terrain.hpp:9, terrain.cpp:20)
// to turn a level PNG into the simulation-resolution flag mask, exactly as
// UbootGlApp does (ubootgl_app.hpp:29-30).  Used by tests/golden/make_golden.py.
#include "terrain.hpp"
#include <cstring>

extern "C" {
void *ref_terrain_create(const char *png, int scale) {
  return new Terrain(png, scale);
}
void ref_terrain_destroy(void *t) { delete (Terrain *)t; }
void ref_terrain_size(void *t, int *w, int *h) {
  *w = ((Terrain *)t)->flagSimRes.width;
  *h = ((Terrain *)t)->flagSimRes.height;
}
void ref_terrain_flag(void *t_, float *dst) {
  auto *t = (Terrain *)t_;
  std::memcpy(dst, t->flagSimRes.data(),
              sizeof(float) * (size_t)t->flagSimRes.width * t->flagSimRes.height);
}
void ref_terrain_draw_circle(void *t, float x, float y, int diam, float val) {
  ((Terrain *)t)->drawCircle(glm::vec2(x, y), diam, val);
}
}
