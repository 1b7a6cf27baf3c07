// advect_floating_items.cpp -- Simulation::advectFloatingItems / advectFloatingItemsSimple
// (te42kyfo/ubootgl simulation.hpp:116-117, called every frame by ubootgl_app.cpp:129-130)
// for the B200 drop-in: same signatures, the bodies run on the GPU.
//
// The reference walks registry.view<CoItem, CoKinematics[Simple]>() serially on the render
// thread, sampling vx / vy / p / flag and scattering reaction forces into vx_accum / vy_accum
// under accum_mutex (advect_floating_items.cpp:16-146, :148-274).  Here the view is copied,
// in iteration order (the order the reference's neighbour loop and its rounding depend on),
// into ubgl_item records (= CoItem + CoKinematics[Simple], components.hpp:6-43, field for
// field), ubgl_items_advect[_simple] advances them against the device-resident fields and
// adds the reaction forces to the DEVICE accumulators, and the records are written back to
// the components.  Nothing of the accumulators crosses PCIe.
//
// Compiled only inside a tree that has the reference's components.hpp and the vendored entt
// on the include path (UBGL_HAVE_REGISTRY, see simulation.hpp); this repository's own build
// has neither.  A maintainer who prefers the CPU item path keeps the reference's own
// advect_floating_items.cpp instead of this file: it compiles and links against the same
// header (host mirrors, SyncMode::MIRROR).
#include "simulation.hpp"
#ifdef UBGL_HAVE_REGISTRY
#include "../../include/ubgl.h"
#include <stdexcept>
#include <string>
#include <vector>

namespace {
void ck(int rc, const char *what) {
  if (rc != UBGL_OK)
    throw std::runtime_error(std::string(what) + ": libubgl error " + std::to_string(rc) + ": " +
                             ubgl_last_error());
}

template <class Kin> void pack(const CoItem &it, const Kin &k, ubgl_item &r) {
  r.size[0] = it.size.x; r.size[1] = it.size.y;
  r.pos[0] = it.pos.x; r.pos[1] = it.pos.y;
  r.rotation = it.rotation;
  r.mass = k.mass;
  r.vel[0] = k.vel.x; r.vel[1] = k.vel.y;
  r.force[0] = k.force.x; r.force[1] = k.force.y;
  r.angVel = k.angVel;
  r.angForce = k.angForce;
  r.bumpCount = k.bumpCount;
}
template <class Kin> void unpack(const ubgl_item &r, CoItem &it, Kin &k) {
  it.pos = glm::vec2(r.pos[0], r.pos[1]);
  it.rotation = r.rotation;
  k.vel = glm::vec2(r.vel[0], r.vel[1]);
  k.force = glm::vec2(r.force[0], r.force[1]);
  k.angVel = r.angVel;
  k.angForce = r.angForce;
  k.bumpCount = r.bumpCount;
}

template <class Kin, class Advect>
void run(Simulation &sim, std::shared_ptr<ubgl_items> &set, entt::registry &registry, float gameDT,
         Advect advect) {
  auto view = registry.view<CoItem, Kin>();
  std::vector<ubgl_item> rec;
  std::vector<entt::entity> who;
  for (auto e : view) {
    rec.emplace_back();
    pack(view.template get<CoItem>(e), view.template get<Kin>(e), rec.back());
    who.push_back(e);
  }
  if (rec.empty()) return;
  // host edits since the last step (shiftMap, setGrids, a CPU scatter into the accumulators)
  // reach the device before the items sample it
  sim.syncToDevice();
  if (!set) {
    ubgl_items_t *raw = nullptr;
    ck(ubgl_items_create(0, &raw), "ubgl_items_create");
    set.reset(raw, [](ubgl_items_t *p) { ubgl_items_destroy(p); });
  }
  ck(ubgl_items_upload(set.get(), rec.data(), (int)rec.size()), "ubgl_items_upload");
  ck(advect(set.get(), sim.handle(), gameDT), "ubgl_items_advect");
  int n = 0;
  ck(ubgl_items_download(set.get(), rec.data(), (int)rec.size(), &n), "ubgl_items_download");
  for (int i = 0; i < n; i++)
    unpack(rec[i], view.template get<CoItem>(who[i]), view.template get<Kin>(who[i]));
}
} // namespace

// input in standard grid space (advect_floating_items.cpp:11-14, interpolators.hpp:11-27);
// host-side O(1) read of the velocity mirrors for callers outside the item loop
glm::vec2 Simulation::bilinearVel(glm::vec2 c) {
  auto sample = [](const DoubleBuffered2DGrid &g, float cx, float cy) {
    cx = std::fmin(std::fmax(cx, 0.0f), g.width - 1.1f);
    cy = std::fmin(std::fmax(cy, 0.0f), g.height - 1.1f);
    const int ix = (int)cx, iy = (int)cy;
    const float sx = cx - std::floor(cx), sy = cy - std::floor(cy);
    const float v1 = g(ix, iy), v2 = g(ix + 1, iy), v3 = g(ix, iy + 1), v4 = g(ix + 1, iy + 1);
    const float vm1 = v1 + (v2 - v1) * sx, vm2 = v3 + (v4 - v3) * sx;
    return vm1 + (vm2 - vm1) * sy;
  };
  const DoubleBuffered2DGrid &ux = vx, &uy = vy;
  return glm::vec2(sample(ux, c.x - 0.5f, c.y), sample(uy, c.x, c.y - 0.5f));
}

void Simulation::advectFloatingItems(entt::registry &registry, float gameDT) {
  run<CoKinematics>(*this, items_[0], registry, gameDT, ubgl_items_advect);
}

void Simulation::advectFloatingItemsSimple(entt::registry &registry, float gameDT) {
  run<CoKinematicsSimple>(*this, items_[1], registry, gameDT, ubgl_items_advect_simple);
}
#endif // UBGL_HAVE_REGISTRY
