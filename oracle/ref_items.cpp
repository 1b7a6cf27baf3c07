// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// Forwards to the UNMODIFIED reference Simulation::advectFloatingItemsSimple and
// Simulation::advectFloatingItems (advect_floating_items.cpp:148-274 / :16-146, compiled
// where it lies by oracle/Makefile) through an entt registry, to pin the restatements
// orc_items_advect_simple / orc_items_advect (oracle/ubgl_oracle_next.c).  Nothing here
// restates the algorithm.
#include <cstring>
#include <vector>

#include "components.hpp"
#include "simulation.hpp"

namespace {
struct ItemRec { // same layout as orc_item (oracle/ubgl_oracle.h)
  float size[2], pos[2], rotation;
  float mass, vel[2], force[2], angVel, angForce;
  int bumpCount;
};
} // namespace

extern "C" {
// items[0..n) in, advanced items out.  The reference visits its
// view<CoItem, CoKinematicsSimple> in entt's packed order (last created first),
// so the entities are created in REVERSE array order: the view then walks the
// array front to back like orc_items_advect_simple.
void ref_items_advect_simple(void *sim_, void *items_, int n, float game_dt) {
  auto *sim = (Simulation *)sim_;
  auto *items = (ItemRec *)items_;
  entt::registry reg;
  std::vector<entt::entity> ent(n);
  for (int i = n - 1; i >= 0; i--) {
    const ItemRec &r = items[i];
    auto e = reg.create();
    reg.emplace<CoItem>(e, glm::vec2(r.size[0], r.size[1]), glm::vec2(r.pos[0], r.pos[1]), r.rotation);
    auto &k = reg.emplace<CoKinematicsSimple>(e, r.mass, glm::vec2(r.vel[0], r.vel[1]), r.angVel);
    k.force = glm::vec2(r.force[0], r.force[1]);
    k.angForce = r.angForce;
    k.bumpCount = r.bumpCount;
    ent[i] = e;
  }
  sim->advectFloatingItemsSimple(reg, game_dt);
  for (int i = 0; i < n; i++) {
    const auto &it = reg.get<CoItem>(ent[i]);
    const auto &k = reg.get<CoKinematicsSimple>(ent[i]);
    ItemRec &r = items[i];
    r.size[0] = it.size.x; r.size[1] = it.size.y;
    r.pos[0] = it.pos.x; r.pos[1] = it.pos.y;
    r.rotation = it.rotation;
    r.mass = k.mass;
    r.vel[0] = k.vel.x; r.vel[1] = k.vel.y;
    r.force[0] = k.force.x; r.force[1] = k.force.y;
    r.angVel = k.angVel; r.angForce = k.angForce;
    r.bumpCount = k.bumpCount;
  }
}

// The same for Simulation::advectFloatingItems (advect_floating_items.cpp:16-146), whose
// view is <CoItem, CoKinematics> (same fields as CoKinematicsSimple, components.hpp:12-43).
void ref_items_advect(void *sim_, void *items_, int n, float game_dt) {
  auto *sim = (Simulation *)sim_;
  auto *items = (ItemRec *)items_;
  entt::registry reg;
  std::vector<entt::entity> ent(n);
  for (int i = n - 1; i >= 0; i--) {
    const ItemRec &r = items[i];
    auto e = reg.create();
    reg.emplace<CoItem>(e, glm::vec2(r.size[0], r.size[1]), glm::vec2(r.pos[0], r.pos[1]), r.rotation);
    auto &k = reg.emplace<CoKinematics>(e, r.mass, glm::vec2(r.vel[0], r.vel[1]), r.angVel);
    k.force = glm::vec2(r.force[0], r.force[1]);
    k.angForce = r.angForce;
    k.bumpCount = r.bumpCount;
    ent[i] = e;
  }
  sim->advectFloatingItems(reg, game_dt);
  for (int i = 0; i < n; i++) {
    const auto &it = reg.get<CoItem>(ent[i]);
    const auto &k = reg.get<CoKinematics>(ent[i]);
    ItemRec &r = items[i];
    r.size[0] = it.size.x; r.size[1] = it.size.y;
    r.pos[0] = it.pos.x; r.pos[1] = it.pos.y;
    r.rotation = it.rotation;
    r.mass = k.mass;
    r.vel[0] = k.vel.x; r.vel[1] = k.vel.y;
    r.force[0] = k.force.x; r.force[1] = k.force.y;
    r.angVel = k.angVel; r.angForce = k.angForce;
    r.bumpCount = k.bumpCount;
  }
}

// order in which the reference's view visits entities created as above
// (array indices), for the test that checks the ordering assumption
void ref_items_view_order(int n, int *order) {
  entt::registry reg;
  std::vector<entt::entity> ent(n);
  for (int i = n - 1; i >= 0; i--) {
    auto e = reg.create();
    reg.emplace<CoItem>(e, glm::vec2(1, 1), glm::vec2((float)i, 0.0f), 0.0f);
    reg.emplace<CoKinematicsSimple>(e, 1.0f, glm::vec2(0, 0), 0.0f);
    ent[i] = e;
  }
  int k = 0;
  auto view = reg.view<CoItem, CoKinematicsSimple>();
  for (auto e : view) order[k++] = (int)view.get<CoItem>(e).pos.x;
}
}
