// TEST PROGRAM of the C++ drop-in (ubootgl_b200/host): drives
// Simulation::advectFloatingItems / advectFloatingItemsSimple through an entt registry exactly
// as ubootgl_app.cpp:129-130 does, then one Simulation::step, and dumps the results for
// tests/test_dropin.py to compare with the UNMODIFIED reference run on the same input
// (oracle/_ref: ref_items_advect*, oracle/ref_items.cpp).
//
// Built inside a stand-in of the reference tree (tests/dropin/Makefile): the reference's own
// components.hpp and vendored entt on the include path, ubootgl_b200/host/*.hpp in place of
// its simulation.hpp / pressure_solver.hpp / db2dgrid.hpp.
//   usage: itemsdemo in_prefix out_prefix
//   in_prefix.meta : "W H n kind frames game_dt step_dt"   (kind 0: CoKinematics, 1: ...Simple)
//   in_prefix.{flag,vx,vy,p} raw fp32, in_prefix.items raw 52-byte records (ubgl_item layout)
#include "components.hpp"
#include "simulation.hpp"
#include "../../include/ubgl.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#ifndef UBGL_HAVE_REGISTRY
#error "itemsdemo must be built where components.hpp and entt are on the include path"
#endif

static std::vector<char> slurp(const std::string &path) {
  FILE *fp = std::fopen(path.c_str(), "rb");
  if (!fp) { std::fprintf(stderr, "cannot read %s\n", path.c_str()); std::exit(2); }
  std::fseek(fp, 0, SEEK_END);
  long n = std::ftell(fp);
  std::fseek(fp, 0, SEEK_SET);
  std::vector<char> b((size_t)n);
  if (n && std::fread(b.data(), 1, (size_t)n, fp) != (size_t)n) std::exit(2);
  std::fclose(fp);
  return b;
}
static void dump(const std::string &path, const void *d, size_t bytes) {
  FILE *fp = std::fopen(path.c_str(), "wb");
  if (!fp) std::exit(2);
  std::fwrite(d, 1, bytes, fp);
  std::fclose(fp);
}

template <class Kin> static void fill(entt::registry &reg, std::vector<entt::entity> &ent, const ubgl_item *r, int n) {
  // entt walks a view in packed order, last created first: create in reverse so the view
  // visits the array front to back (same convention as oracle/ref_items.cpp)
  for (int i = n - 1; i >= 0; i--) {
    auto e = reg.create();
    reg.emplace<CoItem>(e, glm::vec2(r[i].size[0], r[i].size[1]), glm::vec2(r[i].pos[0], r[i].pos[1]), r[i].rotation);
    auto &k = reg.emplace<Kin>(e, r[i].mass, glm::vec2(r[i].vel[0], r[i].vel[1]), r[i].angVel);
    k.force = glm::vec2(r[i].force[0], r[i].force[1]);
    k.angForce = r[i].angForce;
    k.bumpCount = r[i].bumpCount;
    ent[i] = e;
  }
}
template <class Kin> static void read_back(entt::registry &reg, const std::vector<entt::entity> &ent, ubgl_item *r, int n) {
  for (int i = 0; i < n; i++) {
    const auto &it = reg.get<CoItem>(ent[i]);
    const auto &k = reg.get<Kin>(ent[i]);
    r[i].size[0] = it.size.x; r[i].size[1] = it.size.y;
    r[i].pos[0] = it.pos.x; r[i].pos[1] = it.pos.y;
    r[i].rotation = it.rotation;
    r[i].mass = k.mass;
    r[i].vel[0] = k.vel.x; r[i].vel[1] = k.vel.y;
    r[i].force[0] = k.force.x; r[i].force[1] = k.force.y;
    r[i].angVel = k.angVel; r[i].angForce = k.angForce;
    r[i].bumpCount = k.bumpCount;
  }
}

int main(int argc, char **argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: itemsdemo in_prefix out_prefix\n"); return 1; }
  const std::string in = argv[1], out = argv[2];
  int W, H, n, kind, frames;
  float game_dt, step_dt;
  {
    auto m = slurp(in + ".meta");
    m.push_back(0);
    if (std::sscanf(m.data(), "%d %d %d %d %d %f %f", &W, &H, &n, &kind, &frames, &game_dt, &step_dt) != 7) return 2;
  }
  Single2DGrid flag(W, H);
  { auto b = slurp(in + ".flag"); std::memcpy(flag.data(), b.data(), b.size()); }
  Simulation sim(flag, 0.8f, 0.001f);
  { auto b = slurp(in + ".vx"); std::memcpy(sim.vx.data(), b.data(), b.size()); }   // marks the mirrors dirty:
  { auto b = slurp(in + ".vy"); std::memcpy(sim.vy.data(), b.data(), b.size()); }   // uploaded by the next call
  { auto b = slurp(in + ".p"); std::memcpy(sim.p.data(), b.data(), b.size()); }
  auto ib = slurp(in + ".items");
  std::vector<ubgl_item> rec((size_t)n);
  std::memcpy(rec.data(), ib.data(), sizeof(ubgl_item) * (size_t)n);

  entt::registry reg;
  std::vector<entt::entity> ent((size_t)n);
  if (kind == 0) fill<CoKinematics>(reg, ent, rec.data(), n);
  else fill<CoKinematicsSimple>(reg, ent, rec.data(), n);
  for (int f = 0; f < frames; f++) { // ubootgl_app.cpp:129-130
    if (kind == 0) sim.advectFloatingItems(reg, game_dt);
    else sim.advectFloatingItemsSimple(reg, game_dt);
  }
  if (kind == 0) read_back<CoKinematics>(reg, ent, rec.data(), n);
  else read_back<CoKinematicsSimple>(reg, ent, rec.data(), n);
  dump(out + ".items", rec.data(), sizeof(ubgl_item) * (size_t)n);
  // the reaction forces are in the DEVICE accumulators (nothing was uploaded)
  std::vector<float> ax((size_t)(W - 1) * H), ay((size_t)W * (H - 1));
  if (ubgl_sim_download(sim.handle(), UBGL_VX_ACCUM, ax.data()) || ubgl_sim_download(sim.handle(), UBGL_VY_ACCUM, ay.data())) {
    std::fprintf(stderr, "%s\n", ubgl_last_error());
    return 3;
  }
  dump(out + ".ax", ax.data(), sizeof(float) * ax.size());
  dump(out + ".ay", ay.data(), sizeof(float) * ay.size());
  sim.step(step_dt); // sim_loop.cpp:29: the accumulated forces feed the next fluid step
  const DoubleBuffered2DGrid &vx = sim.vx, &vy = sim.vy;
  const Single2DGrid &p = sim.p;
  dump(out + ".vx", vx.data(), sizeof(float) * (size_t)(W - 1) * H);
  dump(out + ".vy", vy.data(), sizeof(float) * (size_t)W * (H - 1));
  dump(out + ".p", p.data(), sizeof(float) * (size_t)W * H);
  std::printf("itemsdemo %dx%d: %d %s, %d frames, %lld kernel launches\n", W, H, n,
              kind == 0 ? "rigid bodies" : "simple items", frames, sim.kernelLaunches());
  return 0;
}
