// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// extern "C" handle API around the UNMODIFIED reference translation units
// (pressure_solver.cpp, simulation.cpp under /root/reference), compiled where
// they lie by oracle/Makefile into oracle/_ref/libubgl_ref.so.  Nothing here
// restates any algorithm: every function forwards to the reference's own
// classes / free functions (Single2DGrid db2dgrid.hpp:12, MG
// pressure_solver.hpp:13, Simulation simulation.hpp:18).  Used by tests/ to
// pin the C restatement (oracle/ubgl_oracle.c) and to generate tests/golden/,
// and by bench.py's cpu_baseline / --impl reference legs.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <functional>
#include <iomanip>
#include <iostream>
#include <mutex>
#include <random>
#include <sstream>
#include <string>
#include <vector>
#include <omp.h>

// The MG level pyramid (flagcs, rs, ...) is private in the reference
// (pressure_solver.hpp:64-75); the parity tests need to read the coarse flag
// masks, so open the class up for this wrapper TU only.
#include "components.hpp" // pulls entt + glm before the access override below
#define private public
#include "simulation.hpp" // includes pressure_solver.hpp (which has no include guard)
#undef private

// Free functions defined in pressure_solver.cpp with external linkage but not
// declared in its header (pressure_solver.cpp:118,134,174,183).
void restrict(Single2DGrid &r, Single2DGrid &rc);
void prolongate(Single2DGrid &r, Single2DGrid &rc, Single2DGrid &flagc,
                Single2DGrid &flag);
void correct(Single2DGrid &p, Single2DGrid &e);
void setZeroGradientBC(Single2DGrid &p);

namespace {
Single2DGrid mk(const float *src, int w, int h) {
  Single2DGrid g(w, h);
  if (src)
    std::memcpy(g.data(), src, sizeof(float) * (size_t)w * h);
  return g;
}
void out(const Single2DGrid &g, float *dst) {
  std::memcpy(dst, g.data(), sizeof(float) * (size_t)g.width * g.height);
}
} // namespace

extern "C" {

void ref_set_threads(int n) { omp_set_num_threads(n); }
int ref_max_threads() { return omp_get_max_threads(); }
int ref_num_procs() { return omp_get_num_procs(); }

// ---- pressure_solver.cpp free functions --------------------------------
void ref_rbgs(float *p, const float *f, const float *flag, int w, int h,
              float hh, float alpha, int sweeps) {
  auto P = mk(p, w, h), F = mk(f, w, h), G = mk(flag, w, h);
  for (int i = 0; i < sweeps; i++)
    rbgs(P, F, G, hh, alpha);
  out(P, p);
}

float ref_residual(const float *p, const float *f, const float *flag, float *r,
                   int w, int h, float hh) {
  auto P = mk(p, w, h), F = mk(f, w, h), G = mk(flag, w, h), R = mk(r, w, h);
  float l2 = calculateResidualField(P, F, G, R, hh);
  out(R, r);
  return l2;
}

void ref_restrict(const float *r, int w, int h, float *rc, int wc, int hc) {
  auto R = mk(r, w, h), RC = mk(rc, wc, hc);
  restrict(R, RC);
  out(RC, rc);
}

void ref_prolongate(float *e, int w, int h, const float *ec, const float *flagc,
                    int wc, int hc, const float *flag) {
  auto E = mk(e, w, h), EC = mk(ec, wc, hc), GC = mk(flagc, wc, hc),
       G = mk(flag, w, h);
  prolongate(E, EC, GC, G);
  out(E, e);
}

void ref_correct(float *p, const float *e, int w, int h) {
  auto P = mk(p, w, h), E = mk(e, w, h);
  correct(P, E);
  out(P, p);
}

void ref_zero_gradient_bc(float *p, int w, int h) {
  auto P = mk(p, w, h);
  setZeroGradientBC(P);
  out(P, p);
}

// ---- class MG -------------------------------------------------------------
struct RefMG {
  MG mg;
  Single2DGrid p, f, flag;
};

void *ref_mg_create(int w, int h) {
  auto *m = new RefMG{MG(w, h), Single2DGrid(w, h), Single2DGrid(w, h),
                      Single2DGrid(w, h)};
  m->flag.fill(1.0f);
  return m;
}
void ref_mg_destroy(void *h) { delete (RefMG *)h; }
int ref_mg_levels(void *h) { return ((RefMG *)h)->mg.levels; }
void ref_mg_level_size(void *h, int l, int *w, int *hh) {
  auto &g = ((RefMG *)h)->mg.flagcs[l];
  *w = g.width;
  *hh = g.height;
}
void ref_mg_update_fields(void *h, const float *flag) {
  auto *m = (RefMG *)h;
  std::memcpy(m->flag.data(), flag,
              sizeof(float) * (size_t)m->flag.width * m->flag.height);
  m->mg.updateFields(m->flag);
}
void ref_mg_get_flagc(void *h, int l, float *dst) {
  out(((RefMG *)h)->mg.flagcs[l], dst);
}
// Fields stay resident in the handle so that timing loops do not pay copies.
void ref_mg_set(void *h, const float *p, const float *f, const float *flag) {
  auto *m = (RefMG *)h;
  size_t n = sizeof(float) * (size_t)m->p.width * m->p.height;
  if (p) std::memcpy(m->p.data(), p, n);
  if (f) std::memcpy(m->f.data(), f, n);
  if (flag) std::memcpy(m->flag.data(), flag, n);
}
void ref_mg_get_p(void *h, float *p) { out(((RefMG *)h)->p, p); }
void ref_mg_solve(void *h, float hh, int zero_gradient_bc) {
  auto *m = (RefMG *)h;
  m->mg.solve(m->p, m->f, m->flag, hh, zero_gradient_bc != 0);
}
float ref_mg_residual(void *h, float hh) {
  auto *m = (RefMG *)h;
  Single2DGrid r(m->p.width, m->p.height);
  return calculateResidualField(m->p, m->f, m->flag, r, hh);
}

// ---- class Simulation -------------------------------------------------------
enum {
  REF_FLAG = 0, REF_VX, REF_VY, REF_VXB, REF_VYB, REF_P, REF_F,
  REF_VX_ACCUM, REF_VY_ACCUM, REF_R, REF_VX_CURRENT, REF_VY_CURRENT
};
enum {
  REF_ST_ACCUM = 0, REF_ST_DIFFUSE, REF_ST_ADVECT, REF_ST_SETVBCS,
  REF_ST_PROJECT, REF_ST_SAVE
};

void *ref_sim_create(const float *flag, int w, int h, float pwidth, float mu) {
  auto G = mk(flag, w, h);
  return new Simulation(G, pwidth, mu);
}
void ref_sim_destroy(void *s) { delete (Simulation *)s; }

static float *sim_field(Simulation *s, int field, int *w, int *h) {
  switch (field) {
  case REF_FLAG: *w = s->flag.width; *h = s->flag.height; return s->flag.data();
  case REF_VX: *w = s->vx.width; *h = s->vx.height; return s->vx.data();
  case REF_VY: *w = s->vy.width; *h = s->vy.height; return s->vy.data();
  case REF_VXB: *w = s->vx.width; *h = s->vx.height; return s->vx.back_data();
  case REF_VYB: *w = s->vy.width; *h = s->vy.height; return s->vy.back_data();
  case REF_P: *w = s->p.width; *h = s->p.height; return s->p.data();
  case REF_F: *w = s->f.width; *h = s->f.height; return s->f.data();
  case REF_VX_ACCUM: *w = s->vx_accum.width; *h = s->vx_accum.height; return s->vx_accum.data();
  case REF_VY_ACCUM: *w = s->vy_accum.width; *h = s->vy_accum.height; return s->vy_accum.data();
  case REF_R: *w = s->r.width; *h = s->r.height; return s->r.data();
  case REF_VX_CURRENT: *w = s->vx_current.width; *h = s->vx_current.height; return s->vx_current.data();
  case REF_VY_CURRENT: *w = s->vy_current.width; *h = s->vy_current.height; return s->vy_current.data();
  }
  return nullptr;
}

int ref_sim_get(void *s, int field, float *dst) {
  int w, h;
  float *src = sim_field((Simulation *)s, field, &w, &h);
  if (!src) return -1;
  std::memcpy(dst, src, sizeof(float) * (size_t)w * h);
  return 0;
}
int ref_sim_set(void *s, int field, const float *src) {
  int w, h;
  float *dst = sim_field((Simulation *)s, field, &w, &h);
  if (!dst) return -1;
  std::memcpy(dst, src, sizeof(float) * (size_t)w * h);
  return 0;
}
// ubootgl_app.cpp:111-112: memcpy into sim.flag, then mg.updateFields.
void ref_sim_update_flag(void *s_, const float *flag) {
  auto *s = (Simulation *)s_;
  std::memcpy(s->flag.data(), flag,
              sizeof(float) * (size_t)s->flag.width * s->flag.height);
  s->mg.updateFields(s->flag);
}
void ref_sim_set_bc(void *s_, int west, int east, int north, int south) {
  auto *s = (Simulation *)s_;
  s->bcWest = (Simulation::BC)west;
  s->bcEast = (Simulation::BC)east;
  s->bcNorth = (Simulation::BC)north;
  s->bcSouth = (Simulation::BC)south;
}
void ref_sim_add_sink(void *s, float x, float y, float z) {
  ((Simulation *)s)->sinks.push_back(glm::vec3(x, y, z));
}
int ref_sim_num_sinks(void *s) { return (int)((Simulation *)s)->sinks.size(); }
void ref_sim_get_sinks(void *s_, float *xyz) {
  auto *s = (Simulation *)s_;
  for (size_t i = 0; i < s->sinks.size(); i++) {
    xyz[3 * i + 0] = s->sinks[i].x;
    xyz[3 * i + 1] = s->sinks[i].y;
    xyz[3 * i + 2] = s->sinks[i].z;
  }
}
float ref_sim_h(void *s) { return ((Simulation *)s)->h; }
void ref_sim_step(void *s, float dt) { ((Simulation *)s)->step(dt); }
void ref_sim_stage(void *s_, int stage, float dt) {
  auto *s = (Simulation *)s_;
  s->dt = dt;
  switch (stage) {
  case REF_ST_ACCUM: s->applyAccumulatedVelocity(); break;
  case REF_ST_DIFFUSE: s->diffuse(); break;
  case REF_ST_ADVECT: s->advect(); break;
  case REF_ST_SETVBCS: s->setVBCs(); break;
  case REF_ST_PROJECT: s->project(); break;
  case REF_ST_SAVE: s->saveCurrentVelocityFields(); break;
  }
}
int ref_sim_mg_levels(void *s) { return ((Simulation *)s)->mg.levels; }
void ref_sim_mg_level_size(void *s, int l, int *w, int *h) {
  auto &g = ((Simulation *)s)->mg.flagcs[l];
  *w = g.width;
  *h = g.height;
}
void ref_sim_mg_get_flagc(void *s, int l, float *dst) {
  out(((Simulation *)s)->mg.flagcs[l], dst);
}

} // extern "C"
