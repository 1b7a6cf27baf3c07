// pressure_solver.cpp -- host side of the MG drop-in; all arithmetic is in
// libubgl.so (csrc/mg.cu, csrc/mg_fused.cu).
#include "pressure_solver.hpp"
#include "../../include/ubgl.h"
#include <stdexcept>
#include <string>

namespace {
void ck(int rc, const char *what) {
  if (rc != UBGL_OK)
    throw std::runtime_error(std::string(what) + ": libubgl error " + std::to_string(rc) + ": " +
                             ubgl_last_error());
}
} // namespace

void rbgs(Single2DGrid &p, Single2DGrid &f, Single2DGrid &flag, float h, float alpha) {
  ck(ubgl_rbgs(p.data(), f.data(), flag.data(), p.width, p.height, h, alpha, 1), "rbgs");
}

float calculateResidualField(Single2DGrid &p, Single2DGrid &f, Single2DGrid &flag, Single2DGrid &r,
                             float h) {
  float l2 = 0.0f;
  ck(ubgl_residual(p.data(), f.data(), flag.data(), r.data(), p.width, p.height, h, &l2),
     "calculateResidualField");
  return l2;
}

MG::MG(int w, int h, int device) : width(w), height(h) {
  ubgl_mg_t *raw = nullptr;
  ck(ubgl_mg_create(w, h, device, &raw), "MG::MG");
  dev_.reset(raw, [](ubgl_mg_t *m) { ubgl_mg_destroy(m); });
}

MG MG::attached(std::shared_ptr<ubgl_sim> sim, int w, int h) {
  MG m;
  m.sim_ = std::move(sim);
  m.width = w;
  m.height = h;
  return m;
}

void MG::updateFields(Single2DGrid &flag) {
  if (sim_) {
    ck(ubgl_sim_update_flag(sim_.get(), flag.data()), "MG::updateFields");
    flag.mirror().clean();
    return;
  }
  if (!dev_) throw std::runtime_error("MG::updateFields on a default-constructed MG");
  ck(ubgl_mg_update_fields(dev_.get(), flag.data()), "MG::updateFields");
}

void MG::solve(Single2DGrid &p, Single2DGrid &f, Single2DGrid &flag, float h, bool zeroGradientBC) {
  if (sim_) { // solve on the simulation's resident pyramid with the caller's grids
    ck(ubgl_sim_upload(sim_.get(), UBGL_P, p.data()), "MG::solve");
    ck(ubgl_sim_upload(sim_.get(), UBGL_F, f.data()), "MG::solve");
    ck(ubgl_sim_upload(sim_.get(), UBGL_FLAG, flag.data()), "MG::solve");
    ck(ubgl_sim_mg_solve_ex(sim_.get(), h, zeroGradientBC ? 1 : 0, 1), "MG::solve");
    ck(ubgl_sim_download(sim_.get(), UBGL_P, p.data()), "MG::solve");
    return;
  }
  if (!dev_) throw std::runtime_error("MG::solve on a default-constructed MG");
  ck(ubgl_mg_solve_host(dev_.get(), p.data(), f.data(), flag.data(), h, zeroGradientBC ? 1 : 0),
     "MG::solve");
}

int MG::numLevels() const {
  if (sim_) return ubgl_sim_mg_levels(sim_.get());
  return dev_ ? ubgl_mg_levels(dev_.get()) : 0;
}

Single2DGrid MG::coarseFlag(int level) const {
  int w = 0, h = 0;
  if (sim_) {
    ck(ubgl_sim_mg_level_size(sim_.get(), level, &w, &h), "MG::coarseFlag");
    Single2DGrid g(w, h);
    ck(ubgl_sim_mg_get_flagc(sim_.get(), level, g.data()), "MG::coarseFlag");
    return g;
  }
  ck(ubgl_mg_level_size(dev_.get(), level, &w, &h), "MG::coarseFlag");
  Single2DGrid g(w, h);
  ck(ubgl_mg_get_flagc(dev_.get(), level, g.data()), "MG::coarseFlag");
  return g;
}
