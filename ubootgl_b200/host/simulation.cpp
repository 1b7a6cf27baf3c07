// simulation.cpp -- host side of the Simulation drop-in.  All field arithmetic
// is in libubgl.so (csrc/sim.cu, csrc/sim_fused.cu, csrc/mg*.cu); this file only
// keeps the host mirrors coherent and restates the O(1) helpers the game calls
// on the host (flag samplers, setGrids).
#include "simulation.hpp"
#include "../../include/ubgl.h"
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>

namespace {
void ck(int rc, const char *what) {
  if (rc != UBGL_OK)
    throw std::runtime_error(std::string(what) + ": libubgl error " + std::to_string(rc) + ": " +
                             ubgl_last_error());
}
} // namespace

Simulation::Simulation(float pwidth_, float mu_, int w, int h_)
    : pwidth(pwidth_), mu(mu_), width(w), height(h_), vx(w - 1, h_), vy(w, h_ - 1),
      vx_accum(w - 1, h_), vy_accum(w, h_ - 1), vx_current(w - 1, h_), vy_current(w, h_ - 1),
      p(w, h_), f(w, h_), flag(w, h_), r(w, h_), h(pwidth_ / (w - 1.0f)), disx(0.0f, 1.0f),
      disy(0.0f, (float)h_ / w) {
  // simulation.hpp:20-30 leaves flag all-zero (all solid) and the BC members
  // uninitialised; here they take the defaults of the other constructor.
  create();
}

Simulation::Simulation(const Single2DGrid &flagInput, float pwidth_, float mu_)
    : pwidth(pwidth_), mu(mu_), width(flagInput.width), height(flagInput.height) {
  vx = DoubleBuffered2DGrid(width - 1, height);
  vy = DoubleBuffered2DGrid(width, height - 1);
  vx_accum = Single2DGrid(width - 1, height);
  vy_accum = Single2DGrid(width, height - 1);
  vx_current = Single2DGrid(width - 1, height);
  vy_current = Single2DGrid(width, height - 1);
  p = Single2DGrid(width, height);
  f = Single2DGrid(width, height);
  r = Single2DGrid(width, height);
  flag = flagInput;
  for (int y = 0; y < vx.height; y++) vx.f(0, y) = vx.b(0, y) = 1.0f; // simulation.hpp:58-60
  h = pwidth / (width - 1.0f);
  disx = std::uniform_real_distribution<float>(0.0f, 1.0f);
  disy = std::uniform_real_distribution<float>(0.0f, (float)height / width);
  create();
}

Simulation::~Simulation() = default;

void Simulation::create() {
  ubgl_sim_t *raw = nullptr;
  ck(ubgl_sim_create(flag.data(), width, height, pwidth, mu, 0, &raw), "Simulation");
  dev_.reset(raw, [](ubgl_sim_t *s) { ubgl_sim_destroy(s); });
  flag.mirror().clean();
  mg = MG::attached(dev_, width, height);
  // lazy pulls of SyncMode::RESIDENT: a stale mirror fetches its field on first access
  auto pull = [this](int field) {
    return [this, field](float *dst) { ck(ubgl_sim_download(dev_.get(), field, dst), "pull"); };
  };
  for (int k = 0; k < 2; k++) {
    vx.store(k).set_pull([this, k](float *dst) {
      ck(ubgl_sim_download(dev_.get(), vx.front_index() == k ? UBGL_VX : UBGL_VXB, dst), "pull");
    });
    vy.store(k).set_pull([this, k](float *dst) {
      ck(ubgl_sim_download(dev_.get(), vy.front_index() == k ? UBGL_VY : UBGL_VYB, dst), "pull");
    });
  }
  p.mirror().set_pull(pull(UBGL_P));
  f.mirror().set_pull(pull(UBGL_F));
  vx_current.mirror().set_pull(pull(UBGL_VX_CURRENT));
  vy_current.mirror().set_pull(pull(UBGL_VY_CURRENT));
  syncToDevice();
}

void Simulation::markStale() {
  for (int k = 0; k < 2; k++) {
    vx.store(k).mark_stale();
    vy.store(k).mark_stale();
  }
  p.mirror().mark_stale();
  f.mirror().mark_stale();
  vx_current.mirror().mark_stale();
  vy_current.mirror().mark_stale();
}

void Simulation::pushBCs() {
  const BC now[4] = {bcWest, bcEast, bcNorth, bcSouth};
  if (std::memcmp(now, sentBC_, sizeof now) != 0) {
    ck(ubgl_sim_set_bc(dev_.get(), (int)bcWest, (int)bcEast, (int)bcNorth, (int)bcSouth), "set_bc");
    std::memcpy(sentBC_, now, sizeof now);
  }
}

void Simulation::syncToDevice() {
  pushBCs();
  auto up = [&](int field, ubgl_host::MirrorStore &m) {
    if (!m.dirty()) return;
    ck(ubgl_sim_upload(dev_.get(), field, m.ro()), "upload");
    m.clean();
  };
  if (flag.mirror().dirty()) {
    // a bare write to sim.flag (memcpy ubootgl_app.cpp:111, setGrids simulation.hpp:85) changes
    // level 0 only; the coarse flags follow at mg.updateFields(flag), as in the reference
    ck(ubgl_sim_upload(dev_.get(), UBGL_FLAG, flag.mirror().ro()), "upload flag");
    flag.mirror().clean();
  }
  up(UBGL_VX, vx.front_mirror());
  up(UBGL_VXB, vx.back_mirror());
  up(UBGL_VY, vy.front_mirror());
  up(UBGL_VYB, vy.back_mirror());
  up(UBGL_P, p.mirror());
  {
    // accumulators: consumed under the same mutex the item advection scatters
    // under (simulation.cpp:377); the host copies are zeroed like :384,:392
    std::lock_guard<std::mutex> lock(accum_mutex);
    auto take = [&](int field, Single2DGrid &g) {
      if (!g.mirror().dirty()) return;
      // += : the device accumulators may already hold what the items kernels scattered
      ck(ubgl_sim_upload_add(dev_.get(), field, g.mirror().ro()), "upload accum");
      float *a = g.mirror().raw();
      for (int y = 1; y < g.height - 1; y++)
        std::memset(a + (size_t)y * g.width + 1, 0, sizeof(float) * (g.width - 2));
      g.mirror().clean();
    };
    take(UBGL_VX_ACCUM, vx_accum);
    take(UBGL_VY_ACCUM, vy_accum);
  }
  static_assert(sizeof(glm::vec3) == 3 * sizeof(float), "sinks are xyz float triples");
  ck(ubgl_sim_set_sinks(dev_.get(), sinks.empty() ? nullptr : &sinks[0].x, (int)sinks.size()),
     "set_sinks");
}

void Simulation::download(int field, ubgl_host::MirrorStore &m) {
  ck(ubgl_sim_download(dev_.get(), field, m.raw()), "download");
  m.clean();
  m.mark_fresh();
}

void Simulation::syncToHost() {
  download(UBGL_VX, vx.front_mirror());
  download(UBGL_VXB, vx.back_mirror());
  download(UBGL_VY, vy.front_mirror());
  download(UBGL_VYB, vy.back_mirror());
  download(UBGL_P, p.mirror());
  download(UBGL_F, f.mirror());
  download(UBGL_VX_CURRENT, vx_current.mirror());
  download(UBGL_VY_CURRENT, vy_current.mirror());
  int n = 0;
  ck(ubgl_sim_get_sinks(dev_.get(), nullptr, 0, &n), "get_sinks");
  sinks.resize(n);
  if (n) ck(ubgl_sim_get_sinks(dev_.get(), &sinks[0].x, n, &n), "get_sinks");
}

// Simulation::step (simulation.cpp:356-374)
void Simulation::step(float timestep) {
  dt = timestep;
  syncToDevice();
  ck(ubgl_sim_step(dev_.get(), timestep), "step");
  if (mode_ == SyncMode::MIRROR) {
    syncToHost();
  } else {
    markStale();
    int n = 0; // the sink list is host-side state in either mode
    ck(ubgl_sim_get_sinks(dev_.get(), nullptr, 0, &n), "get_sinks");
    sinks.resize(n);
    if (n) ck(ubgl_sim_get_sinks(dev_.get(), &sinks[0].x, n, &n), "get_sinks");
  }
  diag.str("");
  diag << "step dt=" << timestep << " on device, " << kernelLaunches() << " kernel launches so far\n";
}

void Simulation::runStage(int stage) {
  syncToDevice();
  ck(ubgl_sim_stage(dev_.get(), stage, dt), "stage");
  if (mode_ == SyncMode::MIRROR) syncToHost();
  else markStale();
}
void Simulation::applyAccumulatedVelocity() { runStage(UBGL_ST_ACCUM); }
void Simulation::diffuse() { runStage(UBGL_ST_DIFFUSE); }
void Simulation::advect() { runStage(UBGL_ST_ADVECT); }
void Simulation::setVBCs() { runStage(UBGL_ST_SETVBCS); }
void Simulation::project() { runStage(UBGL_ST_PROJECT); }
void Simulation::saveCurrentVelocityFields() { runStage(UBGL_ST_SAVE); }

// simulation.cpp:36-45, on the host mirror and on the device copy
void Simulation::setPBC() {
  for (int y = 0; y < height; y++) {
    p(0, y) = singlePBC(bcWest, p(1, y));
    p(width - 1, y) = singlePBC(bcEast, p(width - 2, y));
  }
  for (int x = 0; x < width; x++) {
    p(x, 0) = singlePBC(bcSouth, p(x, 1));
    p(x, height - 1) = singlePBC(bcNorth, p(x, height - 2));
  }
}

float Simulation::singlePBC(BC bc, float b) { return bc == BC::OUTFLOW_ZERO_PRESSURE ? -b : b; }
float Simulation::VBCPar(BC bc, float a, float b) {
  if (bc == BC::INFLOW) return b;
  if (bc == BC::NOSLIP) return 0.0f;
  return std::fmax(a, 0.0f);
}
float Simulation::VBCPer(BC bc, float a, float b) {
  if (bc == BC::INFLOW) return b;
  if (bc == BC::NOSLIP) return -a;
  return std::fmax(a, 0.0f);
}

float *Simulation::getR() {
  residualNorm();
  download(UBGL_R, r.mirror());
  return r.data();
}

float Simulation::residualNorm() {
  syncToDevice();
  float l2 = 0.0f;
  ck(ubgl_sim_residual(dev_.get(), &l2), "residual");
  return l2;
}

long long Simulation::kernelLaunches() const { return ubgl_sim_launch_count(dev_.get()); }

// simulation.hpp:82-98: terrain edits zero the faces and the pressure of a
// cell that turns solid; the mirrors go dirty and are uploaded at the next step
void Simulation::setGrids(glm::ivec2 c, float val) {
  if (c.x < 0 || c.y < 0 || c.x >= width || c.y >= height) return;
  flag(c.x, c.y) = val;
  if (val != 0) return;
  if (c.x < vx.width) vx(c.x, c.y) = 0.0f;
  if (c.x > 0) vx(c.x - 1, c.y) = 0.0f;
  if (c.y < vy.height) vy(c.x, c.y) = 0.0f;
  if (c.y > 0) vy(c.x, c.y - 1) = 0.0f;
  p(c.x, c.y) = 0.0f;
}

// simulation.cpp:398-425 -- O(1) host reads of the flag mirror
float Simulation::psampleFlagLinear(glm::vec2 pc) {
  const float cell = pwidth / flag.width;
  const float cx = pc.x / cell - 0.5f, cy = pc.y / cell - 0.5f;
  const Single2DGrid &fl = flag;
  const int ix = std::max(0, std::min(fl.width - 2, (int)cx));
  const int iy = std::max(0, std::min(fl.height - 2, (int)cy));
  const float sx = cx - std::floor(cx), sy = cy - std::floor(cy);
  const float lo = glm::mix(fl(ix, iy), fl(ix + 1, iy), sx);
  const float hi = glm::mix(fl(ix, iy + 1), fl(ix + 1, iy + 1), sx);
  return glm::mix(lo, hi, sy);
}

glm::vec2 Simulation::psampleFlagNormal(glm::vec2 pc) {
  const float nw = psampleFlagLinear(glm::vec2(pc.x - h, pc.y + h));
  const float ne = psampleFlagLinear(glm::vec2(pc.x + h, pc.y + h));
  const float sw = psampleFlagLinear(glm::vec2(pc.x - h, pc.y - h));
  const float se = psampleFlagLinear(glm::vec2(pc.x + h, pc.y - h));
  return glm::vec2(ne + se - nw - sw, nw + ne - sw - se);
}

float Simulation::psampleFlagNearest(glm::vec2 pc) {
  const float cell = pwidth / flag.width;
  const Single2DGrid &fl = flag;
  return fl((int)(pc.x / cell), (int)(pc.y / cell));
}

// simulation.cpp:210-224 (unused by step): mean-free pressure on the interior
void Simulation::centerP() {
  double sum = 0.0;
  const Single2DGrid &pc = p;
  for (int y = 1; y < height - 1; y++)
    for (int x = 1; x < width - 1; x++) sum += pc(x, y);
  const float shift = (float)(sum / width / height);
  for (int y = 1; y < height - 1; y++)
    for (int x = 1; x < width - 1; x++) p(x, y) -= shift;
}

// simulation.cpp:226-238 (unused by step)
float Simulation::getDT() {
  float vmax = 1.0e-7f;
  const DoubleBuffered2DGrid &ux = vx, &uy = vy;
  for (int y = 1; y < height - 1; y++)
    for (int x = 1; x < width - 1; x++) {
      if (x < ux.width) vmax = std::fmax(vmax, ux(x, y));
      if (y < uy.height) vmax = std::fmax(vmax, uy(x, y));
    }
  const float rDT = pwidth / (width - 1.0f) / vmax * 2.5f;
  diag << "SET_DT: Vmax=" << vmax << ", dt=" << rDT << "\n";
  return rDT;
}
