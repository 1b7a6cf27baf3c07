// simulation.hpp -- drop-in for the reference's class Simulation
// (te42kyfo/ubootgl simulation.hpp:18-141) whose step() runs on a B200.
//
// The public data members the game reads and writes directly (flag, vx, vy, p,
// vx_accum, vy_accum, vx_current, vy_current, sinks, mg, h, ... -- SURVEY.md
// 8b) are kept as HOST MIRRORS of the device-resident state:
//   step() entry: every mirror the host touched since the last step is
//                 uploaded (flag -> also rebuilds the MG flag pyramid;
//                 accumulators are consumed and zeroed under accum_mutex;
//                 sinks are handed over);
//   step() exit : vx, vy (front and back), p, f, vx_current, vy_current and the
//                 surviving sinks are downloaded.
// setSyncMode(RESIDENT) switches the automatic downloads off: the mirrors are then
// marked stale and pulled from the device lazily, by the first host access to each.
//
// Floating items (simulation.hpp:116-117, called by ubootgl_app.cpp:129-130): inside the
// reference tree -- where components.hpp and the vendored entt are on the include path --
// advectFloatingItems / advectFloatingItemsSimple are declared exactly as in the reference.
// Two definitions fit them: this directory's advect_floating_items.cpp (copies the
// registry view into ubgl_item records and runs ubgl_items_advect[_simple] on the GPU), or
// the reference's own advect_floating_items.cpp unchanged (CPU, on the host mirrors; it
// compiles against this header -- tests/test_dropin_compile.py does exactly that).
#pragma once
#if __has_include("components.hpp") && __has_include("entt/entity/registry.hpp")
#define UBGL_HAVE_REGISTRY 1
#include "components.hpp" // CoItem, CoKinematics, CoKinematicsSimple (components.hpp:6-43)
#include "entt/entity/registry.hpp"
#endif
#include "db2dgrid.hpp"
#include "pressure_solver.hpp"
#include <memory>
#include <mutex>
#include <random>
#include <sstream>
#include <vector>

struct ubgl_sim;

class Simulation {
public:
  enum class BC { INFLOW, OUTFLOW, OUTFLOW_ZERO_PRESSURE, NOSLIP }; // simulation.hpp:69
  enum class SyncMode { MIRROR, RESIDENT };

  Simulation(float pwidth, float mu, int width, int height);         // simulation.hpp:20-30
  Simulation(const Single2DGrid &flagInput, float pwidth, float mu); // simulation.hpp:32-67
  ~Simulation();

  float singlePBC(BC bc, float b);
  float VBCPar(BC bc, float a, float b);
  float VBCPer(BC bc, float a, float b);
  void setPBC();
  void setVBCs();

  float *getFlag() { return flag.data(); }
  float *getP() { return p.data(); }
  float *getR(); // computes the residual field on the device first

  void setGrids(glm::ivec2 c, float val); // simulation.hpp:82-98

  glm::vec2 bilinearVel(glm::vec2 c); // simulation.hpp:100, defined by the items TU
  float psampleFlagNearest(glm::vec2 pc);
  float psampleFlagLinear(glm::vec2 pc);
  glm::vec2 psampleFlagNormal(glm::vec2 pc);

  void diffuse();
  void project();
  void centerP();
  float getDT();
  void advect();
  void applyAccumulatedVelocity();
  void saveCurrentVelocityFields();
  void step(float timestep);
  void interpolateFields();      // simulation.hpp:77, declared and never defined there either
  float diffusion_l2_residual(); // simulation.hpp:105, likewise

#ifdef UBGL_HAVE_REGISTRY
  void advectFloatingItems(entt::registry &registry, float gameDT);       // simulation.hpp:116
  void advectFloatingItemsSimple(entt::registry &registry, float gameDT); // simulation.hpp:117
#endif

  // ---- additions of the drop-in ----
  void setSyncMode(SyncMode m) { mode_ = m; }
  void syncToDevice();  // upload every dirty mirror now
  void syncToHost();    // download vx, vy, p, f, vx_current, vy_current, sinks
  float residualNorm(); // calculateResidualField on the resident fields
  long long kernelLaunches() const;
  ubgl_sim *handle() { return dev_.get(); }

  float pwidth;
  float mu;
  int width, height;
  float dt = 0.0f;

  BC bcWest = BC::INFLOW, bcEast = BC::OUTFLOW_ZERO_PRESSURE, bcNorth = BC::NOSLIP,
     bcSouth = BC::NOSLIP;

  std::stringstream diag;

  DoubleBuffered2DGrid vx, vy;
  Single2DGrid vx_accum, vy_accum;
  Single2DGrid vx_current, vy_current;
  std::mutex accum_mutex;
  Single2DGrid p, f, flag, r;
  MG mg;
  float h;

  std::default_random_engine gen;
  std::uniform_real_distribution<float> disx;
  std::uniform_real_distribution<float> disy;

  std::vector<glm::vec3> sinks;

private:
  void create();
  void runStage(int stage);
  void pushBCs();
  void download(int field, ubgl_host::MirrorStore &m);
  void markStale();
  std::shared_ptr<ubgl_sim> dev_;
  std::shared_ptr<struct ubgl_items> items_[2]; // device item sets of advectFloatingItems / ...Simple
  SyncMode mode_ = SyncMode::MIRROR;
  BC sentBC_[4] = {BC::INFLOW, BC::OUTFLOW_ZERO_PRESSURE, BC::NOSLIP, BC::NOSLIP};
};
