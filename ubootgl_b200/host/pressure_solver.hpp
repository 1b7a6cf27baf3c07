// pressure_solver.hpp -- drop-in for the reference's multigrid interface
// (te42kyfo/ubootgl pressure_solver.hpp:3-76) on top of the C ABI (ubgl.h).
// The V-cycle runs in hand-written sm_100a kernels; there is no CPU fallback:
// every call throws std::runtime_error if libubgl reports an error.
#pragma once
#include "db2dgrid.hpp"
#include <memory>

struct ubgl_mg;
struct ubgl_sim;

// pressure_solver.cpp:49-89 (canonical red-black order) / :91-116
void rbgs(Single2DGrid &p, Single2DGrid &f, Single2DGrid &flag, float h, float alpha);
float calculateResidualField(Single2DGrid &p, Single2DGrid &f, Single2DGrid &flag, Single2DGrid &r,
                             float h);

class MG {
public:
  MG() = default;
  MG(int width, int height, int device = 0); // pressure_solver.hpp:16-31
  explicit MG(Single2DGrid &flag, int device = 0) : MG(flag.width, flag.height, device) {
    updateFields(flag);
  }
  // MG is copy-assigned by the reference (simulation.hpp:62); copies share the
  // device state.
  void updateFields(Single2DGrid &flag); // pressure_solver.hpp:34-57
  void solve(Single2DGrid &p, Single2DGrid &f, Single2DGrid &flag, float h,
             bool zeroGradientBC = false); // pressure_solver.hpp:59-62
  // The `mg` member of a Simulation: shares the simulation's device state, so
  // sim.mg.updateFields(sim.flag) (ubootgl_app.cpp:112,296) rebuilds the pyramid
  // the next step() uses.
  static MG attached(std::shared_ptr<ubgl_sim> sim, int width, int height);
  int numLevels() const;
  Single2DGrid coarseFlag(int level) const; // flagcs[level], bit-exact with the reference

private:
  std::shared_ptr<ubgl_mg> dev_;
  std::shared_ptr<ubgl_sim> sim_;
  int width = 0, height = 0;
};
