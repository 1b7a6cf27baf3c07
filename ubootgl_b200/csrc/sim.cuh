// sim.cuh -- device-resident state of the reference's class Simulation
// (simulation.hpp:18-141) and its step() pipeline (simulation.cpp:356-374).
#pragma once
#include "common.cuh"
#include "mg.cuh"
#include <memory>
#include <vector>

namespace ubgl {

enum { F_FLAG = 0, F_VX, F_VY, F_VXB, F_VYB, F_P, F_F, F_VX_ACCUM, F_VY_ACCUM, F_R,
       F_VX_CURRENT, F_VY_CURRENT, F_COUNT };
enum { ST_ACCUM = 0, ST_DIFFUSE, ST_ADVECT, ST_SETVBCS, ST_PROJECT, ST_SAVE, ST_COUNT };

struct Sink {
  float x, y, z;
};

class DeviceSim {
public:
  DeviceSim(const float *flag, int W, int H, float pwidth, float mu, int device);
  ~DeviceSim();
  DeviceSim(const DeviceSim &) = delete;
  DeviceSim &operator=(const DeviceSim &) = delete;

  Grid field(int id); // current device grid of a public member (front/back resolved)
  void field_size(int id, int *w, int *h) const;
  void upload(int id, const float *host);
  void download(int id, float *host);
  void update_flag(const float *host_flag);
  void flag_changed(bool pyramid);

  void stage(int st, float dt);
  void step(float dt);
  float residual();
  void mg_solve(int cycles);
  void mg_solve_ex(float hh, bool zgbc, int cycles);
  void sync();

  // stages (plain path)
  void apply_accum();
  void diffuse();
  void advect();
  void set_vbcs();
  void project();
  void save_current();

  int W, H, pitch;
  float pwidth, mu, h, dt = 0.0f;
  int bcW, bcE, bcN, bcS;
  int vcycles = 2;
  bool fused = true, use_graph = true, timing = false;
  std::vector<Sink> sinks;
  int device;
  cudaStream_t stream = nullptr;
  LaunchCounter lc;
  std::unique_ptr<DeviceMG> mg;
  float stage_ms[ST_COUNT] = {0, 0, 0, 0, 0, 0};

private:
  // Three buffers per velocity component in the roles front / back
  // (db2dgrid.hpp:52-109) / vx_current (simulation.hpp:130).  swap() exchanges
  // front and back; the fused diffuse writes its result into the buffer that
  // holds the (dead until save) vx_current role and rotates the roles.
  Grid vxb[3], vyb[3];
  int ixf = 0, ixb = 1, ixc = 2, iyf = 0, iyb = 1, iyc = 2;
  Grid vx_accum, vy_accum, p, f, flag, r;
  void project_sinks();
  // sim_fused.cu
  void fused_prestep();
  void fused_borders(bool with_p, bool with_current);
  void fused_divergence();
  void fused_gradient_save();
  float *d_sinks = nullptr; // ix, iy, z triples stamped into f
  int cap_sinks = 0;
  bool r_alloc = false;
};

} // namespace ubgl
