// mg.cuh -- device-resident geometric multigrid (the reference's class MG,
// pressure_solver.hpp:13-76, and the free operators of pressure_solver.cpp).
#pragma once
#include "common.cuh"
#include <vector>

namespace ubgl {

struct MGLevel {
  int w = 0, h = 0, pitch = 0;
  Grid flagc; // flagcs[l]  (level 0: copy of the flag given to updateFields)
  Grid rc;    // rcs[l]     restricted residual = rhs of this level (l >= 1)
  Grid ec;    // ecs[l]     error / solution of this level            (l >= 1)
  Grid r;     // rs[l]      residual scratch, only materialised by the plain path
  Grid eb;    // ping-pong partner of ec for the fused tile kernels (l >= 1)
  uint8_t *mask = nullptr; // 5-bit stencil mask of flagc (pitch bytes per row), l >= 1
};

class DeviceMG {
public:
  DeviceMG(int W, int H, int device, cudaStream_t stream, LaunchCounter *lc);
  ~DeviceMG();
  DeviceMG(const DeviceMG &) = delete;
  DeviceMG &operator=(const DeviceMG &) = delete;

  int levels() const { return (int)lv.size(); }
  const MGLevel &level(int l) const { return lv[l]; }

  // MG::updateFields (pressure_solver.hpp:34-57); flag0 is a device grid.
  void update_fields(const Grid &flag0);
  bool update_fields_discs(const Grid &flag0, const float *d_xyd, int n, int max_diam); // mg.cu
  // MG::solve -> solveLevel(.., level 0) (pressure_solver.cpp:201-248)
  void solve(const Grid &p, const Grid &f, const Grid &flag, float hh, bool zgbc);

  // individual operators (device grids), used by the plain path and the C ABI
  void rbgs(const Grid &p, const Grid &f, const Grid &flag, float hh, float alpha);
  void zero_gradient_bc(const Grid &p);
  void residual(const Grid &p, const Grid &f, const Grid &flag, const Grid &r,
                float hh, bool want_norm);
  float residual_norm_result(); // syncs the stream
  void restrict_fw(const Grid &r, const Grid &rc);
  void prolongate(const Grid &e, const Grid &ec, const Grid &flagc, const Grid &flag);
  void correct(const Grid &p, const Grid &e);
  void prolongate_correct(const Grid &p, const Grid &ec, const Grid &flagc,
                          const Grid &flag);

  // The fused path needs binary flags; level 0 uses the caller's flag grid
  // (pressure_solver.hpp:59-62), whose mask is rebuilt when it changes.
  void invalidate_mask0() { mask0_src = nullptr; }
  void prepare_mask0(const Grid &flag, bool known_binary = false);
  const uint8_t *mask0_ptr() const { return mask0; }
  bool mask0_is_binary() const { return mask0_src != nullptr && mask0_binary; }

  bool fused = true;
  int cur_level = 0; // MG level the operator launches are attributed to (profile)
  int device;
  cudaStream_t stream;
  LaunchCounter *lc;

private:
  void solve_level(const Grid &p, const Grid &f, const Grid &flag, float hh,
                   int level, bool zgbc);
  void ensure_r(int level);
  void solve_fused(const Grid &p, const Grid &f, const Grid &flag, float hh, bool zgbc);
  uint8_t *mask0 = nullptr;        // stencil mask of the level-0 flag
  const float *mask0_src = nullptr; // flag buffer mask0 was built from
  bool mask0_binary = false;
  Grid scratch0;                   // level-0 ping-pong partner of the caller's p
  int *d_nonbinary = nullptr;
  std::vector<MGLevel> lv;
  double *d_partials = nullptr; // per-block partial sums of r^2
  double *d_norm = nullptr;
  int n_partials = 0;
};

// mg_fused.cu
void launch_mg_pre(const float *p_in, float *p_out, const Grid &f, const uint8_t *mask,
                   const Grid &rc, float hh, bool zgbc, cudaStream_t stream, LaunchCounter *lc,
                   int level, const Rows *rows = nullptr);
void launch_mg_post(const float *p_in, float *p_out, const Grid &f, const uint8_t *mask,
                    const Grid &ec, const uint8_t *maskc, float hh, bool zgbc, cudaStream_t stream,
                    LaunchCounter *lc, int level, const Rows *rows = nullptr,
                    const Rows *crows = nullptr);
void set_tile_variant(int v); // 2: k_mg_run (default), 1: k_mg_tile
int tile_variant();
void launch_mg_smooth5(float *p_out, const Grid &f, const uint8_t *mask, float hh,
                       cudaStream_t stream, LaunchCounter *lc, int level);
// k_mg_tail: levels t..L of the V-cycle in one single-CTA launch (shared-memory resident)
struct TailLevel {
  int w, h, pitch;
  const uint8_t *mask;
};
int mg_tail_first_level(const std::vector<TailLevel> &lv, int t_min); // 0: none fits
void launch_mg_tail(const std::vector<TailLevel> &lv, int t, const float *hh, const float *rhs,
                    float *out, cudaStream_t stream, LaunchCounter *lc);
void launch_make_mask(const Grid &flag, uint8_t *mask, int *d_nonbinary, cudaStream_t stream,
                      LaunchCounter *lc, int level, const Rows *rows = nullptr,
                    const Rows *crows = nullptr);
// MG::updateFields level step on coarse rows [r_lo, r_hi) (mg.cu)
// the mask bytes in the neighbourhoods of edited discs only (mg.cu: update_fields_discs)
void launch_make_mask_discs(const Grid &flag, uint8_t *mask, const float *d_xyd, dim3 grid, int level,
                            cudaStream_t stream, LaunchCounter *lc);
void launch_coarsen_flag(const Grid &fine, const Grid &fc, int r_lo, int r_hi, cudaStream_t stream,
                         LaunchCounter *lc, int level);

Grid alloc_grid(int w, int h, int pitch_floats, bool zero = true);
void free_grid(Grid &g);
// unpadded host <-> pitched device copies on a stream
void upload_grid(const Grid &g, const float *host, int w, int h, cudaStream_t s);
void download_grid(const Grid &g, float *host, int w, int h, cudaStream_t s);
void fill_grid(const Grid &g, float v, cudaStream_t s, LaunchCounter *lc);

} // namespace ubgl
