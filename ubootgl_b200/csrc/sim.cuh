// sim.cuh -- device-resident state of the reference's class Simulation
// (simulation.hpp:18-141) and its step() pipeline (simulation.cpp:356-374).
#pragma once
#include "common.cuh"
#include "mg.cuh"
#include <map>
#include <memory>
#include <vector>

namespace ubgl {

enum { F_FLAG = 0, F_VX, F_VY, F_VXB, F_VYB, F_P, F_F, F_VX_ACCUM, F_VY_ACCUM, F_R,
       F_VX_CURRENT, F_VY_CURRENT, F_COUNT };
enum { ST_ACCUM = 0, ST_DIFFUSE, ST_ADVECT, ST_SETVBCS, ST_PROJECT, ST_SAVE, ST_COUNT };

// ---- kernel argument blocks and launchers shared by DeviceSim and the slab
// ---- driver (definitions in sim_fused.cu / sim.cu) ----
struct PrestepArgs {
  const float *A;   // front buffer (input)
  const float *K;   // buffer whose BORDER cells are the "kept" values of the
                    // first setVBCs after pass 1 (vx: the old back buffer)
  float *acc;       // accumulator (read; zeroed later by k_divergence4)
  float *B;         // old back buffer: receives the pass-1 result (interior)
  float *Cout;      // receives the pass-2 result (interior) + border values
  const uint8_t *mask;
  int gw, gh;       // size of this staggered grid
  int H;            // rows of the cell grid (mask)
  int pitch;
  float a, rden;
  int bcLo, bcHi;   // BC of the column sides (W, E)
  int bcS, bcN;
  int st_lo, st_hi, own_lo, own_hi; // row slab (cell-grid rows), see common.cuh Rows
  // frame mode (sim_fused.cu): a 1-D grid over the tiles the register-run kernel does not take --
  // tile rows by < f_byl and by >= f_byf whole, in between only bx == 0 and bx >= f_bxf
  int frame, f_nbx, f_byl, f_byf, f_bxf;
};


struct BorderArgs {
  Grid xf, xb, yf, yb; // velocity front / back buffers (both are written)
  Grid xc, yc;         // optional third copy (vx_current / vy_current), d == nullptr: none
  Grid p;              // optional: setPBC on p, d == nullptr: none
  int bcW, bcE, bcN, bcS;
  int y_lo, y_hi;   // rows whose W/E border cells are set here
  int do_s, do_n;   // this GPU owns the bottom / top border row
};


void launch_prestep(int comp, const PrestepArgs &g, cudaStream_t stream, LaunchCounter *lc);
void launch_borders(const BorderArgs &g, cudaStream_t stream, LaunchCounter *lc);
void launch_pbc(const Grid &p, int bcW, int bcE, int bcN, int bcS, int y_lo, int y_hi, bool do_s,
                bool do_n, cudaStream_t stream, LaunchCounter *lc);
void launch_divergence4(const Grid &vx, const Grid &vy, const Grid &f, const Grid &ax, const Grid &ay,
                        float ih, int y_lo, int y_hi, cudaStream_t stream, LaunchCounter *lc);
void launch_gradient_save(const Grid &vx, const Grid &vy, const Grid &p, const uint8_t *mask,
                          const Grid &cx, const Grid &cy, float ih, int y_lo, int y_hi,
                          cudaStream_t stream, LaunchCounter *lc);
// advect (sim.cu): faces of rows [y_lo, y_hi).  Slab mode (peers != nullptr): tap
// rows outside the locally stored [st_lo, st_hi) are loaded from the neighbours'
// front buffers over NVLink; rows outside [peer_lo, peer_hi) raise *err.
struct AdvectPeers {
  int st_lo, st_hi, peer_lo, peer_hi;
  const float *vx_lo, *vx_hi, *vy_lo, *vy_hi; // neighbours' vx / vy fronts (virtual row 0); null at the ends
  int *err;
};
// mask != nullptr (binary flags): the merged packed-fp32 kernel k_advect_xy may be used; with
// ax / ay it also zeroes the accumulator interiors of those rows and the call returns true.
enum { ADV_ZEROED = 1, ADV_DIV = 2 };
int launch_advect(const Grid &vx, const Grid &vy, const Grid &vxb, const Grid &vyb, const Grid &flag,
                  float half, float full, int y_lo, int y_hi, const AdvectPeers *peers,
                  cudaStream_t stream, LaunchCounter *lc, const uint8_t *mask = nullptr,
                  float *ax = nullptr, float *ay = nullptr, float *fdiv = nullptr, float ih = 0.0f);
// the cells the advect epilogue leaves out (CTA edges, the ring next to the BC faces), after setVBCs
void launch_divergence_edges(const Grid &vx, const Grid &vy, const Grid &f, float ih, int y_lo, int y_hi,
                             cudaStream_t stream, LaunchCounter *lc);
// sinks (sim.cu): 3x3 stamps restricted to rows [y_lo, y_hi)
void launch_stamp_sinks(const Grid &f, const float *d_sinks, int n, int y_lo, int y_hi,
                        cudaStream_t stream, LaunchCounter *lc);

void launch_pack_rows(const float *src, int pitch, int w, int rows, float *dst, cudaStream_t stream,
                      LaunchCounter *lc);

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
  // vx_current / vy_current are byte copies of the final front buffers
  // (saveCurrentVelocityFields, simulation.cpp:16-19).  The fused step does not write them a
  // second time (8 B/cell of HBM stores): after it the *_current role ALIASES the front buffer
  // (cur_alias) and field() resolves it so; will_write(id) -- called by everything that is about
  // to change a front or *_current buffer outside step() -- makes the real copy first.
  void will_write(int id);
  bool lazy_current = true, cur_alias = false;
  void field_size(int id, int *w, int *h) const;
  void upload(int id, const float *host);
  void upload_add(int id, const float *host); // field += host grid
  void download(int id, float *host);
  void update_flag(const float *host_flag);
  void flag_changed(bool pyramid, bool binary_edit = false);
  // a device-side edit of 0.0 / 1.0 values inside the boxes of n discs (cx, cy, diam): patch the flag
  // pyramid and the stencil masks around them instead of rebuilding every level
  void flag_edited_discs(const float *d_xyd, int n, int max_diam);

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
  int advect_impl(bool fused_step); // ADV_* flags of launch_advect
  void set_vbcs();
  void project();
  void save_current();

  int W, H, pitch;
  float pwidth, mu, h, dt = 0.0f;
  int bcW, bcE, bcN, bcS;
  int vcycles = 2;
  // tolerance mode of the pressure solves (tol <= 0: the reference's fixed count)
  float tol = 0.0f, stag = 0.9f, fnorm = 0.0f;
  int max_cycles = 20, cycles_done = 0;
  std::vector<float> res_hist; // ||r|| before the first and after every cycle (tolerance mode)
  bool fused = true, use_graph = true, timing = false;
  std::vector<Sink> sinks;
  int device;
  cudaStream_t stream = nullptr;
  LaunchCounter lc;
  std::unique_ptr<DeviceMG> mg;
  float stage_ms[ST_COUNT] = {0, 0, 0, 0, 0, 0};
  // co-located velocity texture of interp_shader.cs (next.cu), allocated on first use
  float *d_vxy = nullptr, *d_mag = nullptr;
  // pinned staging for small per-step host lists (see sim.cu)
  float *stage_host(size_t nfloats);
  void stage_done();
  float *h_stage = nullptr;
  size_t cap_stage = 0;
  cudaEvent_t ev_stage = nullptr;
  // D->H of a large field for the host mirrors: rows packed to the reference's unpadded
  // layout on the device, then ONE contiguous DMA (a pitched 2-D copy of 8191-float rows
  // runs at 40 GB/s on this box, a contiguous one at 45 GB/s; tools/pcie_probe.py)
  float *packed(const Grid &g); // stream-ordered; the buffer is reused by the next call
  // CUDA graphs of the fused step for small grids (UBGL_OPT_GRAPH, default on): below ~4 M
  // cells a step is ~26 dependent launches of ~10 us each and the gaps between them count.
  // Part A = prestep .. divergence (depends on dt and on the buffer-role rotation), part B =
  // V-cycles .. save (rotation only); the sink stamps between them stay ordinary launches.
  // A graph is captured from the stream the third time its key is seen and replayed after.
  struct StepGraph {
    cudaGraphExec_t exec = nullptr;
    long long launches = 0;
    float dt = 0.0f;
    int seen = 0;
    int post[6] = {0, 0, 0, 0, 0, 0}; // buffer roles after the part
  };
  void drop_graphs();
  float *d_pack = nullptr;
  size_t cap_pack = 0;
  // ubgl_sim_step_host_pipelined (capi.cu): copy streams, the staged (packed) outputs of the last
  // step and the events that order upload -> step -> pack -> download across calls
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaEvent_t ev_up = nullptr, ev_pack = nullptr, ev_down = nullptr, ev_main = nullptr;
  float *d_out = nullptr;
  bool pipe_pending = false;

private:
  // Three buffers per velocity component in the roles front / back
  // (db2dgrid.hpp:52-109) / vx_current (simulation.hpp:130).  swap() exchanges
  // front and back; the fused diffuse writes its result into the buffer that
  // holds the (dead until save) vx_current role and rotates the roles.
  Grid vxb[3], vyb[3];
  int ixf = 0, ixb = 1, ixc = 2, iyf = 0, iyb = 1, iyc = 2;
  Grid vx_accum, vy_accum, p, f, flag, r;
  void project_sinks();
  void solve_cycles();
  double *d_fnorm = nullptr;
  // sim_fused.cu
  void fused_prestep();
  void fused_borders(bool with_p, bool with_current);
  void fused_divergence(bool zero_accum);
  void fused_gradient_save();
  std::map<int, StepGraph> graphs_a, graphs_b; // key: buffer roles before the part
  template <class F> void run_part(std::map<int, StepGraph> &cache, bool dt_dependent, bool graphable, F &&enqueue);
  float *d_sinks = nullptr; // ix, iy, z triples stamped into f
  int cap_sinks = 0;
  bool r_alloc = false;
};

} // namespace ubgl
