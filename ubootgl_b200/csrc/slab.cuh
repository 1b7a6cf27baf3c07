// slab.cuh -- Simulation::step row-slab decomposed over the GPUs of one NVSwitch
// box, one process per GPU.
//
//  * GPU r owns the cell rows [cut_r, cut_{r+1}) of level 0 and the rows
//    [cut_r >> l, cut_{r+1} >> l) of every DISTRIBUTED multigrid level l < ndist
//    (cuts are multiples of 2^ndist, so coarse row yc <-> fine row 2yc never
//    straddles a cut).  Levels >= ndist are REPLICATED: every GPU holds them
//    whole and runs the identical (bitwise deterministic) kernels, after an
//    all-gather of the restricted residual -- the "agglomeration" of the coarse
//    levels without a scatter on the way up.
//  * Every slab array carries `ghost` extra rows per side and is addressed with
//    GLOBAL row indices (Grid::d is the virtual address of row 0), so the
//    single-GPU kernels run unchanged on [own_lo, own_hi).
//  * Halo rows travel by in-kernel PEER STORES over NVLink: k_halo_push copies
//    the boundary rows of up to 8 arrays straight into the neighbours' ghost
//    rows (cudaIpc-mapped arena), fences, and the last block to finish releases
//    a sequence number into the neighbour's signal slot; k_halo_wait acquires it
//    on the consumer's stream.  No host synchronisation inside a step; no
//    collective library on the data path.
#pragma once
#include "common.cuh"
#include "mg.cuh"
#include "sim.cuh"
#include <vector>

namespace ubgl {

struct SlabPlan {
  int W = 0, H = 0, nranks = 1, rank = 0;
  int levels = 0; // multigrid levels (pressure_solver.hpp:16-31)
  int ndist = 0;  // levels [0, ndist) are slab-decomposed, the rest replicated
  int ghost = 16; // ghost rows per side on every distributed level
  std::vector<int> cuts; // nranks + 1 level-0 row cuts
  std::vector<int> lw, lh; // level sizes

  // rows of level l (cell grid) stored / owned by rank r
  Rows rows(int l, int r) const;
  Rows rows(int l) const { return rows(l, rank); }
  int max_stored_rows(int l) const;
};
// Pure host arithmetic (no CUDA): throws ArgError if the grid is too small to
// give every rank a slab.
SlabPlan make_slab_plan(int W, int H, int nranks, int rank);
// relative cost of every level-0 row for the plans made from now on (nullptr: equal rows)
void set_slab_row_weights(const float *w, int H);

constexpr int SLAB_MAXSEG = 16;
constexpr int SLAB_MAXRANKS = 8;

struct HaloSeg {
  const uint4 *src;
  uint4 *dst;
  unsigned long long n16;
};
struct HaloPush {
  HaloSeg seg[SLAB_MAXSEG];
  int nseg;
  unsigned *counter;               // local: blocks finished
  unsigned *sig[SLAB_MAXRANKS];    // peer signal slots to release (nullptr: none)
  int nsig;
  unsigned seq;
  // the same launch then waits for the neighbours' releases of `seq` (one kernel per
  // exchange instead of a push + a wait launch)
  unsigned *wait[SLAB_MAXRANKS];   // local signal slots to wait on
  int nwait;
  int *err;
};

class SlabSim {
public:
  SlabSim(const float *flag_slab, int W, int H, float pwidth, float mu, int device, int rank,
          int nranks);
  ~SlabSim();
  SlabSim(const SlabSim &) = delete;
  SlabSim &operator=(const SlabSim &) = delete;

  // bootstrap: exchange ipc_export() blobs between the processes, then connect()
  void ipc_export(void *blob64) const;
  void connect(const void *blobs64); // nranks x 64 bytes, indexed by rank
  void finish_setup();               // flag pyramid + masks (needs the peers)

  Grid field(int id);
  void field_rows(int id, int *row_lo, int *nrows, int *w) const; // stored rows of a field
  void upload(int id, const float *host);   // host covers the stored rows, unpadded
  void download(int id, float *host);
  void step(float dt);
  // Simulation::step as a host caller sees it: accumulator mirrors in (stored rows, then their
  // own rows are cleared), vx, vy, p, vx_current, vy_current mirrors out (own rows, written at
  // their place inside arrays that cover the stored rows); any pointer may be null
  struct HostRows {
    float *vx_accum, *vy_accum, *vx, *vy, *p, *vx_current, *vy_current;
  };
  void step_host(float dt, const HostRows &m);
  void sync();
  double residual_sumsq(); // sum of r^2 over the own rows (caller adds the ranks and takes sqrt)

  SlabPlan plan;
  int W, H, pitch;
  float pwidth, mu, h, dt = 0.0f;
  int bcW = 0, bcE = 2, bcN = 3, bcS = 3;
  int vcycles = 2;
  std::vector<Sink> sinks;
  int device;
  cudaStream_t stream = nullptr;
  LaunchCounter lc;
  long long exchanges = 0;
  size_t halo_bytes = 0; // bytes pushed to peers so far

private:
  struct Level {
    int w = 0, h = 0, pitch = 0;
    bool dist = false;
    Rows rows{};
    Grid flagc, rc, ec, eb;
    uint8_t *mask = nullptr; // virtual base like the grids
  };
  // arena
  char *arena = nullptr;
  size_t arena_bytes = 0, arena_off = 0;
  char *peer_arena[SLAB_MAXRANKS] = {};
  bool connected = false;
  char *take(size_t bytes);
  Grid arena_grid(int w, int h, int level_pitch, int level, bool dist);
  uint8_t *arena_mask(int h, int level_pitch, int level, bool dist);

  // control block at the start of the arena
  unsigned *sig = nullptr;     // [SLAB_MAXRANKS] last sequence number received from rank j
  unsigned *counter = nullptr; // k_halo_push block counter
  int *err = nullptr;          // 1: advect left the halo, 2: halo wait timed out
  unsigned seq = 0;

  template <typename T> T *peer_ptr(int r, const T *local_virtual, int level, size_t elem_pitch_bytes) const;

  struct XField {
    const void *base; // virtual base (row 0)
    int level;
    size_t row_bytes;
    int rows_hi_clip; // rows of this array (staggered vy: H-1), -1: level height
  };
  XField xf(const Grid &g, int level) const { return XField{g.d, level, sizeof(float) * (size_t)g.pitch, g.h}; }
  XField xm(const uint8_t *m, int level) const { return XField{m, level, (size_t)lv[level].pitch, lv[level].h}; }
  void exchange(const std::vector<XField> &fields, int depth);
  void allgather(const std::vector<XField> &fields); // own rows of replicated arrays to every peer
  void push_and_wait(HaloPush &a, const std::vector<int> &peers);
  void check_err();

  void update_fields();
  void mg_solve();
  void project_sinks();

  std::vector<Level> lv;
  Grid vxb[3], vyb[3];
  int ixf = 0, ixb = 1, ixc = 2, iyf = 0, iyb = 1, iyc = 2;
  // *_current aliases the front buffers after a step until somebody writes either (DeviceSim::cur_alias)
  bool lazy_current = true, cur_alias = false;
  void will_write(int id);
  Grid vx_accum, vy_accum, p, scratch0, f, flag, r;
  uint8_t *mask0 = nullptr;
  int *d_nonbinary = nullptr;
  float *d_sinks = nullptr;
  int cap_sinks = 0;
  double *d_partials = nullptr, *d_sum = nullptr;
  float *d_pack = nullptr; // own rows of one field, unpadded: the source of contiguous D->H copies
};

// halo kernels (slab.cu)
void launch_halo_push(const HaloPush &a, size_t total16, cudaStream_t stream, LaunchCounter *lc);
void launch_halo_wait(unsigned *const *slots, int n, unsigned seq, int *err, cudaStream_t stream,
                      LaunchCounter *lc);

} // namespace ubgl
