// mg.cu -- multigrid V-cycle on the device.  Reference: pressure_solver.cpp /
// pressure_solver.hpp of te42kyfo/ubootgl (file:line cited per kernel).
#include "mg.cuh"
#include "stencils.cuh"

namespace ubgl {

#define LVL cur_level

// ---------------------------------------------------------------------------
// grid helpers
// ---------------------------------------------------------------------------
Grid alloc_grid(int w, int h, int pitch, bool zero) {
  Grid g;
  g.w = w;
  g.h = h;
  g.pitch = pitch;
  UBGL_CUDA(cudaMalloc(&g.d, g.bytes()));
  if (zero) {
    // device memsets are asynchronous and the library's streams are non-blocking: finish it
    // before any stream can touch the array
    UBGL_CUDA(cudaMemset(g.d, 0, g.bytes()));
    UBGL_CUDA(cudaDeviceSynchronize());
  }
  return g;
}
void free_grid(Grid &g) {
  if (g.d) cudaFree(g.d);
  g.d = nullptr;
}
void upload_grid(const Grid &g, const float *host, int w, int h, cudaStream_t s) {
  UBGL_CUDA(cudaMemcpy2DAsync(g.d, sizeof(float) * g.pitch, host, sizeof(float) * w,
                              sizeof(float) * w, h, cudaMemcpyHostToDevice, s));
}
void download_grid(const Grid &g, float *host, int w, int h, cudaStream_t s) {
  UBGL_CUDA(cudaMemcpy2DAsync(host, sizeof(float) * w, g.d, sizeof(float) * g.pitch,
                              sizeof(float) * w, h, cudaMemcpyDeviceToHost, s));
}

__global__ void k_fill(float *d, size_t n, float v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) d[i] = v;
}
void fill_grid(const Grid &g, float v, cudaStream_t s, LaunchCounter *lc) {
  size_t n = (size_t)g.pitch * g.h;
  if (v == 0.0f) {
    UBGL_CUDA(cudaMemsetAsync(g.d, 0, n * sizeof(float), s));
    return;
  }
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  LaunchCounter dummy;
  UBGL_LAUNCH(lc ? lc : &dummy, K_FILL, 0, s, k_fill<<<blocks, 256, 0, s>>>(g.d, n, v));
}

// ---------------------------------------------------------------------------
// plain kernels: one per reference operator
// ---------------------------------------------------------------------------

// rbgs_red_line / rbgs_black_line (pressure_solver.cpp:35-47): color 0 ("red")
// starts each row at x = 1 + (y%2), color 1 ("black") at x = 1 + ((y+1)%2).
__global__ void k_rbgs_half(Grid p, Grid f, Grid flag, float hh, float alpha, int color) {
  int y = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int x = 1 + ((y + color) & 1) + 2 * i;
  if (y >= p.h - 1 || x >= p.w - 1) return;
  size_t c = (size_t)y * p.pitch + x;
  const float *fl = flag.d + (size_t)y * flag.pitch + x;
  float fh2 = fh2_of(f.d[(size_t)y * f.pitch + x], hh);
  p.d[c] = smooth_cell(p.d[c], p.d[c - 1], p.d[c + 1], p.d[c - p.pitch], p.d[c + p.pitch],
                       fl[0], fl[-1], fl[1], fl[-flag.pitch], fl[flag.pitch], fh2, alpha);
}

// setZeroGradientBC (pressure_solver.cpp:183-192): the column loop touches only
// y in [1,h-2] and the row loop only x in [1,w-2], so they are independent.
__global__ void k_zero_gradient_bc(Grid p) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 1 && i < p.h - 1) {
    p.at(0, i) = p.at(1, i);
    p.at(p.w - 1, i) = p.at(p.w - 2, i);
  }
  if (i >= 1 && i < p.w - 1) {
    p.at(i, 0) = p.at(i, 1);
    p.at(i, p.h - 1) = p.at(i, p.h - 2);
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// calculateResidualField (pressure_solver.cpp:91-116).  Writes r on the whole
// grid (0 on the border, as after the reference's r.fill(0.0) :218) and, if
// partials != nullptr, one double partial sum of r^2 per block (warp-shuffle
// reduction; summed by k_finish_norm in a fixed order => deterministic).
__global__ void k_residual(Grid p, Grid f, Grid flag, Grid r, float ihsq, double *partials) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y * blockDim.y + threadIdx.y;
  float rv = 0.0f;
  if (x < p.w && y < p.h) {
    if (x >= 1 && y >= 1 && x < p.w - 1 && y < p.h - 1) {
      size_t c = (size_t)y * p.pitch + x;
      const float *fl = flag.d + (size_t)y * flag.pitch + x;
      rv = residual_cell(p.d[c], p.d[c - 1], p.d[c + 1], p.d[c - p.pitch], p.d[c + p.pitch],
                         fl[0], fl[-1], fl[1], fl[-flag.pitch], fl[flag.pitch],
                         f.d[(size_t)y * f.pitch + x], ihsq);
    }
    r.d[(size_t)y * r.pitch + x] = rv;
  }
  if (partials) {
    __shared__ double wsum[32];
    double s = warp_sum((double)rv * (double)rv);
    int tid = threadIdx.y * blockDim.x + threadIdx.x;
    int nw = (blockDim.x * blockDim.y + 31) / 32;
    if ((tid & 31) == 0) wsum[tid >> 5] = s;
    __syncthreads();
    if (tid < 32) {
      double t = tid < nw ? wsum[tid] : 0.0;
      t = warp_sum(t);
      if (tid == 0) partials[blockIdx.y * gridDim.x + blockIdx.x] = t;
    }
  }
}

__global__ void k_finish_norm(const double *partials, int n, double *out) {
  __shared__ double wsum[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += partials[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = threadIdx.x < (blockDim.x >> 5) ? wsum[threadIdx.x] : 0.0;
    t = warp_sum(t);
    if (threadIdx.x == 0) *out = sqrt(t);
  }
}

// restrict (pressure_solver.cpp:118-132) + the rc.fill(0.0) before it (:222):
// the coarse border is written as 0.
__global__ void k_restrict(Grid r, Grid rc) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= rc.w || y >= rc.h) return;
  float v = 0.0f;
  if (x >= 1 && y >= 1 && x < rc.w - 1 && y < rc.h - 1) {
    const float *a = r.d + (size_t)(2 * y - 1) * r.pitch + 2 * x;
    const float *b = a + r.pitch, *c = b + r.pitch;
    v = fw9(a[-1], a[0], a[1], b[-1], b[0], b[1], c[-1], c[0], c[1]);
  }
  rc.d[(size_t)y * rc.pitch + x] = v;
}

// MG::updateFields level step (pressure_solver.hpp:36-55): threshold of the
// full-weighted fine flag at 0.2 (double compare); border cells stay 1.0.
__global__ void k_coarsen_flag(Grid fine, Grid fc, int r_lo, int r_hi) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = r_lo + blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= fc.w || y >= r_hi) return;
  float v = 1.0f;
  if (x >= 1 && y >= 1 && x < fc.w - 1 && y < fc.h - 1) {
    const float *a = fine.d + (size_t)(2 * y - 1) * fine.pitch + 2 * x;
    const float *b = a + fine.pitch, *c = b + fine.pitch;
    float s = fw9(a[-1], a[0], a[1], b[-1], b[0], b[1], c[-1], c[0], c[1]);
    v = ((double)s > 0.2) ? 1.0f : 0.0f;
  }
  fc.d[(size_t)y * fc.pitch + x] = v;
}

// ---------------------------------------------------------------------------
// The same update restricted to the neighbourhoods of a list of edited discs (Terrain::drawCircle
// craters, terrain.cpp:213-234; one (cx, cy, diam) triple each): a crater touches (2 diam + 1)^2
// level-0 cells, the whole-field rebuild above reads and writes every level.  A level-l cell can
// change only if its 3 x 3 fine stencil meets the dirty rectangle of level l-1, and a mask byte only
// if the cell or one of its four neighbours changed: with the level-0 rectangle [lo, hi] per axis
// both sets lie inside [(lo >> l) - 2, (hi >> l) + 3].  blockIdx.z = disc; cells outside the
// rectangles keep their (unchanged) values, cells inside are recomputed to what k_coarsen_flag /
// k_make_mask would write; overlapping rectangles write the same values twice.
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool disc_rect_cell(const float *xyd, int level, int w, int h, int &x, int &y) {
  const int c = blockIdx.z;
  const int d = (int)xyd[3 * c + 2];
  const int lox = (int)xyd[3 * c] - d - 1, loy = (int)xyd[3 * c + 1] - d - 1, side = ((2 * d + 3) >> level) + 7;
  const int tx = blockIdx.x * blockDim.x + threadIdx.x, ty = blockIdx.y * blockDim.y + threadIdx.y;
  x = (lox >> level) - 2 + tx;
  y = (loy >> level) - 2 + ty;
  return tx < side && ty < side && x >= 0 && y >= 0 && x < w && y < h;
}
// level 0: the caller's flag field into the solver's own copy (MG::updateFields' flagcs[0] = flag)
__global__ void k_copy_flag_rects(Grid flag, Grid fc, const float *xyd) {
  ubgl_pdl_prologue();
  int x, y;
  if (!disc_rect_cell(xyd, 0, fc.w, fc.h, x, y)) return;
  fc.at(x, y) = flag.at(x, y);
}
__global__ void k_coarsen_flag_rects(Grid fine, Grid fc, const float *xyd, int level) {
  ubgl_pdl_prologue();
  int x, y;
  if (!disc_rect_cell(xyd, level, fc.w, fc.h, x, y)) return;
  if (x < 1 || y < 1 || x >= fc.w - 1 || y >= fc.h - 1) return; // border cells stay 1.0
  const float *a = fine.d + (size_t)(2 * y - 1) * fine.pitch + 2 * x;
  const float *b = a + fine.pitch, *c = b + fine.pitch;
  const float s = fw9(a[-1], a[0], a[1], b[-1], b[0], b[1], c[-1], c[0], c[1]);
  fc.d[(size_t)y * fc.pitch + x] = ((double)s > 0.2) ? 1.0f : 0.0f;
}

// prolongate (pressure_solver.cpp:134-172): e on the whole fine grid.
__global__ void k_prolongate(Grid e, Grid ec, Grid flagc, Grid flag) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= e.w || y >= e.h) return;
  e.d[(size_t)y * e.pitch + x] =
      prolong_cell(ec.d, flagc.d, ec.pitch, flag.d[(size_t)y * flag.pitch + x], x, y, e.w, e.h);
}

// correct (pressure_solver.cpp:174-181)
__global__ void k_correct(Grid p, Grid e) {
  int x = 1 + blockIdx.x * blockDim.x + threadIdx.x;
  int y = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= p.w - 1 || y >= p.h - 1) return;
  size_t c = (size_t)y * p.pitch + x;
  p.d[c] = __fadd_rn(p.d[c], e.d[(size_t)y * e.pitch + x]);
}

// prolongate + correct in one pass (e never materialised)
__global__ void k_prolongate_correct(Grid p, Grid ec, Grid flagc, Grid flag) {
  int x = 1 + blockIdx.x * blockDim.x + threadIdx.x;
  int y = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= p.w - 1 || y >= p.h - 1) return;
  size_t c = (size_t)y * p.pitch + x;
  float e = prolong_cell(ec.d, flagc.d, ec.pitch, flag.d[(size_t)y * flag.pitch + x], x, y,
                         p.w, p.h);
  p.d[c] = __fadd_rn(p.d[c], e);
}

// ---------------------------------------------------------------------------
// DeviceMG
// ---------------------------------------------------------------------------
static inline dim3 blk2d() { return dim3(32, 8); }
static inline dim3 grd2d(int w, int h) { return dim3(ceil_div(w, 32), ceil_div(h, 8)); }

void launch_coarsen_flag(const Grid &fine, const Grid &fc, int r_lo, int r_hi, cudaStream_t stream,
                         LaunchCounter *lc, int level) {
  UBGL_LAUNCH(lc, K_COARSEN, level, stream, k_coarsen_flag<<<grd2d(fc.w, r_hi - r_lo), blk2d(), 0, stream>>>(fine, fc, r_lo, r_hi));
}

DeviceMG::DeviceMG(int W, int H, int device_, cudaStream_t stream_, LaunchCounter *lc_)
    : device(device_), stream(stream_), lc(lc_) {
  // MG::MG(int,int) pressure_solver.hpp:16-31
  int cw = W, ch = H;
  while (cw > 3 && ch > 3) {
    MGLevel L;
    L.w = cw;
    L.h = ch;
    L.pitch = round_up(cw, 32);
    lv.push_back(L);
    cw /= 2;
    ch /= 2;
  }
  UBGL_REQUIRE(lv.size() >= 2, "MG needs at least two levels (W,H >= 8)");
  for (size_t l = 0; l < lv.size(); l++) {
    MGLevel &L = lv[l];
    L.flagc = alloc_grid(L.w, L.h, L.pitch, false);
    fill_grid(L.flagc, 1.0f, stream, nullptr);
    if (l >= 1 && l + 1 < lv.size()) { // level levels-1 is never visited (:203)
      L.rc = alloc_grid(L.w, L.h, L.pitch);
      L.ec = alloc_grid(L.w, L.h, L.pitch);
      L.eb = alloc_grid(L.w, L.h, L.pitch);
      UBGL_CUDA(cudaMalloc(&L.mask, (size_t)L.pitch * L.h));
    }
  }
  UBGL_CUDA(cudaMalloc(&mask0, (size_t)lv[0].pitch * lv[0].h));
  UBGL_CUDA(cudaMalloc(&d_nonbinary, sizeof(int)));
  UBGL_CUDA(cudaMemset(d_nonbinary, 0, sizeof(int)));
  UBGL_CUDA(cudaDeviceSynchronize());
  for (size_t l = 1; l + 1 < lv.size(); l++) // all-ones coarse flags of MG(int,int)
    launch_make_mask(lv[l].flagc, lv[l].mask, d_nonbinary, stream, lc, (int)l);
  dim3 g = grd2d(W, H);
  n_partials = g.x * g.y;
  UBGL_CUDA(cudaMalloc(&d_partials, sizeof(double) * n_partials));
  UBGL_CUDA(cudaMalloc(&d_norm, sizeof(double)));
  UBGL_CUDA(cudaStreamSynchronize(stream));
}

DeviceMG::~DeviceMG() {
  for (auto &L : lv) {
    free_grid(L.flagc);
    free_grid(L.rc);
    free_grid(L.ec);
    free_grid(L.r);
    free_grid(L.eb);
    if (L.mask) cudaFree(L.mask);
  }
  free_grid(scratch0);
  if (mask0) cudaFree(mask0);
  if (d_nonbinary) cudaFree(d_nonbinary);
  if (d_partials) cudaFree(d_partials);
  if (d_norm) cudaFree(d_norm);
}

void DeviceMG::ensure_r(int level) {
  MGLevel &L = lv[level];
  if (!L.r.d) L.r = alloc_grid(L.w, L.h, L.pitch);
}

void DeviceMG::update_fields(const Grid &flag0) {
  UBGL_REQUIRE(flag0.w == lv[0].w && flag0.h == lv[0].h, "updateFields: flag size mismatch");
  if (flag0.d != lv[0].flagc.d)
    UBGL_CUDA(cudaMemcpy2DAsync(lv[0].flagc.d, sizeof(float) * lv[0].pitch, flag0.d,
                                sizeof(float) * flag0.pitch, sizeof(float) * flag0.w, flag0.h,
                                cudaMemcpyDeviceToDevice, stream));
  for (size_t l = 1; l < lv.size(); l++) {
    launch_coarsen_flag(lv[l - 1].flagc, lv[l].flagc, 0, lv[l].h, stream, lc, (int)l);
    if (l + 1 < lv.size())
      launch_make_mask(lv[l].flagc, lv[l].mask, d_nonbinary, stream, lc, (int)l);
  }
}

// update_fields + the level-0 mask for an edit that wrote only 0.0 / 1.0 inside the boxes of n discs
// (device list of (cx, cy, diam), diam <= max_diam) into a flag field whose masks are valid and
// binary: ~2 small launches per level instead of a pass over every level.  Returns false (nothing
// done) when the preconditions do not hold; the caller then rebuilds everything.
bool DeviceMG::update_fields_discs(const Grid &flag0, const float *d_xyd, int n, int max_diam) {
  if (n <= 0 || mask0_src != flag0.d || !mask0_binary) return false;
  if (flag0.w != lv[0].w || flag0.h != lv[0].h || flag0.pitch != lv[0].pitch) return false;
  auto grid = [&](int level) {
    const int side = ((2 * max_diam + 3) >> level) + 7;
    return dim3(ceil_div(side, 32), ceil_div(side, 8), n);
  };
  if (flag0.d != lv[0].flagc.d)
    UBGL_LAUNCH(lc, K_COARSEN, 0, stream, launch_k(k_copy_flag_rects, grid(0), blk2d(), 0, stream, flag0, lv[0].flagc, d_xyd));
  launch_make_mask_discs(flag0, mask0, d_xyd, grid(0), 0, stream, lc);
  for (size_t l = 1; l < lv.size(); l++) {
    UBGL_LAUNCH(lc, K_COARSEN, (int)l, stream,
                launch_k(k_coarsen_flag_rects, grid((int)l), blk2d(), 0, stream, lv[l - 1].flagc, lv[l].flagc, d_xyd, (int)l));
    if (l + 1 < lv.size()) launch_make_mask_discs(lv[l].flagc, lv[l].mask, d_xyd, grid((int)l), (int)l, stream, lc);
  }
  return true;
}

// (Re)build the level-0 stencil mask from the flag grid the caller solves with.
// Synchronises the stream (reads back the "non-binary flag seen" bit), so the
// owners call it when the flag changes, not inside the step.
void DeviceMG::prepare_mask0(const Grid &flag, bool known_binary) {
  if (mask0_src == flag.d) return;
  UBGL_REQUIRE(flag.w == lv[0].w && flag.h == lv[0].h && flag.pitch == lv[0].pitch,
               "flag grid does not match the MG level-0 layout");
  UBGL_CUDA(cudaMemsetAsync(d_nonbinary, 0, sizeof(int), stream));
  launch_make_mask(flag, mask0, d_nonbinary, stream, lc, 0);
  if (known_binary) {
    // device-side edit that only writes 0.0 / 1.0 into a field that was binary: no need to
    // read the verdict back (a host sync per terrain edit would serialise the frame loop)
    mask0_binary = true;
  } else {
    int nb = 0;
    UBGL_CUDA(cudaMemcpyAsync(&nb, d_nonbinary, sizeof(int), cudaMemcpyDeviceToHost, stream));
    UBGL_CUDA(cudaStreamSynchronize(stream));
    mask0_binary = (nb == 0);
  }
  mask0_src = flag.d;
}

void DeviceMG::rbgs(const Grid &p, const Grid &f, const Grid &flag, float hh, float alpha) {
  dim3 b = blk2d();
  dim3 g(ceil_div((p.w - 2 + 1) / 2, b.x), ceil_div(p.h - 2, b.y));
  for (int color = 0; color < 2; color++) {
    UBGL_LAUNCH(lc, K_RBGS, LVL, stream, k_rbgs_half<<<g, b, 0, stream>>>(p, f, flag, hh, alpha, color));
  }
}

void DeviceMG::zero_gradient_bc(const Grid &p) {
  int n = p.w > p.h ? p.w : p.h;
  UBGL_LAUNCH(lc, K_ZGBC, LVL, stream, k_zero_gradient_bc<<<ceil_div(n, 256), 256, 0, stream>>>(p));
}

void DeviceMG::residual(const Grid &p, const Grid &f, const Grid &flag, const Grid &r, float hh,
                        bool want_norm) {
  dim3 g = grd2d(p.w, p.h);
  UBGL_REQUIRE(!want_norm || (int)(g.x * g.y) <= n_partials, "residual: grid larger than MG");
  float ihsq = 1.0f / hh / hh;
  UBGL_LAUNCH(lc, K_RESIDUAL, LVL, stream, k_residual<<<g, blk2d(), 0, stream>>>(p, f, flag, r, ihsq, want_norm ? d_partials : nullptr));
  if (want_norm) {
    UBGL_LAUNCH(lc, K_NORM, LVL, stream, k_finish_norm<<<1, 1024, 0, stream>>>(d_partials, g.x * g.y, d_norm));
  }
}

float DeviceMG::residual_norm_result() {
  double v = 0.0;
  UBGL_CUDA(cudaMemcpyAsync(&v, d_norm, sizeof(double), cudaMemcpyDeviceToHost, stream));
  UBGL_CUDA(cudaStreamSynchronize(stream));
  return (float)v;
}

void DeviceMG::restrict_fw(const Grid &r, const Grid &rc) {
  UBGL_LAUNCH(lc, K_RESTRICT, LVL, stream, k_restrict<<<grd2d(rc.w, rc.h), blk2d(), 0, stream>>>(r, rc));
}

void DeviceMG::prolongate(const Grid &e, const Grid &ec, const Grid &flagc, const Grid &flag) {
  UBGL_LAUNCH(lc, K_PROLONG, LVL, stream, k_prolongate<<<grd2d(e.w, e.h), blk2d(), 0, stream>>>(e, ec, flagc, flag));
}

void DeviceMG::correct(const Grid &p, const Grid &e) {
  UBGL_LAUNCH(lc, K_PROLONG, LVL, stream, k_correct<<<grd2d(p.w - 2, p.h - 2), blk2d(), 0, stream>>>(p, e));
}

void DeviceMG::prolongate_correct(const Grid &p, const Grid &ec, const Grid &flagc,
                                  const Grid &flag) {
  UBGL_LAUNCH(lc, K_PROLONG, LVL, stream, k_prolongate_correct<<<grd2d(p.w - 2, p.h - 2), blk2d(), 0, stream>>>(p, ec, flagc, flag));
}

void DeviceMG::solve(const Grid &p, const Grid &f, const Grid &flag, float hh, bool zgbc) {
  UBGL_REQUIRE(p.w == lv[0].w && p.h == lv[0].h, "solve: grid size mismatch");
  if (fused && levels() >= 3 && p.pitch == lv[0].pitch && f.pitch == lv[0].pitch) {
    prepare_mask0(flag);
    if (mask0_binary) {
      solve_fused(p, f, flag, hh, zgbc);
      return;
    }
  }
  solve_level(p, f, flag, hh, 0, zgbc);
  cur_level = 0;
}

// MG::solveLevel (pressure_solver.cpp:201-248) with the temporally blocked tile
// kernels of mg_fused.cu: per level one PRE pass (3 sweeps + residual +
// restriction) on the way down, one 5-sweep pass on the coarsest used level
// (levels-2), one POST pass (prolongation + correction + 3 sweeps) on the way up.
void DeviceMG::solve_fused(const Grid &p, const Grid &f, const Grid &flag, float hh0, bool zgbc) {
  (void)flag;
  const int L = levels() - 2;
  if (!scratch0.d) scratch0 = alloc_grid(lv[0].w, lv[0].h, lv[0].pitch);
  std::vector<float> hh(L + 1);
  hh[0] = hh0;
  for (int l = 0; l < L; l++) // :229
    hh[l + 1] = hh[l] * ((float)lv[l].w - 1.0f) / ((float)lv[l + 1].w - 1.0f);
  // levels t..L run in one launch (k_mg_tail) when they fit a CTA's shared memory
  std::vector<TailLevel> tv;
  for (auto &V : lv) tv.push_back(TailLevel{V.w, V.h, V.pitch, V.mask});
  const int t = mg_tail_first_level(tv, 1);
  const int Ld = t ? t : L; // levels [0, Ld) use the PRE / POST passes
  for (int l = 0; l < Ld; l++) {
    const Grid &fl = (l == 0) ? f : lv[l].rc;
    launch_mg_pre(l == 0 ? p.d : nullptr, l == 0 ? scratch0.d : lv[l].eb.d, fl,
                  l == 0 ? mask0 : lv[l].mask, lv[l + 1].rc, hh[l], zgbc && l == 0, stream, lc, l);
  }
  if (t)
    launch_mg_tail(tv, t, hh.data(), lv[t].rc.d, lv[t].ec.d, stream, lc);
  else
    launch_mg_smooth5(lv[L].ec.d, lv[L].rc, lv[L].mask, hh[L], stream, lc, L);
  for (int l = Ld - 1; l >= 0; l--) {
    const Grid &fl = (l == 0) ? f : lv[l].rc;
    launch_mg_post(l == 0 ? scratch0.d : lv[l].eb.d, l == 0 ? p.d : lv[l].ec.d, fl,
                   l == 0 ? mask0 : lv[l].mask, lv[l + 1].ec, lv[l + 1].mask, hh[l],
                   zgbc && l == 0, stream, lc, l);
  }
}

// MG::solveLevel (pressure_solver.cpp:201-248), plain path: one kernel per
// reference operator, same order.
void DeviceMG::solve_level(const Grid &p, const Grid &f, const Grid &flag, float hh, int level,
                           bool zgbc) {
  const int nl = levels();
  cur_level = level;
  if (level == nl - 2) {
    for (int i = 0; i < 5; i++) rbgs(p, f, flag, hh, 1.0f);
    return;
  }
  const bool bc = (level == 0 && zgbc);
  for (int i = 0; i < 3; i++) {
    rbgs(p, f, flag, hh, 1.0f);
    if (bc) zero_gradient_bc(p);
  }
  MGLevel &C = lv[level + 1];
  ensure_r(level);
  residual(p, f, flag, lv[level].r, hh, false);
  restrict_fw(lv[level].r, C.rc);
  fill_grid(C.ec, 0.0f, stream, lc);
  float hc = hh * ((float)p.w - 1.0f) / ((float)C.w - 1.0f);
  solve_level(C.ec, C.rc, C.flagc, hc, level + 1, false);
  cur_level = level;
  prolongate_correct(p, C.ec, C.flagc, flag);
  if (bc) zero_gradient_bc(p);
  for (int i = 0; i < 3; i++) {
    rbgs(p, f, flag, hh, 1.0f);
    if (bc) zero_gradient_bc(p);
  }
}

} // namespace ubgl
