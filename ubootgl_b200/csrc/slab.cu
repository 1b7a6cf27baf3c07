// slab.cu -- row-slab decomposed Simulation::step (see slab.cuh).
#include "slab.cuh"
#include "hostwork.cuh"
#include "stencils.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace ubgl {

// ---------------------------------------------------------------------------
// plan (pure host)
// ---------------------------------------------------------------------------
// own rows per rank on the coarsest distributed level; coarser levels are replicated.
// A distributed coarse level costs a ~14 us kernel + a ~20 us exchange per leg whatever its
// size, a replicated one only the kernel (plus a larger all-gather, once): measured at 8 GPUs on
// 32768^2 (tools/gpu_t.sh) 9.35 / 9.31 / 9.27 / 9.36 / 9.88 ms per step for 64 / 128 / 256 / 512 /
// 1024 rows.  Tall slabs therefore stop distributing at 256 rows (which also aligns the cuts to 32
// rows instead of 128: finer load balancing); short ones keep 64 so that small grids still split.
static int slab_min_rows(int H, int nranks) {
  const char *e = getenv("UBGL_SLAB_MIN_ROWS");
  const int v = e ? atoi(e) : (H / nranks >= 2048 ? 256 : 64);
  return v >= 32 ? v : 32;
}

// Row weights of the next plans (ubgl_slab_set_row_weights): relative cost of a level-0 row.
// The advect pass skips octets without fluid (simulation.cpp:254-256), so rows through obstacles
// are cheaper than open-channel rows; with equal-height slabs the obstacle-free ranks are the
// slowest and every halo exchange waits for them.  Process-wide, like the option env vars.
static std::vector<double> g_row_weight_prefix; // prefix sums, H + 1 entries; empty: equal rows
void set_slab_row_weights(const float *w, int H) {
  g_row_weight_prefix.clear();
  if (!w || H <= 0) return;
  g_row_weight_prefix.resize((size_t)H + 1, 0.0);
  for (int y = 0; y < H; y++) {
    UBGL_REQUIRE(w[y] > 0.0f, "slab: row weights must be positive");
    g_row_weight_prefix[y + 1] = g_row_weight_prefix[y] + (double)w[y];
  }
}

SlabPlan make_slab_plan(int W, int H, int nranks, int rank) {
  UBGL_REQUIRE(nranks >= 1 && nranks <= SLAB_MAXRANKS, "slab: 1..8 ranks");
  UBGL_REQUIRE(rank >= 0 && rank < nranks, "slab: bad rank");
  SlabPlan P;
  P.W = W; P.H = H; P.nranks = nranks; P.rank = rank;
  int cw = W, ch = H;
  while (cw > 3 && ch > 3) { // pressure_solver.hpp:20-29
    P.lw.push_back(cw);
    P.lh.push_back(ch);
    cw /= 2;
    ch /= 2;
  }
  P.levels = (int)P.lw.size();
  UBGL_REQUIRE(P.levels >= 3, "slab: grid too small for a multigrid pyramid");
  const int L = P.levels - 2; // coarsest used level (pressure_solver.cpp:203)
  int n = 0;
  const int min_rows = slab_min_rows(H, nranks);
  while (n < L && ((H >> n) / nranks) >= min_rows) n++;
  UBGL_REQUIRE(n >= 1, "slab: fewer than 64 rows per GPU at level 0 -- use fewer GPUs");
  P.ndist = n;
  const int align = 1 << n;
  P.cuts.resize(nranks + 1);
  for (int r = 0; r < nranks; r++) P.cuts[r] = (int)((long long)H * r / nranks) / align * align;
  P.cuts[nranks] = H;
  if ((int)g_row_weight_prefix.size() == H + 1 && nranks > 1) {
    // equal WEIGHT per rank: cut r at the aligned row whose prefix weight is nearest r/nranks of the total
    const std::vector<double> &pre = g_row_weight_prefix;
    for (int r = 1; r < nranks; r++) {
      const double target = pre[H] * r / nranks;
      int y = (int)(std::lower_bound(pre.begin(), pre.end(), target) - pre.begin());
      int c = (y + align / 2) / align * align;
      c = std::max(c, P.cuts[r - 1] + align);
      c = std::min(c, H - (nranks - r) * align);
      P.cuts[r] = c;
    }
  }
  for (int r = 0; r < nranks; r++)
    UBGL_REQUIRE(((P.cuts[r + 1] - P.cuts[r]) >> (n - 1)) >= 2 * P.ghost,
                 "slab: slab thinner than two halos on the coarsest distributed level");
  return P;
}

Rows SlabPlan::rows(int l, int r) const {
  Rows R;
  const int hl = lh[l];
  if (l >= ndist) {
    R.st_lo = R.own_lo = 0;
    R.st_hi = R.own_hi = hl;
    return R;
  }
  R.own_lo = cuts[r] >> l;
  R.own_hi = (r == nranks - 1) ? hl : (cuts[r + 1] >> l);
  R.st_lo = std::max(0, R.own_lo - ghost);
  R.st_hi = std::min(hl, R.own_hi + ghost);
  return R;
}

int SlabPlan::max_stored_rows(int l) const {
  int m = 0;
  for (int r = 0; r < nranks; r++) {
    Rows R = rows(l, r);
    m = std::max(m, R.st_hi - R.st_lo);
  }
  return m;
}

// ---------------------------------------------------------------------------
// halo kernels
// ---------------------------------------------------------------------------
// Copies every segment with 128-bit loads/stores (dst is peer memory mapped over
// NVLink), then publishes: every thread fences its stores system-wide, the last
// block to finish releases `seq` into the peers' signal slots, one slot per thread.
// The critical path of an exchange is launch -> stores -> fence -> (block count) -> release
// -> the neighbour's spin; a step makes ~34 of them, so nothing else sits on it: a launch of a
// single block (the small exchanges of the coarse levels) skips the block counter altogether, the counter is
// re-armed AFTER the release, and no fence follows the release (nothing later in this kernel
// depends on it; the peer polls its own memory).
__global__ void __launch_bounds__(256) k_halo_push(HaloPush a) {
  ubgl_pdl_prologue();
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nthreads = (size_t)gridDim.x * blockDim.x;
  for (int s = 0; s < a.nseg; s++) {
    const uint4 *__restrict__ src = a.seg[s].src;
    uint4 *__restrict__ dst = a.seg[s].dst;
    const size_t n = a.seg[s].n16;
    for (size_t i = tid; i < n; i += nthreads) dst[i] = src[i];
  }
  __threadfence_system(); // this thread's peer stores are performed before anything below
  __syncthreads();
  __shared__ int is_last;
  if (gridDim.x == 1) {
    if (threadIdx.x == 0) is_last = 1;
  } else if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(a.counter, 1u);
    is_last = done == gridDim.x - 1;
    if (is_last) __threadfence_system(); // the other blocks' stores (fenced before their atomicAdd) before the release
  }
  __syncthreads();
  if (!is_last) return;
  if ((int)threadIdx.x < a.nsig && a.sig[threadIdx.x])
    *reinterpret_cast<volatile unsigned *>(a.sig[threadIdx.x]) = a.seq;
  if (threadIdx.x == 0 && gridDim.x > 1) *a.counter = 0; // ready for the next launch on this stream
  // the last block stays until the neighbours have released the same sequence number into
  // this rank's slots (bounded spin, see k_halo_wait); every rank releases before it waits
  if ((int)threadIdx.x < a.nwait) {
    volatile unsigned *s = a.wait[threadIdx.x];
    const long long t0 = clock64();
    while ((int)(*s - a.seq) < 0) {
      if (clock64() - t0 > 40000000000LL) {
        *a.err = 2;
        break;
      }
      __nanosleep(40);
    }
    __threadfence_system(); // acquire: the ghost rows the neighbour stored before its release
  }
}

// Waits until every slot has reached `seq` (slots only ever increase).  Bounded:
// after ~20 s of spinning the error word is raised instead of hanging the GPU.
__global__ void k_halo_wait(unsigned *s0, unsigned *s1, unsigned *s2, unsigned *s3, unsigned *s4,
                            unsigned *s5, unsigned *s6, unsigned *s7, int n, unsigned seq, int *err) {
  unsigned *slots[8] = {s0, s1, s2, s3, s4, s5, s6, s7};
  const int i = threadIdx.x;
  if (i < n) {
    volatile unsigned *s = slots[i];
    const long long t0 = clock64();
    while ((int)(*s - seq) < 0) {
      if (clock64() - t0 > 40000000000LL) {
        *err = 2;
        break;
      }
      __nanosleep(200);
    }
  }
  __threadfence_system();
}

void launch_halo_push(const HaloPush &a, size_t total16, cudaStream_t stream, LaunchCounter *lc) {
  // up to 4 K 16-byte words (64 KB, the coarse distributed levels): one block, 16 independent
  // load/store pairs per thread, no block counter; larger: ~4 per thread over up to 2 blocks per SM
  int blocks = total16 <= (size_t)(4 << 10) ? 1 : (int)std::min<size_t>(148 * 2, (total16 + 1023) / 1024);
  if (blocks < 1) blocks = 1;
  UBGL_LAUNCH(lc, K_HALO_PUSH, 0, stream, launch_k(k_halo_push, blocks, 256, 0, stream, a));
}

void launch_halo_wait(unsigned *const *slots, int n, unsigned seq, int *err, cudaStream_t stream,
                      LaunchCounter *lc) {
  unsigned *s[8] = {};
  for (int i = 0; i < n; i++) s[i] = slots[i];
  UBGL_LAUNCH(lc, K_HALO_WAIT, 0, stream, k_halo_wait<<<1, 32, 0, stream>>>(s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], n, seq, err));
}

// residual sum of squares over rows [y_lo, y_hi) (calculateResidualField,
// pressure_solver.cpp:91-116, fp32 flags)
__global__ void k_slab_residual(Grid p, Grid f, Grid flag, float ihsq, int y_lo, int y_hi, double *sum) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = y_lo + blockIdx.y * blockDim.y + threadIdx.y;
  float rv = 0.0f;
  if (x >= 1 && x < p.w - 1 && y >= 1 && y < p.h - 1 && y < y_hi) {
    const size_t c = (size_t)y * p.pitch + x;
    const float *fl = flag.d + c;
    rv = residual_cell(p.d[c], p.d[c - 1], p.d[c + 1], p.d[c - p.pitch], p.d[c + p.pitch], fl[0],
                       fl[-1], fl[1], fl[-flag.pitch], fl[flag.pitch], f.d[c], ihsq);
  }
  double s = (double)rv * (double)rv;
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  __shared__ double ws[8];
  const int t = threadIdx.y * blockDim.x + threadIdx.x;
  if ((t & 31) == 0) ws[t >> 5] = s;
  __syncthreads();
  if (t == 0) {
    double tot = 0.0;
    for (int i = 0; i < 8; i++) tot += ws[i];
    atomicAdd(sum, tot);
  }
}

// ---------------------------------------------------------------------------
// SlabSim: allocation
// ---------------------------------------------------------------------------
char *SlabSim::take(size_t bytes) {
  bytes = (bytes + 255) / 256 * 256;
  UBGL_REQUIRE(arena_off + bytes <= arena_bytes, "slab arena exhausted (internal sizing error)");
  char *q = arena + arena_off;
  arena_off += bytes;
  return q;
}

Grid SlabSim::arena_grid(int w, int hh, int level_pitch, int level, bool dist) {
  Grid g;
  g.w = w;
  g.h = hh;
  g.pitch = level_pitch;
  const int nst = dist ? plan.max_stored_rows(level) : plan.lh[level];
  char *real = take(sizeof(float) * (size_t)level_pitch * nst);
  const int st_lo = dist ? plan.rows(level).st_lo : 0;
  g.d = reinterpret_cast<float *>(real) - (ptrdiff_t)st_lo * level_pitch;
  return g;
}

uint8_t *SlabSim::arena_mask(int hh, int level_pitch, int level, bool dist) {
  (void)hh;
  const int nst = dist ? plan.max_stored_rows(level) : plan.lh[level];
  char *real = take((size_t)level_pitch * nst);
  const int st_lo = dist ? plan.rows(level).st_lo : 0;
  return reinterpret_cast<uint8_t *>(real) - (ptrdiff_t)st_lo * level_pitch;
}

SlabSim::SlabSim(const float *flag_slab, int W_, int H_, float pwidth_, float mu_, int device_,
                 int rank, int nranks)
    : W(W_), H(H_), pwidth(pwidth_), mu(mu_), device(device_) {
  UBGL_REQUIRE(flag_slab != nullptr, "flag must not be null");
  plan = make_slab_plan(W, H, nranks, rank);
  UBGL_CUDA(cudaSetDevice(device));
  UBGL_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  pitch = round_up(W, 32);
  h = pwidth / ((float)W - 1.0f);
  if (const char *e = getenv("UBGL_LAZY_CURRENT")) lazy_current = e[0] != '0';

  // arena size: control block + 14 level-0 fp32 slabs + mask, the distributed
  // pyramid (4 fp32 + mask per level) and the replicated levels (same arrays, whole)
  size_t bytes = 4096;
  auto lvl_bytes = [&](int l, bool dist, int nf32) {
    const size_t rows = dist ? plan.max_stored_rows(l) : plan.lh[l];
    const size_t lp = round_up(plan.lw[l], 32);
    return (size_t)nf32 * (rows * lp * 4 + 256) + rows * lp + 256;
  };
  bytes += lvl_bytes(0, true, 14);
  for (int l = 1; l < plan.levels; l++) bytes += lvl_bytes(l, l < plan.ndist, 4);
  arena_bytes = bytes;
  UBGL_CUDA(cudaMalloc(&arena, arena_bytes));
  // on OUR stream: a legacy-stream cudaMemset is asynchronous and does not order with a
  // non-blocking stream -- it would race with the uploads below and zero fresh data
  UBGL_CUDA(cudaMemsetAsync(arena, 0, arena_bytes, stream));
  UBGL_CUDA(cudaStreamSynchronize(stream));
  char *ctl = take(4096);
  sig = reinterpret_cast<unsigned *>(ctl);
  counter = reinterpret_cast<unsigned *>(ctl + 256);
  err = reinterpret_cast<int *>(ctl + 512);
  d_nonbinary = reinterpret_cast<int *>(ctl + 768);
  d_sum = reinterpret_cast<double *>(ctl + 1024);

  for (int b = 0; b < 3; b++) {
    vxb[b] = arena_grid(W - 1, H, pitch, 0, true);
    vyb[b] = arena_grid(W, H - 1, pitch, 0, true);
  }
  vx_accum = arena_grid(W - 1, H, pitch, 0, true);
  vy_accum = arena_grid(W, H - 1, pitch, 0, true);
  p = arena_grid(W, H, pitch, 0, true);
  scratch0 = arena_grid(W, H, pitch, 0, true);
  f = arena_grid(W, H, pitch, 0, true);
  flag = arena_grid(W, H, pitch, 0, true);
  mask0 = arena_mask(H, pitch, 0, true);
  lv.resize(plan.levels);
  for (int l = 0; l < plan.levels; l++) {
    Level &L = lv[l];
    L.w = plan.lw[l];
    L.h = plan.lh[l];
    L.pitch = round_up(L.w, 32);
    L.dist = l < plan.ndist;
    L.rows = plan.rows(l);
    if (l == 0) {
      L.flagc = flag; // the caller's flag IS flagcs[0] (pressure_solver.hpp:35)
      L.mask = mask0;
      continue;
    }
    L.flagc = arena_grid(L.w, L.h, L.pitch, l, L.dist);
    if (l + 1 < plan.levels) {
      L.rc = arena_grid(L.w, L.h, L.pitch, l, L.dist);
      L.ec = arena_grid(L.w, L.h, L.pitch, l, L.dist);
      L.eb = arena_grid(L.w, L.h, L.pitch, l, L.dist);
      L.mask = arena_mask(L.h, L.pitch, l, L.dist);
    }
  }
  // Simulation(flag, pwidth, mu), simulation.hpp:32-67
  upload(F_FLAG, flag_slab);
  {
    const Rows R = plan.rows(0);
    std::vector<float> col(R.st_hi - R.st_lo, 1.0f); // vx.f(0,y) = vx.b(0,y) = 1 (:58-60)
    for (int b = 0; b < 2; b++)
      UBGL_CUDA(cudaMemcpy2DAsync(&vxb[b].at(0, R.st_lo), sizeof(float) * pitch, col.data(),
                                  sizeof(float), sizeof(float), col.size(), cudaMemcpyHostToDevice,
                                  stream));
    UBGL_CUDA(cudaStreamSynchronize(stream));
  }
}

SlabSim::~SlabSim() {
  cudaSetDevice(device);
  if (stream) cudaStreamSynchronize(stream);
  for (int r = 0; r < SLAB_MAXRANKS; r++)
    if (peer_arena[r] && r != plan.rank) cudaIpcCloseMemHandle(peer_arena[r]);
  if (d_sinks) cudaFree(d_sinks);
  if (d_pack) cudaFree(d_pack);
  if (arena) cudaFree(arena);
  if (stream) cudaStreamDestroy(stream);
}

void SlabSim::ipc_export(void *blob64) const {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
  cudaIpcMemHandle_t hd;
  UBGL_CUDA(cudaIpcGetMemHandle(&hd, arena));
  std::memcpy(blob64, &hd, 64);
}

void SlabSim::connect(const void *blobs64) {
  UBGL_CUDA(cudaSetDevice(device));
  for (int r = 0; r < plan.nranks; r++) {
    if (r == plan.rank) {
      peer_arena[r] = arena;
      continue;
    }
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, (const char *)blobs64 + 64 * r, 64);
    void *q = nullptr;
    UBGL_CUDA(cudaIpcOpenMemHandle(&q, hd, cudaIpcMemLazyEnablePeerAccess));
    peer_arena[r] = (char *)q;
  }
  connected = true;
}

// the address in rank r's arena of the array whose LOCAL virtual base is given
template <typename T>
T *SlabSim::peer_ptr(int r, const T *local_virtual, int level, size_t row_bytes) const {
  const bool dist = level < plan.ndist;
  const ptrdiff_t my_lo = dist ? plan.rows(level).st_lo : 0;
  const ptrdiff_t his_lo = dist ? plan.rows(level, r).st_lo : 0;
  const char *my_real = (const char *)local_virtual + my_lo * (ptrdiff_t)row_bytes;
  const ptrdiff_t off = my_real - arena; // identical layout on every rank
  char *his_real = peer_arena[r] + off;
  return (T *)(his_real - his_lo * (ptrdiff_t)row_bytes);
}

void SlabSim::push_and_wait(HaloPush &a, const std::vector<int> &peers) {
  UBGL_REQUIRE(connected, "slab: connect() the peers first");
  seq++;
  a.seq = seq;
  a.counter = counter;
  a.nsig = 0;
  unsigned *slots[SLAB_MAXRANKS];
  int nslots = 0;
  for (int r : peers) {
    // my slot on rank r is sig[plan.rank] in ITS control block (offset 0 of the arena)
    a.sig[a.nsig++] = reinterpret_cast<unsigned *>(peer_arena[r]) + plan.rank;
    slots[nslots++] = sig + r;
  }
  size_t total = 0;
  for (int s = 0; s < a.nseg; s++) total += a.seg[s].n16;
  halo_bytes += total * 16;
  exchanges++;
  a.nwait = nslots;
  for (int i = 0; i < nslots; i++) a.wait[i] = slots[i];
  a.err = err;
  launch_halo_push(a, total, stream, &lc);
}

// Neighbour exchange: my first `depth` own rows go into the lower neighbour's
// upper ghost rows, my last `depth` own rows into the upper neighbour's lower
// ghost rows (same global row indices on both sides).
void SlabSim::exchange(const std::vector<XField> &fields, int depth) {
  if (plan.nranks == 1) return;
  HaloPush a{};
  std::vector<int> peers;
  const int lo = plan.rank - 1, hi = plan.rank + 1;
  if (lo >= 0) peers.push_back(lo);
  if (hi < plan.nranks) peers.push_back(hi);
  for (const XField &x : fields) {
    const Rows R = plan.rows(x.level);
    UBGL_REQUIRE(x.level < plan.ndist && depth <= plan.ghost && depth <= R.own_hi - R.own_lo,
                 "slab: bad exchange");
    const int top = std::min(R.own_hi, x.rows_hi_clip); // staggered vy has one row less
    auto add = [&](int peer, int row0, int nrows) {
      if (nrows <= 0) return;
      UBGL_REQUIRE(a.nseg < SLAB_MAXSEG, "slab: too many halo segments");
      const char *src = (const char *)x.base + (size_t)row0 * x.row_bytes;
      char *dst = (char *)peer_ptr(peer, (const char *)x.base, x.level, x.row_bytes) +
                  (size_t)row0 * x.row_bytes;
      a.seg[a.nseg++] = HaloSeg{(const uint4 *)src, (uint4 *)dst, (size_t)nrows * x.row_bytes / 16};
    };
    if (lo >= 0) add(lo, R.own_lo, depth);
    if (hi < plan.nranks) add(hi, top - depth, depth);
  }
  push_and_wait(a, peers);
}

// Own rows of replicated arrays (levels >= ndist) to every peer.
void SlabSim::allgather(const std::vector<XField> &fields) {
  if (plan.nranks == 1) return;
  std::vector<int> peers;
  for (int r = 0; r < plan.nranks; r++)
    if (r != plan.rank) peers.push_back(r);
  HaloPush a{};
  for (const XField &x : fields) {
    UBGL_REQUIRE(x.level >= plan.ndist, "slab: allgather is for replicated levels");
    const int hl = plan.lh[x.level];
    const int own_lo = plan.cuts[plan.rank] >> x.level;
    const int own_hi = plan.rank == plan.nranks - 1 ? hl : (plan.cuts[plan.rank + 1] >> x.level);
    for (int r : peers) {
      UBGL_REQUIRE(a.nseg < SLAB_MAXSEG, "slab: too many halo segments");
      const char *src = (const char *)x.base + (size_t)own_lo * x.row_bytes;
      char *dst = (char *)peer_ptr(r, (const char *)x.base, x.level, x.row_bytes) +
                  (size_t)own_lo * x.row_bytes;
      a.seg[a.nseg++] =
          HaloSeg{(const uint4 *)src, (uint4 *)dst, (size_t)(own_hi - own_lo) * x.row_bytes / 16};
    }
  }
  push_and_wait(a, peers);
}

void SlabSim::check_err() {
  int e = 0;
  UBGL_CUDA(cudaMemcpyAsync(&e, err, sizeof(int), cudaMemcpyDeviceToHost, stream));
  UBGL_CUDA(cudaStreamSynchronize(stream));
  if (e) {
    UBGL_CUDA(cudaMemsetAsync(err, 0, sizeof(int), stream));
    throw ArgError{e == 2 ? "slab: halo wait timed out (a peer rank stopped?)"
                          : "slab: advect back-trace is longer than the neighbouring slab (CFL > rows per GPU)"};
  }
}

// ---------------------------------------------------------------------------
// fields
// ---------------------------------------------------------------------------
Grid SlabSim::field(int id) {
  switch (id) {
  case F_FLAG: return flag;
  case F_VX: return vxb[ixf];
  case F_VY: return vyb[iyf];
  case F_VXB: return vxb[ixb];
  case F_VYB: return vyb[iyb];
  case F_P: return p;
  case F_F: return f;
  case F_VX_ACCUM: return vx_accum;
  case F_VY_ACCUM: return vy_accum;
  case F_VX_CURRENT: return vxb[cur_alias ? ixf : ixc];
  case F_VY_CURRENT: return vyb[cur_alias ? iyf : iyc];
  }
  throw ArgError{"unknown / unsupported field id in slab mode"};
}

void SlabSim::will_write(int id) {
  if (!cur_alias) return;
  if (id != F_VX && id != F_VY && id != F_VX_CURRENT && id != F_VY_CURRENT) return;
  cur_alias = false;
  const Rows R = plan.rows(0);
  auto copy = [&](const Grid &src, const Grid &dst) { // stored rows, whole pitch
    const int n = std::min(R.st_hi, src.h) - R.st_lo;
    if (n > 0)
      UBGL_CUDA(cudaMemcpyAsync(&dst.at(0, R.st_lo), &src.at(0, R.st_lo), sizeof(float) * (size_t)src.pitch * n,
                                cudaMemcpyDeviceToDevice, stream));
  };
  copy(vxb[ixf], vxb[ixc]);
  copy(vyb[iyf], vyb[iyc]);
}

void SlabSim::field_rows(int id, int *row_lo, int *nrows, int *w) const {
  const Rows R = plan.rows(0);
  int gh = H, gw = W;
  switch (id) {
  case F_VX: case F_VXB: case F_VX_ACCUM: case F_VX_CURRENT: gw = W - 1; break;
  case F_VY: case F_VYB: case F_VY_ACCUM: case F_VY_CURRENT: gh = H - 1; break;
  default: break;
  }
  *row_lo = R.st_lo;
  *nrows = std::min(R.st_hi, gh) - R.st_lo;
  *w = gw;
}

void SlabSim::upload(int id, const float *host) {
  UBGL_REQUIRE(host != nullptr, "upload: null host pointer");
  will_write(id);
  Grid g = field(id);
  int r0, n, w;
  field_rows(id, &r0, &n, &w);
  UBGL_CUDA(cudaMemcpy2DAsync(&g.at(0, r0), sizeof(float) * g.pitch, host, sizeof(float) * w,
                              sizeof(float) * w, n, cudaMemcpyHostToDevice, stream));
  UBGL_CUDA(cudaStreamSynchronize(stream));
}

void SlabSim::download(int id, float *host) {
  UBGL_REQUIRE(host != nullptr, "download: null host pointer");
  Grid g = field(id);
  int r0, n, w;
  field_rows(id, &r0, &n, &w);
  UBGL_CUDA(cudaMemcpy2DAsync(host, sizeof(float) * w, &g.at(0, r0), sizeof(float) * g.pitch,
                              sizeof(float) * w, n, cudaMemcpyDeviceToHost, stream));
  UBGL_CUDA(cudaStreamSynchronize(stream));
  check_err();
}

void SlabSim::sync() {
  UBGL_CUDA(cudaStreamSynchronize(stream));
  check_err();
}

// ---------------------------------------------------------------------------
// MG::updateFields (pressure_solver.hpp:34-57) + stencil masks, slab-wise
// ---------------------------------------------------------------------------
void SlabSim::finish_setup() { update_fields(); }

void SlabSim::update_fields() {
  UBGL_CUDA(cudaMemsetAsync(d_nonbinary, 0, sizeof(int), stream));
  const int G = plan.ghost;
  // level 0: flag ghost rows came with the upload; mask own rows, then ghost from the owners
  launch_make_mask(flag, mask0, d_nonbinary, stream, &lc, 0, &lv[0].rows);
  exchange({xm(mask0, 0)}, G);
  for (int l = 1; l < plan.levels; l++) {
    Level &L = lv[l];
    if (L.dist) {
      launch_coarsen_flag(lv[l - 1].flagc, L.flagc, L.rows.own_lo, L.rows.own_hi, stream, &lc, l);
      exchange({xf(L.flagc, l)}, G);
    } else if (l == plan.ndist) {
      // first replicated level: own coarse rows from the distributed level above, then all-gather
      const int own_lo = plan.cuts[plan.rank] >> l;
      const int own_hi = plan.rank == plan.nranks - 1 ? L.h : (plan.cuts[plan.rank + 1] >> l);
      launch_coarsen_flag(lv[l - 1].flagc, L.flagc, own_lo, own_hi, stream, &lc, l);
      allgather({xf(L.flagc, l)});
    } else {
      launch_coarsen_flag(lv[l - 1].flagc, L.flagc, 0, L.h, stream, &lc, l);
    }
    if (L.mask) {
      launch_make_mask(L.flagc, L.mask, d_nonbinary, stream, &lc, l, &L.rows);
      if (L.dist) exchange({xm(L.mask, l)}, G);
    }
  }
  int nb = 0;
  UBGL_CUDA(cudaMemcpyAsync(&nb, d_nonbinary, sizeof(int), cudaMemcpyDeviceToHost, stream));
  UBGL_CUDA(cudaStreamSynchronize(stream));
  check_err();
  UBGL_REQUIRE(nb == 0, "slab mode needs binary flags (0.0 / 1.0)");
}

// ---------------------------------------------------------------------------
// MG::solveLevel (pressure_solver.cpp:201-248) with the fused tile kernels
// ---------------------------------------------------------------------------
void SlabSim::mg_solve() {
  const int L = plan.levels - 2;
  const int nd = plan.ndist;
  std::vector<float> hh(L + 1);
  hh[0] = h;
  for (int l = 0; l < L; l++) // :229
    hh[l + 1] = hh[l] * ((float)lv[l].w - 1.0f) / ((float)lv[l + 1].w - 1.0f);
  // replicated levels t..L that fit one CTA's shared memory run in one launch (k_mg_tail),
  // identically on every rank
  std::vector<TailLevel> tv;
  for (auto &V : lv) tv.push_back(TailLevel{V.w, V.h, V.pitch, V.mask});
  const int t = mg_tail_first_level(tv, nd > 1 ? nd : 1);
  const int Ld = t ? t : L;
  for (int l = 0; l < Ld; l++) {
    const Grid &fl = (l == 0) ? f : lv[l].rc;
    const Grid &pout = (l == 0) ? scratch0 : lv[l].eb;
    const Rows *rows = lv[l].dist ? &lv[l].rows : nullptr;
    launch_mg_pre(l == 0 ? p.d : nullptr, pout.d, fl, lv[l].mask, lv[l + 1].rc, hh[l], l == 0, stream,
                  &lc, l, rows);
    if (lv[l].dist) {
      if (l + 1 < nd) {
        exchange({xf(lv[l + 1].rc, l + 1), xf(pout, l)}, 8);
      } else {
        exchange({xf(pout, l)}, 8);
        allgather({xf(lv[l + 1].rc, l + 1)});
      }
    }
  }
  if (t)
    launch_mg_tail(tv, t, hh.data(), lv[t].rc.d, lv[t].ec.d, stream, &lc);
  else
    launch_mg_smooth5(lv[L].ec.d, lv[L].rc, lv[L].mask, hh[L], stream, &lc, L);
  for (int l = Ld - 1; l >= 0; l--) {
    const Grid &fl = (l == 0) ? f : lv[l].rc;
    const Grid &pin = (l == 0) ? scratch0 : lv[l].eb;
    const Grid &pout = (l == 0) ? p : lv[l].ec;
    const Rows *rows = lv[l].dist ? &lv[l].rows : nullptr;
    const Rows *crows = lv[l + 1].dist ? &lv[l + 1].rows : nullptr;
    launch_mg_post(pin.d, pout.d, fl, lv[l].mask, lv[l + 1].ec, lv[l + 1].mask, hh[l], l == 0, stream,
                   &lc, l, rows, crows);
    if (lv[l].dist && l > 0) exchange({xf(pout, l)}, 8);
    // level 0: the caller exchanges p (after setPBC on the last cycle)
  }
}

void SlabSim::project_sinks() {
  // simulation.cpp:173-187, identical on every rank; each stamps its own rows
  std::vector<float> stamps;
  for (auto &s : sinks) {
    float gx = s.x / h + 0.5f, gy = s.y / h + 0.5f;
    if (gx <= 3 || gx > (float)(W - 3) || gy <= 3 || gy > (float)(H - 3)) continue;
    stamps.push_back((float)(int)gx);
    stamps.push_back((float)(int)gy);
    stamps.push_back(s.z);
    s.z = (float)((double)s.z * std::pow(0.000001, (double)(dt * 50)));
  }
  size_t n = 0;
  for (size_t k = 0; k < sinks.size(); k++)
    if (!(sinks[k].z < 0.05f)) sinks[n++] = sinks[k];
  sinks.resize(n);
  if (stamps.empty()) return;
  const int ns = (int)stamps.size() / 3;
  if (ns > cap_sinks) {
    if (d_sinks) UBGL_CUDA(cudaFree(d_sinks));
    cap_sinks = ns * 2;
    UBGL_CUDA(cudaMalloc(&d_sinks, sizeof(float) * 3 * cap_sinks));
  }
  UBGL_CUDA(cudaMemcpyAsync(d_sinks, stamps.data(), sizeof(float) * stamps.size(),
                            cudaMemcpyHostToDevice, stream));
  UBGL_CUDA(cudaStreamSynchronize(stream));
  const Rows R = plan.rows(0);
  launch_stamp_sinks(f, d_sinks, ns, R.own_lo, R.own_hi, stream, &lc);
}

// ---------------------------------------------------------------------------
// Simulation::step (simulation.cpp:356-374), slab-wise.  Entry invariant: the
// front velocity buffers and p are valid on own rows + ghost rows.
// ---------------------------------------------------------------------------
void SlabSim::step(float dt_) {
  dt = dt_;
  cur_alias = false; // the third buffers are scratch during the step
  const Rows R = plan.rows(0);
  const bool first = plan.rank == 0, last = plan.rank == plan.nranks - 1;
  const float ih = 1.0f / h;

  auto borders = [&](bool with_p, bool with_current) {
    BorderArgs g{};
    g.xf = vxb[ixf]; g.xb = vxb[ixb]; g.yf = vyb[iyf]; g.yb = vyb[iyb];
    if (with_current) {
      g.xc = vxb[ixc];
      g.yc = vyb[iyc];
    }
    if (with_p) g.p = p;
    g.bcW = bcW; g.bcE = bcE; g.bcN = bcN; g.bcS = bcS;
    g.y_lo = R.own_lo; g.y_hi = R.own_hi; g.do_s = first; g.do_n = last;
    launch_borders(g, stream, &lc);
  };

  // applyAccumulatedVelocity + diffuse (fused), then setVBCs
  {
    const float a = dt * mu * ((float)W - 1.0f) / pwidth; // simulation.cpp:105
    PrestepArgs g{};
    g.mask = mask0;
    g.H = H; g.pitch = pitch; g.a = a; g.rden = 1.0f / (1.0f + 4.0f * a);
    g.bcLo = bcW; g.bcHi = bcE; g.bcS = bcS; g.bcN = bcN;
    g.st_lo = R.st_lo; g.st_hi = R.st_hi; g.own_lo = R.own_lo; g.own_hi = R.own_hi;
    g.A = vxb[ixf].d; g.K = vxb[ixb].d; g.acc = vx_accum.d; g.B = vxb[ixb].d; g.Cout = vxb[ixc].d;
    g.gw = W - 1; g.gh = H;
    launch_prestep(0, g, stream, &lc);
    std::swap(ixf, ixc);
    g.A = vyb[iyf].d; g.K = vyb[iyf].d; g.acc = vy_accum.d; g.B = vyb[iyb].d; g.Cout = vyb[iyc].d;
    g.gw = W; g.gh = H - 1;
    launch_prestep(1, g, stream, &lc);
    std::swap(iyf, iyc);
  }
  borders(false, false);
  exchange({xf(vxb[ixf], 0), xf(vyb[iyf], 0)}, plan.ghost);

  // advect (reads the fronts incl. ghost rows, writes own rows of the backs)
  int adv = 0;
  {
    AdvectPeers ap{};
    ap.st_lo = R.st_lo; ap.st_hi = R.st_hi; ap.peer_lo = R.st_lo; ap.peer_hi = R.st_hi;
    ap.err = err;
    const size_t rb = sizeof(float) * (size_t)pitch;
    if (!first) {
      ap.peer_lo = plan.rows(0, plan.rank - 1).own_lo;
      ap.vx_lo = peer_ptr(plan.rank - 1, vxb[ixf].d, 0, rb);
      ap.vy_lo = peer_ptr(plan.rank - 1, vyb[iyf].d, 0, rb);
    }
    if (!last) {
      ap.peer_hi = plan.rows(0, plan.rank + 1).own_hi;
      ap.vx_hi = peer_ptr(plan.rank + 1, vxb[ixf].d, 0, rb);
      ap.vy_hi = peer_ptr(plan.rank + 1, vyb[iyf].d, 0, rb);
    }
    adv = launch_advect(vxb[ixf], vyb[iyf], vxb[ixb], vyb[iyb], flag, 0.5f * dt * ih, dt * ih, R.own_lo, R.own_hi,
                        plan.nranks > 1 ? &ap : nullptr, stream, &lc, mask0, vx_accum.d, vy_accum.d, f.d, ih);
  }
  std::swap(ixf, ixb);
  std::swap(iyf, iyb);
  borders(false, false);
  exchange({xf(vyb[iyf], 0)}, 1); // divergence reads vy(x, y-1)

  // project: divergence, sinks, V-cycles, setPBC, gradient
  {
    // the divergence pass also zeroes the accumulator interiors of the own rows
    // (simulation.cpp:384,392); the few ghost rows are cleared by memsets
    const Grid none{};
    const bool acc_zeroed = adv & ADV_ZEROED;
    if (adv & ADV_DIV) // the advect epilogue wrote f except on CTA edges and next to the BC faces
      launch_divergence_edges(vxb[ixf], vyb[iyf], f, ih, R.own_lo, R.own_hi, stream, &lc);
    else
      launch_divergence4(vxb[ixf], vyb[iyf], f, acc_zeroed ? none : vx_accum, acc_zeroed ? none : vy_accum, ih,
                         R.own_lo, R.own_hi, stream, &lc);
    auto clear_rows = [&](int y0, int y1) {
      y0 = std::max(1, y0);
      const int yx = std::min(H - 1, y1), yy = std::min(H - 2, y1);
      if (yx > y0)
        UBGL_CUDA(cudaMemset2DAsync(&vx_accum.at(1, y0), sizeof(float) * pitch, 0,
                                    sizeof(float) * (W - 3), yx - y0, stream));
      if (yy > y0)
        UBGL_CUDA(cudaMemset2DAsync(&vy_accum.at(1, y0), sizeof(float) * pitch, 0,
                                    sizeof(float) * (W - 2), yy - y0, stream));
    };
    clear_rows(R.st_lo, R.own_lo);
    clear_rows(R.own_hi, R.st_hi);
  }
  project_sinks();
  exchange({xf(f, 0)}, 8);
  for (int c = 0; c < vcycles; c++) {
    mg_solve();
    if (c + 1 < vcycles) exchange({xf(p, 0)}, 8);
  }
  // setPBC (columns of the own rows, rows on the edge ranks), THEN the p halo: the
  // neighbours' ghost copies must hold the BC'd border columns
  launch_pbc(p, bcW, bcE, bcN, bcS, R.own_lo, R.own_hi, first, last, stream, &lc);
  if (vcycles > 0) exchange({xf(p, 0)}, 8);
  const Grid none{};
  launch_gradient_save(vxb[ixf], vyb[iyf], p, mask0, lazy_current ? none : vxb[ixc], lazy_current ? none : vyb[iyc],
                       ih, R.own_lo, R.own_hi, stream, &lc);
  borders(false, !lazy_current);
  cur_alias = lazy_current;
  exchange({xf(vxb[ixf], 0), xf(vyb[iyf], 0)}, 4); // entry invariant of the next step
}

// The slab analogue of ubgl_sim_step_host (capi.cu), one rank's share: every rank drives its own
// PCIe link.  H->D: the accumulator mirrors, stored rows (ghost rows straight from the caller's
// arrays, no exchange).  D->H: the OWN rows of vx, vy, p, packed on the device to the
// reference's unpadded rows and sent as contiguous DMAs (a pitched 2-D copy is slower,
// tools/pcie_probe.py); vx_current / vy_current are byte copies of the final vx / vy
// (saveCurrentVelocityFields is a memcpy, simulation.cpp:16-19): host threads fill those
// mirrors from the vx / vy bands as they land instead of a second PCIe crossing, and clear the
// accumulator mirrors (:384,392) while the GPU steps.
void SlabSim::step_host(float dt_, const HostRows &m) {
  const Rows R = plan.rows(0);
  auto up = [&](const Grid &g, const float *host) {
    const int n = std::min(R.st_hi, g.h) - R.st_lo;
    UBGL_CUDA(cudaMemcpy2DAsync(&g.at(0, R.st_lo), sizeof(float) * g.pitch, host, sizeof(float) * g.w,
                                sizeof(float) * g.w, n, cudaMemcpyHostToDevice, stream));
  };
  if (m.vx_accum) up(vx_accum, m.vx_accum);
  if (m.vy_accum) up(vy_accum, m.vy_accum);
  cudaEvent_t uploaded = nullptr;
  std::vector<HostBand> bands;
  auto cleanup = [&]() {
    if (uploaded) cudaEventDestroy(uploaded);
    for (auto &b : bands)
      if (b.ready) cudaEventDestroy(b.ready);
  };
  try {
    if (m.vx_accum || m.vy_accum) {
      UBGL_CUDA(cudaEventCreateWithFlags(&uploaded, cudaEventDisableTiming));
      UBGL_CUDA(cudaEventRecord(uploaded, stream));
    }
    step(dt_);
    if (!d_pack && (m.vx || m.vy || m.p))
      UBGL_CUDA(cudaMalloc(&d_pack, sizeof(float) * (size_t)W * (R.own_hi - R.own_lo)));
    struct Out { const Grid *g; float *dst, *cur; };
    const Out outs[3] = {{&vxb[ixf], m.vx, m.vx_current}, {&vyb[iyf], m.vy, m.vy_current}, {&p, m.p, nullptr}};
    for (const Out &o : outs) {
      if (!o.dst && !o.cur) continue;
      const Grid &g = *o.g;
      const int y_lo = R.own_lo, y_hi = std::min(R.own_hi, g.h), n = y_hi - y_lo;
      if (n <= 0) continue;
      // the packed buffer is reused field after field: stream order keeps pack k+1 behind copy k
      launch_pack_rows(&g.at(0, y_lo), g.pitch, g.w, n, d_pack, stream, &lc);
      float *dst = o.dst ? o.dst : o.cur; // only the *_current mirror wanted: it takes the DMA itself
      const size_t off = (size_t)(y_lo - R.st_lo) * g.w, row = sizeof(float) * (size_t)g.w;
      const int nb = (o.dst && o.cur) ? std::min(16, n) : 1;
      for (int b = 0; b < nb; b++) {
        const int r0 = (int)((long long)n * b / nb), r1 = (int)((long long)n * (b + 1) / nb);
        UBGL_CUDA(cudaMemcpyAsync(dst + off + (size_t)r0 * g.w, d_pack + (size_t)r0 * g.w, row * (size_t)(r1 - r0),
                                  cudaMemcpyDeviceToHost, stream));
        if (o.dst && o.cur) {
          HostBand hb;
          UBGL_CUDA(cudaEventCreateWithFlags(&hb.ready, cudaEventDisableTiming));
          bands.push_back(hb);
          HostBand &k = bands.back();
          UBGL_CUDA(cudaEventRecord(k.ready, stream));
          k.src = o.dst + off + (size_t)r0 * g.w;
          k.dst = o.cur + off + (size_t)r0 * g.w;
          k.bytes = row * (size_t)(r1 - r0);
        }
      }
    }
    cudaError_t herr = cudaSuccess;
    if (uploaded || !bands.empty())
      host_side_work(device, uploaded, m.vx_accum, m.vy_accum, W, H, R.st_lo, R.own_lo, R.own_hi, bands, &herr);
    UBGL_CUDA(herr);
    UBGL_CUDA(cudaStreamSynchronize(stream));
    check_err();
  } catch (...) {
    cleanup();
    throw;
  }
  cleanup();
}

double SlabSim::residual_sumsq() {
  const Rows R = plan.rows(0);
  UBGL_CUDA(cudaMemsetAsync(d_sum, 0, sizeof(double), stream));
  const int y_lo = std::max(1, R.own_lo), y_hi = std::min(H - 1, R.own_hi);
  dim3 b(32, 8), g(ceil_div(W, 32), ceil_div(y_hi - y_lo, 8));
  UBGL_LAUNCH(&lc, K_RESIDUAL, 0, stream, k_slab_residual<<<g, b, 0, stream>>>(p, f, flag, 1.0f / h / h, y_lo, y_hi, d_sum));
  double v = 0.0;
  UBGL_CUDA(cudaMemcpyAsync(&v, d_sum, sizeof(double), cudaMemcpyDeviceToHost, stream));
  UBGL_CUDA(cudaStreamSynchronize(stream));
  check_err();
  return v;
}

} // namespace ubgl
