// capi.cu -- the extern "C" boundary declared in include/ubgl.h.  Everything
// behind it is C++/CUDA; nothing but plain C types and opaque handles crosses.
#include "../../include/ubgl.h"
#include "mg.cuh"
#include "sim.cuh"
#include "slab.cuh"
#include "capi_internal.cuh"
#include "hostwork.cuh"
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <thread>
#include <vector>

namespace ubgl {
const char *kind_name(int k) {
  static const char *names[K_COUNT] = {
      "other", "fill", "accum", "diffuse", "vbc", "advect", "divergence", "sinks", "rbgs_half",
      "zero_gradient_bc", "residual", "norm", "restrict", "prolong_correct", "coarsen_flag",
      "pbc", "gradient", "prestep_fused", "advect_div_fused", "mg_pre_fused", "mg_post_fused",
      "mg_coarse_fused", "finish_fused", "halo_push", "halo_wait", "colocate", "tracers", "items",
      "terrain"};
  return (k >= 0 && k < K_COUNT) ? names[k] : "?";
}
static thread_local std::string g_err;
void set_error(const std::string &m) { g_err = m; }
const char *get_error() { return g_err.c_str(); }
} // namespace ubgl

using namespace ubgl;

struct ubgl_mg {
  int W = 0, H = 0, device = 0;
  cudaStream_t stream = nullptr;
  LaunchCounter lc;
  std::unique_ptr<DeviceMG> mg;
  Grid p, f, flag;
  ~ubgl_mg() {
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    mg.reset();
    free_grid(p);
    free_grid(f);
    free_grid(flag);
    if (stream) cudaStreamDestroy(stream);
  }
};

extern "C" {

int ubgl_version(void) { return UBGL_VERSION; }
const char *ubgl_last_error(void) { return get_error(); }
int ubgl_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

// ---- Simulation -------------------------------------------------------------
int ubgl_sim_create(const float *flag, int W, int H, float pwidth, float mu, int device,
                    ubgl_sim_t **out) {
  UBGL_TRY
  NEED(out, "out");
  *out = nullptr;
  NEED(flag, "flag");
  require_device(device);
  std::unique_ptr<ubgl_sim> h(new ubgl_sim);
  h->s.reset(new DeviceSim(flag, W, H, pwidth, mu, device));
  *out = h.release();
  UBGL_CATCH
}

int ubgl_sim_destroy(ubgl_sim_t *sim) {
  UBGL_TRY
  delete sim;
  UBGL_CATCH
}

int ubgl_sim_set_option(ubgl_sim_t *sim, int option, int value) {
  UBGL_TRY
  SIM(sim);
  switch (option) {
  case UBGL_OPT_VCYCLES:
    UBGL_REQUIRE(value >= 0, "vcycles must be >= 0");
    S.vcycles = value;
    break;
  case UBGL_OPT_FUSED: S.fused = value != 0; S.mg->fused = value != 0; if (value) set_tile_variant(value); break;
  case UBGL_OPT_GRAPH: S.use_graph = value != 0; break;
  case UBGL_OPT_TIMING: S.timing = value != 0; break;
  default: throw ArgError{"unknown option"};
  }
  S.drop_graphs(); // captured launch sequences depend on every option
  UBGL_CATCH
}

int ubgl_sim_set_bc(ubgl_sim_t *sim, int west, int east, int north, int south) {
  UBGL_TRY
  SIM(sim);
  auto ok = [](int b) { return b >= 0 && b <= 3; };
  UBGL_REQUIRE(ok(west) && ok(east) && ok(north) && ok(south), "bad BC id");
  S.bcW = west; S.bcE = east; S.bcN = north; S.bcS = south;
  S.drop_graphs();
  UBGL_CATCH
}

int ubgl_sim_upload(ubgl_sim_t *sim, int field, const float *host) {
  UBGL_TRY
  SIM(sim);
  UBGL_REQUIRE(field >= 0 && field < UBGL_NUM_FIELDS, "bad field id");
  S.upload(field, host);
  UBGL_CATCH
}

int ubgl_sim_upload_add(ubgl_sim_t *sim, int field, const float *host) {
  UBGL_TRY
  SIM(sim);
  UBGL_REQUIRE(field >= 0 && field < UBGL_NUM_FIELDS, "bad field id");
  S.upload_add(field, host);
  UBGL_CATCH
}

int ubgl_sim_download(ubgl_sim_t *sim, int field, float *host) {
  UBGL_TRY
  SIM(sim);
  UBGL_REQUIRE(field >= 0 && field < UBGL_NUM_FIELDS, "bad field id");
  S.download(field, host);
  UBGL_CATCH
}

int ubgl_sim_update_flag(ubgl_sim_t *sim, const float *flag) {
  UBGL_TRY
  SIM(sim);
  S.update_flag(flag);
  UBGL_CATCH
}

int ubgl_sim_mg_levels(ubgl_sim_t *sim) { return sim ? sim->s->mg->levels() : UBGL_E_ARG; }

int ubgl_sim_mg_level_size(ubgl_sim_t *sim, int level, int *w, int *h) {
  UBGL_TRY
  SIM(sim);
  UBGL_REQUIRE(level >= 0 && level < S.mg->levels(), "bad level");
  NEED(w, "w");
  NEED(h, "h");
  *w = S.mg->level(level).w;
  *h = S.mg->level(level).h;
  UBGL_CATCH
}

int ubgl_sim_mg_get_flagc(ubgl_sim_t *sim, int level, float *host) {
  UBGL_TRY
  SIM(sim);
  UBGL_REQUIRE(level >= 0 && level < S.mg->levels(), "bad level");
  NEED(host, "host");
  const MGLevel &L = S.mg->level(level);
  download_grid(L.flagc, host, L.w, L.h, S.stream);
  S.sync();
  UBGL_CATCH
}

int ubgl_sim_set_sinks(ubgl_sim_t *sim, const float *xyz, int n) {
  UBGL_TRY
  SIM(sim);
  UBGL_REQUIRE(n >= 0 && (n == 0 || xyz), "bad sink list");
  S.sinks.resize(n);
  for (int i = 0; i < n; i++) S.sinks[i] = Sink{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
  UBGL_CATCH
}

int ubgl_sim_get_sinks(ubgl_sim_t *sim, float *xyz, int cap, int *n) {
  UBGL_TRY
  SIM(sim);
  NEED(n, "n");
  *n = (int)S.sinks.size();
  for (int i = 0; i < *n && i < cap && xyz; i++) {
    xyz[3 * i] = S.sinks[i].x;
    xyz[3 * i + 1] = S.sinks[i].y;
    xyz[3 * i + 2] = S.sinks[i].z;
  }
  UBGL_CATCH
}

int ubgl_sim_step(ubgl_sim_t *sim, float dt) {
  UBGL_TRY
  SIM(sim);
  S.step(dt);
  UBGL_CATCH
}

int ubgl_sim_stage(ubgl_sim_t *sim, int stage, float dt) {
  UBGL_TRY
  SIM(sim);
  S.stage(stage, dt);
  UBGL_CATCH
}

int ubgl_sim_step_host(ubgl_sim_t *sim, float dt, const ubgl_host_mirrors *m) {
  UBGL_TRY
  SIM(sim);
  NEED(m, "mirrors");
  if (m->flag) {
    Grid g = S.field(F_FLAG);
    upload_grid(g, m->flag, g.w, g.h, S.stream);
    S.flag_changed(true);
  }
  if (m->vx_accum) {
    Grid g = S.field(F_VX_ACCUM);
    upload_grid(g, m->vx_accum, g.w, g.h, S.stream);
  }
  if (m->vy_accum) {
    Grid g = S.field(F_VY_ACCUM);
    upload_grid(g, m->vy_accum, g.w, g.h, S.stream);
  }
  cudaEvent_t uploaded = nullptr;
  std::vector<HostBand> bands;
  auto cleanup = [&]() {
    if (uploaded) cudaEventDestroy(uploaded);
    for (auto &b : bands)
      if (b.ready) cudaEventDestroy(b.ready);
  };
  try {
    if (m->vx_accum || m->vy_accum) {
      UBGL_CUDA(cudaEventCreateWithFlags(&uploaded, cudaEventDisableTiming));
      UBGL_CUDA(cudaEventRecord(uploaded, S.stream));
    }
    S.step(dt); // asynchronous: every kernel of the step is now queued behind the uploads
    // velocity mirrors first: a component whose *_current mirror is wanted too comes down
    // in row bands, each followed by an event the host copy threads wait on; p last, so
    // the host copies of the final bands hide behind its transfer
    struct { int id, cur_id; float *dst, *cur; } vel[] = {{F_VX, F_VX_CURRENT, m->vx, m->vx_current},
                                                          {F_VY, F_VY_CURRENT, m->vy, m->vy_current}};
    for (auto &o : vel) {
      const bool big = o.dst && (size_t)S.field(o.id).h * S.field(o.id).w * sizeof(float) >= (size_t)(16 << 20);
      if (o.dst && o.cur && big) {
        // large field: one PCIe crossing, the *_current mirror is a multi-threaded host copy of
        // the bands as they land (for a few MB a second small DMA beats a single-core memcpy)
        Grid g = S.field(o.id);
        const size_t row = sizeof(float) * (size_t)g.w;
        const int nb = std::min(32, g.h);
        const float *pk = S.packed(g); // unpadded rows: contiguous DMA
        for (int b = 0; b < nb; b++) {
          const int y0 = (int)((long long)g.h * b / nb), y1 = (int)((long long)g.h * (b + 1) / nb);
          UBGL_CUDA(cudaMemcpyAsync(o.dst + (size_t)y0 * g.w, pk + (size_t)y0 * g.w, row * (size_t)(y1 - y0),
                                    cudaMemcpyDeviceToHost, S.stream));
          HostBand hb;
          UBGL_CUDA(cudaEventCreateWithFlags(&hb.ready, cudaEventDisableTiming));
          bands.push_back(hb);
          HostBand &k = bands.back();
          UBGL_CUDA(cudaEventRecord(k.ready, S.stream));
          k.src = o.dst + (size_t)y0 * g.w;
          k.dst = o.cur + (size_t)y0 * g.w;
          k.bytes = row * (size_t)(y1 - y0);
        }
      } else {
        if (o.dst) {
          Grid g = S.field(o.id);
          download_grid(g, o.dst, g.w, g.h, S.stream);
        }
        if (o.cur) {
          Grid g = S.field(o.cur_id);
          download_grid(g, o.cur, g.w, g.h, S.stream);
        }
      }
    }
    if (m->p) {
      Grid g = S.field(F_P);
      if (g.bytes() >= (size_t)(16 << 20) && g.w != g.pitch)
        UBGL_CUDA(cudaMemcpyAsync(m->p, S.packed(g), sizeof(float) * (size_t)g.w * g.h, cudaMemcpyDeviceToHost,
                                  S.stream));
      else
        download_grid(g, m->p, g.w, g.h, S.stream);
    }
    cudaError_t herr = cudaSuccess;
    if (uploaded || !bands.empty())
      host_side_work(S.device, uploaded, m->vx_accum, m->vy_accum, S.W, S.H, 0, 0, S.H, bands, &herr);
    UBGL_CUDA(herr);
    S.sync();
  } catch (...) {
    cleanup();
    throw;
  }
  cleanup();
  UBGL_CATCH
}

// ---------------------------------------------------------------------------
// Pipelined host-mirror step (optional mode).  ubgl_sim_step_host is PCIe-serial: accumulators
// up (10.5 ms at 8192^2), the step (4.5 ms), fields down (18 ms).  Here the three overlap across
// calls on PCIe's two directions: call n sends step n-1's staged outputs down on one copy stream
// while step n's accumulators go up on another and step n runs; the mirrors a call returns are
// therefore ONE STEP LATE (after call n they hold the fields of step n-1; the fields themselves
// are those of the synchronous call, bit for bit).  That is the coherence the reference's render
// thread already lives with: it reads sim.vx_current / sim.p while the simulation thread is
// somewhere inside step() (draw.cpp:101 vs sim_loop.cpp:29).
// ubgl_sim_step_host_flush brings the last step's fields down.
// ---------------------------------------------------------------------------
namespace {
void pipe_setup(DeviceSim &S) {
  if (S.s_in) return;
  UBGL_CUDA(cudaStreamCreateWithFlags(&S.s_in, cudaStreamNonBlocking));
  UBGL_CUDA(cudaStreamCreateWithFlags(&S.s_out, cudaStreamNonBlocking));
  UBGL_CUDA(cudaEventCreateWithFlags(&S.ev_up, cudaEventDisableTiming));
  UBGL_CUDA(cudaEventCreateWithFlags(&S.ev_pack, cudaEventDisableTiming));
  UBGL_CUDA(cudaEventCreateWithFlags(&S.ev_down, cudaEventDisableTiming));
  UBGL_CUDA(cudaEventCreateWithFlags(&S.ev_main, cudaEventDisableTiming));
  const size_t n = (size_t)(S.W - 1) * S.H + (size_t)S.W * (S.H - 1) + (size_t)S.W * S.H;
  UBGL_CUDA(cudaMalloc(&S.d_out, sizeof(float) * n));
}

// D->H of the staged outputs of the previous step into the caller's mirrors (copy stream s_out),
// vx / vy in bands with events so that host threads can fill the *_current mirrors behind them
void pipe_download(DeviceSim &S, const ubgl_host_mirrors *m, std::vector<HostBand> &bands) {
  UBGL_CUDA(cudaStreamWaitEvent(S.s_out, S.ev_pack, 0));
  const size_t nvx = (size_t)(S.W - 1) * S.H, nvy = (size_t)S.W * (S.H - 1), np = (size_t)S.W * S.H;
  struct { const float *src; float *dst, *cur; int w, h; } out[3] = {
      {S.d_out, m->vx, m->vx_current, S.W - 1, S.H},
      {S.d_out + nvx, m->vy, m->vy_current, S.W, S.H - 1},
      {S.d_out + nvx + nvy, m->p, nullptr, S.W, S.H}};
  (void)np;
  for (auto &o : out) {
    float *dst = o.dst ? o.dst : o.cur;
    if (!dst) continue;
    const size_t row = sizeof(float) * (size_t)o.w;
    const int nb = (o.dst && o.cur) ? std::min(32, o.h) : 1;
    for (int b = 0; b < nb; b++) {
      const int y0 = (int)((long long)o.h * b / nb), y1 = (int)((long long)o.h * (b + 1) / nb);
      UBGL_CUDA(cudaMemcpyAsync(dst + (size_t)y0 * o.w, o.src + (size_t)y0 * o.w, row * (size_t)(y1 - y0),
                                cudaMemcpyDeviceToHost, S.s_out));
      if (o.dst && o.cur) {
        HostBand hb;
        UBGL_CUDA(cudaEventCreateWithFlags(&hb.ready, cudaEventDisableTiming));
        bands.push_back(hb);
        HostBand &k = bands.back();
        UBGL_CUDA(cudaEventRecord(k.ready, S.s_out));
        k.src = o.dst + (size_t)y0 * o.w;
        k.dst = o.cur + (size_t)y0 * o.w;
        k.bytes = row * (size_t)(y1 - y0);
      }
    }
  }
  UBGL_CUDA(cudaEventRecord(S.ev_down, S.s_out));
}
} // namespace

int ubgl_sim_step_host_pipelined(ubgl_sim_t *sim, float dt, const ubgl_host_mirrors *m) {
  UBGL_TRY
  SIM(sim);
  NEED(m, "mirrors");
  UBGL_REQUIRE(m->flag == nullptr, "pipelined step: upload a new flag with ubgl_sim_update_flag between calls");
  pipe_setup(S);
  std::vector<HostBand> bands;
  auto cleanup = [&]() {
    for (auto &b : bands)
      if (b.ready) cudaEventDestroy(b.ready);
  };
  try {
    const bool had = S.pipe_pending;
    if (had) pipe_download(S, m, bands); // step n-1 goes down ...
    bool up = false;
    // the accumulators may be overwritten only after everything queued so far (step n-1 reads and
    // clears them) is through
    UBGL_CUDA(cudaEventRecord(S.ev_main, S.stream));
    UBGL_CUDA(cudaStreamWaitEvent(S.s_in, S.ev_main, 0));
    if (m->vx_accum) { // ... while step n's accumulators come up on the other PCIe direction
      Grid g = S.field(F_VX_ACCUM);
      upload_grid(g, m->vx_accum, g.w, g.h, S.s_in);
      up = true;
    }
    if (m->vy_accum) {
      Grid g = S.field(F_VY_ACCUM);
      upload_grid(g, m->vy_accum, g.w, g.h, S.s_in);
      up = true;
    }
    if (up) {
      UBGL_CUDA(cudaEventRecord(S.ev_up, S.s_in));
      UBGL_CUDA(cudaStreamWaitEvent(S.stream, S.ev_up, 0));
    }
    S.step(dt);
    // the staging buffer is free once the previous download has read it
    if (had) UBGL_CUDA(cudaStreamWaitEvent(S.stream, S.ev_down, 0));
    {
      const size_t nvx = (size_t)(S.W - 1) * S.H, nvy = (size_t)S.W * (S.H - 1);
      Grid gx = S.field(F_VX), gy = S.field(F_VY), gp = S.field(F_P);
      launch_pack_rows(gx.d, gx.pitch, gx.w, gx.h, S.d_out, S.stream, &S.lc);
      launch_pack_rows(gy.d, gy.pitch, gy.w, gy.h, S.d_out + nvx, S.stream, &S.lc);
      launch_pack_rows(gp.d, gp.pitch, gp.w, gp.h, S.d_out + nvx + nvy, S.stream, &S.lc);
      UBGL_CUDA(cudaEventRecord(S.ev_pack, S.stream));
      S.pipe_pending = true;
    }
    static const bool dbg = getenv("UBGL_PIPE_DEBUG") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    auto ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
    double t_up = 0, t_host = 0, t_down = 0;
    if (dbg && up) {
      cudaEventSynchronize(S.ev_up);
      t_up = ms();
    }
    cudaError_t herr = cudaSuccess;
    if (up || !bands.empty())
      host_side_work(S.device, up ? S.ev_up : nullptr, m->vx_accum, m->vy_accum, S.W, S.H, 0, 0, S.H, bands, &herr);
    UBGL_CUDA(herr);
    t_host = ms();
    if (had) UBGL_CUDA(cudaEventSynchronize(S.ev_down)); // the mirrors now hold step n-1; step n keeps running
    else if (up) UBGL_CUDA(cudaEventSynchronize(S.ev_up));
    t_down = ms();
    if (dbg) fprintf(stderr, "[pipe] upload done %.2f ms, host work done %.2f ms, download done %.2f ms\n", t_up, t_host, t_down);
  } catch (...) {
    cleanup();
    throw;
  }
  cleanup();
  UBGL_CATCH
}

int ubgl_sim_step_host_flush(ubgl_sim_t *sim, const ubgl_host_mirrors *m) {
  UBGL_TRY
  SIM(sim);
  NEED(m, "mirrors");
  if (!S.pipe_pending) return UBGL_OK;
  std::vector<HostBand> bands;
  auto cleanup = [&]() {
    for (auto &b : bands)
      if (b.ready) cudaEventDestroy(b.ready);
  };
  try {
    pipe_download(S, m, bands);
    cudaError_t herr = cudaSuccess;
    if (!bands.empty()) host_side_work(S.device, nullptr, nullptr, nullptr, S.W, S.H, 0, 0, 0, bands, &herr);
    UBGL_CUDA(herr);
    UBGL_CUDA(cudaEventSynchronize(S.ev_down));
    S.pipe_pending = false;
    S.sync();
  } catch (...) {
    cleanup();
    throw;
  }
  cleanup();
  UBGL_CATCH
}

int ubgl_sim_set_tolerance(ubgl_sim_t *sim, float rel_tol, int max_cycles, float stagnation) {
  UBGL_TRY
  SIM(sim);
  UBGL_REQUIRE(rel_tol <= 0.0f || (max_cycles >= 1 && stagnation > 0.0f && stagnation <= 1.0f),
               "tolerance mode needs max_cycles >= 1 and 0 < stagnation <= 1");
  S.tol = rel_tol;
  if (rel_tol > 0.0f) {
    S.max_cycles = max_cycles;
    S.stag = stagnation;
  }
  UBGL_CATCH
}

int ubgl_sim_solve_info(ubgl_sim_t *sim, int *cycles_done, float *fnorm, float *res_hist, int cap,
                        int *n_hist) {
  UBGL_TRY
  SIM(sim);
  if (cycles_done) *cycles_done = S.cycles_done;
  if (fnorm) *fnorm = S.fnorm;
  if (n_hist) *n_hist = (int)S.res_hist.size();
  for (int i = 0; res_hist && i < cap && i < (int)S.res_hist.size(); i++) res_hist[i] = S.res_hist[i];
  UBGL_CATCH
}

int ubgl_sim_sync(ubgl_sim_t *sim) {
  UBGL_TRY
  SIM(sim);
  S.sync();
  UBGL_CATCH
}

int ubgl_sim_residual(ubgl_sim_t *sim, float *l2) {
  UBGL_TRY
  SIM(sim);
  NEED(l2, "l2");
  *l2 = S.residual();
  UBGL_CATCH
}

int ubgl_sim_mg_solve(ubgl_sim_t *sim, int cycles) {
  UBGL_TRY
  SIM(sim);
  UBGL_REQUIRE(cycles >= 0, "cycles must be >= 0");
  S.mg_solve(cycles);
  UBGL_CATCH
}

int ubgl_sim_mg_solve_ex(ubgl_sim_t *sim, float h, int zero_gradient_bc, int cycles) {
  UBGL_TRY
  SIM(sim);
  UBGL_REQUIRE(cycles >= 0 && h > 0.0f, "bad cycles / h");
  S.mg_solve_ex(h, zero_gradient_bc != 0, cycles);
  UBGL_CATCH
}

int ubgl_sim_device_ptr(ubgl_sim_t *sim, int field, void **dptr, int *pitch) {
  UBGL_TRY
  SIM(sim);
  NEED(dptr, "dptr");
  UBGL_REQUIRE(field >= 0 && field < UBGL_NUM_FIELDS, "bad field id");
  S.will_write(field); // the caller may write through the pointer: front and *_current get their own buffers
  Grid g = S.field(field);
  *dptr = g.d;
  if (pitch) *pitch = g.pitch;
  UBGL_CATCH
}

int ubgl_sim_stage_ms(ubgl_sim_t *sim, int stage, float *ms) {
  UBGL_TRY
  SIM(sim);
  NEED(ms, "ms");
  UBGL_REQUIRE(stage >= 0 && stage < ST_COUNT, "bad stage id");
  *ms = S.stage_ms[stage];
  UBGL_CATCH
}

long long ubgl_sim_launch_count(ubgl_sim_t *sim) { return sim ? sim->s->lc.n : -1; }

int ubgl_sim_profile(ubgl_sim_t *sim, int on) {
  UBGL_TRY
  SIM(sim);
  S.sync();
  S.lc.collect();
  if (on) S.lc.reset_stats();
  S.lc.prof = on != 0;
  UBGL_CATCH
}

int ubgl_num_kernel_kinds(void) { return K_COUNT; }
const char *ubgl_kernel_kind_name(int kind) { return kind_name(kind); }

int ubgl_sim_kernel_stats(ubgl_sim_t *sim, int kind, int level, long long *count, double *ms) {
  UBGL_TRY
  SIM(sim);
  UBGL_REQUIRE(kind >= 0 && kind < K_COUNT && level >= 0 && level < LaunchCounter::MAXLVL,
               "bad kind/level");
  S.sync();
  S.lc.collect();
  if (count) *count = S.lc.cnt[kind][level];
  if (ms) *ms = S.lc.ms[kind][level];
  UBGL_CATCH
}
void *ubgl_sim_stream(ubgl_sim_t *sim) { return sim ? (void *)sim->s->stream : nullptr; }

// ---- MG ---------------------------------------------------------------------
int ubgl_mg_create(int W, int H, int device, ubgl_mg_t **out) {
  UBGL_TRY
  NEED(out, "out");
  *out = nullptr;
  UBGL_REQUIRE(W >= 8 && H >= 8, "MG needs W,H >= 8 (two levels)");
  require_device(device);
  std::unique_ptr<ubgl_mg> m(new ubgl_mg);
  m->W = W;
  m->H = H;
  m->device = device;
  UBGL_CUDA(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
  m->mg.reset(new DeviceMG(W, H, device, m->stream, &m->lc));
  int pitch = m->mg->level(0).pitch;
  m->p = alloc_grid(W, H, pitch);
  m->f = alloc_grid(W, H, pitch);
  m->flag = alloc_grid(W, H, pitch, false);
  fill_grid(m->flag, 1.0f, m->stream, nullptr);
  UBGL_CUDA(cudaStreamSynchronize(m->stream));
  *out = m.release();
  UBGL_CATCH
}

int ubgl_mg_destroy(ubgl_mg_t *mg) {
  UBGL_TRY
  delete mg;
  UBGL_CATCH
}

#define MGH(mg)                                                                \
  NEED(mg, "mg");                                                              \
  ubgl_mg &M = *(mg);                                                          \
  UBGL_CUDA(cudaSetDevice(M.device));

int ubgl_mg_set_option(ubgl_mg_t *mg, int option, int value) {
  UBGL_TRY
  MGH(mg);
  switch (option) {
  case UBGL_OPT_FUSED: M.mg->fused = value != 0; if (value) set_tile_variant(value); break;
  case UBGL_OPT_GRAPH: case UBGL_OPT_TIMING: case UBGL_OPT_VCYCLES: break;
  default: throw ArgError{"unknown option"};
  }
  UBGL_CATCH
}

int ubgl_mg_levels(ubgl_mg_t *mg) { return mg ? mg->mg->levels() : UBGL_E_ARG; }

int ubgl_mg_level_size(ubgl_mg_t *mg, int level, int *w, int *h) {
  UBGL_TRY
  MGH(mg);
  UBGL_REQUIRE(level >= 0 && level < M.mg->levels(), "bad level");
  NEED(w, "w");
  NEED(h, "h");
  *w = M.mg->level(level).w;
  *h = M.mg->level(level).h;
  UBGL_CATCH
}

int ubgl_mg_update_fields(ubgl_mg_t *mg, const float *flag) {
  UBGL_TRY
  MGH(mg);
  NEED(flag, "flag");
  upload_grid(M.flag, flag, M.W, M.H, M.stream);
  M.mg->update_fields(M.flag);
  M.mg->invalidate_mask0();
  UBGL_CUDA(cudaStreamSynchronize(M.stream));
  UBGL_CATCH
}

int ubgl_mg_get_flagc(ubgl_mg_t *mg, int level, float *host) {
  UBGL_TRY
  MGH(mg);
  UBGL_REQUIRE(level >= 0 && level < M.mg->levels(), "bad level");
  NEED(host, "host");
  const MGLevel &L = M.mg->level(level);
  download_grid(L.flagc, host, L.w, L.h, M.stream);
  UBGL_CUDA(cudaStreamSynchronize(M.stream));
  UBGL_CATCH
}

int ubgl_mg_upload(ubgl_mg_t *mg, const float *p, const float *f, const float *flag) {
  UBGL_TRY
  MGH(mg);
  if (p) upload_grid(M.p, p, M.W, M.H, M.stream);
  if (f) upload_grid(M.f, f, M.W, M.H, M.stream);
  if (flag) {
    upload_grid(M.flag, flag, M.W, M.H, M.stream);
    M.mg->invalidate_mask0();
  }
  UBGL_CUDA(cudaStreamSynchronize(M.stream));
  UBGL_CATCH
}

int ubgl_mg_download_p(ubgl_mg_t *mg, float *p) {
  UBGL_TRY
  MGH(mg);
  NEED(p, "p");
  download_grid(M.p, p, M.W, M.H, M.stream);
  UBGL_CUDA(cudaStreamSynchronize(M.stream));
  UBGL_CATCH
}

int ubgl_mg_solve(ubgl_mg_t *mg, float h, int zero_gradient_bc, int cycles) {
  UBGL_TRY
  MGH(mg);
  UBGL_REQUIRE(cycles >= 0, "cycles must be >= 0");
  for (int c = 0; c < cycles; c++) M.mg->solve(M.p, M.f, M.flag, h, zero_gradient_bc != 0);
  UBGL_CATCH
}

int ubgl_mg_solve_host(ubgl_mg_t *mg, float *p, const float *f, const float *flag, float h,
                       int zero_gradient_bc) {
  UBGL_TRY
  MGH(mg);
  NEED(p, "p");
  NEED(f, "f");
  NEED(flag, "flag");
  upload_grid(M.p, p, M.W, M.H, M.stream);
  upload_grid(M.f, f, M.W, M.H, M.stream);
  upload_grid(M.flag, flag, M.W, M.H, M.stream);
  M.mg->invalidate_mask0();
  M.mg->solve(M.p, M.f, M.flag, h, zero_gradient_bc != 0);
  download_grid(M.p, p, M.W, M.H, M.stream);
  UBGL_CUDA(cudaStreamSynchronize(M.stream));
  UBGL_CATCH
}

int ubgl_mg_residual(ubgl_mg_t *mg, float h, float *l2) {
  UBGL_TRY
  MGH(mg);
  NEED(l2, "l2");
  Grid r = alloc_grid(M.W, M.H, M.p.pitch, false);
  try {
    M.mg->residual(M.p, M.f, M.flag, r, h, true);
    *l2 = M.mg->residual_norm_result();
  } catch (...) {
    free_grid(r);
    throw;
  }
  free_grid(r);
  UBGL_CATCH
}

int ubgl_mg_sync(ubgl_mg_t *mg) {
  UBGL_TRY
  MGH(mg);
  UBGL_CUDA(cudaStreamSynchronize(M.stream));
  UBGL_CATCH
}

long long ubgl_mg_launch_count(ubgl_mg_t *mg) { return mg ? mg->lc.n : -1; }
void *ubgl_mg_stream(ubgl_mg_t *mg) { return mg ? (void *)mg->stream : nullptr; }

// ---- slabs ------------------------------------------------------------------
struct ubgl_slab {
  std::unique_ptr<SlabSim> s;
};
#define SLAB(h)                                                                \
  NEED(h, "slab");                                                             \
  SlabSim &S = *(h)->s;                                                        \
  UBGL_CUDA(cudaSetDevice(S.device));

int ubgl_slab_plan(int W, int H, int nranks, int rank, int *dist_levels, int *ghost, int *own_lo,
                   int *own_hi, int *st_lo, int *st_hi) {
  UBGL_TRY
  SlabPlan P = make_slab_plan(W, H, nranks, rank);
  Rows R = P.rows(0);
  if (dist_levels) *dist_levels = P.ndist;
  if (ghost) *ghost = P.ghost;
  if (own_lo) *own_lo = R.own_lo;
  if (own_hi) *own_hi = R.own_hi;
  if (st_lo) *st_lo = R.st_lo;
  if (st_hi) *st_hi = R.st_hi;
  UBGL_CATCH
}

int ubgl_slab_set_row_weights(const float *weights, int H) {
  UBGL_TRY
  set_slab_row_weights(weights, H);
  UBGL_CATCH
}

int ubgl_slab_create(const float *flag, int W, int H, float pwidth, float mu, int device, int rank,
                     int nranks, ubgl_slab_t **out) {
  UBGL_TRY
  NEED(out, "out");
  *out = nullptr;
  NEED(flag, "flag");
  require_device(device);
  std::unique_ptr<ubgl_slab> h(new ubgl_slab);
  h->s.reset(new SlabSim(flag, W, H, pwidth, mu, device, rank, nranks));
  *out = h.release();
  UBGL_CATCH
}

int ubgl_slab_destroy(ubgl_slab_t *s) {
  UBGL_TRY
  delete s;
  UBGL_CATCH
}

int ubgl_slab_ipc_size(void) { return 64; }

int ubgl_slab_ipc_export(ubgl_slab_t *s, void *blob) {
  UBGL_TRY
  SLAB(s);
  NEED(blob, "blob");
  S.ipc_export(blob);
  UBGL_CATCH
}

int ubgl_slab_connect(ubgl_slab_t *s, const void *blobs) {
  UBGL_TRY
  SLAB(s);
  NEED(blobs, "blobs");
  S.connect(blobs);
  S.finish_setup();
  UBGL_CATCH
}

int ubgl_slab_field_rows(ubgl_slab_t *s, int field, int *row_lo, int *nrows, int *w) {
  UBGL_TRY
  SLAB(s);
  NEED(row_lo, "row_lo"); NEED(nrows, "nrows"); NEED(w, "w");
  UBGL_REQUIRE(field >= 0 && field < UBGL_NUM_FIELDS, "bad field id");
  S.field_rows(field, row_lo, nrows, w);
  UBGL_CATCH
}

int ubgl_slab_upload(ubgl_slab_t *s, int field, const float *host) {
  UBGL_TRY
  SLAB(s);
  S.upload(field, host);
  UBGL_CATCH
}

int ubgl_slab_download(ubgl_slab_t *s, int field, float *host) {
  UBGL_TRY
  SLAB(s);
  S.download(field, host);
  UBGL_CATCH
}

int ubgl_slab_step_host(ubgl_slab_t *s, float dt, const ubgl_host_mirrors *m) {
  UBGL_TRY
  SLAB(s);
  NEED(m, "mirrors");
  UBGL_REQUIRE(m->flag == nullptr, "slab: a flag change is ubgl_slab_upload(UBGL_FLAG) + reconnect, not part of step_host");
  S.step_host(dt, SlabSim::HostRows{m->vx_accum, m->vy_accum, m->vx, m->vy, m->p, m->vx_current, m->vy_current});
  UBGL_CATCH
}

int ubgl_slab_set_option(ubgl_slab_t *s, int option, int value) {
  UBGL_TRY
  SLAB(s);
  UBGL_REQUIRE(option == UBGL_OPT_VCYCLES && value >= 0, "slab: only UBGL_OPT_VCYCLES >= 0");
  S.vcycles = value;
  UBGL_CATCH
}

int ubgl_slab_set_sinks(ubgl_slab_t *s, const float *xyz, int n) {
  UBGL_TRY
  SLAB(s);
  UBGL_REQUIRE(n >= 0 && (n == 0 || xyz), "bad sink list");
  S.sinks.resize(n);
  for (int i = 0; i < n; i++) S.sinks[i] = Sink{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
  UBGL_CATCH
}

int ubgl_slab_step(ubgl_slab_t *s, float dt) {
  UBGL_TRY
  SLAB(s);
  S.step(dt);
  UBGL_CATCH
}

int ubgl_slab_sync(ubgl_slab_t *s) {
  UBGL_TRY
  SLAB(s);
  S.sync();
  UBGL_CATCH
}

int ubgl_slab_residual_sumsq(ubgl_slab_t *s, double *sumsq) {
  UBGL_TRY
  SLAB(s);
  NEED(sumsq, "sumsq");
  *sumsq = S.residual_sumsq();
  UBGL_CATCH
}

long long ubgl_slab_launch_count(ubgl_slab_t *s) { return s ? s->s->lc.n : -1; }
void *ubgl_slab_stream(ubgl_slab_t *s) { return s ? (void *)s->s->stream : nullptr; }

int ubgl_slab_stats(ubgl_slab_t *s, long long *exchanges, long long *halo_bytes) {
  UBGL_TRY
  SLAB(s);
  if (exchanges) *exchanges = S.exchanges;
  if (halo_bytes) *halo_bytes = (long long)S.halo_bytes;
  UBGL_CATCH
}

int ubgl_slab_profile(ubgl_slab_t *s, int on) {
  UBGL_TRY
  SLAB(s);
  S.sync();
  S.lc.collect();
  if (on) S.lc.reset_stats();
  S.lc.prof = on != 0;
  UBGL_CATCH
}

int ubgl_slab_kernel_stats(ubgl_slab_t *s, int kind, int level, long long *count, double *ms) {
  UBGL_TRY
  SLAB(s);
  UBGL_REQUIRE(kind >= 0 && kind < K_COUNT && level >= 0 && level < LaunchCounter::MAXLVL,
               "bad kind/level");
  S.sync();
  S.lc.collect();
  if (count) *count = S.lc.cnt[kind][level];
  if (ms) *ms = S.lc.ms[kind][level];
  UBGL_CATCH
}

// ---- free operators with host grids ------------------------------------------
namespace {
// small RAII scratch context for the one-shot operator calls
struct Scratch {
  cudaStream_t stream = nullptr;
  LaunchCounter lc;
  std::vector<Grid> grids;
  std::unique_ptr<DeviceMG> mg;
  Scratch(int w, int h) {
    require_device(0);
    UBGL_REQUIRE(w >= 8 && h >= 8, "operator grids need w,h >= 8");
    UBGL_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    mg.reset(new DeviceMG(w, h, 0, stream, &lc));
  }
  Grid up(const float *host, int w, int h) {
    Grid g = alloc_grid(w, h, round_up(w, 32));
    grids.push_back(g);
    if (host) upload_grid(g, host, w, h, stream);
    return g;
  }
  void down(const Grid &g, float *host) {
    download_grid(g, host, g.w, g.h, stream);
    UBGL_CUDA(cudaStreamSynchronize(stream));
  }
  ~Scratch() {
    if (stream) cudaStreamSynchronize(stream);
    mg.reset();
    for (auto &g : grids) free_grid(g);
    if (stream) cudaStreamDestroy(stream);
  }
};
} // namespace

int ubgl_rbgs(float *p, const float *f, const float *flag, int w, int h, float hh, float alpha,
              int sweeps) {
  UBGL_TRY
  NEED(p, "p"); NEED(f, "f"); NEED(flag, "flag");
  Scratch sc(w, h);
  Grid P = sc.up(p, w, h), F = sc.up(f, w, h), G = sc.up(flag, w, h);
  for (int i = 0; i < sweeps; i++) sc.mg->rbgs(P, F, G, hh, alpha);
  sc.down(P, p);
  UBGL_CATCH
}

int ubgl_residual(const float *p, const float *f, const float *flag, float *r, int w, int h,
                  float hh, float *l2) {
  UBGL_TRY
  NEED(p, "p"); NEED(f, "f"); NEED(flag, "flag"); NEED(r, "r");
  Scratch sc(w, h);
  Grid P = sc.up(p, w, h), F = sc.up(f, w, h), G = sc.up(flag, w, h), R = sc.up(nullptr, w, h);
  sc.mg->residual(P, F, G, R, hh, true);
  float n = sc.mg->residual_norm_result();
  if (l2) *l2 = n;
  sc.down(R, r);
  UBGL_CATCH
}

int ubgl_restrict(const float *r, int w, int h, float *rc) {
  UBGL_TRY
  NEED(r, "r"); NEED(rc, "rc");
  Scratch sc(w, h);
  Grid R = sc.up(r, w, h), RC = sc.up(nullptr, w / 2, h / 2);
  sc.mg->restrict_fw(R, RC);
  sc.down(RC, rc);
  UBGL_CATCH
}

int ubgl_prolongate(float *e, int w, int h, const float *ec, const float *flagc,
                    const float *flag) {
  UBGL_TRY
  NEED(e, "e"); NEED(ec, "ec"); NEED(flagc, "flagc"); NEED(flag, "flag");
  Scratch sc(w, h);
  Grid E = sc.up(nullptr, w, h), EC = sc.up(ec, w / 2, h / 2), GC = sc.up(flagc, w / 2, h / 2),
       G = sc.up(flag, w, h);
  sc.mg->prolongate(E, EC, GC, G);
  sc.down(E, e);
  UBGL_CATCH
}

int ubgl_correct(float *p, const float *e, int w, int h) {
  UBGL_TRY
  NEED(p, "p"); NEED(e, "e");
  Scratch sc(w, h);
  Grid P = sc.up(p, w, h), E = sc.up(e, w, h);
  sc.mg->correct(P, E);
  sc.down(P, p);
  UBGL_CATCH
}

int ubgl_zero_gradient_bc(float *p, int w, int h) {
  UBGL_TRY
  NEED(p, "p");
  Scratch sc(w, h);
  Grid P = sc.up(p, w, h);
  sc.mg->zero_gradient_bc(P);
  sc.down(P, p);
  UBGL_CATCH
}

} // extern "C"
