// common.cuh -- device-grid descriptor, error plumbing and launch helpers shared
// by every translation unit of libubgl.so (sm_100a only, no CPU fallback).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <utility>
#include <vector>
#include <atomic>

namespace ubgl {

// A 2-D fp32 field resident in HBM.  Row-major like the reference's
// Single2DGrid (db2dgrid.hpp:19) but PITCHED: every row starts on a 128-byte
// boundary (pitch is a multiple of 32 floats) so that x = 0 mod 4 is 16-byte
// aligned and warps read whole 128-byte lines.
struct Grid {
  float *d = nullptr;
  int w = 0, h = 0, pitch = 0;
  __host__ __device__ __forceinline__ float &at(int x, int y) const {
    return d[(size_t)y * pitch + x];
  }
  size_t bytes() const { return sizeof(float) * (size_t)pitch * h; }
};

// Row-slab decomposition (csrc/slab.cu): of a level's rows, [st_lo, st_hi) are
// stored on this GPU (own rows + ghost rows) and [own_lo, own_hi) are computed
// here.  Slab grids keep GLOBAL row indices: Grid::d is the virtual address of
// row 0, so kernels index them exactly like single-GPU grids.
struct Rows {
  int st_lo, st_hi, own_lo, own_hi;
};

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

void set_error(const std::string &msg);
const char *get_error();

struct CudaError {
  cudaError_t code;
  const char *file;
  int line;
};

#define UBGL_CUDA(call)                                                        \
  do {                                                                         \
    cudaError_t e_ = (call);                                                   \
    if (e_ != cudaSuccess) throw ::ubgl::CudaError{e_, __FILE__, __LINE__};    \
  } while (0)

#define UBGL_CHECK_LAUNCH() UBGL_CUDA(cudaGetLastError())

struct ArgError {
  std::string msg;
};
#define UBGL_REQUIRE(cond, text)                                               \
  do {                                                                         \
    if (!(cond)) throw ::ubgl::ArgError{std::string(text)};                    \
  } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute of a kernel: a handle
// created on another GPU of the same process needs its own opt-in.  One bit per device ordinal;
// every C-ABI entry point has made the handle's device current before a launcher runs.
template <class K>
inline void ensure_dyn_smem(K kernel, size_t bytes, std::atomic<unsigned long long> &done) {
  int dev = 0;
  UBGL_CUDA(cudaGetDevice(&dev));
  const unsigned long long bit = 1ull << (dev & 63);
  if (done.load(std::memory_order_acquire) & bit) return;
  UBGL_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  done.fetch_or(bit, std::memory_order_release);
}

// Kernel kinds for the per-kernel profile (ubgl_sim_kernel_stats).
enum Kind {
  K_OTHER = 0, K_FILL, K_ACCUM, K_DIFFUSE, K_VBC, K_ADVECT, K_DIVERGENCE, K_SINKS, K_RBGS,
  K_ZGBC, K_RESIDUAL, K_NORM, K_RESTRICT, K_PROLONG, K_COARSEN, K_PBC, K_GRADIENT,
  K_PRESTEP /* fused accum+diffuse */, K_ADVDIV /* fused advect+BC+divergence */,
  K_MG_PRE /* fused smooth+residual+restrict */, K_MG_POST /* fused prolong+correct+smooth */,
  K_MG_COARSE /* all coarse levels in one kernel */, K_FINISH /* fused pBC+gradient+vBC+save */,
  K_HALO_PUSH /* peer-store halo rows + signal */, K_HALO_WAIT,
  K_COLOCATE /* staggered -> co-located velocity texture */, K_TRACERS, K_ITEMS /* floating items */,
  K_TERRAIN /* flag edits, scroll */,
  K_COUNT
};
const char *kind_name(int kind);

// Counts kernel launches per handle ("gpu_launches" in bench.py) and, when
// prof is on, brackets every launch with CUDA events on the launching stream so
// that bench.py can report per-kernel device time (kind x MG level).
struct LaunchCounter {
  static const int MAXLVL = 16;
  long long n = 0;
  bool prof = false;
  struct Rec {
    cudaEvent_t a, b;
    int kind, level;
  };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  double ms[K_COUNT][MAXLVL] = {};
  long long cnt[K_COUNT][MAXLVL] = {};

  cudaEvent_t get_event() {
    if (!pool.empty()) {
      cudaEvent_t e = pool.back();
      pool.pop_back();
      return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
  void pre(int kind, int level, cudaStream_t s) {
    n++;
    if (!prof) return;
    Rec r{get_event(), get_event(), kind, level < MAXLVL ? level : MAXLVL - 1};
    cudaEventRecord(r.a, s);
    recs.push_back(r);
  }
  void post(cudaStream_t s) {
    if (prof) cudaEventRecord(recs.back().b, s);
  }
  // call after a stream synchronize
  void collect() {
    // UBGL_TIMELINE=<prefix>: start offset and duration of every profiled launch of this collect
    // (<prefix>.<RANK>.csv, rewritten each time) -- where a rank idles between its kernels
    if (const char *tl = getenv("UBGL_TIMELINE")) {
      if (!recs.empty()) {
        const char *rk = getenv("RANK");
        std::string path = std::string(tl) + "." + (rk ? rk : "0") + ".csv";
        if (FILE *fp = fopen(path.c_str(), "w")) {
          fprintf(fp, "kind,level,start_ms,dur_ms\n");
          for (auto &r : recs) {
            float t0 = 0.f, t = 0.f;
            if (cudaEventElapsedTime(&t0, recs[0].a, r.a) == cudaSuccess &&
                cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess)
              fprintf(fp, "%d,%d,%.4f,%.4f\n", r.kind, r.level, t0, t);
          }
          fclose(fp);
        }
      }
    }
    for (auto &r : recs) {
      float t = 0.f;
      if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
        ms[r.kind][r.level] += t;
        cnt[r.kind][r.level]++;
      }
      pool.push_back(r.a);
      pool.push_back(r.b);
    }
    recs.clear();
  }
  void reset_stats() {
    for (int k = 0; k < K_COUNT; k++)
      for (int l = 0; l < MAXLVL; l++) ms[k][l] = 0.0, cnt[k][l] = 0;
  }
  ~LaunchCounter() {
    for (auto &r : recs) {
      cudaEventDestroy(r.a);
      cudaEventDestroy(r.b);
    }
    for (auto e : pool) cudaEventDestroy(e);
  }
};

// ---------------------------------------------------------------------------
// Programmatic dependent launch (sm_90+).  A step is a chain of 45-130 dependent kernels in one
// stream; between two of them the GPU otherwise drains, launches, fills.  Kernels launched through
// launch_k() carry cudaLaunchAttributeProgrammaticStreamSerialization: their CTAs are scheduled as
// soon as every CTA of the previous kernel has executed griddepcontrol.launch_dependents (or
// exited) -- i.e. into the SM slots the previous kernel's tail frees -- and park at
// griddepcontrol.wait, which returns once the previous grid has completed and its stores are
// visible.  ubgl_pdl_prologue() is the FIRST statement of every kernel launched that way (wait,
// then trigger: at most two kernels of the chain overlap).  A kernel without the prologue behind
// one with it, or the other way round, simply serialises as before.  UBGL_PDL=0 switches it off.
// Used for the latency-bound links of the chain: the multigrid passes (k_mg_run, k_mg_tail), the
// border kernels, the sink stamps and the slab halo pushes.  Measured: 8192^2 step 4.41 -> 4.35 ms
// (V-cycle 1.194 -> 1.163 ms), game level 0.2325 -> 0.2264 ms (V-cycle 0.104 -> 0.091 ms).  NOT used
// for the long streaming kernels (prestep, advect, divergence, gradient): with them in the chain the
// same step took 4.45 ms -- parked CTAs of the next kernel take the slots of an occupancy-bound tail.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void ubgl_pdl_prologue() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;");
}
inline bool pdl_enabled() {
  static const bool on = [] {
    const char *e = getenv("UBGL_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args &&...args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// Every kernel launch in the library goes through this macro.
#define UBGL_LAUNCH(lc, kind, level, stream, ...)                              \
  do {                                                                         \
    (lc)->pre((kind), (level), (stream));                                      \
    __VA_ARGS__;                                                               \
    UBGL_CHECK_LAUNCH();                                                       \
    (lc)->post((stream));                                                      \
  } while (0)

} // namespace ubgl
