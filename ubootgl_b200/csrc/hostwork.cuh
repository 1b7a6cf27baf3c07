// hostwork.cuh -- the host-thread side of the *_step_host entry points (single GPU: capi.cu,
// row slabs: slab.cu).
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace ubgl {

// Host side of ubgl_sim_step_host.  Two jobs run on a few host threads while the
// GPU steps and the DMA engines copy:
//  * applyAccumulatedVelocity's "accum = 0" (simulation.cpp:384,392): interior rows
//    1..H-2 x cols 1..W-3 of vx_accum ((W-1) x H) and rows 1..H-3 x cols 1..W-2 of
//    vy_accum (W x (H-1)), as soon as the upload has consumed the mirrors;
//  * saveCurrentVelocityFields (simulation.cpp:16-19) for the mirrors: vx_current /
//    vy_current are byte copies of the final front vx / vy, so they are filled from
//    the freshly downloaded vx / vy mirror band by band (a host memcpy behind the
//    D->H copy) instead of crossing PCIe a second time.
struct HostBand {
  cudaEvent_t ready = nullptr; // the D->H copy of this band has landed
  const float *src = nullptr;
  float *dst = nullptr;
  size_t bytes = 0;
};

// ax / ay are mirrors of the global rows [row0, ...) of the accumulators (row0 = 0: the whole
// grid; a slab rank: its first stored row); rows [clr_lo, clr_hi) of them are cleared.
inline void host_side_work(int device, cudaEvent_t uploaded, float *ax, float *ay, int W, int H, int row0,
                           int clr_lo, int clr_hi, std::vector<HostBand> &bands, cudaError_t *first_err) {
  const size_t big = (size_t)(16 << 20);
  size_t total = 0;
  for (auto &b : bands) total += b.bytes;
  if (uploaded) total += sizeof(float) * (size_t)W * (clr_hi - clr_lo) * ((ax ? 1 : 0) + (ay ? 1 : 0));
  unsigned nt = std::thread::hardware_concurrency();
  const unsigned cap = [] { // UBGL_HOST_THREADS: host copy threads of the *_step_host calls (read per call)
    const char *e = getenv("UBGL_HOST_THREADS");
    return e ? (unsigned)std::max(1, atoi(e)) : 16u;
  }();
  nt = std::min(nt ? nt : 1u, cap);
  if (total < big) nt = 1;
  std::vector<cudaError_t> errs(nt, cudaSuccess);
  auto work = [&](unsigned t) {
    cudaSetDevice(device);
    if (uploaded) {
      cudaError_t e = cudaEventSynchronize(uploaded);
      if (e != cudaSuccess) errs[t] = e;
      const int nr = clr_hi - clr_lo;
      const int y0 = clr_lo + (int)((long long)nr * t / nt), y1 = clr_lo + (int)((long long)nr * (t + 1) / nt);
      if (ax)
        for (int y = std::max(y0, 1); y < std::min(y1, H - 1); y++)
          std::memset(ax + (size_t)(y - row0) * (W - 1) + 1, 0, sizeof(float) * (W - 3));
      if (ay)
        for (int y = std::max(y0, 1); y < std::min(y1, H - 2); y++)
          std::memset(ay + (size_t)(y - row0) * W + 1, 0, sizeof(float) * (W - 2));
    }
    for (size_t b = t; b < bands.size(); b += nt) {
      cudaError_t e = cudaEventSynchronize(bands[b].ready);
      if (e != cudaSuccess) {
        errs[t] = e;
        continue;
      }
      std::memcpy(bands[b].dst, bands[b].src, bands[b].bytes);
    }
  };
  if (nt < 2) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++) th.emplace_back(work, t);
    for (auto &t : th) t.join();
  }
  for (auto e : errs)
    if (e != cudaSuccess && *first_err == cudaSuccess) *first_err = e;
}


} // namespace ubgl
