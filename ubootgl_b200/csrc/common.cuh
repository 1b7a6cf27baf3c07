// common.cuh -- device-grid descriptor, error plumbing and launch helpers shared
// by every translation unit of libubgl.so (sm_100a only, no CPU fallback).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <string>

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

// Counts kernel launches per handle ("gpu_launches" in bench.py).
struct LaunchCounter {
  long long n = 0;
};

} // namespace ubgl
