// capi_internal.cuh -- handle definitions and the try/catch shell shared by the
// translation units that implement include/ubgl.h (capi.cu, next.cu).
#pragma once
#include "../../include/ubgl.h"
#include "sim.cuh"
#include <memory>
#include <new>
#include <string>

struct ubgl_sim {
  std::unique_ptr<::ubgl::DeviceSim> s;
};


#define UBGL_TRY try {
#define UBGL_CATCH                                                             \
  }                                                                            \
  catch (const ::ubgl::CudaError &e) {                                                 \
    char buf[512];                                                             \
    snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d", (int)e.code,      \
             cudaGetErrorString(e.code), e.file, e.line);                      \
    ::ubgl::set_error(buf);                                                            \
    cudaGetLastError();                                                        \
    return e.code == cudaErrorMemoryAllocation ? UBGL_E_NOMEM : UBGL_E_CUDA;   \
  }                                                                            \
  catch (const ::ubgl::ArgError &e) {                                                  \
    ::ubgl::set_error(e.msg);                                                          \
    return UBGL_E_ARG;                                                         \
  }                                                                            \
  catch (const std::bad_alloc &) {                                             \
    ::ubgl::set_error("host allocation failed");                                       \
    return UBGL_E_NOMEM;                                                       \
  }                                                                            \
  catch (...) {                                                                \
    ::ubgl::set_error("unknown internal error");                                       \
    return UBGL_E_STATE;                                                       \
  }                                                                            \
  return UBGL_OK;

#define NEED(ptr, what) UBGL_REQUIRE((ptr) != nullptr, what " must not be null")

inline void require_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) throw ::ubgl::CudaError{e, __FILE__, __LINE__};
  UBGL_REQUIRE(device >= 0 && device < n, "no such CUDA device (libubgl has no CPU fallback)");
  UBGL_CUDA(cudaSetDevice(device));
}

#define SIM(sim)                                                               \
  NEED(sim, "sim");                                                            \
  ::ubgl::DeviceSim &S = *(sim)->s;                                                    \
  UBGL_CUDA(cudaSetDevice(S.device));
