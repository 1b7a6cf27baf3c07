// packed.cuh -- two fp32 lanes per instruction (sm_100: fma / mul / add .rn.f32x2, SASS FFMA2 /
// FMUL2 / FADD2).  Each lane is an IEEE round-to-nearest operation, so a packed chain is bit for
// bit the scalar chain of stencils.cuh; the point is the instruction count of issue-bound kernels.
#pragma once
#include <cuda_runtime.h>

namespace ubgl {

struct f2 { // two fp32 in one 64-bit register pair
  unsigned long long v;
};
__device__ __forceinline__ f2 pk(float lo, float hi) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float lo(f2 a) {
  float x, y;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v));
  return x;
}
__device__ __forceinline__ float hi(f2 a) {
  float x, y;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a.v));
  return y;
}
__device__ __forceinline__ f2 bc(float c) { return pk(c, c); } // broadcast
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  f2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  f2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}

} // namespace ubgl
