// stencils.cuh -- per-cell arithmetic of every stage, written ONCE with explicit
// rounding intrinsics (no compiler-chosen FMA contraction) so that the plain
// one-kernel-per-stage path and the fused / temporally blocked kernels produce
// bit-identical fields.  Citations: file:line under te42kyfo/ubootgl.
#pragma once
#include <cuda_runtime.h>

namespace ubgl {

// 1/sum for the stencil denominators.  With binary flags sum is 0..4 and the
// quotient val/sum is evaluated as val * (1/sum): exact for 1, 2, 4 and within
// 1 ulp for 3 (the reference itself is built with -Ofast, where gcc vectorises
// the division as rcpps + one Newton step, so IEEE division is not what pins
// parity -- the 1e-5 relative-L2 tolerance is).  sum == 0 gives 0 like the
// reference's `if (sum == 0) val = 0` (pressure_solver.cpp:19).  Any other sum
// (non-binary flags, plain path only) divides.
__device__ __forceinline__ float rcp_count(int c) {
  return c == 4 ? 0.25f : c == 3 ? (1.0f / 3.0f) : c == 2 ? 0.5f : c == 1 ? 1.0f : 0.0f;
}
__device__ __forceinline__ float div_by_sum(float val, float sum) {
  if (sum == 4.0f) return __fmul_rn(val, 0.25f);
  if (sum == 3.0f) return __fmul_rn(val, 1.0f / 3.0f);
  if (sum == 2.0f) return __fmul_rn(val, 0.5f);
  if (sum == 1.0f) return val;
  if (sum == 0.0f) return 0.0f;
  return __fdiv_rn(val, sum);
}

// smoothingKernel (pressure_solver.cpp:10-24).
// fh2 = f(x,y)*h*h, evaluated by the caller as (f*h)*h.
__device__ __forceinline__ float smooth_cell(float pC, float pW, float pE,
                                             float pS, float pN, float fC,
                                             float fW, float fE, float fS,
                                             float fN, float fh2, float alpha) {
  float sum = __fadd_rn(__fadd_rn(__fadd_rn(fW, fE), fN), fS);
  float val = __fmul_rn(pW, fW);
  val = __fmaf_rn(pE, fE, val);
  val = __fmaf_rn(pS, fS, val);
  val = __fmaf_rn(pN, fN, val);
  val = __fadd_rn(val, fh2);
  val = div_by_sum(val, sum);
  float mix = __fmaf_rn(alpha, val, __fmul_rn(__fsub_rn(1.0f, alpha), pC));
  return __fmul_rn(fC, mix);
}

__device__ __forceinline__ float fh2_of(float f, float hh) {
  return __fmul_rn(__fmul_rn(f, hh), hh);
}

// calculateResidualField body (pressure_solver.cpp:101-111); ihsq = 1/h/h.
__device__ __forceinline__ float residual_cell(float pC, float pW, float pE,
                                               float pS, float pN, float fC,
                                               float fW, float fE, float fS,
                                               float fN, float f, float ihsq) {
  float val = __fmaf_rn(pW, fW, __fmul_rn(pC, __fsub_rn(1.0f, fW)));
  val = __fadd_rn(val, __fmaf_rn(pE, fE, __fmul_rn(pC, __fsub_rn(1.0f, fE))));
  val = __fadd_rn(val, __fmaf_rn(pS, fS, __fmul_rn(pC, __fsub_rn(1.0f, fS))));
  val = __fadd_rn(val, __fmaf_rn(pN, fN, __fmul_rn(pC, __fsub_rn(1.0f, fN))));
  val = __fmaf_rn(-4.0f, pC, val);
  val = __fmul_rn(val, ihsq);
  return __fmul_rn(__fadd_rn(f, val), fC);
}

// 9-point full weighting (restrict pressure_solver.cpp:122-129, updateFields
// pressure_solver.hpp:42-49); rows are y-1, y, y+1 of the fine grid.
__device__ __forceinline__ float fw9(float a0, float a1, float a2, float b0,
                                     float b1, float b2, float c0, float c1,
                                     float c2) {
  float v = __fadd_rn(__fmaf_rn(a1, 2.0f, a0), a2);
  v = __fadd_rn(v, __fadd_rn(__fmaf_rn(b1, 4.0f, __fmul_rn(b0, 2.0f)),
                             __fmul_rn(b2, 2.0f)));
  v = __fadd_rn(v, __fadd_rn(__fmaf_rn(c1, 2.0f, c0), c2));
  return __fmul_rn(v, 0.0625f);
}

// Denominators of prolongate (pressure_solver.cpp:149,157,164): flagc sums of
// 0..4 plus 0.0001 (a double literal in the reference).  Evaluated as a multiply
// by the rounded reciprocal, see rcp_count above.
__device__ __forceinline__ float prolong_rcp(int c) {
  return c == 4   ? (float)(1.0 / 4.0001)
         : c == 3 ? (float)(1.0 / 3.0001)
         : c == 2 ? (float)(1.0 / 2.0001)
         : c == 1 ? (float)(1.0 / 1.0001)
                  : (float)(1.0 / 0.0001);
}
__device__ __forceinline__ float prolong_div(float num, float fs) {
  if (fs == 4.0f || fs == 3.0f || fs == 2.0f || fs == 1.0f || fs == 0.0f)
    return __fmul_rn(num, prolong_rcp((int)fs));
  return __fdiv_rn(num, (float)((double)fs + 0.0001));
}

// prolongate (pressure_solver.cpp:134-172) evaluated per fine cell; returns the
// value the reference leaves in e(x,y) (0 where no loop writes it).  ec / flagc
// are the coarse error and coarse flag grids with pitch pc.
__device__ __forceinline__ float prolong_cell(const float *__restrict__ ec,
                                              const float *__restrict__ flagc,
                                              int pc, float flagf, int x, int y,
                                              int w, int h) {
  const int ox = x & 1, oy = y & 1;
  const int xc = x >> 1, yc = y >> 1;
  if (!ox && !oy) {
    if (x < 2 || y < 2 || x >= w - 1 || y >= h - 1) return 0.0f;
    return __fmul_rn(ec[(size_t)yc * pc + xc], flagf);
  }
  if (ox && !oy) {
    if (y < 2 || x >= w - 2 || y >= h - 1) return 0.0f;
    size_t i = (size_t)yc * pc + xc;
    return prolong_div(__fmul_rn(flagf, __fadd_rn(ec[i], ec[i + 1])),
                       __fadd_rn(flagc[i], flagc[i + 1]));
  }
  if (!ox && oy) {
    if (x < 2 || x >= w - 1 || y >= h - 2) return 0.0f;
    size_t i = (size_t)yc * pc + xc;
    return prolong_div(__fmul_rn(flagf, __fadd_rn(ec[i], ec[i + pc])),
                       __fadd_rn(flagc[i], flagc[i + pc]));
  }
  if (x >= w - 2 || y >= h - 2) return 0.0f;
  size_t i = (size_t)yc * pc + xc;
  // :163-164  flagc(x/2,y/2) + flagc(x/2+1,y/2+1) + flagc(x/2+1,y/2) + flagc(x/2,y/2+1)
  float fs = __fadd_rn(__fadd_rn(__fadd_rn(flagc[i], flagc[i + pc + 1]), flagc[i + 1]),
                       flagc[i + pc]);
  float es = __fadd_rn(__fadd_rn(__fadd_rn(ec[i], ec[i + pc + 1]), ec[i + 1]), ec[i + pc]);
  return prolong_div(__fmul_rn(flagf, es), fs);
}

// CubicHermite / Catmull-Rom (interpolators.hpp:78-85)
__device__ __forceinline__ float cubic_hermite(float t, float A, float B, float C,
                                               float D) {
  float a = __fadd_rn(__fsub_rn(__fadd_rn(__fmul_rn(-0.5f, A), __fmul_rn(1.5f, B)),
                                __fmul_rn(1.5f, C)),
                      __fmul_rn(0.5f, D));
  float b = __fsub_rn(__fadd_rn(__fsub_rn(A, __fmul_rn(2.5f, B)), __fmul_rn(2.0f, C)),
                      __fmul_rn(0.5f, D));
  float c = __fadd_rn(__fmul_rn(-0.5f, A), __fmul_rn(0.5f, C));
  // a*t*t*t + b*t*t + c*t + d, in Horner form with fused multiply-adds
  return __fmaf_rn(__fmaf_rn(__fmaf_rn(a, t, b), t, c), t, B);
}

// diffuse body for vx (simulation.cpp:117-127): c = vx.f(x,y), E/W/N/S its
// neighbours; fEE = flag(x+1,y)flag(x+2,y) etc. are evaluated by the caller.
__device__ __forceinline__ float diffuse_cell(float c, float vA, float mA, float vB,
                                              float mB, float vN, float fvn,
                                              float vS, float fvs, float mC,
                                              float a, float rden) {
  float val = __fmul_rn(vA, mA);
  val = __fmaf_rn(vB, mB, val);
  val = __fadd_rn(val, __fmaf_rn(vN, fvn, __fmul_rn(__fsub_rn(1.0f, fvn), -c)));
  val = __fadd_rn(val, __fmaf_rn(vS, fvs, __fmul_rn(__fsub_rn(1.0f, fvs), -c)));
  // (v + a*val) / (1 + 4a) as a multiply by the rounded reciprocal (see rcp_count)
  return __fmul_rn(__fmul_rn(mC, __fmaf_rn(a, val, c)), rden);
}

// VBCPar / VBCPer (simulation.cpp:50-78); bc numbering = Simulation::BC.
__device__ __forceinline__ float vbc_par(int bc, float a, float b) {
  if (bc == 0) return b;
  if (bc == 1 || bc == 2) return fmaxf(a, 0.0f);
  return 0.0f;
}
__device__ __forceinline__ float vbc_per(int bc, float a, float b) {
  if (bc == 0) return b;
  if (bc == 1 || bc == 2) return fmaxf(a, 0.0f);
  return -a;
}
// singlePBC (simulation.cpp:23-34)
__device__ __forceinline__ float single_pbc(int bc, float a) { return bc == 2 ? -a : a; }

} // namespace ubgl
