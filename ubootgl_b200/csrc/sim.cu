// sim.cu -- Simulation::step on the device.  Reference: simulation.cpp /
// simulation.hpp / interpolators.hpp of te42kyfo/ubootgl (file:line per kernel).
#include "sim.cuh"
#include "stencils.cuh"
#include "packed.cuh"
#include <cstdlib>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <utility>

namespace ubgl {

#define LVL 0

static inline dim3 blk2d() { return dim3(32, 8); }
static inline dim3 grd2d(int w, int h) { return dim3(ceil_div(w, 32), ceil_div(h, 8)); }

// ---------------------------------------------------------------------------
// plain kernels: one per reference stage
// ---------------------------------------------------------------------------

// applyAccumulatedVelocity (simulation.cpp:376-396), one component
__global__ void k_apply_accum(Grid v, Grid acc) {
  int x = 1 + blockIdx.x * blockDim.x + threadIdx.x;
  int y = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= v.w - 1 || y >= v.h - 1) return;
  size_t i = (size_t)y * v.pitch + x;
  v.d[i] = __fadd_rn(v.d[i], acc.d[i]);
  acc.d[i] = 0.0f;
}

// diffuse, x-velocity pass (simulation.cpp:113-129)
__global__ void k_diffuse_vx(Grid src, Grid dst, Grid flag, float a, float rden) {
  int x = 1 + blockIdx.x * blockDim.x + threadIdx.x;
  int y = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= src.w - 1 || y >= src.h - 1) return;
  const float *v = src.d + (size_t)y * src.pitch + x;
  const float *fl = flag.d + (size_t)y * flag.pitch + x;
  const int vp = src.pitch, fp = flag.pitch;
  float mE = __fmul_rn(fl[1], fl[2]);
  float mW = __fmul_rn(fl[0], fl[-1]);
  float fvn = __fmul_rn(fl[fp], fl[fp - 1]);
  float fvs = __fmul_rn(fl[-fp], fl[-fp - 1]);
  float mC = __fmul_rn(fl[0], fl[1]);
  dst.d[(size_t)y * dst.pitch + x] =
      diffuse_cell(v[0], v[1], mE, v[-1], mW, v[vp], fvn, v[-vp], fvs, mC, a, rden);
}

// diffuse, y-velocity pass (simulation.cpp:139-155)
__global__ void k_diffuse_vy(Grid src, Grid dst, Grid flag, float a, float rden) {
  int x = 1 + blockIdx.x * blockDim.x + threadIdx.x;
  int y = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= src.w - 1 || y >= src.h - 1) return;
  const float *v = src.d + (size_t)y * src.pitch + x;
  const float *fl = flag.d + (size_t)y * flag.pitch + x;
  const int vp = src.pitch, fp = flag.pitch;
  float mS = __fmul_rn(fl[0], fl[-fp]);
  float mN = __fmul_rn(fl[fp], fl[2 * fp]);
  float fve = __fmul_rn(fl[1], fl[fp + 1]);
  float fvw = __fmul_rn(fl[-1], fl[fp - 1]);
  float mC = __fmul_rn(fl[0], fl[fp]);
  dst.d[(size_t)y * dst.pitch + x] =
      diffuse_cell(v[0], v[-vp], mS, v[vp], mN, v[1], fve, v[-1], fvw, mC, a, rden);
}

// setVBCs (simulation.cpp:80-102): column loops, then row loops (the row loops
// read the column results at x = 0 / w-1, and win at the corners), written to
// the front AND back buffers.  One block; __syncthreads separates the phases.
__global__ void k_set_vbcs(Grid xf, Grid xb, Grid yf, Grid yb, int bcW, int bcE, int bcN,
                           int bcS) {
  for (int y = threadIdx.x; y < xf.h; y += blockDim.x) {
    float v = vbc_par(bcW, xf.at(1, y), xf.at(0, y));
    xf.at(0, y) = v;
    xb.at(0, y) = v;
    v = vbc_par(bcE, xf.at(xf.w - 2, y), xf.at(xf.w - 1, y));
    xf.at(xf.w - 1, y) = v;
    xb.at(xf.w - 1, y) = v;
  }
  for (int y = threadIdx.x; y < yf.h; y += blockDim.x) {
    float v = vbc_per(bcW, yf.at(1, y), yf.at(0, y));
    yf.at(0, y) = v;
    yb.at(0, y) = v;
    v = vbc_per(bcE, yf.at(yf.w - 2, y), yf.at(yf.w - 1, y));
    yf.at(yf.w - 1, y) = v;
    yb.at(yf.w - 1, y) = v;
  }
  __syncthreads();
  for (int x = threadIdx.x; x < xf.w; x += blockDim.x) {
    float v = vbc_per(bcS, xf.at(x, 1), xf.at(x, 0));
    xf.at(x, 0) = v;
    xb.at(x, 0) = v;
    v = vbc_per(bcN, xf.at(x, xf.h - 2), xf.at(x, xf.h - 1));
    xf.at(x, xf.h - 1) = v;
    xb.at(x, xf.h - 1) = v;
  }
  for (int x = threadIdx.x; x < yf.w; x += blockDim.x) {
    float v = vbc_par(bcS, yf.at(x, 1), yf.at(x, 0));
    yf.at(x, 0) = v;
    yb.at(x, 0) = v;
    v = vbc_par(bcN, yf.at(x, yf.h - 2), yf.at(x, yf.h - 1));
    yf.at(x, yf.h - 1) = v;
    yb.at(x, yf.h - 1) = v;
  }
}

// Catmull-Rom weights of the four taps: CubicHermite (interpolators.hpp:78-85)
// a*t^3 + b*t^2 + c*t + d regrouped by tap, i.e. the same cubic evaluated as a
// weighted sum (differs from the reference's Horner form only in rounding,
// ~1e-7 relative; the kernel is issue bound and this form needs 40 instead of
// 95 floating-point instructions per bicubic sample).
__device__ __forceinline__ void cr_weights(float t, float &w0, float &w1, float &w2, float &w3) {
  const float t2 = __fmul_rn(t, t);
  w0 = __fmul_rn(__fmaf_rn(__fmaf_rn(-0.5f, t, 1.0f), t, -0.5f), t);
  w1 = __fmaf_rn(__fmaf_rn(1.5f, t, -2.5f), t2, 1.0f);
  w2 = __fmul_rn(__fmaf_rn(__fmaf_rn(-1.5f, t, 2.0f), t, 0.5f), t);
  w3 = __fmul_rn(__fmaf_rn(0.5f, t, -0.5f), t2);
}

// bicubicSample (interpolators.hpp:92-206): clamp to [3, w-3] x [3, h-3]
// (:94-97), truncate (:99-103), 4x4 taps at rows iy-1..iy+2 / cols ix-1..ix+2,
// vertical Hermite per column first, then horizontal (:132-204).
// Row-slab runs (csrc/slab.cu) store only rows [lo, hi) of a field.  A tap row
// outside them is read straight from the NEIGHBOUR's slab through its
// NVLink-mapped arena (peer loads; NVSwitch makes every peer uniform, SURVEY.md
// 8e), so the halo depth bounds nothing: any CFL works, the rare far tap just
// costs an NVLink round trip.  The neighbour's front buffers are complete when
// this kernel starts (its halo push is stream-ordered after them and was waited
// for) and are not rewritten before the next exchange.  Only a tap beyond the
// neighbour's own rows (back-trace longer than a whole slab) raises *err.
struct TapRows {
  int lo = 0, hi = 0;             // rows stored locally
  int plo = 0, phi = 0;           // rows reachable through the neighbours [plo, phi)
  const float *x_lo = nullptr, *x_hi = nullptr; // vx front of the lower / upper neighbour (virtual row 0)
  const float *y_lo = nullptr, *y_hi = nullptr; // vy front likewise
  int *err = nullptr;
  // k_advect_xy's one-compare form of "all four tap rows are stored locally":
  // (unsigned)(icy - lo1) <= span, lo1 = lo + 1, span = min(hi, grid height) - 3 - lo1
  int lo1 = 0;
  unsigned span_x = 0, span_y = 0;
};

// The rare slab case -- a tap row that is not stored locally -- is kept out of line so that
// the common path of the slab kernels stays as lean (registers, instruction count) as the
// single-GPU one.  Same arithmetic, rows fetched from the neighbour's arena.
__device__ __noinline__ float bicubic_far(const float *g, int pitch, int icx, int icy, float stx,
                                          float sty, int h, int lo, int hi_, int plo, int phi,
                                          const float *g_lo, const float *g_hi, int *err) {
  if (icy - 1 < plo || icy + 2 >= min(phi, h)) {
    *err = 1;
    icy = max(plo + 1, min(icy, min(phi, h) - 3));
  }
  const int hi = min(hi_, h);
  float y0, y1, y2, y3, x0, x1, x2, x3;
  cr_weights(sty, y0, y1, y2, y3);
  cr_weights(stx, x0, x1, x2, x3);
  auto rowp = [&](int y) {
    const float *b = y < lo ? g_lo : (y >= hi ? g_hi : g);
    return b + ((size_t)y * pitch + (icx - 1));
  };
  const float *r0 = rowp(icy - 1), *r1 = rowp(icy), *r2 = rowp(icy + 1), *r3 = rowp(icy + 2);
  auto col = [&](int j) {
    float c = __fmul_rn(y0, __ldg(r0 + j));
    c = __fmaf_rn(y1, __ldg(r1 + j), c);
    c = __fmaf_rn(y2, __ldg(r2 + j), c);
    return __fmaf_rn(y3, __ldg(r3 + j), c);
  };
  float v = __fmul_rn(x0, col(0));
  v = __fmaf_rn(x1, col(1), v);
  v = __fmaf_rn(x2, col(2), v);
  return __fmaf_rn(x3, col(3), v);
}

template <bool SLAB>
__device__ __forceinline__ float bicubic(const float *__restrict__ g, int pitch, int w, int h,
                                         float cx, float cy, const TapRows &tr,
                                         const float *g_lo, const float *g_hi) {
  cx = fmaxf(fminf(cx, (float)w - 3.0f), 3.0f);
  cy = fmaxf(fminf(cy, (float)h - 3.0f), 3.0f);
  const int icx = (int)cx;
  const int icy = (int)cy;
  const float stx = __fsub_rn(cx, (float)icx), sty = __fsub_rn(cy, truncf(cy));
  if (SLAB && (icy - 1 < tr.lo || icy + 2 >= min(tr.hi, h)))
    return bicubic_far(g, pitch, icx, icy, stx, sty, h, tr.lo, tr.hi, tr.plo, tr.phi, g_lo, g_hi, tr.err);
  float y0, y1, y2, y3, x0, x1, x2, x3;
  cr_weights(sty, y0, y1, y2, y3);
  cr_weights(stx, x0, x1, x2, x3);
  const float *r0 = g + ((icy - 1) * pitch + (icx - 1));
  const float *r1 = r0 + pitch, *r2 = r1 + pitch, *r3 = r2 + pitch;
  auto col = [&](int j) {
    float c = __fmul_rn(y0, __ldg(r0 + j));
    c = __fmaf_rn(y1, __ldg(r1 + j), c);
    c = __fmaf_rn(y2, __ldg(r2 + j), c);
    return __fmaf_rn(y3, __ldg(r3 + j), c);
  };
  float v = __fmul_rn(x0, col(0));
  v = __fmaf_rn(x1, col(1), v);
  v = __fmaf_rn(x2, col(2), v);
  return __fmaf_rn(x3, col(3), v);
}

// advect, x-velocity faces (simulation.cpp:246-297).  blockDim.x == 32 and the
// warp's first face is x = 1 (mod 32), so lanes 8k..8k+7 are exactly one of the
// reference's AVX2 octets (x = 1, 9, 17, ...).  Reproduced quirks:
//  * octets exist only while x0 < vx.width - 8 (:248) -> last columns untouched;
//  * whole-octet skip unless some lane has flag(x-1+i,y)+flag(x+i,y) == 2 (:254);
//  * untouched entries keep whatever the back buffer holds.
template <bool SLAB>
__global__ void __launch_bounds__(256, 8) k_advect_vx(Grid vx, Grid vy, Grid vxb, Grid flag, float half, float full,
                            int y_lo, int y_hi, TapRows tr) {
  const int lane = threadIdx.x;
  const int xi = 1 + blockIdx.x * 32 + lane;
  const int y = y_lo + blockIdx.y * blockDim.y + threadIdx.y; // y_lo >= 1, y_hi <= H-1
  const bool row_ok = y < y_hi;
  const int x0 = xi - (lane & 7);
  const bool oct_ok = row_ok && (x0 < vx.w - 8);
  bool cond = false;
  float f0 = 0.0f, f1 = 0.0f;
  if (oct_ok) {
    const float *fl = flag.d + (size_t)y * flag.pitch + xi;
    f0 = fl[0];
    f1 = fl[1];
    cond = (__fadd_rn(fl[-1], f0) == 2.0f);
  }
  unsigned ball = __ballot_sync(0xffffffffu, cond);
  if (!oct_ok || ((ball >> (lane & ~7)) & 0xffu) == 0) return;

  float posx = __fadd_rn((float)xi, 0.5f), posy = (float)y;
  float vx1 = vx.at(xi, y);
  float vy1 = __fmul_rn(__fadd_rn(__fadd_rn(vy.at(xi, y), vy.at(xi, y - 1)),
                                  __fadd_rn(vy.at(xi + 1, y), vy.at(xi + 1, y - 1))),
                        0.25f);
  float midx = __fmaf_rn(-vx1, half, posx), midy = __fmaf_rn(-vy1, half, posy);
  float vx2 = bicubic<SLAB>(vx.d, vx.pitch, vx.w, vx.h, __fsub_rn(midx, 0.5f), midy, tr, tr.x_lo, tr.x_hi);
  float vy2 = bicubic<SLAB>(vy.d, vy.pitch, vy.w, vy.h, midx, __fsub_rn(midy, 0.5f), tr, tr.y_lo, tr.y_hi);
  float endx = __fmaf_rn(-vx2, full, posx), endy = __fmaf_rn(-vy2, full, posy);
  float xvel = bicubic<SLAB>(vx.d, vx.pitch, vx.w, vx.h, __fsub_rn(endx, 0.5f), endy, tr, tr.x_lo, tr.x_hi);
  vxb.at(xi, y) = __fmul_rn(__fmul_rn(xvel, f0), f1);
}

// advect, y-velocity faces (simulation.cpp:300-347); rows 1..H-2 like the vx
// loop (so vy's top boundary row IS written, :247).  The four vx loads at
// (x+1, .) are FLAT in the reference (:317-325): when x+1 == vx.width they read
// the first element of the next row; reproduced through vx_flat().
__device__ __forceinline__ float vx_flat(const Grid &vx, int x, int y) {
  if (x >= vx.w) {
    x -= vx.w;
    y += 1;
  }
  return vx.at(x, y);
}
template <bool SLAB>
__global__ void __launch_bounds__(256, 8) k_advect_vy(Grid vx, Grid vy, Grid vyb, Grid flag, float half, float full,
                            int y_lo, int y_hi, TapRows tr) {
  const int lane = threadIdx.x;
  const int xi = 1 + blockIdx.x * 32 + lane;
  const int y = y_lo + blockIdx.y * blockDim.y + threadIdx.y;
  const bool row_ok = y < y_hi; // y in [1, H-2] like the vx loop (simulation.cpp:247)
  const int x0 = xi - (lane & 7);
  const bool oct_ok = row_ok && (x0 < vy.w - 8);
  bool cond = false;
  float f0 = 0.0f, f1 = 0.0f;
  if (oct_ok) {
    const float *fl = flag.d + (size_t)y * flag.pitch + xi;
    f0 = fl[0];
    f1 = fl[flag.pitch];
    cond = (__fadd_rn(f0, fl[-flag.pitch]) == 2.0f);
  }
  unsigned ball = __ballot_sync(0xffffffffu, cond);
  if (!oct_ok || ((ball >> (lane & ~7)) & 0xffu) == 0) return;

  float posy = __fadd_rn((float)y, 0.5f), posx = (float)xi;
  float vy1 = vy.at(xi, y);
  float vx1 = __fmul_rn(__fadd_rn(__fadd_rn(vx.at(xi, y), vx.at(xi, y - 1)),
                                  __fadd_rn(vx_flat(vx, xi + 1, y), vx_flat(vx, xi + 1, y - 1))),
                        0.25f);
  float midx = __fmaf_rn(-vx1, half, posx), midy = __fmaf_rn(-vy1, half, posy);
  float vx2 = bicubic<SLAB>(vx.d, vx.pitch, vx.w, vx.h, __fsub_rn(midx, 0.5f), midy, tr, tr.x_lo, tr.x_hi);
  float vy2 = bicubic<SLAB>(vy.d, vy.pitch, vy.w, vy.h, midx, __fsub_rn(midy, 0.5f), tr, tr.y_lo, tr.y_hi);
  float endx = __fmaf_rn(-vx2, full, posx), endy = __fmaf_rn(-vy2, full, posy);
  float yvel = bicubic<SLAB>(vy.d, vy.pitch, vy.w, vy.h, endx, __fsub_rn(endy, 0.5f), tr, tr.y_lo, tr.y_hi);
  vyb.at(xi, y) = __fmul_rn(__fmul_rn(yvel, f0), f1);
}

// ---------------------------------------------------------------------------
// Row-pair advect (EXPERIMENT, off by default: UBGL_ADVECT_VARIANT=2).  Question: are
// k_advect_vx / vy above bound by the L1 wavefront rate (56 loads per face, each touching
// two 128-byte lines) or by instruction issue?  Here a thread advects the faces (x, y) and
// (x, y+1): their back-traced 4x4 tap windows are, almost everywhere, the same columns and
// rows shifted by one, so the pair needs 5 tap rows instead of 2 x 4 (and 6 instead of 8
// loads for the RK2 start velocity): 38 loads per face.  Where the pair's windows do not
// line up (strong shear, clamping at the border, a far tap of a slab run) the two samples
// are taken separately with bicubic<SLAB>().  Per-face arithmetic is unchanged, the results
// are bit-identical to the one-face kernels (tests/test_gpu_advect_variants.py).
// Answer (B200, 8192^2): 0.789 vs 0.776 ms per launch -- a third fewer loads buys nothing,
// the kernel is bound by instruction issue (154 FP + ~150 address / conversion / control
// instructions per face), so the simpler one-face kernels stay the default.
// ---------------------------------------------------------------------------
template <bool SLAB>
__device__ __forceinline__ void bicubic2(const float *__restrict__ g, int pitch, int w, int h,
                                         float cxA, float cyA, float cxB, float cyB, const TapRows &tr,
                                         const float *g_lo, const float *g_hi, float &outA, float &outB) {
  const float ax = fmaxf(fminf(cxA, (float)w - 3.0f), 3.0f), ay = fmaxf(fminf(cyA, (float)h - 3.0f), 3.0f);
  const float bx = fmaxf(fminf(cxB, (float)w - 3.0f), 3.0f), by = fmaxf(fminf(cyB, (float)h - 3.0f), 3.0f);
  const int iax = (int)ax, iay = (int)ay, ibx = (int)bx, iby = (int)by;
  bool together = iax == ibx && iby == iay + 1;
  if (SLAB) together = together && !(iay - 1 < tr.lo || iby + 2 >= min(tr.hi, h));
  if (!together) {
    outA = bicubic<SLAB>(g, pitch, w, h, cxA, cyA, tr, g_lo, g_hi);
    outB = bicubic<SLAB>(g, pitch, w, h, cxB, cyB, tr, g_lo, g_hi);
    return;
  }
  float a0, a1, a2, a3, b0, b1, b2, b3, xa0, xa1, xa2, xa3, xb0, xb1, xb2, xb3;
  cr_weights(__fsub_rn(ay, truncf(ay)), a0, a1, a2, a3);
  cr_weights(__fsub_rn(ax, (float)iax), xa0, xa1, xa2, xa3);
  cr_weights(__fsub_rn(by, truncf(by)), b0, b1, b2, b3);
  cr_weights(__fsub_rn(bx, (float)ibx), xb0, xb1, xb2, xb3);
  const float *r0 = g + ((iay - 1) * pitch + (iax - 1));
  const float *r1 = r0 + pitch, *r2 = r1 + pitch, *r3 = r2 + pitch, *r4 = r3 + pitch;
  float ca[4], cb[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const float t0 = __ldg(r0 + j), t1 = __ldg(r1 + j), t2 = __ldg(r2 + j), t3 = __ldg(r3 + j), t4 = __ldg(r4 + j);
    float c = __fmul_rn(a0, t0);
    c = __fmaf_rn(a1, t1, c);
    c = __fmaf_rn(a2, t2, c);
    ca[j] = __fmaf_rn(a3, t3, c);
    c = __fmul_rn(b0, t1);
    c = __fmaf_rn(b1, t2, c);
    c = __fmaf_rn(b2, t3, c);
    cb[j] = __fmaf_rn(b3, t4, c);
  }
  float v = __fmul_rn(xa0, ca[0]);
  v = __fmaf_rn(xa1, ca[1], v);
  v = __fmaf_rn(xa2, ca[2], v);
  outA = __fmaf_rn(xa3, ca[3], v);
  v = __fmul_rn(xb0, cb[0]);
  v = __fmaf_rn(xb1, cb[1], v);
  v = __fmaf_rn(xb2, cb[2], v);
  outB = __fmaf_rn(xb3, cb[3], v);
}

// COMP 0: x-velocity faces (simulation.cpp:246-297), COMP 1: y-velocity faces (:300-347).
// blockDim = (32, 8): a block covers 32 faces x 16 rows, thread (lane, ty) the rows
// y and y+1 with y = y_lo + 16 blockIdx.y + 2 ty.  Octet logic per row as in the
// one-face kernels.
template <bool SLAB, int COMP>
__global__ void __launch_bounds__(256, 4) k_advect_pair(Grid vx, Grid vy, Grid out, Grid flag, float half, float full,
                                                        int y_lo, int y_hi, TapRows tr) {
  const int lane = threadIdx.x;
  const int xi = 1 + blockIdx.x * 32 + lane;
  const int yA = y_lo + (blockIdx.y * blockDim.y + threadIdx.y) * 2, yB = yA + 1;
  const Grid &self = COMP == 0 ? vx : vy;
  const int x0 = xi - (lane & 7);
  const bool octA = yA < y_hi && (x0 < self.w - 8), octB = yB < y_hi && (x0 < self.w - 8);
  bool condA = false, condB = false;
  float fA0 = 0.f, fA1 = 0.f, fB0 = 0.f, fB1 = 0.f;
  const int fp = flag.pitch;
  if (octA) {
    const float *fl = flag.d + (size_t)yA * fp + xi;
    fA0 = fl[0];
    if (COMP == 0) {
      fA1 = fl[1];
      condA = (__fadd_rn(fl[-1], fA0) == 2.0f);
    } else {
      fA1 = fl[fp];
      condA = (__fadd_rn(fA0, fl[-fp]) == 2.0f);
      if (octB) { // row B's flags: (x, y+1) = fA1, (x, y+2), and its test pairs (x, y+1) with (x, y)
        fB0 = fA1;
        fB1 = fl[2 * fp];
        condB = (__fadd_rn(fB0, fA0) == 2.0f);
      }
    }
  }
  if (COMP == 0 && octB) {
    const float *fl = flag.d + (size_t)yB * fp + xi;
    fB0 = fl[0];
    fB1 = fl[1];
    condB = (__fadd_rn(fl[-1], fB0) == 2.0f);
  }
  const unsigned ballA = __ballot_sync(0xffffffffu, condA), ballB = __ballot_sync(0xffffffffu, condB);
  const bool actA = octA && ((ballA >> (lane & ~7)) & 0xffu) != 0;
  const bool actB = octB && ((ballB >> (lane & ~7)) & 0xffu) != 0;
  if (!actA && !actB) return;
  // an inactive face of the pair is computed on the coordinates of the active one (never stored)
  const int ya = actA ? yA : yB, yb = actB ? yB : yA;

  float posxA, posyA, posxB, posyB, vx1A, vy1A, vx1B, vy1B;
  if (COMP == 0) {
    posxA = posxB = __fadd_rn((float)xi, 0.5f);
    posyA = (float)ya;
    posyB = (float)yb;
    vx1A = vx.at(xi, ya);
    vx1B = vx.at(xi, yb);
    // vy rows ya-1, ya (and yb = ya+1 when the pair is whole) at x and x+1
    const float p0 = vy.at(xi, ya - 1), q0 = vy.at(xi + 1, ya - 1), p1 = vy.at(xi, ya), q1 = vy.at(xi + 1, ya);
    vy1A = __fmul_rn(__fadd_rn(__fadd_rn(p1, p0), __fadd_rn(q1, q0)), 0.25f);
    if (yb == ya + 1) {
      const float p2 = vy.at(xi, yb), q2 = vy.at(xi + 1, yb);
      vy1B = __fmul_rn(__fadd_rn(__fadd_rn(p2, p1), __fadd_rn(q2, q1)), 0.25f);
    } else {
      vy1B = vy1A;
    }
  } else {
    posxA = posxB = (float)xi;
    posyA = __fadd_rn((float)ya, 0.5f);
    posyB = __fadd_rn((float)yb, 0.5f);
    vy1A = vy.at(xi, ya);
    vy1B = vy.at(xi, yb);
    const float p0 = vx.at(xi, ya - 1), q0 = vx_flat(vx, xi + 1, ya - 1), p1 = vx.at(xi, ya), q1 = vx_flat(vx, xi + 1, ya);
    vx1A = __fmul_rn(__fadd_rn(__fadd_rn(p1, p0), __fadd_rn(q1, q0)), 0.25f);
    if (yb == ya + 1) {
      const float p2 = vx.at(xi, yb), q2 = vx_flat(vx, xi + 1, yb);
      vx1B = __fmul_rn(__fadd_rn(__fadd_rn(p2, p1), __fadd_rn(q2, q1)), 0.25f);
    } else {
      vx1B = vx1A;
    }
  }
  const float midxA = __fmaf_rn(-vx1A, half, posxA), midyA = __fmaf_rn(-vy1A, half, posyA);
  const float midxB = __fmaf_rn(-vx1B, half, posxB), midyB = __fmaf_rn(-vy1B, half, posyB);
  float vx2A, vx2B, vy2A, vy2B;
  bicubic2<SLAB>(vx.d, vx.pitch, vx.w, vx.h, __fsub_rn(midxA, 0.5f), midyA, __fsub_rn(midxB, 0.5f), midyB, tr,
                 tr.x_lo, tr.x_hi, vx2A, vx2B);
  bicubic2<SLAB>(vy.d, vy.pitch, vy.w, vy.h, midxA, __fsub_rn(midyA, 0.5f), midxB, __fsub_rn(midyB, 0.5f), tr,
                 tr.y_lo, tr.y_hi, vy2A, vy2B);
  const float endxA = __fmaf_rn(-vx2A, full, posxA), endyA = __fmaf_rn(-vy2A, full, posyA);
  const float endxB = __fmaf_rn(-vx2B, full, posxB), endyB = __fmaf_rn(-vy2B, full, posyB);
  float rA, rB;
  if (COMP == 0)
    bicubic2<SLAB>(vx.d, vx.pitch, vx.w, vx.h, __fsub_rn(endxA, 0.5f), endyA, __fsub_rn(endxB, 0.5f), endyB, tr,
                   tr.x_lo, tr.x_hi, rA, rB);
  else
    bicubic2<SLAB>(vy.d, vy.pitch, vy.w, vy.h, endxA, __fsub_rn(endyA, 0.5f), endxB, __fsub_rn(endyB, 0.5f), tr,
                   tr.y_lo, tr.y_hi, rA, rB);
  if (actA) out.at(xi, yA) = __fmul_rn(__fmul_rn(rA, fA0), fA1);
  if (actB) out.at(xi, yB) = __fmul_rn(__fmul_rn(rB, fB0), fB1);
}

// ---------------------------------------------------------------------------
// k_advect_xy -- the advect of the fused step: BOTH components in one launch, packed fp32.
//
// k_advect_vx / k_advect_vy above are bound by instruction issue (ncu, 8192^2: issue slots
// 77 % busy, 360 SASS instructions per face for 48 taps, DRAM 15 %).  This kernel does the
// same arithmetic, bit for bit, in about half the instructions:
//  * fma.rn.f32x2 / mul.rn.f32x2 (SASS FFMA2 / FMUL2, new on sm_100): the four vertical
//    Catmull-Rom column sums of a sample are two packed chains (8 instead of 16 FP
//    instructions), the tap weights two packed polynomials (6 instead of 10 per coordinate);
//  * truncation by a round-toward-zero add of 2^23 (FADD.RZ on the FMA pipe) instead of
//    F2I + I2F + FRND on the quarter-rate conversion pipe: the integer part is the low
//    mantissa of the sum, its float value one more add;
//  * one thread advects the vx face AND the vy face of its cell: the RK2 start velocities
//    share loads (8 instead of 10), the five flags the two octet tests and masks need come
//    from ONE byte of the level-0 stencil mask (C, W, E, S, N bits), and two independent
//    dependency chains are in flight per thread;
//  * the accumulator interiors are zeroed here (simulation.cpp:384,392) with two stores
//    that cost this issue-bound kernel nothing in DRAM terms, instead of 8 B/cell of the
//    DRAM-bound divergence pass.
// Semantics (octets x = 1, 9, 17 ..., whole-octet skip, untouched last columns, flat
// (x+1) loads of the vy loop, rows 1..H-2 for both components) as k_advect_vx / vy;
// binary flags only (the fused step's precondition).
// ---------------------------------------------------------------------------
// cr_weights() of tx (low halves) and ty (high halves) at once: the same operations lane for
// lane, constants as broadcast scalars (w3 = fma(.., t2, +0) instead of a multiply: equal up to
// the sign of a zero weight)
__device__ __forceinline__ void cr_weights_xy(float tx, float ty, f2 &w0, f2 &w1, f2 &w2, f2 &w3) {
  const f2 t = pk(tx, ty);
  const f2 t2 = mul2(t, t);
  w0 = mul2(fma2(fma2(bc(-0.5f), t, bc(1.0f)), t, bc(-0.5f)), t);
  w1 = fma2(fma2(bc(1.5f), t, bc(-2.5f)), t2, bc(1.0f));
  w2 = mul2(fma2(fma2(bc(-1.5f), t, bc(2.0f)), t, bc(0.5f)), t);
  w3 = fma2(fma2(bc(0.5f), t, bc(-0.5f)), t2, bc(0.0f));
}

// floor of a clamped (>= 3) coordinate: c + 2^23 rounded toward zero is 2^23 + floor(c)
// exactly; returns the integer part, *fl its float value
__device__ __forceinline__ int floor_rz(float c, float &fl) {
  const float m = __fadd_rz(c, 8388608.0f);
  fl = __fsub_rn(m, 8388608.0f);
  return __float_as_int(m) - 0x4B000000;
}

// wm3 / hm3: (float)w - 3, (float)h - 3 of the sampled grid (kernel parameters: the clamps read
// them from the constant bank instead of converting w and h for every sample)
template <bool SLAB>
__device__ __forceinline__ float bicubic_pk(const float *__restrict__ g, int pitch, float wm3, float hm3, int h,
                                            float cx, float cy, const TapRows &tr, unsigned span,
                                            const float *g_lo, const float *g_hi) {
  cx = fmaxf(fminf(cx, wm3), 3.0f);
  cy = fmaxf(fminf(cy, hm3), 3.0f);
  float flx, fly;
  const int icx = floor_rz(cx, flx), icy = floor_rz(cy, fly);
  const float stx = __fsub_rn(cx, flx), sty = __fsub_rn(cy, fly);
  if (SLAB && (unsigned)(icy - tr.lo1) > span) // icy - 1 < tr.lo || icy + 2 >= min(tr.hi, h)
    return bicubic_far(g, pitch, icx, icy, stx, sty, h, tr.lo, tr.hi, tr.plo, tr.phi, g_lo, g_hi, tr.err);
  f2 w0, w1, w2, w3; // (x weight, y weight) of tap 0..3
  cr_weights_xy(stx, sty, w0, w1, w2, w3);
  const float *r0 = g + ((icy - 1) * pitch + (icx - 1));
  const float *r1 = r0 + pitch, *r2 = r1 + pitch, *r3 = r2 + pitch;
  // columns (0, 2) and (1, 3): c_j = ((y0 t0j + y1 t1j) + y2 t2j) + y3 t3j, as bicubic()
  f2 c02 = mul2(bc(hi(w0)), pk(__ldg(r0), __ldg(r0 + 2)));
  f2 c13 = mul2(bc(hi(w0)), pk(__ldg(r0 + 1), __ldg(r0 + 3)));
  c02 = fma2(bc(hi(w1)), pk(__ldg(r1), __ldg(r1 + 2)), c02);
  c13 = fma2(bc(hi(w1)), pk(__ldg(r1 + 1), __ldg(r1 + 3)), c13);
  c02 = fma2(bc(hi(w2)), pk(__ldg(r2), __ldg(r2 + 2)), c02);
  c13 = fma2(bc(hi(w2)), pk(__ldg(r2 + 1), __ldg(r2 + 3)), c13);
  c02 = fma2(bc(hi(w3)), pk(__ldg(r3), __ldg(r3 + 2)), c02);
  c13 = fma2(bc(hi(w3)), pk(__ldg(r3 + 1), __ldg(r3 + 3)), c13);
  float v = __fmul_rn(lo(w0), lo(c02));
  v = __fmaf_rn(lo(w1), lo(c13), v);
  v = __fmaf_rn(lo(w2), hi(c02), v);
  return __fmaf_rn(lo(w3), hi(c13), v);
}

enum { AMB_C = 1, AMB_W = 2, AMB_E = 32, AMB_S = 64, AMB_N = 128 }; // stencil mask bits (mg_fused.cu)

// DIV: the divergence of the advected field (simulation.cpp:166-171) in the epilogue, for the cells
// whose four faces this CTA holds: f(x, y) needs vx(x-1, y) -- the left lane's face, by shuffle -- and
// vy(x, y-1) -- the face of the thread one row down, through 1 KB of shared memory.  Faces this
// launch does not advect (skipped octets, last columns) are read back from the buffer they stay in.
// The CTA's first row and first column, and the ring of cells next to the border faces setVBCs
// rewrites afterwards, are left to k_divergence_edges (after setVBCs).  Same operation order as
// k_divergence: bit-identical f.
template <bool SLAB, int OCC, bool DIV>
__global__ void __launch_bounds__(256, OCC) k_advect_xy(Grid vx, Grid vy, Grid vxb, Grid vyb,
                                                      const uint8_t *__restrict__ mask, float *ax, float *ay,
                                                      float half, float full, float4 lim /* vx.w-3, vx.h-3, vy.w-3, vy.h-3 */,
                                                      int y_lo, int y_hi, TapRows tr, float *fdiv, float nih) {
  const int lane = threadIdx.x;
  const int xi = 1 + blockIdx.x * 32 + lane;
  const int y = y_lo + blockIdx.y * blockDim.y + threadIdx.y; // y_lo >= 1, y_hi <= H-1
  const int W = vy.w, H = vx.h;
  const bool row_ok = y < y_hi;
  const int x0 = xi - (lane & 7);
  const bool octx = row_ok && (x0 < vx.w - 8), octy = row_ok && (x0 < vy.w - 8);
  const int pitch = vx.pitch; // shared by every level-0 grid
  const int o = y * pitch + xi;
  // The stencil-mask byte and the eight RK2 start velocities -- vx / vy at (x, y), (x, y-1),
  // (x+1, y), (x+1, y-1); the (x+1) loads of vx are flat (simulation.cpp:317-325: x+1 == vx.width
  // wraps into the next row) -- are requested TOGETHER, before the octet test: the kernel is bound
  // by the latency of its first-touch loads (ncu: long scoreboard on exactly these two groups),
  // one round trip instead of two.  Faces in skipped octets load eight values for nothing.
  unsigned m = 0;
  float ux00 = 0.f, uy00 = 0.f, ux0m = 0.f, uy0m = 0.f, uy10 = 0.f, uy1m = 0.f, ux10 = 0.f, ux1m = 0.f;
  if (row_ok && xi <= W - 2) {
    m = mask[o];
    const float *px = vx.d + o, *py = vy.d + o;
    ux00 = px[0]; uy00 = py[0];
    ux0m = px[-pitch]; uy0m = py[-pitch];
    uy10 = py[1]; uy1m = py[1 - pitch];
    if (xi + 1 >= vx.w) {
      ux10 = vx.d[(y + 1) * pitch + (xi + 1 - vx.w)];
      ux1m = vx.d[y * pitch + (xi + 1 - vx.w)];
    } else {
      ux10 = px[1];
      ux1m = px[1 - pitch];
    }
    // accumulators: interiors 1..W-3 x 1..H-2 (vx) and 1..W-2 x 1..H-3 (vy), simulation.cpp:380-394
    if (ax) {
      if (xi <= W - 3) ax[o] = 0.0f;
      if (y <= H - 3) ay[o] = 0.0f;
    }
  }
  const bool condx = octx && (m & (AMB_C | AMB_W)) == (AMB_C | AMB_W); // flag(x-1,y)+flag(x,y) == 2
  const bool condy = octy && (m & (AMB_C | AMB_S)) == (AMB_C | AMB_S); // flag(x,y)+flag(x,y-1) == 2
  const unsigned bx = __ballot_sync(0xffffffffu, condx), by = __ballot_sync(0xffffffffu, condy);
  const bool actx = octx && ((bx >> (lane & ~7)) & 0xffu) != 0;
  const bool acty = octy && ((by >> (lane & ~7)) & 0xffu) != 0;
  if (!DIV && !actx && !acty) return;

  const float fC = (m & AMB_C) ? 1.0f : 0.0f, fE = (m & AMB_E) ? 1.0f : 0.0f, fN = (m & AMB_N) ? 1.0f : 0.0f;
  const float fx = (float)xi, fy = (float)y;
  float nvx = 0.0f, nvy = 0.0f; // DIV: the faces (xi, y) of the advected field

  if (actx) { // simulation.cpp:246-297
    const float posx = __fadd_rn(fx, 0.5f), posy = fy;
    const float vx1 = ux00;
    const float vy1 = __fmul_rn(__fadd_rn(__fadd_rn(uy00, uy0m), __fadd_rn(uy10, uy1m)), 0.25f);
    const float midx = __fmaf_rn(-vx1, half, posx), midy = __fmaf_rn(-vy1, half, posy);
    const float vx2 = bicubic_pk<SLAB>(vx.d, pitch, lim.x, lim.y, vx.h, __fsub_rn(midx, 0.5f), midy, tr, tr.span_x, tr.x_lo, tr.x_hi);
    const float vy2 = bicubic_pk<SLAB>(vy.d, pitch, lim.z, lim.w, vy.h, midx, __fsub_rn(midy, 0.5f), tr, tr.span_y, tr.y_lo, tr.y_hi);
    const float endx = __fmaf_rn(-vx2, full, posx), endy = __fmaf_rn(-vy2, full, posy);
    const float xvel = bicubic_pk<SLAB>(vx.d, pitch, lim.x, lim.y, vx.h, __fsub_rn(endx, 0.5f), endy, tr, tr.span_x, tr.x_lo, tr.x_hi);
    nvx = __fmul_rn(__fmul_rn(xvel, fC), fE);
    vxb.d[o] = nvx;
  }
  if (acty) { // simulation.cpp:300-347
    const float posy = __fadd_rn(fy, 0.5f), posx = fx;
    const float vy1 = uy00;
    const float vx1 = __fmul_rn(__fadd_rn(__fadd_rn(ux00, ux0m), __fadd_rn(ux10, ux1m)), 0.25f);
    const float midx = __fmaf_rn(-vx1, half, posx), midy = __fmaf_rn(-vy1, half, posy);
    const float vx2 = bicubic_pk<SLAB>(vx.d, pitch, lim.x, lim.y, vx.h, __fsub_rn(midx, 0.5f), midy, tr, tr.span_x, tr.x_lo, tr.x_hi);
    const float vy2 = bicubic_pk<SLAB>(vy.d, pitch, lim.z, lim.w, vy.h, midx, __fsub_rn(midy, 0.5f), tr, tr.span_y, tr.y_lo, tr.y_hi);
    const float endx = __fmaf_rn(-vx2, full, posx), endy = __fmaf_rn(-vy2, full, posy);
    const float yvel = bicubic_pk<SLAB>(vy.d, pitch, lim.z, lim.w, vy.h, endx, __fsub_rn(endy, 0.5f), tr, tr.span_y, tr.y_lo, tr.y_hi);
    nvy = __fmul_rn(__fmul_rn(yvel, fC), fN);
    vyb.d[o] = nvy;
  }
  if (DIV) {
    __shared__ float south[8][32];
    const bool in = row_ok && xi <= W - 2;
    if (in && !actx) nvx = vxb.d[o];
    if (in && !acty) nvy = vyb.d[o];
    const float west = __shfl_up_sync(0xffffffffu, nvx, 1);
    south[threadIdx.y][lane] = nvy;
    __syncthreads();
    if (in && lane > 0 && threadIdx.y > 0 && xi <= W - 3 && y <= H - 3)
      fdiv[o] = __fmul_rn(nih, __fsub_rn(__fadd_rn(__fsub_rn(nvx, west), nvy), south[threadIdx.y - 1][lane]));
  }
}

// The cells k_advect_xy<DIV> leaves out: x = 1 (mod 32) (a CTA's first column), rows y_lo + 8k (its
// first row), and x = W-2, y = H-2, whose east / north faces setVBCs writes.  One launch, 256-thread
// blocks: the first n_row_blocks take 256-cell segments of the edge rows, the rest 32 list columns
// (1, 33, 65, ..., and W-2) x 8 rows each; cells on both lists are written twice with the same value.
__global__ void k_divergence_edges(Grid vx, Grid vy, Grid f, float nih, int y_lo, int y_hi, int n_row_blocks,
                                   int segs, int n_edge_rows, int col_blocks) {
  ubgl_pdl_prologue();
  const int W = f.w, H = f.h;
  auto cell = [&](int x, int y) {
    const float d = __fsub_rn(__fadd_rn(__fsub_rn(vx.at(x, y), vx.at(x - 1, y)), vy.at(x, y)), vy.at(x, y - 1));
    f.at(x, y) = __fmul_rn(nih, d);
  };
  const int tid = threadIdx.y * 32 + threadIdx.x;
  if ((int)blockIdx.x < n_row_blocks) {
    const int er = blockIdx.x / segs, seg = blockIdx.x - er * segs;
    const int y = er < n_edge_rows - 1 ? y_lo + 8 * er : H - 2; // the last entry of the list is row H-2
    const int x = 1 + seg * 256 + tid;
    if (y >= y_lo && y < y_hi && x <= W - 2) cell(x, y);
    return;
  }
  const int b = blockIdx.x - n_row_blocks;
  const int rb = b / col_blocks, cb = b - rb * col_blocks;
  const int y = y_lo + 8 * rb + threadIdx.y;
  const int c = cb * 32 + threadIdx.x;
  const int ncol = (W - 2 + 31) / 32; // x = 1 + 32 c <= W-2
  if (y >= y_hi) return;
  if (c < ncol) cell(1 + 32 * c, y);
  else if (c == ncol) cell(W - 2, y);
}

// y_lo / y_hi: the row range of the advect launch that computed the rest (after its clamps)
void launch_divergence_edges(const Grid &vx, const Grid &vy, const Grid &f, float ih, int y_lo, int y_hi,
                             cudaStream_t stream, LaunchCounter *lc) {
  y_lo = std::max(y_lo, 1);
  y_hi = std::min(y_hi, f.h - 1);
  if (y_hi <= y_lo) return;
  const int W = f.w;
  const int segs = ceil_div(W - 2, 256);
  const int n_edge_rows = ceil_div(y_hi - y_lo, 8) + 1; // rows y_lo + 8k, then H-2 (out of range on a slab that does not hold it)
  const int ncol = (W - 2 + 31) / 32 + 1;
  const int col_blocks = ceil_div(ncol, 32);
  const int n_row_blocks = segs * n_edge_rows;
  const int n_col_blocks = col_blocks * ceil_div(y_hi - y_lo, 8);
  UBGL_LAUNCH(lc, K_DIVERGENCE, 0, stream,
              (launch_k(k_divergence_edges, n_row_blocks + n_col_blocks, dim3(32, 8), 0, stream, vx, vy, f, -ih, y_lo, y_hi, n_row_blocks, segs,
                                                                                          n_edge_rows, col_blocks)));
}

// project part 1 (simulation.cpp:166-171): f = -(1/h) div v on the interior
__global__ void k_divergence(Grid vx, Grid vy, Grid f, float ih) {
  int x = 1 + blockIdx.x * blockDim.x + threadIdx.x;
  int y = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= f.w - 1 || y >= f.h - 1) return;
  float d = __fsub_rn(__fadd_rn(__fsub_rn(vx.at(x, y), vx.at(x - 1, y)), vy.at(x, y)),
                      vy.at(x, y - 1));
  f.at(x, y) = __fmul_rn(-ih, d);
}

// sinks (simulation.cpp:173-182): 3x3 stamps, in list order (later sinks win)
__global__ void k_stamp_sinks(Grid f, const float *sinks, int n, int y_lo, int y_hi) {
  ubgl_pdl_prologue();
  int dx = (int)(threadIdx.x % 3) - 1, dy = (int)(threadIdx.x / 3) - 1;
  for (int k = 0; k < n; k++) {
    if (threadIdx.x < 9) {
      int ix = (int)sinks[3 * k], iy = (int)sinks[3 * k + 1];
      if (iy + dy >= y_lo && iy + dy < y_hi) f.at(ix + dx, iy + dy) = sinks[3 * k + 2];
    }
    __syncthreads();
  }
}

void launch_stamp_sinks(const Grid &f, const float *d_sinks, int n, int y_lo, int y_hi,
                        cudaStream_t stream, LaunchCounter *lc) {
  UBGL_LAUNCH(lc, K_SINKS, LVL, stream, launch_k(k_stamp_sinks, 1, 32, 0, stream, f, d_sinks, n, y_lo, y_hi));
}

// setPBC (simulation.cpp:36-45): columns for all y, then rows for all x.
__global__ void k_set_pbc(Grid p, int bcW, int bcE, int bcN, int bcS) {
  for (int y = threadIdx.x; y < p.h; y += blockDim.x) {
    p.at(0, y) = single_pbc(bcW, p.at(1, y));
    p.at(p.w - 1, y) = single_pbc(bcE, p.at(p.w - 2, y));
  }
  __syncthreads();
  for (int x = threadIdx.x; x < p.w; x += blockDim.x) {
    p.at(x, 0) = single_pbc(bcS, p.at(x, 1));
    p.at(x, p.h - 1) = single_pbc(bcN, p.at(x, p.h - 2));
  }
}

// gradient subtraction (simulation.cpp:196-207)
__global__ void k_gradient(Grid vx, Grid vy, Grid p, Grid flag, float ih) {
  int x = 1 + blockIdx.x * blockDim.x + threadIdx.x;
  int y = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  const int W = p.w, H = p.h;
  if (x >= W - 1 || y >= H - 1) return;
  float pc = p.at(x, y), fc = flag.at(x, y);
  if (x < W - 2) {
    float m = __fmul_rn(__fmul_rn(fc, flag.at(x + 1, y)), ih);
    vx.at(x, y) = __fmaf_rn(-m, __fsub_rn(p.at(x + 1, y), pc), vx.at(x, y));
  }
  if (y < H - 2) {
    float m = __fmul_rn(__fmul_rn(fc, flag.at(x, y + 1)), ih);
    vy.at(x, y) = __fmaf_rn(-m, __fsub_rn(p.at(x, y + 1), pc), vy.at(x, y));
  }
}

// ---------------------------------------------------------------------------
// DeviceSim
// ---------------------------------------------------------------------------
DeviceSim::DeviceSim(const float *host_flag, int W_, int H_, float pwidth_, float mu_, int device_)
    : W(W_), H(H_), pwidth(pwidth_), mu(mu_), device(device_) {
  UBGL_REQUIRE(W >= 8 && H >= 8, "Simulation needs W,H >= 8");
  UBGL_REQUIRE(host_flag != nullptr, "flag must not be null");
  UBGL_CUDA(cudaSetDevice(device));
  UBGL_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  pitch = round_up(W, 32);
  for (int b = 0; b < 3; b++) {
    vxb[b] = alloc_grid(W - 1, H, pitch);
    vyb[b] = alloc_grid(W, H - 1, pitch);
  }
  vx_accum = alloc_grid(W - 1, H, pitch);
  vy_accum = alloc_grid(W, H - 1, pitch);
  p = alloc_grid(W, H, pitch);
  f = alloc_grid(W, H, pitch);
  flag = alloc_grid(W, H, pitch);
  // Simulation(flag,pwidth,mu) simulation.hpp:32-67
  bcS = 3; bcN = 3; bcW = 0; bcE = 2;
  h = pwidth / ((float)W - 1.0f);
  if (const char *e = getenv("UBGL_LAZY_CURRENT")) lazy_current = e[0] != '0'; // A/B of the aliasing
  mg.reset(new DeviceMG(W, H, device, stream, &lc));
  upload_grid(flag, host_flag, W, H, stream);
  {
    std::vector<float> col(H, 1.0f); // vx.f(0,y) = vx.b(0,y) = 1 (:58-60)
    for (int b = 0; b < 2; b++)
      UBGL_CUDA(cudaMemcpy2DAsync(vxb[b].d, sizeof(float) * pitch, col.data(), sizeof(float),
                                  sizeof(float), H, cudaMemcpyHostToDevice, stream));
    UBGL_CUDA(cudaStreamSynchronize(stream));
  }
  mg->update_fields(flag);
  mg->prepare_mask0(flag);
  UBGL_CUDA(cudaStreamSynchronize(stream));
}

DeviceSim::~DeviceSim() {
  cudaSetDevice(device);
  if (stream) cudaStreamSynchronize(stream);
  mg.reset();
  for (int b = 0; b < 3; b++) {
    free_grid(vxb[b]);
    free_grid(vyb[b]);
  }
  free_grid(vx_accum); free_grid(vy_accum);
  free_grid(p); free_grid(f); free_grid(flag); free_grid(r);
  drop_graphs();
  if (d_sinks) cudaFree(d_sinks);
  if (d_fnorm) cudaFree(d_fnorm);
  if (h_stage) cudaFreeHost(h_stage);
  if (d_pack) cudaFree(d_pack);
  if (d_out) cudaFree(d_out);
  if (s_in) { cudaStreamSynchronize(s_in); cudaStreamDestroy(s_in); }
  if (s_out) { cudaStreamSynchronize(s_out); cudaStreamDestroy(s_out); }
  for (cudaEvent_t e : {ev_up, ev_pack, ev_down, ev_main})
    if (e) cudaEventDestroy(e);
  if (ev_stage) cudaEventDestroy(ev_stage);
  if (d_vxy) cudaFree(d_vxy);
  if (d_mag) cudaFree(d_mag);
  if (stream) cudaStreamDestroy(stream);
}

Grid DeviceSim::field(int id) {
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
  case F_R:
    if (!r.d) r = alloc_grid(W, H, pitch);
    return r;
  case F_VX_CURRENT: return vxb[cur_alias ? ixf : ixc];
  case F_VY_CURRENT: return vyb[cur_alias ? iyf : iyc];
  }
  throw ArgError{"unknown field id"};
}

void DeviceSim::will_write(int id) {
  if (!cur_alias) return;
  if (id != F_VX && id != F_VY && id != F_VX_CURRENT && id != F_VY_CURRENT) return;
  cur_alias = false;
  save_current(); // stream-ordered device copy front -> the spare third buffer
}

void DeviceSim::field_size(int id, int *w, int *hh) const {
  switch (id) {
  case F_VX: case F_VXB: case F_VX_ACCUM: case F_VX_CURRENT: *w = W - 1; *hh = H; return;
  case F_VY: case F_VYB: case F_VY_ACCUM: case F_VY_CURRENT: *w = W; *hh = H - 1; return;
  default: *w = W; *hh = H; return;
  }
}

void DeviceSim::upload(int id, const float *host) {
  UBGL_REQUIRE(host != nullptr, "upload: null host pointer");
  will_write(id);
  Grid g = field(id);
  upload_grid(g, host, g.w, g.h, stream);
  UBGL_CUDA(cudaStreamSynchronize(stream)); // host buffer is only borrowed for the call
  if (id == F_FLAG) flag_changed(false);
}

// dst(x, y) += src[y * w + x]: host contributions to a device-resident accumulator
__global__ void k_add_packed(Grid dst, const float *__restrict__ src, int border) {
  const size_t y = blockIdx.y + border;
  for (int x = border + blockIdx.x * blockDim.x + threadIdx.x; x < dst.w - border; x += gridDim.x * blockDim.x)
    dst.d[y * dst.pitch + x] = __fadd_rn(dst.d[y * dst.pitch + x], src[y * dst.w + x]);
}

// += of a host grid onto a device field.  The accumulators are written from both sides of the
// C ABI: the items kernels scatter into the device copies with atomicAdd (next.cu), a host
// caller (advect_floating_items.cpp:118-120 running on the CPU) adds into its mirror; a plain
// upload of the mirror would erase what the device has collected since the last step.
void DeviceSim::upload_add(int id, const float *host) {
  UBGL_REQUIRE(host != nullptr, "upload_add: null host pointer");
  UBGL_REQUIRE(id != F_FLAG, "upload_add: not meaningful for the flag field");
  will_write(id);
  Grid g = field(id);
  const size_t n = (size_t)g.w * g.h;
  if (n > cap_pack) {
    if (d_pack) UBGL_CUDA(cudaFree(d_pack));
    d_pack = nullptr;
    cap_pack = 0;
    UBGL_CUDA(cudaMalloc(&d_pack, sizeof(float) * (size_t)W * H));
    cap_pack = (size_t)W * H;
  }
  UBGL_CUDA(cudaMemcpyAsync(d_pack, host, sizeof(float) * n, cudaMemcpyHostToDevice, stream));
  // the accumulators exist on the interior only (simulation.cpp:380-394 applies and clears
  // 1..w-2 x 1..h-2); their border cells are not part of the sum
  const int border = (id == F_VX_ACCUM || id == F_VY_ACCUM) ? 1 : 0;
  dim3 grid(std::min(ceil_div(g.w, 256), 8), g.h - 2 * border);
  UBGL_LAUNCH(&lc, K_ACCUM, 0, stream, k_add_packed<<<grid, 256, 0, stream>>>(g, d_pack, border));
  UBGL_CUDA(cudaStreamSynchronize(stream)); // host buffer is only borrowed for the call
}

// The level-0 flag changed: refresh what is derived from it.  `pyramid` also
// rebuilds the coarse flags (MG::updateFields); a bare write to sim.flag does
// not, exactly as in the reference (simulation.hpp:85 vs ubootgl_app.cpp:111-112).
void DeviceSim::flag_changed(bool pyramid, bool binary_edit) {
  if (pyramid) mg->update_fields(flag);
  drop_graphs(); // the fused / plain decision and the mask build are outside the graphs
  const bool keep = binary_edit && mg->mask0_is_binary(); // the edit wrote only 0.0 / 1.0
  mg->invalidate_mask0();
  mg->prepare_mask0(flag, keep);
}

void DeviceSim::flag_edited_discs(const float *d_xyd, int n, int max_diam) {
  static const bool on = [] {
    const char *e = getenv("UBGL_DISC_UPDATE");
    return !(e && e[0] == '0');
  }();
  drop_graphs();
  if (on && mg->update_fields_discs(flag, d_xyd, n, max_diam)) return;
  flag_changed(true, true);
}

void DeviceSim::download(int id, float *host) {
  UBGL_REQUIRE(host != nullptr, "download: null host pointer");
  Grid g = field(id);
  download_grid(g, host, g.w, g.h, stream);
  UBGL_CUDA(cudaStreamSynchronize(stream));
}

void DeviceSim::update_flag(const float *host_flag) {
  UBGL_REQUIRE(host_flag != nullptr, "update_flag: null host pointer");
  upload_grid(flag, host_flag, W, H, stream);
  flag_changed(true);
  UBGL_CUDA(cudaStreamSynchronize(stream));
}

void DeviceSim::sync() { UBGL_CUDA(cudaStreamSynchronize(stream)); }

void DeviceSim::apply_accum() {
  UBGL_LAUNCH(&lc, K_ACCUM, LVL, stream, k_apply_accum<<<grd2d(W - 3, H - 2), blk2d(), 0, stream>>>(vxb[ixf], vx_accum));
  UBGL_LAUNCH(&lc, K_ACCUM, LVL, stream, k_apply_accum<<<grd2d(W - 2, H - 3), blk2d(), 0, stream>>>(vyb[iyf], vy_accum));
}

void DeviceSim::set_vbcs() {
  UBGL_LAUNCH(&lc, K_VBC, LVL, stream, k_set_vbcs<<<1, 1024, 0, stream>>>(vxb[ixf], vxb[ixb], vyb[iyf], vyb[iyb], bcW, bcE, bcN,
                                     bcS));
}

void DeviceSim::diffuse() {
  float a = dt * mu * ((float)W - 1.0f) / pwidth; // simulation.cpp:105
  float rden = 1.0f / (1.0f + 4.0f * a);
  for (int i = 1; i < 3; i++) {
    UBGL_LAUNCH(&lc, K_DIFFUSE, LVL, stream, k_diffuse_vx<<<grd2d(W - 3, H - 2), blk2d(), 0, stream>>>(vxb[ixf], vxb[ixb], flag, a, rden));
    std::swap(ixf, ixb);
    set_vbcs();
  }
  for (int i = 1; i < 3; i++) {
    UBGL_LAUNCH(&lc, K_DIFFUSE, LVL, stream, k_diffuse_vy<<<grd2d(W - 2, H - 3), blk2d(), 0, stream>>>(vyb[iyf], vyb[iyb], flag, a, rden));
    std::swap(iyf, iyb);
    set_vbcs();
  }
}

// 1: one face per thread (k_advect_vx / vy); 2: row-pair experiment; 3 (default): k_advect_xy
static int advect_variant() {
  static const int variant = [] {
    const char *e = getenv("UBGL_ADVECT_VARIANT");
    return (e && e[0] >= '1' && e[0] <= '3') ? e[0] - '0' : 3;
  }();
  return variant;
}

// Returns ADV_ZEROED when the launch also zeroed the accumulator interiors of rows [y_lo, y_hi)
// (k_advect_xy, needs the stencil mask, i.e. binary flags), ADV_DIV when it also wrote the
// divergence f = -ih div v of its cells except the edge set (fdiv != null; the caller then runs
// launch_divergence_edges after setVBCs instead of the divergence pass).
int launch_advect(const Grid &vx, const Grid &vy, const Grid &vxbk, const Grid &vybk,
                  const Grid &flag, float half, float full, int y_lo, int y_hi, const AdvectPeers *peers,
                  cudaStream_t stream, LaunchCounter *lc, const uint8_t *mask, float *ax, float *ay, float *fdiv,
                  float ih) {
  const int W = flag.w, H = flag.h;
  y_lo = std::max(y_lo, 1);
  y_hi = std::min(y_hi, H - 1);
  if (y_hi <= y_lo) return 0;
  dim3 b(32, 8);
  dim3 g(ceil_div(W - 2, 32), ceil_div(y_hi - y_lo, 8));
  const int variant = advect_variant();
  TapRows tr{};
  if (peers) {
    tr.lo = peers->st_lo; tr.hi = peers->st_hi; tr.plo = peers->peer_lo; tr.phi = peers->peer_hi;
    tr.x_lo = peers->vx_lo; tr.x_hi = peers->vx_hi; tr.y_lo = peers->vy_lo; tr.y_hi = peers->vy_hi;
    tr.err = peers->err;
    tr.lo1 = tr.lo + 1;
    // a slab thinner than 4 rows leaves no local sample at all: span wraps to "never"
    const int sx = std::min(tr.hi, vx.h) - 3 - tr.lo1, sy = std::min(tr.hi, vy.h) - 3 - tr.lo1;
    tr.span_x = sx >= 0 ? (unsigned)sx : 0u;
    tr.span_y = sy >= 0 ? (unsigned)sy : 0u;
  }
  if (variant == 3 && mask) {
    const float4 lim = make_float4((float)vx.w - 3.0f, (float)vx.h - 3.0f, (float)vy.w - 3.0f, (float)vy.h - 3.0f);
    // resident CTAs per SM the register budget is cut for.  Measured at 8192^2 (B200): 3: 1.64,
    // 4: 1.28, 5: 1.28, 6: 1.23, 8: 1.24 ms -- the kernel is latency bound, occupancy wins
    static const int occ = [] {
      const char *e = getenv("UBGL_ADVECT_OCC");
      return e ? atoi(e) : 6;
    }();
    // UBGL_ADVECT_FORCE_SLAB=1: the slab build of the kernel on a single GPU (every row is local),
    // to measure what its per-sample row test costs
    static const bool force_slab = [] {
      const char *e = getenv("UBGL_ADVECT_FORCE_SLAB");
      return e && e[0] == '1';
    }();
    if (force_slab && !peers) {
      tr.lo = tr.plo = 0; tr.hi = tr.phi = vx.h;
      tr.lo1 = 1;
      tr.span_x = (unsigned)(vx.h - 4); tr.span_y = (unsigned)(vy.h - 4);
    }
    // UBGL_ADVECT_DIV=1: the divergence in the advect epilogue (k_advect_xy<DIV> + k_divergence_edges)
    // instead of its own pass.  Bit-identical (tests/test_gpu_advect_variants.py) and OFF by default:
    // measured at 8192^2 it LOSES 0.12 ms per step -- the advect kernel goes from 1.230 to 1.319 ms (every
    // thread stays to the end, a block barrier, the retained faces re-read) and the edge pass (the
    // 1/32 of columns on CTA edges are sector-sized accesses) takes 0.153 ms against 0.121 ms for the
    // whole k_divergence4, which runs at 85 % of the HBM peak.
    static const bool div_on = [] {
      const char *e = getenv("UBGL_ADVECT_DIV");
      return e && e[0] == '1';
    }();
    const bool div = div_on && fdiv != nullptr && (occ == 6 || (peers && occ != 4));
    const float nih = -ih;
#define UBGL_ADV_XY(S_, O_) UBGL_LAUNCH(lc, K_ADVECT, LVL, stream, (k_advect_xy<S_, O_, false><<<g, b, 0, stream>>>(vx, vy, vxbk, vybk, mask, ax, ay, half, full, lim, y_lo, y_hi, tr, nullptr, 0.0f)))
#define UBGL_ADV_XY_DIV(S_) UBGL_LAUNCH(lc, K_ADVECT, LVL, stream, (k_advect_xy<S_, 6, true><<<g, b, 0, stream>>>(vx, vy, vxbk, vybk, mask, ax, ay, half, full, lim, y_lo, y_hi, tr, fdiv, nih)))
    if (div) {
      if (peers || force_slab) {
        UBGL_ADV_XY_DIV(true);
      } else {
        UBGL_ADV_XY_DIV(false);
      }
      return (ax != nullptr ? ADV_ZEROED : 0) | ADV_DIV;
    }
    if (peers || force_slab) {
      if (occ == 4) {
        UBGL_ADV_XY(true, 4);
      } else {
        UBGL_ADV_XY(true, 6);
      }
    } else if (occ == 3) {
      UBGL_ADV_XY(false, 3);
    } else if (occ == 5) {
      UBGL_ADV_XY(false, 5);
    } else if (occ == 6) {
      UBGL_ADV_XY(false, 6);
    } else if (occ == 7) {
      UBGL_ADV_XY(false, 7);
    } else if (occ == 8) {
      UBGL_ADV_XY(false, 8);
    } else {
      UBGL_ADV_XY(false, 4);
    }
#undef UBGL_ADV_XY
#undef UBGL_ADV_XY_DIV
    return ax != nullptr ? ADV_ZEROED : 0;
  }
  if (variant == 2) {
    dim3 g2(g.x, ceil_div(y_hi - y_lo, 16));
    if (peers) {
      UBGL_LAUNCH(lc, K_ADVECT, LVL, stream, (k_advect_pair<true, 0><<<g2, b, 0, stream>>>(vx, vy, vxbk, flag, half, full, y_lo, y_hi, tr)));
      UBGL_LAUNCH(lc, K_ADVECT, LVL, stream, (k_advect_pair<true, 1><<<g2, b, 0, stream>>>(vx, vy, vybk, flag, half, full, y_lo, y_hi, tr)));
    } else {
      UBGL_LAUNCH(lc, K_ADVECT, LVL, stream, (k_advect_pair<false, 0><<<g2, b, 0, stream>>>(vx, vy, vxbk, flag, half, full, y_lo, y_hi, tr)));
      UBGL_LAUNCH(lc, K_ADVECT, LVL, stream, (k_advect_pair<false, 1><<<g2, b, 0, stream>>>(vx, vy, vybk, flag, half, full, y_lo, y_hi, tr)));
    }
    return 0;
  }
  if (!peers) {
    UBGL_LAUNCH(lc, K_ADVECT, LVL, stream, k_advect_vx<false><<<g, b, 0, stream>>>(vx, vy, vxbk, flag, half, full, y_lo, y_hi, tr));
    UBGL_LAUNCH(lc, K_ADVECT, LVL, stream, k_advect_vy<false><<<g, b, 0, stream>>>(vx, vy, vybk, flag, half, full, y_lo, y_hi, tr));
    return 0;
  }
  UBGL_LAUNCH(lc, K_ADVECT, LVL, stream, k_advect_vx<true><<<g, b, 0, stream>>>(vx, vy, vxbk, flag, half, full, y_lo, y_hi, tr));
  UBGL_LAUNCH(lc, K_ADVECT, LVL, stream, k_advect_vy<true><<<g, b, 0, stream>>>(vx, vy, vybk, flag, half, full, y_lo, y_hi, tr));
  return 0;
}

void DeviceSim::advect() { advect_impl(false); }

// fused_step: the stencil mask is valid (binary flags) and the accumulators are consumed, so the
// merged kernel may run and clear them; returns whether it did
int DeviceSim::advect_impl(bool fused_step) {
  float ih = 1.0f / h;
  float half = 0.5f * dt * ih, full = dt * ih; // simulation.cpp:276,286
  const int did =
      launch_advect(vxb[ixf], vyb[iyf], vxb[ixb], vyb[iyb], flag, half, full, 1, H - 1, nullptr, stream, &lc,
                    fused_step ? mg->mask0_ptr() : nullptr, fused_step ? vx_accum.d : nullptr,
                    fused_step ? vy_accum.d : nullptr, fused_step ? f.d : nullptr, ih);
  std::swap(ixf, ixb);
  std::swap(iyf, iyb);
  return did;
}

// sum over the interior of (f * flag)^2: the residual norm of p = 0, the
// reference value of the relative residual in tolerance mode
__global__ void k_fnorm_sq(Grid f, Grid flag, double *out) {
  double acc = 0.0;
  for (int y = 1 + blockIdx.x; y < f.h - 1; y += gridDim.x)
    for (int x = 1 + threadIdx.x; x < f.w - 1; x += blockDim.x) {
      float v = f.at(x, y) * flag.at(x, y);
      acc += (double)v * (double)v;
    }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  __shared__ double part[32];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) atomicAdd(out, acc);
  }
}

// The pressure solves of project() (simulation.cpp:189-190).  Reference mode:
// a fixed number of warm-started V-cycles, no convergence test.  Tolerance mode
// (tol > 0): cycle until ||r|| <= tol * ||f*flag||, or the
// residual stagnates (a cycle removes less than `stag` of it), or max_cycles.
void DeviceSim::solve_cycles() {
  cycles_done = 0;
  res_hist.clear();
  if (tol <= 0.0f) {
    for (int c = 0; c < vcycles; c++) mg->solve(p, f, flag, h, true);
    cycles_done = vcycles;
    return;
  }
  if (!d_fnorm) UBGL_CUDA(cudaMalloc(&d_fnorm, sizeof(double)));
  UBGL_CUDA(cudaMemsetAsync(d_fnorm, 0, sizeof(double), stream));
  UBGL_LAUNCH(&lc, K_NORM, LVL, stream, k_fnorm_sq<<<std::min(H - 2, 148 * 8), 256, 0, stream>>>(f, flag, d_fnorm));
  double fn = 0.0;
  UBGL_CUDA(cudaMemcpyAsync(&fn, d_fnorm, sizeof(double), cudaMemcpyDeviceToHost, stream));
  float r = residual(); // warm start (syncs the stream)
  fnorm = (float)std::sqrt(fn);
  res_hist.push_back(r);
  const float target = tol * fnorm;
  while (cycles_done < max_cycles && r > target) {
    mg->solve(p, f, flag, h, true);
    cycles_done++;
    float rn = residual();
    res_hist.push_back(rn);
    const bool stalled = !(rn < stag * r);
    r = rn;
    if (stalled) break;
  }
}

void DeviceSim::project() {
  float ih = 1.0f / h;
  UBGL_LAUNCH(&lc, K_DIVERGENCE, LVL, stream, k_divergence<<<grd2d(W - 2, H - 2), blk2d(), 0, stream>>>(vxb[ixf], vyb[iyf], f, ih));
  project_sinks();
  solve_cycles();

  UBGL_LAUNCH(&lc, K_PBC, LVL, stream, k_set_pbc<<<1, 1024, 0, stream>>>(p, bcW, bcE, bcN, bcS));
  UBGL_LAUNCH(&lc, K_GRADIENT, LVL, stream, k_gradient<<<grd2d(W - 2, H - 2), blk2d(), 0, stream>>>(vxb[ixf], vyb[iyf], p, flag, ih));
}

// Pinned host staging buffer for small per-step lists (sink stamps, crater lists).  One
// buffer, reused: stage_host() waits until the previous copy out of it has completed.
float *DeviceSim::stage_host(size_t nfloats) {
  if (ev_stage) UBGL_CUDA(cudaEventSynchronize(ev_stage));
  else UBGL_CUDA(cudaEventCreateWithFlags(&ev_stage, cudaEventDisableTiming));
  if (nfloats > cap_stage) {
    if (h_stage) UBGL_CUDA(cudaFreeHost(h_stage));
    cap_stage = std::max<size_t>(nfloats * 2, 4096);
    UBGL_CUDA(cudaMallocHost(&h_stage, sizeof(float) * cap_stage));
  }
  return h_stage;
}
void DeviceSim::stage_done() { UBGL_CUDA(cudaEventRecord(ev_stage, stream)); }

__global__ void k_pack_rows(const float *__restrict__ src, int pitch, int w, float *__restrict__ dst) {
  const size_t y = blockIdx.y;
  for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < w; x += gridDim.x * blockDim.x)
    dst[y * w + x] = src[y * pitch + x];
}
// rows [0, rows) of a pitched array (src = its first row) packed to w floats per row
void launch_pack_rows(const float *src, int pitch, int w, int rows, float *dst, cudaStream_t stream,
                      LaunchCounter *lc) {
  if (rows <= 0) return;
  dim3 grid(std::min(ceil_div(w, 256), 8), rows);
  UBGL_LAUNCH(lc, K_OTHER, 0, stream, k_pack_rows<<<grid, 256, 0, stream>>>(src, pitch, w, dst));
}
float *DeviceSim::packed(const Grid &g) {
  const size_t n = (size_t)g.w * g.h;
  if (n > cap_pack) {
    if (d_pack) UBGL_CUDA(cudaFree(d_pack)); // synchronises: no copy out of the old buffer is left
    d_pack = nullptr;
    cap_pack = 0;
    UBGL_CUDA(cudaMalloc(&d_pack, sizeof(float) * (size_t)W * H));
    cap_pack = (size_t)W * H;
  }
  dim3 grid(std::min(ceil_div(g.w, 256), 8), g.h);
  UBGL_LAUNCH(&lc, K_OTHER, 0, stream, k_pack_rows<<<grid, 256, 0, stream>>>(g.d, g.pitch, g.w, d_pack));
  return d_pack;
}

void DeviceSim::project_sinks() {
  // sinks (simulation.cpp:173-187): grid position, border skip, decay and erase
  // are host-side list work exactly as in the reference; only the 3x3 stamps
  // touch device memory.
  std::vector<float> stamps;
  for (auto &s : sinks) {
    float gx = s.x / h + 0.5f, gy = s.y / h + 0.5f;
    if (gx <= 3 || gx > (float)(W - 3) || gy <= 3 || gy > (float)(H - 3)) continue;
    stamps.push_back((float)(int)gx);
    stamps.push_back((float)(int)gy);
    stamps.push_back(s.z);
    s.z = (float)((double)s.z * std::pow(0.000001, (double)(dt * 50)));
  }
  {
    size_t n = 0;
    for (size_t k = 0; k < sinks.size(); k++)
      if (!(sinks[k].z < 0.05f)) sinks[n++] = sinks[k];
    sinks.resize(n);
  }
  if (!stamps.empty()) {
    int n = (int)stamps.size() / 3;
    if (n > cap_sinks) {
      if (d_sinks) UBGL_CUDA(cudaFree(d_sinks));
      cap_sinks = n * 2;
      UBGL_CUDA(cudaMalloc(&d_sinks, sizeof(float) * 3 * cap_sinks));
    }
    // pinned staging, so the copy is asynchronous and the host keeps running ahead of the GPU
    float *hs = stage_host(stamps.size());
    std::memcpy(hs, stamps.data(), sizeof(float) * stamps.size());
    UBGL_CUDA(cudaMemcpyAsync(d_sinks, hs, sizeof(float) * stamps.size(), cudaMemcpyHostToDevice, stream));
    stage_done();
    launch_stamp_sinks(f, d_sinks, n, 0, H, stream, &lc);
  }
}

void DeviceSim::save_current() {
  UBGL_CUDA(cudaMemcpyAsync(vxb[ixc].d, vxb[ixf].d, vxb[ixc].bytes(), cudaMemcpyDeviceToDevice,
                            stream));
  UBGL_CUDA(cudaMemcpyAsync(vyb[iyc].d, vyb[iyf].d, vyb[iyc].bytes(), cudaMemcpyDeviceToDevice,
                            stream));
}

void DeviceSim::stage(int st, float dt_) {
  dt = dt_;
  will_write(F_VX); // the stages work on the front buffers in place
  switch (st) {
  case ST_ACCUM: apply_accum(); break;
  case ST_DIFFUSE: diffuse(); break;
  case ST_ADVECT: advect(); break;
  case ST_SETVBCS: set_vbcs(); break;
  case ST_PROJECT: project(); break;
  case ST_SAVE: save_current(); break;
  default: throw ArgError{"unknown stage id"};
  }
}

void DeviceSim::drop_graphs() {
  for (auto *c : {&graphs_a, &graphs_b}) {
    for (auto &kv : *c)
      if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    c->clear();
  }
}

// Runs one part of the fused step: replays its graph when one exists for the current buffer
// roles (and dt), captures one the third time the key is seen, otherwise just enqueues.
template <class F>
void DeviceSim::run_part(std::map<int, StepGraph> &cache, bool dt_dependent, bool graphable, F &&enqueue) {
  if (!graphable) {
    enqueue();
    return;
  }
  const int key = ixf | (ixb << 2) | (ixc << 4) | (iyf << 6) | (iyb << 8) | (iyc << 10);
  StepGraph &g = cache[key];
  if (g.exec && (!dt_dependent || g.dt == dt)) {
    UBGL_CUDA(cudaGraphLaunch(g.exec, stream));
    lc.n += g.launches;
    ixf = g.post[0]; ixb = g.post[1]; ixc = g.post[2]; iyf = g.post[3]; iyb = g.post[4]; iyc = g.post[5];
    return;
  }
  if (dt_dependent && g.dt != dt) { // a new time step: wait until it repeats before capturing
    if (g.exec) { // captured with the old dt's constants baked in: must never be replayed again
      cudaGraphExecDestroy(g.exec);
      g.exec = nullptr;
    }
    g.dt = dt;
    g.seen = 0;
  }
  if (++g.seen < 3) {
    enqueue();
    return;
  }
  if (g.exec) {
    cudaGraphExecDestroy(g.exec);
    g.exec = nullptr;
  }
  const long long n0 = lc.n;
  cudaGraph_t graph = nullptr;
  UBGL_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
  bool ok = true;
  try {
    enqueue();
  } catch (...) {
    ok = false;
  }
  cudaError_t e = cudaStreamEndCapture(stream, &graph);
  if (ok && e == cudaSuccess && graph && cudaGraphInstantiate(&g.exec, graph, 0) == cudaSuccess) {
    g.launches = lc.n - n0;
    g.dt = dt;
    g.post[0] = ixf; g.post[1] = ixb; g.post[2] = ixc; g.post[3] = iyf; g.post[4] = iyb; g.post[5] = iyc;
    cudaGraphDestroy(graph);
    UBGL_CUDA(cudaGraphLaunch(g.exec, stream)); // the captured work has not run yet
    return;
  }
  // capture failed: nothing was enqueued; give up on graphs for this handle and run directly
  if (graph) cudaGraphDestroy(graph);
  cudaGetLastError();
  g.exec = nullptr;
  use_graph = false;
  throw ArgError{"internal: CUDA graph capture of the step failed (UBGL_OPT_GRAPH is now off for this handle)"};
}

// Simulation::step (simulation.cpp:356-374)
void DeviceSim::step(float dt_) {
  dt = dt_;
  const bool fz = fused && mg->mask0_is_binary();
  cudaEvent_t ev[8];
  if (timing)
    for (auto &e : ev) UBGL_CUDA(cudaEventCreate(&e));
  int k = 0;
  auto mark = [&]() { if (timing) UBGL_CUDA(cudaEventRecord(ev[k++], stream)); };
  mark();
  if (!fz) will_write(F_VX);
  if (fz) {
    cur_alias = false; // the third buffers are scratch during the fused step, as before
    // small grids replay CUDA graphs of the two launch sequences (see sim.cuh)
    const bool graphable = use_graph && !timing && !lc.prof && tol <= 0.0f && (size_t)W * H <= ((size_t)1 << 22);
    mark(); // applyAccumulatedVelocity is part of the fused diffuse pass
    run_part(graphs_a, true, graphable, [&]() {
      fused_prestep();
      fused_borders(false, false);
      mark();
      const int adv = advect_impl(true);
      mark();
      fused_borders(false, false);
      mark();
      if (adv & ADV_DIV) // the advect epilogue wrote f except on CTA edges and next to the BC faces
        launch_divergence_edges(vxb[ixf], vyb[iyf], f, 1.0f / h, 1, H - 1, stream, &lc);
      else
        fused_divergence(!(adv & ADV_ZEROED));
    });
    project_sinks();
    run_part(graphs_b, false, graphable, [&]() {
      solve_cycles();
      fused_gradient_save(); // reads only interior p: independent of setPBC
      mark();
      fused_borders(true, !lazy_current); // setPBC + setVBCs, also into vx_current / vy_current
    });
    if (graphable) { // what solve_cycles() records on the host
      cycles_done = vcycles;
      res_hist.clear();
    }
    cur_alias = lazy_current;
    mark();
    mark();
  } else {
    apply_accum();
    mark();
    diffuse();
    mark();
    advect();
    mark();
    set_vbcs();
    mark();
    project();
    mark();
    set_vbcs();
    mark();
    save_current();
    mark();
  }
  if (timing) {
    UBGL_CUDA(cudaStreamSynchronize(stream));
    float ms[7];
    for (int i = 0; i < 7; i++) UBGL_CUDA(cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]));
    stage_ms[ST_ACCUM] = ms[0];
    stage_ms[ST_DIFFUSE] = ms[1];
    stage_ms[ST_ADVECT] = ms[2];
    stage_ms[ST_SETVBCS] = ms[3] + ms[5];
    stage_ms[ST_PROJECT] = ms[4];
    stage_ms[ST_SAVE] = ms[6];
    for (auto &e : ev) cudaEventDestroy(e);
  }
}

float DeviceSim::residual() {
  Grid rr = field(F_R);
  mg->residual(p, f, flag, rr, h, true);
  return mg->residual_norm_result();
}

void DeviceSim::mg_solve_ex(float hh, bool zgbc, int cycles) {
  for (int c = 0; c < cycles; c++) mg->solve(p, f, flag, hh, zgbc);
}

void DeviceSim::mg_solve(int cycles) {
  for (int c = 0; c < cycles; c++) mg->solve(p, f, flag, h, true);
}

} // namespace ubgl
