// mg_fused.cu -- temporally blocked multigrid kernels (sm_100a).
//
// Three schedules of the same passes, bit-identical to each other and to the plain
// one-kernel-per-operator path of mg.cu (tests/test_gpu_fused.py):
//   k_mg_run  (default)  PRE / POST with the thread's cells held in REGISTERS: a thread owns
//                        4 window rows x 8 cells, keeps f*h*h and the smoothing weights in
//                        registers over all six half-sweeps, reads the other colour through
//                        a rolling set of six 128-bit loads and gets its lane-edge neighbour
//                        by shuffle; POST applies prolongation + correction in registers
//                        while staging; the next window is prefetched into L2.
//   k_mg_tail (default)  levels below ~8K cells down to the coarsest: the whole bottom of
//                        the V-cycle in one shared-memory resident single-CTA launch.
//   k_mg_tile            the first schedule (window of p, f, mask in shared memory, 4 cells
//                        per thread and sweep); still used for the 5-sweep coarsest pass
//                        when the tail is off, selectable with UBGL_OPT_FUSED = 1 as the
//                        cross-check.
//
// A pass stages a (128 x LH) window of a level (the valid region shrinks by one cell per
// half-sweep, so the window carries a halo of 2S (+2) cells) and fuses the neighbouring
// multigrid operators:
//   MODE_PRE    : S x (rbgs [+ zero-gradient BC]) -> residual -> full-weighting
//                 restriction.  HBM traffic per cell: read p, f, mask (9 B),
//                 write p (4 B) + rc (1 B); the residual field is never stored.
//   MODE_POST   : prolongate + correct [+ BC] -> S x (rbgs [+ BC]).
//   MODE_SMOOTH : S x rbgs from a zero initial guess (coarsest level).
// p is ping-ponged between two buffers (tiles read their neighbours' cells as
// halo, so an in-place update would race).
//
// Common conventions:
//  * red and black cells live in separate shared-memory planes (plane =
//    (x+y)&1, column x>>1), so a half-sweep touches consecutive words;
//  * solid cells are stored as 0 in the window (their value is never used by
//    the reference either: every read is multiplied by the cell's flag,
//    pressure_solver.cpp:13-17,101-108), so the 5-point sum needs no flag
//    multiplies; the divide by the fluid-neighbour count and the centre flag
//    collapse into one multiply by a table weight (stencils.cuh rcp_count);
//  * border cells (never smoothed) keep flag*value in the window for their
//    neighbours and are re-materialised from p_in / the zero-gradient copy at
//    write-back.
//
// Reference semantics: pressure_solver.cpp:10-24 (smoothingKernel), :35-72
// (canonical red-black order), :91-116 (residual), :118-132 (restrict),
// :134-181 (prolongate, correct), :183-192 (setZeroGradientBC), :201-248
// (solveLevel).
#include "mg.cuh"
#include "stencils.cuh"
#include "packed.cuh"
#include <cstdlib>
#include <type_traits>

namespace ubgl {

enum { MODE_PRE = 0, MODE_POST = 1, MODE_SMOOTH = 2 };

// stencil mask byte: bit0 flag(c), bit1 W, bits2-4 weight code (0 = solid or no
// fluid neighbour, else the fluid-neighbour count 1..4), bit5 E, bit6 S, bit7 N
enum { MB_C = 1, MB_W = 2, MB_E = 32, MB_S = 64, MB_N = 128 };

struct TileArgs {
  const float *p_in; // nullptr: initial guess is 0 (levels >= 1 start from ec.fill(0))
  float *p_out;
  const float *f;
  const uint8_t *mask;
  int w, h, pitch;
  float *rc;          // MODE_PRE: restricted residual (coarse grid)
  const float *ec;    // MODE_POST: coarse error
  const uint8_t *maskc; // MODE_POST: stencil mask of the coarse level (coarse flag sums)
  int wc, hc, pc;
  float hh, ihsq;
  int zgbc;
  // row-slab decomposition (csrc/slab.cu): rows [st_lo, st_hi) of this level are
  // stored on this GPU (p_in, p_out, f, mask are addressed with GLOBAL row
  // indices), rows [own_lo, own_hi) are computed and written.  Single GPU: 0, h.
  int st_lo, st_hi, own_lo, own_hi;
  int c_lo, c_hi; // MODE_POST: rows of the coarse level stored here (single GPU: 0, hc)
  int pf_dist;    // k_mg_run: L2-prefetch the window this many CTAs ahead (0: off)
  int pair_sync;  // k_mg_run: neighbour-warp named barriers between half-sweeps (0: block barriers)
  int dbg;        // timing attribution only (UBGL_MG_DBG): 1 skips the sweeps, 2 the residual + restriction
};

constexpr int LW = 128;    // staged window width in cells
constexpr int HW = LW / 2; // cells per colour per row
constexpr int RS = 72;     // shared row stride (floats / bytes): HW + 4 left + 4 right pad
constexpr int XO = 4;      // column of xh = 0 inside a padded row

template <int S, int MODE> struct TileGeom {
  // halo: 2 cells per sweep, +2 for residual(+1) and restriction(+1); rounded
  // up to a multiple of 4 so that window rows start 16-byte aligned
  static constexpr int NEED = 2 * S + (MODE == MODE_PRE ? 2 : 0);
  static constexpr int HALO = (NEED + 3) / 4 * 4;
  static constexpr int TX = LW - 2 * HALO;
};

template <int LH> struct TileSmem {
  float P[2][LH][RS];
  float F[2][LH][RS];
  uint8_t M[2][LH][RS];
  float rcpt[8];
};

__device__ __forceinline__ float sel0(unsigned m, unsigned bit, float v) {
  return (m & bit) ? v : 0.0f;
}

template <int S, int MODE, int LH, int NT>
__global__ void __launch_bounds__(NT, 2) k_mg_tile(TileArgs a) {
  constexpr int HALO = TileGeom<S, MODE>::HALO;
  constexpr int TX = TileGeom<S, MODE>::TX;
  constexpr int TY = LH - 2 * HALO;
  constexpr int NW = NT / 32;
  constexpr int RPP = 2 * NW; // rows per pass of the 4-cells-per-thread loops
  static_assert(TY > 0 && TY % 2 == 0 && TX % 8 == 0 && NW % 2 == 0, "tile geometry");

  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem<LH> &sm = *reinterpret_cast<TileSmem<LH> *>(smem_raw);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x0 = blockIdx.x * TX, y0 = a.own_lo + blockIdx.y * TY;
  const int X0 = x0 - HALO, Y0 = y0 - HALO; // X0 % 4 == 0, Y0 even: local parity == global parity
  const int w = a.w, h = a.h;

  if (threadIdx.x < 8) sm.rcpt[threadIdx.x] = rcp_count(threadIdx.x);

  // ---- stage the window: coalesced 128-bit loads, de-interleaved by colour ----
  for (int r = warp; r < LH; r += NW) {
    const int gy = Y0 + r, gx = X0 + 4 * lane;
    float4 pv = make_float4(0.f, 0.f, 0.f, 0.f), fv = pv;
    uchar4 mv = make_uchar4(0, 0, 0, 0);
    if (gy >= a.st_lo && gy < a.st_hi && gx >= 0 && gx < a.pitch) {
      const size_t o = (size_t)gy * a.pitch + gx;
      if (a.p_in) pv = *reinterpret_cast<const float4 *>(a.p_in + o);
      fv = __ldg(reinterpret_cast<const float4 *>(a.f + o));
      mv = __ldg(reinterpret_cast<const uchar4 *>(a.mask + o));
    }
    pv.x = sel0(mv.x, MB_C, pv.x);
    pv.y = sel0(mv.y, MB_C, pv.y);
    pv.z = sel0(mv.z, MB_C, pv.z);
    pv.w = sel0(mv.w, MB_C, pv.w);
    if (MODE != MODE_PRE) { // the sweeps only need f*h*h; MODE_PRE keeps f for the residual
      fv.x = fh2_of(fv.x, a.hh);
      fv.y = fh2_of(fv.y, a.hh);
      fv.z = fh2_of(fv.z, a.hh);
      fv.w = fh2_of(fv.w, a.hh);
    }
    const int pr = r & 1, c = XO + 2 * lane;
    *reinterpret_cast<float2 *>(&sm.P[pr][r][c]) = make_float2(pv.x, pv.z);
    *reinterpret_cast<float2 *>(&sm.P[pr ^ 1][r][c]) = make_float2(pv.y, pv.w);
    *reinterpret_cast<float2 *>(&sm.F[pr][r][c]) = make_float2(fv.x, fv.z);
    *reinterpret_cast<float2 *>(&sm.F[pr ^ 1][r][c]) = make_float2(fv.y, fv.w);
    *reinterpret_cast<uchar2 *>(&sm.M[pr][r][c]) = make_uchar2(mv.x, mv.z);
    *reinterpret_cast<uchar2 *>(&sm.M[pr ^ 1][r][c]) = make_uchar2(mv.y, mv.w);
    if (lane < 2) { // pads read by the first / last group of a row
      const int pc_ = lane ? XO + HW : 0;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4 *>(&sm.P[0][r][pc_]) = z;
      *reinterpret_cast<float4 *>(&sm.P[1][r][pc_]) = z;
    }
  }
  __syncthreads();

  // 4-cells-per-thread mapping: 16 threads cover the 64 same-colour cells of a
  // row; a warp works on rows r and r+2 so that the x-parity q of its cells is
  // warp-uniform.
  const int tg = lane & 15;                                        // group within the row
  const int rslot = 4 * (warp >> 1) + (warp & 1) + 2 * (lane >> 4); // 0 .. RPP-1
  const int ci = XO + 4 * tg;                                      // padded column of the group

  // Per-thread 4-bit masks over its group for x-parity q = 0 / 1:
  // fz: cells on the global W/E border column (never updated);
  // in: cells with 1 <= gx <= w-2.
  unsigned fzb = 0, inb = 0; // bits 0-3: q = 0, bits 4-7: q = 1
#pragma unroll
  for (int q = 0; q < 2; q++) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int gx = X0 + 2 * (4 * tg + j) + q;
      if (gx == 0 || gx == w - 1) fzb |= 1u << (4 * q + j);
      if (gx >= 1 && gx <= w - 2) inb |= 1u << (4 * q + j);
    }
  }

  auto lo = [](int origin, int k) { return max(1, origin + k); };
  auto hi = [](int origin, int len, int n, int k) { return min(n - 1, origin + len - k); };

  // one colour of one sweep over the rows that are still exact at time k
  auto half_sweep = [&](int k, int cpar) {
    const int r_lo = lo(Y0, k) - Y0, r_hi = hi(Y0, LH, h, k) - Y0;
    for (int r = r_lo + rslot; r < r_hi; r += RPP) {
      const int q = (cpar + r) & 1;
      const float *po = &sm.P[cpar ^ 1][r][ci];
      const float4 A = *reinterpret_cast<const float4 *>(po);
      const float4 Sv = *reinterpret_cast<const float4 *>(po - RS);
      const float4 Nv = *reinterpret_cast<const float4 *>(po + RS);
      const float4 Fv = *reinterpret_cast<const float4 *>(&sm.F[cpar][r][ci]);
      const unsigned mw = *reinterpret_cast<const unsigned *>(&sm.M[cpar][r][ci]);
      float4 Wv, Ev;
      if (q == 0) {
        Wv = make_float4(po[-1], A.x, A.y, A.z);
        Ev = A;
      } else {
        Wv = A;
        Ev = make_float4(A.y, A.z, A.w, po[4]);
      }
      auto upd = [&](float pw, float pe, float ps, float pn, float f, int j) {
        float v = __fadd_rn(__fadd_rn(__fadd_rn(pw, pe), ps), pn);
        v = __fadd_rn(v, MODE == MODE_PRE ? fh2_of(f, a.hh) : f);
        return __fmul_rn(v, sm.rcpt[(mw >> (8 * j + 2)) & 7u]);
      };
      float4 v;
      v.x = upd(Wv.x, Ev.x, Sv.x, Nv.x, Fv.x, 0);
      v.y = upd(Wv.y, Ev.y, Sv.y, Nv.y, Fv.y, 1);
      v.z = upd(Wv.z, Ev.z, Sv.z, Nv.z, Fv.z, 2);
      v.w = upd(Wv.w, Ev.w, Sv.w, Nv.w, Fv.w, 3);
      float *pd = &sm.P[cpar][r][ci];
      const unsigned z = (fzb >> (4 * q)) & 15u;
      if (z == 0) {
        *reinterpret_cast<float4 *>(pd) = v;
      } else {
        if (!(z & 1)) pd[0] = v.x;
        if (!(z & 2)) pd[1] = v.y;
        if (!(z & 4)) pd[2] = v.z;
        if (!(z & 8)) pd[3] = v.w;
      }
    }
  };

  auto cell = [&](int gx, int gy) -> float & {
    const int lx = gx - X0, ly = gy - Y0;
    return sm.P[(lx + ly) & 1][ly][XO + (lx >> 1)];
  };
  auto mbyte = [&](int gx, int gy) -> unsigned {
    const int lx = gx - X0, ly = gy - Y0;
    return sm.M[(lx + ly) & 1][ly][XO + (lx >> 1)];
  };

  // setZeroGradientBC restricted to the border cells whose interior neighbour is
  // still exact at time k (corners are never touched); the window keeps
  // flag * value for border cells
  auto zero_gradient = [&](int k) {
    const int gx_lo = lo(X0, k), gx_hi = hi(X0, LW, w, k);
    const int gy_lo = lo(Y0, k), gy_hi = hi(Y0, LH, h, k);
    const int t = threadIdx.x;
    if (X0 <= 0 && gx_lo == 1)
      for (int gy = gy_lo + t; gy < gy_hi; gy += NT) cell(0, gy) = sel0(mbyte(0, gy), MB_C, cell(1, gy));
    if (w - 1 < X0 + LW && gx_hi == w - 1)
      for (int gy = gy_lo + t; gy < gy_hi; gy += NT)
        cell(w - 1, gy) = sel0(mbyte(w - 1, gy), MB_C, cell(w - 2, gy));
    if (Y0 <= 0 && gy_lo == 1)
      for (int gx = gx_lo + t; gx < gx_hi; gx += NT) cell(gx, 0) = sel0(mbyte(gx, 0), MB_C, cell(gx, 1));
    if (h - 1 < Y0 + LH && gy_hi == h - 1)
      for (int gx = gx_lo + t; gx < gx_hi; gx += NT)
        cell(gx, h - 1) = sel0(mbyte(gx, h - 1), MB_C, cell(gx, h - 2));
  };

  if (MODE == MODE_POST) {
    // prolongate + correct (pressure_solver.cpp:134-181) on every cell of the
    // window: one thread per COARSE cell (xc,yc) updates its four fine cells
    // (2xc,2yc) (2xc+1,2yc) (2xc,2yc+1) (2xc+1,2yc+1); same rounding sequence
    // as prolong_cell (stencils.cuh), the coarse flag sums come from the coarse
    // level's stencil mask (C, E, N bits of cell i and the E bit of the cell
    // above it).
    const int xcb = X0 >> 1, ycb = Y0 >> 1;
    for (int j = warp; j < LH / 2; j += NW) {
      const int yc = ycb + j, y = 2 * yc;
      if (yc < a.c_lo || yc >= a.c_hi) continue;
      const bool y_e = y >= 2 && y <= h - 2;                  // rows of the even-y loops (:140,:146)
      const bool y_o = y + 1 <= h - 3 && yc + 1 < a.c_hi;     // rows of the odd-y loops  (:153,:161)
      const float *ecr = a.ec + (size_t)yc * a.pc;
      const uint8_t *mcr = a.maskc + (size_t)yc * a.pc;
      for (int k = lane; k < HW; k += 32) {
        const int xc = xcb + k, x = 2 * xc;
        if (xc < 0 || xc >= a.wc) continue;
        const bool x_e = x >= 2 && x <= w - 2, x_o = x + 1 <= w - 3;
        const float e00 = __ldg(ecr + xc);
        const float e10 = x_o ? __ldg(ecr + xc + 1) : 0.0f;
        const float e01 = y_o ? __ldg(ecr + a.pc + xc) : 0.0f;
        const float e11 = (x_o && y_o) ? __ldg(ecr + a.pc + xc + 1) : 0.0f;
        const unsigned mc = __ldg(mcr + xc);
        const unsigned mn = y_o ? __ldg(mcr + a.pc + xc) : 0u;
        const int fC = mc & 1, fE = (mc >> 5) & 1, fN = (mc >> 7) & 1, fNE = (mn >> 5) & 1;
        const int c = XO + k;
        if (y_e) {
          if (x_e) {
            const float e = sel0(sm.M[0][2 * j][c], MB_C, e00);
            sm.P[0][2 * j][c] = __fadd_rn(sm.P[0][2 * j][c], e);
          }
          if (x_o) {
            const float e = __fmul_rn(sel0(sm.M[1][2 * j][c], MB_C, __fadd_rn(e00, e10)),
                                      prolong_rcp(fC + fE));
            sm.P[1][2 * j][c] = __fadd_rn(sm.P[1][2 * j][c], e);
          }
        }
        if (y_o) {
          if (x_e) {
            const float e = __fmul_rn(sel0(sm.M[1][2 * j + 1][c], MB_C, __fadd_rn(e00, e01)),
                                      prolong_rcp(fC + fN));
            sm.P[1][2 * j + 1][c] = __fadd_rn(sm.P[1][2 * j + 1][c], e);
          }
          if (x_o) {
            const float es = __fadd_rn(__fadd_rn(__fadd_rn(e00, e11), e10), e01);
            const float e = __fmul_rn(sel0(sm.M[0][2 * j + 1][c], MB_C, es),
                                      prolong_rcp(fC + fNE + fE + fN));
            sm.P[0][2 * j + 1][c] = __fadd_rn(sm.P[0][2 * j + 1][c], e);
          }
        }
      }
    }
    __syncthreads();
    if (a.zgbc) {
      zero_gradient(0);
      __syncthreads();
    }
  }

#pragma unroll 1
  for (int s = 0; s < S; s++) {
    half_sweep(2 * s + 1, 1); // "red":   (x+y) odd,  pressure_solver.cpp:35-40
    __syncthreads();
    half_sweep(2 * s + 2, 0); // "black": (x+y) even, pressure_solver.cpp:42-47
    __syncthreads();
    if (a.zgbc) {
      zero_gradient(2 * s + 2);
      __syncthreads();
    }
  }

  if (MODE == MODE_PRE) {
    // residual on (tile + 1 ring), written over f; 0 outside the interior
    const int r_lo = HALO - 1, r_hi = HALO + TY + 1;
#pragma unroll 1
    for (int cpar = 0; cpar < 2; cpar++) {
      for (int r = r_lo + rslot; r < r_hi; r += RPP) {
        const int q = (cpar + r) & 1;
        const int gy = Y0 + r;
        const bool rowin = gy >= 1 && gy <= h - 2;
        const float *po = &sm.P[cpar ^ 1][r][ci];
        const float4 A = *reinterpret_cast<const float4 *>(po);
        const float4 Sv = *reinterpret_cast<const float4 *>(po - RS);
        const float4 Nv = *reinterpret_cast<const float4 *>(po + RS);
        const float4 Cv = *reinterpret_cast<const float4 *>(&sm.P[cpar][r][ci]);
        const float4 Fv = *reinterpret_cast<const float4 *>(&sm.F[cpar][r][ci]);
        const unsigned mw = *reinterpret_cast<const unsigned *>(&sm.M[cpar][r][ci]);
        float4 Wv, Ev;
        if (q == 0) {
          Wv = make_float4(po[-1], A.x, A.y, A.z);
          Ev = A;
        } else {
          Wv = A;
          Ev = make_float4(A.y, A.z, A.w, po[4]);
        }
        const unsigned in = rowin ? (inb >> (4 * q)) & 15u : 0u;
        // residual_cell (stencils.cuh) with binary flags: every term
        // p_nb*flag_nb + p_c*(1-flag_nb) is exactly p_nb or p_c
        auto res = [&](float pc_, float pw, float pe, float ps, float pn, float f, int j) {
          const unsigned m = mw >> (8 * j);
          float val = (m & MB_W) ? pw : pc_;
          val = __fadd_rn(val, (m & MB_E) ? pe : pc_);
          val = __fadd_rn(val, (m & MB_S) ? ps : pc_);
          val = __fadd_rn(val, (m & MB_N) ? pn : pc_);
          val = __fmaf_rn(-4.0f, pc_, val);
          val = __fmul_rn(val, a.ihsq);
          val = __fadd_rn(f, val);
          return ((m & MB_C) && ((in >> j) & 1u)) ? val : 0.0f;
        };
        float4 v;
        v.x = res(Cv.x, Wv.x, Ev.x, Sv.x, Nv.x, Fv.x, 0);
        v.y = res(Cv.y, Wv.y, Ev.y, Sv.y, Nv.y, Fv.y, 1);
        v.z = res(Cv.z, Wv.z, Ev.z, Sv.z, Nv.z, Fv.z, 2);
        v.w = res(Cv.w, Wv.w, Ev.w, Sv.w, Nv.w, Fv.w, 3);
        *reinterpret_cast<float4 *>(&sm.F[cpar][r][ci]) = v;
      }
    }
    __syncthreads();
    // full-weighting restriction of the coarse cells whose fine centre (2xc,2yc)
    // lies in this tile; coarse border = 0 (rc.fill(0.0), pressure_solver.cpp:222)
    const int xc0 = x0 >> 1, yc0 = y0 >> 1;
    for (int j = warp; j < TY / 2; j += NW) {
      const int yc = yc0 + j;
      if (yc >= a.hc || 2 * yc >= a.own_hi) break;
      const int ly = 2 * yc - Y0;
      for (int i = lane; i < TX / 2; i += 32) {
        const int xc = xc0 + i;
        if (xc >= a.wc) break;
        float v = 0.0f;
        if (xc >= 1 && yc >= 1 && xc < a.wc - 1 && yc < a.hc - 1) {
          const int c = XO + ((2 * xc - X0) >> 1); // column of the (even,even) centre
          // rows ly-1 / ly+1: corners in plane 0, middle in plane 1; row ly: the opposite
          v = fw9(sm.F[0][ly - 1][c - 1], sm.F[1][ly - 1][c], sm.F[0][ly - 1][c],
                  sm.F[1][ly][c - 1], sm.F[0][ly][c], sm.F[1][ly][c], sm.F[0][ly + 1][c - 1],
                  sm.F[1][ly + 1][c], sm.F[0][ly + 1][c]);
        }
        a.rc[(size_t)yc * a.pc + xc] = v;
      }
    }
  }

  // ---- write the tile back (p ping-pong buffer), 128-bit stores ----
  for (int ly = HALO + warp; ly < HALO + TY; ly += NW) {
    const int gy = Y0 + ly;
    if (gy >= a.own_hi) break;
    const int pr = ly & 1;
    for (int qd = lane; qd < TX / 4; qd += 32) {
      const int lx = HALO + 4 * qd, gx = X0 + lx;
      if (gx >= w) break;
      const float2 e = *reinterpret_cast<const float2 *>(&sm.P[pr][ly][XO + (lx >> 1)]);
      const float2 o = *reinterpret_cast<const float2 *>(&sm.P[pr ^ 1][ly][XO + (lx >> 1)]);
      *reinterpret_cast<float4 *>(a.p_out + (size_t)gy * a.pitch + gx) =
          make_float4(e.x, o.x, e.y, o.y);
    }
  }

  // ---- border cells of this tile: the window held flag*value; the grid gets
  // the zero-gradient copy of the (now final) interior neighbour or, without
  // that BC and at the four corners, the unchanged input value ----
  const bool bx0 = x0 == 0, bx1 = (w - 1 >= x0 && w - 1 < x0 + TX);
  const bool by0 = y0 == 0, by1 = (h - 1 >= y0 && h - 1 < y0 + TY && h - 1 < a.own_hi);
  if (bx0 || bx1 || by0 || by1) {
    __syncthreads();
    const int t = threadIdx.x;
    auto orig = [&](int gx, int gy) {
      return a.p_in ? a.p_in[(size_t)gy * a.pitch + gx] : 0.0f;
    };
    auto put = [&](int gx, int gy, int nx, int ny) {
      const bool corner = (gx == 0 || gx == w - 1) && (gy == 0 || gy == h - 1);
      a.p_out[(size_t)gy * a.pitch + gx] = (a.zgbc && !corner) ? cell(nx, ny) : orig(gx, gy);
    };
    const int ty_hi = min(a.own_hi, y0 + TY), tx_hi = min(w, x0 + TX);
    // columns first, rows second: at the corners both write orig()
    if (bx0)
      for (int gy = y0 + t; gy < ty_hi; gy += NT) put(0, gy, 1, gy);
    if (bx1)
      for (int gy = y0 + t; gy < ty_hi; gy += NT) put(w - 1, gy, w - 2, gy);
    if (by0)
      for (int gx = x0 + t; gx < tx_hi; gx += NT) put(gx, 0, gx, 1);
    if (by1)
      for (int gx = x0 + t; gx < tx_hi; gx += NT) put(gx, h - 1, gx, h - 2);
  }
}

// ---------------------------------------------------------------------------
// k_mg_run -- the same PRE / POST passes (S = 3) with the thread's cells held in
// REGISTERS.  k_mg_tile above is bound by the shared-memory pipe (ncu: L1/shared
// 70 %, issue 50 %, DRAM 19 %): every 4-cell update re-reads f, the weight table
// and three rows of the other colour from shared memory.  Here a thread owns a
// run of R consecutive window rows x 8 consecutive cells (4 of each colour):
//  * f*h*h and the smoothing weights of its 8R cells are loaded from HBM ONCE into
//    registers and reused by all 6 half-sweeps (no F / weight traffic in the loop);
//  * a half-sweep walks the run top to bottom, so the other colour's rows r-1, r,
//    r+1 are a rolling register window: (R+2)/R instead of 3 128-bit loads per row;
//  * the W / E neighbour outside the thread's 8 cells comes from the adjacent lane
//    by shuffle instead of an unaligned 32-bit shared load.
// Shared-memory traffic per 4-cell update drops from 26 to ~11 crossbar cycles and
// the instruction count from ~50 to ~30.  R is even and window rows start even, so
// the colour of every register slot is known at compile time.  Per-cell arithmetic
// (order of additions, table weights) is unchanged: results are bit-identical to
// k_mg_tile and to the plain path (tests/test_gpu_fused.py).
// ---------------------------------------------------------------------------
// R = rows per thread (template parameter): 4 -> 16 row chunks, 256 threads, ~122 registers, 16 warps
// per SM; 2 -> 32 chunks, 512 threads, 64 registers, 32 warps per SM (twice the warps to hide the
// latencies this kernel is bound by, for 1.33x the other-colour loads per cell)
#ifndef UBGL_MG_ROWS_DEFAULT
#define UBGL_MG_ROWS_DEFAULT 4
#endif
constexpr int RUN_LH = 64;               // window rows
constexpr int RUN_HX = 8;                // x halo: whole 8-cell column groups

template <int MODE> struct RunGeom {
  static constexpr int S = 3;
  static constexpr int HY = (MODE == MODE_PRE) ? 8 : 6; // 2S (+2: residual, restriction)
  static constexpr int TX = LW - 2 * RUN_HX;
  static constexpr int TY = RUN_LH - 2 * HY;
  static constexpr size_t p_floats = 2 * (RUN_LH + 2) * RS; // one pad row above and below
  static constexpr size_t r_floats = (MODE == MODE_PRE) ? 2 * RUN_LH * RS : 0;
  static constexpr size_t smem = sizeof(float) * (p_floats + r_floats + 16) + 2 * RUN_LH * RS;
};

__device__ __forceinline__ float4 lds4(const float *p) { return *reinterpret_cast<const float4 *>(p); }

template <int MODE, int R>
__global__ void __launch_bounds__(16 * (RUN_LH / R), 2) k_mg_run(TileArgs a) {
  using G = RunGeom<MODE>;
  constexpr int S = G::S, LH = RUN_LH, NT = 16 * (RUN_LH / R), HX = RUN_HX, HY = G::HY;
  constexpr int TX = G::TX, TY = G::TY, NW = NT / 32;
  static_assert(R % 2 == 0 && HY % 2 == 0 && TX % 8 == 0 && TY % 2 == 0, "run geometry");

  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *sP = reinterpret_cast<float *>(smem_raw);
  float *sR = sP + G::p_floats;
  float *rcpt = sR + G::r_floats;
  float *prt = rcpt + 8;
  uint8_t *sM = reinterpret_cast<uint8_t *>(rcpt + 16);
  auto P = [&](int pl, int r) -> float * { return sP + (pl * (LH + 2) + r + 1) * RS; };
  auto Rr = [&](int pl, int r) -> float * { return sR + (pl * LH + r) * RS; };
  auto M = [&](int pl, int r) -> uint8_t * { return sM + (pl * LH + r) * RS; };

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tg = lane & 15, chunk = 2 * warp + (lane >> 4);
  const int r0 = R * chunk, ci = XO + 4 * tg;
  const unsigned hmask = 0xFFFFu << (lane & 16); // the 16 lanes that share this row chunk
  const int x0 = blockIdx.x * TX, y0 = a.own_lo + blockIdx.y * TY;
  const int X0 = x0 - HX, Y0 = y0 - HY; // X0 % 8 == 0, Y0 even: local parity == global parity
  const int w = a.w, h = a.h;
  const int gx8 = X0 + 8 * tg;

  if (threadIdx.x < 8) {
    rcpt[threadIdx.x] = rcp_count(threadIdx.x);
    prt[threadIdx.x] = prolong_rcp(threadIdx.x);
  }
  // programmatic dependent launch (common.cuh): this grid may have been scheduled while the previous
  // kernel of the stream was still draining; everything above touches no global memory
  ubgl_pdl_prologue();

  // MODE_POST: the coarse error and coarse stencil mask under this thread's cells -- coarse
  // rows ycb .. ycb+2, columns xcb .. xcb+4 -- for prolongation + correction
  // (pressure_solver.cpp:134-181), applied to p in registers on its way into shared memory
  float ecr[3][5];
  unsigned mcr[3];
  const int ycb = (Y0 + r0) >> 1;
  if (MODE == MODE_POST) {
    const int xcb = gx8 >> 1;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const int yc = ycb + j;
#pragma unroll
      for (int k = 0; k < 5; k++) ecr[j][k] = 0.0f;
      mcr[j] = 0u;
      if (yc >= a.c_lo && yc < a.c_hi && gx8 >= 0 && xcb < a.pc) {
        const size_t oc = (size_t)yc * a.pc + xcb;
        const float4 t = __ldg(reinterpret_cast<const float4 *>(a.ec + oc));
        ecr[j][0] = t.x; ecr[j][1] = t.y; ecr[j][2] = t.z; ecr[j][3] = t.w;
        if (xcb + 4 < a.pc) ecr[j][4] = __ldg(a.ec + oc + 4);
        mcr[j] = __ldg(reinterpret_cast<const unsigned *>(a.maskc + oc));
      }
    }
    __syncthreads(); // prt[] is read while staging
  }

  // ---- stage: p and the mask go to shared memory (colour planes), f*h*h stays in
  // registers; slot E = the thread's cells 0,2,4,6, slot O = cells 1,3,5,7 ----
  float4 FE[R], FO[R], WE[R], WO[R];
  {
    unsigned ME[R], MO[R];
#pragma unroll
    for (int i = 0; i < R; i++) {
      const int r = r0 + i, gy = Y0 + r;
      float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = p0, f0 = p0, f1 = p0;
      uint2 mv = make_uint2(0u, 0u);
      if (gy >= a.st_lo && gy < a.st_hi && gx8 >= 0 && gx8 < a.pitch) {
        const size_t o = (size_t)gy * a.pitch + gx8;
        if (a.p_in) {
          p0 = *reinterpret_cast<const float4 *>(a.p_in + o);
          p1 = *reinterpret_cast<const float4 *>(a.p_in + o + 4);
        }
        f0 = __ldg(reinterpret_cast<const float4 *>(a.f + o));
        f1 = __ldg(reinterpret_cast<const float4 *>(a.f + o + 4));
        mv = __ldg(reinterpret_cast<const uint2 *>(a.mask + o));
      }
      ME[i] = __byte_perm(mv.x, mv.y, 0x6420);
      MO[i] = __byte_perm(mv.x, mv.y, 0x7531);
      float4 pe = make_float4(sel0(ME[i], MB_C, p0.x), sel0(ME[i] >> 8, MB_C, p0.z),
                              sel0(ME[i] >> 16, MB_C, p1.x), sel0(ME[i] >> 24, MB_C, p1.z));
      float4 po = make_float4(sel0(MO[i], MB_C, p0.y), sel0(MO[i] >> 8, MB_C, p0.w),
                              sel0(MO[i] >> 16, MB_C, p1.y), sel0(MO[i] >> 24, MB_C, p1.w));
      if (MODE == MODE_POST) {
        // same per-cell rounding sequence as prolong_cell (stencils.cuh) + correct; the coarse
        // flag sums come from the coarse mask (C, E, N bits of (xc,yc), E bit of (xc,yc+1))
        const int j = i >> 1, yc = ycb + j, y = 2 * yc;
        const bool row_ok = yc >= a.c_lo && yc < a.c_hi;
        const bool y_o = y + 1 <= h - 3 && yc + 1 < a.c_hi;
        const bool act = row_ok && ((i & 1) ? y_o : (y >= 2 && y <= h - 2));
        float pev[4] = {pe.x, pe.y, pe.z, pe.w}, pov[4] = {po.x, po.y, po.z, po.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const int x = gx8 + 2 * k;
          const bool x_e = x >= 2 && x <= w - 2, x_o = x + 1 <= w - 3 && gx8 >= 0;
          const unsigned mc = mcr[j] >> (8 * k), mn = mcr[j + 1] >> (8 * k);
          const int fC = mc & 1, fE = (mc >> 5) & 1, fN = (mc >> 7) & 1, fNE = (mn >> 5) & 1;
          const float e00 = ecr[j][k], e10 = ecr[j][k + 1], e01 = ecr[j + 1][k], e11 = ecr[j + 1][k + 1];
          if ((i & 1) == 0) {
            if (act && x_e) pev[k] = __fadd_rn(pev[k], sel0(ME[i] >> (8 * k), MB_C, e00));
            if (act && x_o)
              pov[k] = __fadd_rn(pov[k], __fmul_rn(sel0(MO[i] >> (8 * k), MB_C, __fadd_rn(e00, e10)), prt[fC + fE]));
          } else {
            if (act && x_e)
              pev[k] = __fadd_rn(pev[k], __fmul_rn(sel0(ME[i] >> (8 * k), MB_C, __fadd_rn(e00, e01)), prt[fC + fN]));
            if (act && x_o) {
              const float es = __fadd_rn(__fadd_rn(__fadd_rn(e00, e11), e10), e01);
              pov[k] = __fadd_rn(pov[k], __fmul_rn(sel0(MO[i] >> (8 * k), MB_C, es), prt[fC + fNE + fE + fN]));
            }
          }
        }
        pe = make_float4(pev[0], pev[1], pev[2], pev[3]);
        po = make_float4(pov[0], pov[1], pov[2], pov[3]);
      }
      // cell j of row r has colour (j + r) & 1 = (j + i) & 1
      *reinterpret_cast<float4 *>(P(i & 1, r) + ci) = pe;
      *reinterpret_cast<float4 *>(P((i & 1) ^ 1, r) + ci) = po;
      *reinterpret_cast<unsigned *>(M(i & 1, r) + ci) = ME[i];
      *reinterpret_cast<unsigned *>(M((i & 1) ^ 1, r) + ci) = MO[i];
      if (MODE == MODE_PRE) { // raw f waits in the (still unused) residual planes for the residual pass
        *reinterpret_cast<float4 *>(Rr(i & 1, r) + ci) = make_float4(f0.x, f0.z, f1.x, f1.z);
        *reinterpret_cast<float4 *>(Rr((i & 1) ^ 1, r) + ci) = make_float4(f0.y, f0.w, f1.y, f1.w);
      }
      FE[i] = make_float4(fh2_of(f0.x, a.hh), fh2_of(f0.z, a.hh), fh2_of(f1.x, a.hh), fh2_of(f1.z, a.hh));
      FO[i] = make_float4(fh2_of(f0.y, a.hh), fh2_of(f0.w, a.hh), fh2_of(f1.y, a.hh), fh2_of(f1.w, a.hh));
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < R; i++) {
      WE[i] = make_float4(rcpt[(ME[i] >> 2) & 7u], rcpt[(ME[i] >> 10) & 7u], rcpt[(ME[i] >> 18) & 7u],
                          rcpt[(ME[i] >> 26) & 7u]);
      WO[i] = make_float4(rcpt[(MO[i] >> 2) & 7u], rcpt[(MO[i] >> 10) & 7u], rcpt[(MO[i] >> 18) & 7u],
                          rcpt[(MO[i] >> 26) & 7u]);
    }
  }

  // L2 prefetch of the window the CTA that follows this one on the SM will stage (about one
  // generation of resident CTAs ahead in launch order): with 2 CTAs x 8 warps per SM the
  // staging loads are the largest stall of the kernel (ncu: long_sb 28 %), an L2 hit
  // instead of an HBM miss shortens it.  No extra DRAM traffic: each line is still
  // fetched once and used within ~one CTA lifetime (296 windows x 72 KB << 126 MB L2).
  if (a.pf_dist > 0) {
    const long long lb = (long long)blockIdx.y * gridDim.x + blockIdx.x + a.pf_dist;
    if (lb < (long long)gridDim.x * gridDim.y) {
      const int by = (int)(lb / gridDim.x), bx = (int)(lb - (long long)by * gridDim.x);
      const int pgx = bx * TX - HX + 8 * tg, pY0 = a.own_lo + by * TY - HY;
      if (pgx >= 0 && pgx < a.pitch) {
#pragma unroll
        for (int i = 0; i < R; i++) {
          const int gy = pY0 + r0 + i;
          if (gy < a.st_lo || gy >= a.st_hi) continue;
          const size_t o = (size_t)gy * a.pitch + pgx;
          if (a.p_in) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.p_in + o));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(a.f + o));
          if ((tg & 3) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.mask + o));
        }
      }
    }
  }

  // Per-thread 4-bit masks over its 4 cells of slot q (bits 4q .. 4q+3):
  // fz: cells on the global W/E border column (never updated); in: 1 <= gx <= w-2.
  unsigned fzb = 0, inb = 0;
#pragma unroll
  for (int q = 0; q < 2; q++) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int gx = gx8 + 2 * j + q;
      if (gx == 0 || gx == w - 1) fzb |= 1u << (4 * q + j);
      if (gx >= 1 && gx <= w - 2) inb |= 1u << (4 * q + j);
    }
  }

  auto lo_ = [](int origin, int k) { return max(1, origin + k); };
  auto hi_ = [](int origin, int len, int n, int k) { return min(n - 1, origin + len - k); };

  // W / E neighbours of the 4 cells of slot q, given the other colour's 4 cells A of
  // the same row: q == 0: W = (left lane's A.w, A.x, A.y, A.z), E = A;
  //               q == 1: W = A, E = (A.y, A.z, A.w, right lane's A.x).
  // The window edge reads 0 (the cell beyond it is outside the valid trapezoid anyway).
  auto edge = [&](const float4 &A, int q) {
    float e = q == 0 ? __shfl_up_sync(hmask, A.w, 1, 16) : __shfl_down_sync(hmask, A.x, 1, 16);
    return (tg == (q == 0 ? 0 : 15)) ? 0.0f : e;
  };

  // Most threads (row chunks 2 .. NCH-3 of an interior window, no cell on a global W / E border
  // column) have all four rows inside the exact region of every half-sweep: for them the
  // half-sweep runs without per-row tests and with packed fp32 (two cells per FADD2 / FMUL2;
  // lane for lane the additions and the multiply of upd() below, in its order).
  // (uniform over the 16 lanes that share a row chunk and shuffle with each other)
  const bool no_fz = (__ballot_sync(0xffffffffu, fzb != 0) & hmask) == 0;
  const bool fast = no_fz && r0 >= 2 * S && r0 >= 1 - Y0 && r0 + R <= LH - 2 * S && r0 + R <= h - 1 - Y0;

  // one colour of one sweep over the rows of this thread's run that are still exact at time k
  auto half_sweep = [&](int k, auto cpar_tag) {
    constexpr int CPAR = decltype(cpar_tag)::value;
    const float *po = P(CPAR ^ 1, r0 - 1) + ci;
    float *pd = P(CPAR, r0) + ci;
    if (fast) {
      float4 O[R + 2];
#pragma unroll
      for (int j = 0; j < R + 2; j++) O[j] = lds4(po + j * RS);
      float ed[R];
#pragma unroll
      for (int i = 0; i < R; i++)
        ed[i] = ((CPAR + i) & 1) == 0 ? __shfl_up_sync(hmask, O[i + 1].w, 1, 16) : __shfl_down_sync(hmask, O[i + 1].x, 1, 16);
#pragma unroll
      for (int i = 0; i < R; i++) {
        const int q = (CPAR + i) & 1;
        const float4 Sv = O[i], A = O[i + 1], Nv = O[i + 2];
        const float4 Fv = q == 0 ? FE[i] : FO[i];
        const float4 Wt = q == 0 ? WE[i] : WO[i];
        // pw + pe of the four cells: q == 0: (ed, A.x, A.y, A.z) + A;  q == 1: A + (A.y, A.z, A.w, ed)
        f2 a = q == 0 ? add2(pk(ed[i], A.x), pk(A.x, A.y)) : add2(pk(A.x, A.y), pk(A.y, A.z));
        f2 b = q == 0 ? add2(pk(A.y, A.z), pk(A.z, A.w)) : add2(pk(A.z, A.w), pk(A.w, ed[i]));
        a = add2(add2(a, pk(Sv.x, Sv.y)), pk(Nv.x, Nv.y));
        b = add2(add2(b, pk(Sv.z, Sv.w)), pk(Nv.z, Nv.w));
        a = mul2(add2(a, pk(Fv.x, Fv.y)), pk(Wt.x, Wt.y));
        b = mul2(add2(b, pk(Fv.z, Fv.w)), pk(Wt.z, Wt.w));
        *reinterpret_cast<float4 *>(pd + i * RS) = make_float4(lo(a), hi(a), lo(b), hi(b));
      }
      return;
    }
    const int r_lo = lo_(Y0, k) - Y0, r_hi = hi_(Y0, LH, h, k) - Y0;
    if (r0 + R <= r_lo || r0 >= r_hi) return;
    // all R+2 rows of the other colour and the R lane-edge values first (independent loads
    // and shuffles in flight together), then the arithmetic
    float4 O[R + 2];
#pragma unroll
    for (int j = 0; j < R + 2; j++) O[j] = lds4(po + j * RS);
    float ed[R];
#pragma unroll
    for (int i = 0; i < R; i++) ed[i] = edge(O[i + 1], (CPAR + i) & 1);
#pragma unroll
    for (int i = 0; i < R; i++) {
      const int q = (CPAR + i) & 1;
      const float4 Sv = O[i], A = O[i + 1], Nv = O[i + 2];
      if (r0 + i >= r_lo && r0 + i < r_hi) {
        const float4 Wv = q == 0 ? make_float4(ed[i], A.x, A.y, A.z) : A;
        const float4 Ev = q == 0 ? A : make_float4(A.y, A.z, A.w, ed[i]);
        const float4 Fv = q == 0 ? FE[i] : FO[i];
        const float4 Wt = q == 0 ? WE[i] : WO[i];
        auto upd = [&](float pw, float pe, float ps, float pn, float f, float wt) {
          float v = __fadd_rn(__fadd_rn(__fadd_rn(pw, pe), ps), pn);
          return __fmul_rn(__fadd_rn(v, f), wt);
        };
        float4 v;
        v.x = upd(Wv.x, Ev.x, Sv.x, Nv.x, Fv.x, Wt.x);
        v.y = upd(Wv.y, Ev.y, Sv.y, Nv.y, Fv.y, Wt.y);
        v.z = upd(Wv.z, Ev.z, Sv.z, Nv.z, Fv.z, Wt.z);
        v.w = upd(Wv.w, Ev.w, Sv.w, Nv.w, Fv.w, Wt.w);
        float *d = pd + i * RS;
        const unsigned z = (fzb >> (4 * q)) & 15u;
        if (z == 0) {
          *reinterpret_cast<float4 *>(d) = v;
        } else {
          if (!(z & 1)) d[0] = v.x;
          if (!(z & 2)) d[1] = v.y;
          if (!(z & 4)) d[2] = v.z;
          if (!(z & 8)) d[3] = v.w;
        }
      }
    }
  };

  auto cell = [&](int gx, int gy) -> float & {
    const int lx = gx - X0, ly = gy - Y0;
    return P((lx + ly) & 1, ly)[XO + (lx >> 1)];
  };
  auto mbyte = [&](int gx, int gy) -> unsigned {
    const int lx = gx - X0, ly = gy - Y0;
    return M((lx + ly) & 1, ly)[XO + (lx >> 1)];
  };

  // setZeroGradientBC on the border cells whose interior neighbour is still exact at
  // time k (see k_mg_tile)
  auto zero_gradient = [&](int k) {
    const int gx_lo = lo_(X0, k), gx_hi = hi_(X0, LW, w, k);
    const int gy_lo = lo_(Y0, k), gy_hi = hi_(Y0, LH, h, k);
    const int t = threadIdx.x;
    if (X0 <= 0 && gx_lo == 1)
      for (int gy = gy_lo + t; gy < gy_hi; gy += NT) cell(0, gy) = sel0(mbyte(0, gy), MB_C, cell(1, gy));
    if (w - 1 < X0 + LW && gx_hi == w - 1)
      for (int gy = gy_lo + t; gy < gy_hi; gy += NT)
        cell(w - 1, gy) = sel0(mbyte(w - 1, gy), MB_C, cell(w - 2, gy));
    if (Y0 <= 0 && gy_lo == 1)
      for (int gx = gx_lo + t; gx < gx_hi; gx += NT) cell(gx, 0) = sel0(mbyte(gx, 0), MB_C, cell(gx, 1));
    if (h - 1 < Y0 + LH && gy_hi == h - 1)
      for (int gx = gx_lo + t; gx < gx_hi; gx += NT)
        cell(gx, h - 1) = sel0(mbyte(gx, h - 1), MB_C, cell(gx, h - 2));
  };

  // Synchronisation between half-sweeps.  A half-sweep reads, of other warps' cells, only the
  // row directly below its lower row chunk and the row directly above its upper one: rows of
  // warp - 1 and warp + 1.  So a warp does not wait for the whole CTA but meets its two
  // neighbours at NAMED barriers (id = upper warp of the pair, 64 threads each, lower pair
  // first: no cycle).  Warps of a CTA may then be a half-sweep apart, and one warp's
  // shared-memory latency overlaps another's arithmetic -- with a block barrier all eight start
  // every phase together and the 4 warps per scheduler stall together (ncu: no pipe above
  // 65 %, barrier + short-scoreboard stalls).  Windows that hold border cells of a level-0
  // zero-gradient solve run zero_gradient() between sweeps, a block-wide pattern: they keep the
  // block barriers; every other window skips them.
  const bool zg_here = a.zgbc && (X0 <= 0 || w - 1 < X0 + LW || Y0 <= 0 || h - 1 < Y0 + LH);
  const bool g_pair_sync_off = a.pair_sync == 0;
  auto pair_sync = [&]() {
    if (g_pair_sync_off) {
      __syncthreads();
      return;
    }
    if (warp > 0) asm volatile("bar.sync %0, 64;" ::"r"(warp) : "memory");
    if (warp < NW - 1) asm volatile("bar.sync %0, 64;" ::"r"(warp + 1) : "memory");
  };

  if (MODE == MODE_POST && zg_here) { // setZeroGradientBC after correct (pressure_solver.cpp:236-239)
    zero_gradient(0);
    __syncthreads();
  }

#pragma unroll 1
  for (int s = 0; s < ((a.dbg & 1) ? 0 : S); s++) {
    half_sweep(2 * s + 1, std::integral_constant<int, 1>()); // "red":   (x+y) odd,  pressure_solver.cpp:35-40
    if (zg_here) __syncthreads(); else pair_sync();
    half_sweep(2 * s + 2, std::integral_constant<int, 0>()); // "black": (x+y) even, pressure_solver.cpp:42-47
    if (zg_here) {
      __syncthreads();
      zero_gradient(2 * s + 2);
      __syncthreads();
    } else {
      pair_sync();
    }
  }

  if (MODE == MODE_PRE && !(a.dbg & 2)) {
    // residual (pressure_solver.cpp:101-111, binary flags) of the thread's own cells on
    // rows [HY-1, HY+TY+1), both colours, written over the raw f parked in the R planes
    const int rr_lo = HY - 1, rr_hi = HY + TY + 1;
    if (r0 + R > rr_lo && r0 < rr_hi) {
#pragma unroll
      for (int cpar = 0; cpar < 2; cpar++) {
        const float *po = P(cpar ^ 1, r0 - 1) + ci;
        float4 O[R + 2];
#pragma unroll
        for (int j = 0; j < R + 2; j++) O[j] = lds4(po + j * RS);
        float edv[R];
#pragma unroll
        for (int i = 0; i < R; i++) edv[i] = edge(O[i + 1], (cpar + i) & 1);
#pragma unroll
        for (int i = 0; i < R; i++) {
          const float4 Sv = O[i], A = O[i + 1], Nv = O[i + 2];
          const int q = (cpar + i) & 1;
          const float ed = edv[i];
          const int r = r0 + i, gy = Y0 + r;
          if (r >= rr_lo && r < rr_hi) {
            const float4 Wv = q == 0 ? make_float4(ed, A.x, A.y, A.z) : A;
            const float4 Ev = q == 0 ? A : make_float4(A.y, A.z, A.w, ed);
            const float4 Cv = lds4(P(cpar, r) + ci);
            const unsigned mw = *reinterpret_cast<const unsigned *>(M(cpar, r) + ci);
            const float4 Fv = lds4(Rr(cpar, r) + ci); // raw f of these cells, parked at staging
            const bool rowin = gy >= 1 && gy <= h - 2;
            const unsigned in = rowin ? (inb >> (4 * q)) & 15u : 0u;
            auto res = [&](float pc_, float pw, float pe, float ps, float pn, float f, int j) {
              const unsigned m = mw >> (8 * j);
              float val = (m & MB_W) ? pw : pc_;
              val = __fadd_rn(val, (m & MB_E) ? pe : pc_);
              val = __fadd_rn(val, (m & MB_S) ? ps : pc_);
              val = __fadd_rn(val, (m & MB_N) ? pn : pc_);
              val = __fmaf_rn(-4.0f, pc_, val);
              val = __fmul_rn(val, a.ihsq);
              val = __fadd_rn(f, val);
              return ((m & MB_C) && ((in >> j) & 1u)) ? val : 0.0f;
            };
            float4 v;
            v.x = res(Cv.x, Wv.x, Ev.x, Sv.x, Nv.x, Fv.x, 0);
            v.y = res(Cv.y, Wv.y, Ev.y, Sv.y, Nv.y, Fv.y, 1);
            v.z = res(Cv.z, Wv.z, Ev.z, Sv.z, Nv.z, Fv.z, 2);
            v.w = res(Cv.w, Wv.w, Ev.w, Sv.w, Nv.w, Fv.w, 3);
            *reinterpret_cast<float4 *>(Rr(cpar, r) + ci) = v;
          }
        }
      }
    }
    __syncthreads();
    // full-weighting restriction of the coarse cells whose fine centre (2xc,2yc) lies in
    // this tile; coarse border = 0 (rc.fill(0.0), pressure_solver.cpp:222)
    const int xc0 = x0 >> 1, yc0 = y0 >> 1;
    for (int j = warp; j < TY / 2; j += NW) {
      const int yc = yc0 + j;
      if (yc >= a.hc || 2 * yc >= a.own_hi) break;
      const int ly = 2 * yc - Y0;
      const float *F0m = Rr(0, ly - 1), *F1m = Rr(1, ly - 1), *F0c = Rr(0, ly), *F1c = Rr(1, ly),
                  *F0p = Rr(0, ly + 1), *F1p = Rr(1, ly + 1);
      for (int i = lane; i < TX / 2; i += 32) {
        const int xc = xc0 + i;
        if (xc >= a.wc) break;
        float v = 0.0f;
        if (xc >= 1 && yc >= 1 && xc < a.wc - 1 && yc < a.hc - 1) {
          const int c = XO + ((2 * xc - X0) >> 1); // column of the (even,even) centre
          v = fw9(F0m[c - 1], F1m[c], F0m[c], F1c[c - 1], F0c[c], F1c[c], F0p[c - 1], F1p[c], F0p[c]);
        }
        a.rc[(size_t)yc * a.pc + xc] = v;
      }
    }
  }

  // ---- write the thread's own tile cells back (p ping-pong buffer), 128-bit stores ----
  if (tg >= HX / 8 && tg < (HX + TX) / 8) {
#pragma unroll
    for (int i = 0; i < R; i++) {
      const int r = r0 + i, gy = Y0 + r;
      if (r < HY || r >= HY + TY || gy >= a.own_hi || gx8 >= w) continue;
      const float4 e = lds4(P(i & 1, r) + ci), o = lds4(P((i & 1) ^ 1, r) + ci);
      float *dst = a.p_out + (size_t)gy * a.pitch + gx8;
      *reinterpret_cast<float4 *>(dst) = make_float4(e.x, o.x, e.y, o.y);
      if (gx8 + 4 < w) *reinterpret_cast<float4 *>(dst + 4) = make_float4(e.z, o.z, e.w, o.w);
    }
  }

  // ---- border cells of this tile (see k_mg_tile) ----
  const bool bx0 = x0 == 0, bx1 = (w - 1 >= x0 && w - 1 < x0 + TX);
  const bool by0 = y0 == 0, by1 = (h - 1 >= y0 && h - 1 < y0 + TY && h - 1 < a.own_hi);
  if (bx0 || bx1 || by0 || by1) {
    __syncthreads();
    const int t = threadIdx.x;
    auto orig = [&](int gx, int gy) { return a.p_in ? a.p_in[(size_t)gy * a.pitch + gx] : 0.0f; };
    auto put = [&](int gx, int gy, int nx, int ny) {
      const bool corner = (gx == 0 || gx == w - 1) && (gy == 0 || gy == h - 1);
      a.p_out[(size_t)gy * a.pitch + gx] = (a.zgbc && !corner) ? cell(nx, ny) : orig(gx, gy);
    };
    const int ty_hi = min(a.own_hi, y0 + TY), tx_hi = min(w, x0 + TX);
    if (bx0)
      for (int gy = y0 + t; gy < ty_hi; gy += NT) put(0, gy, 1, gy);
    if (bx1)
      for (int gy = y0 + t; gy < ty_hi; gy += NT) put(w - 1, gy, w - 2, gy);
    if (by0)
      for (int gx = x0 + t; gx < tx_hi; gx += NT) put(gx, 0, gx, 1);
    if (by1)
      for (int gx = x0 + t; gx < tx_hi; gx += NT) put(gx, h - 1, gx, h - 2);
  }
}

// ---------------------------------------------------------------------------
// k_mg_tail -- the bottom of the V-cycle in ONE launch.  Below ~128^2 cells a level
// is a single window and every PRE / POST launch costs its ~12 us latency chain
// (staging, 6 barriers-separated half-sweeps, write-back) whatever the size: at
// 8192^2 levels 6-10 take 117 us of a 1.66 ms V-cycle, on the 1090x436 game level
// a third of the step.  Here one CTA keeps levels t..L (L = levels-2, the coarsest
// one solveLevel uses) entirely in shared memory -- p, the rhs and the stencil mask
// of each level, 9 B/cell -- and runs solveLevel(t) (pressure_solver.cpp:201-248)
// from the zero initial guess: 3 sweeps, residual + full weighting (residuals are
// evaluated inside the restriction stencil, never stored), recursion, prolongation
// + correction, 3 sweeps; 5 sweeps on level L.  The only HBM traffic is the rhs of
// level t in, its solution out.  Per-cell arithmetic is the tile kernels' (solid
// cells stored as 0, table weights, same order of additions): bit-identical.
// ---------------------------------------------------------------------------
constexpr int TAIL_MAXLV = 8, TAIL_NT = 1024;
struct TailArgs {
  int n;                       // levels t .. t+n-1
  int w[TAIL_MAXLV], h[TAIL_MAXLV], gp[TAIL_MAXLV]; // size and GLOBAL pitch of each level
  int sp[TAIL_MAXLV];          // shared-memory row pitch: w rounded up to 4 cells
  int off[TAIL_MAXLV];         // first cell of the level in the shared arrays
  int cells;                   // total (padded) cells of all tail levels
  float hh[TAIL_MAXLV];
  const uint8_t *mask[TAIL_MAXLV];
  const float *rhs;            // rc of level t
  float *out;                  // ec of level t
};

__global__ void __launch_bounds__(TAIL_NT, 1) k_mg_tail(TailArgs a) {
  ubgl_pdl_prologue();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *sPa = reinterpret_cast<float *>(smem_raw);
  float *sFa = sPa + a.cells;
  uint8_t *sMa = reinterpret_cast<uint8_t *>(sFa + a.cells);
  __shared__ float rcpt[8], prt[8];
  const int tid = threadIdx.x;
  constexpr int NT = TAIL_NT, UB = 4; // staging keeps UB rows per warp in flight
  if (tid < 8) {
    rcpt[tid] = rcp_count(tid);
    prt[tid] = prolong_rcp(tid);
  }
  // ---- stage: masks of every level, rhs of level t, p = 0 everywhere; 4 cells per access
  // (shared rows are padded to a multiple of 4 cells, global rows to 32) ----
  const int lane = tid & 31, warp = tid >> 5;
  constexpr int NW = NT / 32;
  for (int l = 0; l < a.n; l++) {
    const int sp = a.sp[l], wq = sp >> 2, h = a.h[l];
    for (int y0 = warp; y0 < h; y0 += UB * NW)
      for (int xq = lane; xq < wq; xq += 32) {
        unsigned m[UB];
        float4 f[UB];
#pragma unroll
        for (int k = 0; k < UB; k++) {
          const int y = y0 + k * NW;
          m[k] = 0u;
          f[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (y < h) {
            m[k] = __ldg(reinterpret_cast<const unsigned *>(a.mask[l] + (size_t)y * a.gp[l] + 4 * xq));
            if (l == 0) f[k] = __ldg(reinterpret_cast<const float4 *>(a.rhs + (size_t)y * a.gp[0] + 4 * xq));
          }
        }
#pragma unroll
        for (int k = 0; k < UB; k++) {
          const int y = y0 + k * NW;
          if (y < h) {
            const int i = a.off[l] + y * sp + 4 * xq;
            *reinterpret_cast<unsigned *>(sMa + i) = m[k];
            *reinterpret_cast<float4 *>(sFa + i) = f[k];
            *reinterpret_cast<float4 *>(sPa + i) = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
  }
  __syncthreads();

  // `count` red-black sweeps of level l: a warp per row, lanes over the row's cells of one colour
  auto sweeps = [&](int l, int count) {
    const int w = a.w[l], h = a.h[l], sp = a.sp[l];
    float *P = sPa + a.off[l];
    const float *F = sFa + a.off[l];
    const uint8_t *M = sMa + a.off[l];
    const float hh = a.hh[l];
    for (int s = 0; s < 2 * count; s++) {
      const int cpar = (s & 1) ^ 1; // red (x+y odd) first, pressure_solver.cpp:35-47
      for (int y = 1 + warp; y < h - 1; y += NW)
        for (int x = 1 + ((y + cpar + 1) & 1) + 2 * lane; x < w - 1; x += 64) {
          const int i = y * sp + x;
          float v = __fadd_rn(__fadd_rn(__fadd_rn(P[i - 1], P[i + 1]), P[i - sp]), P[i + sp]);
          v = __fadd_rn(v, fh2_of(F[i], hh));
          P[i] = __fmul_rn(v, rcpt[(M[i] >> 2) & 7u]);
        }
      __syncthreads();
    }
  };
  // residual_cell with binary flags (see k_mg_tile); 0 outside the interior
  auto residual = [&](const float *P, const float *F, const uint8_t *M, int w, int h, int sp, int x,
                      int y, float ihsq) {
    if (x < 1 || y < 1 || x > w - 2 || y > h - 2) return 0.0f;
    const int i = y * sp + x;
    const unsigned m = M[i];
    const float pc = P[i];
    float val = (m & MB_W) ? P[i - 1] : pc;
    val = __fadd_rn(val, (m & MB_E) ? P[i + 1] : pc);
    val = __fadd_rn(val, (m & MB_S) ? P[i - sp] : pc);
    val = __fadd_rn(val, (m & MB_N) ? P[i + sp] : pc);
    val = __fmaf_rn(-4.0f, pc, val);
    val = __fmul_rn(val, ihsq);
    val = __fadd_rn(F[i], val);
    return (m & MB_C) ? val : 0.0f;
  };

  // ---- down ----
  for (int l = 0; l + 1 < a.n; l++) {
    sweeps(l, 3);
    const int w = a.w[l], h = a.h[l], sp = a.sp[l], wc = a.w[l + 1], hc = a.h[l + 1], spc = a.sp[l + 1];
    const float *P = sPa + a.off[l], *F = sFa + a.off[l];
    const uint8_t *M = sMa + a.off[l];
    float *Fc = sFa + a.off[l + 1];
    const float ihsq = 1.0f / a.hh[l] / a.hh[l];
    for (int yc = 1 + warp; yc < hc - 1; yc += NW)
      for (int xc = 1 + lane; xc < wc - 1; xc += 32) {
        const int x = 2 * xc, y = 2 * yc;
        auto r = [&](int dx, int dy) { return residual(P, F, M, w, h, sp, x + dx, y + dy, ihsq); };
        Fc[yc * spc + xc] = fw9(r(-1, -1), r(0, -1), r(1, -1), r(-1, 0), r(0, 0), r(1, 0), r(-1, 1), r(0, 1), r(1, 1));
      }
    __syncthreads();
  }
  sweeps(a.n - 1, 5); // level L: 5 sweeps, pressure_solver.cpp:203-206
  // ---- up ----
  for (int l = a.n - 2; l >= 0; l--) {
    const int w = a.w[l], h = a.h[l], sp = a.sp[l], wc = a.w[l + 1], hc = a.h[l + 1], spc = a.sp[l + 1];
    float *P = sPa + a.off[l];
    const uint8_t *M = sMa + a.off[l];
    const float *E = sPa + a.off[l + 1];
    const uint8_t *Mc = sMa + a.off[l + 1];
    // prolongate + correct, one thread per coarse cell (as in k_mg_tile)
    for (int yc = warp; yc < hc; yc += NW)
    for (int xc = lane; xc < wc; xc += 32) {
      const int x = 2 * xc, y = 2 * yc;
      const bool y_e = y >= 2 && y <= h - 2, y_o = y + 1 <= h - 3 && yc + 1 < hc;
      const bool x_e = x >= 2 && x <= w - 2, x_o = x + 1 <= w - 3;
      const int ic = yc * spc + xc;
      const float e00 = E[ic];
      const float e10 = x_o ? E[ic + 1] : 0.0f;
      const float e01 = y_o ? E[ic + spc] : 0.0f;
      const float e11 = (x_o && y_o) ? E[ic + spc + 1] : 0.0f;
      const unsigned mc = Mc[ic];
      const unsigned mn = y_o ? Mc[ic + spc] : 0u;
      const int fC = mc & 1, fE = (mc >> 5) & 1, fN = (mc >> 7) & 1, fNE = (mn >> 5) & 1;
      const int i = y * sp + x;
      if (y_e) {
        if (x_e) P[i] = __fadd_rn(P[i], sel0(M[i], MB_C, e00));
        if (x_o)
          P[i + 1] = __fadd_rn(P[i + 1], __fmul_rn(sel0(M[i + 1], MB_C, __fadd_rn(e00, e10)), prt[fC + fE]));
      }
      if (y_o) {
        if (x_e)
          P[i + sp] = __fadd_rn(P[i + sp], __fmul_rn(sel0(M[i + sp], MB_C, __fadd_rn(e00, e01)), prt[fC + fN]));
        if (x_o) {
          const float es = __fadd_rn(__fadd_rn(__fadd_rn(e00, e11), e10), e01);
          P[i + sp + 1] = __fadd_rn(P[i + sp + 1], __fmul_rn(sel0(M[i + sp + 1], MB_C, es), prt[fC + fNE + fE + fN]));
        }
      }
    }
    __syncthreads();
    sweeps(l, 3);
  }
  // ---- solution of level t (whole padded rows: the pad columns hold 0) ----
  {
    const int sp = a.sp[0], wq = sp >> 2, h = a.h[0];
    for (int y = warp; y < h; y += NW)
      for (int xq = lane; xq < wq; xq += 32)
        *reinterpret_cast<float4 *>(a.out + (size_t)y * a.gp[0] + 4 * xq) =
            *reinterpret_cast<const float4 *>(sPa + y * sp + 4 * xq);
  }
}

// Stencil mask of a flag grid (layout: enum MB_* above; neighbours outside the
// grid count as solid).  *nonbinary is raised if any flag is neither 0.0 nor 1.0
// (then the bit form is not equivalent and the plain path is used).
// Rows [r_lo, r_hi) are written; flag rows outside [st_lo, st_hi) are not stored
// on this GPU and read as solid (the slab code refreshes those mask rows from
// their owner afterwards).
__global__ void k_make_mask(Grid flag, uint8_t *mask, int *nonbinary, int r_lo, int r_hi, int st_lo,
                            int st_hi) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = r_lo + blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= flag.pitch || y >= r_hi) return;
  unsigned m = 0;
  if (x < flag.w) {
    auto bit = [&](int xx, int yy) -> unsigned {
      if (xx < 0 || yy < st_lo || xx >= flag.w || yy >= st_hi) return 0u;
      return flag.at(xx, yy) != 0.0f ? 1u : 0u;
    };
    float c = flag.at(x, y);
    if (c != 0.0f && c != 1.0f) *nonbinary = 1;
    const unsigned bc = bit(x, y), bw = bit(x - 1, y), be = bit(x + 1, y), bs = bit(x, y - 1),
                   bn = bit(x, y + 1);
    const unsigned code = bc ? (bw + be + bs + bn) : 0u;
    m = bc * MB_C | bw * MB_W | (code << 2) | be * MB_E | bs * MB_S | bn * MB_N;
  }
  mask[(size_t)y * flag.pitch + x] = (uint8_t)m;
}

// k_make_mask for the cells around edited discs (see disc_rect_cell in mg.cu: same rectangle)
__global__ void k_make_mask_discs(Grid flag, uint8_t *mask, const float *xyd, int level) {
  ubgl_pdl_prologue();
  const int c = blockIdx.z;
  const int d = (int)xyd[3 * c + 2];
  const int lox = (int)xyd[3 * c] - d - 1, loy = (int)xyd[3 * c + 1] - d - 1, side = ((2 * d + 3) >> level) + 7;
  const int tx = blockIdx.x * blockDim.x + threadIdx.x, ty = blockIdx.y * blockDim.y + threadIdx.y;
  const int x = (lox >> level) - 2 + tx, y = (loy >> level) - 2 + ty;
  if (tx >= side || ty >= side || x < 0 || y < 0 || x >= flag.w || y >= flag.h) return;
  auto bit = [&](int xx, int yy) -> unsigned {
    if (xx < 0 || yy < 0 || xx >= flag.w || yy >= flag.h) return 0u;
    return flag.at(xx, yy) != 0.0f ? 1u : 0u;
  };
  const unsigned bc = bit(x, y), bw = bit(x - 1, y), be = bit(x + 1, y), bs = bit(x, y - 1), bn = bit(x, y + 1);
  const unsigned code = bc ? (bw + be + bs + bn) : 0u;
  mask[(size_t)y * flag.pitch + x] = (uint8_t)(bc * MB_C | bw * MB_W | (code << 2) | be * MB_E | bs * MB_S | bn * MB_N);
}
void launch_make_mask_discs(const Grid &flag, uint8_t *mask, const float *d_xyd, dim3 grid, int level,
                            cudaStream_t stream, LaunchCounter *lc) {
  UBGL_LAUNCH(lc, K_COARSEN, level, stream, launch_k(k_make_mask_discs, grid, dim3(32, 8), 0, stream, flag, mask, d_xyd, level));
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
template <int S, int MODE, int LH, int NT>
static void launch_tile(const TileArgs &a, cudaStream_t stream, LaunchCounter *lc, int kind,
                        int level) {
  constexpr int HALO = TileGeom<S, MODE>::HALO;
  constexpr int TX = TileGeom<S, MODE>::TX;
  constexpr int TY = LH - 2 * HALO;
  constexpr size_t smem = sizeof(TileSmem<LH>);
  static std::atomic<unsigned long long> attr_done{0};
  ensure_dyn_smem(k_mg_tile<S, MODE, LH, NT>, smem, attr_done);
  dim3 grid(ceil_div(a.w, TX), ceil_div(a.own_hi - a.own_lo, TY));
  UBGL_LAUNCH(lc, kind, level, stream, k_mg_tile<S, MODE, LH, NT><<<grid, NT, smem, stream>>>(a));
}

constexpr int LH_MAIN = 80, NT_MAIN = 256;

// 2 (default): k_mg_run (register runs); 1: k_mg_tile (shared-memory tiles).  Process
// wide; set through UBGL_OPT_FUSED so the tests can compare the two schedules.
static int env_tile_variant() {
  const char *e = getenv("UBGL_TILE_VARIANT"); // A/B runs of bench.py; tests use UBGL_OPT_FUSED
  return (e && e[0] == '1') ? 1 : 2;
}
static int g_tile_variant = env_tile_variant();
static int env_prefetch_dist() {
  const char *e = getenv("UBGL_MG_PREFETCH"); // CTAs ahead; default one generation of 148 SMs x 2
  return e ? atoi(e) : 296;
}
static int g_prefetch_dist = env_prefetch_dist();
static size_t env_tail_cells() {
  const char *e = getenv("UBGL_MG_TAIL_CELLS");
  return e ? (size_t)atoll(e) : 8192;
}
static size_t g_tail_max_cells = env_tail_cells();
void set_tile_variant(int v) { g_tile_variant = (v == 1) ? 1 : 2; }
int tile_variant() { return g_tile_variant; }

template <int MODE>
static void launch_run(const TileArgs &a, cudaStream_t stream, LaunchCounter *lc, int kind, int level) {
  using G = RunGeom<MODE>;
  dim3 grid(ceil_div(a.w, G::TX), ceil_div(a.own_hi - a.own_lo, G::TY));
  TileArgs b = a;
  b.pf_dist = g_prefetch_dist;
  static const int pair = [] {
    const char *e = getenv("UBGL_MG_PAIR_SYNC");
    return (e && e[0] == '0') ? 0 : 1;
  }();
  b.pair_sync = pair;
  static const int dbg = [] {
    const char *e = getenv("UBGL_MG_DBG");
    return e ? atoi(e) : 0;
  }();
  b.dbg = dbg;
  static const int rows_per_thread = [] { // UBGL_MG_ROWS: 4 or 2 (see RUN_LH)
    const char *e = getenv("UBGL_MG_ROWS");
    return (e && e[0] == '4') ? 4 : ((e && e[0] == '2') ? 2 : UBGL_MG_ROWS_DEFAULT);
  }();
  if (rows_per_thread == 2) {
    static std::atomic<unsigned long long> attr_done2{0};
    ensure_dyn_smem(k_mg_run<MODE, 2>, G::smem, attr_done2);
    UBGL_LAUNCH(lc, kind, level, stream, (launch_k(k_mg_run<MODE, 2>, grid, 16 * (RUN_LH / 2), G::smem, stream, b)));
  } else {
    static std::atomic<unsigned long long> attr_done4{0};
    ensure_dyn_smem(k_mg_run<MODE, 4>, G::smem, attr_done4);
    UBGL_LAUNCH(lc, kind, level, stream, (launch_k(k_mg_run<MODE, 4>, grid, 16 * (RUN_LH / 4), G::smem, stream, b)));
  }
}

static void set_rows(TileArgs &a, const Rows *rows) {
  if (rows) {
    a.st_lo = rows->st_lo; a.st_hi = rows->st_hi; a.own_lo = rows->own_lo; a.own_hi = rows->own_hi;
  } else {
    a.st_lo = 0; a.st_hi = a.h; a.own_lo = 0; a.own_hi = a.h;
  }
  a.c_lo = 0; a.c_hi = a.hc;
}

void launch_mg_pre(const float *p_in, float *p_out, const Grid &f, const uint8_t *mask,
                   const Grid &rc, float hh, bool zgbc, cudaStream_t stream, LaunchCounter *lc,
                   int level, const Rows *rows) {
  TileArgs a{};
  a.p_in = p_in; a.p_out = p_out; a.f = f.d; a.mask = mask;
  a.w = f.w; a.h = f.h; a.pitch = f.pitch;
  a.rc = rc.d; a.wc = rc.w; a.hc = rc.h; a.pc = rc.pitch;
  a.hh = hh; a.ihsq = 1.0f / hh / hh; a.zgbc = zgbc ? 1 : 0;
  set_rows(a, rows);
  if (g_tile_variant == 2)
    launch_run<MODE_PRE>(a, stream, lc, K_MG_PRE, level);
  else
    launch_tile<3, MODE_PRE, LH_MAIN, NT_MAIN>(a, stream, lc, K_MG_PRE, level);
}

void launch_mg_post(const float *p_in, float *p_out, const Grid &f, const uint8_t *mask,
                    const Grid &ec, const uint8_t *maskc, float hh, bool zgbc, cudaStream_t stream,
                    LaunchCounter *lc, int level, const Rows *rows, const Rows *crows) {
  TileArgs a{};
  a.p_in = p_in; a.p_out = p_out; a.f = f.d; a.mask = mask;
  a.w = f.w; a.h = f.h; a.pitch = f.pitch;
  a.ec = ec.d; a.maskc = maskc; a.wc = ec.w; a.hc = ec.h; a.pc = ec.pitch;
  a.hh = hh; a.ihsq = 1.0f / hh / hh; a.zgbc = zgbc ? 1 : 0;
  set_rows(a, rows);
  if (crows) {
    a.c_lo = crows->st_lo;
    a.c_hi = crows->st_hi;
  }
  if (g_tile_variant == 2)
    launch_run<MODE_POST>(a, stream, lc, K_MG_POST, level);
  else
    launch_tile<3, MODE_POST, LH_MAIN, NT_MAIN>(a, stream, lc, K_MG_POST, level);
}

void launch_mg_smooth5(float *p_out, const Grid &f, const uint8_t *mask, float hh,
                       cudaStream_t stream, LaunchCounter *lc, int level) {
  TileArgs a{};
  a.p_in = nullptr; a.p_out = p_out; a.f = f.d; a.mask = mask;
  a.w = f.w; a.h = f.h; a.pitch = f.pitch;
  a.hh = hh; a.ihsq = 1.0f / hh / hh; a.zgbc = 0;
  set_rows(a, nullptr);
  launch_tile<5, MODE_SMOOTH, LH_MAIN, NT_MAIN>(a, stream, lc, K_MG_COARSE, level);
}

// First level t >= max(1, t_min) from which levels t..L fit one CTA's shared memory (0: none).
constexpr size_t TAIL_SMEM_MAX = 220 * 1024;
int mg_tail_first_level(const std::vector<TailLevel> &lv, int t_min) {
  const int L = (int)lv.size() - 2;
  if (g_tile_variant != 2 || L < 1) return 0;
  for (int t = t_min > 1 ? t_min : 1; t <= L; t++) {
    size_t cells = 0;
    for (int l = t; l <= L; l++) cells += (size_t)round_up(lv[l].w, 4) * lv[l].h;
    // one SM runs the tail: above ~8K cells on its first level the multi-CTA passes are faster
    if (L - t + 1 <= TAIL_MAXLV && cells * 9 + 16 <= TAIL_SMEM_MAX && (size_t)lv[t].w * lv[t].h <= g_tail_max_cells)
      return t;
  }
  return 0;
}

// solveLevel(t) from the zero guess: rhs = rc of level t, solution -> ec of level t
void launch_mg_tail(const std::vector<TailLevel> &lv, int t, const float *hh, const float *rhs,
                    float *out, cudaStream_t stream, LaunchCounter *lc) {
  const int L = (int)lv.size() - 2;
  TailArgs a{};
  a.n = L - t + 1;
  int cells = 0;
  for (int i = 0; i < a.n; i++) {
    const TailLevel &V = lv[t + i];
    a.w[i] = V.w; a.h[i] = V.h; a.gp[i] = V.pitch; a.sp[i] = round_up(V.w, 4); a.off[i] = cells;
    a.hh[i] = hh[t + i];
    a.mask[i] = V.mask;
    cells += a.sp[i] * V.h;
  }
  a.cells = cells;
  a.rhs = rhs;
  a.out = out;
  const size_t smem = (size_t)cells * 9 + 16;
  static std::atomic<unsigned long long> attr_done{0};
  ensure_dyn_smem(k_mg_tail, TAIL_SMEM_MAX, attr_done);
  UBGL_LAUNCH(lc, K_MG_COARSE, t, stream, launch_k(k_mg_tail, 1, TAIL_NT, smem, stream, a));
}

void launch_make_mask(const Grid &flag, uint8_t *mask, int *d_nonbinary, cudaStream_t stream,
                      LaunchCounter *lc, int level, const Rows *rows, const Rows *crows) {
  const int r_lo = rows ? rows->own_lo : 0, r_hi = rows ? rows->own_hi : flag.h;
  const int st_lo = rows ? rows->st_lo : 0, st_hi = rows ? rows->st_hi : flag.h;
  dim3 b(32, 8), g(ceil_div(flag.pitch, 32), ceil_div(r_hi - r_lo, 8));
  UBGL_LAUNCH(lc, K_COARSEN, level, stream, k_make_mask<<<g, b, 0, stream>>>(flag, mask, d_nonbinary, r_lo, r_hi, st_lo, st_hi));
}

} // namespace ubgl
