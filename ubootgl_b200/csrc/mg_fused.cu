// mg_fused.cu -- temporally blocked multigrid tile kernels (sm_100a).
//
// One CTA stages a (128 x LH) window of p, f and the 1-byte stencil mask in
// shared memory, runs S complete red-black Gauss-Seidel sweeps on it (the valid
// region shrinks by one cell per half-sweep, so the window carries a halo of
// 2S (+2) cells) and fuses the neighbouring multigrid operators into the same
// pass:
//   MODE_PRE    : S x (rbgs [+ zero-gradient BC]) -> residual -> full-weighting
//                 restriction.  HBM traffic per cell: read p, f, mask (9 B),
//                 write p (4 B) + rc (1 B); the residual field is never stored.
//   MODE_POST   : prolongate + correct [+ BC] -> S x (rbgs [+ BC]).
//   MODE_SMOOTH : S x rbgs from a zero initial guess (coarsest level).
// p is ping-ponged between two buffers (tiles read their neighbours' cells as
// halo, so an in-place update would race).
//
// The kernels are instruction-issue bound, not DRAM bound (ncu, profiles/), so
// the inner loops are written for instruction count:
//  * red and black cells live in separate shared-memory planes (plane =
//    (x+y)&1, column x>>1); a thread updates 4 consecutive same-colour cells
//    with 128-bit LDS/STS;
//  * solid cells are stored as 0 in the window (their value is never used by
//    the reference either: every read is multiplied by the cell's flag,
//    pressure_solver.cpp:13-17,101-108), so the 5-point sum needs no flag
//    multiplies; the divide by the fluid-neighbour count and the centre flag
//    collapse into one multiply by a table weight (stencils.cuh rcp_count);
//  * border cells (never smoothed) keep flag*value in the window for their
//    neighbours and are re-materialised from p_in / the zero-gradient copy at
//    write-back.
// Every per-cell rounding sequence equals the plain one-kernel-per-operator
// path in mg.cu for binary flags: the fused result is bit-identical to the
// plain one up to the sign of zeros (tests/test_gpu_fused.py).
//
// Reference semantics: pressure_solver.cpp:10-24 (smoothingKernel), :35-72
// (canonical red-black order), :91-116 (residual), :118-132 (restrict),
// :134-181 (prolongate, correct), :183-192 (setZeroGradientBC), :201-248
// (solveLevel).
#include "mg.cuh"
#include "stencils.cuh"

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
  static bool attr_set = false;
  if (!attr_set) {
    UBGL_CUDA(cudaFuncSetAttribute(k_mg_tile<S, MODE, LH, NT>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  dim3 grid(ceil_div(a.w, TX), ceil_div(a.own_hi - a.own_lo, TY));
  UBGL_LAUNCH(lc, kind, level, stream, k_mg_tile<S, MODE, LH, NT><<<grid, NT, smem, stream>>>(a));
}

constexpr int LH_MAIN = 80, NT_MAIN = 256;

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

void launch_make_mask(const Grid &flag, uint8_t *mask, int *d_nonbinary, cudaStream_t stream,
                      LaunchCounter *lc, int level, const Rows *rows, const Rows *crows) {
  const int r_lo = rows ? rows->own_lo : 0, r_hi = rows ? rows->own_hi : flag.h;
  const int st_lo = rows ? rows->st_lo : 0, st_hi = rows ? rows->st_hi : flag.h;
  dim3 b(32, 8), g(ceil_div(flag.pitch, 32), ceil_div(r_hi - r_lo, 8));
  UBGL_LAUNCH(lc, K_COARSEN, level, stream, k_make_mask<<<g, b, 0, stream>>>(flag, mask, d_nonbinary, r_lo, r_hi, st_lo, st_hi));
}

} // namespace ubgl
