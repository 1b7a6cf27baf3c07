// sim_fused.cu -- fused non-multigrid kernels of Simulation::step (sm_100a).
//
//   k_prestep<COMP> : applyAccumulatedVelocity + both diffuse passes of one
//                     velocity component in one tile pass (simulation.cpp:
//                     376-396, 104-162), the setVBCs between the passes applied
//                     to the window in shared memory.
//   k_vbc_cols/rows : setVBCs (simulation.cpp:80-102) and setPBC (:36-45), fully
//                     parallel over the perimeter (column phase, then row phase).
//   k_divergence4   : project part 1 (:166-171), 128-bit accesses, also zeroes
//                     the accumulator interiors (:384,392).
//   k_gradient_save : gradient subtraction + saveCurrentVelocityFields
//                     (:196-207, :16-19) in one pass.
//
// Flags are read as the level-0 stencil mask (1 B/cell, mg_fused.cu MB_*), so
// the fused step requires binary flags; otherwise DeviceSim uses the plain
// kernels of sim.cu.  Per-cell arithmetic = stencils.cuh, i.e. bit-identical to
// the plain path (tests/test_gpu_fused.py).
#include "sim.cuh"
#include "stencils.cuh"
#include <algorithm>
#include <cstdlib>

namespace ubgl {

enum { MB_C = 1, MB_W = 2, MB_E = 32, MB_S = 64, MB_N = 128 };

// ---------------------------------------------------------------------------
// k_prestep
// ---------------------------------------------------------------------------
constexpr int PW = 128;        // window width (one float4 per lane)
constexpr int PTX = PW - 8;    // tile width  (halo 2, rounded to 4 for alignment)
constexpr int PTY = 32;        // tile height
constexpr int PWH = PTY + 4;   // window height (halo 2)
constexpr int PRS = PW + 8;    // shared row stride, data at column 4
constexpr int PNT = 256;

struct PrestepSmem {
  float V[PWH][PRS]; // v0 = front + accum, later the pass-2 result
  float D[PWH][PRS]; // pass-1 result with its boundary values
  uint8_t M[PWH][PRS];
};

// BC value of a border cell given its interior neighbour a and its own value b.
// vx: columns are "parallel" sides, rows "perpendicular" (simulation.cpp:83-91);
// vy the other way round (:92-100).
template <int COMP> __device__ __forceinline__ float bc_col(int bc, float a, float b) {
  return COMP == 0 ? vbc_par(bc, a, b) : vbc_per(bc, a, b);
}
template <int COMP> __device__ __forceinline__ float bc_row(int bc, float a, float b) {
  return COMP == 0 ? vbc_per(bc, a, b) : vbc_par(bc, a, b);
}

template <int COMP> __global__ void __launch_bounds__(PNT) k_prestep(PrestepArgs g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PrestepSmem &sm = *reinterpret_cast<PrestepSmem *>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NW = PNT / 32;
  int tbx = blockIdx.x, tby = blockIdx.y;
  if (g.frame) { // linear id -> frame tile (see PrestepArgs)
    int id = blockIdx.x;
    const int low = g.f_byl * g.f_nbx, cw = 1 + g.f_nbx - g.f_bxf, mid = (g.f_byf - g.f_byl) * cw;
    if (id < low) {
      tby = id / g.f_nbx;
      tbx = id - tby * g.f_nbx;
    } else if (id < low + mid) {
      id -= low;
      tby = g.f_byl + id / cw;
      const int c = id % cw;
      tbx = c == 0 ? 0 : g.f_bxf + c - 1;
    } else {
      id -= low + mid;
      tby = g.f_byf + id / g.f_nbx;
      tbx = id % g.f_nbx;
    }
  }
  const int x0 = tbx * PTX, y0 = g.own_lo + tby * PTY;
  const int X0 = x0 - 4, Y0 = y0 - 2;
  const int gw = g.gw, gh = g.gh;
  const int s_hi = min(gh, g.st_hi), m_hi = min(g.H, g.st_hi), o_hi = min(gh, g.own_hi);

  auto interior = [&](int gx, int gy) { return gx >= 1 && gx <= gw - 2 && gy >= 1 && gy <= gh - 2; };

  // ---- stage: v0 = A (+ acc on the interior), keep-values, mask ----
  for (int r = warp; r < PWH; r += NW) {
    const int gy = Y0 + r, gx = X0 + 4 * lane;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f), k = v;
    uchar4 m = make_uchar4(0, 0, 0, 0);
    if (gy >= g.st_lo && gx >= 0 && gx < g.pitch) {
      const size_t o = (size_t)gy * g.pitch + gx;
      if (gy < s_hi) {
        v = *reinterpret_cast<const float4 *>(g.A + o);
        const bool rowin = gy >= 1 && gy <= gh - 2;
        if (rowin) {
          const float4 ac = *reinterpret_cast<const float4 *>(g.acc + o);
          if (gx >= 1 && gx <= gw - 2) v.x = __fadd_rn(v.x, ac.x);
          if (gx + 1 >= 1 && gx + 1 <= gw - 2) v.y = __fadd_rn(v.y, ac.y);
          if (gx + 2 >= 1 && gx + 2 <= gw - 2) v.z = __fadd_rn(v.z, ac.z);
          if (gx + 3 >= 1 && gx + 3 <= gw - 2) v.w = __fadd_rn(v.w, ac.w);
        }
        k = v;
        if (COMP == 0 && (!rowin || gx == 0 || (gw - 1 >= gx && gw - 1 <= gx + 3)))
          k = *reinterpret_cast<const float4 *>(g.K + o);
      }
      if (gy < m_hi) m = __ldg(reinterpret_cast<const uchar4 *>(g.mask + o));
    }
    *reinterpret_cast<float4 *>(&sm.V[r][4 + 4 * lane]) = v;
    *reinterpret_cast<float4 *>(&sm.D[r][4 + 4 * lane]) = k;
    *reinterpret_cast<uchar4 *>(&sm.M[r][4 + 4 * lane]) = m;
    if (lane < 2) {
      const int c = lane ? 4 + PW : 0;
      *reinterpret_cast<float4 *>(&sm.V[r][c]) = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4 *>(&sm.D[r][c]) = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<uchar4 *>(&sm.M[r][c]) = make_uchar4(0, 0, 0, 0);
    }
  }
  __syncthreads();

  // setVBCs on the window held in `arr` (non-corner border cells whose interior
  // neighbour is inside the window; corners are never read by the 5-point pass)
  auto fix_borders = [&](float(*arr)[PRS]) {
    const int t = threadIdx.x;
    const int ly_lo = max(0, 1 - Y0), ly_hi = min(PWH, gh - 1 - Y0); // rows with 1 <= gy <= gh-2
    const int lx_lo = max(0, 1 - X0), lx_hi = min(PW, gw - 1 - X0);
    if (X0 <= 0 && -X0 + 1 < PW) {
      const int c = 4 - X0;
      for (int r = ly_lo + t; r < ly_hi; r += PNT) arr[r][c] = bc_col<COMP>(g.bcLo, arr[r][c + 1], arr[r][c]);
    }
    if (gw - 1 - X0 >= 1 && gw - 1 - X0 < PW) {
      const int c = 4 + gw - 1 - X0;
      for (int r = ly_lo + t; r < ly_hi; r += PNT) arr[r][c] = bc_col<COMP>(g.bcHi, arr[r][c - 1], arr[r][c]);
    }
    if (Y0 <= 0 && -Y0 + 1 < PWH) {
      const int r = -Y0;
      for (int c = lx_lo + t; c < lx_hi; c += PNT) arr[r][4 + c] = bc_row<COMP>(g.bcS, arr[r + 1][4 + c], arr[r][4 + c]);
    }
    if (gh - 1 - Y0 >= 1 && gh - 1 - Y0 < PWH) {
      const int r = gh - 1 - Y0;
      for (int c = lx_lo + t; c < lx_hi; c += PNT) arr[r][4 + c] = bc_row<COMP>(g.bcN, arr[r - 1][4 + c], arr[r][4 + c]);
    }
  };

  const bool edge = X0 <= 0 || Y0 <= 0 || gw - 1 - X0 < PW || gh - 1 - Y0 < PWH;
  if (COMP == 1) {
    // vy: the setVBCs calls issued during the vx passes (simulation.cpp:132) have
    // already refreshed vy's border from its post-accumulation interior
    if (edge) {
      fix_borders(sm.V);
      __syncthreads();
      // those values are what the BC after vy's first pass keeps
      for (int i = threadIdx.x; i < PWH * (PW / 4); i += PNT) {
        const int r = i / (PW / 4), c = 4 + 4 * (i % (PW / 4));
        const int gy = Y0 + r, gx = X0 + c - 4;
        if (gy == 0 || gy == gh - 1 || gx == 0 || (gw - 1 >= gx && gw - 1 <= gx + 3))
          *reinterpret_cast<float4 *>(&sm.D[r][c]) = *reinterpret_cast<const float4 *>(&sm.V[r][c]);
      }
      __syncthreads();
    }
  }

  // one diffuse pass over window rows [r_lo, r_hi), 4 cells per thread
  auto pass = [&](float(*src)[PRS], float(*dst)[PRS], int r_lo, int r_hi, int g_lo, int g_hi) {
    const int ng = g_hi - g_lo;
    for (int i = threadIdx.x; i < (r_hi - r_lo) * ng; i += PNT) {
      const int r = r_lo + i / ng, c = 4 + 4 * (g_lo + i % ng);
      const int gy = Y0 + r, gx = X0 + c - 4;
      if (gy < 1 || gy > gh - 2) continue;
      const float4 C4 = *reinterpret_cast<const float4 *>(&src[r][c]);
      const float4 N4 = *reinterpret_cast<const float4 *>(&src[r + 1][c]);
      const float4 S4 = *reinterpret_cast<const float4 *>(&src[r - 1][c]);
      const float wv = src[r][c - 1], ev = src[r][c + 4];
      const unsigned mc = *reinterpret_cast<const unsigned *>(&sm.M[r][c]);
      const unsigned mn = *reinterpret_cast<const unsigned *>(&sm.M[r + 1][c]);
      const unsigned ms = *reinterpret_cast<const unsigned *>(&sm.M[r - 1][c]);
      const unsigned mE = sm.M[r][c + 4], mW = sm.M[r][c - 1];
      const float cc[4] = {C4.x, C4.y, C4.z, C4.w};
      const float nn[4] = {N4.x, N4.y, N4.z, N4.w};
      const float ss[4] = {S4.x, S4.y, S4.z, S4.w};
      const float ww[4] = {wv, C4.x, C4.y, C4.z};
      const float ee[4] = {C4.y, C4.z, C4.w, ev};
      float out[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const unsigned bC = (mc >> (8 * j)) & 255u;
        const unsigned bE = j < 3 ? (mc >> (8 * j + 8)) & 255u : mE;
        const unsigned bW = j > 0 ? (mc >> (8 * j - 8)) & 255u : mW;
        const unsigned bN = (mn >> (8 * j)) & 255u, bS = (ms >> (8 * j)) & 255u;
        auto both = [](unsigned b, unsigned bits) { return (b & bits) == bits; };
        const float c0 = cc[j];
        float val;
        bool mC;
        if (COMP == 0) { // simulation.cpp:117-127
          val = both(bE, MB_C | MB_E) ? ee[j] : 0.0f;
          val = __fadd_rn(both(bC, MB_C | MB_W) ? ww[j] : 0.0f, val);
          val = __fadd_rn(val, both(bN, MB_C | MB_W) ? nn[j] : -c0);
          val = __fadd_rn(val, both(bS, MB_C | MB_W) ? ss[j] : -c0);
          mC = both(bC, MB_C | MB_E);
        } else { // simulation.cpp:143-153
          val = both(bC, MB_C | MB_S) ? ss[j] : 0.0f;
          val = __fadd_rn(both(bN, MB_C | MB_N) ? nn[j] : 0.0f, val);
          val = __fadd_rn(val, both(bE, MB_C | MB_N) ? ee[j] : -c0);
          val = __fadd_rn(val, both(bW, MB_C | MB_N) ? ww[j] : -c0);
          mC = both(bC, MB_C | MB_N);
        }
        out[j] = mC ? __fmul_rn(__fmaf_rn(g.a, val, c0), g.rden) : 0.0f;
      }
      if (gx >= 1 && gx + 3 <= gw - 2) {
        *reinterpret_cast<float4 *>(&dst[r][c]) = make_float4(out[0], out[1], out[2], out[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; j++)
          if (gx + j >= 1 && gx + j <= gw - 2) dst[r][c + j] = out[j];
      }
    }
  };

  pass(sm.V, sm.D, 1, PWH - 1, 0, PW / 4); // pass 1 (i = 1), simulation.cpp:113-129 / :139-155
  __syncthreads();
  if (edge) {
    fix_borders(sm.D); // swap + setVBCs, simulation.cpp:130-132
    __syncthreads();
  }
  pass(sm.D, sm.V, 2, PWH - 2, 1, PW / 4 - 1); // pass 2 (i = 2) on the tile proper
  __syncthreads();

  // ---- write back: B <- pass 1 (interior), Cout <- pass 2 (interior) and the
  // boundary values the next setVBCs keeps ----
  for (int i = threadIdx.x; i < PTY * (PTX / 4); i += PNT) {
    const int r = 2 + i / (PTX / 4), c = 8 + 4 * (i % (PTX / 4));
    const int gy = Y0 + r, gx = X0 + c - 4;
    if (gy >= o_hi || gx >= gw) continue;
    const size_t o = (size_t)gy * g.pitch + gx;
    const float4 d1 = *reinterpret_cast<const float4 *>(&sm.D[r][c]);
    const float4 d2 = *reinterpret_cast<const float4 *>(&sm.V[r][c]);
    if (gy >= 1 && gy <= gh - 2 && gx >= 1 && gx + 3 <= gw - 2) {
      *reinterpret_cast<float4 *>(g.B + o) = d1;
      *reinterpret_cast<float4 *>(g.Cout + o) = d2;
    } else {
      const float a1[4] = {d1.x, d1.y, d1.z, d1.w}, a2[4] = {d2.x, d2.y, d2.z, d2.w};
#pragma unroll
      for (int j = 0; j < 4; j++) {
        if (gx + j >= gw) break;
        if (interior(gx + j, gy)) {
          g.B[o + j] = a1[j];
          g.Cout[o + j] = a2[j];
        } else {
          g.Cout[o + j] = a1[j];
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// k_prestep_run<COMP> -- the same two passes for the tiles whose windows touch no border
// cell (all but a one-tile frame of the grid), entirely in REGISTERS.
//
// k_prestep above is bound by the shared-memory pipe (ncu, 8192^2: L1/shared 80 %, DRAM 42 %):
// every 4-cell group re-reads three 128-bit rows, two scalars and five mask words per pass.
// Here a warp owns a 128-column strip (lane = 4 consecutive cells, 120 useful) and walks RW rows
// bottom to top: the rows of v0 = front + accumulator, of the pass-1 result and of the stencil
// mask it needs are a rolling window of registers (a row is loaded from HBM once, as one
// 128-bit load per array), the W / E neighbours outside a lane's four cells come from the
// adjacent lanes by shuffle, and the pass-2 row trails the pass-1 row by one.  No shared
// memory, no block barrier; the price is pass 1 on two extra rows per strip and 8 idle columns
// per 128.  Away from the border no setVBCs intervenes between the passes, so there is no BC
// logic here: the frame tiles keep running k_prestep (launch_prestep).  Per-cell arithmetic and
// predicate tests are those of k_prestep's pass(): bit-identical results.
// ---------------------------------------------------------------------------
template <int COMP>
__device__ __forceinline__ float4 prestep_row(const float4 &S4, const float4 &C4, const float4 &N4, float wv, float ev,
                                              unsigned ms, unsigned mc, unsigned mn, unsigned mW, unsigned mE,
                                              float a, float rden) {
  const float cc[4] = {C4.x, C4.y, C4.z, C4.w};
  const float nn[4] = {N4.x, N4.y, N4.z, N4.w};
  const float ss[4] = {S4.x, S4.y, S4.z, S4.w};
  const float ww[4] = {wv, C4.x, C4.y, C4.z};
  const float ee[4] = {C4.y, C4.z, C4.w, ev};
  float out[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const unsigned bC = (mc >> (8 * j)) & 255u;
    const unsigned bE = j < 3 ? (mc >> (8 * j + 8)) & 255u : mE;
    const unsigned bW = j > 0 ? (mc >> (8 * j - 8)) & 255u : mW;
    const unsigned bN = (mn >> (8 * j)) & 255u, bS = (ms >> (8 * j)) & 255u;
    auto both = [](unsigned b, unsigned bits) { return (b & bits) == bits; };
    const float c0 = cc[j];
    float val;
    bool mC;
    if (COMP == 0) { // simulation.cpp:117-127
      val = both(bE, MB_C | MB_E) ? ee[j] : 0.0f;
      val = __fadd_rn(both(bC, MB_C | MB_W) ? ww[j] : 0.0f, val);
      val = __fadd_rn(val, both(bN, MB_C | MB_W) ? nn[j] : -c0);
      val = __fadd_rn(val, both(bS, MB_C | MB_W) ? ss[j] : -c0);
      mC = both(bC, MB_C | MB_E);
    } else { // simulation.cpp:143-153
      val = both(bC, MB_C | MB_S) ? ss[j] : 0.0f;
      val = __fadd_rn(both(bN, MB_C | MB_N) ? nn[j] : 0.0f, val);
      val = __fadd_rn(val, both(bE, MB_C | MB_N) ? ee[j] : -c0);
      val = __fadd_rn(val, both(bW, MB_C | MB_N) ? ww[j] : -c0);
      mC = both(bC, MB_C | MB_N);
    }
    out[j] = mC ? __fmul_rn(__fmaf_rn(a, val, c0), rden) : 0.0f;
  }
  return make_float4(out[0], out[1], out[2], out[3]);
}

struct PrestepRunGeom {
  int bx0, bx1; // interior tile columns [bx0, bx1) of PTX cells
  int ya, yb;   // interior rows [ya, yb)
};

// One WARP per block: the strip's row range depends on blockIdx only, so every loop bound and
// row test below is provably warp-uniform (uniform registers, plain branches around the shuffles).
// PRW = rows per warp strip: 64 on large grids (pass 1 runs on PRW + 2 rows), 16 where that would
// leave the GPU short of warps.
template <int COMP, int PRW>
__global__ void __launch_bounds__(32, 32) k_prestep_run(PrestepArgs g, PrestepRunGeom q) {
  const int lane = threadIdx.x;
  const int bx = q.bx0 + blockIdx.x;
  const int c0 = q.ya + blockIdx.y * PRW; // first output row of this warp
  const int c1 = min(c0 + PRW, q.yb);
  const int gx = bx * PTX - 4 + 4 * lane; // first of this lane's four cells
  const bool col_ok = gx >= 0 && gx < g.pitch;
  const bool useful = lane >= 1 && lane <= 30;
  const int s_hi = min(g.gh, g.st_hi), m_hi = min(g.H, g.st_hi);
  const int pitch = g.pitch;

  float4 v[4], d[4]; // rolling rows: slot = row & 3
  unsigned m[4], mWn[4], mEn[4];
  // A row is fetched in two halves, one loop iteration apart: issue_row() starts the three
  // 128-/32-bit loads of row r + 2 into nA / nC / nM, commit_row() of the NEXT iteration turns
  // them into v0 = front + accumulator and the neighbour mask bytes.  Between the two sits one
  // iteration of arithmetic (both passes of one row), so the warp no longer stalls on its own
  // loads (ncu r02: long scoreboard was 72 % of this kernel's stall samples at 62 % DRAM).
  // Measured at 8192^2, both components: 0.492 -> 0.460 ms with the register budget held at 64
  // (__launch_bounds__(32, 32): 32 warps per SM; at 68 registers / 28 warps it LOST: 0.522 ms);
  // an additional prefetch.global.L2 four rows ahead lost as well (0.524 ms) and was removed.
  float4 nA, nC;
  unsigned nM;
  auto issue_row = [&](int r) {
    nA = nC = make_float4(0.f, 0.f, 0.f, 0.f);
    nM = 0u;
    if (col_ok && r >= g.st_lo) {
      const size_t o = (size_t)r * pitch + gx;
      if (r < s_hi) {
        nA = *reinterpret_cast<const float4 *>(g.A + o);
        nC = *reinterpret_cast<const float4 *>(g.acc + o); // every row here is interior
      }
      if (r < m_hi) nM = __ldg(reinterpret_cast<const unsigned *>(g.mask + o));
    }
  };
  auto commit_row = [&](int k) {
    v[k] = make_float4(__fadd_rn(nA.x, nC.x), __fadd_rn(nA.y, nC.y), __fadd_rn(nA.z, nC.z), __fadd_rn(nA.w, nC.w));
    m[k] = nM;
    // mask bytes of the cells left and right of this lane's four
    mWn[k] = __shfl_up_sync(0xffffffffu, nM, 1) >> 24;
    mEn[k] = __shfl_down_sync(0xffffffffu, nM, 1) & 255u;
  };
  auto west = [&](const float4 &x) { return __shfl_up_sync(0xffffffffu, x.w, 1); };
  auto east = [&](const float4 &x) { return __shfl_down_sync(0xffffffffu, x.x, 1); };

  // Row r lives in slot r & 3.  The loop starts on a multiple of 4 so that inside the unrolled
  // body every slot index is a compile-time constant (the arrays stay in registers); the up to
  // three rows before the first needed one are skipped by uniform tests.
  // Iteration r: row r+1 arrives (its loads were issued one iteration earlier), the loads of row
  // r+2 start; pass 1 on row r (rows c0-1 .. c1); pass 2 on row r-1 (rows c0 .. c1-1).
  issue_row(c0 - 2);
  for (int rb = (c0 - 3) & ~3; rb <= c1; rb += 4) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int r = rb + u;
      const int kC = u, kS = (u + 3) & 3, kN = (u + 1) & 3, kS2 = (u + 2) & 3;
      if (r + 1 >= c0 - 2 && r <= c1) commit_row(kN);
      if (r + 2 >= c0 - 1 && r + 2 <= c1 + 1) issue_row(r + 2);
      if (r >= c0 - 1 && r <= c1) {
        const float4 D = prestep_row<COMP>(v[kS], v[kC], v[kN], west(v[kC]), east(v[kC]), m[kS], m[kC], m[kN],
                                           mWn[kC], mEn[kC], g.a, g.rden);
        d[kC] = D;
        if (useful && r >= c0 && r < c1) *reinterpret_cast<float4 *>(g.B + (size_t)r * pitch + gx) = D;
        if (r - 1 >= c0) { // pass 2 on row r-1: pass-1 rows r-2, r-1, r
          const float4 O = prestep_row<COMP>(d[kS2], d[kS], D, west(d[kS]), east(d[kS]), m[kS2], m[kS], m[kC],
                                             mWn[kS], mEn[kS], g.a, g.rden);
          if (useful) *reinterpret_cast<float4 *>(g.Cout + (size_t)(r - 1) * pitch + gx) = O;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// setVBCs / setPBC, parallel over the perimeter.  Column phase and row phase are
// separate launches: the row phase reads the column results at x = 0 / w-1 and
// wins at the corners, exactly as the reference's loop order.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void put3(const Grid &a, const Grid &b, const Grid &c, int x, int y, float v) {
  a.at(x, y) = v;
  b.at(x, y) = v;
  if (c.d) c.at(x, y) = v;
}

// Columns and rows in ONE launch (the game level's step is ~25 launches of ~8 us; three setVBCs
// of two launches each were a fifth of it).  The reference's row loops run after its column loops
// and read the column results at x = 0 / w-1 (simulation.cpp:80-102); the only row cells that
// depend on them are the four corners, whose interior neighbour (0, 1), (w-1, 1), (0, h-2),
// (w-1, h-2) is itself a column-phase cell.  A corner thread therefore evaluates that column
// value itself from the interior cell next to it: whether its read of the cell's own value
// (only INFLOW uses it, and keeps it) sees the column thread's store or not, the value is the same.
struct BcGrid {
  Grid f, b, c; // front, back, optional third copy
  int comp;     // 0: vx (columns parallel, rows perpendicular), 1: vy
};
__device__ __forceinline__ float bcv(int comp, bool col, int bc, float a, float own) {
  return (comp == 0) == col ? vbc_par(bc, a, own) : vbc_per(bc, a, own);
}
__device__ __forceinline__ void vbc_col_cell(const BcGrid &q, const BorderArgs &g, int y) {
  const int w = q.f.w;
  put3(q.f, q.b, q.c, 0, y, bcv(q.comp, true, g.bcW, q.f.at(1, y), q.f.at(0, y)));
  put3(q.f, q.b, q.c, w - 1, y, bcv(q.comp, true, g.bcE, q.f.at(w - 2, y), q.f.at(w - 1, y)));
}
__device__ __forceinline__ void vbc_row_cell(const BcGrid &q, const BorderArgs &g, int x) {
  const int w = q.f.w, h = q.f.h;
  // The corner threads also do their interior neighbour (x = 1 / w-2), whose thread does
  // nothing: the corner value may depend on that cell's value BEFORE the row phase (its own
  // column-phase value, kept when the row side is INFLOW), read here before anybody writes it.
  if (x == 1 || x == w - 2) return;
  const bool corner = x == 0 || x == w - 1;
  const int xn = x == 0 ? 1 : w - 2, bcx = x == 0 ? g.bcW : g.bcE;
  auto row = [&](int y, int yi, int bc) { // border row y, interior row yi next to it
    if (!corner) {
      put3(q.f, q.b, q.c, x, y, bcv(q.comp, false, bc, q.f.at(x, yi), q.f.at(x, y)));
      return;
    }
    const float nb_old = q.f.at(xn, y);                                              // (xn, y) before the row phase
    const float own_col = bcv(q.comp, true, bcx, nb_old, q.f.at(x, y));              // (x, y) after the column phase
    const float in_col = bcv(q.comp, true, bcx, q.f.at(xn, yi), q.f.at(x, yi));      // (x, yi) after the column phase
    const float nb_new = bcv(q.comp, false, bc, q.f.at(xn, yi), nb_old);
    put3(q.f, q.b, q.c, x, y, bcv(q.comp, false, bc, in_col, own_col));
    put3(q.f, q.b, q.c, xn, y, nb_new);
  };
  if (g.do_s) row(0, 1, g.bcS);
  if (g.do_n) row(h - 1, h - 2, g.bcN);
}

// thread t < ncol: column cells of row y_lo + t; else row cells of column t - ncol
__global__ void k_vbc_all(BorderArgs g, int ncol) {
  ubgl_pdl_prologue();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const BcGrid qx{g.xf, g.xb, g.xc, 0}, qy{g.yf, g.yb, g.yc, 1};
  if (t < ncol) {
    const int y = g.y_lo + t;
    // rows 0 and h-1 of a grid belong to the row phase (it runs last in the reference); the
    // column phase writes them too there, with values the row phase overwrites: skipped here
    if (y >= 1 && y < g.xf.h - 1) vbc_col_cell(qx, g, y);
    if (y >= 1 && y < g.yf.h - 1) vbc_col_cell(qy, g, y);
    if (g.p.d && y >= 1 && y < g.p.h - 1) {
      g.p.at(0, y) = single_pbc(g.bcW, g.p.at(1, y));
      g.p.at(g.p.w - 1, y) = single_pbc(g.bcE, g.p.at(g.p.w - 2, y));
    }
    return;
  }
  const int x = t - ncol;
  if (!(g.do_s || g.do_n)) return;
  if (x < g.xf.w) vbc_row_cell(qx, g, x);
  if (x < g.yf.w) vbc_row_cell(qy, g, x);
  if (g.p.d && x < g.p.w) {
    const int w = g.p.w, h = g.p.h;
    auto inner = [&](int yy) {
      if (x == 0) return single_pbc(g.bcW, g.p.at(1, yy));
      if (x == w - 1) return single_pbc(g.bcE, g.p.at(w - 2, yy));
      return g.p.at(x, yy);
    };
    if (g.do_s) g.p.at(x, 0) = single_pbc(g.bcS, inner(1));
    if (g.do_n) g.p.at(x, h - 1) = single_pbc(g.bcN, inner(h - 2));
  }
}

// ---------------------------------------------------------------------------
// divergence (simulation.cpp:166-171) + zeroing of the accumulator interiors
// (:384,:392).  One thread = 4 consecutive cells of a row.
// ---------------------------------------------------------------------------
__global__ void k_divergence4(Grid vx, Grid vy, Grid f, Grid ax, Grid ay, float ih, int y_lo,
                              int y_hi) {
  const int x = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  const int y = y_lo + blockIdx.y * blockDim.y + threadIdx.y; // y_lo >= 1, y_hi <= H-1
  const int W = f.w, H = f.h;
  if (x >= W - 1 || y >= y_hi) return;
  const size_t o = (size_t)y * f.pitch + x; // all grids share the pitch
  const float4 a = *reinterpret_cast<const float4 *>(vx.d + o);
  const float aw = x > 0 ? vx.d[o - 1] : 0.0f;
  const float4 b = *reinterpret_cast<const float4 *>(vy.d + o);
  const float4 c = *reinterpret_cast<const float4 *>(vy.d + o - f.pitch);
  auto dv = [&](float xe, float xw, float yn, float ys) {
    return __fmul_rn(-ih, __fsub_rn(__fadd_rn(__fsub_rn(xe, xw), yn), ys));
  };
  const float r[4] = {dv(a.x, aw, b.x, c.x), dv(a.y, a.x, b.y, c.y), dv(a.z, a.y, b.z, c.z),
                      dv(a.w, a.z, b.w, c.w)};
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  if (x >= 1 && x + 3 <= W - 3 && y <= H - 3) { // every cell interior in f, vx and vy
    *reinterpret_cast<float4 *>(f.d + o) = make_float4(r[0], r[1], r[2], r[3]);
    if (ax.d) {
      *reinterpret_cast<float4 *>(ax.d + o) = z;
      *reinterpret_cast<float4 *>(ay.d + o) = z;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int xx = x + j;
      if (xx < 1 || xx > W - 2) continue;
      f.d[o + j] = r[j];
      if (!ax.d) continue;
      if (xx <= W - 3) ax.d[o + j] = 0.0f; // vx interior: 1..W-3 x 1..H-2
      if (y <= H - 3) ay.d[o + j] = 0.0f;  // vy interior: 1..W-2 x 1..H-3
    }
  }
}

// ---------------------------------------------------------------------------
// gradient subtraction (simulation.cpp:196-207) + saveCurrentVelocityFields
// (:16-19) for the interior faces; the border faces of vx_current / vy_current
// are written by the setVBCs kernels that follow.
// ---------------------------------------------------------------------------
__global__ void k_gradient_save(Grid vx, Grid vy, Grid p, const uint8_t *mask, Grid cx, Grid cy,
                                float ih, int y_lo, int y_hi) {
  const int x = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  const int y = y_lo + blockIdx.y * blockDim.y + threadIdx.y; // y_lo >= 1, y_hi <= H-1
  const int W = p.w, H = p.h;
  if (x >= W - 1 || y >= y_hi) return;
  const size_t o = (size_t)y * p.pitch + x;
  const float4 pc = *reinterpret_cast<const float4 *>(p.d + o);
  const float4 pn = *reinterpret_cast<const float4 *>(p.d + o + p.pitch);
  const float pe = x + 4 < p.pitch ? p.d[o + 4] : 0.0f;
  const unsigned m = *reinterpret_cast<const unsigned *>(mask + o);
  float4 ux = *reinterpret_cast<const float4 *>(vx.d + o);
  float4 uy = *reinterpret_cast<const float4 *>(vy.d + o);
  auto gx = [&](float v, float p1, float p0, int j) {
    const unsigned b = (m >> (8 * j)) & (MB_C | MB_E);
    return b == (MB_C | MB_E) ? __fmaf_rn(-ih, __fsub_rn(p1, p0), v) : v;
  };
  auto gy = [&](float v, float p1, float p0, int j) {
    const unsigned b = (m >> (8 * j)) & (MB_C | MB_N);
    return b == (MB_C | MB_N) ? __fmaf_rn(-ih, __fsub_rn(p1, p0), v) : v;
  };
  ux.x = gx(ux.x, pc.y, pc.x, 0);
  ux.y = gx(ux.y, pc.z, pc.y, 1);
  ux.z = gx(ux.z, pc.w, pc.z, 2);
  ux.w = gx(ux.w, pe, pc.w, 3);
  uy.x = gy(uy.x, pn.x, pc.x, 0);
  uy.y = gy(uy.y, pn.y, pc.y, 1);
  uy.z = gy(uy.z, pn.z, pc.z, 2);
  uy.w = gy(uy.w, pn.w, pc.w, 3);
  if (x >= 1 && x + 3 <= W - 3 && y <= H - 3) {
    *reinterpret_cast<float4 *>(vx.d + o) = ux;
    *reinterpret_cast<float4 *>(vy.d + o) = uy;
    if (cx.d) { // null: the caller aliases *_current to the front buffers (DeviceSim::cur_alias)
      *reinterpret_cast<float4 *>(cx.d + o) = ux;
      *reinterpret_cast<float4 *>(cy.d + o) = uy;
    }
  } else {
    const float ax[4] = {ux.x, ux.y, ux.z, ux.w}, ay[4] = {uy.x, uy.y, uy.z, uy.w};
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int xx = x + j;
      if (xx < 1 || xx > W - 2) continue;
      if (xx <= W - 3) { // vx faces 1..W-3 x 1..H-2
        vx.d[o + j] = ax[j];
        if (cx.d) cx.d[o + j] = ax[j];
      }
      if (y <= H - 3) { // vy faces 1..W-2 x 1..H-3
        vy.d[o + j] = ay[j];
        if (cx.d) cy.d[o + j] = ay[j];
      }
    }
  }
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
// 0: k_prestep everywhere; 1 (default): register-run kernel inside, k_prestep on the frame
static int prestep_variant() {
  static const int v = [] {
    const char *e = getenv("UBGL_PRESTEP_VARIANT");
    return (e && e[0] == '0') ? 0 : 1;
  }();
  return v;
}

void launch_prestep(int comp, const PrestepArgs &g0, cudaStream_t stream, LaunchCounter *lc) {
  static std::atomic<unsigned long long> attr_done0{0}, attr_done1{0};
  ensure_dyn_smem(k_prestep<0>, sizeof(PrestepSmem), attr_done0);
  ensure_dyn_smem(k_prestep<1>, sizeof(PrestepSmem), attr_done1);
  PrestepArgs g = g0;
  const int rows = std::min(g.gh, g.own_hi) - g.own_lo;
  if (rows <= 0) return;
  const int nbx = ceil_div(g.gw, PTX), nby = ceil_div(rows, PTY);
  dim3 grid(nbx, nby);
  g.frame = 0;
  // tiles whose 128 x 36 window holds no border cell: columns [1, bxf), tile rows [byl, byf)
  const int bxf = g.gw >= 125 ? std::min(nbx, (g.gw - 125) / PTX + 1) : 0;
  const int byl = g.own_lo <= 1 ? 1 : 0;
  const int top = g.gh - 35 - g.own_lo;
  const int byf = top >= 0 ? std::min(nby, top / PTY + 1) : 0;
  PrestepRunGeom q{1, bxf, g.own_lo + PTY * byl, std::min(g.own_lo + PTY * byf, std::min(g.gh, g.own_hi))};
  // one warp per strip: worth it only where the strips fill the machine (8 warps on each of the
  // 148 SMs at least); the game level (1090 x 436: 184 strips) stays with k_prestep alone
  const bool can = prestep_variant() == 1 && bxf > 1 && byf > byl && q.yb - q.ya >= 2 * PTY;
  const int nx = can ? q.bx1 - q.bx0 : 0;
  const int prw = !can ? 0 : nx * ceil_div(q.yb - q.ya, 64) >= 2 * 1184 ? 64 : nx * ceil_div(q.yb - q.ya, 16) >= 1184 ? 16 : 0;
  if (prw) {
    g.frame = 1; g.f_nbx = nbx; g.f_byl = byl; g.f_byf = byf; g.f_bxf = bxf;
    grid = dim3(nbx * nby - (bxf - 1) * (byf - byl), 1);
    dim3 rg(nx, ceil_div(q.yb - q.ya, prw));
    if (comp == 0 && prw == 64)
      UBGL_LAUNCH(lc, K_PRESTEP, 0, stream, (k_prestep_run<0, 64><<<rg, 32, 0, stream>>>(g, q)));
    else if (comp == 0)
      UBGL_LAUNCH(lc, K_PRESTEP, 0, stream, (k_prestep_run<0, 16><<<rg, 32, 0, stream>>>(g, q)));
    else if (prw == 64)
      UBGL_LAUNCH(lc, K_PRESTEP, 0, stream, (k_prestep_run<1, 64><<<rg, 32, 0, stream>>>(g, q)));
    else
      UBGL_LAUNCH(lc, K_PRESTEP, 0, stream, (k_prestep_run<1, 16><<<rg, 32, 0, stream>>>(g, q)));
  }
  if (comp == 0)
    UBGL_LAUNCH(lc, K_PRESTEP, 0, stream, k_prestep<0><<<grid, PNT, sizeof(PrestepSmem), stream>>>(g));
  else
    UBGL_LAUNCH(lc, K_PRESTEP, 0, stream, k_prestep<1><<<grid, PNT, sizeof(PrestepSmem), stream>>>(g));
}

void launch_borders(const BorderArgs &g, cudaStream_t stream, LaunchCounter *lc) {
  const int ncol = std::max(0, g.y_hi - g.y_lo), nrow = (g.do_s || g.do_n) ? g.yf.w : 0;
  if (ncol + nrow > 0)
    UBGL_LAUNCH(lc, K_VBC, 0, stream, launch_k(k_vbc_all, ceil_div(ncol + nrow, 128), 128, 0, stream, g, ncol));
}

// setPBC alone (simulation.cpp:36-45) for the slab driver
__global__ void k_pbc_cols(Grid p, int bcW, int bcE, int y_lo, int y_hi) {
  ubgl_pdl_prologue();
  const int y = y_lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (y >= y_hi || y >= p.h) return;
  p.at(0, y) = single_pbc(bcW, p.at(1, y));
  p.at(p.w - 1, y) = single_pbc(bcE, p.at(p.w - 2, y));
}
__global__ void k_pbc_rows(Grid p, int bcS, int bcN, int do_s, int do_n) {
  ubgl_pdl_prologue();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= p.w) return;
  if (do_s) p.at(x, 0) = single_pbc(bcS, p.at(x, 1));
  if (do_n) p.at(x, p.h - 1) = single_pbc(bcN, p.at(x, p.h - 2));
}
void launch_pbc(const Grid &p, int bcW, int bcE, int bcN, int bcS, int y_lo, int y_hi, bool do_s,
                bool do_n, cudaStream_t stream, LaunchCounter *lc) {
  if (y_hi > y_lo)
    UBGL_LAUNCH(lc, K_PBC, 0, stream, launch_k(k_pbc_cols, ceil_div(y_hi - y_lo, 128), 128, 0, stream, p, bcW, bcE, y_lo, y_hi));
  if (do_s || do_n)
    UBGL_LAUNCH(lc, K_PBC, 0, stream, launch_k(k_pbc_rows, ceil_div(p.w, 128), 128, 0, stream, p, bcS, bcN, do_s, do_n));
}

void launch_divergence4(const Grid &vx, const Grid &vy, const Grid &f, const Grid &ax, const Grid &ay,
                        float ih, int y_lo, int y_hi, cudaStream_t stream, LaunchCounter *lc) {
  y_lo = std::max(y_lo, 1);
  y_hi = std::min(y_hi, f.h - 1);
  if (y_hi <= y_lo) return;
  dim3 b(32, 8), g(ceil_div(ceil_div(f.w - 1, 4), 32), ceil_div(y_hi - y_lo, 8));
  UBGL_LAUNCH(lc, K_DIVERGENCE, 0, stream, k_divergence4<<<g, b, 0, stream>>>(vx, vy, f, ax, ay, ih, y_lo, y_hi));
}

void launch_gradient_save(const Grid &vx, const Grid &vy, const Grid &p, const uint8_t *mask,
                          const Grid &cx, const Grid &cy, float ih, int y_lo, int y_hi,
                          cudaStream_t stream, LaunchCounter *lc) {
  y_lo = std::max(y_lo, 1);
  y_hi = std::min(y_hi, p.h - 1);
  if (y_hi <= y_lo) return;
  dim3 b(32, 8), g(ceil_div(ceil_div(p.w - 1, 4), 32), ceil_div(y_hi - y_lo, 8));
  UBGL_LAUNCH(lc, K_FINISH, 0, stream, k_gradient_save<<<g, b, 0, stream>>>(vx, vy, p, mask, cx, cy, ih, y_lo, y_hi));
}

// ---------------------------------------------------------------------------
// DeviceSim members (single GPU: every row is stored and owned)
// ---------------------------------------------------------------------------
void DeviceSim::fused_prestep() {
  const float a = dt * mu * ((float)W - 1.0f) / pwidth; // simulation.cpp:105
  PrestepArgs g{};
  g.mask = mg->mask0_ptr();
  g.H = H; g.pitch = pitch; g.a = a; g.rden = 1.0f / (1.0f + 4.0f * a);
  g.bcLo = bcW; g.bcHi = bcE; g.bcS = bcS; g.bcN = bcN;
  g.st_lo = 0; g.st_hi = H; g.own_lo = 0; g.own_hi = H;
  // vx: A = front, B = back, Cout = the vx_current buffer; afterwards the roles
  // rotate: front <- Cout, back stays, vx_current <- A (dead until save)
  g.A = vxb[ixf].d; g.K = vxb[ixb].d; g.acc = vx_accum.d; g.B = vxb[ixb].d; g.Cout = vxb[ixc].d;
  g.gw = W - 1; g.gh = H;
  launch_prestep(0, g, stream, &lc);
  std::swap(ixf, ixc);
  g.A = vyb[iyf].d; g.K = vyb[iyf].d; g.acc = vy_accum.d; g.B = vyb[iyb].d; g.Cout = vyb[iyc].d;
  g.gw = W; g.gh = H - 1;
  launch_prestep(1, g, stream, &lc);
  std::swap(iyf, iyc);
}

void DeviceSim::fused_borders(bool with_p, bool with_current) {
  BorderArgs g{};
  g.xf = vxb[ixf]; g.xb = vxb[ixb]; g.yf = vyb[iyf]; g.yb = vyb[iyb];
  if (with_current) {
    g.xc = vxb[ixc];
    g.yc = vyb[iyc];
  }
  if (with_p) g.p = p;
  g.bcW = bcW; g.bcE = bcE; g.bcN = bcN; g.bcS = bcS;
  g.y_lo = 0; g.y_hi = H; g.do_s = 1; g.do_n = 1;
  launch_borders(g, stream, &lc);
}

void DeviceSim::fused_divergence(bool zero_accum) {
  const Grid none{};
  launch_divergence4(vxb[ixf], vyb[iyf], f, zero_accum ? vx_accum : none, zero_accum ? vy_accum : none,
                     1.0f / h, 1, H - 1, stream, &lc);
}

void DeviceSim::fused_gradient_save() {
  const Grid none{};
  launch_gradient_save(vxb[ixf], vyb[iyf], p, mg->mask0_ptr(), lazy_current ? none : vxb[ixc],
                       lazy_current ? none : vyb[iyc], 1.0f / h, 1, H - 1, stream, &lc);
}

} // namespace ubgl
