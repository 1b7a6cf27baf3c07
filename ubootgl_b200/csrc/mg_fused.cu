// mg_fused.cu -- temporally blocked multigrid tile kernels (sm_100a).
//
// One CTA stages a (128 x LH) window of p, f and the 5-bit stencil mask in
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
// Every per-cell formula comes from stencils.cuh, i.e. the same rounding
// sequence as the plain one-kernel-per-operator path in mg.cu: the fused result
// is bit-identical to the plain one (tests/test_gpu_fused.py).
//
// Shared-memory layout: red and black cells live in separate planes
// (plane = (x+y)&1, column index x>>1) so that a half-sweep, which touches every
// other cell, reads and writes consecutive words (no bank conflicts).
//
// Reference semantics: pressure_solver.cpp:10-24 (smoothingKernel), :35-72
// (canonical red-black order), :91-116 (residual), :118-132 (restrict),
// :134-181 (prolongate, correct), :183-192 (setZeroGradientBC), :201-248
// (solveLevel).
#include "mg.cuh"
#include "stencils.cuh"

namespace ubgl {

enum { MODE_PRE = 0, MODE_POST = 1, MODE_SMOOTH = 2 };

struct TileArgs {
  const float *p_in; // nullptr: initial guess is 0 (levels >= 1 start from ec.fill(0))
  float *p_out;
  const float *f;
  const uint8_t *mask;
  int w, h, pitch;
  float *rc;          // MODE_PRE: restricted residual (coarse grid)
  const float *ec;    // MODE_POST: coarse error
  const float *flagc; // MODE_POST: coarse flags (fp32, as the reference sums them)
  int wc, hc, pc;
  float hh, ihsq;
  int zgbc;
};

constexpr int LW = 128; // staged window width in cells (one float4 per lane)
constexpr int HW = LW / 2;

template <int S, int MODE> struct TileGeom {
  // halo: 2 cells per sweep, +2 for residual(+1) and restriction(+1); rounded
  // up to a multiple of 4 so that window rows start 16-byte aligned
  static constexpr int NEED = 2 * S + (MODE == MODE_PRE ? 2 : 0);
  static constexpr int HALO = (NEED + 3) / 4 * 4;
  static constexpr int TX = LW - 2 * HALO;
};

__device__ __forceinline__ float bitf(unsigned m, int b) { return (float)((m >> b) & 1u); }

template <int S, int MODE, int LH, int NT>
__global__ void __launch_bounds__(NT) k_mg_tile(TileArgs a) {
  constexpr int HALO = TileGeom<S, MODE>::HALO;
  constexpr int TX = TileGeom<S, MODE>::TX;
  constexpr int TY = LH - 2 * HALO;
  constexpr int NW = NT / 32;
  static_assert(TY > 0 && TY % 2 == 0 && TX % 4 == 0, "tile geometry");

  extern __shared__ __align__(16) unsigned char smem_raw[];
  float(*P)[LH][HW] = reinterpret_cast<float(*)[LH][HW]>(smem_raw);
  float(*F)[LH][HW] = reinterpret_cast<float(*)[LH][HW]>(smem_raw + sizeof(float) * 2 * LH * HW);
  uint8_t(*M)[LW] = reinterpret_cast<uint8_t(*)[LW]>(smem_raw + sizeof(float) * 4 * LH * HW);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
  const int X0 = x0 - HALO, Y0 = y0 - HALO; // both even: local parity == global parity
  const int w = a.w, h = a.h;

  // ---- stage the window: coalesced 128-bit loads, de-interleaved by colour ----
  for (int r = warp; r < LH; r += NW) {
    const int gy = Y0 + r, gx = X0 + 4 * lane;
    float4 pv = make_float4(0.f, 0.f, 0.f, 0.f), fv = pv;
    uchar4 mv = make_uchar4(0, 0, 0, 0);
    if (gy >= 0 && gy < h && gx >= 0 && gx < a.pitch) {
      const size_t o = (size_t)gy * a.pitch + gx;
      if (a.p_in) pv = *reinterpret_cast<const float4 *>(a.p_in + o);
      fv = __ldg(reinterpret_cast<const float4 *>(a.f + o));
      mv = __ldg(reinterpret_cast<const uchar4 *>(a.mask + o));
    }
    const int pr = gy & 1;
    *reinterpret_cast<float2 *>(&P[pr][r][2 * lane]) = make_float2(pv.x, pv.z);
    *reinterpret_cast<float2 *>(&P[pr ^ 1][r][2 * lane]) = make_float2(pv.y, pv.w);
    *reinterpret_cast<float2 *>(&F[pr][r][2 * lane]) = make_float2(fv.x, fv.z);
    *reinterpret_cast<float2 *>(&F[pr ^ 1][r][2 * lane]) = make_float2(fv.y, fv.w);
    *reinterpret_cast<uchar4 *>(&M[r][4 * lane]) = mv;
  }
  __syncthreads();

  // valid region after k half-sweeps (global coordinates, interior only)
  auto lo = [](int origin, int k) { return max(1, origin + k); };
  auto hi = [](int origin, int len, int n, int k) { return min(n - 1, origin + len - k); };

  // one colour of one sweep over the region that is still exact at time k
  auto half_sweep = [&](int k, int cpar) {
    const int gx_lo = lo(X0, k), gx_hi = hi(X0, LW, w, k);
    const int gy_lo = lo(Y0, k), gy_hi = hi(Y0, LH, h, k);
    for (int gy = gy_lo + warp; gy < gy_hi; gy += NW) {
      const int r = gy - Y0;
      const int gxs = gx_lo + (((gx_lo + gy) & 1) ^ cpar);
      const float *po = &P[cpar ^ 1][r][0];
      const float *ps = &P[cpar ^ 1][r - 1][0];
      const float *pn = &P[cpar ^ 1][r + 1][0];
      for (int gx = gxs + 2 * lane; gx < gx_hi; gx += 64) {
        const int lx = gx - X0, xh = lx >> 1, wo = (lx & 1) - 1;
        const unsigned m = M[r][lx];
        const float v = smooth_cell1(po[xh + wo], po[xh + wo + 1], ps[xh], pn[xh], bitf(m, 0),
                                     bitf(m, 1), bitf(m, 2), bitf(m, 3), bitf(m, 4),
                                     fh2_of(F[cpar][r][xh], a.hh));
        P[cpar][r][xh] = v;
      }
    }
  };

  auto cell = [&](int gx, int gy) -> float & {
    const int lx = gx - X0, ly = gy - Y0;
    return P[(lx + ly) & 1][ly][lx >> 1];
  };

  // setZeroGradientBC restricted to the border cells whose interior neighbour is
  // still exact at time k (corners are never touched)
  auto zero_gradient = [&](int k) {
    const int gx_lo = lo(X0, k), gx_hi = hi(X0, LW, w, k);
    const int gy_lo = lo(Y0, k), gy_hi = hi(Y0, LH, h, k);
    const int t = threadIdx.x;
    if (X0 <= 0 && gx_lo == 1)
      for (int gy = gy_lo + t; gy < gy_hi; gy += NT) cell(0, gy) = cell(1, gy);
    if (w - 1 < X0 + LW && gx_hi == w - 1)
      for (int gy = gy_lo + t; gy < gy_hi; gy += NT) cell(w - 1, gy) = cell(w - 2, gy);
    if (Y0 <= 0 && gy_lo == 1)
      for (int gx = gx_lo + t; gx < gx_hi; gx += NT) cell(gx, 0) = cell(gx, 1);
    if (h - 1 < Y0 + LH && gy_hi == h - 1)
      for (int gx = gx_lo + t; gx < gx_hi; gx += NT) cell(gx, h - 1) = cell(gx, h - 2);
  };

  if (MODE == MODE_POST) {
    // prolongate + correct on the whole window (interior cells), colour by
    // colour so that a warp sees a single parity case of prolong_cell
    for (int cpar = 0; cpar < 2; cpar++) {
      const int gx_lo = lo(X0, 0), gx_hi = hi(X0, LW, w, 0);
      const int gy_lo = lo(Y0, 0), gy_hi = hi(Y0, LH, h, 0);
      for (int gy = gy_lo + warp; gy < gy_hi; gy += NW) {
        const int r = gy - Y0;
        const int gxs = gx_lo + (((gx_lo + gy) & 1) ^ cpar);
        for (int gx = gxs + 2 * lane; gx < gx_hi; gx += 64) {
          const int lx = gx - X0, xh = lx >> 1;
          const float e = prolong_cell(a.ec, a.flagc, a.pc, bitf(M[r][lx], 0), gx, gy, w, h);
          P[cpar][r][xh] = __fadd_rn(P[cpar][r][xh], e);
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
    // residual on (tile + 1) in place of f; r = 0 outside the interior
    const int fx_lo = HALO - 1, fx_hi = HALO + TX + 1, fy_lo = HALO - 1, fy_hi = HALO + TY + 1;
    for (int cpar = 0; cpar < 2; cpar++) {
      for (int ly = fy_lo + warp; ly < fy_hi; ly += NW) {
        const int gy = Y0 + ly;
        const int lxs = fx_lo + (((fx_lo + ly) & 1) ^ cpar);
        const float *po = &P[cpar ^ 1][ly][0];
        for (int lx = lxs + 2 * lane; lx < fx_hi; lx += 64) {
          const int gx = X0 + lx, xh = lx >> 1, wo = (lx & 1) - 1;
          float rv = 0.0f;
          if (gx >= 1 && gx < w - 1 && gy >= 1 && gy < h - 1) {
            const unsigned m = M[ly][lx];
            rv = residual_cell(P[cpar][ly][xh], po[xh + wo], po[xh + wo + 1],
                               P[cpar ^ 1][ly - 1][xh], P[cpar ^ 1][ly + 1][xh], bitf(m, 0),
                               bitf(m, 1), bitf(m, 2), bitf(m, 3), bitf(m, 4), F[cpar][ly][xh],
                               a.ihsq);
          }
          F[cpar][ly][xh] = rv;
        }
      }
    }
    __syncthreads();
    // full-weighting restriction of the coarse cells whose fine centre (2xc,2yc)
    // lies in this tile; coarse border = 0 (rc.fill(0.0), pressure_solver.cpp:222)
    const int xc0 = x0 >> 1, yc0 = y0 >> 1;
    for (int j = warp; j < TY / 2; j += NW) {
      const int yc = yc0 + j;
      if (yc >= a.hc) break;
      const int ly = 2 * yc - Y0;
      for (int i = lane; i < TX / 2; i += 32) {
        const int xc = xc0 + i;
        if (xc >= a.wc) break;
        float v = 0.0f;
        if (xc >= 1 && yc >= 1 && xc < a.wc - 1 && yc < a.hc - 1) {
          const int c = (2 * xc - X0) >> 1; // column index of the (even,even) centre
          // rows ly-1 / ly+1: corners in plane 0, middle in plane 1; row ly: the opposite
          v = fw9(F[0][ly - 1][c - 1], F[1][ly - 1][c], F[0][ly - 1][c], F[1][ly][c - 1],
                  F[0][ly][c], F[1][ly][c], F[0][ly + 1][c - 1], F[1][ly + 1][c],
                  F[0][ly + 1][c]);
        }
        a.rc[(size_t)yc * a.pc + xc] = v;
      }
    }
  }

  // ---- write the tile back (p ping-pong buffer), 128-bit stores ----
  for (int ly = HALO + warp; ly < HALO + TY; ly += NW) {
    const int gy = Y0 + ly;
    if (gy >= h) break;
    const int pr = ly & 1;
    for (int q = lane; q < TX / 4; q += 32) {
      const int lx = HALO + 4 * q, gx = X0 + lx;
      if (gx >= w) break;
      const float2 e = *reinterpret_cast<const float2 *>(&P[pr][ly][lx >> 1]);
      const float2 o = *reinterpret_cast<const float2 *>(&P[pr ^ 1][ly][lx >> 1]);
      *reinterpret_cast<float4 *>(a.p_out + (size_t)gy * a.pitch + gx) =
          make_float4(e.x, o.x, e.y, o.y);
    }
  }
}

// 5-bit stencil mask of a flag grid: bit0 = flag(x,y), bit1..4 = W, E, S, N
// neighbour (0 outside the grid).  *nonbinary is raised if any flag is neither
// 0.0 nor 1.0 (then the bit form is not equivalent and the plain path is used).
__global__ void k_make_mask(Grid flag, uint8_t *mask, int *nonbinary) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= flag.pitch || y >= flag.h) return;
  unsigned m = 0;
  if (x < flag.w) {
    auto bit = [&](int xx, int yy) -> unsigned {
      if (xx < 0 || yy < 0 || xx >= flag.w || yy >= flag.h) return 0u;
      return flag.at(xx, yy) != 0.0f ? 1u : 0u;
    };
    float c = flag.at(x, y);
    if (c != 0.0f && c != 1.0f) *nonbinary = 1;
    m = bit(x, y) | (bit(x - 1, y) << 1) | (bit(x + 1, y) << 2) | (bit(x, y - 1) << 3) |
        (bit(x, y + 1) << 4);
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
  constexpr size_t smem = sizeof(float) * 4 * LH * HW + (size_t)LH * LW;
  static bool attr_set = false;
  if (!attr_set) {
    UBGL_CUDA(cudaFuncSetAttribute(k_mg_tile<S, MODE, LH, NT>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  dim3 grid(ceil_div(a.w, TX), ceil_div(a.h, TY));
  UBGL_LAUNCH(lc, kind, level, stream, k_mg_tile<S, MODE, LH, NT><<<grid, NT, smem, stream>>>(a));
}

constexpr int LH_MAIN = 64, NT_MAIN = 256;

void launch_mg_pre(const float *p_in, float *p_out, const Grid &f, const uint8_t *mask,
                   const Grid &rc, float hh, bool zgbc, cudaStream_t stream, LaunchCounter *lc,
                   int level) {
  TileArgs a{};
  a.p_in = p_in; a.p_out = p_out; a.f = f.d; a.mask = mask;
  a.w = f.w; a.h = f.h; a.pitch = f.pitch;
  a.rc = rc.d; a.wc = rc.w; a.hc = rc.h; a.pc = rc.pitch;
  a.hh = hh; a.ihsq = 1.0f / hh / hh; a.zgbc = zgbc ? 1 : 0;
  launch_tile<3, MODE_PRE, LH_MAIN, NT_MAIN>(a, stream, lc, K_MG_PRE, level);
}

void launch_mg_post(const float *p_in, float *p_out, const Grid &f, const uint8_t *mask,
                    const Grid &ec, const Grid &flagc, float hh, bool zgbc, cudaStream_t stream,
                    LaunchCounter *lc, int level) {
  TileArgs a{};
  a.p_in = p_in; a.p_out = p_out; a.f = f.d; a.mask = mask;
  a.w = f.w; a.h = f.h; a.pitch = f.pitch;
  a.ec = ec.d; a.flagc = flagc.d; a.wc = ec.w; a.hc = ec.h; a.pc = ec.pitch;
  a.hh = hh; a.ihsq = 1.0f / hh / hh; a.zgbc = zgbc ? 1 : 0;
  launch_tile<3, MODE_POST, LH_MAIN, NT_MAIN>(a, stream, lc, K_MG_POST, level);
}

void launch_mg_smooth5(float *p_out, const Grid &f, const uint8_t *mask, float hh,
                       cudaStream_t stream, LaunchCounter *lc, int level) {
  TileArgs a{};
  a.p_in = nullptr; a.p_out = p_out; a.f = f.d; a.mask = mask;
  a.w = f.w; a.h = f.h; a.pitch = f.pitch;
  a.hh = hh; a.ihsq = 1.0f / hh / hh; a.zgbc = 0;
  launch_tile<5, MODE_SMOOTH, LH_MAIN, NT_MAIN>(a, stream, lc, K_MG_COARSE, level);
}

void launch_make_mask(const Grid &flag, uint8_t *mask, int *d_nonbinary, cudaStream_t stream,
                      LaunchCounter *lc, int level) {
  dim3 b(32, 8), g(ceil_div(flag.pitch, 32), ceil_div(flag.h, 8));
  UBGL_LAUNCH(lc, K_COARSEN, level, stream, k_make_mask<<<g, b, 0, stream>>>(flag, mask, d_nonbinary));
}

} // namespace ubgl
