// next.cu -- the callers either side of Simulation::step, on the device
// (SURVEY.md 8f): fluid tracers + the co-located velocity texture they sample
// (GLSL compute shaders in the reference), simple floating items with their
// force scatter into vx_accum / vy_accum, and terrain edits / scrolling of the
// resident fields.  Compiled with --fmad=false (see ubootgl_b200/Makefile): this
// file is gather/scatter and branch bound, and un-contracted arithmetic keeps the
// kernels operation-for-operation comparable with the scalar CPU checker.
//
// All citations are file:line under te42kyfo/ubootgl.
#include "capi_internal.cuh"
#include "common.cuh"
#include "sim.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace ubgl {
namespace {

// ---------------------------------------------------------------------------
// Texture sampling as the reference's shaders see it: GL_LINEAR magnification,
// GL_REPEAT wrap (the texture-object defaults; velocity_textures.cpp never sets
// a sampler parameter), level 0.  u = s*w - 1/2, i0 = floor(u), alpha = frac(u).
// ---------------------------------------------------------------------------
__device__ __forceinline__ int wrap_repeat(int i, int n) {
  int m = i % n;
  return m < 0 ? m + n : m;
}

struct LinearTaps {
  int i0, i1, j0, j1;
  float w00, w10, w01, w11;
};
__device__ __forceinline__ LinearTaps linear_taps(int w, int h, float s, float t) {
  const float u = s * (float)w - 0.5f, v = t * (float)h - 0.5f;
  const float fu = floorf(u), fv = floorf(v);
  const float a = u - fu, b = v - fv;
  LinearTaps q;
  q.i0 = wrap_repeat((int)fu, w);
  q.i1 = wrap_repeat((int)fu + 1, w);
  q.j0 = wrap_repeat((int)fv, h);
  q.j1 = wrap_repeat((int)fv + 1, h);
  q.w00 = (1.0f - a) * (1.0f - b);
  q.w10 = a * (1.0f - b);
  q.w01 = (1.0f - a) * b;
  q.w11 = a * b;
  return q;
}
__device__ __forceinline__ float blend(const LinearTaps &q, float t00, float t10, float t01, float t11) {
  return ((q.w00 * t00 + q.w10 * t10) + q.w01 * t01) + q.w11 * t11;
}
// R32F texture stored as a pitched grid
__device__ __forceinline__ float tex_r32f(const Grid &g, float s, float t) {
  const LinearTaps q = linear_taps(g.w, g.h, s, t);
  return blend(q, __ldg(&g.at(q.i0, q.j0)), __ldg(&g.at(q.i1, q.j0)), __ldg(&g.at(q.i0, q.j1)),
               __ldg(&g.at(q.i1, q.j1)));
}

// One texel of the co-located velocity texture, interp_shader.cs:15-35: vx is
// the (nx-1) x ny staggered texture, vy nx x (ny-1); the output texture is
// (2nx-1) x (2ny-1).
__device__ __forceinline__ float2 vxy_texel(const Grid &vx, const Grid &vy, int nx, int ny, int gx, int gy) {
  const float sx = (float)gx / (2.0f * (float)nx - 2.0f);
  const float sy = (float)(gy + 1) / (2.0f * (float)ny);
  const float tx = (float)(gx + 1) / (2.0f * (float)nx);
  const float ty = (float)gy / (2.0f * (float)ny - 2.0f);
  return make_float2(tex_r32f(vx, sx, sy), tex_r32f(vy, tx, ty));
}

// texture(tex_vxy, st): the four RG32F texels are evaluated on the fly from the
// staggered fields instead of being read from a materialised (2nx-1)x(2ny-1)
// texture -- 36 B/cell of HBM traffic per frame saved; a tracer touches 8 texels.
__device__ __forceinline__ float2 vxy_sample(const Grid &vx, const Grid &vy, int nx, int ny, float s, float t) {
  const LinearTaps q = linear_taps(2 * nx - 1, 2 * ny - 1, s, t);
  const float2 t00 = vxy_texel(vx, vy, nx, ny, q.i0, q.j0), t10 = vxy_texel(vx, vy, nx, ny, q.i1, q.j0);
  const float2 t01 = vxy_texel(vx, vy, nx, ny, q.i0, q.j1), t11 = vxy_texel(vx, vy, nx, ny, q.i1, q.j1);
  return make_float2(blend(q, t00.x, t10.x, t01.x, t11.x), blend(q, t00.y, t10.y, t01.y, t11.y));
}

// interp_shader.cs as a kernel (the materialised texture, for display/interop
// consumers and for parity of the on-the-fly path above)
__global__ void k_colocate(Grid vx, Grid vy, int nx, int ny, float2 *vxy, float *mag) {
  const int gx = blockIdx.x * blockDim.x + threadIdx.x, gy = blockIdx.y * blockDim.y + threadIdx.y;
  const int tw = 2 * nx - 1, th = 2 * ny - 1;
  if (gx >= tw || gy >= th) return;
  const float2 v = vxy_texel(vx, vy, nx, ny, gx, gy);
  vxy[(size_t)gy * tw + gx] = v;
  if (mag) mag[(size_t)gy * tw + gx] = sqrtf(v.x * v.x + v.y * v.y);
}

// ---------------------------------------------------------------------------
// Display export.  The reference crosses PCIe three times per frame for the
// display: vx_current / vy_current into two R32F textures that interp_shader.cs turns into the
// RG32F velocity and R32F magnitude textures (velocity_textures.cpp:63-93), and p into a fresh
// R32F texture (draw.cpp:101 -> draw_2dbuf.cpp:181-209).  Here the same texels are written from
// the resident fields straight into CUDA arrays -- which is what a GL texture registered with
// cudaGraphicsGLRegisterImage is once mapped (cudaGraphicsSubResourceGetMappedArray) -- by surface
// stores; nothing leaves the device.  Mip levels stay the caller's glGenerateMipmap (GPU side).
// ---------------------------------------------------------------------------
__global__ void k_export_vxy(Grid vx, Grid vy, int nx, int ny, cudaSurfaceObject_t s_vxy, cudaSurfaceObject_t s_mag) {
  const int gx = blockIdx.x * blockDim.x + threadIdx.x, gy = blockIdx.y * blockDim.y + threadIdx.y;
  if (gx >= 2 * nx - 1 || gy >= 2 * ny - 1) return;
  const float2 v = vxy_texel(vx, vy, nx, ny, gx, gy);
  if (s_vxy) surf2Dwrite(v, s_vxy, gx * (int)sizeof(float2), gy);
  if (s_mag) surf2Dwrite(sqrtf(v.x * v.x + v.y * v.y), s_mag, gx * (int)sizeof(float), gy);
}
__global__ void k_export_scalar(Grid g, cudaSurfaceObject_t s) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= g.w || y >= g.h) return;
  surf2Dwrite(g.at(x, y), s, x * (int)sizeof(float), y);
}

namespace {
struct Surface { // surface object over a caller-owned array, checked against the expected texel layout
  cudaSurfaceObject_t obj = 0;
  Surface(void *array, int w, int h, int channels, const char *what) {
    if (!array) return;
    cudaChannelFormatDesc d{};
    cudaExtent e{};
    unsigned flags = 0;
    UBGL_CUDA(cudaArrayGetInfo(&d, &e, &flags, (cudaArray_t)array));
    const bool fmt = d.f == cudaChannelFormatKindFloat && d.x == 32 && (channels == 2 ? d.y == 32 : d.y == 0) &&
                     d.z == 0 && d.w == 0;
    UBGL_REQUIRE(fmt && (int)e.width == w && (int)e.height == h, what);
    cudaResourceDesc r{};
    r.resType = cudaResourceTypeArray;
    r.res.array.array = (cudaArray_t)array;
    UBGL_CUDA(cudaCreateSurfaceObject(&obj, &r));
  }
  ~Surface() {
    if (obj) cudaDestroySurfaceObject(obj);
  }
  Surface(const Surface &) = delete;
  Surface &operator=(const Surface &) = delete;
};
} // namespace

__device__ __forceinline__ unsigned wang_hash(unsigned seed) { // advect_tracer_points.cs:20-27
  seed = (seed ^ 61u) ^ (seed >> 16);
  seed *= 9u;
  seed = seed ^ (seed >> 4);
  seed *= 0x27d4eb2du;
  seed = seed ^ (seed >> 15);
  return seed;
}

// advect_tracer_points.cs:42-82, one thread per tracer
__global__ void k_tracers_advect(float2 *points, unsigned *start, unsigned *end, float *ages, int ntracers,
                                 int npoints, float dt, float pdx, float pdy, unsigned rand_seed, Grid vx,
                                 Grid vy, int nx, int ny, Grid flagtex) {
  const unsigned gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (unsigned)ntracers) return;
  unsigned rng = wang_hash(gid + rand_seed);
  const size_t base = (size_t)gid * npoints;
  const unsigned e = end[gid];
  const unsigned curr = e % (unsigned)npoints, next = (e + 1u) % (unsigned)npoints;
  const float2 c = points[base + curr];
  const float sx = c.x / pdx, sy = c.y / pdy;
  const float2 v1 = vxy_sample(vx, vy, nx, ny, sx, sy);
  const float mx = c.x + (v1.x * dt) * 0.5f, my = c.y + (v1.y * dt) * 0.5f;
  const float2 v2 = vxy_sample(vx, vy, nx, ny, mx / pdx, my / pdy);
  float2 n = make_float2(c.x + v2.x * dt, c.y + v2.y * dt);
  float age = ages[gid];
  if (c.x < 0.0f || c.y < 0.0f || c.x > pdx || c.y > pdy || tex_r32f(flagtex, sx, sy) < 0.6f) {
    n = c;
    age += 0.1f;
  }
  if (age > 2.0f * 3.141f) {
    start[gid] = 0;
    end[gid] = 0;
    rng = 1664525u * rng + 1013904223u;
    n.x = ((float)(rng % 100000u) / 100000.0f) * pdx;
    rng = 1664525u * rng + 1013904223u;
    n.y = ((float)(rng % 100000u) / 100000.0f) * pdy;
    points[base] = n;
    ages[gid] = 0.0f;
  } else {
    end[gid] = next;
    const unsigned s = start[gid];
    if (s == next) start[gid] = (s + 1u) % (unsigned)npoints;
    points[base + next] = n;
    ages[gid] = age + 0.02f;
  }
}

__global__ void k_tracers_shift(float2 *points, size_t n, float shift) { // shift_tracers.cs:18-27
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n) points[g].x += shift;
}

// ---------------------------------------------------------------------------
// floating items (advect_floating_items.cpp:148-274)
// ---------------------------------------------------------------------------
struct Item { // == ubgl_item (include/ubgl.h) == CoItem + CoKinematicsSimple
  float size[2], pos[2], rotation;
  float mass, vel[2], force[2], angVel, angForce;
  int bumpCount;
};
static_assert(sizeof(Item) == 52, "ubgl_item layout");

__global__ void k_items_bin(const Item *items, int n, unsigned char *bin, int *idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // ((int)(pos.x * 100)) % bins.size() with bins.size() an unsigned long (:157)
  bin[i] = (unsigned char)((unsigned long long)(long long)(int)(items[i].pos[0] * 100.0f) % 100ull);
  idx[i] = i;
}
// off[b] = first sorted slot of bin b (off[100] = n); empty bins inherit the next start
__global__ void k_items_offsets(const unsigned char *sbin, int n, int *off) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int b = sbin[k], pb = k ? sbin[k - 1] : -1;
  for (int q = pb + 1; q <= b; q++) off[q] = k;
  if (k == n - 1)
    for (int q = b + 1; q <= 100; q++) off[q] = n;
}
__global__ void k_items_gather_pos(const Item *items, const int *order, int n, float2 *spos) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) spos[k] = make_float2(items[order[k]].pos[0], items[order[k]].pos[1]);
}

__device__ __forceinline__ float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }

// Simulation::psampleFlagLinear, simulation.cpp:398-412
__device__ __forceinline__ float psample_flag_linear(const Grid &flag, float pwidth, float px, float py) {
  const float s = pwidth / (float)flag.w;
  const float cx = px / s - 0.5f, cy = py / s - 0.5f;
  int ix = min((int)cx, flag.w - 2), iy = min((int)cy, flag.h - 2);
  ix = max(ix, 0);
  iy = max(iy, 0);
  const float sx = cx - floorf(cx), sy = cy - floorf(cy);
  const float p01 = __ldg(&flag.at(ix, iy + 1)), p11 = __ldg(&flag.at(ix + 1, iy + 1));
  const float p00 = __ldg(&flag.at(ix, iy)), p10 = __ldg(&flag.at(ix + 1, iy));
  return mixf(mixf(p00, p10, sx), mixf(p01, p11, sx), sy);
}
// bilinearSample, interpolators.hpp:11-27
__device__ __forceinline__ float bilinear_sample(const Grid &g, float cx, float cy) {
  cx = fminf(fmaxf(cx, 0.0f), (float)g.w - 1.1f);
  cy = fminf(fmaxf(cy, 0.0f), (float)g.h - 1.1f);
  const int ix = (int)cx, iy = (int)cy;
  const float sx = cx - floorf(cx), sy = cy - floorf(cy);
  const float v1 = __ldg(&g.at(ix, iy)), v2 = __ldg(&g.at(ix + 1, iy));
  const float v3 = __ldg(&g.at(ix, iy + 1)), v4 = __ldg(&g.at(ix + 1, iy + 1));
  const float vm1 = v1 + (v2 - v1) * sx, vm2 = v3 + (v4 - v3) * sx;
  return vm1 + (vm2 - vm1) * sy;
}
// bilinearScatter, interpolators.hpp:29-40 -- the reference runs its items
// serially under accum_mutex; here every += is an atomicAdd (the sum is the same
// up to fp32 association)
__device__ __forceinline__ void bilinear_scatter(const Grid &g, float cx, float cy, float v) {
  cx = fminf(fmaxf(cx, 0.0f), (float)g.w - 1.1f);
  cy = fminf(fmaxf(cy, 0.0f), (float)g.h - 1.1f);
  const int ix = (int)cx, iy = (int)cy;
  const float sx = cx - floorf(cx), sy = cy - floorf(cy);
  const float ax = 1.0f - sx, ay = 1.0f - sy;
  atomicAdd(&g.at(ix + 1, iy + 1), sx * sy * v);
  atomicAdd(&g.at(ix, iy + 1), ax * sy * v);
  atomicAdd(&g.at(ix + 1, iy), sx * ay * v);
  atomicAdd(&g.at(ix, iy), ax * ay * v);
}

// ---- two-level bins (UBGL_ITEMS_VARIANT = 2, default) -----------------------------------
// The reference tests an item against EVERY position of its x-bin (:167-181): O(N^2 / 100),
// 10^10 pair tests for 10^6 items, 65 % of the round-1 explosion frame.  Only pairs closer than
// sqrt(size.x size.y 0.4) of the testing item contribute, so the x-bins are cut into y-cells at
// least as tall as the largest such radius: an item then needs the cells y-1, y, y+1 of its
// x-bin only.  Sort key = (x-bin, y-cell); the stable sort keeps array order inside a cell, and
// the three cells are walked as a 3-way MERGE by array index, so the contacts of an item are
// still added in the reference's push_back order and the sums round as in the reference.
constexpr int ITEMS_NY = 2048;
// r2max = max over the items of size.x * size.y * 0.4 (positive floats order like their bits)
__global__ void k_items_r2max(const Item *items, int n, unsigned *r2max_bits) {
  float m = 0.0f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    m = fmaxf(m, items[i].size[0] * items[i].size[1] * 0.4f);
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(r2max_bits, __float_as_uint(m));
}
// height of a y-cell: the largest interaction radius, but no more than ITEMS_NY cells over the domain
__device__ __forceinline__ float items_cell_height(unsigned r2max_bits, float yrange) {
  return fmaxf(sqrtf(__uint_as_float(r2max_bits)) * 1.0001f, yrange / (float)ITEMS_NY);
}
__device__ __forceinline__ int items_ycell(float y, float cy) {
  const float c = floorf(y / cy);
  return (int)fminf(fmaxf(c, 0.0f), (float)(ITEMS_NY - 1)); // monotone in y: |dy| < cy  =>  cells differ by <= 1
}
__global__ void k_items_key2(const Item *items, int n, const unsigned *r2max_bits, float yrange, unsigned *key, int *idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float cy = items_cell_height(*r2max_bits, yrange);
  const unsigned xb = (unsigned)((unsigned long long)(long long)(int)(items[i].pos[0] * 100.0f) % 100ull); // :157
  key[i] = xb * ITEMS_NY + (unsigned)items_ycell(items[i].pos[1], cy);
  idx[i] = i;
}
// off[q] = first sorted slot whose key is >= q, q = 0 .. nkeys (off[nkeys] = n)
__global__ void k_items_offsets2(const unsigned *skey, int n, int nkeys, int *off) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q > nkeys) return;
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (skey[mid] < (unsigned)q) lo = mid + 1; else hi = mid;
  }
  off[q] = lo;
}

// One thread per SORTED slot, so the lanes of a warp share a bin: the repulsion
// loop (:167-181) walks the bin's positions in array order (stable sort == the
// reference's push_back order).  It is the O(N^2/100) part of the reference
// algorithm (10^10 pair tests for 10^6 items), so the positions of the bin(s) a
// block touches are streamed through shared memory in tiles and every thread
// reads them as warp-uniform broadcasts; each thread still visits exactly its own
// bin's entries, in order, so the sums round as before.
constexpr int ITEMS_NT = 128, ITEMS_TILE = 512;
template <int VARIANT>
__global__ void __launch_bounds__(ITEMS_NT) k_items_advect(Item *items, const int *order, const unsigned char *sbin,
                                                           const unsigned *skey, const int *off, const float2 *spos,
                                                           int n, float game_dt, Grid flag, Grid vx, Grid vy, Grid p,
                                                           Grid ax, Grid ay, float pwidth, float h) {
  __shared__ float2 tile[VARIANT == 1 ? ITEMS_TILE : 1];
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = k < n;
  Item it;
  int qs = 0, qe = 0;
  if (live) {
    it = items[order[k]];
    if (VARIANT == 1) {
      const int b = sbin[k];
      qs = off[b];
      qe = off[b + 1];
    }
  }
  float rfx = 0.0f, rfy = 0.0f;
  int contacts = 0;
  const float r2 = live ? it.size[0] * it.size[1] * 0.4f : 0.0f, lmin = live ? 0.1f * it.size[0] : 0.0f;
  const float px0 = live ? it.pos[0] : 0.0f, py0 = live ? it.pos[1] : 0.0f;
  if (VARIANT == 2) {
    if (live) {
      // cells (xb, yc-1), (xb, yc), (xb, yc+1): three runs of sorted slots, each in array order
      const unsigned key = skey[k];
      const int yc = (int)(key % ITEMS_NY), kb = (int)(key - yc);
      int cur[3], end[3], val[3];
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const int y = yc - 1 + c;
        const bool ok = y >= 0 && y < ITEMS_NY;
        cur[c] = ok ? off[kb + y] : 0;
        end[c] = ok ? off[kb + y + 1] : 0;
        val[c] = cur[c] < end[c] ? order[cur[c]] : 0x7fffffff;
      }
      while (true) {
        const int m = min(val[0], min(val[1], val[2]));
        if (m == 0x7fffffff) break;
        const int c = val[0] == m ? 0 : (val[1] == m ? 1 : 2);
        const int j = c == 0 ? cur[0] : (c == 1 ? cur[1] : cur[2]);
        const float2 o = __ldg(&spos[j]);
        const float dx = px0 - o.x, dy = py0 - o.y;
        const float d2 = dx * dx + dy * dy;
        if (d2 < r2) { // :170-180
          const float len = fmaxf(lmin, sqrtf(d2));
          rfx += 0.0001f * (dx / len / len);
          rfy += 0.0001f * (dy / len / len);
          contacts++;
        }
        const int nj = j + 1;
#pragma unroll
        for (int q = 0; q < 3; q++)
          if (q == c) {
            cur[q] = nj;
            val[q] = nj < end[q] ? order[nj] : 0x7fffffff;
          }
      }
    }
  } else {
    const int kf = blockIdx.x * blockDim.x, kl = min(kf + (int)blockDim.x, n) - 1;
    const int qlo = off[sbin[kf]], qhi = off[sbin[kl] + 1]; // slots any thread of this block visits
    for (int base = qlo; base < qhi; base += ITEMS_TILE) {
      __syncthreads();
      for (int t = threadIdx.x; t < ITEMS_TILE; t += ITEMS_NT)
        tile[t] = base + t < qhi ? __ldg(&spos[base + t]) : make_float2(0.0f, 0.0f);
      __syncthreads();
      const int j1 = min(qe, base + ITEMS_TILE) - base;
      int j = max(qs, base) - base;
      auto hit = [&](float dx, float dy, float d2) { // :170-180, taken for a handful of the pairs
        const float len = fmaxf(lmin, sqrtf(d2));
        rfx += 0.0001f * (dx / len / len);
        rfy += 0.0001f * (dy / len / len);
        contacts++;
      };
      // 8 pair tests at a time, branch-free; the (rare) contacts of a batch are then applied in
      // array order, so the sums round exactly as in the one-by-one loop
      for (; j + 8 <= j1; j += 8) {
        float dx[8], dy[8], d2[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const float2 o = tile[j + k];
          dx[k] = px0 - o.x;
          dy[k] = py0 - o.y;
          d2[k] = dx[k] * dx[k] + dy[k] * dy[k];
        }
        const float m = fminf(fminf(fminf(d2[0], d2[1]), fminf(d2[2], d2[3])),
                              fminf(fminf(d2[4], d2[5]), fminf(d2[6], d2[7])));
        if (m < r2) {
#pragma unroll
          for (int k = 0; k < 8; k++)
            if (d2[k] < r2) hit(dx[k], dy[k], d2[k]);
        }
      }
      for (; j < j1; j++) {
        const float2 o = tile[j];
        const float dx = px0 - o.x, dy = py0 - o.y;
        const float d2 = dx * dx + dy * dy;
        if (d2 < r2) hit(dx, dy, d2);
      }
    }
  }
  if (!live) return;
  const float cden = (float)max(contacts, 1);
  rfx /= cden;
  rfy /= cden;
  it.force[0] += rfx * 10.0f;
  it.force[1] += rfy * 10.0f;

  const int W = flag.w, H = flag.h;
  const int steps = (int)fmin(15.0, fmax(1.0, (double)(fmaxf(fabsf(it.vel[0]), fabsf(it.vel[1])) * game_dt / h) * 2.5));
  const float sub = game_dt / (float)steps;
  const float gs = pwidth / (float)W;
  for (int s = 0; s < steps; s++) {
    it.pos[0] += sub * it.vel[0];
    it.pos[1] += sub * it.vel[1];
    it.rotation = (float)fmod((double)(it.rotation + sub * it.angVel) + 2 * M_PI, 2 * M_PI);
    const float gpx = it.pos[0] / gs, gpy = it.pos[1] / gs;
    if (gpx >= (float)(W - 2) || gpx <= 1.0f || gpy >= (float)(H - 2) || gpy <= 1.0f) continue;
    if (psample_flag_linear(flag, pwidth, it.pos[0], it.pos[1]) < 1.0f) {
      const float px = it.pos[0], py = it.pos[1];
      const float p01 = psample_flag_linear(flag, pwidth, px - h, py + h);
      const float p11 = psample_flag_linear(flag, pwidth, px + h, py + h);
      const float p00 = psample_flag_linear(flag, pwidth, px - h, py - h);
      const float p10 = psample_flag_linear(flag, pwidth, px + h, py - h);
      float nx = p11 + p10 - p01 - p00, ny = p01 + p11 - p00 - p10; // psampleFlagNormal
      const float nl = sqrtf(nx * nx + ny * ny);
      if (nl > 0.0f) {
        nx /= nl;
        ny /= nl;
        float d = it.vel[0] * nx + it.vel[1] * ny;
        if (d < 0.0f) { // reflect(v, n) * 0.7
          d = nx * it.vel[0] + ny * it.vel[1];
          it.vel[0] = (it.vel[0] - nx * d * 2.0f) * 0.7f;
          it.vel[1] = (it.vel[1] - ny * d * 2.0f) * 0.7f;
        }
        d = it.force[0] * nx + it.force[1] * ny;
        if (d < 0.0f) {
          d = nx * it.force[0] + ny * it.force[1];
          it.force[0] = (it.force[0] - nx * d * 2.0f) * 0.7f;
          it.force[1] = (it.force[1] - ny * d * 2.0f) * 0.7f;
        }
        const float ang = 0.5f * (float)M_PI;
        const float ca = cosf(ang), sa = sinf(ang);
        const float lx = nx * ca - ny * sa, ly = nx * sa + ny * ca; // glm::rotate(n, pi/2)
        const float lat_vel = lx * it.vel[0] + ly * it.vel[1];
        const float rot_vel = it.angVel * (it.size[0] + it.size[1]) * 0.5f;
        const float lat_diff = lat_vel - rot_vel;
        it.force[0] += lat_diff * lx * 1.0f;
        it.force[1] += lat_diff * ly * 1.0f;
      } else {
        it.vel[0] = 0.0f;
        it.vel[1] = 0.0f;
      }
      it.bumpCount++;
    }
    float efx = 0.0f * it.mass + it.force[0], efy = -0.5f * it.mass + it.force[1];
    const float gx = it.pos[0] / h, gy = it.pos[1] / h;
    const float dvx = bilinear_sample(vx, gx - 0.5f, gy) - it.vel[0]; // bilinearVel :11-14
    const float dvy = bilinear_sample(vy, gx, gy - 0.5f) - it.vel[1];
    const float drag = 2000.0f * (it.size[0] + it.size[1]);
    efx += drag * dvx;
    efy += drag * dvy;
    if (gx > 1.0f && gx < (float)vx.w - 2.0f && gy < 1.0f && gy < (float)vx.h - 2.0f) { // :240-241 (sic)
      const float ddx = dvx * it.size[0] * it.size[1], ddy = dvy * it.size[0] * it.size[1];
      bilinear_scatter(ax, gx - 0.5f, gy, -ddx);
      bilinear_scatter(ay, gx, gy - 0.5f, -ddy);
    }
    it.vel[0] += sub * efx / it.mass;
    it.vel[1] += sub * efy / it.mass;
    const int ix = (int)gpx, iy = (int)gpy;
    const float fluid_ang = -((__ldg(&vx.at(ix, iy)) - __ldg(&vx.at(ix, iy - 1))) -
                              (__ldg(&vy.at(ix, iy)) - __ldg(&vy.at(ix - 1, iy)))) / h / 2.0f;
    const float ang_mass = it.size[0] * it.size[1] * it.mass * (1.0f / 12.0f);
    const float kk = fminf(1.0f, sub / ang_mass * 0.0005f * (it.size[0] + it.size[1]) / 4.0f);
    it.angVel += kk * (fluid_ang - it.angVel);
    it.angVel += sub / ang_mass * it.angForce;
    atomicAdd(&ax.at(ix, iy), -(it.angVel * kk * 0.01f * __ldg(&p.at(ix, iy))));
    atomicAdd(&ax.at(ix, iy - 1), it.angVel * kk * 0.01f * __ldg(&p.at(ix, iy - 1)));
    atomicAdd(&ay.at(ix, iy), it.angVel * kk * 0.01f * __ldg(&p.at(ix, iy)));
    atomicAdd(&ay.at(ix - 1, iy), -(it.angVel * kk * 0.01f * __ldg(&p.at(ix - 1, iy))));
  }
  it.angForce = 0.0f;
  it.force[0] = 0.0f;
  it.force[1] = 0.0f;
  items[order[k]] = it;
}

// Simulation::advectFloatingItems (advect_floating_items.cpp:16-146): the rigid rectangular
// bodies (CoItem + CoKinematics -- submarines, torpedoes).  Bodies do not interact, so it is
// one thread per body; per sub-step five terrain probes (each up to six bilinear flag samples),
// then drag sampled at max(2, side/h) points along each of the four sides with the reaction
// scattered into the device accumulators by atomicAdd, like the simple items above.
// glm::rotate(vec2, angle).  The reference's cos / sin are glibc's cosf / sinf, correctly rounded
// for practically every argument; CUDA's cosf / sinf are good to 1-2 ulp, and a last-bit difference
// in a probe position flips `psampleFlagLinear(...) < 0.5` for bodies that sit exactly on a
// terrain edge.  There are only a few thousand rigid bodies, so the rotation takes the
// double-precision functions and rounds once: the float the reference computes.
__device__ __forceinline__ void rot2(float x, float y, float ang, float &ox, float &oy) {
  const float c = (float)cos((double)ang), s = (float)sin((double)ang);
  ox = x * c - y * s;
  oy = x * s + y * c;
}
__global__ void __launch_bounds__(64) k_items_advect_rigid(Item *items, int n, float game_dt, Grid flag, Grid vx,
                                                          Grid vy, Grid ax, Grid ay, float pwidth, float h) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  Item it = items[q];
  const int W = flag.w, H = flag.h;
  const float SPX[5] = {1.0f, -1.0f, 1.0f, -1.0f, 0.0f}, SPY[5] = {1.0f, 1.0f, -1.0f, -1.0f, 0.0f}; // :48-50
  const float SFX[4] = {-1.0f, 1.0f, 0.0f, 0.0f}, SFY[4] = {0.0f, 0.0f, -1.0f, 1.0f};               // :80-81
  const int steps = (int)fmin(15.0, fmax(1.0, (double)(fmaxf(fabsf(it.vel[0]), fabsf(it.vel[1])) * game_dt / h) * 2.5));
  const float sub = game_dt / (float)steps;
  const float gs = pwidth / (float)W;
  for (int st = 0; st < steps; st++) {
    const float bx = it.pos[0], by = it.pos[1]; // posBefore
    it.pos[0] += sub * it.vel[0];
    it.pos[1] += sub * it.vel[1];
    it.rotation = (float)fmod((double)(it.rotation + sub * it.angVel) + 2 * M_PI, 2 * M_PI);
    const float gpx = it.pos[0] / gs, gpy = it.pos[1] / gs;
    if (gpx >= (float)(W - 2) || gpx <= 1.0f || gpy >= (float)(H - 2) || gpy <= 1.0f) continue; // :42-45
    it.force[0] += 0.0f * it.mass; // :47
    it.force[1] += -0.5f * it.mass;
#pragma unroll 1
    for (int k = 0; k < 5; k++) { // terrain probes, :52-72
      float spx, spy;
      rot2(SPX[k] * 0.5f * it.size[0], SPY[k] * 0.5f * it.size[1], it.rotation, spx, spy);
      if (psample_flag_linear(flag, pwidth, it.pos[0] + spx, it.pos[1] + spy) < 0.5f) {
        const float mx = 0.5f * (bx + it.pos[0]) + spx, my = 0.5f * (by + it.pos[1]) + spy;
        const float p01 = psample_flag_linear(flag, pwidth, mx - h, my + h);
        const float p11 = psample_flag_linear(flag, pwidth, mx + h, my + h);
        const float p00 = psample_flag_linear(flag, pwidth, mx - h, my - h);
        const float p10 = psample_flag_linear(flag, pwidth, mx + h, my - h);
        float nx = p11 + p10 - p01 - p00, ny = p01 + p11 - p00 - p10; // psampleFlagNormal
        float nl = sqrtf(nx * nx + ny * ny);
        nx /= nl; // normalize(): NaN when the normal vanishes, then the block below is skipped
        ny /= nl;
        if (psample_flag_linear(flag, pwidth, bx + spx, by + spy) > 0.5f) {
          it.pos[0] = bx;
          it.pos[1] = by;
        }
        nl = sqrtf(nx * nx + ny * ny);
        if (nl > 0.0f) {
          nx /= nl;
          ny /= nl;
          if (it.vel[0] * nx + it.vel[1] * ny < 0.0f) { // reflect(v, n) * 0.7
            const float d = nx * it.vel[0] + ny * it.vel[1];
            it.vel[0] = (it.vel[0] - nx * d * 2.0f) * 0.7f;
            it.vel[1] = (it.vel[1] - ny * d * 2.0f) * 0.7f;
          }
          const float fl = psample_flag_linear(flag, pwidth, it.pos[0] + spx, it.pos[1] + spy);
          it.vel[0] += nx * 0.07f * fl;
          it.vel[1] += ny * 0.07f * fl;
          const float df = it.force[0] * nx + it.force[1] * ny;
          if (df < 0.0f) {
            it.force[0] += 1.1f * df * nx;
            it.force[1] += 1.1f * df * ny;
          }
        }
        it.bumpCount++;
      }
    }
    const float efx = 0.0f * it.mass + it.force[0], efy = -0.5f * it.mass + it.force[1]; // :74
    float cfx = 0.0f, cfy = 0.0f;
    const float ang_force = it.angForce;
#pragma unroll 1
    for (int i = 0; i < 4; i++) {
      const float side = i < 2 ? it.size[1] : it.size[0]; // :83
      const int nsp = (int)fmaxf(2.0f, side / h);         // :86
      float ox, oy;
      rot2(SFX[i], SFY[i], it.rotation, ox, oy);
      for (int k = 0; k < nsp; k++) {
        const float tpar = 1.0f - (float)k * 2.0f / (float)(nsp - 1);
        const float sx = SFX[i] + fabsf(SFY[i]) * tpar, sy = SFY[i] + fabsf(SFX[i]) * tpar; // :89-91
        const float lx = sx * it.size[0] * 0.5f, ly = sy * it.size[1] * 0.5f;
        float tx, ty, rx, ry, nx, ny;
        rot2(lx, ly, it.rotation, tx, ty);
        tx += it.pos[0];
        ty += it.pos[1];
        const float gx = tx / h, gy = ty / h;
        rot2(lx, ly, it.rotation + 0.5f * 3.141f, rx, ry);
        const float dvx = bilinear_sample(vx, gx - 0.5f, gy) - (it.vel[0] + 3.141f * rx * it.angVel);
        const float dvy = bilinear_sample(vy, gx, gy - 0.5f) - (it.vel[1] + 3.141f * ry * it.angVel);
        const float sl = sqrtf(sx * sx + sy * sy);
        rot2(sx / sl, sy / sl, it.rotation, nx, ny); // rotate(normalize(sp), rotation)
        const float pr = fminf(0.0f, dvx * ox + dvy * oy);
        const float fx = nx * pr, fy = ny * pr;
        cfx += fx * 400000.0f * (0.003f + side) * side / (float)nsp; // :110-111
        cfy += fy * 400000.0f * (0.003f + side) * side / (float)nsp;
        if (gx < 1.0f || gx > (float)vx.w - 2.0f || gy < 1.0f || gy > (float)vx.h - 2.0f) continue; // :113-115
        const float ddx = fx * (0.003f + side) * side / (float)nsp * sub * 18000000.0f;
        const float ddy = fy * (0.003f + side) * side / (float)nsp * sub * 18000000.0f;
        bilinear_scatter(ax, gx - 0.5f, gy, -ddx);
        bilinear_scatter(ay, gx, gy - 0.5f, -ddy);
      }
    }
    {
      const float gx = it.pos[0] / h, gy = it.pos[1] / h; // :125-126
      const float kk = 1000.0f * (it.size[0] + it.size[1]);
      cfx += kk * (bilinear_sample(vx, gx - 0.5f, gy) - it.vel[0]);
      cfy += kk * (bilinear_sample(vy, gx, gy - 0.5f) - it.vel[1]);
    }
    it.vel[0] += sub * (efx + cfx) / it.mass; // :129
    it.vel[1] += sub * (efy + cfy) / it.mass;
    const float ang_mass = it.size[0] * it.size[1] * it.mass * (1.0f / 12.0f);
    it.angVel += sub * ang_force / ang_mass;
    it.angVel = (float)((double)it.angVel * 0.98); // :135
  }
  it.angForce = 0.0f;
  it.force[0] = 0.0f;
  it.force[1] = 0.0f;
  items[q] = it;
}

// ---------------------------------------------------------------------------
// terrain edits on the resident simulation-resolution mask
// ---------------------------------------------------------------------------
// Terrain::drawCircle (terrain.cpp:213-234) at terrain scale 1, where flagSimRes
// is flagFullRes thresholded at 0.99 (subSample :3-11): every cell of the
// (2 diam + 1)^2 box that passes the clip test gets (val > 0.99).  One block per
// circle; circles of one call carry the same value, so their order is immaterial.
__global__ void k_draw_circles(Grid flag, const float *xyd, int n, float val) {
  const int c = blockIdx.x;
  const float cx = xyd[3 * c], cy = xyd[3 * c + 1];
  const int diam = (int)xyd[3 * c + 2];
  const int side = 2 * diam + 1;
  const float v = val > 0.99f ? 1.0f : 0.0f;
  for (int t = threadIdx.x; t < side * side; t += blockDim.x) {
    const int x = t % side - diam, y = t / side - diam;
    if (x * x + y * y > diam * diam || cx + (float)x < 0.0f || (float)x + cx > (float)flag.w ||
        (float)y + cy < 2.0f || (float)y + cy >= (float)(flag.h - 3))
      continue;
    flag.at((int)((float)x + cx), (int)((float)y + cy)) = v;
  }
}

// Simulation::setGrids (simulation.hpp:82-98) for every cell with the mask `nf`
// (ubootgl_app.cpp:274-278): flag = nf; around a solid cell the four adjacent
// faces of the FRONT buffers and p are zeroed (all writes are zeros: no race).
// nf == flag.d is allowed (re-apply the resident mask).
__global__ void k_set_grids_all(Grid flag, Grid vx, Grid vy, Grid p, const float *nf, int nf_pitch) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= flag.w || y >= flag.h) return;
  const float v = nf[(size_t)y * nf_pitch + x];
  flag.at(x, y) = v;
  if (v == 0.0f) {
    if (x < vx.w) vx.at(x, y) = 0.0f;
    if (x > 0) vx.at(x - 1, y) = 0.0f;
    if (y < vy.h) vy.at(x, y) = 0.0f;
    if (y > 0) vy.at(x, y - 1) = 0.0f;
    p.at(x, y) = 0.0f;
  }
}

// In-place scroll of one row by one column to the left (ubootgl_app.cpp:254-272,
// terrain.cpp:114-118): dst[x-1] = src[x] for x in [x_first, w).  One block per
// row marches over the row in chunks; a chunk is read completely before it is
// written, and chunk k+1 reads only elements chunk k did not write.
__global__ void k_shift_rows(Grid a, Grid mirror, int x_first, const float *last_col) {
  const int y = blockIdx.x;
  if (y >= a.h) return;
  float *row = &a.at(0, y);
  float *mrow = mirror.d ? &mirror.at(0, y) : nullptr;
  for (int x0 = x_first; x0 < a.w; x0 += blockDim.x) {
    const int x = x0 + threadIdx.x;
    float v = 0.0f;
    if (x < a.w) v = row[x];
    __syncthreads();
    if (x < a.w) {
      row[x - 1] = v;
      if (mrow) mrow[x - 1] = v;
    }
    __syncthreads();
  }
  if (last_col && threadIdx.x == 0) row[a.w - 1] = last_col[y];
}

// inlet column after a scroll (ubootgl_app.cpp:280-289): inletArea = 1 + sum of
// flag(0, y) over y < H-1 (a sum of 0/1 values: exact in any order);
// vx.f(0,y) = vx.b(0,y) = 0.07 * H / inletArea * flag(0,y)
__global__ void k_inlet(Grid flag, Grid vxf, Grid vxb) {
  __shared__ float part[32];
  float acc = 0.0f;
  for (int y = threadIdx.x; y < flag.h - 1; y += blockDim.x) acc += flag.at(0, y);
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.0f;
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) part[0] = 1.0f + acc;
  }
  __syncthreads();
  const float v = 0.07f * (float)flag.h / part[0];
  for (int y = threadIdx.x; y < flag.h; y += blockDim.x) {
    const float q = v * flag.at(0, y);
    vxf.at(0, y) = q;
    vxb.at(0, y) = q;
  }
}

dim3 blk(int x = 32, int y = 8) { return dim3(x, y); }
dim3 grd(int w, int h, int bx = 32, int by = 8) { return dim3(ceil_div(w, bx), ceil_div(h, by)); }

} // namespace
} // namespace ubgl

using namespace ubgl;

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
struct ubgl_tracers {
  int device = 0, ntracers = 0, npoints = 0;
  float2 *points = nullptr;
  unsigned *start = nullptr, *end = nullptr;
  float *ages = nullptr;
  Grid flagtex{}; // optional full-resolution flag texture (velocity_textures.cpp:95-101)
  ~ubgl_tracers() {
    cudaSetDevice(device);
    cudaDeviceSynchronize();
    cudaFree(points); cudaFree(start); cudaFree(end); cudaFree(ages);
    free_grid(flagtex);
  }
};

struct ubgl_items {
  int device = 0, n = 0, cap = 0;
  Item *items = nullptr;
  unsigned char *bin = nullptr, *sbin = nullptr;
  int *idx = nullptr, *order = nullptr, *off = nullptr;
  float2 *spos = nullptr;
  unsigned *key = nullptr, *skey = nullptr, *r2max = nullptr; // two-level bins
  int *off2 = nullptr;
  void *tmp = nullptr;
  size_t tmp_bytes = 0;
  void release() {
    cudaFree(items); cudaFree(bin); cudaFree(sbin); cudaFree(idx); cudaFree(order); cudaFree(off);
    cudaFree(spos); cudaFree(tmp); cudaFree(key); cudaFree(skey); cudaFree(r2max); cudaFree(off2);
    items = nullptr; bin = sbin = nullptr; idx = order = off = nullptr; spos = nullptr; tmp = nullptr;
    key = skey = r2max = nullptr; off2 = nullptr;
    tmp_bytes = 0; cap = 0;
  }
  ~ubgl_items() {
    cudaSetDevice(device);
    cudaDeviceSynchronize();
    release();
  }
};

extern "C" {

int ubgl_sim_colocate_velocity(ubgl_sim_t *sim, float *vxy_host, float *mag_host) {
  UBGL_TRY
  SIM(sim);
  const int tw = 2 * S.W - 1, th = 2 * S.H - 1;
  const size_t n = (size_t)tw * th;
  if (!S.d_vxy) UBGL_CUDA(cudaMalloc(&S.d_vxy, sizeof(float) * 2 * n));
  if (!S.d_mag) UBGL_CUDA(cudaMalloc(&S.d_mag, sizeof(float) * n));
  UBGL_LAUNCH(&S.lc, K_COLOCATE, 0, S.stream,
              k_colocate<<<grd(tw, th), blk(), 0, S.stream>>>(S.field(F_VX_CURRENT), S.field(F_VY_CURRENT), S.W,
                                                               S.H, (float2 *)S.d_vxy, S.d_mag));
  if (vxy_host)
    UBGL_CUDA(cudaMemcpyAsync(vxy_host, S.d_vxy, sizeof(float) * 2 * n, cudaMemcpyDeviceToHost, S.stream));
  if (mag_host)
    UBGL_CUDA(cudaMemcpyAsync(mag_host, S.d_mag, sizeof(float) * n, cudaMemcpyDeviceToHost, S.stream));
  if (vxy_host || mag_host) UBGL_CUDA(cudaStreamSynchronize(S.stream));
  UBGL_CATCH
}

int ubgl_sim_export_display(ubgl_sim_t *sim, void *vxy_array, void *mag_array, void *p_array) {
  UBGL_TRY
  SIM(sim);
  const int tw = 2 * S.W - 1, th = 2 * S.H - 1;
  Surface sv(vxy_array, tw, th, 2, "export_display: vxy must be a (2W-1) x (2H-1) array of 2 x 32-bit float texels (RG32F)");
  Surface sm(mag_array, tw, th, 1, "export_display: mag must be a (2W-1) x (2H-1) array of 32-bit float texels (R32F)");
  Surface sp(p_array, S.W, S.H, 1, "export_display: p must be a W x H array of 32-bit float texels (R32F)");
  if (sv.obj || sm.obj)
    UBGL_LAUNCH(&S.lc, K_COLOCATE, 0, S.stream,
                k_export_vxy<<<grd(tw, th), blk(), 0, S.stream>>>(S.field(F_VX_CURRENT), S.field(F_VY_CURRENT), S.W,
                                                                 S.H, sv.obj, sm.obj));
  if (sp.obj)
    UBGL_LAUNCH(&S.lc, K_COLOCATE, 0, S.stream, k_export_scalar<<<grd(S.W, S.H), blk(), 0, S.stream>>>(S.field(F_P), sp.obj));
  // the surface objects die with this call: the stores must have been issued against live handles
  UBGL_CUDA(cudaStreamSynchronize(S.stream));
  UBGL_CATCH
}

int ubgl_display_array_create(int w, int h, int channels, int device, void **array_out) {
  UBGL_TRY
  NEED(array_out, "array_out");
  *array_out = nullptr;
  UBGL_REQUIRE(w >= 1 && h >= 1 && (channels == 1 || channels == 2), "display_array: w, h >= 1, channels 1 or 2");
  require_device(device);
  UBGL_CUDA(cudaSetDevice(device));
  const cudaChannelFormatDesc d = cudaCreateChannelDesc(32, channels == 2 ? 32 : 0, 0, 0, cudaChannelFormatKindFloat);
  cudaArray_t a = nullptr;
  UBGL_CUDA(cudaMallocArray(&a, &d, (size_t)w, (size_t)h, cudaArraySurfaceLoadStore));
  *array_out = a;
  UBGL_CATCH
}

int ubgl_display_array_read(void *array, float *host) {
  UBGL_TRY
  NEED(array, "array");
  NEED(host, "host");
  cudaChannelFormatDesc d{};
  cudaExtent e{};
  unsigned flags = 0;
  UBGL_CUDA(cudaArrayGetInfo(&d, &e, &flags, (cudaArray_t)array));
  const size_t row = e.width * (size_t)((d.x + d.y + d.z + d.w) / 8);
  UBGL_CUDA(cudaMemcpy2DFromArray(host, row, (cudaArray_t)array, 0, 0, row, e.height, cudaMemcpyDeviceToHost));
  UBGL_CATCH
}

int ubgl_display_array_destroy(void *array) {
  UBGL_TRY
  if (array) UBGL_CUDA(cudaFreeArray((cudaArray_t)array));
  UBGL_CATCH
}

int ubgl_tracers_create(int ntracers, int npoints, int device, ubgl_tracers_t **out) {
  UBGL_TRY
  NEED(out, "out");
  *out = nullptr;
  UBGL_REQUIRE(ntracers >= 1 && npoints >= 1, "tracers: ntracers, npoints must be >= 1");
  require_device(device);
  std::unique_ptr<ubgl_tracers> t(new ubgl_tracers);
  t->device = device;
  t->ntracers = ntracers;
  t->npoints = npoints;
  const size_t np = (size_t)ntracers * npoints;
  UBGL_CUDA(cudaMalloc(&t->points, sizeof(float2) * np));
  UBGL_CUDA(cudaMalloc(&t->start, sizeof(unsigned) * ntracers));
  UBGL_CUDA(cudaMalloc(&t->end, sizeof(unsigned) * ntracers));
  UBGL_CUDA(cudaMalloc(&t->ages, sizeof(float) * ntracers));
  // GLTracers::init, draw_tracers_cs.cpp:42-63: zero points and pointers, ages 2*3.1
  UBGL_CUDA(cudaMemset(t->points, 0, sizeof(float2) * np));
  UBGL_CUDA(cudaMemset(t->start, 0, sizeof(unsigned) * ntracers));
  UBGL_CUDA(cudaMemset(t->end, 0, sizeof(unsigned) * ntracers));
  std::vector<float> a(ntracers, (float)(2 * 3.1));
  UBGL_CUDA(cudaMemcpy(t->ages, a.data(), sizeof(float) * ntracers, cudaMemcpyHostToDevice));
  *out = t.release();
  UBGL_CATCH
}

int ubgl_tracers_destroy(ubgl_tracers_t *t) {
  UBGL_TRY
  delete t;
  UBGL_CATCH
}

#define TRC(t)                                                                 \
  NEED(t, "tracers");                                                          \
  UBGL_CUDA(cudaSetDevice((t)->device));

int ubgl_tracers_upload(ubgl_tracers_t *t, const float *points, const unsigned *start, const unsigned *end,
                        const float *ages) {
  UBGL_TRY
  TRC(t);
  const size_t np = (size_t)t->ntracers * t->npoints;
  if (points) UBGL_CUDA(cudaMemcpy(t->points, points, sizeof(float2) * np, cudaMemcpyHostToDevice));
  if (start) UBGL_CUDA(cudaMemcpy(t->start, start, sizeof(unsigned) * t->ntracers, cudaMemcpyHostToDevice));
  if (end) UBGL_CUDA(cudaMemcpy(t->end, end, sizeof(unsigned) * t->ntracers, cudaMemcpyHostToDevice));
  if (ages) UBGL_CUDA(cudaMemcpy(t->ages, ages, sizeof(float) * t->ntracers, cudaMemcpyHostToDevice));
  UBGL_CATCH
}

int ubgl_tracers_download(ubgl_tracers_t *t, float *points, unsigned *start, unsigned *end, float *ages) {
  UBGL_TRY
  TRC(t);
  UBGL_CUDA(cudaDeviceSynchronize());
  const size_t np = (size_t)t->ntracers * t->npoints;
  if (points) UBGL_CUDA(cudaMemcpy(points, t->points, sizeof(float2) * np, cudaMemcpyDeviceToHost));
  if (start) UBGL_CUDA(cudaMemcpy(start, t->start, sizeof(unsigned) * t->ntracers, cudaMemcpyDeviceToHost));
  if (end) UBGL_CUDA(cudaMemcpy(end, t->end, sizeof(unsigned) * t->ntracers, cudaMemcpyDeviceToHost));
  if (ages) UBGL_CUDA(cudaMemcpy(ages, t->ages, sizeof(float) * t->ntracers, cudaMemcpyDeviceToHost));
  UBGL_CATCH
}

int ubgl_tracers_set_flag_texture(ubgl_tracers_t *t, const float *flag, int w, int h) {
  UBGL_TRY
  TRC(t);
  UBGL_CUDA(cudaDeviceSynchronize());
  free_grid(t->flagtex);
  if (flag) {
    UBGL_REQUIRE(w >= 1 && h >= 1, "flag texture size");
    t->flagtex = alloc_grid(w, h, round_up(w, 32), false);
    upload_grid(t->flagtex, flag, w, h, nullptr);
    UBGL_CUDA(cudaDeviceSynchronize());
  }
  UBGL_CATCH
}

int ubgl_tracers_advect(ubgl_tracers_t *t, ubgl_sim_t *sim, float dt, unsigned rand_seed) {
  UBGL_TRY
  TRC(t);
  SIM(sim);
  UBGL_REQUIRE(S.device == t->device, "tracers and simulation live on different devices");
  const Grid ft = t->flagtex.d ? t->flagtex : S.field(F_FLAG);
  const float pdx = S.pwidth, pdy = S.pwidth * (float)S.H / (float)S.W; // draw_tracers_cs.cpp:143
  UBGL_LAUNCH(&S.lc, K_TRACERS, 0, S.stream,
              k_tracers_advect<<<ceil_div(t->ntracers, 256), 256, 0, S.stream>>>(
                  t->points, t->start, t->end, t->ages, t->ntracers, t->npoints, dt, pdx, pdy, rand_seed,
                  S.field(F_VX_CURRENT), S.field(F_VY_CURRENT), S.W, S.H, ft));
  UBGL_CATCH
}

int ubgl_tracers_shift(ubgl_tracers_t *t, float shift) {
  UBGL_TRY
  TRC(t);
  const size_t n = (size_t)t->ntracers * t->npoints;
  k_tracers_shift<<<(unsigned)((n + 255) / 256), 256>>>(t->points, n, shift);
  UBGL_CHECK_LAUNCH();
  UBGL_CATCH
}

// ---- floating items ---------------------------------------------------------
int ubgl_items_create(int device, ubgl_items_t **out) {
  UBGL_TRY
  NEED(out, "out");
  *out = nullptr;
  require_device(device);
  std::unique_ptr<ubgl_items> h(new ubgl_items);
  h->device = device;
  *out = h.release();
  UBGL_CATCH
}

int ubgl_items_destroy(ubgl_items_t *it) {
  UBGL_TRY
  delete it;
  UBGL_CATCH
}

int ubgl_items_upload(ubgl_items_t *it, const ubgl_item *items, int n) {
  UBGL_TRY
  NEED(it, "items");
  UBGL_REQUIRE(n >= 0 && (n == 0 || items), "bad item list");
  UBGL_CUDA(cudaSetDevice(it->device));
  UBGL_CUDA(cudaDeviceSynchronize());
  if (n > it->cap) {
    it->release();
    it->cap = n;
    UBGL_CUDA(cudaMalloc(&it->items, sizeof(Item) * n));
    UBGL_CUDA(cudaMalloc(&it->bin, n));
    UBGL_CUDA(cudaMalloc(&it->sbin, n));
    UBGL_CUDA(cudaMalloc(&it->idx, sizeof(int) * n));
    UBGL_CUDA(cudaMalloc(&it->order, sizeof(int) * n));
    UBGL_CUDA(cudaMalloc(&it->off, sizeof(int) * 101));
    UBGL_CUDA(cudaMalloc(&it->spos, sizeof(float2) * n));
    UBGL_CUDA(cudaMalloc(&it->key, sizeof(unsigned) * n));
    UBGL_CUDA(cudaMalloc(&it->skey, sizeof(unsigned) * n));
    UBGL_CUDA(cudaMalloc(&it->r2max, sizeof(unsigned)));
    UBGL_CUDA(cudaMalloc(&it->off2, sizeof(int) * (100 * ITEMS_NY + 1)));
    size_t t1 = 0, t2 = 0;
    UBGL_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, t1, it->bin, it->sbin, it->idx, it->order, n, 0, 7));
    UBGL_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, t2, it->key, it->skey, it->idx, it->order, n, 0, 18));
    it->tmp_bytes = std::max(t1, t2);
    UBGL_CUDA(cudaMalloc(&it->tmp, it->tmp_bytes));
  }
  it->n = n;
  if (n) UBGL_CUDA(cudaMemcpy(it->items, items, sizeof(Item) * n, cudaMemcpyHostToDevice));
  UBGL_CATCH
}

int ubgl_items_download(ubgl_items_t *it, ubgl_item *items, int cap, int *n) {
  UBGL_TRY
  NEED(it, "items");
  UBGL_CUDA(cudaSetDevice(it->device));
  UBGL_CUDA(cudaDeviceSynchronize());
  if (n) *n = it->n;
  const int m = std::min(cap, it->n);
  if (items && m > 0) UBGL_CUDA(cudaMemcpy(items, it->items, sizeof(Item) * m, cudaMemcpyDeviceToHost));
  UBGL_CATCH
}

int ubgl_items_advect_simple(ubgl_items_t *it, ubgl_sim_t *sim, float game_dt) {
  UBGL_TRY
  NEED(it, "items");
  SIM(sim);
  UBGL_REQUIRE(S.device == it->device, "items and simulation live on different devices");
  const int n = it->n;
  if (n == 0) return UBGL_OK;
  cudaStream_t st = S.stream;
  const int g = ceil_div(n, 256);
  static const int variant = [] { // 1: one level of x-bins, every pair of a bin tested; 2 (default): + y-cells
    const char *e = getenv("UBGL_ITEMS_VARIANT");
    return (e && e[0] == '1') ? 1 : 2;
  }();
  if (variant == 2) {
    static_assert(100 * ITEMS_NY < (1 << 18), "sort key bits");
    const int nkeys = 100 * ITEMS_NY;
    const float yrange = S.pwidth * (float)S.H / (float)S.W;
    size_t tb = it->tmp_bytes;
    UBGL_CUDA(cudaMemsetAsync(it->r2max, 0, sizeof(unsigned), st));
    UBGL_LAUNCH(&S.lc, K_ITEMS, 0, st, k_items_r2max<<<std::min(g, 1184), 256, 0, st>>>(it->items, n, it->r2max));
    UBGL_LAUNCH(&S.lc, K_ITEMS, 0, st, k_items_key2<<<g, 256, 0, st>>>(it->items, n, it->r2max, yrange, it->key, it->idx));
    S.lc.n += 1; // the radix sort below (library kernels, not counted one by one)
    UBGL_CUDA(cub::DeviceRadixSort::SortPairs(it->tmp, tb, it->key, it->skey, it->idx, it->order, n, 0, 18, st));
    UBGL_LAUNCH(&S.lc, K_ITEMS, 0, st, k_items_offsets2<<<ceil_div(nkeys + 1, 256), 256, 0, st>>>(it->skey, n, nkeys, it->off2));
    UBGL_LAUNCH(&S.lc, K_ITEMS, 0, st, k_items_gather_pos<<<g, 256, 0, st>>>(it->items, it->order, n, it->spos));
    UBGL_LAUNCH(&S.lc, K_ITEMS, 1, st,
                (k_items_advect<2><<<ceil_div(n, 128), 128, 0, st>>>(it->items, it->order, nullptr, it->skey, it->off2,
                                                                    it->spos, n, game_dt, S.field(F_FLAG), S.field(F_VX),
                                                                    S.field(F_VY), S.field(F_P), S.field(F_VX_ACCUM),
                                                                    S.field(F_VY_ACCUM), S.pwidth, S.h)));
    return UBGL_OK;
  }
  size_t tb = it->tmp_bytes;
  UBGL_LAUNCH(&S.lc, K_ITEMS, 0, st, k_items_bin<<<g, 256, 0, st>>>(it->items, n, it->bin, it->idx));
  S.lc.n += 1; // the radix sort below (library kernels, not counted one by one)
  UBGL_CUDA(cub::DeviceRadixSort::SortPairs(it->tmp, tb, it->bin, it->sbin, it->idx, it->order, n, 0, 7, st));
  UBGL_LAUNCH(&S.lc, K_ITEMS, 0, st, k_items_offsets<<<g, 256, 0, st>>>(it->sbin, n, it->off));
  UBGL_LAUNCH(&S.lc, K_ITEMS, 0, st, k_items_gather_pos<<<g, 256, 0, st>>>(it->items, it->order, n, it->spos));
  UBGL_LAUNCH(&S.lc, K_ITEMS, 1, st,
              (k_items_advect<1><<<ceil_div(n, 128), 128, 0, st>>>(it->items, it->order, it->sbin, nullptr, it->off,
                                                                  it->spos, n, game_dt, S.field(F_FLAG), S.field(F_VX),
                                                                  S.field(F_VY), S.field(F_P), S.field(F_VX_ACCUM),
                                                                  S.field(F_VY_ACCUM), S.pwidth, S.h)));
  UBGL_CATCH
}

int ubgl_items_advect(ubgl_items_t *it, ubgl_sim_t *sim, float game_dt) {
  UBGL_TRY
  NEED(it, "items");
  SIM(sim);
  UBGL_REQUIRE(S.device == it->device, "items and simulation live on different devices");
  const int n = it->n;
  if (n == 0) return UBGL_OK;
  UBGL_LAUNCH(&S.lc, K_ITEMS, 2, S.stream,
              k_items_advect_rigid<<<ceil_div(n, 64), 64, 0, S.stream>>>(it->items, n, game_dt, S.field(F_FLAG),
                                                                       S.field(F_VX), S.field(F_VY),
                                                                       S.field(F_VX_ACCUM), S.field(F_VY_ACCUM),
                                                                       S.pwidth, S.h));
  UBGL_CATCH
}

// ---- terrain edits ----------------------------------------------------------
int ubgl_sim_draw_circles(ubgl_sim_t *sim, const float *xyd, int n, float val) {
  UBGL_TRY
  SIM(sim);
  UBGL_REQUIRE(n >= 0 && (n == 0 || xyd), "bad circle list");
  if (n == 0) return UBGL_OK;
  for (int c = 0; c < n; c++) {
    // the reference's second loop (terrain.cpp:228-233) has no bounds test: keep the
    // (2 diam + 1)^2 box inside the grid like every call site does
    const float cx = xyd[3 * c], cy = xyd[3 * c + 1];
    const int d = (int)xyd[3 * c + 2];
    UBGL_REQUIRE(d >= 0 && d <= 4096 && (int)cx - d >= 0 && (int)cx + d < S.W && (int)cy - d >= 0 &&
                     (int)cy + d < S.H,
                 "draw_circles: circle box leaves the grid");
  }
  float *d_xyd = nullptr;
  UBGL_CUDA(cudaMallocAsync(&d_xyd, sizeof(float) * 3 * n, S.stream));
  float *hs = S.stage_host((size_t)3 * n); // xyd is caller memory: stage it, stay asynchronous
  std::memcpy(hs, xyd, sizeof(float) * 3 * n);
  UBGL_CUDA(cudaMemcpyAsync(d_xyd, hs, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, S.stream));
  S.stage_done();
  UBGL_LAUNCH(&S.lc, K_TERRAIN, 0, S.stream, k_draw_circles<<<n, 256, 0, S.stream>>>(S.field(F_FLAG), d_xyd, n, val));
  // MG::updateFields + stencil masks (ubootgl_app.cpp:111-112).  The edit wrote 0.0 / 1.0 inside the discs'
  // boxes only: every level is patched around them (UBGL_DISC_UPDATE=0, non-binary flags: rebuilt whole)
  int max_d = 0;
  for (int c = 0; c < n; c++) max_d = std::max(max_d, (int)xyd[3 * c + 2]);
  S.flag_edited_discs(d_xyd, n, max_d);
  UBGL_CUDA(cudaFreeAsync(d_xyd, S.stream));
  UBGL_CATCH
}

int ubgl_sim_set_grids_all(ubgl_sim_t *sim, const float *newflag) {
  UBGL_TRY
  SIM(sim);
  Grid fl = S.field(F_FLAG);
  const float *nf = fl.d;
  int nfp = fl.pitch;
  Grid tmp{};
  if (newflag) {
    tmp = alloc_grid(S.W, S.H, fl.pitch, false);
    upload_grid(tmp, newflag, S.W, S.H, S.stream);
    nf = tmp.d;
  }
  S.will_write(F_VX); // setGrids zeroes the front velocities in solids, not *_current
  UBGL_LAUNCH(&S.lc, K_TERRAIN, 0, S.stream,
              k_set_grids_all<<<grd(S.W, S.H), blk(), 0, S.stream>>>(fl, S.field(F_VX), S.field(F_VY), S.field(F_P), nf, nfp));
  if (newflag) {
    UBGL_CUDA(cudaStreamSynchronize(S.stream));
    free_grid(tmp);
  }
  S.flag_changed(false); // like the reference, setGrids does not rebuild the pyramid
  UBGL_CATCH
}

int ubgl_sim_shift_map(ubgl_sim_t *sim, const float *new_last_column) {
  UBGL_TRY
  SIM(sim);
  NEED(new_last_column, "new_last_column");
  float *d_col = nullptr;
  UBGL_CUDA(cudaMallocAsync(&d_col, sizeof(float) * S.H, S.stream));
  float *hs = S.stage_host((size_t)S.H);
  std::memcpy(hs, new_last_column, sizeof(float) * S.H);
  UBGL_CUDA(cudaMemcpyAsync(d_col, hs, sizeof(float) * S.H, cudaMemcpyHostToDevice, S.stream));
  S.stage_done();
  Grid none{};
  cudaStream_t st = S.stream;
  // velocities: front shifted, back receives the same values (ubootgl_app.cpp:254-266)
  S.will_write(F_VX); // *_current keeps the unshifted field, like the reference's separate copy
  UBGL_LAUNCH(&S.lc, K_TERRAIN, 0, st, k_shift_rows<<<S.H, 256, 0, st>>>(S.field(F_VX), S.field(F_VXB), 2, nullptr));
  UBGL_LAUNCH(&S.lc, K_TERRAIN, 0, st, k_shift_rows<<<S.H - 1, 256, 0, st>>>(S.field(F_VY), S.field(F_VYB), 2, nullptr));
  UBGL_LAUNCH(&S.lc, K_TERRAIN, 0, st, k_shift_rows<<<S.H, 256, 0, st>>>(S.field(F_P), none, 1, nullptr)); // :268-272
  // terrain mask: flagSimRes scrolls by one column, the generated column enters on the right
  // (terrain.cpp:114-118,141-153; the procedural generator itself stays on the host)
  UBGL_LAUNCH(&S.lc, K_TERRAIN, 0, st, k_shift_rows<<<S.H, 256, 0, st>>>(S.field(F_FLAG), none, 1, d_col));
  Grid fl = S.field(F_FLAG);
  UBGL_LAUNCH(&S.lc, K_TERRAIN, 0, st,
              k_set_grids_all<<<grd(S.W, S.H), blk(), 0, st>>>(fl, S.field(F_VX), S.field(F_VY), S.field(F_P), fl.d, fl.pitch)); // :274-278
  UBGL_LAUNCH(&S.lc, K_TERRAIN, 0, st, k_inlet<<<1, 1024, 0, st>>>(fl, S.field(F_VX), S.field(F_VXB))); // :280-289
  UBGL_CUDA(cudaFreeAsync(d_col, st));
  S.save_current();    // :293
  S.flag_changed(true); // sim.mg.updateFields(sim.flag), :296
  UBGL_CATCH
}

} // extern "C"
