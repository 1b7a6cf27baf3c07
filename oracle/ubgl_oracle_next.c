/* TEST INFRASTRUCTURE ONLY -- see ubgl_oracle.h.  CPU restatement (plain C) of
 * the callers either side of the fluid step (SURVEY.md 8f "next" rows):
 *
 *  1. co-located velocity texture + fluid tracer advection -- GLSL compute
 *     shaders in the reference (interp_shader.cs, advect_tracer_points.cs,
 *     shift_tracers.cs).  No GL implementation exists in this image, so the
 *     shaders are executed as C++: oracle/Makefile compiles the UNMODIFIED .cs
 *     sources where they lie through oracle/shim/glsl_shim.hpp into
 *     oracle/_ref/libubgl_glsl.so, and tests/test_oracle_next.py pins this
 *     restatement on them bit for bit (colocate at 4 sizes, 120 tracer frames
 *     incl. respawn, ring wrap and freezing).  PINNED on the shader source;
 *     what stays this project's reading is the texture FILTER arithmetic, which
 *     the specification leaves open: the restatement and the shim both follow
 *     the OpenGL 4.5 core texture-filtering rules (spec 8.14.2
 *     "coordinate wrapping and texel selection": u = s*w - 1/2, i0 = floor(u),
 *     alpha = frac(u), wrap = the texture-object default GL_REPEAT, LOD 0 in a
 *     compute shader => magnification filter, whose default is GL_LINEAR) in
 *     exact fp32, where real GPUs use ~8-bit filter weights.
 *  2. Simulation::advectFloatingItemsSimple (advect_floating_items.cpp:148-274) and
 *     Simulation::advectFloatingItems (:16-146)
 *     with bilinearSample / bilinearScatter (interpolators.hpp:11-40) and
 *     psampleFlagLinear / psampleFlagNormal (simulation.cpp:398-420).
 *     PARITY PINNED against the unmodified reference TU in oracle/_ref
 *     (tests/test_oracle_next.py).
 *  3. Terrain::drawCircle at scale 1 (terrain.cpp:213-234), Simulation::setGrids
 *     (simulation.hpp:82-98) and the field part of UbootGlApp::shiftMap
 *     (ubootgl_app.cpp:252-293).  drawCircle PINNED against oracle/_ref;
 *     shiftMap's field part restates caller code of the (unbuildable) game app.
 *
 * All citations are file:line under /root/reference.  Scalar fp32,
 * -ffp-contract=off.
 */
#include "ubgl_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define IDX(x, y, w) ((size_t)(y) * (size_t)(w) + (size_t)(x))

/* ------------------------------------------------------------------------
 * 1. textures, co-located velocity, tracers
 * ---------------------------------------------------------------------- */
static inline int wrap_repeat(int i, int n) {
  int m = i % n;
  return m < 0 ? m + n : m;
}

/* GL_LINEAR, GL_REPEAT fetch of component `c` of a texture with `nc`
 * interleaved fp32 components */
static float tex_linear(const float *t, int w, int h, int nc, int c, float s, float tt) {
  const float u = s * (float)w - 0.5f, v = tt * (float)h - 0.5f;
  const float fu = floorf(u), fv = floorf(v);
  const float a = u - fu, b = v - fv;
  const int i0 = wrap_repeat((int)fu, w), i1 = wrap_repeat((int)fu + 1, w);
  const int j0 = wrap_repeat((int)fv, h), j1 = wrap_repeat((int)fv + 1, h);
  const float t00 = t[IDX(i0, j0, w) * nc + c], t10 = t[IDX(i1, j0, w) * nc + c];
  const float t01 = t[IDX(i0, j1, w) * nc + c], t11 = t[IDX(i1, j1, w) * nc + c];
  const float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b);
  const float w01 = (1.0f - a) * b, w11 = a * b;
  return ((w00 * t00 + w10 * t10) + w01 * t01) + w11 * t11;
}

/* interp_shader.cs:15-35 fed by VelocityTextures::updateFromStaggered
 * (velocity_textures.cpp:63-93): vx is (nx-1) x ny, vy nx x (ny-1); outputs
 * are (2nx-1) x (2ny-1): vxy interleaved (RG32F), mag (R32F). */
void orc_colocate(const float *vx, const float *vy, int nx, int ny, float *vxy, float *mag) {
  const int tw = 2 * nx - 1, th = 2 * ny - 1;
  for (int gy = 0; gy < th; gy++)
    for (int gx = 0; gx < tw; gx++) {
      const float sx = (float)gx / (2.0f * (float)nx - 2.0f); /* :18 */
      const float sy = (float)(gy + 1) / (2.0f * (float)ny);  /* :19 */
      const float tx = (float)(gx + 1) / (2.0f * (float)nx);  /* :21 */
      const float ty = (float)gy / (2.0f * (float)ny - 2.0f); /* :22 */
      const float a = tex_linear(vx, nx - 1, ny, 1, 0, sx, sy);
      const float b = tex_linear(vy, nx, ny - 1, 1, 0, tx, ty);
      vxy[IDX(gx, gy, tw) * 2 + 0] = a;
      vxy[IDX(gx, gy, tw) * 2 + 1] = b;
      if (mag) mag[IDX(gx, gy, tw)] = sqrtf(a * a + b * b); /* length(), :27 */
    }
}

static unsigned wang_hash(unsigned seed) { /* advect_tracer_points.cs:20-27 */
  seed = (seed ^ 61u) ^ (seed >> 16);
  seed *= 9u;
  seed = seed ^ (seed >> 4);
  seed *= 0x27d4eb2du;
  seed = seed ^ (seed >> 15);
  return seed;
}

/* advect_tracer_points.cs:42-82, one invocation per tracer.  points is
 * [ntracers][npoints] vec2; tex_vxy is the (tw x th) output of orc_colocate;
 * tex_flag an (fw x fh) R32F texture (velocity_textures.cpp:95-101, level 0). */
void orc_tracers_advect(float *points, unsigned *start, unsigned *end, float *ages, int ntracers,
                        int npoints, float dt, float pdx, float pdy, unsigned rand_seed,
                        const float *vxy, int tw, int th, const float *flagtex, int fw, int fh) {
  for (unsigned gid = 0; gid < (unsigned)ntracers; gid++) {
    unsigned rng = wang_hash(gid + rand_seed); /* :46 */
    const unsigned base = gid * (unsigned)npoints;
    const unsigned curr = end[gid] % (unsigned)npoints;
    const unsigned next = (end[gid] + 1u) % (unsigned)npoints;
    const float cx = points[(base + curr) * 2], cy = points[(base + curr) * 2 + 1];
    /* RK2 / midpoint rule, :56-59 */
    const float sx = cx / pdx, sy = cy / pdy;
    const float v1x = tex_linear(vxy, tw, th, 2, 0, sx, sy), v1y = tex_linear(vxy, tw, th, 2, 1, sx, sy);
    const float mx = cx + (v1x * dt) * 0.5f, my = cy + (v1y * dt) * 0.5f;
    const float v2x = tex_linear(vxy, tw, th, 2, 0, mx / pdx, my / pdy);
    const float v2y = tex_linear(vxy, tw, th, 2, 1, mx / pdx, my / pdy);
    float nx = cx + v2x * dt, ny = cy + v2y * dt;
    /* out of bounds or inside terrain: freeze and age faster, :62-65 */
    if (cx < 0.0f || cy < 0.0f || cx > pdx || cy > pdy ||
        tex_linear(flagtex, fw, fh, 1, 0, sx, sy) < 0.6f) {
      nx = cx;
      ny = cy;
      ages[gid] += 0.1f;
    }
    if (ages[gid] > 2.0f * 3.141f) { /* respawn, :68-74 */
      start[gid] = 0;
      end[gid] = 0;
      rng = 1664525u * rng + 1013904223u;
      nx = ((float)(rng % 100000u) / 100000.0f) * pdx;
      rng = 1664525u * rng + 1013904223u;
      ny = ((float)(rng % 100000u) / 100000.0f) * pdy;
      points[(base + 0) * 2] = nx;
      points[(base + 0) * 2 + 1] = ny;
      ages[gid] = 0.0f;
    } else { /* :75-81 */
      end[gid] = next;
      if (start[gid] == end[gid]) start[gid] = (start[gid] + 1u) % (unsigned)npoints;
      points[(base + next) * 2] = nx;
      points[(base + next) * 2 + 1] = ny;
      ages[gid] = ages[gid] + 0.02f;
    }
  }
}

/* shift_tracers.cs:18-27 (the ribbon vertices it also shifts are rendering state) */
void orc_tracers_shift(float *points, int ntracers, int npoints, float shift) {
  for (size_t g = 0; g < (size_t)ntracers * npoints; g++) points[g * 2] += shift;
}

/* ------------------------------------------------------------------------
 * 2. floating items (simple kinematics)
 * ---------------------------------------------------------------------- */
/* bilinearSample, interpolators.hpp:11-27 */
static float bilinear_sample(const float *g, int w, int h, float cx, float cy) {
  cx = fminf(fmaxf(cx, 0.0f), (float)w - 1.1f);
  cy = fminf(fmaxf(cy, 0.0f), (float)h - 1.1f);
  const int ix = (int)cx, iy = (int)cy;
  const float sx = cx - floorf(cx), sy = cy - floorf(cy);
  const float v1 = g[IDX(ix, iy, w)], v2 = g[IDX(ix + 1, iy, w)];
  const float v3 = g[IDX(ix, iy + 1, w)], v4 = g[IDX(ix + 1, iy + 1, w)];
  const float vm1 = v1 + (v2 - v1) * sx, vm2 = v3 + (v4 - v3) * sx;
  return vm1 + (vm2 - vm1) * sy;
}

/* bilinearScatter, interpolators.hpp:29-40 */
static void bilinear_scatter(float *g, int w, int h, float cx, float cy, float v) {
  cx = fminf(fmaxf(cx, 0.0f), (float)w - 1.1f);
  cy = fminf(fmaxf(cy, 0.0f), (float)h - 1.1f);
  const int ix = (int)cx, iy = (int)cy;
  const float sx = cx - floorf(cx), sy = cy - floorf(cy);
  const float ax = 1.0f - sx, ay = 1.0f - sy;
  g[IDX(ix + 1, iy + 1, w)] += sx * sy * v;
  g[IDX(ix, iy + 1, w)] += ax * sy * v;
  g[IDX(ix + 1, iy, w)] += sx * ay * v;
  g[IDX(ix, iy, w)] += ax * ay * v;
}

static inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }

/* Simulation::psampleFlagLinear, simulation.cpp:398-412 */
static float psample_flag_linear(const float *flag, int w, int h, float pwidth, float px, float py) {
  const float s = pwidth / (float)w;
  const float cx = px / s - 0.5f, cy = py / s - 0.5f;
  int ix = (int)cx, iy = (int)cy;
  ix = ix > w - 2 ? w - 2 : ix;
  iy = iy > h - 2 ? h - 2 : iy;
  ix = ix < 0 ? 0 : ix;
  iy = iy < 0 ? 0 : iy;
  const float sx = cx - floorf(cx), sy = cy - floorf(cy);
  const float p01 = flag[IDX(ix, iy + 1, w)], p11 = flag[IDX(ix + 1, iy + 1, w)];
  const float p00 = flag[IDX(ix, iy, w)], p10 = flag[IDX(ix + 1, iy, w)];
  return mixf(mixf(p00, p10, sx), mixf(p01, p11, sx), sy);
}

/* Simulation::advectFloatingItemsSimple, advect_floating_items.cpp:148-274.
 * Items are visited in array order (the reference visits its entt view order;
 * the wrapper in oracle/ref_items.cpp lays the entities out so that the two
 * orders agree).  vx, vy are the FRONT velocity buffers. */
void orc_items_advect_simple(orc_item *items, int n, float game_dt, const float *flag, const float *vx,
                             const float *vy, const float *p, float *vx_accum, float *vy_accum, int W,
                             int H, float pwidth) {
  const float h = pwidth / ((float)W - 1.0f); /* simulation.hpp:60 */
  /* bins of positions by (int)(pos.x*100) % 100, :153-159 */
  int cnt[100] = {0}, off[101];
  int *bin_of = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; i++) {
    bin_of[i] = (int)((unsigned long long)(long long)(int)(items[i].pos[0] * 100.0f) % 100ull);
    cnt[bin_of[i]]++;
  }
  off[0] = 0;
  for (int b = 0; b < 100; b++) off[b + 1] = off[b] + cnt[b];
  float *bx = (float *)malloc(sizeof(float) * 2 * (size_t)(n > 0 ? n : 1));
  int fill[100];
  memcpy(fill, off, sizeof(fill));
  for (int i = 0; i < n; i++) {
    const int k = fill[bin_of[i]]++;
    bx[2 * k] = items[i].pos[0];
    bx[2 * k + 1] = items[i].pos[1];
  }
  for (int i = 0; i < n; i++) {
    orc_item *it = &items[i];
    /* neighbour repulsion inside the bin, :167-181 */
    float rfx = 0.0f, rfy = 0.0f;
    int contacts = 0;
    const int b = bin_of[i];
    for (int k = off[b]; k < off[b + 1]; k++) {
      const float dx = it->pos[0] - bx[2 * k], dy = it->pos[1] - bx[2 * k + 1];
      if (dx * dx + dy * dy < (it->size[0] * it->size[1] * 0.4f)) {
        const float len = fmaxf(0.1f * it->size[0], sqrtf(dx * dx + dy * dy));
        rfx += 0.0001f * (dx / len / len);
        rfy += 0.0001f * (dy / len / len);
        contacts++;
      }
    }
    const float cden = (float)(contacts > 1 ? contacts : 1);
    rfx /= cden;
    rfy /= cden;
    it->force[0] += rfx * 10.0f;
    it->force[1] += rfy * 10.0f;

    const int steps = (int)fmin(15.0f, fmax(1.0f, (double)(fmaxf(fabsf(it->vel[0]), fabsf(it->vel[1])) *
                                                           game_dt / h) * 2.5)); /* :183-185 */
    const float sub = game_dt / (float)steps;
    for (int s = 0; s < steps; s++) {
      it->pos[0] += sub * it->vel[0]; /* :193 */
      it->pos[1] += sub * it->vel[1];
      it->rotation = (float)fmod((double)(it->rotation + sub * it->angVel) + 2 * M_PI, 2 * M_PI); /* :196 */
      const float gs = pwidth / (float)W;
      const float gpx = it->pos[0] / gs, gpy = it->pos[1] / gs; /* :199 */
      if (gpx >= (float)(W - 2) || gpx <= 1.0f || gpy >= (float)(H - 2) || gpy <= 1.0f) continue; /* :202 */
      if (psample_flag_linear(flag, W, H, pwidth, it->pos[0], it->pos[1]) < 1.0f) { /* :208-233 */
        const float px = it->pos[0], py = it->pos[1];
        const float p01 = psample_flag_linear(flag, W, H, pwidth, px - h, py + h);
        const float p11 = psample_flag_linear(flag, W, H, pwidth, px + h, py + h);
        const float p00 = psample_flag_linear(flag, W, H, pwidth, px - h, py - h);
        const float p10 = psample_flag_linear(flag, W, H, pwidth, px + h, py - h);
        float nx = p11 + p10 - p01 - p00, ny = p01 + p11 - p00 - p10; /* simulation.cpp:414-420 */
        float nl = sqrtf(nx * nx + ny * ny);
        if (nl > 0.0f) {
          nx /= nl;
          ny /= nl;
          float d = it->vel[0] * nx + it->vel[1] * ny;
          if (d < 0.0f) { /* reflect(v, n) * 0.7 */
            d = nx * it->vel[0] + ny * it->vel[1];
            it->vel[0] = (it->vel[0] - nx * d * 2.0f) * 0.7f;
            it->vel[1] = (it->vel[1] - ny * d * 2.0f) * 0.7f;
          }
          d = it->force[0] * nx + it->force[1] * ny;
          if (d < 0.0f) {
            d = nx * it->force[0] + ny * it->force[1];
            it->force[0] = (it->force[0] - nx * d * 2.0f) * 0.7f;
            it->force[1] = (it->force[1] - ny * d * 2.0f) * 0.7f;
          }
          const float ang = 0.5f * (float)M_PI;
          const float ca = cosf(ang), sa = sinf(ang);
          const float lx = nx * ca - ny * sa, ly = nx * sa + ny * ca; /* glm::rotate(n, pi/2) */
          const float lat_vel = lx * it->vel[0] + ly * it->vel[1];
          const float rot_vel = it->angVel * (it->size[0] + it->size[1]) * 0.5f;
          const float lat_diff = lat_vel - rot_vel;
          it->force[0] += lat_diff * lx * 1.0f;
          it->force[1] += lat_diff * ly * 1.0f;
        } else {
          it->vel[0] = 0.0f;
          it->vel[1] = 0.0f;
        }
        it->bumpCount++;
      }
      float efx = 0.0f * it->mass + it->force[0], efy = -0.5f * it->mass + it->force[1]; /* :235 */
      const float gx = it->pos[0] / h, gy = it->pos[1] / h;
      /* bilinearVel, advect_floating_items.cpp:11-14 */
      const float dvx = bilinear_sample(vx, W - 1, H, gx - 0.5f, gy) - it->vel[0];
      const float dvy = bilinear_sample(vy, W, H - 1, gx, gy - 0.5f) - it->vel[1];
      const float drag = 2000.0f * (it->size[0] + it->size[1]);
      efx += drag * dvx;
      efy += drag * dvy;
      if (gx > 1.0f && gx < (float)(W - 1) - 2.0f && gy < 1.0f && gy < (float)H - 2.0f) { /* :240-248 (sic) */
        const float ddx = dvx * it->size[0] * it->size[1], ddy = dvy * it->size[0] * it->size[1];
        bilinear_scatter(vx_accum, W - 1, H, gx - 0.5f, gy, -ddx);
        bilinear_scatter(vy_accum, W, H - 1, gx, gy - 0.5f, -ddy);
      }
      it->vel[0] += sub * efx / it->mass; /* :251 */
      it->vel[1] += sub * efy / it->mass;
      const int ix = (int)gpx, iy = (int)gpy; /* :254 */
      const float fluid_ang = -((vx[IDX(ix, iy, W - 1)] - vx[IDX(ix, iy - 1, W - 1)]) -
                                (vy[IDX(ix, iy, W)] - vy[IDX(ix - 1, iy, W)])) / h / 2.0f; /* :255-257 */
      const float ang_mass = it->size[0] * it->size[1] * it->mass * (1.0f / 12.0f);
      const float k = fminf(1.0f, sub / ang_mass * 0.0005f * (it->size[0] + it->size[1]) / 4.0f);
      it->angVel += k * (fluid_ang - it->angVel);
      it->angVel += sub / ang_mass * it->angForce;
      vx_accum[IDX(ix, iy, W - 1)] -= it->angVel * k * 0.01f * p[IDX(ix, iy, W)]; /* :266-272 */
      vx_accum[IDX(ix, iy - 1, W - 1)] += it->angVel * k * 0.01f * p[IDX(ix, iy - 1, W)];
      vy_accum[IDX(ix, iy, W)] += it->angVel * k * 0.01f * p[IDX(ix, iy, W)];
      vy_accum[IDX(ix - 1, iy, W)] -= it->angVel * k * 0.01f * p[IDX(ix - 1, iy, W)];
    }
    it->angForce = 0.0f;
    it->force[0] = 0.0f;
    it->force[1] = 0.0f;
  }
  free(bin_of);
  free(bx);
}

/* Simulation::advectFloatingItems, advect_floating_items.cpp:16-146: rigid rectangular
 * bodies (CoItem + CoKinematics: the submarines and torpedoes) -- five terrain probes per
 * sub-step, drag from the fluid sampled along the four sides, reaction scattered into the
 * accumulators.  glm::rotate / normalize / reflect as GLM defines them (2-D rotation,
 * a / length(a), I - N dot(N,I) 2).  vx, vy are the FRONT buffers. */
static void rot2(float x, float y, float ang, float *ox, float *oy) {
  const float c = cosf(ang), s = sinf(ang);
  *ox = x * c - y * s;
  *oy = x * s + y * c;
}
void orc_items_advect(orc_item *items, int n, float game_dt, const float *flag, const float *vx,
                      const float *vy, float *vx_accum, float *vy_accum, int W, int H, float pwidth) {
  const float h = pwidth / ((float)W - 1.0f); /* simulation.hpp:60 */
  static const float SPX[5] = {1.0f, -1.0f, 1.0f, -1.0f, 0.0f}, SPY[5] = {1.0f, 1.0f, -1.0f, -1.0f, 0.0f}; /* :48-50 */
  static const float SFX[4] = {-1.0f, 1.0f, 0.0f, 0.0f}, SFY[4] = {0.0f, 0.0f, -1.0f, 1.0f};               /* :80-81 */
  for (int q = 0; q < n; q++) {
    orc_item *it = &items[q];
    const int steps = (int)fmin(15.0f, fmax(1.0f, (double)(fmaxf(fabsf(it->vel[0]), fabsf(it->vel[1])) *
                                                           game_dt / h) * 2.5)); /* :23-25 */
    const float sub = game_dt / (float)steps;
    for (int st = 0; st < steps; st++) {
      const float bx = it->pos[0], by = it->pos[1]; /* posBefore */
      it->pos[0] += sub * it->vel[0];               /* :33 */
      it->pos[1] += sub * it->vel[1];
      it->rotation = (float)fmod((double)(it->rotation + sub * it->angVel) + 2 * M_PI, 2 * M_PI); /* :36-37 */
      const float gs = pwidth / (float)W;
      const float gpx = it->pos[0] / gs, gpy = it->pos[1] / gs;
      if (gpx >= (float)(W - 2) || gpx <= 1.0f || gpy >= (float)(H - 2) || gpy <= 1.0f) continue; /* :42-45 */
      it->force[0] += 0.0f * it->mass; /* :47 */
      it->force[1] += -0.5f * it->mass;
      for (int k = 0; k < 5; k++) { /* terrain probes, :52-72 */
        float spx, spy;
        rot2(SPX[k] * 0.5f * it->size[0], SPY[k] * 0.5f * it->size[1], it->rotation, &spx, &spy);
        if (psample_flag_linear(flag, W, H, pwidth, it->pos[0] + spx, it->pos[1] + spy) < 0.5f) {
          const float mx = 0.5f * (bx + it->pos[0]) + spx, my = 0.5f * (by + it->pos[1]) + spy;
          const float p01 = psample_flag_linear(flag, W, H, pwidth, mx - h, my + h);
          const float p11 = psample_flag_linear(flag, W, H, pwidth, mx + h, my + h);
          const float p00 = psample_flag_linear(flag, W, H, pwidth, mx - h, my - h);
          const float p10 = psample_flag_linear(flag, W, H, pwidth, mx + h, my - h);
          float nx = p11 + p10 - p01 - p00, ny = p01 + p11 - p00 - p10; /* simulation.cpp:414-420 */
          float nl = sqrtf(nx * nx + ny * ny);
          nx /= nl; /* normalize(): 0/0 = NaN when the normal vanishes */
          ny /= nl;
          if (psample_flag_linear(flag, W, H, pwidth, bx + spx, by + spy) > 0.5f) {
            it->pos[0] = bx;
            it->pos[1] = by;
          }
          nl = sqrtf(nx * nx + ny * ny);
          if (nl > 0.0f) { /* false for NaN */
            nx /= nl;
            ny /= nl;
            if (it->vel[0] * nx + it->vel[1] * ny < 0.0f) { /* reflect(v, n) * 0.7 */
              const float d = nx * it->vel[0] + ny * it->vel[1];
              it->vel[0] = (it->vel[0] - nx * d * 2.0f) * 0.7f;
              it->vel[1] = (it->vel[1] - ny * d * 2.0f) * 0.7f;
            }
            const float fl = psample_flag_linear(flag, W, H, pwidth, it->pos[0] + spx, it->pos[1] + spy);
            it->vel[0] += nx * 0.07f * fl;
            it->vel[1] += ny * 0.07f * fl;
            const float df = it->force[0] * nx + it->force[1] * ny;
            if (df < 0.0f) {
              it->force[0] += 1.1f * df * nx;
              it->force[1] += 1.1f * df * ny;
            }
          }
          it->bumpCount++;
        }
      }
      const float efx = 0.0f * it->mass + it->force[0], efy = -0.5f * it->mass + it->force[1]; /* :74 */
      float cfx = 0.0f, cfy = 0.0f;
      const float ang_force = it->angForce;
      const float side[4] = {it->size[1], it->size[1], it->size[0], it->size[0]}; /* :83 */
      for (int i = 0; i < 4; i++) {
        const int nsp = (int)fmaxf(2.0f, side[i] / h); /* :86 */
        for (int k = 0; k < nsp; k++) {
          const float tpar = 1.0f - (float)k * 2.0f / (float)(nsp - 1);
          const float sx = SFX[i] + fabsf(SFY[i]) * tpar, sy = SFY[i] + fabsf(SFX[i]) * tpar; /* :89-91 */
          const float lx = sx * it->size[0] * 0.5f, ly = sy * it->size[1] * 0.5f;
          float tx, ty;
          rot2(lx, ly, it->rotation, &tx, &ty);
          tx += it->pos[0];
          ty += it->pos[1];
          const float gx = tx / h, gy = ty / h;
          float rx, ry;
          rot2(lx, ly, it->rotation + 0.5f * 3.141f, &rx, &ry);
          const float dvx = bilinear_sample(vx, W - 1, H, gx - 0.5f, gy) - (it->vel[0] + 3.141f * rx * it->angVel);
          const float dvy = bilinear_sample(vy, W, H - 1, gx, gy - 0.5f) - (it->vel[1] + 3.141f * ry * it->angVel);
          const float sl = sqrtf(sx * sx + sy * sy);
          float nx, ny, ox, oy;
          rot2(sx / sl, sy / sl, it->rotation, &nx, &ny);  /* rotate(normalize(sp), rotation) */
          rot2(SFX[i], SFY[i], it->rotation, &ox, &oy);
          const float pr = fminf(0.0f, dvx * ox + dvy * oy);
          const float fx = nx * pr, fy = ny * pr;
          cfx += fx * 400000.0f * (0.003f + side[i]) * side[i] / (float)nsp; /* :110-111 */
          cfy += fy * 400000.0f * (0.003f + side[i]) * side[i] / (float)nsp;
          if (gx < 1.0f || gx > (float)(W - 1) - 2.0f || gy < 1.0f || gy > (float)H - 2.0f) continue; /* :113-115 */
          const float ddx = fx * (0.003f + side[i]) * side[i] / (float)nsp * sub * 18000000.0f;
          const float ddy = fy * (0.003f + side[i]) * side[i] / (float)nsp * sub * 18000000.0f;
          bilinear_scatter(vx_accum, W - 1, H, gx - 0.5f, gy, -ddx);
          bilinear_scatter(vy_accum, W, H - 1, gx, gy - 0.5f, -ddy);
        }
      }
      {
        const float gx = it->pos[0] / h, gy = it->pos[1] / h; /* :125-126 */
        const float kk = 1000.0f * (it->size[0] + it->size[1]);
        cfx += kk * (bilinear_sample(vx, W - 1, H, gx - 0.5f, gy) - it->vel[0]);
        cfy += kk * (bilinear_sample(vy, W, H - 1, gx, gy - 0.5f) - it->vel[1]);
      }
      it->vel[0] += sub * (efx + cfx) / it->mass; /* :129 */
      it->vel[1] += sub * (efy + cfy) / it->mass;
      const float ang_mass = it->size[0] * it->size[1] * it->mass * (1.0f / 12.0f);
      it->angVel += sub * ang_force / ang_mass;
      it->angVel = (float)((double)it->angVel * 0.98); /* :135 */
    }
    it->angForce = 0.0f;
    it->force[0] = 0.0f;
    it->force[1] = 0.0f;
  }
}

/* ------------------------------------------------------------------------
 * 3. terrain edits on the simulation-resolution flag (terrain scale 1)
 * ---------------------------------------------------------------------- */
/* Terrain::drawCircle, terrain.cpp:213-234, for scale == 1 where flagFullRes and
 * flagSimRes have the same size: the disc is written with `val`, then the
 * (2 diam + 1)^2 box around it is re-thresholded by subSample (:3-11: > 0.99).
 * `flag` is the sim-resolution mask the solver sees (terrain.flagSimRes). */
void orc_draw_circle(float *flag_full, float *flag_sim, int w, int h, float cx, float cy, int diam,
                     float val) {
  for (int y = -diam; y <= diam; y++)
    for (int x = -diam; x <= diam; x++) {
      if (x * x + y * y > diam * diam || cx + (float)x < 0.0f || (float)x + cx > (float)w ||
          (float)y + cy < 2.0f || (float)y + cy >= (float)(h - 3))
        continue;
      flag_full[IDX((int)((float)x + cx), (int)((float)y + cy), w)] = val;
    }
  const int sd = (diam - 1) / 1 + 1;
  for (int y = -sd; y <= sd; y++)
    for (int x = -sd; x <= sd; x++) {
      const int ix = (int)(cx / 1.0f + (float)x), iy = (int)(cy / 1.0f + (float)y);
      flag_sim[IDX(ix, iy, w)] = flag_full[IDX(ix, iy, w)] / 1.0f / 1.0f > 0.99f ? 1.0f : 0.0f;
    }
}

/* Simulation::setGrids for every cell, as UbootGlApp::shiftMap does
 * (ubootgl_app.cpp:274-278 calling simulation.hpp:82-98); vx, vy = FRONT buffers */
void orc_set_grids_all(float *flag, float *vx, float *vy, float *p, const float *newflag, int W, int H) {
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      const float v = newflag[IDX(x, y, W)];
      flag[IDX(x, y, W)] = v;
      if (v == 0.0f) {
        if (x < W - 1) vx[IDX(x, y, W - 1)] = 0.0f;
        if (x > 0) vx[IDX(x - 1, y, W - 1)] = 0.0f;
        if (y < H - 1) vy[IDX(x, y, W)] = 0.0f;
        if (y > 0) vy[IDX(x, y - 1, W)] = 0.0f;
        p[IDX(x, y, W)] = 0.0f;
      }
    }
}

/* Field part of UbootGlApp::shiftMap (ubootgl_app.cpp:252-293): scroll vx, vy
 * (front and back) and p one column to the left, apply the scrolled terrain
 * mask with setGrids, reset the inlet column, saveCurrentVelocityFields. */
void orc_shift_map(float *flag, float *vxf, float *vxb, float *vyf, float *vyb, float *p, float *vxc,
                   float *vyc, const float *newflag, int W, int H) {
  for (int y = 0; y < H; y++)
    for (int x = 2; x < W - 1; x++) { /* :254-259 */
      vxf[IDX(x - 1, y, W - 1)] = vxf[IDX(x, y, W - 1)];
      vxb[IDX(x - 1, y, W - 1)] = vxf[IDX(x, y, W - 1)];
    }
  for (int y = 0; y < H - 1; y++)
    for (int x = 2; x < W; x++) { /* :261-266 */
      vyf[IDX(x - 1, y, W)] = vyf[IDX(x, y, W)];
      vyb[IDX(x - 1, y, W)] = vyf[IDX(x, y, W)];
    }
  for (int y = 0; y < H; y++)
    for (int x = 1; x < W; x++) p[IDX(x - 1, y, W)] = p[IDX(x, y, W)]; /* :268-272 */
  orc_set_grids_all(flag, vxf, vyf, p, newflag, W, H);                 /* :274-278 */
  float inlet_area = 1.0f;
  for (int y = 0; y < H - 1; y++) inlet_area += flag[IDX(0, y, W)]; /* :281-284 */
  for (int y = 0; y < H; y++) {
    const float v = 0.07f * (float)H / inlet_area; /* :287 */
    vxf[IDX(0, y, W - 1)] = vxb[IDX(0, y, W - 1)] = v * flag[IDX(0, y, W)];
  }
  memcpy(vxc, vxf, sizeof(float) * (size_t)(W - 1) * H); /* :293 -> simulation.cpp:16-19 */
  memcpy(vyc, vyf, sizeof(float) * (size_t)W * (H - 1));
}
