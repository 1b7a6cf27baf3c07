/* TEST INFRASTRUCTURE ONLY -- see ubgl_oracle.h.  CPU restatement (plain C) of
 * the reference's fluid-step hot path.  All citations are file:line under
 * /root/reference (te42kyfo/ubootgl).  Arithmetic is scalar fp32 with
 * -ffp-contract=off; the reference is -Ofast AVX2/FMA, so agreement with it is
 * to rounding (rel-L2 ~1e-7 per stage), pinned by tests/test_oracle_*.py.
 *
 * PARITY PINNED against oracle/_ref (the unmodified reference TUs) and the
 * fixtures in tests/golden/.
 */
#include "ubgl_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void orc_set_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
int orc_num_procs(void) {
#ifdef _OPENMP
  return omp_get_num_procs();
#else
  return 1;
#endif
}

#define IDX(x, y, w) ((size_t)(y) * (size_t)(w) + (size_t)(x))

/* ------------------------------------------------------------------------
 * pressure_solver.cpp
 * ---------------------------------------------------------------------- */

/* smoothingKernel, pressure_solver.cpp:10-24 */
static inline float smoothing_kernel(const float *p, const float *f,
                                     const float *flag, int w, float hh,
                                     float alpha, int x, int y) {
  float val = 0.0f, sum = 0.0f;
  sum += flag[IDX(x - 1, y, w)] + flag[IDX(x + 1, y, w)] +
         flag[IDX(x, y + 1, w)] + flag[IDX(x, y - 1, w)];
  val += p[IDX(x - 1, y, w)] * flag[IDX(x - 1, y, w)];
  val += p[IDX(x + 1, y, w)] * flag[IDX(x + 1, y, w)];
  val += p[IDX(x, y - 1, w)] * flag[IDX(x, y - 1, w)];
  val += p[IDX(x, y + 1, w)] * flag[IDX(x, y + 1, w)];
  val += f[IDX(x, y, w)] * hh * hh;
  val /= sum;
  if (sum == 0.0f)
    val = 0.0f;
  return flag[IDX(x, y, w)] * (alpha * val + (1.0f - alpha) * p[IDX(x, y, w)]);
}

/* rbgs, canonical red-black order (paths 1 and 2 of pressure_solver.cpp:49-72;
 * path 3 :73-87 is the racy pipelined variant and is deliberately NOT
 * restated -- SURVEY.md section 8a M3).  "red" = x starts at 1 + (y%2)
 * (:35-40), "black" = x starts at 1 + ((y+1)%2) (:42-47). */
static void rbgs_once(float *p, const float *f, const float *flag, int w, int h,
                      float hh, float alpha) {
#pragma omp parallel for schedule(static)
  for (int y = 1; y < h - 1; y++)
    for (int x = 1 + (y % 2); x < w - 1; x += 2)
      p[IDX(x, y, w)] = smoothing_kernel(p, f, flag, w, hh, alpha, x, y);
#pragma omp parallel for schedule(static)
  for (int y = 1; y < h - 1; y++)
    for (int x = 1 + ((y + 1) % 2); x < w - 1; x += 2)
      p[IDX(x, y, w)] = smoothing_kernel(p, f, flag, w, hh, alpha, x, y);
}

void orc_rbgs(float *p, const float *f, const float *flag, int w, int h,
              float hh, float alpha, int sweeps) {
  for (int i = 0; i < sweeps; i++)
    rbgs_once(p, f, flag, w, h, hh, alpha);
}

/* calculateResidualField, pressure_solver.cpp:91-116.  The reference sums r^2
 * in fp32 with an OpenMP reduction (order unspecified); the restatement sums in
 * double so that it is the better-conditioned side of the comparison. */
float orc_residual(const float *p, const float *f, const float *flag, float *r,
                   int w, int h, float hh) {
  double l2r = 0.0;
  float ihsq = 1.0f / hh / hh;
#pragma omp parallel for schedule(static) reduction(+ : l2r)
  for (int y = 1; y < h - 1; y++) {
    for (int x = 1; x < w - 1; x++) {
      float pc = p[IDX(x, y, w)];
      float val = 0.0f;
      val += p[IDX(x - 1, y, w)] * flag[IDX(x - 1, y, w)] +
             pc * (1.0f - flag[IDX(x - 1, y, w)]);
      val += p[IDX(x + 1, y, w)] * flag[IDX(x + 1, y, w)] +
             pc * (1.0f - flag[IDX(x + 1, y, w)]);
      val += p[IDX(x, y - 1, w)] * flag[IDX(x, y - 1, w)] +
             pc * (1.0f - flag[IDX(x, y - 1, w)]);
      val += p[IDX(x, y + 1, w)] * flag[IDX(x, y + 1, w)] +
             pc * (1.0f - flag[IDX(x, y + 1, w)]);
      val += -4.0f * pc;
      val *= ihsq;
      float rv = (f[IDX(x, y, w)] + val) * flag[IDX(x, y, w)];
      r[IDX(x, y, w)] = rv;
      l2r += (double)rv * (double)rv;
    }
  }
  return (float)sqrt(l2r);
}

/* 9-point full weighting used by both restrict (pressure_solver.cpp:118-132)
 * and MG::updateFields (pressure_solver.hpp:42-49); same association order. */
static inline float fw9(const float *r, int w, int x, int y) {
  float v;
  v = r[IDX(2 * x - 1, 2 * y - 1, w)] * 1 + r[IDX(2 * x + 0, 2 * y - 1, w)] * 2 +
      r[IDX(2 * x + 1, 2 * y - 1, w)] * 1;
  v += r[IDX(2 * x - 1, 2 * y + 0, w)] * 2 + r[IDX(2 * x + 0, 2 * y + 0, w)] * 4 +
       r[IDX(2 * x + 1, 2 * y + 0, w)] * 2;
  v += r[IDX(2 * x - 1, 2 * y + 1, w)] * 1 + r[IDX(2 * x + 0, 2 * y + 1, w)] * 2 +
       r[IDX(2 * x + 1, 2 * y + 1, w)] * 1;
  return v * (1.0f / 16.0f);
}

/* restrict, pressure_solver.cpp:118-132 */
void orc_restrict(const float *r, int w, int h, float *rc, int wc, int hc) {
  (void)h;
#pragma omp parallel for schedule(static)
  for (int y = 1; y < hc - 1; y++)
    for (int x = 1; x < wc - 1; x++)
      rc[IDX(x, y, wc)] = fw9(r, w, x, y);
}

/* prolongate, pressure_solver.cpp:134-172 (zero-fills the target first :136).
 * The 0.0001 literal is a double in the reference, so `sum` is evaluated in
 * double and then narrowed to float (float sum = ... + 0.0001). */
void orc_prolongate(float *e, int w, int h, const float *ec, const float *flagc,
                    int wc, int hc, const float *flag) {
  (void)hc;
  memset(e, 0, sizeof(float) * (size_t)w * h);
#pragma omp parallel for schedule(static)
  for (int y = 2; y < h - 1; y += 2)
    for (int x = 2; x < w - 1; x += 2)
      e[IDX(x, y, w)] = ec[IDX(x / 2, y / 2, wc)] * flag[IDX(x, y, w)];
#pragma omp parallel for schedule(static)
  for (int y = 2; y < h - 1; y += 2)
    for (int x = 1; x < w - 2; x += 2) {
      float sum = (float)((double)(flagc[IDX(x / 2, y / 2, wc)] +
                                   flagc[IDX(x / 2 + 1, y / 2, wc)]) + 0.0001);
      e[IDX(x, y, w)] = flag[IDX(x, y, w)] *
                        (ec[IDX(x / 2, y / 2, wc)] + ec[IDX(x / 2 + 1, y / 2, wc)]) / sum;
    }
#pragma omp parallel for schedule(static)
  for (int y = 1; y < h - 2; y += 2)
    for (int x = 2; x < w - 1; x += 2) {
      float sum = (float)((double)(flagc[IDX(x / 2, y / 2, wc)] +
                                   flagc[IDX(x / 2, y / 2 + 1, wc)]) + 0.0001);
      e[IDX(x, y, w)] = flag[IDX(x, y, w)] *
                        (ec[IDX(x / 2, y / 2, wc)] + ec[IDX(x / 2, y / 2 + 1, wc)]) / sum;
    }
#pragma omp parallel for schedule(static)
  for (int y = 1; y < h - 2; y += 2)
    for (int x = 1; x < w - 2; x += 2) {
      float sum = (float)((double)(flagc[IDX(x / 2, y / 2, wc)] +
                                   flagc[IDX(x / 2 + 1, y / 2 + 1, wc)] +
                                   flagc[IDX(x / 2 + 1, y / 2, wc)] +
                                   flagc[IDX(x / 2, y / 2 + 1, wc)]) + 0.0001);
      e[IDX(x, y, w)] = flag[IDX(x, y, w)] *
                        (ec[IDX(x / 2, y / 2, wc)] + ec[IDX(x / 2 + 1, y / 2 + 1, wc)] +
                         ec[IDX(x / 2 + 1, y / 2, wc)] + ec[IDX(x / 2, y / 2 + 1, wc)]) / sum;
    }
}

/* correct, pressure_solver.cpp:174-181 */
void orc_correct(float *p, const float *e, int w, int h) {
#pragma omp parallel for schedule(static)
  for (int y = 1; y < h - 1; y++)
    for (int x = 1; x < w - 1; x++)
      p[IDX(x, y, w)] += e[IDX(x, y, w)] * 1.0f;
}

/* setZeroGradientBC, pressure_solver.cpp:183-192 (corners untouched) */
void orc_zero_gradient_bc(float *p, int w, int h) {
  for (int y = 1; y < h - 1; y++) {
    p[IDX(0, y, w)] = p[IDX(1, y, w)];
    p[IDX(w - 1, y, w)] = p[IDX(w - 2, y, w)];
  }
  for (int x = 1; x < w - 1; x++) {
    p[IDX(x, 0, w)] = p[IDX(x, 1, w)];
    p[IDX(x, h - 1, w)] = p[IDX(x, h - 2, w)];
  }
}

/* ---- class MG, pressure_solver.hpp:13-76 ------------------------------- */
typedef struct {
  int levels, width, height;
  int *lw, *lh;
  float **rs, **rcs, **ecs, **es, **flagcs;
  float *p, *f, *flag; /* resident user fields for the handle API */
} OrcMG;

static float *zalloc(size_t n) { return (float *)calloc(n ? n : 1, sizeof(float)); }

/* MG::MG(int,int), pressure_solver.hpp:16-31 */
static void mg_init(OrcMG *m, int w, int h) {
  memset(m, 0, sizeof(*m));
  m->width = w;
  m->height = h;
  int cw = w, ch = h, n = 0;
  while (cw > 3 && ch > 3) { n++; cw /= 2; ch /= 2; }
  m->levels = n;
  m->lw = (int *)malloc(sizeof(int) * (n + 1));
  m->lh = (int *)malloc(sizeof(int) * (n + 1));
  m->rs = (float **)malloc(sizeof(float *) * (n + 1));
  m->rcs = (float **)malloc(sizeof(float *) * (n + 1));
  m->ecs = (float **)malloc(sizeof(float *) * (n + 1));
  m->es = (float **)malloc(sizeof(float *) * (n + 1));
  m->flagcs = (float **)malloc(sizeof(float *) * (n + 1));
  cw = w; ch = h;
  for (int l = 0; l < n; l++) {
    size_t cells = (size_t)cw * ch;
    m->lw[l] = cw; m->lh[l] = ch;
    m->rs[l] = zalloc(cells); m->rcs[l] = zalloc(cells);
    m->ecs[l] = zalloc(cells); m->es[l] = zalloc(cells);
    m->flagcs[l] = zalloc(cells);
    for (size_t i = 0; i < cells; i++) m->flagcs[l][i] = 1.0f;
    cw /= 2; ch /= 2;
  }
}

static void mg_free(OrcMG *m) {
  for (int l = 0; l < m->levels; l++) {
    free(m->rs[l]); free(m->rcs[l]); free(m->ecs[l]); free(m->es[l]); free(m->flagcs[l]);
  }
  free(m->lw); free(m->lh); free(m->rs); free(m->rcs); free(m->ecs); free(m->es);
  free(m->flagcs); free(m->p); free(m->f); free(m->flag);
}

/* MG::updateFields, pressure_solver.hpp:34-57: threshold of the full-weighted
 * fine flag at 0.2 (double literal: float v promoted), borders stay 1.0. */
static void mg_update_fields(OrcMG *m, const float *flag) {
  memcpy(m->flagcs[0], flag, sizeof(float) * (size_t)m->width * m->height);
  for (int l = 1; l < m->levels; l++) {
    float *fc = m->flagcs[l];
    const float *fl = m->flagcs[l - 1];
    int wc = m->lw[l], hc = m->lh[l], wf = m->lw[l - 1];
    for (size_t i = 0; i < (size_t)wc * hc; i++) fc[i] = 1.0f;
    for (int y = 1; y < hc - 1; y++)
      for (int x = 1; x < wc - 1; x++) {
        float v = fw9(fl, wf, x, y);
        fc[IDX(x, y, wc)] = ((double)v > 0.2) ? 1.0f : 0.0f;
      }
  }
}

/* MG::solveLevel, pressure_solver.cpp:201-248 */
static void mg_solve_level(OrcMG *m, float *p, const float *f, const float *flag,
                           float hh, int level, int zgbc) {
  int w = m->lw[level], h = m->lh[level];
  if (level == m->levels - 2) {
    for (int i = 0; i < 5; i++) rbgs_once(p, f, flag, w, h, hh, 1.0f);
    return;
  }
  for (int i = 0; i < 3; i++) {
    rbgs_once(p, f, flag, w, h, hh, 1.0f);
    if (level == 0 && zgbc) orc_zero_gradient_bc(p, w, h);
  }
  float *r = m->rs[level];
  memset(r, 0, sizeof(float) * (size_t)w * h);
  orc_residual(p, f, flag, r, w, h, hh);

  int wc = m->lw[level + 1], hc = m->lh[level + 1];
  float *rc = m->rcs[level + 1];
  memset(rc, 0, sizeof(float) * (size_t)wc * hc);
  orc_restrict(r, w, h, rc, wc, hc);

  float *ec = m->ecs[level + 1];
  memset(ec, 0, sizeof(float) * (size_t)wc * hc);
  const float *flagc = m->flagcs[level + 1];

  /* :229  h * (r.width - 1.0f) / (rc.width - 1.0f); recursion drops the BC flag */
  mg_solve_level(m, ec, rc, flagc, hh * ((float)w - 1.0f) / ((float)wc - 1.0f),
                 level + 1, 0);

  float *e = m->es[level];
  orc_prolongate(e, w, h, ec, flagc, wc, hc, flag);
  orc_correct(p, e, w, h);
  if (level == 0 && zgbc) orc_zero_gradient_bc(p, w, h);
  for (int i = 0; i < 3; i++) {
    rbgs_once(p, f, flag, w, h, hh, 1.0f);
    if (level == 0 && zgbc) orc_zero_gradient_bc(p, w, h);
  }
}

void *orc_mg_create(int w, int h) {
  OrcMG *m = (OrcMG *)malloc(sizeof(OrcMG));
  mg_init(m, w, h);
  size_t n = (size_t)w * h;
  m->p = zalloc(n); m->f = zalloc(n); m->flag = zalloc(n);
  for (size_t i = 0; i < n; i++) m->flag[i] = 1.0f;
  return m;
}
void orc_mg_destroy(void *mg) { mg_free((OrcMG *)mg); free(mg); }
int orc_mg_levels(void *mg) { return ((OrcMG *)mg)->levels; }
void orc_mg_level_size(void *mg, int l, int *w, int *h) {
  *w = ((OrcMG *)mg)->lw[l]; *h = ((OrcMG *)mg)->lh[l];
}
void orc_mg_update_fields(void *mg, const float *flag) {
  OrcMG *m = (OrcMG *)mg;
  memcpy(m->flag, flag, sizeof(float) * (size_t)m->width * m->height);
  mg_update_fields(m, m->flag);
}
void orc_mg_get_flagc(void *mg, int l, float *dst) {
  OrcMG *m = (OrcMG *)mg;
  memcpy(dst, m->flagcs[l], sizeof(float) * (size_t)m->lw[l] * m->lh[l]);
}
void orc_mg_set(void *mg, const float *p, const float *f, const float *flag) {
  OrcMG *m = (OrcMG *)mg;
  size_t n = sizeof(float) * (size_t)m->width * m->height;
  if (p) memcpy(m->p, p, n);
  if (f) memcpy(m->f, f, n);
  if (flag) memcpy(m->flag, flag, n);
}
void orc_mg_get_p(void *mg, float *p) {
  OrcMG *m = (OrcMG *)mg;
  memcpy(p, m->p, sizeof(float) * (size_t)m->width * m->height);
}
void orc_mg_solve(void *mg, float hh, int zgbc) {
  OrcMG *m = (OrcMG *)mg;
  mg_solve_level(m, m->p, m->f, m->flag, hh, 0, zgbc);
}
float orc_mg_residual(void *mg, float hh) {
  OrcMG *m = (OrcMG *)mg;
  float *r = zalloc((size_t)m->width * m->height);
  float l2 = orc_residual(m->p, m->f, m->flag, r, m->width, m->height, hh);
  free(r);
  return l2;
}

/* ------------------------------------------------------------------------
 * simulation.{hpp,cpp}
 * ---------------------------------------------------------------------- */
typedef struct {
  float pwidth, mu, dt, h;
  int width, height;
  int bcW, bcE, bcN, bcS;
  float *vx[2], *vy[2];
  int vxf, vyf; /* front index; back = 1 - front (db2dgrid.hpp:64) */
  float *vx_accum, *vy_accum, *vx_current, *vy_current;
  float *p, *f, *flag, *r;
  OrcMG mg;
  float *sinks; /* xyz triples */
  int nsinks, capsinks;
} OrcSim;

#define VXW(s) ((s)->width - 1)
#define VXH(s) ((s)->height)
#define VYW(s) ((s)->width)
#define VYH(s) ((s)->height - 1)

/* Simulation(flag,pwidth,mu), simulation.hpp:32-67 */
void *orc_sim_create(const float *flag, int w, int h, float pwidth, float mu) {
  OrcSim *s = (OrcSim *)calloc(1, sizeof(OrcSim));
  s->pwidth = pwidth; s->mu = mu; s->width = w; s->height = h;
  size_t nx = (size_t)(w - 1) * h, ny = (size_t)w * (h - 1), n = (size_t)w * h;
  for (int b = 0; b < 2; b++) { s->vx[b] = zalloc(nx); s->vy[b] = zalloc(ny); }
  s->vx_accum = zalloc(nx); s->vy_accum = zalloc(ny);
  s->vx_current = zalloc(nx); s->vy_current = zalloc(ny);
  s->p = zalloc(n); s->f = zalloc(n); s->flag = zalloc(n); s->r = zalloc(n);
  memcpy(s->flag, flag, sizeof(float) * n);
  s->bcS = ORC_BC_NOSLIP; s->bcN = ORC_BC_NOSLIP;
  s->bcW = ORC_BC_INFLOW; s->bcE = ORC_BC_OUTFLOW_ZERO_PRESSURE;
  for (int y = 0; y < h; y++) /* :58-60 */
    s->vx[0][IDX(0, y, w - 1)] = s->vx[1][IDX(0, y, w - 1)] = 1.0f;
  mg_init(&s->mg, w, h);
  mg_update_fields(&s->mg, s->flag);
  s->h = pwidth / ((float)w - 1.0f);
  return s;
}

void orc_sim_destroy(void *sim) {
  OrcSim *s = (OrcSim *)sim;
  for (int b = 0; b < 2; b++) { free(s->vx[b]); free(s->vy[b]); }
  free(s->vx_accum); free(s->vy_accum); free(s->vx_current); free(s->vy_current);
  free(s->p); free(s->f); free(s->flag); free(s->r); free(s->sinks);
  mg_free(&s->mg);
  free(s);
}

static float *sim_field(OrcSim *s, int field, size_t *n) {
  size_t nx = (size_t)VXW(s) * VXH(s), ny = (size_t)VYW(s) * VYH(s),
         nc = (size_t)s->width * s->height;
  switch (field) {
  case ORC_FLAG: *n = nc; return s->flag;
  case ORC_VX: *n = nx; return s->vx[s->vxf];
  case ORC_VY: *n = ny; return s->vy[s->vyf];
  case ORC_VXB: *n = nx; return s->vx[1 - s->vxf];
  case ORC_VYB: *n = ny; return s->vy[1 - s->vyf];
  case ORC_P: *n = nc; return s->p;
  case ORC_F: *n = nc; return s->f;
  case ORC_VX_ACCUM: *n = nx; return s->vx_accum;
  case ORC_VY_ACCUM: *n = ny; return s->vy_accum;
  case ORC_R: *n = nc; return s->r;
  case ORC_VX_CURRENT: *n = nx; return s->vx_current;
  case ORC_VY_CURRENT: *n = ny; return s->vy_current;
  }
  return NULL;
}
int orc_sim_get(void *sim, int field, float *dst) {
  size_t n; float *src = sim_field((OrcSim *)sim, field, &n);
  if (!src) return -1;
  memcpy(dst, src, n * sizeof(float));
  return 0;
}
int orc_sim_set(void *sim, int field, const float *src) {
  size_t n; float *dst = sim_field((OrcSim *)sim, field, &n);
  if (!dst) return -1;
  memcpy(dst, src, n * sizeof(float));
  return 0;
}
void orc_sim_update_flag(void *sim, const float *flag) {
  OrcSim *s = (OrcSim *)sim;
  memcpy(s->flag, flag, sizeof(float) * (size_t)s->width * s->height);
  mg_update_fields(&s->mg, s->flag);
}
void orc_sim_set_bc(void *sim, int west, int east, int north, int south) {
  OrcSim *s = (OrcSim *)sim;
  s->bcW = west; s->bcE = east; s->bcN = north; s->bcS = south;
}
void orc_sim_add_sink(void *sim, float x, float y, float z) {
  OrcSim *s = (OrcSim *)sim;
  if (s->nsinks == s->capsinks) {
    s->capsinks = s->capsinks ? 2 * s->capsinks : 16;
    s->sinks = (float *)realloc(s->sinks, sizeof(float) * 3 * s->capsinks);
  }
  s->sinks[3 * s->nsinks] = x; s->sinks[3 * s->nsinks + 1] = y; s->sinks[3 * s->nsinks + 2] = z;
  s->nsinks++;
}
int orc_sim_num_sinks(void *sim) { return ((OrcSim *)sim)->nsinks; }
void orc_sim_get_sinks(void *sim, float *xyz) {
  OrcSim *s = (OrcSim *)sim;
  memcpy(xyz, s->sinks, sizeof(float) * 3 * s->nsinks);
}
float orc_sim_h(void *sim) { return ((OrcSim *)sim)->h; }
int orc_sim_mg_levels(void *sim) { return ((OrcSim *)sim)->mg.levels; }
void orc_sim_mg_level_size(void *sim, int l, int *w, int *h) {
  *w = ((OrcSim *)sim)->mg.lw[l]; *h = ((OrcSim *)sim)->mg.lh[l];
}
void orc_sim_mg_get_flagc(void *sim, int l, float *dst) {
  OrcSim *s = (OrcSim *)sim;
  memcpy(dst, s->mg.flagcs[l], sizeof(float) * (size_t)s->mg.lw[l] * s->mg.lh[l]);
}

/* singlePBC / setPBC, simulation.cpp:23-45 */
static float single_pbc(int bc, float a) {
  return bc == ORC_BC_OUTFLOW_ZERO_PRESSURE ? -a : a;
}
static void set_pbc(OrcSim *s) {
  int w = s->width, h = s->height;
  float *p = s->p;
  for (int y = 0; y < h; y++) {
    p[IDX(0, y, w)] = single_pbc(s->bcW, p[IDX(1, y, w)]);
    p[IDX(w - 1, y, w)] = single_pbc(s->bcE, p[IDX(w - 2, y, w)]);
  }
  for (int x = 0; x < w; x++) {
    p[IDX(x, 0, w)] = single_pbc(s->bcS, p[IDX(x, 1, w)]);
    p[IDX(x, h - 1, w)] = single_pbc(s->bcN, p[IDX(x, h - 2, w)]);
  }
}

/* VBCPar / VBCPer, simulation.cpp:50-78 */
static float vbc_par(int bc, float a, float b) {
  if (bc == ORC_BC_INFLOW) return b;
  if (bc == ORC_BC_OUTFLOW || bc == ORC_BC_OUTFLOW_ZERO_PRESSURE) return fmaxf(a, 0.0f);
  return 0.0f; /* NOSLIP */
}
static float vbc_per(int bc, float a, float b) {
  if (bc == ORC_BC_INFLOW) return b;
  if (bc == ORC_BC_OUTFLOW || bc == ORC_BC_OUTFLOW_ZERO_PRESSURE) return fmaxf(a, 0.0f);
  return -a; /* NOSLIP */
}

/* setVBCs, simulation.cpp:80-102: writes front AND back, columns then rows. */
static void set_vbcs(OrcSim *s) {
  float *xf = s->vx[s->vxf], *xb = s->vx[1 - s->vxf];
  float *yf = s->vy[s->vyf], *yb = s->vy[1 - s->vyf];
  int xw = VXW(s), xh = VXH(s), yw = VYW(s), yh = VYH(s);
  for (int y = 0; y < xh; y++) {
    xf[IDX(0, y, xw)] = xb[IDX(0, y, xw)] = vbc_par(s->bcW, xf[IDX(1, y, xw)], xf[IDX(0, y, xw)]);
    xf[IDX(xw - 1, y, xw)] = xb[IDX(xw - 1, y, xw)] =
        vbc_par(s->bcE, xf[IDX(xw - 2, y, xw)], xf[IDX(xw - 1, y, xw)]);
  }
  for (int x = 0; x < xw; x++) {
    xf[IDX(x, 0, xw)] = xb[IDX(x, 0, xw)] = vbc_per(s->bcS, xf[IDX(x, 1, xw)], xf[IDX(x, 0, xw)]);
    xf[IDX(x, xh - 1, xw)] = xb[IDX(x, xh - 1, xw)] =
        vbc_per(s->bcN, xf[IDX(x, xh - 2, xw)], xf[IDX(x, xh - 1, xw)]);
  }
  for (int y = 0; y < yh; y++) {
    yf[IDX(0, y, yw)] = yb[IDX(0, y, yw)] = vbc_per(s->bcW, yf[IDX(1, y, yw)], yf[IDX(0, y, yw)]);
    yf[IDX(yw - 1, y, yw)] = yb[IDX(yw - 1, y, yw)] =
        vbc_per(s->bcE, yf[IDX(yw - 2, y, yw)], yf[IDX(yw - 1, y, yw)]);
  }
  for (int x = 0; x < yw; x++) {
    yf[IDX(x, 0, yw)] = yb[IDX(x, 0, yw)] = vbc_par(s->bcS, yf[IDX(x, 1, yw)], yf[IDX(x, 0, yw)]);
    yf[IDX(x, yh - 1, yw)] = yb[IDX(x, yh - 1, yw)] =
        vbc_par(s->bcN, yf[IDX(x, yh - 2, yw)], yf[IDX(x, yh - 1, yw)]);
  }
}

/* applyAccumulatedVelocity, simulation.cpp:376-396 */
static void apply_accum(OrcSim *s) {
  float *vx = s->vx[s->vxf], *vy = s->vy[s->vyf];
  int xw = VXW(s), xh = VXH(s), yw = VYW(s), yh = VYH(s);
  for (int y = 1; y < xh - 1; y++)
    for (int x = 1; x < xw - 1; x++) {
      vx[IDX(x, y, xw)] += s->vx_accum[IDX(x, y, xw)];
      s->vx_accum[IDX(x, y, xw)] = 0;
    }
  for (int y = 1; y < yh - 1; y++)
    for (int x = 1; x < yw - 1; x++) {
      vy[IDX(x, y, yw)] += s->vy_accum[IDX(x, y, yw)];
      s->vy_accum[IDX(x, y, yw)] = 0;
    }
}

/* diffuse, simulation.cpp:104-162 */
static void diffuse(OrcSim *s) {
  int W = s->width;
  const float *flag = s->flag;
  float a = s->dt * s->mu * ((float)W - 1.0f) / s->pwidth;
  int xw = VXW(s), xh = VXH(s), yw = VYW(s), yh = VYH(s);
  for (int i = 1; i < 3; i++) {
    const float *vf = s->vx[s->vxf];
    float *vb = s->vx[1 - s->vxf];
#pragma omp parallel for schedule(static)
    for (int y = 1; y < xh - 1; y++)
      for (int x = 1; x < xw - 1; x++) {
        float c = vf[IDX(x, y, xw)];
        float val = 0;
        val += vf[IDX(x + 1, y, xw)] * flag[IDX(x + 1, y, W)] * flag[IDX(x + 2, y, W)];
        val += vf[IDX(x - 1, y, xw)] * flag[IDX(x, y, W)] * flag[IDX(x - 1, y, W)];
        float fvn = flag[IDX(x, y + 1, W)] * flag[IDX(x - 1, y + 1, W)];
        val += vf[IDX(x, y + 1, xw)] * fvn + (1.0f - fvn) * -c;
        float fvs = flag[IDX(x, y - 1, W)] * flag[IDX(x - 1, y - 1, W)];
        val += vf[IDX(x, y - 1, xw)] * fvs + (1.0f - fvs) * -c;
        vb[IDX(x, y, xw)] = flag[IDX(x, y, W)] * flag[IDX(x + 1, y, W)] * (c + a * val) /
                            (1.0f + 4.0f * a);
      }
    s->vxf = 1 - s->vxf; /* swap :131 */
    set_vbcs(s);
  }
  for (int i = 1; i < 3; i++) {
    const float *vf = s->vy[s->vyf];
    float *vb = s->vy[1 - s->vyf];
#pragma omp parallel for schedule(static)
    for (int y = 1; y < yh - 1; y++)
      for (int x = 1; x < yw - 1; x++) {
        float c = vf[IDX(x, y, yw)];
        float val = 0;
        val += vf[IDX(x, y - 1, yw)] * flag[IDX(x, y, W)] * flag[IDX(x, y - 1, W)];
        val += vf[IDX(x, y + 1, yw)] * flag[IDX(x, y + 1, W)] * flag[IDX(x, y + 2, W)];
        float fve = flag[IDX(x + 1, y, W)] * flag[IDX(x + 1, y + 1, W)];
        val += vf[IDX(x + 1, y, yw)] * fve + (1.0f - fve) * -c;
        float fvw = flag[IDX(x - 1, y, W)] * flag[IDX(x - 1, y + 1, W)];
        val += vf[IDX(x - 1, y, yw)] * fvw + (1.0f - fvw) * -c;
        vb[IDX(x, y, yw)] = flag[IDX(x, y, W)] * flag[IDX(x, y + 1, W)] * (c + a * val) /
                            (1.0f + 4.0f * a);
      }
    s->vyf = 1 - s->vyf;
    set_vbcs(s);
  }
}

/* CubicHermite (Catmull-Rom), interpolators.hpp:78-85 */
static inline float cubic_hermite(float t, float A, float B, float C, float D) {
  float a = -A / 2.0f + (3.0f * B) / 2.0f - (3.0f * C) / 2.0f + D / 2.0f;
  float b = A - (5.0f * B) / 2.0f + 2.0f * C - D / 2.0f;
  float c = -A / 2.0f + C / 2.0f;
  float d = B;
  return a * t * t * t + b * t * t + c * t + d;
}

/* bicubicSample, interpolators.hpp:92-206: clamp to [3, w-3] x [3, h-3], 4x4
 * taps at rows iy-1..iy+2 / cols ix-1..ix+2 through FLAT indices (:106-111),
 * vertical Hermite per column first (:132-193), then horizontal (:204). */
static inline float bicubic(const float *g, int w, int h, float cx, float cy) {
  cx = fmaxf(fminf(cx, (float)w - 3.0f), 3.0f);
  cy = fmaxf(fminf(cy, (float)h - 3.0f), 3.0f);
  int icx = (int)cx, icy = (int)cy;
  float stx = cx - truncf(cx), sty = cy - truncf(cy);
  long i1 = (long)(icx - 1) + (long)(icy - 1) * w;
  long i2 = i1 + w, i3 = i2 + w, i4 = i3 + w;
  float col[4];
  for (int k = 0; k < 4; k++)
    col[k] = cubic_hermite(sty, g[i1 + k], g[i2 + k], g[i3 + k], g[i4 + k]);
  return cubic_hermite(stx, col[0], col[1], col[2], col[3]);
}

/* advect, simulation.cpp:241-354.  The 8-wide octet structure is observable
 * (whole-octet skip test :254-256 / :302-304, loop bound x < width-8 :248/:300,
 * flat loads that run into the next row), so it is restated lane by lane. */
static void advect(OrcSim *s) {
  float ih = 1.0f / s->h;
  float dt = s->dt;
  int W = s->width;
  const float *flag = s->flag;
  const float *vx = s->vx[s->vxf], *vy = s->vy[s->vyf];
  float *vxb = s->vx[1 - s->vxf], *vyb = s->vy[1 - s->vyf];
  int xw = VXW(s), xh = VXH(s), yw = VYW(s), yh = VYH(s);
  float half = 0.5f * dt * ih, full = dt * ih;
#pragma omp parallel for schedule(dynamic, 16)
  for (int y = 1; y < xh - 1; y++) {
    for (int x = 1; x < xw - 8; x += 8) {
      int any = 0;
      for (int i = 0; i < 8; i++)
        if (flag[IDX(x - 1 + i, y, W)] + flag[IDX(x + i, y, W)] == 2.0f) any = 1;
      if (!any) continue;
      for (int i = 0; i < 8; i++) {
        int xi = x + i;
        float posx = (float)xi + 0.5f, posy = (float)y;
        float vx1 = vx[IDX(xi, y, xw)];
        float vy1 = ((vy[IDX(xi, y, yw)] + vy[IDX(xi, y - 1, yw)]) +
                     (vy[IDX(xi + 1, y, yw)] + vy[IDX(xi + 1, y - 1, yw)])) * 0.25f;
        float midx = posx - vx1 * half, midy = posy - vy1 * half;
        float vx2 = bicubic(vx, xw, xh, midx - 0.5f, midy);
        float vy2 = bicubic(vy, yw, yh, midx, midy - 0.5f);
        float endx = posx - vx2 * full, endy = posy - vy2 * full;
        float xvel = bicubic(vx, xw, xh, endx - 0.5f, endy);
        xvel = xvel * flag[IDX(xi, y, W)] * flag[IDX(xi + 1, y, W)];
        vxb[IDX(xi, y, xw)] = xvel;
      }
    }
    for (int x = 1; x < yw - 8; x += 8) {
      int any = 0;
      for (int i = 0; i < 8; i++)
        if (flag[IDX(x + i, y, W)] + flag[IDX(x + i, y - 1, W)] == 2.0f) any = 1;
      if (!any) continue;
      for (int i = 0; i < 8; i++) {
        int xi = x + i;
        float posy = (float)y + 0.5f, posx = (float)xi;
        float vy1 = vy[IDX(xi, y, yw)];
        /* flat vx loads (:317-325): xi+1 may equal vx.width, i.e. wrap into
         * the first element of the next row -- kept as in the reference */
        float vx1 = ((vx[IDX(xi, y, xw)] + vx[IDX(xi, y - 1, xw)]) +
                     (vx[IDX(xi + 1, y, xw)] + vx[IDX(xi + 1, y - 1, xw)])) * 0.25f;
        float midx = posx - vx1 * half, midy = posy - vy1 * half;
        float vx2 = bicubic(vx, xw, xh, midx - 0.5f, midy);
        float vy2 = bicubic(vy, yw, yh, midx, midy - 0.5f);
        float endx = posx - vx2 * full, endy = posy - vy2 * full;
        float yvel = bicubic(vy, yw, yh, endx, endy - 0.5f);
        yvel = yvel * flag[IDX(xi, y, W)] * flag[IDX(xi, y + 1, W)];
        vyb[IDX(xi, y, yw)] = yvel;
      }
    }
  }
  s->vxf = 1 - s->vxf;
  s->vyf = 1 - s->vyf;
}

/* project, simulation.cpp:164-208 */
static void project(OrcSim *s) {
  int W = s->width, H = s->height;
  float ih = 1.0f / s->h;
  float *vx = s->vx[s->vxf], *vy = s->vy[s->vyf];
  int xw = VXW(s), yw = VYW(s);
  float *f = s->f, *p = s->p;
  const float *flag = s->flag;
#pragma omp parallel for schedule(static)
  for (int y = 1; y < H - 1; y++)
    for (int x = 1; x < W - 1; x++)
      f[IDX(x, y, W)] = -ih * (vx[IDX(x, y, xw)] - vx[IDX(x - 1, y, xw)] +
                               vy[IDX(x, y, yw)] - vy[IDX(x, y - 1, yw)]);

  /* sinks :173-187 (skipped sinks neither stamp nor decay) */
  for (int k = 0; k < s->nsinks; k++) {
    float *sk = s->sinks + 3 * k;
    float gx = sk[0] / s->h + 0.5f, gy = sk[1] / s->h + 0.5f;
    if (gx <= 3 || gx > (float)(W - 3) || gy <= 3 || gy > (float)(H - 3)) continue;
    for (int y = -1; y <= 1; y++)
      for (int x = -1; x <= 1; x++)
        f[IDX((int)(gx + (float)x), (int)(gy + (float)y), W)] = sk[2];
    sk[2] = (float)((double)sk[2] * pow(0.000001, (double)(s->dt * 50)));
  }
  int n = 0;
  for (int k = 0; k < s->nsinks; k++)
    if (!(s->sinks[3 * k + 2] < 0.05f)) {
      memmove(s->sinks + 3 * n, s->sinks + 3 * k, 3 * sizeof(float));
      n++;
    }
  s->nsinks = n;

  mg_solve_level(&s->mg, p, f, flag, s->h, 0, 1);
  mg_solve_level(&s->mg, p, f, flag, s->h, 0, 1);
  set_pbc(s);

#pragma omp parallel for schedule(static)
  for (int y = 1; y < H - 1; y++)
    for (int x = 1; x < W - 2; x++)
      vx[IDX(x, y, xw)] -= flag[IDX(x, y, W)] * flag[IDX(x + 1, y, W)] * ih *
                           (p[IDX(x + 1, y, W)] - p[IDX(x, y, W)]);
#pragma omp parallel for schedule(static)
  for (int y = 1; y < H - 2; y++)
    for (int x = 1; x < W - 1; x++)
      vy[IDX(x, y, yw)] -= flag[IDX(x, y, W)] * flag[IDX(x, y + 1, W)] * ih *
                           (p[IDX(x, y + 1, W)] - p[IDX(x, y, W)]);
}

/* saveCurrentVelocityFields, simulation.cpp:16-19 */
static void save_current(OrcSim *s) {
  memcpy(s->vx_current, s->vx[s->vxf], sizeof(float) * (size_t)VXW(s) * VXH(s));
  memcpy(s->vy_current, s->vy[s->vyf], sizeof(float) * (size_t)VYW(s) * VYH(s));
}

void orc_sim_stage(void *sim, int stage, float dt) {
  OrcSim *s = (OrcSim *)sim;
  s->dt = dt;
  switch (stage) {
  case ORC_ST_ACCUM: apply_accum(s); break;
  case ORC_ST_DIFFUSE: diffuse(s); break;
  case ORC_ST_ADVECT: advect(s); break;
  case ORC_ST_SETVBCS: set_vbcs(s); break;
  case ORC_ST_PROJECT: project(s); break;
  case ORC_ST_SAVE: save_current(s); break;
  }
}

/* step, simulation.cpp:356-374 */
void orc_sim_step(void *sim, float dt) {
  OrcSim *s = (OrcSim *)sim;
  s->dt = dt;
  apply_accum(s);
  diffuse(s);
  advect(s);
  set_vbcs(s);
  project(s);
  set_vbcs(s);
  save_current(s);
}
