/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference fluid-step path.
 *
 * Plain-C, scalar, unpadded row-major (idx = y*width + x, db2dgrid.hpp:19)
 * restatement of te42kyfo/ubootgl's Simulation::step and MG::solve.  Each
 * function in ubgl_oracle.c cites the reference file:line it follows.
 *
 * PARITY PINNED: tests/test_oracle_vs_reference.py checks every function here
 * against the unmodified reference TUs compiled into oracle/_ref (when
 * /root/reference is present) and tests/test_oracle_golden.py checks it
 * against the committed fixtures in tests/golden/ that were generated from
 * those same TUs (tests/golden/make_golden.py), including the survey's
 * known-answer vectors (mgtest residual history, game-level field norms).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * link or call this.  The product path (ubootgl_b200/) must not.
 */
#ifndef UBGL_ORACLE_H
#define UBGL_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_FLAG = 0, ORC_VX, ORC_VY, ORC_VXB, ORC_VYB, ORC_P, ORC_F,
       ORC_VX_ACCUM, ORC_VY_ACCUM, ORC_R, ORC_VX_CURRENT, ORC_VY_CURRENT };
enum { ORC_ST_ACCUM = 0, ORC_ST_DIFFUSE, ORC_ST_ADVECT, ORC_ST_SETVBCS,
       ORC_ST_PROJECT, ORC_ST_SAVE };
enum { ORC_BC_INFLOW = 0, ORC_BC_OUTFLOW, ORC_BC_OUTFLOW_ZERO_PRESSURE,
       ORC_BC_NOSLIP };

void orc_set_threads(int n);
int orc_max_threads(void);
int orc_num_procs(void);

/* pressure_solver.cpp free functions */
void orc_rbgs(float *p, const float *f, const float *flag, int w, int h,
              float hh, float alpha, int sweeps);
float orc_residual(const float *p, const float *f, const float *flag, float *r,
                   int w, int h, float hh);
void orc_restrict(const float *r, int w, int h, float *rc, int wc, int hc);
void orc_prolongate(float *e, int w, int h, const float *ec,
                    const float *flagc, int wc, int hc, const float *flag);
void orc_correct(float *p, const float *e, int w, int h);
void orc_zero_gradient_bc(float *p, int w, int h);

/* class MG */
void *orc_mg_create(int w, int h);
void orc_mg_destroy(void *mg);
int orc_mg_levels(void *mg);
void orc_mg_level_size(void *mg, int level, int *w, int *h);
void orc_mg_update_fields(void *mg, const float *flag);
void orc_mg_get_flagc(void *mg, int level, float *dst);
void orc_mg_set(void *mg, const float *p, const float *f, const float *flag);
void orc_mg_get_p(void *mg, float *p);
void orc_mg_solve(void *mg, float hh, int zero_gradient_bc);
float orc_mg_residual(void *mg, float hh);

/* class Simulation */
void *orc_sim_create(const float *flag, int w, int h, float pwidth, float mu);
void orc_sim_destroy(void *sim);
int orc_sim_get(void *sim, int field, float *dst);
int orc_sim_set(void *sim, int field, const float *src);
void orc_sim_update_flag(void *sim, const float *flag);
void orc_sim_set_bc(void *sim, int west, int east, int north, int south);
void orc_sim_add_sink(void *sim, float x, float y, float z);
int orc_sim_num_sinks(void *sim);
void orc_sim_get_sinks(void *sim, float *xyz);
float orc_sim_h(void *sim);
void orc_sim_step(void *sim, float dt);
void orc_sim_stage(void *sim, int stage, float dt);
int orc_sim_mg_levels(void *sim);
void orc_sim_mg_level_size(void *sim, int level, int *w, int *h);
void orc_sim_mg_get_flagc(void *sim, int level, float *dst);

/* ---- callers either side of the step (ubgl_oracle_next.c) ---- */
/* CoItem + CoKinematicsSimple (components.hpp:6-43), one record per item */
typedef struct orc_item {
  float size[2], pos[2], rotation;       /* CoItem */
  float mass, vel[2], force[2], angVel, angForce; /* CoKinematicsSimple */
  int bumpCount;
} orc_item;
void orc_colocate(const float *vx, const float *vy, int nx, int ny, float *vxy, float *mag);
void orc_tracers_advect(float *points, unsigned *start, unsigned *end, float *ages, int ntracers,
                        int npoints, float dt, float pdx, float pdy, unsigned rand_seed,
                        const float *vxy, int tw, int th, const float *flagtex, int fw, int fh);
void orc_tracers_shift(float *points, int ntracers, int npoints, float shift);
void orc_items_advect_simple(orc_item *items, int n, float game_dt, const float *flag, const float *vx,
                             const float *vy, const float *p, float *vx_accum, float *vy_accum, int W,
                             int H, float pwidth);
void orc_items_advect(orc_item *items, int n, float game_dt, const float *flag, const float *vx,
                      const float *vy, float *vx_accum, float *vy_accum, int W, int H, float pwidth);
void orc_draw_circle(float *flag_full, float *flag_sim, int w, int h, float cx, float cy, int diam,
                     float val);
void orc_set_grids_all(float *flag, float *vx, float *vy, float *p, const float *newflag, int W, int H);
void orc_shift_map(float *flag, float *vxf, float *vxb, float *vyf, float *vyb, float *p, float *vxc,
                   float *vyc, const float *newflag, int W, int H);

#ifdef __cplusplus
}
#endif
#endif
