/* ubgl.h -- C ABI of the B200-native fluid-step hot path (libubgl.so).
 *
 * Drop-in boundary for te42kyfo/ubootgl's `Simulation` / `MG` solver classes:
 * the repo's own C++ mirror of those classes (ubootgl_b200/host/) and any
 * other binding (ctypes in ubootgl_b200/capi.py) call only what is declared
 * here.  Plain C types, opaque handles, `int` status: 0 = ok, negative = error
 * (UBGL_E_*), text via ubgl_last_error().  No exceptions cross the boundary.
 *
 * Conventions (all from the reference, file:line under te42kyfo/ubootgl):
 *  - host grids are UNPADDED row-major fp32, idx = y*width + x (db2dgrid.hpp:19);
 *    staggered sizes: vx (W-1)xH, vy Wx(H-1), cell fields WxH
 *    (simulation.hpp:22-29).  The library re-pitches them on the device.
 *  - a handle is used by one host thread at a time and owns its own CUDA stream;
 *    host pointers are caller-owned and only touched during the call.
 *  - there is NO CPU fallback: every entry point that computes fails with
 *    UBGL_E_CUDA when no sm_100 device is usable.
 */
#ifndef UBGL_H
#define UBGL_H

#ifdef __cplusplus
extern "C" {
#endif

#define UBGL_VERSION 100

/* status codes */
#define UBGL_OK 0
#define UBGL_E_ARG (-1)    /* bad argument (null handle, size, field id, ...) */
#define UBGL_E_CUDA (-2)   /* CUDA runtime error, see ubgl_last_error() */
#define UBGL_E_STATE (-3)  /* call not valid in the handle's current state */
#define UBGL_E_NOMEM (-4)

/* field ids -- the public data members of Simulation (simulation.hpp:128-132) */
enum ubgl_field {
  UBGL_FLAG = 0,        /* Simulation::flag            W x H     */
  UBGL_VX = 1,          /* Simulation::vx front buffer (W-1) x H */
  UBGL_VY = 2,          /* Simulation::vy front buffer W x (H-1) */
  UBGL_VXB = 3,         /* Simulation::vx back buffer  (db2dgrid.hpp:103) */
  UBGL_VYB = 4,         /* Simulation::vy back buffer  */
  UBGL_P = 5,           /* Simulation::p               W x H */
  UBGL_F = 6,           /* Simulation::f (divergence rhs) */
  UBGL_VX_ACCUM = 7,    /* Simulation::vx_accum        */
  UBGL_VY_ACCUM = 8,    /* Simulation::vy_accum        */
  UBGL_R = 9,           /* residual field of the last ubgl_sim_residual() */
  UBGL_VX_CURRENT = 10, /* Simulation::vx_current      */
  UBGL_VY_CURRENT = 11, /* Simulation::vy_current      */
  UBGL_NUM_FIELDS = 12
};

/* Simulation::BC (simulation.hpp:69), same numeric order */
enum ubgl_bc {
  UBGL_BC_INFLOW = 0,
  UBGL_BC_OUTFLOW = 1,
  UBGL_BC_OUTFLOW_ZERO_PRESSURE = 2,
  UBGL_BC_NOSLIP = 3
};

/* stages of Simulation::step (simulation.cpp:356-374) for stage-level parity */
enum ubgl_stage {
  UBGL_ST_ACCUM = 0,   /* applyAccumulatedVelocity  simulation.cpp:376-396 */
  UBGL_ST_DIFFUSE = 1, /* diffuse                   simulation.cpp:104-162 */
  UBGL_ST_ADVECT = 2,  /* advect                    simulation.cpp:241-354 */
  UBGL_ST_SETVBCS = 3, /* setVBCs                   simulation.cpp:80-102  */
  UBGL_ST_PROJECT = 4, /* project                   simulation.cpp:164-208 */
  UBGL_ST_SAVE = 5     /* saveCurrentVelocityFields simulation.cpp:16-19   */
};

/* tunables (ubgl_sim_set_option / ubgl_mg_set_option) */
enum ubgl_option {
  UBGL_OPT_VCYCLES = 0,  /* V-cycles per project(); default 2 (simulation.cpp:189-190) */
  UBGL_OPT_FUSED = 1,    /* non-zero (default 2): fused / temporally blocked kernels -- 2: multigrid
                            passes with register-resident runs (k_mg_run), 1: the shared-memory
                            tile schedule (k_mg_tile; the choice between 1 and 2 is process wide);
                            0: one plain kernel per reference stage.  Same results in all three. */
  UBGL_OPT_GRAPH = 2,    /* 1 (default): grids up to 4 M cells replay CUDA graphs of the fused step's two launch
                            sequences (prestep..divergence, V-cycles..save), captured from the handle's stream
                            once a buffer-role state (and dt) has repeated; larger grids are GPU-bound and are
                            always launched directly.  0: never use graphs.  Same results either way. */
  UBGL_OPT_TIMING = 3    /* 1: record per-stage CUDA events (ubgl_sim_stage_ms) */
};

typedef struct ubgl_sim ubgl_sim_t; /* device-resident Simulation state */
typedef struct ubgl_mg ubgl_mg_t;   /* device-resident stand-alone MG   */

/* ---- library -------------------------------------------------------------- */
int ubgl_version(void);
const char *ubgl_last_error(void); /* thread-local text of the last failure */
int ubgl_device_count(void);       /* usable CUDA devices, 0 if none */

/* ---- class Simulation (simulation.hpp:18-141) ----------------------------- */
/* Simulation(flag, pwidth, mu) simulation.hpp:32-67: copies flag, BCs
 * W=INFLOW E=OUTFLOW_ZERO_PRESSURE N,S=NOSLIP, vx(0,y)=1 on both buffers,
 * builds the MG flag pyramid, h = pwidth/(W-1).  Requires W,H >= 8. */
int ubgl_sim_create(const float *flag, int W, int H, float pwidth, float mu,
                    int device, ubgl_sim_t **out);
int ubgl_sim_destroy(ubgl_sim_t *sim);
int ubgl_sim_set_option(ubgl_sim_t *sim, int option, int value);
/* members bcWest/bcEast/bcNorth/bcSouth (simulation.hpp:124) */
int ubgl_sim_set_bc(ubgl_sim_t *sim, int west, int east, int north, int south);
/* whole-field H->D / D->H of one public member; host layout as the reference */
int ubgl_sim_upload(ubgl_sim_t *sim, int field, const float *host);
int ubgl_sim_download(ubgl_sim_t *sim, int field, float *host);
/* field += host grid.  For vx_accum / vy_accum, which both sides of this boundary add into:
 * the items kernels on the device (atomicAdd) and a host caller running the reference's own
 * advect_floating_items.cpp:118-120 on its mirrors (`sim.vx_accum(...) += ...`). */
int ubgl_sim_upload_add(ubgl_sim_t *sim, int field, const float *host);
/* memcpy(sim.flag.data(), ...) + sim.mg.updateFields(sim.flag)
 * (ubootgl_app.cpp:111-112, pressure_solver.hpp:34-57) */
int ubgl_sim_update_flag(ubgl_sim_t *sim, const float *flag);
/* coarse flag mask of MG level `level` (flagcs[level]); sizes via
 * ubgl_sim_mg_level_size.  Bit-exact with the reference. */
int ubgl_sim_mg_levels(ubgl_sim_t *sim);
int ubgl_sim_mg_level_size(ubgl_sim_t *sim, int level, int *w, int *h);
int ubgl_sim_mg_get_flagc(ubgl_sim_t *sim, int level, float *host);
/* member `sinks` (simulation.hpp:140): xyz triples in physical coordinates as
 * pushed by explosion.cpp:33; stamped into f, decayed and erased inside
 * project exactly as simulation.cpp:173-187.  get returns the survivors. */
int ubgl_sim_set_sinks(ubgl_sim_t *sim, const float *xyz, int n);
int ubgl_sim_get_sinks(ubgl_sim_t *sim, float *xyz, int cap, int *n);
/* Simulation::step(dt) (simulation.cpp:356-374) on the device-resident state.
 * Asynchronous on the handle's stream; ubgl_sim_sync() or a download waits. */
int ubgl_sim_step(ubgl_sim_t *sim, float dt);
/* One stage of step() (enum ubgl_stage), for parity tests. */
int ubgl_sim_stage(ubgl_sim_t *sim, int stage, float dt);
/* step() as the reference's callers see it, with HOST mirrors: uploads flag
 * (if non-null; rebuilds the pyramid), vx_accum/vy_accum (if non-null; the
 * host arrays are zeroed like simulation.cpp:384,392), runs step(dt), then
 * downloads vx, vy, p, vx_current, vy_current (each if non-null). */
typedef struct ubgl_host_mirrors {
  const float *flag;   /* in, optional */
  float *vx_accum;     /* in (then zeroed), optional */
  float *vy_accum;     /* in (then zeroed), optional */
  float *vx, *vy, *p;  /* out, optional */
  float *vx_current;   /* out, optional */
  float *vy_current;   /* out, optional */
} ubgl_host_mirrors;
int ubgl_sim_step_host(ubgl_sim_t *sim, float dt, const ubgl_host_mirrors *m);
/* Optional pipelined form of the same call (m->flag must be NULL): the outputs of step n-1 cross
 * PCIe downwards while the accumulators of step n cross it upwards and step n runs, so the
 * mirrors written by call n are those of step n-1 -- ONE STEP LATE, otherwise bit-identical to
 * ubgl_sim_step_host's.  This is the coherence the reference's render thread already has with its
 * simulation thread (draw.cpp:101 reads sim.p while sim_loop.cpp:29 is inside step()).  The first
 * call writes no mirrors; ubgl_sim_step_host_flush brings the last step's fields down. */
int ubgl_sim_step_host_pipelined(ubgl_sim_t *sim, float dt, const ubgl_host_mirrors *m);
int ubgl_sim_step_host_flush(ubgl_sim_t *sim, const ubgl_host_mirrors *m);
/* Stop rule of the pressure solves inside step() / stage(PROJECT).  rel_tol <= 0
 * (default): the reference's fixed UBGL_OPT_VCYCLES warm-started V-cycles, no
 * convergence test (simulation.cpp:189-190).  rel_tol > 0: V-cycles until
 * ||r||_2 <= rel_tol * ||f*flag||_2 (calculateResidualField norm,
 * pressure_solver.cpp:91-116, relative to the residual of p = 0), or until a cycle
 * leaves more than `stagnation` (e.g. 0.9) of the previous residual -- channel flows
 * stagnate far above 1e-4 -- or max_cycles.  Costs one residual
 * kernel + one host sync per cycle. */
int ubgl_sim_set_tolerance(ubgl_sim_t *sim, float rel_tol, int max_cycles, float stagnation);
/* what the last step()'s solves did: V-cycles run, ||f*flag||, and (tolerance mode)
 * ||r|| before the first and after every cycle */
int ubgl_sim_solve_info(ubgl_sim_t *sim, int *cycles_done, float *fnorm, float *res_hist, int cap,
                        int *n_hist);
int ubgl_sim_sync(ubgl_sim_t *sim);
/* calculateResidualField(p, f, flag, r, h) (pressure_solver.cpp:91-116) on the
 * resident p/f/flag; r readable as field UBGL_R. */
int ubgl_sim_residual(ubgl_sim_t *sim, float *l2);
/* MG::solve(p, f, flag, h, true) on the resident fields, `cycles` times */
int ubgl_sim_mg_solve(ubgl_sim_t *sim, int cycles);
/* MG::solve(p, f, flag, h, zeroGradientBC) with an explicit grid spacing and BC
 * switch, on the resident fields (Simulation::mg used as a stand-alone solver) */
int ubgl_sim_mg_solve_ex(ubgl_sim_t *sim, float h, int zero_gradient_bc, int cycles);
/* device address + pitch (in floats) of a resident field, for zero-copy
 * producers/consumers (tracers, interop, benchmark input generation).  The
 * velocity buffers rotate roles inside step(): query again after every step.
 * After a fused step vx_current / vy_current alias the front buffers (they are byte
 * copies of them, simulation.cpp:16-19, and the copy is not made until somebody is
 * about to write either side); asking for the pointer of vx, vy, vx_current or
 * vy_current makes that copy first, so writes through it behave as in the reference. */
int ubgl_sim_device_ptr(ubgl_sim_t *sim, int field, void **dptr, int *pitch);
/* per-stage device time of the last step when UBGL_OPT_TIMING=1 */
int ubgl_sim_stage_ms(ubgl_sim_t *sim, int stage, float *ms);
/* number of kernels this handle launched so far (graph replays included) */
long long ubgl_sim_launch_count(ubgl_sim_t *sim);
/* the CUDA stream (cudaStream_t) the handle launches on */
void *ubgl_sim_stream(ubgl_sim_t *sim);
/* per-kernel profile: ubgl_sim_profile(sim, 1) clears the statistics and
 * brackets every subsequent launch with CUDA events on the handle's stream
 * (graph replay is bypassed while on); ubgl_sim_profile(sim, 0) stops.
 * ubgl_sim_kernel_stats reads launches / summed device ms of one kernel kind at
 * one MG level (level 0 for non-MG kernels). */
int ubgl_sim_profile(ubgl_sim_t *sim, int on);
int ubgl_num_kernel_kinds(void);
const char *ubgl_kernel_kind_name(int kind);
int ubgl_sim_kernel_stats(ubgl_sim_t *sim, int kind, int level, long long *count, double *ms);

/* ---- class MG (pressure_solver.hpp:13-76) ---------------------------------- */
/* MG(int w, int h): level pyramid by integer halving while w>3 && h>3, all
 * coarse flags 1.0 (pressure_solver.hpp:16-31).  Requires >= 2 levels. */
int ubgl_mg_create(int W, int H, int device, ubgl_mg_t **out);
int ubgl_mg_destroy(ubgl_mg_t *mg);
int ubgl_mg_set_option(ubgl_mg_t *mg, int option, int value);
int ubgl_mg_levels(ubgl_mg_t *mg);
int ubgl_mg_level_size(ubgl_mg_t *mg, int level, int *w, int *h);
/* MG::updateFields(flag) (pressure_solver.hpp:34-57) */
int ubgl_mg_update_fields(ubgl_mg_t *mg, const float *flag);
int ubgl_mg_get_flagc(ubgl_mg_t *mg, int level, float *host);
/* MG::solve(p, f, flag, h, zeroGradientBC) with host grids
 * (pressure_solver.hpp:59-62): uploads p, f, flag, runs one V-cycle, downloads p.
 * As in the reference, level 0 uses the caller's flag and levels >= 1 the
 * pyramid from the last updateFields. */
int ubgl_mg_solve_host(ubgl_mg_t *mg, float *p, const float *f,
                       const float *flag, float h, int zero_gradient_bc);
/* resident variant: upload once, cycle many times, download */
int ubgl_mg_upload(ubgl_mg_t *mg, const float *p, const float *f,
                   const float *flag); /* each optional */
int ubgl_mg_download_p(ubgl_mg_t *mg, float *p);
int ubgl_mg_solve(ubgl_mg_t *mg, float h, int zero_gradient_bc, int cycles);
int ubgl_mg_residual(ubgl_mg_t *mg, float h, float *l2);
int ubgl_mg_sync(ubgl_mg_t *mg);
long long ubgl_mg_launch_count(ubgl_mg_t *mg);
void *ubgl_mg_stream(ubgl_mg_t *mg);

/* ---- Simulation::step row-slab decomposed over the GPUs of one box ---------- */
/* (new).  One process (or thread) per GPU, each with one handle.
 * GPU r owns the cell rows [own_lo, own_hi) and stores [st_lo, st_hi) (own rows
 * plus ghost rows); upload/download move the STORED rows of a field, unpadded
 * (w x nrows as ubgl_slab_field_rows reports).  Bootstrap: every rank creates its
 * handle, the ranks exchange their ubgl_slab_ipc_export() blobs by any means
 * (bench.py: torch.distributed all_gather), then every rank calls
 * ubgl_slab_connect() with all blobs -- collective: it also builds the coarse
 * flag pyramid, which needs the neighbours' rows.  ubgl_slab_step is
 * Simulation::step(dt) (simulation.cpp:356-374); halo rows move by peer stores
 * over NVLink inside the step, without host synchronisation. */
typedef struct ubgl_slab ubgl_slab_t;
/* pure host arithmetic, no CUDA: the decomposition a (W,H,nranks) run uses */
int ubgl_slab_plan(int W, int H, int nranks, int rank, int *dist_levels, int *ghost, int *own_lo,
                   int *own_hi, int *st_lo, int *st_hi);
/* Load balance: relative cost of every level-0 row (H positive floats; NULL: equal rows) for the
 * plans made from now on in this process -- ubgl_slab_plan and ubgl_slab_create place the cuts at
 * equal weight per rank (still multiples of 2^dist_levels).  The reference's advect skips octets
 * without fluid (simulation.cpp:254-256), so a caller weights a row by its fluid fraction.  Every
 * rank must set the same weights. */
int ubgl_slab_set_row_weights(const float *weights, int H);
int ubgl_slab_create(const float *flag_stored_rows, int W, int H, float pwidth, float mu, int device,
                     int rank, int nranks, ubgl_slab_t **out);
int ubgl_slab_destroy(ubgl_slab_t *s);
int ubgl_slab_ipc_size(void); /* bytes of one export blob */
int ubgl_slab_ipc_export(ubgl_slab_t *s, void *blob);
int ubgl_slab_connect(ubgl_slab_t *s, const void *blobs /* nranks blobs, indexed by rank */);
int ubgl_slab_field_rows(ubgl_slab_t *s, int field, int *row_lo, int *nrows, int *w);
int ubgl_slab_upload(ubgl_slab_t *s, int field, const float *host);
int ubgl_slab_download(ubgl_slab_t *s, int field, float *host);
int ubgl_slab_set_option(ubgl_slab_t *s, int option, int value); /* UBGL_OPT_VCYCLES */
int ubgl_slab_set_sinks(ubgl_slab_t *s, const float *xyz, int n);
int ubgl_slab_step(ubgl_slab_t *s, float dt);
/* ubgl_sim_step_host for one rank's slab (Simulation::step as sim_loop.cpp:29's caller sees it,
 * each rank over its own PCIe link): vx_accum / vy_accum mirrors in (arrays over the STORED rows,
 * ubgl_slab_field_rows; their own rows are cleared like simulation.cpp:384,392), then the OWN
 * rows of vx, vy, p, vx_current, vy_current are written at their place inside arrays over the
 * stored rows.  m->flag must be NULL. */
int ubgl_slab_step_host(ubgl_slab_t *s, float dt, const ubgl_host_mirrors *m);
int ubgl_slab_sync(ubgl_slab_t *s);
/* sum of r^2 over the own rows (calculateResidualField); add the ranks, take sqrt */
int ubgl_slab_residual_sumsq(ubgl_slab_t *s, double *sumsq);
long long ubgl_slab_launch_count(ubgl_slab_t *s);
void *ubgl_slab_stream(ubgl_slab_t *s);
/* halo exchanges issued and bytes pushed to peers so far */
int ubgl_slab_stats(ubgl_slab_t *s, long long *exchanges, long long *halo_bytes);
/* per-kernel profile of this rank, as ubgl_sim_profile / ubgl_sim_kernel_stats; an exchange
 * is ONE launch of kind "halo_push": peer stores, release of the sequence number, then the
 * wait for the neighbours' releases (so its time = NVLink stores + load imbalance + latency;
 * the kind "halo_wait" is kept for the stand-alone wait kernel, unused by step) */
int ubgl_slab_profile(ubgl_slab_t *s, int on);
int ubgl_slab_kernel_stats(ubgl_slab_t *s, int kind, int level, long long *count, double *ms);

/* ---- callers either side of the step, on the device ---------- */
/* (1) Fluid tracers.  The reference runs them as GLSL compute shaders on
 * GL textures built from vx_current / vy_current and the full-resolution flag
 * (velocity_textures.cpp:63-101, draw_tracers_cs.cpp:131-156).  Sampling follows
 * the OpenGL texture-object defaults those textures are created with: GL_LINEAR
 * magnification, GL_REPEAT wrap, level 0, evaluated in exact fp32. */
typedef struct ubgl_tracers ubgl_tracers_t;
/* GLTracers::init(npoints, ntracers) (draw_tracers_cs.cpp:29-79): points and ring
 * pointers zero, ages 2*3.1 (every tracer respawns on its first advect) */
int ubgl_tracers_create(int ntracers, int npoints, int device, ubgl_tracers_t **out);
int ubgl_tracers_destroy(ubgl_tracers_t *t);
/* the SSBOs: points [ntracers][npoints] vec2, start/end ring pointers, ages; each optional */
int ubgl_tracers_upload(ubgl_tracers_t *t, const float *points, const unsigned *start,
                        const unsigned *end, const float *ages);
int ubgl_tracers_download(ubgl_tracers_t *t, float *points, unsigned *start, unsigned *end,
                          float *ages);
/* VelocityTextures::uploadFlag (velocity_textures.cpp:95-101): the w x h flag texture the
 * tracers test against (terrain.flagFullRes); NULL: use the simulation's own flag */
int ubgl_tracers_set_flag_texture(ubgl_tracers_t *t, const float *flag, int w, int h);
/* advect_tracer_points.cs:42-82 (RK2 midpoint through the co-located velocity
 * texture of interp_shader.cs, freeze + age in terrain / out of bounds, wang-hash / LCG
 * respawn) on the simulation's resident vx_current / vy_current; pdim = (pwidth,
 * pwidth*H/W); rand_seed is the per-frame rand() of draw_tracers_cs.cpp:140.  Runs on the
 * simulation's stream, asynchronously. */
int ubgl_tracers_advect(ubgl_tracers_t *t, ubgl_sim_t *sim, float dt, unsigned rand_seed);
int ubgl_tracers_shift(ubgl_tracers_t *t, float shift); /* shift_tracers.cs (points only) */
/* interp_shader.cs:15-35 materialised: vxy is (2H-1) x (2W-1) x 2 floats (RG32F), mag
 * (2H-1) x (2W-1); either host pointer may be NULL (device copy only) */
int ubgl_sim_colocate_velocity(ubgl_sim_t *sim, float *vxy_host, float *mag_host);

/* (4) Display without PCIe (CUDA -> OpenGL interop).  Replaces the per-frame texture uploads
 * VelocityTextures::updateFromStaggered (velocity_textures.cpp:63-93: two glTexSubImage2D +
 * interp_shader.cs) and Draw2DBuf::draw(sim.p.data(), ...) (draw.cpp:101 ->
 * draw_2dbuf.cpp:181-209: glTexSubImage2D of p): the texels are written from the resident
 * fields into CUDA arrays by surface stores.  Each argument is a cudaArray_t (NULL: skip) --
 * for a GL texture: cudaGraphicsGLRegisterImage once, then per frame cudaGraphicsMapResources,
 * cudaGraphicsSubResourceGetMappedArray(level 0), this call, cudaGraphicsUnmapResources
 * (INTEGRATION.md).  vxy: (2W-1) x (2H-1) RG32F, mag: (2W-1) x (2H-1) R32F (tex_vxy / tex_mag,
 * velocity_textures.cpp:38-52), p: W x H R32F; wrong sizes or formats are rejected.  Returns
 * after the stores have completed. */
int ubgl_sim_export_display(ubgl_sim_t *sim, void *vxy_array, void *mag_array, void *p_array);
/* Headless stand-in for a mapped GL texture (tests, offscreen consumers): a CUDA array of
 * w x h texels with `channels` (1 or 2) 32-bit floats, surface-writable. */
int ubgl_display_array_create(int w, int h, int channels, int device, void **array_out);
int ubgl_display_array_read(void *array, float *host); /* h x w x channels floats */
int ubgl_display_array_destroy(void *array);

/* (2) Floating items, Simulation::advectFloatingItemsSimple
 * (advect_floating_items.cpp:148-274).  One record = CoItem + CoKinematicsSimple
 * (components.hpp:6-43); array order = the order the reference's view visits. */
typedef struct ubgl_item {
  float size[2], pos[2], rotation;                /* CoItem */
  float mass, vel[2], force[2], angVel, angForce; /* CoKinematicsSimple */
  int bumpCount;
} ubgl_item;
typedef struct ubgl_items ubgl_items_t;
int ubgl_items_create(int device, ubgl_items_t **out);
int ubgl_items_destroy(ubgl_items_t *it);
int ubgl_items_upload(ubgl_items_t *it, const ubgl_item *items, int n);
int ubgl_items_download(ubgl_items_t *it, ubgl_item *items, int cap, int *n);
/* One game frame of every item against the simulation's resident flag / vx / vy / p;
 * reaction forces are added to the DEVICE vx_accum / vy_accum with atomicAdd (the
 * reference scatters serially under accum_mutex, :160), which removes the per-step
 * upload of the accumulator grids.  Asynchronous on the simulation's stream. */
int ubgl_items_advect_simple(ubgl_items_t *it, ubgl_sim_t *sim, float game_dt);
/* Simulation::advectFloatingItems (advect_floating_items.cpp:16-146): the same records
 * advanced as rigid rectangular bodies (the reference's CoItem + CoKinematics entities --
 * submarines, torpedoes): five terrain probes and drag sampled along the four sides per
 * sub-step, reaction forces into the device accumulators.  Bodies do not interact. */
int ubgl_items_advect(ubgl_items_t *it, ubgl_sim_t *sim, float game_dt);

/* (3) Terrain edits on the resident simulation-resolution mask (terrain scale 1).
 * Terrain::drawCircle (terrain.cpp:213-234) for n circles (cx, cy, diam triples, grid
 * coordinates) with one value, followed by MG::updateFields -- what
 * explosion.cpp:60 + ubootgl_app.cpp:111-112 do through the host every frame. */
int ubgl_sim_draw_circles(ubgl_sim_t *sim, const float *xyd, int n, float val);
/* Simulation::setGrids(c, newflag(c)) for every cell (simulation.hpp:82-98 as called by
 * ubootgl_app.cpp:274-278); newflag NULL: re-apply the resident flag.  Like the
 * reference it does not rebuild the coarse flags. */
int ubgl_sim_set_grids_all(ubgl_sim_t *sim, const float *newflag);
/* The field part of UbootGlApp::shiftMap (ubootgl_app.cpp:252-296): scroll vx, vy
 * (front and back), p and the flag one column to the left, enter new_last_column (H
 * values from the host-side terrain generator, terrain.cpp:119-153) on the right, setGrids
 * for every cell, reset the inlet column, saveCurrentVelocityFields, MG::updateFields. */
int ubgl_sim_shift_map(ubgl_sim_t *sim, const float *new_last_column);

/* ---- pressure_solver.cpp free functions, host grids in/out ---------------- */
/* rbgs(p,f,flag,h,alpha) x sweeps, canonical red-black order
 * (pressure_solver.cpp:35-72; the pipelined path :73-87 is not reproduced) */
int ubgl_rbgs(float *p, const float *f, const float *flag, int w, int h,
              float hh, float alpha, int sweeps);
/* calculateResidualField (pressure_solver.cpp:91-116) */
int ubgl_residual(const float *p, const float *f, const float *flag, float *r,
                  int w, int h, float hh, float *l2);
/* restrict (pressure_solver.cpp:118-132); rc is (w/2) x (h/2), border 0 */
int ubgl_restrict(const float *r, int w, int h, float *rc);
/* prolongate (pressure_solver.cpp:134-172); e is w x h, ec/flagc (w/2)x(h/2) */
int ubgl_prolongate(float *e, int w, int h, const float *ec, const float *flagc,
                    const float *flag);
/* correct (pressure_solver.cpp:174-181) */
int ubgl_correct(float *p, const float *e, int w, int h);
/* setZeroGradientBC (pressure_solver.cpp:183-192) */
int ubgl_zero_gradient_bc(float *p, int w, int h);

#ifdef __cplusplus
}
#endif
#endif /* UBGL_H */
