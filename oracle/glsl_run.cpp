// TEST INFRASTRUCTURE ONLY.  Drives the reference's unmodified GLSL compute shaders, compiled
// as C++ through oracle/shim/glsl_shim.hpp (see there and oracle/Makefile), the way the
// reference's host code dispatches them:
//  * glsl_colocate        = VelocityTextures::updateFromStaggered, velocity_textures.cpp:63-93
//                           (interp_shader.cs over (2nx-1) x (2ny-1) invocations);
//  * glsl_tracers_advect  = DrawTracersCS::updateTracers, draw_tracers_cs.cpp:131-156
//                           (advect_tracer_points.cs over ntracers invocations).
// Invocations run one after the other; neither shader reads what another invocation writes.
#include "shim/glsl_shim.hpp"
#undef in

uvec3_t gl_GlobalInvocationID;

namespace sh_interp {
extern uint nx, ny;
extern sampler2D tex_vx_staggered, tex_vy_staggered;
extern image2D img_vxy, img_mag;
void glsl_main();
} // namespace sh_interp

namespace sh_advect {
extern int npoints, ntracers;
extern float dt;
extern vec2_t pdim;
extern uint rand_seed;
extern float angle;
extern sampler2D tex_vxy, tex_flag;
extern vec2_t *points;
extern uint *start_pointers, *end_pointers;
extern float *ages;
extern uint rng_state;
void glsl_main();
} // namespace sh_advect

namespace sh_shift {
extern int npoints, ntracers;
extern float shift;
extern vec2_t *points, *vertices;
void glsl_main();
} // namespace sh_shift

extern "C" {

// DrawTracersCS::shiftTracers, draw_tracers_cs.cpp:272-283 (shift_tracers.cs over ntracers * npoints invocations);
// vertices: 2 vec2 per point (the ribbon geometry, rendering state)
void glsl_tracers_shift(float *points, float *vertices, int ntracers, int npoints, float shift) {
  sh_shift::npoints = npoints;
  sh_shift::ntracers = ntracers;
  sh_shift::shift = shift;
  sh_shift::points = reinterpret_cast<vec2_t *>(points);
  sh_shift::vertices = reinterpret_cast<vec2_t *>(vertices);
  const int groups = (ntracers * npoints - 1) / 256 + 1;
  for (int g = 0; g < groups * 256; g++) {
    gl_GlobalInvocationID.x = (uint)g;
    gl_GlobalInvocationID.y = gl_GlobalInvocationID.z = 0;
    sh_shift::glsl_main();
  }
}

void glsl_colocate(const float *vx, const float *vy, int nx, int ny, float *vxy, float *mag) {
  using namespace sh_interp;
  sh_interp::nx = (uint)nx;
  sh_interp::ny = (uint)ny;
  tex_vx_staggered = sampler2D{vx, nx - 1, ny, 1}; // glTexStorage2D(R32F, nx-1, ny), velocity_textures.cpp:33
  tex_vy_staggered = sampler2D{vy, nx, ny - 1, 1}; // :37
  const int tw = 2 * nx - 1, th = 2 * ny - 1;
  img_vxy = image2D{vxy, tw, th, 2};               // RG32F, :41
  img_mag = image2D{mag, tw, th, 1};               // R32F, :45
  // glDispatchCompute((2nx-2)/32+1, (2ny-2)/8+1, 1) with local size 32 x 8 (:91): whole groups,
  // the invocations beyond the image store nowhere
  const int gxn = ((tw - 1) / 32 + 1) * 32, gyn = ((th - 1) / 8 + 1) * 8;
  for (int y = 0; y < gyn; y++)
    for (int x = 0; x < gxn; x++) {
      gl_GlobalInvocationID.x = (uint)x;
      gl_GlobalInvocationID.y = (uint)y;
      gl_GlobalInvocationID.z = 0;
      if (x < tw && y < th) sh_interp::glsl_main(); // texture() of an out-of-image invocation is harmless but pointless
    }
}

void glsl_tracers_advect(float *points, unsigned *start, unsigned *end, float *ages, int ntracers, int npoints,
                         float dt, float pdx, float pdy, unsigned rand_seed, const float *vxy, int tw, int th,
                         const float *flagtex, int fw, int fh) {
  sh_advect::npoints = npoints;
  sh_advect::ntracers = ntracers;
  sh_advect::dt = dt;
  sh_advect::pdim = vec2_t{pdx, pdy};
  sh_advect::rand_seed = rand_seed;
  sh_advect::angle = 0.0f;
  sh_advect::tex_vxy = sampler2D{vxy, tw, th, 2};
  sh_advect::tex_flag = sampler2D{flagtex, fw, fh, 1};
  sh_advect::points = reinterpret_cast<vec2_t *>(points);
  sh_advect::start_pointers = start;
  sh_advect::end_pointers = end;
  sh_advect::ages = ages;
  const int groups = (ntracers - 1) / 256 + 1; // draw_tracers_cs.cpp:155
  for (int g = 0; g < groups * 256; g++) {
    gl_GlobalInvocationID.x = (uint)g;
    gl_GlobalInvocationID.y = gl_GlobalInvocationID.z = 0;
    sh_advect::rng_state = rand_seed; // `uint rng_state = rand_seed;` is per-invocation state in GLSL
    sh_advect::glsl_main();
  }
}

} // extern "C"
