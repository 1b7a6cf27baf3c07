// mgtest -- the reference's stand-alone multigrid known-answer test
// (te42kyfo/ubootgl mgtest/mgtest.cpp:9-71) against the drop-in MG: Laplace
// problem on N x N with the analytic solution sinh(pi y) sin(pi x), 5 V-cycles
// with the residual printed after each, the scaled L2 error, then the mean time
// of 10 more V-cycles.  (The reference file assigns floats to grids, :13-16,
// which no longer compiles; fill() is used instead.)
//   usage: mgtest [N=1025] [--resident]
#include "dtime.hpp"
#include "pressure_solver.hpp"
#include "../../include/ubgl.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

int main(int argc, char **argv) {
  int N = 1025;
  bool resident = false;
  for (int i = 1; i < argc; i++) {
    if (!std::strcmp(argv[i], "--resident")) resident = true;
    else N = std::atoi(argv[i]);
  }
  const float h = 1.0 / (N - 1);
  Single2DGrid u(N, N), rhs(N, N), flag(N, N), r(N, N), reference(N, N);
  flag.fill(1.0f);
  for (int x = 1; x < N - 1; x++) u(x, N - 1) = std::sinh(M_PI) * std::sin(x / (N - 1.0) * M_PI);
  for (int y = 0; y < N; y++)
    for (int x = 1; x < N - 1; x++) reference(x, y) = std::sinh(y * h * M_PI) * std::sin(x * h * M_PI);

  std::printf("Initial residual: %g\n", calculateResidualField(u, rhs, flag, r, h));
  MG mg(N, N);
  for (int i = 0; i < 5; i++) {
    mg.solve(u, rhs, flag, h);
    std::printf("%g\n", calculateResidualField(u, rhs, flag, r, h));
  }
  double err = 0.0;
  {
    const Single2DGrid &cu = u, &cr = reference;
    for (int y = 0; y < N; y++)
      for (int x = 0; x < N; x++) {
        const float e = cr(x, y) - cu(x, y);
        err += (double)e * e;
      }
  }
  std::printf("%g\n", std::sqrt(err) / N / N);

  const int iterations = 10;
  double t1 = dtime();
  for (int i = 0; i < iterations; i++) mg.solve(u, rhs, flag, h);
  double t2 = dtime();
  std::printf("%gms per MG::solve through host grids (upload p,f,flag + V-cycle + download p)\n",
              (t2 - t1) / iterations * 1000);

  if (resident) { // the same V-cycle with the grids resident on the device
    ubgl_mg_t *m = nullptr;
    if (ubgl_mg_create(N, N, 0, &m) != UBGL_OK) return 1;
    ubgl_mg_upload(m, u.data(), rhs.data(), flag.data());
    ubgl_mg_solve(m, h, 0, 2);
    ubgl_mg_sync(m);
    t1 = dtime();
    ubgl_mg_solve(m, h, 0, iterations);
    ubgl_mg_sync(m);
    t2 = dtime();
    std::printf("%gms per V-cycle, grids resident\n", (t2 - t1) / iterations * 1000);
    ubgl_mg_destroy(m);
  }
  return 0;
}
