// simdemo -- drives the drop-in Simulation the way the game does
// (sim_loop.cpp:26-50, ubootgl_app.cpp:109-115,245-298, explosion.cpp:33,60):
// step(), item forces scattered into the accumulators under accum_mutex,
// explosions pushing pressure sinks and carving terrain through setGrids +
// mg.updateFields, and reads of the public members.  Writes the resulting
// fields as raw fp32 so tests/test_host_mirror.py can compare them with the
// oracle run of the same script.
//   usage: simdemo W H steps out_prefix
#include "simulation.hpp"
#include <cstdio>
#include <cstdlib>
#include <string>

static void dump(const std::string &path, const float *d, size_t n) {
  FILE *fp = std::fopen(path.c_str(), "wb");
  if (!fp) std::exit(2);
  std::fwrite(d, sizeof(float), n, fp);
  std::fclose(fp);
}

int main(int argc, char **argv) {
  if (argc < 5) {
    std::fprintf(stderr, "usage: simdemo W H steps out_prefix\n");
    return 1;
  }
  const int W = std::atoi(argv[1]), H = std::atoi(argv[2]), steps = std::atoi(argv[3]);
  const std::string out = argv[4];
  // channel with two rectangular obstacles (same rule as tests/test_host_mirror.py)
  Single2DGrid flag(W, H);
  flag.fill(1.0f);
  for (int x = 0; x < W; x++) flag(x, 0) = flag(x, H - 1) = 0.0f;
  for (int y = H / 4; y < H / 2; y++)
    for (int x = W / 5; x < W / 5 + W / 10; x++) flag(x, y) = 0.0f;
  for (int y = H / 2; y < 3 * H / 4; y++)
    for (int x = W / 2; x < W / 2 + W / 12; x++) flag(x, y) = 0.0f;

  Simulation sim(flag, 0.8f, 0.001f);
  const float dt = 0.001f;
  for (int s = 0; s < steps; s++) {
    {
      // a floating item pushes on the fluid (advect_floating_items.cpp:118-120)
      std::lock_guard<std::mutex> lock(sim.accum_mutex);
      sim.vx_accum(W / 3, H / 3) += 0.02f;
      sim.vy_accum(W / 3, H / 3) -= 0.01f;
    }
    if (s == 1) sim.sinks.push_back(glm::vec3(0.5f * 0.8f, 0.5f * 0.8f * H / W, 120.0f)); // explosion.cpp:33
    if (s == 2) { // crater: terrain turns to fluid / rubble to solid (explosion.cpp:60, ubootgl_app.cpp:276)
      for (int y = H / 4; y < H / 4 + 4; y++)
        for (int x = W / 5; x < W / 5 + 4; x++) sim.setGrids(glm::ivec2(x, y), 1.0f);
      for (int y = 3 * H / 5; y < 3 * H / 5 + 3; y++)
        for (int x = 3 * W / 4; x < 3 * W / 4 + 3; x++) sim.setGrids(glm::ivec2(x, y), 0.0f);
      sim.mg.updateFields(sim.flag); // ubootgl_app.cpp:112
    }
    sim.step(dt);
  }
  const DoubleBuffered2DGrid &vx = sim.vx, &vy = sim.vy;
  const Single2DGrid &p = sim.p, &cx = sim.vx_current, &fl = sim.flag;
  dump(out + ".vx", vx.data(), (size_t)(W - 1) * H);
  dump(out + ".vy", vy.data(), (size_t)W * (H - 1));
  dump(out + ".p", p.data(), (size_t)W * H);
  dump(out + ".vxc", cx.data(), (size_t)(W - 1) * H);
  dump(out + ".flag", fl.data(), (size_t)W * H);
  std::printf("simdemo %dx%d %d steps: %lld kernel launches, %zu sinks alive, wall-sample flag %.2f\n", W, H,
              steps, sim.kernelLaunches(), sim.sinks.size(),
              sim.psampleFlagLinear(glm::vec2(0.4f, 0.2f)));
  return 0;
}
