// Test driver for the generated Hydro host class: the parameter block and loop of the reference's
// examples/Hydro/main-kh.cpp:33-60 for a fixed number of steps, printing time and interior sums.
#include <cstdio>
#include <cstdlib>
#include "Hydro.hpp"

int main(int argc, char** argv) {
  const int steps = argc > 1 ? atoi(argv[1]) : 5;
  Hydro sim;
  const int W = sim.om_size_0(), H = sim.om_size_1();
  sim.time() = 0;
  sim.cfl() = 0.5;
  sim.extent0() = 1.0;
  sim.extent1() = 1.0;
  sim.dR0() = sim.extent0() / W;
  sim.dR1() = sim.extent1() / H;
  sim.init();
  for (int t = 0; t < steps; ++t) sim.proceed();
  double sd = 0, sp = 0, sv = 0;
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) { sd += sim.density(x, y); sp += sim.pressure(x, y); sv += sim.velocity0(x, y); }
  printf("%d %d %.17g %.17g %.17g %.17g\n", W, H, sim.time(), sd, sp, sv);
  return 0;
}
