// Test driver for the generated rank-3 host class of Diff3 (paraiso_b200/examples/rank3.py: loadIndex of all axes,
// loadSize, an asymmetric axis-2 reach, a Max reduce feeding a second stage): init, a host write through the accessor
// (margin coordinates included on Open axes), steps, then every cell of the memory box read back through `u(x, y, z)`.
#include <cstdio>
#include <cstdlib>
#include "Diff3.hpp"

int main(int argc, char** argv) {
  const int steps = argc > 1 ? atoi(argv[1]) : 2;
  Diff3 sim;
  const int W = sim.om_size_0(), H = sim.om_size_1(), D = sim.om_size_2();
  sim.init();
  sim.u(1, 2, 3) = 0.75;
  sim.u(W - 1, H - 1, D - 1) = -0.5;
  for (int t = 0; t < steps; ++t) {
    sim.proceed();
    if (t == 0) sim.u(0, 0, t % D) += 0.125;        // a host write between kernels (lazy mirror round trip)
  }
  printf("%d %d %d %.17g\n", W, H, D, sim.peak());
  for (int z = -sim.om_lower_margin_2(); z < D + sim.om_upper_margin_2(); ++z)
    for (int y = -sim.om_lower_margin_1(); y < H + sim.om_upper_margin_1(); ++y)
      for (int x = -sim.om_lower_margin_0(); x < W + sim.om_upper_margin_0(); ++x)
        printf("%.17g\n", sim.u(x, y, z));
  return 0;
}
