// Test driver for the generated host class (same accessor usage as the reference's examples/Life/main.cpp:16-51,
// without the terminal animation): seed cells through `cell(x,y) = 1`, step, print generation/population and a
// checksum of the cells read back through `cell(x,y)`.
#include <cstdio>
#include <cstdlib>
#include "Life.hpp"

int main(int argc, char** argv) {
  const int steps = argc > 1 ? atoi(argv[1]) : 10;
  Life sim;
  const int W = sim.om_size_0(), H = sim.om_size_1();
  sim.init();
  unsigned long long s = 20261017ull;
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      s = s * 6364136223846793005ull + 1442695040888963407ull;
      if ((s >> 33) % 100 < 35) sim.cell(x, y) = 1;
    }
  for (int t = 0; t < steps; ++t) {
    sim.proceed();
    if (t % 4 == 3) sim.cell(t % W, t % H) = 1;        // a host write between kernels (lazy mirror round trip)
  }
  unsigned long long sum = 0, hash = 1469598103934665603ull;
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const int c = sim.cell(x, y);
      sum += c;
      hash = (hash ^ (unsigned long long)c) * 1099511628211ull;
    }
  printf("%d %d %d %llu %llu\n", W, H, sim.generation(), sum, hash);
  printf("population %d\n", sim.population());
  if (getenv("OM_PRINT_EARLY")) fprintf(stderr, "early %ld\n", sim.om_early_exchanges());
  return 0;
}
