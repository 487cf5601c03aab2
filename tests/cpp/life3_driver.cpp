// Test driver for a generated rank-3 host class (Life3, paraiso_b200/examples/rank3.py): seed through
// `cell(x, y, z) = 1`, step, print generation / population and a checksum read back through `cell(x, y, z)`.
#include <cstdio>
#include <cstdlib>
#include "Life3.hpp"

int main(int argc, char** argv) {
  const int steps = argc > 1 ? atoi(argv[1]) : 5;
  Life3 sim;
  const int W = sim.om_size_0(), H = sim.om_size_1(), D = sim.om_size_2();
  unsigned long long s = 20261017ull;
  for (int z = 0; z < D; ++z)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        if ((s >> 33) % 100 < 30) sim.cell(x, y, z) = 1;
      }
  for (int t = 0; t < steps; ++t) {
    sim.proceed();
    if (t == 2) sim.cell(t % W, t % H, t % D) = 1;        // a host write between kernels (lazy mirror round trip)
  }
  unsigned long long sum = 0, hash = 1469598103934665603ull;
  for (int z = 0; z < D; ++z)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        const int c = sim.cell(x, y, z);
        sum += c;
        hash = (hash ^ (unsigned long long)c) * 1099511628211ull;
      }
  printf("%d %d %d %d %llu %llu\n", W, H, D, sim.generation(), sum, hash);
  printf("population %d\n", sim.population());
  return 0;
}
