// Generates tests/golden/rng_mt19937_1887_normal.json with the C++ standard library itself:
// the reference fills its test vectors with std::mt19937(1887) + std::normal_distribution<>(0,1)
// (dune/hpdg/test/randomvector.hh:11-21).  Build: g++ -O2 gen_rng_golden.cc -o gen && ./gen > rng_...json
#include <cstdio>
#include <random>
int main() {
  std::mt19937 mt; mt.seed(1887);
  std::normal_distribution<> g{0, 1};
  printf("{\"seed\": 1887, \"generator\": \"libstdc++ std::mt19937 + std::normal_distribution<>\", \"values\": [");
  for (int i = 0; i < 2000; i++) printf("%s%.17g", i ? ", " : "", g(mt));
  printf("]}\n");
}
