#include <thrust/execution_policy.h>
#include <thrust/random.h>
#include <thrust/shuffle.h>
#include <cstdio>
#include <vector>
int main() {
  const unsigned seeds[] = {1u, 12345u, 2147483647u, 0u, 4294967295u, 77u};
  const int sizes[] = {1, 2, 5, 16, 17, 40, 256, 257, 1000};
  std::printf("{\"cases\": [");
  bool first = true;
  for (unsigned seed : seeds)
    for (int n : sizes) {
      std::vector<unsigned> v(n);
      for (int i = 0; i < n; ++i) v[i] = (unsigned)i;
      thrust::default_random_engine rng;
      rng.seed(seed);
      thrust::shuffle(thrust::host, v.begin(), v.end(), rng);
      std::printf("%s{\"seed\": %u, \"n\": %d, \"perm\": [", first ? "" : ", ", seed, n);
      for (int i = 0; i < n; ++i) std::printf("%s%u", i ? ", " : "", v[i]);
      std::printf("]}");
      first = false;
    }
  std::printf("]}\n");
  return 0;
}
