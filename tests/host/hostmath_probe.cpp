// Prints what the product's host-side number theory (abc_b200/csrc/hostmath.hpp: the tables abc_ctx_create builds) gives
// for SEAL's default parameter sets, one value per line, for tests/test_tools_host.py to compare with the oracle.
#include <cstdio>
#include <cstdlib>
#include "../../abc_b200/csrc/hostmath.hpp"

int main(int argc, char **argv) {
  const unsigned long long N = argc > 1 ? strtoull(argv[1], nullptr, 10) : 8192;
  auto primes = hm::bfv_default_primes(N);
  printf("k %zu\n", primes.size());
  for (auto q : primes) printf("q %llu\n", (unsigned long long)q);
  printf("t %llu\n", (unsigned long long)hm::get_primes(N, 20, 1)[0]);
  for (auto q : primes) printf("psi %llu\n", (unsigned long long)hm::minimal_2nth_root(q, N));
  auto aux = hm::get_primes(N, 61, primes.size() + 1);   // BEHZ auxiliary base candidates (SEAL: 61-bit primes)
  for (auto b : aux) printf("aux %llu\n", (unsigned long long)b);
  printf("inv %llu\n", (unsigned long long)hm::invmod(primes[0] % primes[1], primes[1]));
  printf("brev %u\n", hm::bit_reverse(0x1234u, 13));
  unsigned long long hi, lo;
  { hm::u64 h, l; hm::barrett_ratio(primes[0], h, l); hi = h; lo = l; }
  printf("barrett %llu %llu\n", hi, lo);
  return 0;
}
