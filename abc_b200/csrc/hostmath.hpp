// hostmath.hpp — host-side number theory for context setup (the part of seal::SEALContext /
// util::RNSTool / util::NTTTables construction that ABC triggers at
// /root/reference/src/runtime/SealCiphertextFactory.cpp:74-86).  Independent of oracle/.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace hm {
typedef unsigned long long u64;
typedef unsigned __int128 u128;

inline u64 mulmod(u64 a, u64 b, u64 m) { return (u64)((u128)a * b % m); }
inline u64 powmod(u64 a, u64 e, u64 m) {
  u64 r = 1 % m;
  a %= m;
  for (; e; e >>= 1, a = mulmod(a, a, m))
    if (e & 1) r = mulmod(r, a, m);
  return r;
}
// inverse modulo any m (prime or power of two), by the extended Euclidean algorithm
inline u64 invmod(u64 a, u64 m) {
  __int128 r0 = m, r1 = a % m, s0 = 0, s1 = 1;
  while (r1 != 0) {
    __int128 qt = r0 / r1, t = r0 - qt * r1;
    r0 = r1; r1 = t;
    t = s0 - qt * s1; s0 = s1; s1 = t;
  }
  if (r0 != 1) throw std::runtime_error("invmod: not invertible");
  return (u64)(s0 < 0 ? s0 + (__int128)m : s0);
}
inline u64 shoup(u64 w, u64 q) { return (u64)(((u128)w << 64) / q); }

inline bool is_prime(u64 n) {
  if (n < 2) return false;
  for (u64 p : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull}) {
    if (n == p) return true;
    if (n % p == 0) return false;
  }
  u64 d = n - 1; int s = 0;
  while ((d & 1) == 0) { d >>= 1; ++s; }
  for (u64 a : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull}) {
    u64 x = powmod(a, d, n);
    if (x == 1 || x == n - 1) continue;
    bool witness = true;
    for (int r = 1; r < s && witness; ++r) { x = mulmod(x, x, n); if (x == n - 1) witness = false; }
    if (witness) return false;
  }
  return true;
}

// SEAL rule (util/numth.cpp get_primes): scan down from 2^bits - 2N + 1 in steps of 2N
inline std::vector<u64> get_primes(u64 N, int bits, size_t count) {
  std::vector<u64> out;
  u64 step = 2 * N, v = (1ull << bits) - step + 1, lo = 1ull << (bits - 1);
  for (; out.size() < count && v > lo; v -= step)
    if (is_prime(v)) out.push_back(v);
  if (out.size() != count) throw std::runtime_error("get_primes: not enough primes");
  return out;
}

// SEAL CoeffModulus::BFVDefault(N) (util/globals.cpp, 128-bit security table)
inline std::vector<u64> bfv_default_primes(u64 N) {
  switch (N) {
    case 4096: return {0xffffee001, 0xffffc4001, 0x1ffffe0001};
    case 8192: return {0x7fffffd8001, 0x7fffffc8001, 0xfffffffc001, 0xffffff6c001, 0xfffffebc001};
    case 16384: return {0xfffffffd8001, 0xfffffffa0001, 0xfffffff00001, 0x1fffffff68001, 0x1fffffff50001,
                        0x1ffffffee8001, 0x1ffffffea0001, 0x1ffffffe88001, 0x1ffffffe48001};
    case 32768: return {0x7fffffffe90001, 0x7fffffffbf0001, 0x7fffffffbd0001, 0x7fffffffba0001, 0x7fffffffaa0001,
                        0x7fffffffa50001, 0x7fffffff9f0001, 0x7fffffff7e0001, 0x7fffffff770001, 0x7fffffff380001,
                        0x7fffffff330001, 0x7fffffff2d0001, 0x7fffffff170001, 0x7fffffff150001, 0x7ffffffef00001,
                        0xfffffffff70001};
    default: throw std::runtime_error("BFVDefault: no default coefficient modulus for this poly_degree");
  }
}

inline uint32_t bit_reverse(uint32_t x, int bits) {
  uint32_t r = 0;
  for (int i = 0; i < bits; ++i, x >>= 1) r = (r << 1) | (x & 1);
  return r;
}

// smallest primitive 2N-th root of unity mod q (SEAL try_minimal_primitive_root)
inline u64 minimal_2nth_root(u64 q, u64 N) {
  u64 g = 0;
  for (u64 c = 2;; ++c) {
    g = powmod(c, (q - 1) / (2 * N), q);
    if (powmod(g, N, q) == q - 1) break;
  }
  u64 sq = mulmod(g, g, q), best = g;
  for (u64 i = 1; i < N; ++i) { g = mulmod(g, sq, q); if (g < best) best = g; }
  return best;
}

// prod of v (optionally skipping one index) reduced mod m
inline u64 prod_mod(const std::vector<u64> &v, u64 m, long skip = -1) {
  u64 r = 1 % m;
  for (size_t i = 0; i < v.size(); ++i)
    if ((long)i != skip) r = mulmod(r, v[i] % m, m);
  return r;
}
inline int prod_bits(const std::vector<u64> &v) {
  std::vector<u64> w(1, 1);
  for (u64 p : v) {
    u64 carry = 0;
    for (auto &x : w) { u128 t = (u128)x * p + carry; x = (u64)t; carry = (u64)(t >> 64); }
    if (carry) w.push_back(carry);
  }
  int b = 0;
  for (u64 top = w.back(); top; top >>= 1) ++b;
  return 64 * (int)(w.size() - 1) + b;
}
inline int bits_of(u64 v) { int b = 0; for (; v; v >>= 1) ++b; return b; }
// floor(2^128 / q) as (hi, lo)
inline void barrett_ratio(u64 q, u64 &hi, u64 &lo) {
  // 2^128 / q = ((2^128 - 1) / q) unless q is a power of two (never: q is an odd prime or handled by caller)
  u128 all = ~(u128)0;
  u128 r = all / q;
  if (all % q == q - 1) r += 1;
  hi = (u64)(r >> 64); lo = (u64)r;
}
}  // namespace hm
