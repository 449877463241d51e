// modarith.cuh — 64-bit modular arithmetic for sm_100a (no tensor cores: modular integer work).
// Conventions follow SEAL 3.6.5 util/uintarithsmallmod.h so results are canonical-identical:
// Barrett with const_ratio = floor(2^128/q), Shoup/Harvey operands (w, floor(w*2^64/q)).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long u64;
typedef unsigned int u32;

// One modulus with everything a kernel needs.  Lives in a device array ctx->d_mods.
struct ModInfo {
  u64 q;
  u64 mu_hi, mu_lo;          // floor(2^128 / q)
  u64 ninv, ninv_s;          // N^-1 mod q and Shoup companion
  u64 wl_ninv, wl_ninv_s;    // irp[1] * N^-1 (last inverse stage folded with the scaling)
  const ulonglong2 *tw;      // forward twiddles {w, w'} : tw[bitrev(i)] = psi^i
  const ulonglong2 *itw;     // inverse twiddles at the same index: itw[j] = tw[j]^-1
  // FP64-assisted class (q < 2^49, see ntt.cuh): companions are the bit patterns of double(w/q)
  const ulonglong2 *twf, *itwf;
  u64 ninv_f, wl_ninv_f;     // bits of double(ninv/q), double(wl_ninv/q)
  u64 qinv_bits;             // bits of double(1/q)
  // pure-FP64 class (q < 2^45): twiddles {bits of double(w), bits of double(w/q)}
  const ulonglong2 *twd, *itwd;
  const ulonglong2 *twp, *itwp;  // ... and {double(w), double(w/q)} pairs for the strided passes (indices < 1024)
  u64 ninv_d, wl_ninv_d;     // bits of double(ninv), double(wl_ninv)
  int ar_class;              // AR_SHOUP / AR_FP / AR_FP_LAZY: the fastest class this modulus allows
};

__device__ __forceinline__ u64 csub(u64 x, u64 q) { return x >= q ? x - q : x; }
__device__ __forceinline__ u64 add_mod(u64 a, u64 b, u64 q) { return csub(a + b, q); }
__device__ __forceinline__ u64 sub_mod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }
__device__ __forceinline__ u64 neg_mod(u64 a, u64 q) { return a ? q - a : 0; }

// Harvey lazy product: x any 64-bit value, result in [0, 2q).
__device__ __forceinline__ u64 mul_shoup_lazy(u64 x, u64 w, u64 ws, u64 q) {
  return x * w - __umul64hi(x, ws) * q;
}
__device__ __forceinline__ u64 mul_shoup(u64 x, u64 w, u64 ws, u64 q) {
  return csub(mul_shoup_lazy(x, w, ws, q), q);
}

// SEAL barrett_reduce_64: x mod q for any 64-bit x.
__device__ __forceinline__ u64 barrett64(u64 x, u64 q, u64 mu_hi) {
  return csub(x - __umul64hi(x, mu_hi) * q, q);
}

// SEAL barrett_reduce_128: (hi:lo) mod q; exact Barrett quotient (error <= 1) when (hi:lo)/q < 2^64.
__device__ __forceinline__ u64 barrett128(u64 lo, u64 hi, u64 q, u64 mu_hi, u64 mu_lo) {
  u64 carry = __umul64hi(lo, mu_lo);
  u64 t2lo = lo * mu_hi, t2hi = __umul64hi(lo, mu_hi);
  u64 t1 = t2lo + carry;
  u64 t3 = t2hi + (t1 < t2lo);
  u64 t4lo = hi * mu_lo, t4hi = __umul64hi(hi, mu_lo);
  u64 t5 = t1 + t4lo;
  u64 c2 = t4hi + (t5 < t1);
  u64 qhat = hi * mu_hi + t3 + c2;
  return csub(lo - qhat * q, q);
}
__device__ __forceinline__ u64 mul_mod(u64 a, u64 b, const ModInfo &m) {
  return barrett128(a * b, __umul64hi(a, b), m.q, m.mu_hi, m.mu_lo);
}
__device__ __forceinline__ u64 mul_mod(u64 a, u64 b, u64 q, u64 mu_hi, u64 mu_lo) {
  return barrett128(a * b, __umul64hi(a, b), q, mu_hi, mu_lo);
}

// 128-bit multiply-accumulate: (hi:lo) += a*b
__device__ __forceinline__ void mac128(u64 &lo, u64 &hi, u64 a, u64 b) {
  u64 pl = a * b, ph = __umul64hi(a, b);
  lo += pl;
  hi += ph + (lo < pl);
}

// ---- sampler (DESIGN.md "Sampler"; the oracle carries the same definition).  Randomness is the ChaCha20 key stream
// (D. J. Bernstein's cipher, 20 rounds, the RFC 8439 state layout) under the context's 256-bit key:
//   state = "expand 32-byte k" | key[8] | block counter | nonce[3]
//   nonce = (domain | b << 4, a lo, a hi): one stream per (domain, a, b) — DOM_SK / DOM_PK / DOM_KSK / DOM_ENC, a = key id
//   or encryption counter, b = component / limb — injective for b < 2^28;
//   the 64-bit word `idx` of a stream is words 2*(idx & 7), 2*(idx & 7) + 1 of block idx >> 3.
// The key comes from the OS generator (getrandom) unless the caller fixes a seed for tests.
struct RngKey { u32 k[8]; };
struct RngStream { RngKey key; u32 n0, n1, n2; };
__host__ __device__ __forceinline__ RngStream rng_stream(const RngKey &key, u64 domain, u64 a, u64 b) {
  RngStream s;
  s.key = key;
  s.n0 = (u32)domain | ((u32)b << 4); s.n1 = (u32)a; s.n2 = (u32)(a >> 32);
  return s;
}
__host__ __device__ __forceinline__ u32 rotl32(u32 x, int r) { return (x << r) | (x >> (32 - r)); }
#define ABC_CHACHA_QR(a, b, c, d) \
  a += b; d ^= a; d = rotl32(d, 16); c += d; b ^= c; b = rotl32(b, 12); \
  a += b; d ^= a; d = rotl32(d, 8);  c += d; b ^= c; b = rotl32(b, 7);
// one 64-byte block as eight 64-bit words (word w = x[2w] | x[2w+1] << 32)
__host__ __device__ __forceinline__ void chacha20_block(const RngStream &s, u32 counter, u64 out[8]) {
  const u32 c0 = 0x61707865u, c1 = 0x3320646eu, c2 = 0x79622d32u, c3 = 0x6b206574u;
  u32 x0 = c0, x1 = c1, x2 = c2, x3 = c3;
  u32 x4 = s.key.k[0], x5 = s.key.k[1], x6 = s.key.k[2], x7 = s.key.k[3];
  u32 x8 = s.key.k[4], x9 = s.key.k[5], x10 = s.key.k[6], x11 = s.key.k[7];
  u32 x12 = counter, x13 = s.n0, x14 = s.n1, x15 = s.n2;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    ABC_CHACHA_QR(x0, x4, x8, x12) ABC_CHACHA_QR(x1, x5, x9, x13) ABC_CHACHA_QR(x2, x6, x10, x14) ABC_CHACHA_QR(x3, x7, x11, x15)
    ABC_CHACHA_QR(x0, x5, x10, x15) ABC_CHACHA_QR(x1, x6, x11, x12) ABC_CHACHA_QR(x2, x7, x8, x13) ABC_CHACHA_QR(x3, x4, x9, x14)
  }
  x0 += c0; x1 += c1; x2 += c2; x3 += c3;
  x4 += s.key.k[0]; x5 += s.key.k[1]; x6 += s.key.k[2]; x7 += s.key.k[3];
  x8 += s.key.k[4]; x9 += s.key.k[5]; x10 += s.key.k[6]; x11 += s.key.k[7];
  x12 += counter; x13 += s.n0; x14 += s.n1; x15 += s.n2;
  out[0] = x0 | (u64)x1 << 32; out[1] = x2 | (u64)x3 << 32; out[2] = x4 | (u64)x5 << 32; out[3] = x6 | (u64)x7 << 32;
  out[4] = x8 | (u64)x9 << 32; out[5] = x10 | (u64)x11 << 32; out[6] = x12 | (u64)x13 << 32; out[7] = x14 | (u64)x15 << 32;
}
// word idx of a stream (computes its whole block: the hot paths fill a buffer with k_rng_fill instead)
__host__ __device__ __forceinline__ u64 rng_word(const RngStream &s, u64 idx) {
  u64 w[8];
  chacha20_block(s, (u32)(idx >> 3), w);
  const int j = (int)(idx & 7);
  u64 r = w[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) r = j == i ? w[i] : r;
  return r;
}
// words 2*e2 and 2*e2 + 1 (the same block)
__host__ __device__ __forceinline__ void rng_word_pair(const RngStream &s, u64 e2, u64 &r0, u64 &r1) {
  u64 w[8];
  chacha20_block(s, (u32)(e2 >> 2), w);
  const int j = (int)(e2 & 3);
  r0 = w[0]; r1 = w[1];
#pragma unroll
  for (int i = 1; i < 4; ++i) { r0 = j == i ? w[2 * i] : r0; r1 = j == i ? w[2 * i + 1] : r1; }
}
// the distributions (SEAL util/rlwe.cpp: sample_poly_ternary, sample_poly_cbd) from one 64-bit word
__host__ __device__ __forceinline__ int ternary_of(u64 r) { return (int)(((r >> 32) * 3) >> 32) - 1; }
__host__ __device__ __forceinline__ int popc64(u64 x) {
#ifdef __CUDA_ARCH__
  return __popcll(x);
#else
  return __builtin_popcountll(x);
#endif
}
__host__ __device__ __forceinline__ int cbd_of(u64 r) { return popc64(r & 0x1fffffULL) - popc64((r >> 21) & 0x1fffffULL); }
