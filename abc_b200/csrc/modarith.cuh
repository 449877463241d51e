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

// sampler shared with the oracle spec (DESIGN.md "sampler"): splitmix64 finaliser chain
__host__ __device__ __forceinline__ u64 mix64(u64 z) {
  z += 0x9e3779b97f4a7c15ULL;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ u64 stream_key(u64 seed, u64 domain, u64 a, u64 b) {
  u64 h = mix64(seed ^ (domain * 0xd6e8feb86659fd93ULL));
  h = mix64(h ^ a);
  return mix64(h ^ b);
}
