// ntt.cuh — negacyclic NTT of one RNS limb staged whole in shared memory (N <= 16384), sm_100a.
//
// Replaces SEAL 3.6.5 ntt_negacyclic_harvey / inverse_ntt_negacyclic_harvey (util/ntt.cpp,
// util/dwthandler.h), which ABC reaches inside every multiply / relinearize / rotate / multiply_plain /
// encrypt / decrypt (src/runtime/SealCiphertext.cpp:55,104,105,159; SealCiphertextFactory.cpp:12,150).
// Same transform: forward = Cooley-Tukey, natural in, bit-reversed out, tw[bitrev(i)] = psi^i;
// inverse = Gentleman-Sande, bit-reversed in, natural out, N^-1 folded into the last stage.
// Outputs are canonical residues, so any exact arithmetic inside is bit-compatible with SEAL.
//
// Mapping: every thread owns 8 coefficients per pass and does up to three butterfly stages on them in
// registers (radix-8), so a limb makes 3-4 trips through shared memory instead of log2(N).  For
// N <= 8192 a CTA is 512 threads doing two such groups per pass, so two CTAs share an SM and one CTA's
// barriers and global-memory phases overlap the other's butterflies.  The last forward pass (first inverse
// pass) owns 8 CONTIGUOUS coefficients: gaps 4,2,1 are in-register and gaps 8 (and 16) are warp-shuffle
// butterflies between lane pairs, each lane computing half of the pair's butterflies.  Shared memory is
// XOR-swizzled at 16-byte granularity so the strided and the contiguous passes are bank-conflict free.
//
// Arithmetic classes (AR), chosen per launch from the largest modulus among its rows:
//   AR_SHOUP   64-bit Harvey butterflies, Shoup twiddles (w, floor(w*2^64/q)); any q < 2^62.
//   AR_FP      q < 2^49: the quotient estimate round(y*w/q) comes from the FP64 pipe (one DFMA on the
//              magic-number encoding of y, twiddle companion = double(w/q)); the IMAD pipe only does the two
//              low-half products.  This roughly halves the load on the fmaheavy pipe, which bounds the
//              Shoup version (4 of its 6 IMAD.WIDE are the 64x64 high product).  Harvey guards kept.
//   AR_FP_LAZY q*(2*log2(N)+1) < 2^51: as AR_FP, and the forward transform drops the per-butterfly range
//              guard (values grow by < 2q per stage) and reduces once at the end.
#pragma once
#include "modarith.cuh"

enum { AR_SHOUP = 0, AR_FP = 1, AR_FP_LAZY = 2 };

// element index -> physical index; keeps (even, odd) pairs adjacent so 16-byte accesses stay legal
__device__ __forceinline__ int swz(int e) { return e ^ ((e >> 3) & 14); }

template <int LOGN> struct NttPlan;
// strided passes (radix 2^R each); the contiguous pass (NttLast) takes the remaining stages
template <> struct NttPlan<12> { static constexpr int R0 = 3, R1 = 3, R2 = 2; };   // + 1 shuffle + 3 in-register
template <> struct NttPlan<13> { static constexpr int R0 = 3, R1 = 3, R2 = 3; };   // + 4 in-register (16 per thread)
template <> struct NttPlan<14> { static constexpr int R0 = 3, R1 = 3, R2 = 3; };   // + 1 shuffle + 4 in-register

template <int LOGN> struct NttDims {
  static constexpr int N = 1 << LOGN;
  static constexpr int T = (LOGN >= 14) ? 1024 : ((N / 8 < 512) ? N / 8 : 512);
  static constexpr int MINB = (LOGN >= 14) ? 1 : 2;
  static constexpr int IT = N / 8 / T;
  static constexpr size_t SMEM = (size_t)N * 8;
};

// ---- modular product y*w with a precomputed companion c: lazy result in [0,2q)
// AR_SHOUP: c = floor(w*2^64/q), any 64-bit y.   AR_FP*: c = bits of double(w/q), y < 2^51, result in (0,2q).
// lo64(a*b + c) as one IMAD.WIDE + two IMAD (no separate carry adds): the shape ptxas keeps on the fma pipe
__device__ __forceinline__ u64 mad_lo64(u64 a, u64 b, u64 c) {
  u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32), lo, hi;
  u64 t;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t) : "r"(a0), "r"(b0), "l"(c));
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(t));
  asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(a0), "r"(b1));
  asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(a1), "r"(b0));
  asm("mov.b64 %0, {%1, %2};" : "=l"(t) : "r"(lo), "r"(hi));
  return t;
}
template <int AR> __device__ __forceinline__ u64 mul_tw(u64 y, u64 w, u64 c, u64 q) {
  if (AR == AR_SHOUP) {
    return y * w - __umul64hi(y, c) * q;
  } else {
    const double yd = __longlong_as_double((long long)(y | 0x4330000000000000ULL)) - 4503599627370496.0;
    const double t = fma(yd, __longlong_as_double((long long)c), 4503599627370496.0);
    const u64 qr = (u64)__double_as_longlong(t) & 0x000FFFFFFFFFFFFFULL;  // round(y*w/q), off by <= 1
    return mad_lo64(qr, 0 - q, mad_lo64(y, w, q));                        // y*w - qr*q + q  (mod 2^64)
  }
}
// x mod q for x < 2^51 via the FP64 pipe: result in (0,2q)
__device__ __forceinline__ u64 reduce_fp(u64 x, double qinv, u64 q) {
  const double xd = __longlong_as_double((long long)(x | 0x4330000000000000ULL)) - 4503599627370496.0;
  const double t = fma(xd, qinv, 4503599627370496.0);
  const u64 qr = (u64)__double_as_longlong(t) & 0x000FFFFFFFFFFFFFULL;
  return x - qr * q + q;
}

// forward (Cooley-Tukey) butterfly.  Guarded: x,y in [0,4q) -> [0,4q).  AR_FP_LAZY: no guard, +2q per stage.
template <int AR> __device__ __forceinline__ void bf_fwd(u64 &x, u64 &y, ulonglong2 w, u64 q, u64 q2) {
  const u64 u = (AR == AR_FP_LAZY) ? x : csub(x, q2);
  const u64 v = mul_tw<AR>(y, w.x, w.y, q);
  x = u + v;
  y = u + q2 - v;
}
// inverse (Gentleman-Sande) butterfly: x,y in [0,2q) -> [0,2q)
template <int AR> __device__ __forceinline__ void bf_inv(u64 &x, u64 &y, ulonglong2 w, u64 q, u64 q2) {
  const u64 u = x, v = y;
  x = csub(u + v, q2);
  y = mul_tw<AR>(u + q2 - v, w.x, w.y, q);
}

// ---- strided pass: stages S0 .. S0+R-1 (stage s has 2^s groups, gap N >> (s+1)); twbase = 1 for a whole
// transform, (2^a + block) when this limb is block `block` of the tail of a larger 2^(a+LOGN) transform.
template <int LOGN, int S0, int R, int AR>
__device__ __forceinline__ void ntt_fwd_mid(u64 *sm, const ulonglong2 *__restrict__ tw, u32 twbase, u64 q, u64 q2,
                                            int tid) {
  typedef NttDims<LOGN> D;
  constexpr int LG = LOGN - S0 - R;
#pragma unroll
  for (int it = 0; it < D::IT; ++it) {
    const int vt = tid + it * D::T;
    const int off = vt & ((1 << LG) - 1), blk = vt >> LG;
    const int base = (blk << (LG + 3)) | off;
    const int pbase = swz(base);
    u64 x[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) x[r] = sm[LG >= 7 ? pbase + (r << LG) : swz(base + (r << LG))];
#pragma unroll
    for (int b = R - 1; b >= 0; --b) {
      const int s = S0 + R - 1 - b;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (r & (1 << b)) continue;
        const ulonglong2 w = __ldg(&tw[(twbase << s) + ((u32)((blk << 3) + r) >> (b + 1))]);
        bf_fwd<AR>(x[r], x[r | (1 << b)], w, q, q2);
      }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) sm[LG >= 7 ? pbase + (r << LG) : swz(base + (r << LG))] = x[r];
  }
}

template <int LOGN, int S0, int R, bool FOLD, int AR>
__device__ __forceinline__ void ntt_inv_mid(u64 *sm, const ModInfo &M, u32 twbase, u64 q, u64 q2, int tid) {
  typedef NttDims<LOGN> D;
  constexpr int LG = LOGN - S0 - R;
  const ulonglong2 *__restrict__ tw = (AR == AR_SHOUP) ? M.itw : M.itwf;
#pragma unroll
  for (int it = 0; it < D::IT; ++it) {
    const int vt = tid + it * D::T;
    const int off = vt & ((1 << LG) - 1), blk = vt >> LG;
    const int base = (blk << (LG + 3)) | off;
    const int pbase = swz(base);
    u64 x[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) x[r] = sm[LG >= 7 ? pbase + (r << LG) : swz(base + (r << LG))];
#pragma unroll
    for (int b = 0; b < R; ++b) {
      const int s = S0 + R - 1 - b;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (r & (1 << b)) continue;
        if (FOLD && s == 0) {
          // last stage of the whole transform: fold N^-1 into both outputs (u+v < 4q < 2^51 in the FP classes)
          const u64 u = x[r], v = x[r | (1 << b)];
          x[r] = mul_tw<AR>(u + v, M.ninv, (AR == AR_SHOUP) ? M.ninv_s : M.ninv_f, q);
          x[r | (1 << b)] = mul_tw<AR>(u + q2 - v, M.wl_ninv, (AR == AR_SHOUP) ? M.wl_ninv_s : M.wl_ninv_f, q);
        } else {
          const ulonglong2 w = __ldg(&tw[(twbase << s) + ((u32)((blk << 3) + r) >> (b + 1))]);
          bf_inv<AR>(x[r], x[r | (1 << b)], w, q, q2);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) sm[LG >= 7 ? pbase + (r << LG) : swz(base + (r << LG))] = x[r];
  }
}

// ---- contiguous pass: every thread owns E = 8*IT CONSECUTIVE coefficients (8 or 16): gaps E/2..1 are in-register,
// the NSH stages above them (gaps E, 2E) are warp-shuffle butterflies between lane pairs, each lane computing half
// of the pair's butterflies.  (N = 8192: E = 16 and no shuffle stage at all.)
template <int LOGN> struct NttLast {
  // E = 16 (no shuffle stage at N = 8192) was measured slower than two groups of 8 with one shuffle stage
  // (13.7 vs 15.0 Mrows/s forward): the 32 live data registers cost more than the shuffles save.
  static constexpr int E = 8, LOGE = (E == 16) ? 4 : 3;
  static constexpr int GROUPS = 8 * NttDims<LOGN>::IT / E;
  static constexpr int NSH = LOGN - (NttPlan<LOGN>::R0 + NttPlan<LOGN>::R1 + NttPlan<LOGN>::R2) - LOGE;
  static_assert(E == 8 || E == 16, "contiguous pass handles 8 or 16 coefficients per thread");
  static_assert(NSH >= 0 && NSH <= 2, "stage plan does not add up");
};

template <int LOGN, int AR>
__device__ __forceinline__ void ntt_fwd_last(u64 *sm, const ModInfo &M, u32 twbase, u64 q, u64 q2, int tid) {
  typedef NttLast<LOGN> P;
  constexpr int E = P::E, H = E / 2;
  const ulonglong2 *__restrict__ tw = (AR == AR_SHOUP) ? M.tw : M.twf;
#pragma unroll
  for (int g = 0; g < P::GROUPS; ++g) {
  const int vt = tid + g * NttDims<LOGN>::T;
  u64 x[E];
#pragma unroll
  for (int i = 0; i < H; ++i) {
    ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(&sm[swz(E * vt + 2 * i)]);
    x[2 * i] = v.x; x[2 * i + 1] = v.y;
  }
#pragma unroll
  for (int j = P::NSH - 1; j >= 0; --j) {
    const int s = LOGN - 1 - P::LOGE - j;
    const ulonglong2 w = __ldg(&tw[(twbase << s) + ((u32)vt >> (1 + j))]);
    const bool hi = (vt >> j) & 1;
#pragma unroll
    for (int r = 0; r < H; ++r) {
      u64 recv = __shfl_xor_sync(0xffffffffu, hi ? x[r] : x[H + r], 1 << j);
      u64 a = hi ? recv : x[r];
      u64 b = hi ? x[H + r] : recv;
      bf_fwd<AR>(a, b, w, q, q2);
      u64 got = __shfl_xor_sync(0xffffffffu, hi ? a : b, 1 << j);
      x[r] = hi ? got : a;
      x[H + r] = hi ? b : got;
    }
  }
#pragma unroll
  for (int b = P::LOGE - 1; b >= 0; --b) {
    const int s = LOGN - 1 - b;
#pragma unroll
    for (int r = 0; r < E; ++r) {
      if (r & (1 << b)) continue;
      const ulonglong2 w = __ldg(&tw[(twbase << s) + ((u32)(E * vt + r) >> (b + 1))]);
      bf_fwd<AR>(x[r], x[r | (1 << b)], w, q, q2);
    }
  }
#pragma unroll
  for (int i = 0; i < H; ++i) {
    ulonglong2 v;
    if (AR == AR_FP_LAZY) {
      const double qinv = __longlong_as_double((long long)M.qinv_bits);
      v.x = csub(reduce_fp(x[2 * i], qinv, q), q);
      v.y = csub(reduce_fp(x[2 * i + 1], qinv, q), q);
    } else {
      v.x = csub(csub(x[2 * i], q2), q);
      v.y = csub(csub(x[2 * i + 1], q2), q);
    }
    *reinterpret_cast<ulonglong2 *>(&sm[swz(E * vt + 2 * i)]) = v;
  }
  }
}

template <int LOGN, int AR>
__device__ __forceinline__ void ntt_inv_first(u64 *sm, const ModInfo &M, u32 twbase, u64 q, u64 q2, int tid) {
  typedef NttLast<LOGN> P;
  constexpr int E = P::E, H = E / 2;
  const ulonglong2 *__restrict__ tw = (AR == AR_SHOUP) ? M.itw : M.itwf;
#pragma unroll
  for (int g = 0; g < P::GROUPS; ++g) {
  const int vt = tid + g * NttDims<LOGN>::T;
  u64 x[E];
#pragma unroll
  for (int i = 0; i < H; ++i) {
    ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(&sm[swz(E * vt + 2 * i)]);
    x[2 * i] = v.x; x[2 * i + 1] = v.y;
  }
#pragma unroll
  for (int b = 0; b < P::LOGE; ++b) {
    const int s = LOGN - 1 - b;
#pragma unroll
    for (int r = 0; r < E; ++r) {
      if (r & (1 << b)) continue;
      const ulonglong2 w = __ldg(&tw[(twbase << s) + ((u32)(E * vt + r) >> (b + 1))]);
      bf_inv<AR>(x[r], x[r | (1 << b)], w, q, q2);
    }
  }
#pragma unroll
  for (int j = 0; j < P::NSH; ++j) {
    const int s = LOGN - 1 - P::LOGE - j;
    const ulonglong2 w = __ldg(&tw[(twbase << s) + ((u32)vt >> (1 + j))]);
    const bool hi = (vt >> j) & 1;
#pragma unroll
    for (int r = 0; r < H; ++r) {
      u64 recv = __shfl_xor_sync(0xffffffffu, hi ? x[r] : x[H + r], 1 << j);
      u64 a = hi ? recv : x[r];
      u64 b = hi ? x[H + r] : recv;
      bf_inv<AR>(a, b, w, q, q2);
      u64 got = __shfl_xor_sync(0xffffffffu, hi ? a : b, 1 << j);
      x[r] = hi ? got : a;
      x[H + r] = hi ? b : got;
    }
  }
#pragma unroll
  for (int i = 0; i < H; ++i) {
    ulonglong2 v; v.x = x[2 * i]; v.y = x[2 * i + 1];
    *reinterpret_cast<ulonglong2 *>(&sm[swz(E * vt + 2 * i)]) = v;
  }
  }
}

// ---- whole-limb transforms on a swizzled shared-memory limb.  Caller has filled sm[swz(e)] and synced.
// Forward: input < 4q (canonical for AR_FP_LAZY), output canonical.  Returns after a barrier.
template <int LOGN, int AR>
__device__ __forceinline__ void ntt_fwd_smem(u64 *sm, const ModInfo &M, u32 twbase, int tid) {
  typedef NttPlan<LOGN> P;
  const u64 q = M.q, q2 = 2 * q;
  const ulonglong2 *tw = (AR == AR_SHOUP) ? M.tw : M.twf;
  ntt_fwd_mid<LOGN, 0, P::R0, AR>(sm, tw, twbase, q, q2, tid);
  __syncthreads();
  ntt_fwd_mid<LOGN, P::R0, P::R1, AR>(sm, tw, twbase, q, q2, tid);
  __syncthreads();
  if constexpr (P::R2 > 0) {
    ntt_fwd_mid<LOGN, P::R0 + P::R1, P::R2, AR>(sm, tw, twbase, q, q2, tid);
    __syncthreads();
  }
  ntt_fwd_last<LOGN, AR>(sm, M, twbase, q, q2, tid);
  __syncthreads();
}
// Inverse: input < 2q, output in [0,2q) (the caller's copy-out does the final conditional subtract).
// WHOLE folds N^-1 into the last stage; for a tail block (WHOLE=false) a head pass finishes the transform.
template <int LOGN, bool WHOLE, int AR>
__device__ __forceinline__ void ntt_inv_smem(u64 *sm, const ModInfo &M, u32 twbase, int tid) {
  typedef NttPlan<LOGN> P;
  constexpr int A = (AR == AR_FP_LAZY) ? AR_FP : AR;  // the inverse keeps its guards
  const u64 q = M.q, q2 = 2 * q;
  ntt_inv_first<LOGN, A>(sm, M, twbase, q, q2, tid);
  __syncthreads();
  if constexpr (P::R2 > 0) {
    ntt_inv_mid<LOGN, P::R0 + P::R1, P::R2, false, A>(sm, M, twbase, q, q2, tid);
    __syncthreads();
  }
  ntt_inv_mid<LOGN, P::R0, P::R1, false, A>(sm, M, twbase, q, q2, tid);
  __syncthreads();
  ntt_inv_mid<LOGN, 0, P::R0, WHOLE, A>(sm, M, twbase, q, q2, tid);
  __syncthreads();
}
