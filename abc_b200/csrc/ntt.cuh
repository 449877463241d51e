// ntt.cuh — negacyclic NTT of one RNS limb (or one 8192-coefficient block of a larger limb) in shared memory, sm_100a.
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
// barriers and global-memory phases overlap the other's butterflies.  The contiguous pass owns 8 CONSECUTIVE
// coefficients: gaps 4,2,1 are in-register and the stage above them is a warp-shuffle butterfly between lane
// pairs, each lane computing half of the pair's butterflies.  (Exact-double class at N = 8192: strided passes of 3, 3 and 4
// stages and a shuffle-free contiguous pass, see "radix-16 plan" below.)  Shared memory is XOR-swizzled at 16-byte
// granularity so the strided and the contiguous passes are bank-conflict free; the swizzle is linear over XOR, so a
// strided access costs one XOR with a compile-time constant (swz_strided, swz_row8).
//
// Arithmetic classes (AR), chosen per launch from the largest modulus among its rows:
//   AR_SHOUP   64-bit Harvey butterflies, Shoup twiddles (w, floor(w*2^64/q)); any q < 2^62.  29 SASS
//              instructions per butterfly, bound by the fmaheavy pipe (6 half-rate IMAD.WIDE each).
//   AR_FP      q < 2^49: the quotient estimate round(y*w/q) comes from the FP64 pipe (DADD + DFMA on the
//              magic-number encoding of y, twiddle companion = double(w/q)); the IMAD pipe only does the two
//              low-half products as carry-free IMAD chains.  Harvey range guards kept.
//   AR_FP_LAZY q*(log2(N)+2) < 2^51: as AR_FP, on SIGNED 64-bit values with centred products
//              |y*w - round(y*w/q)*q| <= 0.75q.  The forward butterfly is x+v, x-v with no range guard and no
//              +2q offset (values grow by < 0.75q per stage); the inverse reduces its sum chain twice per
//              transform instead of guarding every butterfly.
//   AR_F64     q < 0.97 * 2^45: the whole transform runs on the FP64 pipe, on exact integer-valued doubles.  A modular product
//              is 6 FP64 instructions: the error-free product y*w = ph + pl (DMUL + DFMA), Q = rint(ph * (1/q)) (DFMA +
//              DADD with the 1.5*2^52 magic: the estimate comes from the rounded product, a twiddle needs no companion),
//              v = (ph - Q*q) + pl (DFMA, exact because the result is an integer below 2^53, + DADD); |v| <= 0.6q (exactness
//              never depends on the estimate, only the range plan).  A butterfly is 8 FP64 instructions and no integer instruction, so the
//              FP64 pipe (64 lanes/clk/SM on B200, otherwise idle) does the arithmetic while the integer pipes do
//              the addressing: 2156 G butterflies/s register-resident vs 863 G (Shoup) and 1544 G (signed-lazy IMAD).
#pragma once
#include "modarith.cuh"

#ifndef ABC_MINB
#define ABC_MINB 2
#endif
#ifndef ABC_F64_TW_PAIRS
#define ABC_F64_TW_PAIRS 0  /* exact-double class, strided passes: 16-byte {w, w/q} twiddles (0: 8-byte w + one DMUL) */
#endif
#ifndef ABC_F64_FRND
#define ABC_F64_FRND 1      /* exact-double class: quotient estimate as DMUL + FRND.F64 (1) or DFMA + DADD on the 1.5*2^52 magic (0) */
#endif
#ifndef ABC_WHATIF
#define ABC_WHATIF 0        /* timing-only experiments (wrong results): see tools/ks_time.py; 0 in every shipped build */
#endif
enum { AR_SHOUP = 0, AR_FP = 1, AR_FP_LAZY = 2, AR_F64 = 3 };
#define ABC_RINT_MAGIC 6755399441055744.0 /* 1.5 * 2^52: x + MAGIC - MAGIC = rint(x) for |x| < 2^51 */
__device__ __forceinline__ double f64_of(u64 bits) { return __longlong_as_double((long long)bits); }
__device__ __forceinline__ u64 bits_of(double d) { return (u64)__double_as_longlong(d); }
// rint(x * c) for the quotient estimates of the exact-double class.  FRND.F64 runs beside the FP64 pipe (ubench:
// 16 lanes/clk/SM against 64 for DFMA), so the estimate costs one FP64-pipe instruction instead of two; the two forms
// may differ by one unit on near-ties, which moves a centred remainder by q inside its range plan and never shows in a
// canonical result.
__device__ __forceinline__ double rint_mul(double x, double c) {
#if ABC_F64_FRND
  double r;
  asm("cvt.rni.f64.f64 %0, %1;" : "=d"(r) : "d"(x * c));
  return r;
#else
  return fma(x, c, ABC_RINT_MAGIC) - ABC_RINT_MAGIC;
#endif
}

// ---- exact-double class on WIDE primes (0.97 * 2^45 <= q < 2^49: SEAL's N = 16384 defaults are 48- and 49-bit).
// The modular product itself is exact for any q < 2^51 once the quotient estimate is FRND(ph * (1/q)) (no 2^51 magic
// range); what a wider prime costs is headroom: every value must stay an exact integer below 2^53 = 16 q, and the
// estimate's three roundings make a centred product |v| <= (0.5 + 3 q |y| / 2^53) q <= (0.5 + 0.1875 |y| / q) q.  Range
// plan (generic stage plan only: N = 16384, and N = 4096 if asked), magnitudes in units of q, input <= 2.1 (a source
// residue of a 49-bit prime under a 48-bit target):
//   forward : 6 stages unreduced reach 10.4; ONE reduction (to 0.5) in the load of the third strided pass; the remaining
//             3 + 5 stages reach 9.9.  Raw outputs for the key switch's ModUp block are reduced once more (0.5), so a
//             sum of L <= 15 key products stays below 15 * 0.6 = 9.
//   inverse : sums double per stage, so at most 3 stages follow a reduction: after the in-register stages of the
//             contiguous pass, in the load of every strided pass, and after the folded last stage (its outputs feed
//             canon_inv / ModDown, which expect |x| < q).
// The extra reductions are taken at run time (`wide`, uniform per row: q of the row's modulus), so one instantiation
// serves rows of both kinds (the BEHZ product mixes 49-bit q rows with 44-bit auxiliary rows in one launch).
#define ABC_F64_NARROW_MAX 34128100000000ull      /* 0.97 * 2^45 */
#define ABC_F64_WIDE_MAX   562949953421312ull     /* 2^49 */
__device__ __forceinline__ bool f64_wide(u64 q) { return ABC_F64_FRND && q >= ABC_F64_NARROW_MAX; }

// element index -> physical index; keeps (even, odd) pairs adjacent so 16-byte accesses stay legal
// (bits 4..6 of e onto bits 1..3: strided and contiguous passes are conflict free; bit 7 onto bit 3: neighbouring
// 128-coefficient blocks are staggered by 64 bytes, which a pass with 8 lanes per block — the radix-16 plan's — needs)
__device__ __forceinline__ int swz(int e) { return e ^ ((e >> 3) & 14) ^ ((e >> 4) & 8); }
// swz(8 * vt + 2 * i) = swz_row8(vt) ^ (2 * i): one XOR per access in the contiguous pass instead of a swizzle each
__device__ __forceinline__ int swz_row8(int vt) { return (8 * vt) ^ (vt & 14) ^ ((vt >> 1) & 8); }
// sm address of element base + (r << LG) of a strided pass.  The swizzle is linear over XOR and base has no bit in
// common with r << LG, so swz(base + (r << LG)) = swz(base) ^ swz(r << LG): ONE xor with a compile-time constant.
template <int LG> __device__ __forceinline__ int swz_strided(int base, int pbase, int r) {
  (void)base;
  return pbase ^ swz(r << LG);
}

// strided passes (radix 2^R each); the contiguous pass (NttLast) takes the remaining stages
template <int LOGN> struct NttPlan;
template <> struct NttPlan<12> { static constexpr int R0 = 3, R1 = 3, R2 = 2; };   // + 1 shuffle + 3 in-register
template <> struct NttPlan<13> { static constexpr int R0 = 3, R1 = 3, R2 = 3; };   // + 1 shuffle + 3 in-register
template <> struct NttPlan<14> { static constexpr int R0 = 3, R1 = 3, R2 = 3; };   // + 2 shuffle + 3 in-register

// TT != 0 overrides the CTA size (the fused key switch runs N = 8192 as ONE 1024-thread CTA per SM: its accumulators
// take the shared memory a second CTA would need)
template <int LOGN, int TT = 0> struct NttDims {
  static constexpr int N = 1 << LOGN;
  static constexpr int T = TT ? TT : ((LOGN >= 14) ? 1024 : ((N / 8 < 512) ? N / 8 : 512));
  static constexpr int MINB = (TT || LOGN >= 14) ? 1 : ABC_MINB;
  static constexpr int IT = N / 8 / T;
  static constexpr size_t SMEM = (size_t)N * 8;
};

// ---- radix-16 plan (exact-double class at N = 8192 with 512-thread CTAs): strided passes of 3, 3 and FOUR stages, then a
// contiguous pass of three in-register stages — no lane-pair shuffle stage, whose 16 shuffles + 48 selects per 8
// coefficients were 7 % of a row's instructions.  The 4-stage pass keeps one group of 16 coefficients per thread (the
// same 32 data registers as two interleaved groups of 8).  Thread mapping: the 128 threads that share a named barrier
// after the second pass (tid >> 7 = g) wrote the 1024-coefficient blocks g and g + 4, so they also run those blocks'
// 4-stage pass (64 groups of 16 each), and every warp of that pass then owns 64 consecutive 8-coefficient blocks,
// which the same warp takes through the contiguous pass: the two new boundaries cost a named barrier and a warp barrier.
#ifndef ABC_PLAN16
#define ABC_PLAN16 1
#endif
template <int LOGN, int AR, int TT> struct UsePlan16 {
  static constexpr bool value = ABC_PLAN16 && LOGN == 13 && AR == 3 /* AR_F64 */ && TT == 0 && NttDims<13, 0>::T == 512;
};
// 1024-coefficient block of thread tid in the 4-stage pass, and its group of 16 inside (0..63)
__device__ __forceinline__ int p16_kblock(int tid) { return (tid >> 7) + 4 * ((tid >> 6) & 1); }
// 8-coefficient block of (tid, g) in the contiguous pass
__device__ __forceinline__ int p16_block8(int tid, int g) {
  return (p16_kblock(tid) << 7) + (((tid >> 5) & 1) << 6) + (g << 5) + (tid & 31);
}

// lo64(a*b + c) as one IMAD.WIDE + two IMAD (no separate carry adds): the shape ptxas keeps on the fma pipe
__device__ __forceinline__ u64 mad_lo64(u64 a, u64 b, u64 c) {
  u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32), lo, hi;
  u64 t;
  // volatile: keeps ptxas from re-associating the chain into IMAD.WIDE(+0) + IADD3 + IMAD.X
  asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t) : "r"(a0), "r"(b0), "l"(c));
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(t));
  asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(a0), "r"(b1));
  asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(a1), "r"(b0));
  asm("mov.b64 %0, {%1, %2};" : "=l"(t) : "r"(lo), "r"(hi));
  return t;
}

// Per-class constants next to q: `aux` is 2q for the guarded classes and MAGIC*q (mod 2^64) for AR_FP_LAZY.
#define ABC_MAGIC_U 0x4330000000000000ULL   /* bits of 2^52         : unsigned encoding, y in [0, 2^52)      */
#define ABC_MAGIC_S 0x4338000000000000ULL   /* bits of 2^52 + 2^51  : signed encoding,   y in [-2^51, 2^51)  */
// `aux`: 2q (guarded classes), MAGIC_S*q mod 2^64 (AR_FP_LAZY), bits of double(q) (AR_F64)
template <int AR> __device__ __forceinline__ u64 ar_aux(u64 q) {
  if (AR == AR_F64) return bits_of((double)q);
  return AR == AR_FP_LAZY ? ABC_MAGIC_S * q : 2 * q;
}
// class representation of a canonical residue (< 2^52), and sums / differences in it
template <int AR> __device__ __forceinline__ u64 ar_from_canon(u64 x) {
  return AR == AR_F64 ? bits_of(f64_of(x | ABC_MAGIC_U) - 4503599627370496.0) : x;
}
template <int AR> __device__ __forceinline__ u64 ar_add(u64 a, u64 b) { return AR == AR_F64 ? bits_of(f64_of(a) + f64_of(b)) : a + b; }
template <int AR> __device__ __forceinline__ u64 ar_sub(u64 a, u64 b) { return AR == AR_F64 ? bits_of(f64_of(a) - f64_of(b)) : a - b; }

// ---- twiddle fetch.  Integer classes: {w, companion} pairs, 16 bytes.  AR_F64, strided passes: the same, from the
// small table twp (stages 0 .. R0+R1+R2-1: at most 512 entries, a warp reads one or two of them per load, L1-resident).
// AR_F64, contiguous pass (every lane its own twiddles): the table twd holds double(w) only
// (8 bytes, half the L1 data-pipe wavefronts; ncu showed that pipe at 81 % with 16-byte twiddles) and the companion
// w/q is one DMUL, w * (1/q): two roundings instead of one, so |product| <= 0.6q instead of 0.53q (range plan below).
// In the AR_F64 table the last two stages are stored lane-contiguously for the contiguous pass: the twiddle thread vt
// needs for its c-th butterfly group of stage s lives at 2^s + c * (N/8) + vt, so a warp's load is one 256-byte run.
template <int AR> __device__ __forceinline__ ulonglong2 tw_get(const ulonglong2 *__restrict__ tw, u32 idx, double qinv) {
  if (AR == AR_F64) {
#if ABC_WHATIF & 1
    idx &= 63u;   // what-if: every twiddle of the contiguous pass L1-hot
#endif
    const u64 w = __ldg(reinterpret_cast<const u64 *>(tw) + idx);
    return make_ulonglong2(w, bits_of(qinv));
  }
  return __ldg(tw + idx);
}

// twiddle of a strided pass
template <int AR> __device__ __forceinline__ ulonglong2 mid_tw(const ulonglong2 *__restrict__ tw, u32 idx, double qinv) {
  if (AR == AR_F64) {
    if (!ABC_F64_TW_PAIRS) return tw_get<AR>(tw, idx, qinv);
    return make_ulonglong2(__ldg(tw + idx).x, bits_of(qinv));
  }
  return __ldg(tw + idx);
}

// ---- modular product y*w with a precomputed companion c
// AR_SHOUP   c = floor(w*2^64/q), any 64-bit y, result in [0,2q).
// AR_FP      c = bits of double(w/q), 0 <= y < 2^51, result in (0,2q).
// AR_FP_LAZY c = bits of double(w/q), |y| < 2^51 (two's complement), result centred: |r| <= 0.75q.
//            t = y*(w/q) + (2^52+2^51) rounds to an integer, so bits(t) = MAGIC_S + qr with qr = round(y*w/q);
//            r = y*w - qr*q = y*w + bits(t)*(-q) + MAGIC_S*q  (mod 2^64): no mask, no offset.
template <int AR> __device__ __forceinline__ u64 mul_tw(u64 y, u64 w, u64 c, u64 q, u64 aux) {
  if (AR == AR_F64) {  // y, w: bits of integer-valued doubles; c: bits of double(1/q) (no per-twiddle companion); aux: bits of double(q)
    // the quotient estimate comes from the rounded product itself, Q = rint(fl(y*w) * fl(1/q)): two roundings like the
    // old companion w * (1/q), so the same |v| <= 0.6q, but no DMUL per twiddle and two registers fewer per twiddle
    const double yd = f64_of(y), wd = f64_of(w);
    const double ph = yd * wd;
    const double Q = rint_mul(ph, f64_of(c));
    const double pl = fma(yd, wd, -ph);
    return bits_of(fma(-Q, f64_of(aux), ph) + pl);
  } else if (AR == AR_SHOUP) {
    return y * w - __umul64hi(y, c) * q;
  } else if (AR == AR_FP) {
    const double yd = __longlong_as_double((long long)(y | ABC_MAGIC_U)) - 4503599627370496.0;
    const double t = fma(yd, __longlong_as_double((long long)c), 4503599627370496.0);
    const u64 qr = (u64)__double_as_longlong(t) & 0x000FFFFFFFFFFFFFULL;  // round(y*w/q), off by <= 1
    return mad_lo64(qr, 0 - q, mad_lo64(y, w, q));                        // y*w - qr*q + q  (mod 2^64)
  } else {
    const double yd = __hiloint2double((int)((u32)(y >> 32) + 0x43380000u), (int)(u32)y) - 6755399441055744.0;
    const double t = fma(yd, __longlong_as_double((long long)c), 6755399441055744.0);
    return mad_lo64((u64)__double_as_longlong(t), 0 - q, mad_lo64(y, w, aux));
  }
}
// x mod q via the FP64 pipe.  signed = false: 0 <= x < 2^51 -> (0,2q).  signed = true: |x| < 2^51 -> |r| <= 0.75q.
template <bool SIGNED> __device__ __forceinline__ u64 reduce_fp(u64 x, double qinv, u64 q, u64 aux) {
  if (SIGNED) {
    const double xd = __hiloint2double((int)((u32)(x >> 32) + 0x43380000u), (int)(u32)x) - 6755399441055744.0;
    const double t = fma(xd, qinv, 6755399441055744.0);
    return mad_lo64((u64)__double_as_longlong(t), 0 - q, x + aux);
  } else {
    const double xd = __longlong_as_double((long long)(x | ABC_MAGIC_U)) - 4503599627370496.0;
    const double t = fma(xd, qinv, 4503599627370496.0);
    const u64 qr = (u64)__double_as_longlong(t) & 0x000FFFFFFFFFFFFFULL;
    return mad_lo64(qr, 0 - q, x + q);
  }
}
// canonical residue of a value as the class leaves it after a transform
// AR_F64: x - rint(x/q)*q, |result| <= 0.51q, exact
__device__ __forceinline__ double reduce_f64(double x, double qinv, double qd) {
  const double Q = rint_mul(x, qinv);
  return fma(-Q, qd, x);
}
// r < 0 ? r + q : r  and  r >= q ? r - q : r  as a compare + ONE predicated add (the C++ ternary compiles to
// add + compare + two selects)
__device__ __forceinline__ double cadd_neg(double r, double q) {
  asm("{\n .reg .pred p;\n setp.lt.f64 p, %0, 0d0000000000000000;\n @p add.f64 %0, %0, %1;\n}" : "+d"(r) : "d"(q));
  return r;
}
__device__ __forceinline__ double csub_ge(double r, double q) {
  asm("{\n .reg .pred p;\n setp.ge.f64 p, %0, %1;\n @p sub.f64 %0, %0, %1;\n}" : "+d"(r) : "d"(q));
  return r;
}
// centred double in (-q, q) -> canonical u64
__device__ __forceinline__ u64 f64_to_canon(double r, double qd) {
  r = cadd_neg(r, qd);
  return bits_of(r + 4503599627370496.0) & 0x000FFFFFFFFFFFFFULL;
}
template <int AR> __device__ __forceinline__ u64 canon_fwd(u64 x, const ModInfo &M, u64 q, u64 aux) {
  if (AR == AR_F64) return f64_to_canon(reduce_f64(f64_of(x), f64_of(M.qinv_bits), f64_of(aux)), f64_of(aux));
  if (AR == AR_FP_LAZY) {
    const u64 r = reduce_fp<true>(x, __longlong_as_double((long long)M.qinv_bits), q, aux);  // |r| <= 0.75q
    return r + (((long long)r >> 63) & q);
  }
  return csub(csub(x, aux), q);  // [0,4q) -> [0,q)
}
template <int AR> __device__ __forceinline__ u64 canon_inv(u64 x, u64 q, u64 aux) {
  if (AR == AR_F64) return f64_to_canon(f64_of(x), f64_of(aux));
  if (AR == AR_FP_LAZY) return x + (((long long)x >> 63) & q);  // centred product of the folded last stage
  return csub(x, q);                                              // [0,2q) -> [0,q)
}

// forward (Cooley-Tukey) butterfly.  Guarded classes: x,y in [0,4q) -> [0,4q).  AR_FP_LAZY: signed, |.| grows by 0.75q.
template <int AR> __device__ __forceinline__ void bf_fwd(u64 &x, u64 &y, ulonglong2 w, u64 q, u64 aux) {
  if (AR == AR_FP_LAZY || AR == AR_F64) {
    const u64 v = mul_tw<AR>(y, w.x, w.y, q, aux);
    y = ar_sub<AR>(x, v);
    x = ar_add<AR>(x, v);
  } else {
    const u64 u = csub(x, aux);
    const u64 v = mul_tw<AR>(y, w.x, w.y, q, aux);
    x = u + v;
    y = u + aux - v;
  }
}
// inverse (Gentleman-Sande) butterfly.  Guarded classes: x,y in [0,2q) -> [0,2q).
// AR_FP_LAZY: signed; the sum doubles per stage (reduced at pass boundaries), the product is centred.
template <int AR> __device__ __forceinline__ void bf_inv(u64 &x, u64 &y, ulonglong2 w, u64 q, u64 aux) {
  const u64 u = x, v = y;
  if (AR == AR_FP_LAZY || AR == AR_F64) {
    x = ar_add<AR>(u, v);
    y = mul_tw<AR>(ar_sub<AR>(u, v), w.x, w.y, q, aux);
  } else {
    x = csub(u + v, aux);
    y = mul_tw<AR>(u + aux - v, w.x, w.y, q, aux);
  }
}

// ---- strided pass: stages S0 .. S0+R-1 (stage s has 2^s groups, gap N >> (s+1)); twbase = 1 for a whole
// transform, (2^a + block) when this limb is block `block` of the tail of a larger 2^(a+LOGN) transform.
// LINSRC (first pass of AR_F64 only): the limb sits in shared memory exactly as it lies in global memory (a bulk
// copy, no swizzle) holding residues of the SOURCE prime qs; the pass reads coefficient e from
// +-sm[e * einv mod 2N] (GaloisTool::apply_galois as a gather, einv = 0: identity), takes NO reduction modulo the
// target prime (an exact-double transform accepts any |x| < 2^45 and canon_fwd reduces at the end), and, because
// it is no longer in place, separates its loads from its stores with a barrier.
// PRECONV: shared memory already holds the class's representation (the first pass of a block whose stage above it was
// done while loading)
template <int LOGN, int S0, int R, int AR, bool LINSRC = false, int TT = 0, bool PRECONV = false>
__device__ __forceinline__ void ntt_fwd_mid(u64 *sm, const ulonglong2 *__restrict__ tw, u32 twbase, u64 q, u64 aux,
                                            int tid, double qinv, u64 qs = 0, u32 einv = 0, bool wide_reduce = false) {
  typedef NttDims<LOGN, TT> D;
  constexpr int LG = LOGN - S0 - R;
  static_assert(!LINSRC || (AR == AR_F64 && S0 == 0), "linear-source gather is the first pass of the exact-double class");
  // AR_F64 is bound by FP64 latency, not issue: its IT groups of 8 are loaded together and their butterflies
  // interleaved (2x the independent chains); the integer classes keep one group live (register pressure).
  constexpr int G = (AR == AR_F64) ? D::IT : 1;
#pragma unroll
  for (int it0 = 0; it0 < D::IT; it0 += G) {
    int blk[G], base[G], pbase[G];
    u64 x[G][8];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int vt = tid + (it0 + g) * D::T;
      blk[g] = vt >> LG;
      base[g] = (blk[g] << (LG + 3)) | (vt & ((1 << LG) - 1));
      pbase[g] = swz(base[g]);
      if (LINSRC) {
        // index e * einv mod 2N: the low log2(N) bits address the source, bit log2(N) is the sign; r0 is left to wrap
        // (2N divides 2^32), and the identity is just einv = 1 (no sign bit ever set: e < N)
        const u32 e1 = einv ? einv : 1u, step = e1 << LG;
        u32 r0 = (u32)base[g] * e1;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          u64 v = sm[r0 & (D::N - 1)];
          // negated position and v != 0: v = qs - v, as two predicated subtractions (the C++ form compiles to two
          // levels of 64-bit selects: 16 instructions per element instead of 10)
          asm("{\n .reg .pred p;\n .reg .b32 lo, hi, t;\n mov.b64 {lo, hi}, %0;\n or.b32 t, lo, hi;\n setp.ne.u32 p, t, 0;\n"
              " and.b32 t, %2, %3;\n setp.ne.and.u32 p, t, 0, p;\n @p sub.u64 %0, %1, %0;\n}"
              : "+l"(v) : "l"(qs), "r"(r0), "r"((u32)D::N));
          x[g][r] = ar_from_canon<AR>(v);
          r0 += step;
        }
        if (qs >= ABC_F64_NARROW_MAX && qs > q) {   // a wide source prime above the target: bring the residues to |x| <= 0.5 q first
#pragma unroll
          for (int r = 0; r < 8; ++r) x[g][r] = bits_of(reduce_f64(f64_of(x[g][r]), qinv, f64_of(aux)));
        }
      } else {
#pragma unroll
        for (int r = 0; r < 8; ++r) x[g][r] = sm[swz_strided<LG>(base[g], pbase[g], r)];
        if (S0 == 0 && !PRECONV) {  // first pass: canonical residues -> the class's representation
#pragma unroll
          for (int r = 0; r < 8; ++r) x[g][r] = ar_from_canon<AR>(x[g][r]);
        }
        if (AR == AR_F64 && S0 != 0 && wide_reduce) {  // wide primes: the one reduction of the forward range plan
#pragma unroll
          for (int r = 0; r < 8; ++r) x[g][r] = bits_of(reduce_f64(f64_of(x[g][r]), qinv, f64_of(aux)));
        }
      }
    }
    if (LINSRC) {
      static_assert(!LINSRC || G == D::IT, "every group is in registers before the first store");
      __syncthreads();
    }
#pragma unroll
    for (int b = R - 1; b >= 0; --b) {
      const int s = S0 + R - 1 - b;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (r & (1 << b)) continue;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const ulonglong2 w = mid_tw<AR>(tw, (twbase << s) + ((u32)((blk[g] << 3) + r) >> (b + 1)), qinv);
          bf_fwd<AR>(x[g][r], x[g][r | (1 << b)], w, q, aux);
        }
      }
    }
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int r = 0; r < 8; ++r) sm[swz_strided<LG>(base[g], pbase[g], r)] = x[g][r];
  }
}

// REDUCE (AR_FP_LAZY only): bring the loaded values back to |x| <= 0.75q before this pass's stages
// EPI: what happens to the pass's results.  StoreSmem (default) writes them back in place; any other type is called as
// epi(e, bits) with the coefficient index e and the value in the class's representation, and nothing is stored (the
// last pass of a transform handing its outputs straight to an epilogue, e.g. the key switch's ModDown).
struct StoreSmem {};
template <class A, class B> struct SameType { static constexpr bool value = false; };
template <class A> struct SameType<A, A> { static constexpr bool value = true; };
template <int LOGN, int S0, int R, bool FOLD, bool REDUCE, int AR, int TT = 0, class EPI = StoreSmem>
__device__ __forceinline__ void ntt_inv_mid(u64 *sm, const ModInfo &M, u32 twbase, u64 q, u64 aux, int tid, EPI epi = EPI()) {
  typedef NttDims<LOGN, TT> D;
  constexpr int LG = LOGN - S0 - R;
  constexpr int G = (AR == AR_F64) ? D::IT : 1;
  const ulonglong2 *__restrict__ tw = (AR == AR_SHOUP) ? M.itw : (AR == AR_F64 ? (ABC_F64_TW_PAIRS ? M.itwp : M.itwd) : M.itwf);
  const double qinv = f64_of(M.qinv_bits);
#pragma unroll
  for (int it0 = 0; it0 < D::IT; it0 += G) {
    int blk[G], base[G], pbase[G];
    u64 x[G][8];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int vt = tid + (it0 + g) * D::T;
      blk[g] = vt >> LG;
      base[g] = (blk[g] << (LG + 3)) | (vt & ((1 << LG) - 1));
      pbase[g] = swz(base[g]);
#pragma unroll
      for (int r = 0; r < 8; ++r) x[g][r] = sm[swz_strided<LG>(base[g], pbase[g], r)];
      if ((AR == AR_FP_LAZY || AR == AR_F64) && (REDUCE || (AR == AR_F64 && f64_wide(q)))) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
          x[g][r] = AR == AR_F64 ? bits_of(reduce_f64(f64_of(x[g][r]), qinv, f64_of(aux))) : reduce_fp<true>(x[g][r], qinv, q, aux);
      }
    }
#pragma unroll
    for (int b = 0; b < R; ++b) {
      const int s = S0 + R - 1 - b;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (r & (1 << b)) continue;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          if (FOLD && s == 0) {
            // last stage of the whole transform: fold N^-1 into both outputs
            const u64 u = x[g][r], v = x[g][r | (1 << b)];
            const u64 d = (AR == AR_FP_LAZY || AR == AR_F64) ? ar_sub<AR>(u, v) : u + aux - v;
            x[g][r] = mul_tw<AR>(ar_add<AR>(u, v), AR == AR_F64 ? M.ninv_d : M.ninv,
                                 (AR == AR_SHOUP) ? M.ninv_s : (AR == AR_F64 ? M.qinv_bits : M.ninv_f), q, aux);
            x[g][r | (1 << b)] = mul_tw<AR>(d, AR == AR_F64 ? M.wl_ninv_d : M.wl_ninv,
                                            (AR == AR_SHOUP) ? M.wl_ninv_s : (AR == AR_F64 ? M.qinv_bits : M.wl_ninv_f), q, aux);
            if (AR == AR_F64 && f64_wide(q)) {   // wide primes: |product| can reach 1.25 q; consumers expect |x| < q
              x[g][r] = bits_of(reduce_f64(f64_of(x[g][r]), qinv, f64_of(aux)));
              x[g][r | (1 << b)] = bits_of(reduce_f64(f64_of(x[g][r | (1 << b)]), qinv, f64_of(aux)));
            }
          } else {
            const ulonglong2 w = mid_tw<AR>(tw, (twbase << s) + ((u32)((blk[g] << 3) + r) >> (b + 1)), qinv);
            bf_inv<AR>(x[g][r], x[g][r | (1 << b)], w, q, aux);
          }
        }
      }
    }
    if constexpr (SameType<EPI, StoreSmem>::value) {
#pragma unroll
      for (int g = 0; g < G; ++g)
#pragma unroll
        for (int r = 0; r < 8; ++r) sm[swz_strided<LG>(base[g], pbase[g], r)] = x[g][r];
    } else {
#pragma unroll
      for (int g = 0; g < G; ++g)
#pragma unroll
        for (int r = 0; r < 8; ++r) epi(base[g] + (r << LG), x[g][r]);
    }
  }
}

// ---- the 4-stage strided pass of the radix-16 plan (stages 6..9 of N = 8192: gap 8 between a thread's 16 coefficients)
template <int LOGN, int AR>
__device__ __forceinline__ void ntt_fwd_mid16(u64 *sm, const ulonglong2 *__restrict__ tw, u64 q, u64 aux, int tid, double qinv,
                                              u32 twbase = 1u, bool wide_reduce = false) {
  constexpr int S0 = 6, LG = LOGN - S0 - 4;
  static_assert(LOGN == 13 && LG == 3, "radix-16 plan is laid out for N = 8192");
  const int blk = (p16_kblock(tid) << 3) + ((tid & 63) >> LG);   // 128-coefficient block
  // swz(base + 8r) = pb ^ C_r with C_r = (8r) ^ (r & 14) known at compile time (base = blk << 7 | column)
  const int pb = ((blk << (LG + 4)) | (tid & ((1 << LG) - 1))) ^ ((blk & 1) << 3);
  u64 x[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) x[r] = sm[pb ^ ((r << LG) ^ (r & 14))];
  if (AR == AR_F64 && wide_reduce) {   // wide primes: the forward range plan's one reduction (7 stages in, 7 to go)
#pragma unroll
    for (int r = 0; r < 16; ++r) x[r] = bits_of(reduce_f64(f64_of(x[r]), qinv, f64_of(aux)));
  }
#pragma unroll
  for (int b = 3; b >= 0; --b) {
    const int s = S0 + 3 - b;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      if (r & (1 << b)) continue;
      const ulonglong2 w = mid_tw<AR>(tw, (twbase << s) + ((u32)((blk << 4) + r) >> (b + 1)), qinv);
      bf_fwd<AR>(x[r], x[r | (1 << b)], w, q, aux);
    }
  }
#pragma unroll
  for (int r = 0; r < 16; ++r) sm[pb ^ ((r << LG) ^ (r & 14))] = x[r];
}
template <int LOGN, int AR, bool REDUCE>
__device__ __forceinline__ void ntt_inv_mid16(u64 *sm, const ulonglong2 *__restrict__ tw, u64 q, u64 aux, int tid, double qinv,
                                              u32 twbase = 1u) {
  constexpr int S0 = 6, LG = LOGN - S0 - 4;
  static_assert(LOGN == 13 && LG == 3 && AR == AR_F64, "radix-16 plan: exact-double class at N = 8192");
  const int blk = (p16_kblock(tid) << 3) + ((tid & 63) >> LG);
  const int pb = ((blk << (LG + 4)) | (tid & ((1 << LG) - 1))) ^ ((blk & 1) << 3);
  u64 x[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) x[r] = sm[pb ^ ((r << LG) ^ (r & 14))];
  if (REDUCE) {
#pragma unroll
    for (int r = 0; r < 16; ++r) x[r] = bits_of(reduce_f64(f64_of(x[r]), qinv, f64_of(aux)));
  }
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const int s = S0 + 3 - b;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      if (r & (1 << b)) continue;
      const ulonglong2 w = mid_tw<AR>(tw, (twbase << s) + ((u32)((blk << 4) + r) >> (b + 1)), qinv);
      bf_inv<AR>(x[r], x[r | (1 << b)], w, q, aux);
    }
  }
#pragma unroll
  for (int r = 0; r < 16; ++r) sm[pb ^ ((r << LG) ^ (r & 14))] = x[r];
}

// ---- contiguous pass: every thread owns E = 8 CONSECUTIVE coefficients per group: gaps 4,2,1 are in-register, the
// NSH stages above them (gaps 8, 16) are warp-shuffle butterflies between lane pairs, each lane computing half of the
// pair's butterflies.
template <int LOGN, int TT = 0> struct NttLast {
  // E = 16 (no shuffle stage at N = 8192) was measured slower than two groups of 8 with one shuffle stage
  // (13.7 vs 15.0 Mrows/s forward): the 32 live data registers cost more than the shuffles save.
  static constexpr int E = 8, LOGE = (E == 16) ? 4 : 3;
  static constexpr int GROUPS = 8 * NttDims<LOGN, TT>::IT / E;
  static constexpr int NSH = LOGN - (NttPlan<LOGN>::R0 + NttPlan<LOGN>::R1 + NttPlan<LOGN>::R2) - LOGE;
  static_assert(E == 8, "contiguous pass handles 8 or 16 coefficients per thread");
  static_assert(NSH >= 0 && NSH <= 2, "stage plan does not add up");
};

// twiddle index of the in-register stage with partner distance 2^b of the contiguous pass (E = 8): natural order
// 2^s + ((8*vt + r) >> (b+1)); AR_F64 stores the last two stages lane-contiguously (see tw_get)
// When this limb is block `blk` of a 2^sub times larger transform (twbase = 2^sub + blk; exact-double class: sub <= 1, the
// split N = 16384 key switch), local stage s is global stage s + sub, the lane-contiguous stride is the larger transform's
// N / 8 and this block's contiguous-pass threads are blk * (N_local / 8) + vt of it.
template <int LOGN, int AR>
__device__ __forceinline__ u32 last_tw_index(u32 twbase, int s, int b, int vt, int r) {
  if (AR == AR_F64 && b < 2) {
    const u32 sub = twbase >> 1 ? 1u : 0u, blk = twbase - (1u << sub);   // twbase in {1, 2, 3}
    return (1u << (s + sub)) + (u32)(r >> (b + 1)) * (u32)((NttDims<LOGN>::N << sub) / 8) + blk * (u32)(NttDims<LOGN>::N / 8) + (u32)vt;
  }
  return (twbase << s) + ((u32)(8 * vt + r) >> (b + 1));
}

// L1 prefetch of the twiddles this thread's contiguous pass will ask for (exact-double class: 7 eight-byte entries per group
// of 8 coefficients, lane-contiguous for the last two stages): issued a pass early, the lines arrive while the pass in between
// does its butterflies, so the contiguous pass's loads hit L1 instead of waiting on L2 (the kernels are latency-bound at 8
// warps per scheduler; timing what-if "twiddles always L1-hot": 3.7 % of the key switch)
#ifndef ABC_TW_PREFETCH
#define ABC_TW_PREFETCH 1
#endif
template <int LOGN, int AR, int TT = 0>
__device__ __forceinline__ void tw_prefetch_last(const ulonglong2 *__restrict__ tw, u32 twbase, int tid) {
#if ABC_TW_PREFETCH
  if constexpr (AR == AR_F64) {
    constexpr bool P16 = UsePlan16<LOGN, AR, TT>::value;
    const u64 *t8 = reinterpret_cast<const u64 *>(tw);
#pragma unroll
    for (int g = 0; g < NttLast<LOGN, TT>::GROUPS; ++g) {
      const int vt = P16 ? p16_block8(tid, g) : tid + g * NttDims<LOGN, TT>::T;
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const int s = LOGN - 1 - b;
#pragma unroll
        for (int r = 0; r < 8; r += (2 << b))
          asm volatile("prefetch.global.L1 [%0];" ::"l"(t8 + last_tw_index<LOGN, AR>(twbase, s, b, vt, r)));
      }
    }
  }
#endif
}

// the butterflies of the contiguous forward pass on one thread's 8 consecutive coefficients (vt = first / 8)
template <int LOGN, int AR, int TT = 0, int NSH = NttLast<LOGN, TT>::NSH>
__device__ __forceinline__ void ntt_fwd_last_math(u64 (&x)[8], const ulonglong2 *__restrict__ tw, u32 twbase, u64 q, u64 aux,
                                                  int vt, double qinv) {
  typedef NttLast<LOGN, TT> P;
  constexpr int E = P::E, H = E / 2;
#pragma unroll
  for (int j = NSH - 1; j >= 0; --j) {
    const int s = LOGN - 1 - P::LOGE - j;
    const ulonglong2 w = tw_get<AR>(tw, (twbase << s) + ((u32)vt >> (1 + j)), qinv);
    const bool hi = (vt >> j) & 1;
#pragma unroll
    for (int r = 0; r < H; ++r) {
      u64 recv = __shfl_xor_sync(0xffffffffu, hi ? x[r] : x[H + r], 1 << j);
      u64 a = hi ? recv : x[r];
      u64 b = hi ? x[H + r] : recv;
      bf_fwd<AR>(a, b, w, q, aux);
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
      const ulonglong2 w = tw_get<AR>(tw, last_tw_index<LOGN, AR>(twbase, s, b, vt, r), qinv);
      bf_fwd<AR>(x[r], x[r | (1 << b)], w, q, aux);
    }
  }
}
// RAW (exact-double class): leave the outputs as the lazy doubles they are (|x| < (log2(N) + 2) q) instead of canonical
// residues — for consumers that multiply them in the same arithmetic (the key switch's ModUp block)
template <int LOGN, int AR, int TT = 0, bool RAW = false>
__device__ __forceinline__ void ntt_fwd_last(u64 *sm, const ModInfo &M, u32 twbase, u64 q, u64 aux, int tid, bool raw_reduce = false) {
  typedef NttLast<LOGN, TT> P;
  constexpr int E = P::E, H = E / 2;
  const ulonglong2 *__restrict__ tw = (AR == AR_SHOUP) ? M.tw : (AR == AR_F64 ? M.twd : M.twf);
  const double qinv = f64_of(M.qinv_bits);
  constexpr bool P16 = UsePlan16<LOGN, AR, TT>::value;
#pragma unroll
  for (int g = 0; g < P::GROUPS; ++g) {
    const int vt = P16 ? p16_block8(tid, g) : tid + g * NttDims<LOGN, TT>::T;
    u64 x[E];
#pragma unroll
    for (int i = 0; i < H; ++i) {
      ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(&sm[swz_row8(vt) ^ (2 * i)]);
      x[2 * i] = v.x; x[2 * i + 1] = v.y;
    }
    ntt_fwd_last_math<LOGN, AR, TT, P16 ? 0 : P::NSH>(x, tw, twbase, q, aux, vt, qinv);
    if (RAW && AR == AR_F64 && (raw_reduce || f64_wide(q))) {   // wide primes (or on request): raw values back to |x| <= 0.5 q
#pragma unroll
      for (int r = 0; r < E; ++r) x[r] = bits_of(reduce_f64(f64_of(x[r]), qinv, f64_of(aux)));
    }
#pragma unroll
    for (int i = 0; i < H; ++i) {
      ulonglong2 v;
      v.x = RAW ? x[2 * i] : canon_fwd<AR>(x[2 * i], M, q, aux);
      v.y = RAW ? x[2 * i + 1] : canon_fwd<AR>(x[2 * i + 1], M, q, aux);
      *reinterpret_cast<ulonglong2 *>(&sm[swz_row8(vt) ^ (2 * i)]) = v;
    }
  }
}

// REPIN: shared memory already holds the class's representation (the fused key-switch inner product stores doubles)
// the butterflies of the contiguous inverse pass on one thread's 8 consecutive coefficients (class representation in x)
template <int LOGN, int AR, int TT = 0, int NSH = NttLast<LOGN, TT>::NSH>
__device__ __forceinline__ void ntt_inv_first_math(u64 (&x)[8], const ulonglong2 *__restrict__ tw, u32 twbase, u64 q, u64 aux,
                                                   int vt, double qinv) {
  typedef NttLast<LOGN, TT> P;
  constexpr int E = P::E, H = E / 2;
#pragma unroll
  for (int b = 0; b < P::LOGE; ++b) {
    const int s = LOGN - 1 - b;
#pragma unroll
    for (int r = 0; r < E; ++r) {
      if (r & (1 << b)) continue;
      const ulonglong2 w = tw_get<AR>(tw, last_tw_index<LOGN, AR>(twbase, s, b, vt, r), qinv);
      bf_inv<AR>(x[r], x[r | (1 << b)], w, q, aux);
    }
  }
  if (AR == AR_F64 && NSH > 0 && f64_wide(q)) {   // wide primes: at most 3 stages between reductions
#pragma unroll
    for (int r = 0; r < E; ++r) x[r] = bits_of(reduce_f64(f64_of(x[r]), qinv, f64_of(aux)));
  }
#pragma unroll
  for (int j = 0; j < NSH; ++j) {
    const int s = LOGN - 1 - P::LOGE - j;
    const ulonglong2 w = tw_get<AR>(tw, (twbase << s) + ((u32)vt >> (1 + j)), qinv);
    const bool hi = (vt >> j) & 1;
#pragma unroll
    for (int r = 0; r < H; ++r) {
      u64 recv = __shfl_xor_sync(0xffffffffu, hi ? x[r] : x[H + r], 1 << j);
      u64 a = hi ? recv : x[r];
      u64 b = hi ? x[H + r] : recv;
      bf_inv<AR>(a, b, w, q, aux);
      u64 got = __shfl_xor_sync(0xffffffffu, hi ? a : b, 1 << j);
      x[r] = hi ? got : a;
      x[H + r] = hi ? b : got;
    }
  }
}
template <int LOGN, int AR, bool REPIN = false, int TT = 0>
__device__ __forceinline__ void ntt_inv_first(u64 *sm, const ModInfo &M, u32 twbase, u64 q, u64 aux, int tid) {
  typedef NttLast<LOGN, TT> P;
  constexpr int E = P::E, H = E / 2;
  const ulonglong2 *__restrict__ tw = (AR == AR_SHOUP) ? M.itw : (AR == AR_F64 ? M.itwd : M.itwf);
  const double qinv = f64_of(M.qinv_bits);
  constexpr bool P16 = UsePlan16<LOGN, AR, TT>::value;
#pragma unroll
  for (int g = 0; g < P::GROUPS; ++g) {
    const int vt = P16 ? p16_block8(tid, g) : tid + g * NttDims<LOGN, TT>::T;
    u64 x[E];
#pragma unroll
    for (int i = 0; i < H; ++i) {
      ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(&sm[swz_row8(vt) ^ (2 * i)]);
      x[2 * i] = REPIN ? v.x : ar_from_canon<AR>(v.x); x[2 * i + 1] = REPIN ? v.y : ar_from_canon<AR>(v.y);
    }
    ntt_inv_first_math<LOGN, AR, TT, P16 ? 0 : P::NSH>(x, tw, twbase, q, aux, vt, qinv);
#pragma unroll
    for (int i = 0; i < H; ++i) {
      ulonglong2 v; v.x = x[2 * i]; v.y = x[2 * i + 1];
      *reinterpret_cast<ulonglong2 *>(&sm[swz_row8(vt) ^ (2 * i)]) = v;
    }
  }
}

// ---- scoped barriers between passes.  After a strided pass with gap 2^LG the transform splits into independent
// blocks of 2^LG coefficients, and the next pass of such a block only reads what the 2^LG threads sharing tid >> LG
// wrote (the second group of a thread, vt = tid + T, lands in a block owned by the same thread set).  So only the
// first boundary needs the whole CTA; the later ones are a named barrier over 2^LG threads or a warp barrier, and the
// rest of the CTA runs ahead instead of queueing up behind the slowest warp.
template <int LG, int T> __device__ __forceinline__ void pass_sync(int tid) {
  if constexpr ((1 << LG) >= T) {
    __syncthreads();
  } else if constexpr (LG <= 5) {
    __syncwarp();
  } else {
    static_assert(T >> LG <= 15, "named barrier ids 1..15");
    asm volatile("bar.sync %0, %1;" ::"r"(1 + (tid >> LG)), "n"(1 << LG) : "memory");
  }
}

// ---- whole-limb transforms on a swizzled shared-memory limb.  Caller has filled sm[swz(e)] and synced.
// Forward: input canonical (guarded classes accept < 4q), output canonical.  Returns after a barrier.
// the strided passes of the forward transform; returns after the barrier in front of the contiguous pass
template <int LOGN, int AR, bool LINSRC = false, int TT = 0, bool PRECONV = false>
__device__ __forceinline__ void ntt_fwd_smem_mids(u64 *sm, const ModInfo &M, u32 twbase, int tid, u64 qs = 0, u32 einv = 0) {
  typedef NttPlan<LOGN> P;
  const u64 q = M.q, aux = ar_aux<AR>(q);
  const ulonglong2 *tw = (AR == AR_SHOUP) ? M.tw : (AR == AR_F64 ? (ABC_F64_TW_PAIRS ? M.twp : M.twd) : M.twf);
  const double qinv = f64_of(M.qinv_bits);
  typedef NttDims<LOGN, TT> D;
#if ABC_WHATIF & 16
  if constexpr (LINSRC) ntt_fwd_mid<LOGN, P::R0, P::R0, AR, false, TT>(sm, tw, twbase, q, aux, tid, qinv);  // what-if: first pass as cheap as a plain one
  else
#endif
  ntt_fwd_mid<LOGN, 0, P::R0, AR, LINSRC, TT, PRECONV>(sm, tw, twbase, q, aux, tid, qinv, qs, einv);
  pass_sync<LOGN - P::R0, D::T>(tid);
  ntt_fwd_mid<LOGN, P::R0, P::R1, AR, false, TT>(sm, tw, twbase, q, aux, tid, qinv);
  pass_sync<LOGN - P::R0 - P::R1, D::T>(tid);
  if constexpr (UsePlan16<LOGN, AR, TT>::value) {
    tw_prefetch_last<LOGN, AR, TT>(M.twd, twbase, tid);
    ntt_fwd_mid16<LOGN, AR>(sm, tw, q, aux, tid, qinv, twbase, AR == AR_F64 && f64_wide(q));
    __syncwarp();
  } else if constexpr (P::R2 > 0) {
    ntt_fwd_mid<LOGN, P::R0 + P::R1, P::R2, AR, false, TT>(sm, tw, twbase, q, aux, tid, qinv, 0, 0, AR == AR_F64 && f64_wide(q));
    pass_sync<LOGN - P::R0 - P::R1 - P::R2, D::T>(tid);
  }
}
template <int LOGN, int AR, bool LINSRC = false, int TT = 0>
__device__ __forceinline__ void ntt_fwd_smem(u64 *sm, const ModInfo &M, u32 twbase, int tid, u64 qs = 0, u32 einv = 0) {
  ntt_fwd_smem_mids<LOGN, AR, LINSRC, TT>(sm, M, twbase, tid, qs, einv);
  ntt_fwd_last<LOGN, AR, TT>(sm, M, twbase, M.q, ar_aux<AR>(M.q), tid);
  __syncthreads();
}
// Inverse: input canonical; output needs canon_inv<AR> (the caller's copy-out applies it).
// WHOLE folds N^-1 into the last stage; for a tail block (WHOLE=false, guarded classes only) the output stays in
// [0,2q) and a head pass finishes the transform.
// AR_FP_LAZY range plan (q*(log2(N)+2) < 2^51, q < 2^45): the sum chain doubles per stage, so it is reduced to
// |x| <= 0.75q after the contiguous pass (4-5 stages) and again before the last strided pass (3 stages left):
// no product ever sees an operand above 2^6 * 2 * 0.75q < 2^51.
// the strided passes of the inverse transform (everything after the contiguous pass); returns after a barrier
template <int LOGN, bool WHOLE, int AR, int TT = 0>
__device__ __forceinline__ void ntt_inv_smem_mids(u64 *sm, const ModInfo &M, u32 twbase, int tid) {
  typedef NttPlan<LOGN> P;
  const u64 q = M.q, aux = ar_aux<AR>(q);
  typedef NttDims<LOGN, TT> D;
  if constexpr (UsePlan16<LOGN, AR, TT>::value) {
    // range plan (q < 0.97 * 2^45, lib.cu fill_mod): 3 contiguous stages leave |x| <= 4.1q; reduced here, the 4 + 3 stages
    // up to the next reduction reach 0.51q * 2^7 = 65.3q, so the last product operand stays below 2^51
    __syncwarp();
    ntt_inv_mid16<LOGN, AR, true>(sm, (ABC_F64_TW_PAIRS ? M.itwp : M.itwd), q, aux, tid, f64_of(M.qinv_bits), twbase);
    pass_sync<LOGN - P::R0 - P::R1, D::T>(tid);
  } else {
    pass_sync<LOGN - P::R0 - P::R1 - P::R2, D::T>(tid);
    if constexpr (P::R2 > 0) {
      ntt_inv_mid<LOGN, P::R0 + P::R1, P::R2, false, true, AR, TT>(sm, M, twbase, q, aux, tid);
      pass_sync<LOGN - P::R0 - P::R1, D::T>(tid);
    }
  }
  ntt_inv_mid<LOGN, P::R0, P::R1, false, P::R2 == 0, AR, TT>(sm, M, twbase, q, aux, tid);
  pass_sync<LOGN - P::R0, D::T>(tid);
  ntt_inv_mid<LOGN, 0, P::R0, WHOLE, true, AR, TT>(sm, M, twbase, q, aux, tid);
  __syncthreads();
}
template <int LOGN, bool WHOLE, int AR, bool REPIN = false, int TT = 0>
__device__ __forceinline__ void ntt_inv_smem(u64 *sm, const ModInfo &M, u32 twbase, int tid) {
  static_assert(WHOLE || (AR != AR_FP_LAZY && AR != AR_F64), "tail blocks use a guarded class");
  ntt_inv_first<LOGN, AR, REPIN, TT>(sm, M, twbase, M.q, ar_aux<AR>(M.q), tid);
  ntt_inv_smem_mids<LOGN, WHOLE, AR, TT>(sm, M, twbase, tid);
}
