// limb.cuh — the limb pipeline: every NTT of the BFV path, fused with the elementwise work either side of it.
//
// N <= 16384: one CTA = one RNS limb staged whole in shared memory (k_limb):
//   load (pre-op) -> [forward NTT] -> [pointwise * mul row] -> [inverse NTT] -> store (post-op)
// N >= 32768 (a limb is 256-512 KiB, more than an SM's shared memory): two passes.  The first A = log2(N) - 13
//   stages have gaps >= 8192 and run as register butterflies straight from global memory, 2^A strided
//   coefficients per thread (k_head_fwd, carrying the pre-op); the remaining 13 stages are confined to 2^A
//   contiguous blocks of 8192 coefficients, each of which is one k_limb CTA in TAIL mode (twiddle base 2^A + block).
//   The inverse runs the tail first and k_head_inv last (carrying N^-1 and the post-op).
// Pre-ops: ModUp reduction, Galois gather + ModUp, plaintext centred lift, ternary / CBD sampling, BatchEncoder
// scatter.  Post-ops: store, + row, BatchEncoder gather, ModDown with rounding + base accumulate.
#pragma once
#include "devconst.cuh"
#include "ntt.cuh"

enum { PRE_LOAD = 0, PRE_REDUCE = 1, PRE_PLAIN_LIFT = 2, PRE_TERNARY = 3, PRE_CBD = 4, PRE_ENCODE = 5, PRE_GALOIS_REDUCE = 6,
       PRE_KS_INNER = 7, PRE_BEHZ_TENSOR = 8 };
enum { POST_STORE = 0, POST_ADD = 1, POST_DECODE = 2, POST_MODDOWN = 3 };

struct LimbJob {
  u64 *dst; const u64 *src; const u64 *mul; const u64 *add;
  u64 *dst2;                                  // POST_MODDOWN with `add`: also store the result WITHOUT the addend here (layout of dst)
  long long dst_is, src_is, mul_is, add_is;  // per-instance strides in words (0 = shared by all instances)
  const int *rowmod;                          // [W] modulus index of row w
  const int *rowdst;                          // [W] destination row, or nullptr = w
  const int *rowsrc;                          // [W] source row, or nullptr = destination row
  const int *rowmul;                          // [W] row of `mul`/`add`, or nullptr = w
  int n;                                      // coefficients per limb (row stride); = kernel N except in TAIL mode
  int sub;                                    // TAIL mode: log2(blocks per limb); blockIdx.x = row << sub | block
  int prefetch_ahead;                         // > 0: L2-prefetch the source row of the CTA this many blocks ahead
  // sampler (PRE_TERNARY / PRE_CBD): stream = rng_stream(rng, domain, a0 + inst, b) (modarith.cuh); rnd, if set, holds the
  // stream's words already generated (k_rng_fill): word idx of instance inst at rnd[inst * rnd_is + idx]
  RngKey rng;
  u64 domain, a0, b;
  const u64 *rnd; long long rnd_is;
  // PRE_ENCODE / POST_DECODE
  const long long *slots_in; long long *slots_out; const u32 *index_map; int n_slots; long long slots_is;
  // PRE_PLAIN_LIFT
  u64 t, t_half_up;
  // PRE_GALOIS_REDUCE / POST_MODDOWN (key switching): automorphism out[j] = +-in[j * einv mod 2N] (einv = 0: none)
  u32 galois_einv;
  const DevConst *C;
  const u64 *tl; long long tl_is;               // accumulator block [2][k][N]; its rows (comp, L) hold INTT_p(acc_L)
  const u64 *base0, *base1; long long base0_is, base1_is;  // polynomial added into component 0 / 1 (nullptr = 0)
  u32 base_einv;                                // automorphism applied to base0/base1 while reading (0: none)
  int L, k;                                     // data limbs / key-level primes of the context
  int i0, nrows;                                // POST_MODDOWN rows: w = comp * nrows + (i - i0), limbs i0 .. i0+nrows-1
  u32 *flags; u32 flag_serial;                  // merged special-row INTT + ModDown launch: flags[inst][comp] == serial when ready
  int skew;                                     // ... special rows run this many instances ahead of their data rows
  // key-switch ModUp block T as the IMAGE of the swizzled shared-memory limb: the forward launch writes each row with one
  // bulk copy (no LDS + STG loop), the tail launch reads pair e2 at t[swz2(e2)] (a permutation inside 128-byte lines, so
  // the loads stay coalesced).  T is private to the key switch, so its element order is ours to choose.
  int t_image;
  // chained ModUp + tail launch (kschain.cu): done[inst][modulus] counts the ModUp rows stored so far (L per key switch)
  u32 *done; u32 done_target;
  u32 *fault;    // host-mapped word raised when a dependency wait gives up (wait_word)
  // ModUp-style rows whose source is already a residue of the row's own modulus (the BEHZ block's forward transforms):
  // no source prime to negate / compare against
  int persist;                   // the CTA processes several rows one after another (persistent chained grid)
  int src_same_mod;
  int raw_reduce;                // t_image rows: reduce the raw outputs to |x| <= 0.5 q before the bulk store (they feed products)
  // PRE_BEHZ_TENSOR: src = the transformed operand block [inst][4][bz_W][N] as raw-double images (polys a0, a1, b0, b1;
  // bz_square: b = a); output row w = poly * bz_W + r takes a0*b0 / a0*b1 + a1*b0 / a1*b1 of row r
  int bz_W, bz_square;
  // forward rows of the BEHZ block (src_same_mod): rows r < L of polynomial p are the operand ciphertexts' own limbs, read
  // where they lie (bz_a: polys 0,1; bz_b: polys 2,3; [inst][2][L][N]) instead of from a copy inside the block
  const u64 *bz_a, *bz_b;
  u32 *ticket; u32 ticket_base;  // grids with internal dependencies: logical block index = atomicAdd(ticket) - ticket_base
  u32 *t_used;   // [inst][modulus] tail rows that have consumed T[inst][modulus][*] (2 per key switch): the second one drops
                 // the rows from L2 (discard.global.L2) so that their dirty lines are never written to DRAM
};
// sm word index of element pair e2 = tid + i * T (T >= 128): the swizzle only looks at bits 1..7, which i * 2T never
// touches, so swz(2 * e2) = swz(2 * tid) + 2 * (e2 - tid) — the swizzle is computed once per thread, not per access
__device__ __forceinline__ int swz_pair(int tid, int e2) { return swz(2 * tid) + 2 * (e2 - tid); }
// physical 16-byte index of element pair e2 in the swizzled image of a limb (swz(2 * e2) / 2)
__device__ __forceinline__ int swz2(int e2) { return e2 ^ ((e2 >> 3) & 7) ^ ((e2 >> 4) & 4); }

// combos of (PRE, FWD, MUL, INV, POST) the library uses
enum {
  LIMB_FWD = 0,            // load, NTT
  LIMB_INV = 1,            // load, INTT
  LIMB_REDUCE_FWD = 2,     // load + reduce mod q (key-switch ModUp), NTT
  LIMB_PLAINLIFT_FWD = 3,  // centred lift of a mod-t plaintext, NTT
  LIMB_TERNARY_FWD = 4,    // sample R_3, NTT
  LIMB_CBD_FWD = 5,        // sample centred binomial noise, NTT
  LIMB_ENCODE_INV = 6,     // BatchEncoder scatter, INTT mod t
  LIMB_FWD_DECODE = 7,     // NTT mod t, BatchEncoder gather
  LIMB_FWD_MUL_INV = 8,    // NTT, * row, INTT                (multiply_plain)
  LIMB_FWD_MUL_INV_ADD = 9,// NTT, * row, INTT, + row         (decrypt: c1*s + c0)
  LIMB_MUL_INV = 10,       // * row, INTT                     (encrypt: pk * u)
  LIMB_GALOIS_REDUCE_FWD = 11,  // Galois gather + reduce mod q, NTT  (rotate: sigma(c1) ModUp)
  LIMB_INV_MODDOWN = 12,   // INTT, ModDown with rounding, + base (key-switch tail)
  LIMB_KSINNER_INV_MODDOWN = 13,  // key-switch inner product over J in the load, INTT, ModDown (AR_F64 only)
  LIMB_BEHZTENSOR_INV = 14,       // BEHZ tensor product of the transformed operand rows in the load, INTT (AR_F64 only)
  LIMB_NCOMBOS = 15
};

__device__ __forceinline__ u64 small_to_mod(int v, u64 q) { return v < 0 ? q - (u64)(-v) : (u64)v; }

// ---- pre-op: coefficients (2*e2, 2*e2+1) of source row `srow`, ready for the forward transform (< q)
// h = sampler stream key (sampling modes); n = coefficients per limb
template <int PRE>
__device__ __forceinline__ ulonglong2 limb_load_pair(const LimbJob &job, const ModInfo &M, const ModInfo *__restrict__ mods,
                                                     int n, int inst, int srow, int e2, const RngStream &rs) {
  const u64 q = M.q;
  ulonglong2 v;
  if (PRE == PRE_TERNARY || PRE == PRE_CBD) {
    u64 r0, r1;
    if (job.rnd) {
      const ulonglong2 w = reinterpret_cast<const ulonglong2 *>(job.rnd + (size_t)inst * job.rnd_is)[e2];
      r0 = w.x; r1 = w.y;
    } else {
      rng_word_pair(rs, (u64)e2, r0, r1);
    }
    const int a = (PRE == PRE_TERNARY) ? ternary_of(r0) : cbd_of(r0);
    const int b = (PRE == PRE_TERNARY) ? ternary_of(r1) : cbd_of(r1);
    v.x = small_to_mod(a, q); v.y = small_to_mod(b, q);
  } else if (PRE == PRE_GALOIS_REDUCE) {
    // GaloisTool::apply_galois as a gather (negation is modulo the SOURCE limb's prime), then the ModUp reduction
    const u64 *src = job.src + (size_t)inst * job.src_is + (size_t)srow * n;
    const u64 qs = mods[srow].q;
    const u32 einv = job.galois_einv, m2 = 2u * n - 1;
    const u32 r0 = ((u32)(2 * e2) * einv) & m2, r1 = (r0 + einv) & m2;
    v.x = src[r0 & (n - 1)]; v.y = src[r1 & (n - 1)];
    if (r0 >= (u32)n) v.x = neg_mod(v.x, qs);
    if (r1 >= (u32)n) v.y = neg_mod(v.y, qs);
    v.x = barrett64(v.x, q, M.mu_hi); v.y = barrett64(v.y, q, M.mu_hi);
  } else {
    v = reinterpret_cast<const ulonglong2 *>(job.src + (size_t)inst * job.src_is + (size_t)srow * n)[e2];
    if (PRE == PRE_REDUCE) { v.x = barrett64(v.x, q, M.mu_hi); v.y = barrett64(v.y, q, M.mu_hi); }
    if (PRE == PRE_PLAIN_LIFT) {
      // multiply_plain_normal: centred lift of a mod-t coefficient into [0,q)
      const u64 th = job.t_half_up, inc = q - job.t;
      v.x = v.x >= th ? v.x + inc : v.x;
      v.y = v.y >= th ? v.y + inc : v.y;
    }
  }
  return v;
}

// ---- post-op on canonical coefficients (2*e2, 2*e2+1) of output row w (destination row drow)
struct ModDownRow { u64 p, p_half, phm, ip, ips; const ulonglong2 *tl; const u64 *base; };
__device__ __forceinline__ ModDownRow moddown_row(const LimbJob &job, int n, int inst, int w) {
  // tail of switch_key_inplace for row (comp, i): dst = base + p^-1 * (acc_i - ([acc_L + p/2]_p mod q_i) + [p/2]_{q_i})
  const DevConst *C = job.C;
  const int comp = w >= job.nrows ? 1 : 0, i = job.i0 + (w - comp * job.nrows);  // w < 2 * nrows
  ModDownRow r;
  r.p = C->p; r.p_half = C->p_half; r.phm = C->p_half_mod_q[i]; r.ip = C->inv_p[i]; r.ips = C->inv_p_s[i];
  r.tl = reinterpret_cast<const ulonglong2 *>(job.tl + (size_t)inst * job.tl_is + (size_t)(comp * job.k + job.L) * n);
  r.base = comp == 0 ? job.base0 : job.base1;
  if (r.base) r.base += (size_t)inst * (comp == 0 ? job.base0_is : job.base1_is) + (size_t)i * n;
  return r;
}
// ModDown on the exact-double class: vd = INTT output as a centred double (|vd| < q), result canonical.
// Same value as the integer formulation below: base + p^-1 * (v - ([t + p/2]_p mod q) + [p/2]_q) mod q.
struct ModDownF64 { double pd, p_half, phm, ipd, ipc, qd, qinv; bool wide; };
__device__ __forceinline__ ModDownF64 moddown_f64(const ModDownRow &md, const ModInfo &M) {
  ModDownF64 f;
  f.pd = (double)md.p; f.p_half = (double)md.p_half; f.phm = (double)md.phm; f.ipd = (double)md.ip;
  f.qd = (double)M.q; f.qinv = f64_of(M.qinv_bits); f.ipc = f.ipd * f.qinv; f.wide = f64_wide(M.q);
  return f;
}
// the rounded division itself: p^-1 * (v - [t + p/2]_p + [p/2]) mod q as a centred double, |r| <= 0.6 q
__device__ __forceinline__ double moddown_core_f64(double vd, u64 t, const ModDownF64 &f) {
  double a = f64_of(ar_from_canon<AR_F64>(t)) + f.p_half;
  a = csub_ge(a, f.pd);                                           // [t + p/2]_p, exact
  const double d = (vd - reduce_f64(a, f.qinv, f.qd)) + f.phm;    // |d| < 2.6q, exact integer
  const double Q = rint_mul(d, f.ipc);
  const double ph = d * f.ipd, pl = fma(d, f.ipd, -ph);
  double r = fma(-Q, f.qd, ph) + pl;                              // d * p^-1 mod q, |r| <= 0.6q
  if (f.wide) r = reduce_f64(r, f.qinv, f.qd);                    // wide primes: the estimate's error can leave |r| ~ q
  return r;
}
__device__ __forceinline__ double moddown_one_f64(double vd, u64 t, bool has_base, u64 b, const ModDownF64 &f) {
  double r = moddown_core_f64(vd, t, f);
  if (has_base) r += f64_of(ar_from_canon<AR_F64>(b));
  r = cadd_neg(r, f.qd);
  return csub_ge(r, f.qd);                                        // canonical, as a double
}
// The additions that follow the division (+ base, + addend, canonical ranges) on the INTEGER pipes, which the key switch
// leaves idle: one DADD turns the centred double into a two's-complement integer (the 1.5 * 2^52 encoding), the rest is
// 64-bit adds and compares instead of 7 .. 11 FP64-pipe instructions per coefficient.  Measured (tools/ab_ks.sh, rotate at
// N = 8192, B = 592): 1.0152 vs 1.0167 ms — nothing: the key switch is bound by latency at 8 warps per scheduler, not by
// FP64 issue, so trading FP64 instructions for integer ones does not show.  Kept as -DABC_MD_INT=1 (the half-limb rows of
// ks14.cu use it); the whole-limb epilogue stays on the FP64 form.
#ifndef ABC_MD_INT
#define ABC_MD_INT 0
#endif
__device__ __forceinline__ u64 moddown_finish_int(double r, bool has_base, u64 b, u64 q) {
  long long x = (long long)(bits_of(r + ABC_RINT_MAGIC) - 0x4338000000000000ULL);   // |r| <= 0.6 q < 2^51
  if (has_base) x += (long long)b;                                                  // (-0.6 q, 1.6 q)
  if (x < 0) x += (long long)q;
  if (x >= (long long)q) x -= (long long)q;
  return (u64)x;
}
__device__ __forceinline__ u64 f64_canon_bits(double r) { return bits_of(r + 4503599627370496.0) & 0x000FFFFFFFFFFFFFULL; }
__device__ __forceinline__ double add_canon_f64(double r, u64 ad, double qd) {
  r += f64_of(ar_from_canon<AR_F64>(ad));
  return csub_ge(r, qd);
}

// ---- ModDown epilogue of one data row on the exact-double class: sm holds the INTT output as centred doubles
// (swizzled); out = base (through the automorphism einv) + p^-1 * (v - [t + p/2]_p + [p/2]) mod q, + addp if given
// (out2 then receives the result without it)
template <int LOGN, int T>
__device__ __forceinline__ void moddown_store_f64(const u64 *sm, const ModInfo &M, const ModDownRow &md, u32 einv,
                                                  ulonglong2 *__restrict__ out, ulonglong2 *__restrict__ out2,
                                                  const ulonglong2 *__restrict__ addp, int tid) {
  constexpr int N = 1 << LOGN, IT = N / 2 / T, CH = IT < 4 ? IT : 4;
  static_assert(IT % CH == 0, "whole chunks");
  const u64 q = M.q;
  const ModDownF64 f = moddown_f64(md, M);
  const u32 m2 = 2u * N - 1;
#pragma unroll 1
  for (int i0 = 0; i0 < IT; i0 += CH) {
    // everything this chunk reads from global memory is requested before the first dependent instruction
    ulonglong2 t[CH], ad[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int e2 = tid + (i0 + i) * T;
      t[i] = __ldcg(md.tl + e2);
      if (addp) ad[i] = addp[e2];
    }
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int e2 = tid + (i0 + i) * T;
      const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(&sm[swz_pair(tid, e2)]);
      ulonglong2 b = make_ulonglong2(0, 0);
      if (md.base) {
        if (einv) {
          const u32 r0 = ((u32)(2 * e2) * einv) & m2, r1 = (r0 + einv) & m2;
          b.x = md.base[r0 & (N - 1)]; b.y = md.base[r1 & (N - 1)];
          if (r0 >= (u32)N) b.x = neg_mod(b.x, q);
          if (r1 >= (u32)N) b.y = neg_mod(b.y, q);
        } else {
          b = reinterpret_cast<const ulonglong2 *>(md.base)[e2];
        }
      }
#if ABC_MD_INT
      u64 rx = moddown_finish_int(moddown_core_f64(f64_of(v.x), t[i].x, f), md.base != nullptr, b.x, q);
      u64 ry = moddown_finish_int(moddown_core_f64(f64_of(v.y), t[i].y, f), md.base != nullptr, b.y, q);
      if (addp) {
        if (out2) out2[e2] = make_ulonglong2(rx, ry);
        rx = add_mod(rx, ad[i].x, q); ry = add_mod(ry, ad[i].y, q);
      }
      out[e2] = make_ulonglong2(rx, ry);
#else
      double rx = moddown_one_f64(f64_of(v.x), t[i].x, md.base != nullptr, b.x, f);
      double ry = moddown_one_f64(f64_of(v.y), t[i].y, md.base != nullptr, b.y, f);
      if (addp) {
        if (out2) out2[e2] = make_ulonglong2(f64_canon_bits(rx), f64_canon_bits(ry));
        rx = add_canon_f64(rx, ad[i].x, f.qd); ry = add_canon_f64(ry, ad[i].y, f.qd);
      }
      out[e2] = make_ulonglong2(f64_canon_bits(rx), f64_canon_bits(ry));
#endif
    }
  }
}

template <int POST>
__device__ __forceinline__ void limb_store_pair(const LimbJob &job, const ModInfo &M, const ModDownRow &md, int n, int inst,
                                                int drow, int arow, int e2, ulonglong2 v) {
  const u64 q = M.q;
  if (POST == POST_MODDOWN) {
    const ulonglong2 t = __ldcg(md.tl + e2);  // written by another CTA of this launch in the merged mode: L2, not L1
    const u64 rx = sub_mod(barrett64(add_mod(t.x, md.p_half, md.p), q, M.mu_hi), md.phm, q);
    const u64 ry = sub_mod(barrett64(add_mod(t.y, md.p_half, md.p), q, M.mu_hi), md.phm, q);
    v.x = mul_shoup(sub_mod(v.x, rx, q), md.ip, md.ips, q);
    v.y = mul_shoup(sub_mod(v.y, ry, q), md.ip, md.ips, q);
    if (md.base) {
      ulonglong2 b;
      const u32 einv = job.base_einv;
      if (einv) {
        const u32 m2 = 2u * n - 1, r0 = ((u32)(2 * e2) * einv) & m2, r1 = (r0 + einv) & m2;
        b.x = md.base[r0 & (n - 1)]; b.y = md.base[r1 & (n - 1)];
        if (r0 >= (u32)n) b.x = neg_mod(b.x, q);
        if (r1 >= (u32)n) b.y = neg_mod(b.y, q);
      } else {
        b = reinterpret_cast<const ulonglong2 *>(md.base)[e2];
      }
      v.x = add_mod(v.x, b.x, q); v.y = add_mod(v.y, b.y, q);
    }
    if (job.add) {  // rotate + add: a whole ciphertext accumulated into the result
      if (job.dst2) reinterpret_cast<ulonglong2 *>(job.dst2 + (size_t)inst * job.dst_is + (size_t)drow * n)[e2] = v;
      const ulonglong2 a = reinterpret_cast<const ulonglong2 *>(job.add + (size_t)inst * job.add_is + (size_t)drow * n)[e2];
      v.x = add_mod(v.x, a.x, q); v.y = add_mod(v.y, a.y, q);
    }
  } else if (POST == POST_ADD) {
    const ulonglong2 a = reinterpret_cast<const ulonglong2 *>(job.add + (size_t)inst * job.add_is + (size_t)arow * n)[e2];
    v.x = add_mod(v.x, a.x, q); v.y = add_mod(v.y, a.y, q);
  }
  reinterpret_cast<ulonglong2 *>(job.dst + (size_t)inst * job.dst_is + (size_t)drow * n)[e2] = v;
}

// ---- bulk asynchronous copy of one row into shared memory (cp.async.bulk, completion on an mbarrier): one thread
// issues it, no registers are staged, and the whole row is in flight from the first cycle of the CTA
// reuse: the CTA will bring in another row through the same barrier object later (persistent grids): invalidate it once every
// thread has seen the copy complete
__device__ __forceinline__ void bulk_row_to_smem(u64 *sm, const u64 *src, u32 bytes, u64 *mbar, int tid, bool reuse = false) {
  const u32 mb = (u32)__cvta_generic_to_shared(mbar), dst = (u32)__cvta_generic_to_shared(sm);
#if ABC_WHATIF & 2
  (void)mb; (void)dst; (void)src; (void)bytes; __syncthreads(); return;   // what-if: the source row costs nothing
#endif
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mb) : "memory");
  }
  __syncthreads();  // the barrier object is initialised before anyone polls it
  u32 ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(mb) : "memory");
  } while (!ok);
  if (reuse) {
    __syncthreads();
    if (tid == 0) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(mb) : "memory");
  }
}

// ---- key-switch inner product in the load of the inverse transform (AR_F64: every prime < 2^45):
// coefficients (2*e2, 2*e2+1) of  sum_J T[inst][I][J] * key[J][comp][I]  mod q_I, as centred doubles.
// The low 64 bits of the sum come from the integer pipe (3 IMAD per term, no carries), the quotient from the FP64
// pipe: Q = rint(fl(sum t*k) / q) is within 1 of the true quotient, so r = lo64(S) - lo64(Q*q) is S - Q*q exactly.
__device__ __forceinline__ double ks_dot_f64(u64 lo, double s, double qinv, u64 q, double qd) {
  const double t = fma(s, qinv, 4503599627370496.0);                       // 2^52 + rint(s/q), s >= 0
  const u64 Q = bits_of(t) & 0x000FFFFFFFFFFFFFULL;
  const u64 r = mad_lo64(Q, 0 - q, lo);                                    // signed, |r| < 2q
  const double rd = __hiloint2double((int)((u32)(r >> 32) + 0x43380000u), (int)(u32)r) - 6755399441055744.0;
  return reduce_f64(rd, qinv, qd);                                         // |.| <= 0.51q
}
__device__ __forceinline__ ulonglong2 ks_inner_pair(const ulonglong2 *__restrict__ t, const ulonglong2 *__restrict__ kp,
                                                    int L, int rowv, int keyv2, const ModInfo &M) {
  u64 lo0 = 0, lo1 = 0;
  double s0 = 0.0, s1 = 0.0;
#pragma unroll 4
  for (int J = 0; J < L; ++J) {
    const ulonglong2 tv = __ldcg(t + (size_t)J * rowv);  // streamed once, possibly written by this very launch: L2
    const ulonglong2 kv = __ldg(kp + (size_t)J * keyv2);
    lo0 = mad_lo64(tv.x, kv.x, lo0); lo1 = mad_lo64(tv.y, kv.y, lo1);
    s0 = fma(f64_of(ar_from_canon<AR_F64>(tv.x)), f64_of(ar_from_canon<AR_F64>(kv.x)), s0);
    s1 = fma(f64_of(ar_from_canon<AR_F64>(tv.y)), f64_of(ar_from_canon<AR_F64>(kv.y)), s1);
  }
  const double qinv = f64_of(M.qinv_bits), qd = (double)M.q;
  ulonglong2 v;
  v.x = bits_of(ks_dot_f64(lo0, s0, qinv, M.q, qd));
  v.y = bits_of(ks_dot_f64(lo1, s1, qinv, M.q, qd));
  return v;
}

// The same inner product when T is the raw-double image the ModUp launch leaves (t_image) and the key is its exact-double
// copy: every term is reduced on the FP64 pipe (mul_tw, |term| <= 0.6q), no conversions and no integer multiplies.
__device__ __forceinline__ ulonglong2 ks_inner_pair_f64(const double2 *__restrict__ t, const double2 *__restrict__ kp, int L,
                                                        int rowv, int keyv2, const ModInfo &M) {
  const double qinv = f64_of(M.qinv_bits), qd = (double)M.q;
  const u64 qb = bits_of(qd);
  double s0 = 0.0, s1 = 0.0;
#pragma unroll 4
  for (int J = 0; J < L; ++J) {
    const double2 tv = __ldcg(t + (size_t)J * rowv);
    const double2 kv = __ldg(kp + (size_t)J * keyv2);
    s0 += f64_of(mul_tw<AR_F64>(bits_of(tv.x), bits_of(kv.x), bits_of(qinv), M.q, qb));
    s1 += f64_of(mul_tw<AR_F64>(bits_of(tv.y), bits_of(kv.y), bits_of(qinv), M.q, qb));
  }
  return make_ulonglong2(bits_of(reduce_f64(s0, qinv, qd)), bits_of(reduce_f64(s1, qinv, qd)));
}

// All pairs of a thread (e2 = tid + i * T), software-pipelined: the T values of pair i + 1 are requested before pair i is
// multiplied, so the load phase of a tail row is not one L2 round trip per pair (ncu: this phase was 20 % of the chained
// grid's stall samples, on the first use of the loaded values).  LT = number of data limbs, compile time.
#ifndef ABC_KS_PIPE_INNER
#define ABC_KS_PIPE_INNER 1
#endif
#ifndef ABC_KS_KEY_CG
#define ABC_KS_KEY_CG 1
#endif
// ACC: add to what shared memory already holds (the sum over an earlier chunk of LT limbs, |x| <= 0.51 q): L = 8 runs as
// two chunks of 4, which keeps the software pipeline inside the register budget
template <int LT, int NIT, int T, bool ACC = false>
__device__ __forceinline__ void ks_inner_rows_f64(u64 *sm, const double2 *__restrict__ t, const double2 *__restrict__ kp,
                                                  int rowv, int keyv2, const ModInfo &M, int tid) {
  const double qinv = f64_of(M.qinv_bits), qd = (double)M.q;
  const u64 qb = bits_of(qd), qib = bits_of(qinv);
  const int p0 = swz(2 * tid);
  const double2 *tp = t + swz2(tid);   // swz2(tid + i * T) = swz2(tid) + i * T: the swizzle looks at bits 0..6 only
  double2 tv[LT], tn[LT];
#pragma unroll
  for (int J = 0; J < LT; ++J) tv[J] = __ldcg(tp + (size_t)J * rowv);
#pragma unroll
  for (int i = 0; i < NIT; ++i) {
    if (i + 1 < NIT) {
#pragma unroll
      for (int J = 0; J < LT; ++J) tn[J] = __ldcg(tp + (i + 1) * T + (size_t)J * rowv);
    }
    double s0 = 0.0, s1 = 0.0;
    if (ACC) {
      const double2 prev = *reinterpret_cast<const double2 *>(&sm[p0 + 2 * i * T]);
      s0 = prev.x; s1 = prev.y;
    }
#pragma unroll
    for (int J = 0; J < LT; ++J) {
#if ABC_KS_KEY_CG
      const double2 kv = __ldcg(kp + tid + i * T + (size_t)J * keyv2);   // L2 only: the key rows stream through, L1 keeps the twiddles
#else
      const double2 kv = __ldg(kp + tid + i * T + (size_t)J * keyv2);
#endif
      s0 += f64_of(mul_tw<AR_F64>(bits_of(tv[J].x), bits_of(kv.x), qib, M.q, qb));
      s1 += f64_of(mul_tw<AR_F64>(bits_of(tv[J].y), bits_of(kv.y), qib, M.q, qb));
    }
    *reinterpret_cast<ulonglong2 *>(&sm[p0 + 2 * i * T]) =
        make_ulonglong2(bits_of(reduce_f64(s0, qinv, qd)), bits_of(reduce_f64(s1, qinv, qd)));
#pragma unroll
    for (int J = 0; J < LT; ++J) tv[J] = tn[J];
  }
}

// ---- one limb (or, TAIL, one 2^LOGN-coefficient block of a larger limb) in shared memory: row w of instance inst
#ifndef ABC_KS_DISCARD_T
#define ABC_KS_DISCARD_T 0
#endif
// second consumer of T[inst][I][0..L): drop the rows from L2 (warp 0; `before` = the consumption count thread 0 took)
template <int N>
__device__ __forceinline__ void discard_consumed_T(const LimbJob &job, int inst, int srow, u32 before, int tid) {
  if (job.t_used && tid < 32) {
    const int I = srow >= job.k ? srow - job.k : srow;
    if (__shfl_sync(0xffffffffu, before, 0) & 1u) {
      const char *p = reinterpret_cast<const char *>(job.src + (size_t)inst * job.src_is + (size_t)I * job.L * N);
      for (int line = tid; line < job.L * (N * 8 / 128); line += 32)
        asm volatile("discard.global.L2 [%0], 128;" ::"l"(p + (size_t)line * 128) : "memory");
    }
  }
}

// Bounded wait for a word another CTA of the same grid sets (a producer dispatched earlier: normally it is set already).
// After about a second it gives up, raises the context's fault word (host-mapped) and lets the row finish with whatever it
// reads: a wrong result that the host reports at its next synchronisation point instead of a hung GPU.  GE: wait until
// (int)(seen - target) >= 0, else until seen == target.
template <bool GE> __device__ __forceinline__ void wait_word(const u32 *p, u32 target, u32 *fault) {
  u32 seen;
  for (u32 spins = 0;; ++spins) {
    asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(seen) : "l"(p) : "memory");
    if (GE ? (int)(seen - target) >= 0 : seen == target) return;
    if (spins > (1u << 22)) { if (fault) *reinterpret_cast<volatile u32 *>(fault) = 1u; return; }
    __nanosleep(64);
  }
}

// Logical block index of a grid whose rows wait for rows "earlier" in the grid.  CUDA does not promise that blocks start in
// blockIdx order, so "earlier" is defined by a ticket taken when the block starts running (as in decoupled look-back
// scans): a block only ever waits for blocks that hold smaller tickets, i.e. that are already running or done, so the
// waits cannot deadlock whatever the dispatch order.  base = tickets handed out by earlier launches on this counter.
__device__ __forceinline__ unsigned grid_ticket(u32 *counter, u32 base) {
  __shared__ unsigned s_ticket;
  if (threadIdx.x == 0) s_ticket = atomicAdd(counter, 1u) - base;
  __syncthreads();
  return s_ticket;
}

// RowIds: the caller already knows modulus / destination row / source row of row w (the chained key switch packs them
// into its schedule: no dependent loads from the row maps in front of the first copy); modidx < 0 = look them up
struct RowIds { int modidx, drow, srow; };
template <int LOGN, int PRE, bool FWD, bool MUL, bool INV, int POST, int AR, bool TAIL>
__device__ __forceinline__ void limb_body(const LimbJob &job, const ModInfo *__restrict__ mods, int inst, int w,
                                          RowIds ids = RowIds{-1, 0, 0}) {
  typedef NttDims<LOGN> D;
  extern __shared__ __align__(16) u64 sm[];
  const int tid = threadIdx.x;
  const int blk = TAIL ? (int)(blockIdx.x & ((1u << job.sub) - 1)) : 0;
  const u32 twbase = TAIL ? ((1u << job.sub) + (u32)blk) : 1u;
  const int n = TAIL ? job.n : D::N;           // coefficients per limb
  const int eoff = blk * (D::N / 2);           // this block's offset in 16-byte units
  const int modidx = ids.modidx >= 0 ? ids.modidx : job.rowmod[w];
  const ModInfo M = mods[modidx];
  const u64 q = M.q;
  const int drow = ids.modidx >= 0 ? ids.drow : (job.rowdst ? job.rowdst[w] : w);
  const int srow = ids.modidx >= 0 ? ids.srow : (job.rowsrc ? job.rowsrc[w] : drow);
  const int mrow = job.rowmul ? job.rowmul[w] : w;

  // ---- L2 prefetch of the row a later CTA of this launch will load (ncu: 35 % of this kernel's stall samples sat on
  // the first shared-memory store, i.e. on the HBM latency of the row load; the prefetch turns it into L2 latency)
  if (!TAIL && (PRE == PRE_LOAD || PRE == PRE_REDUCE || PRE == PRE_PLAIN_LIFT || PRE == PRE_GALOIS_REDUCE) && job.prefetch_ahead > 0) {
    const unsigned nb = gridDim.x * gridDim.y, lb = blockIdx.y * gridDim.x + blockIdx.x + (unsigned)job.prefetch_ahead;
    if (lb < nb) {
      const int w2 = (int)(lb % gridDim.x), inst2 = (int)(lb / gridDim.x);
      const int d2 = job.rowdst ? job.rowdst[w2] : w2, s2 = job.rowsrc ? job.rowsrc[w2] : d2;
      const char *p = reinterpret_cast<const char *>(job.src + (size_t)inst2 * job.src_is + (size_t)s2 * D::N);
      for (int line = tid; line < (int)(D::N * 8 / 128); line += D::T)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p + (size_t)line * 128));
    }
  }

  // AR_F64 ModUp: bulk-copy the source row as it lies, gather / negate / convert in the first pass, no Barrett
  constexpr bool LINSRC = AR == AR_F64 && FWD && !TAIL && (PRE == PRE_REDUCE || PRE == PRE_GALOIS_REDUCE);

  // ---- load
  if constexpr (LINSRC) {
    __shared__ __align__(8) u64 mbar;
    const u64 *sp = job.src + (size_t)inst * job.src_is + (size_t)srow * D::N;
    if (job.bz_a) {
      const int p = w / job.bz_W, r = w - p * job.bz_W;
      if (r < job.L) sp = (p < 2 ? job.bz_a : job.bz_b) + (((size_t)inst * 2 + (p & 1)) * job.L + r) * D::N;
    }
    bulk_row_to_smem(sm, sp, (u32)D::SMEM, &mbar, tid, job.persist != 0);
  } else if (PRE == PRE_KS_INNER) {
    // srow = comp * k + I: row of the accumulator block this CTA produces
    const int comp = srow >= job.k ? 1 : 0, I = srow - comp * job.k;  // srow < 2k
    const ulonglong2 *t = reinterpret_cast<const ulonglong2 *>(job.src + (size_t)inst * job.src_is + (size_t)I * job.L * D::N);
    const ulonglong2 *kp = reinterpret_cast<const ulonglong2 *>(job.mul + (size_t)(comp * job.k + I) * D::N);
    if (job.done) {  // chained launch: T[inst][I][0..L) comes from blocks earlier in this grid
      if (tid == 0) wait_word<true>(job.done + inst * job.k + I, job.done_target, job.fault);
      __syncthreads();
    }
    if (job.prefetch_ahead >= 0 && !job.done) {
      // the L T rows are contiguous and come straight from DRAM (written by the ModUp launch, larger than L2 at
      // bench batch sizes): pull them into L2 now so only the first loads of the loop below pay DRAM latency; same
      // for the rows the ModDown epilogue reads tens of microseconds from now (addend, sigma(c0))
      const char *p = reinterpret_cast<const char *>(t);
      for (int line = tid; line < job.L * (int)(D::N * 8 / 128); line += D::T)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p + (size_t)line * 128));
      if (POST == POST_MODDOWN && w >= 2) {
        const int wq0 = w - 2, comp0 = wq0 >= job.nrows ? 1 : 0, i0 = job.i0 + (wq0 - comp0 * job.nrows);
        if (job.add) {
          const char *pa = reinterpret_cast<const char *>(job.add + (size_t)inst * job.add_is + (size_t)drow * D::N);
          for (int line = tid; line < (int)(D::N * 8 / 128); line += D::T)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pa + (size_t)line * 128));
        }
        if (comp0 == 0 && job.base0) {
          const char *pb = reinterpret_cast<const char *>(job.base0 + (size_t)inst * job.base0_is + (size_t)i0 * D::N);
          for (int line = tid; line < (int)(D::N * 8 / 128); line += D::T)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pb + (size_t)line * 128));
        }
      }
    }
#if ABC_WHATIF & 4
    if (true) {  // what-if: the inner product costs nothing
      for (int e2 = tid; e2 < D::N / 2; e2 += D::T)
        *reinterpret_cast<ulonglong2 *>(&sm[swz_pair(tid, e2)]) = make_ulonglong2(bits_of((double)e2), bits_of((double)tid));
    } else
#endif
    if (job.t_image) tw_prefetch_last<LOGN, AR>(M.itwd, twbase, tid);   // the inverse transform starts with the contiguous pass
    if (job.t_image && ABC_KS_PIPE_INNER && job.L == 4) {
      ks_inner_rows_f64<4, D::N / 2 / D::T, D::T>(sm, reinterpret_cast<const double2 *>(t), reinterpret_cast<const double2 *>(kp),
                                                 D::N / 2, job.k * D::N, M, tid);
    } else if (job.t_image) {
      for (int e2 = tid; e2 < D::N / 2; e2 += D::T)
        *reinterpret_cast<ulonglong2 *>(&sm[swz_pair(tid, e2)]) =
            ks_inner_pair_f64(reinterpret_cast<const double2 *>(t) + swz2(e2), reinterpret_cast<const double2 *>(kp) + e2, job.L,
                              D::N / 2, job.k * D::N, M);
    } else {
      for (int e2 = tid; e2 < D::N / 2; e2 += D::T)
        *reinterpret_cast<ulonglong2 *>(&sm[swz_pair(tid, e2)]) = ks_inner_pair(t + e2, kp + e2, job.L, D::N / 2, job.k * D::N, M);
    }
  } else if (PRE == PRE_BEHZ_TENSOR) {
    // Evaluator::bfv_multiply's dyadic products (behz_ciphertext_product) in the load of the inverse transform: the
    // operand rows are the raw-double images their forward transforms left (|x| <= 0.5 q, same element order as this
    // CTA's shared memory, so image pair j goes straight to shared-memory pair j); every product is six FP64 instructions
    const int Wb = job.bz_W, p = w / Wb, r = w - p * Wb;
    const double2 *a0 = reinterpret_cast<const double2 *>(job.src + (size_t)inst * job.src_is + (size_t)r * D::N);
    const size_t ps = (size_t)Wb * (D::N / 2);
    const double2 *a1 = a0 + ps, *b0 = job.bz_square ? a0 : a0 + 2 * ps, *b1 = job.bz_square ? a1 : a0 + 3 * ps;
    const double qinv = f64_of(M.qinv_bits), qd = (double)q;
    const u64 qb = bits_of(qd);
    auto mul = [&](double x, double y) { return f64_of(mul_tw<AR_F64>(bits_of(x), bits_of(y), M.qinv_bits, q, qb)); };
    const double2 *x = p == 2 ? a1 : a0, *y = p == 0 ? b0 : b1;
    for (int j = tid; j < D::N / 2; j += D::T) {
      const double2 xv = __ldcs(x + j), yv = __ldcs(y + j);
      double2 o = make_double2(mul(xv.x, yv.x), mul(xv.y, yv.y));
      if (p == 1) {   // a0*b1 + a1*b0 (squaring: 2 * a0*a1)
        if (job.bz_square) { o.x += o.x; o.y += o.y; }
        else {
          const double2 uv = __ldcs(a1 + j), vv = __ldcs(b0 + j);
          o.x += mul(uv.x, vv.x); o.y += mul(uv.y, vv.y);
        }
        o.x = reduce_f64(o.x, qinv, qd); o.y = reduce_f64(o.y, qinv, qd);
      }
      *reinterpret_cast<double2 *>(&sm[2 * j]) = o;
    }
  } else if (PRE == PRE_ENCODE) {
    // BatchEncoder::encode (SealCiphertextFactory.cpp:102-132): pad with the last value, scatter by the index map
    const long long *sl = job.slots_in + (size_t)inst * job.slots_is;
    for (int e = tid; e < D::N; e += D::T) {
      long long v = sl[e < job.n_slots ? e : job.n_slots - 1];
      sm[swz((int)job.index_map[e])] = v < 0 ? q + (u64)v : (u64)v;
    }
  } else {
    RngStream rs;
    if (PRE == PRE_TERNARY || PRE == PRE_CBD) rs = rng_stream(job.rng, job.domain, job.a0 + (u64)inst, job.b);
    for (int e2 = tid; e2 < D::N / 2; e2 += D::T)
      *reinterpret_cast<ulonglong2 *>(&sm[swz_pair(tid, e2)]) = limb_load_pair<PRE>(job, M, mods, n, inst, srow, eoff + e2, rs);
  }
  if (!LINSRC) __syncthreads();
  // T[inst][I][0..L) is dead once both components' rows have read it (the barrier above: every thread of this row has).
  // The count is taken now and looked at after the row is stored (its latency hides behind the transform); the row that
  // finds it odd drops T's lines from L2 so that their dirty data never goes to DRAM.
  // (compiled in with -DABC_KS_DISCARD_T=1 only: DRAM traffic of a key switch 1.55 -> 0.77 GB at B = 592, but the extra
  // code costs the tail rows 2 % and DRAM is at 17 % of its peak either way)
  u32 t_used_before = 0;
#if ABC_KS_DISCARD_T
  if constexpr (PRE == PRE_KS_INNER) {
    if (job.t_used && tid == 0) t_used_before = atomicAdd(job.t_used + inst * job.k + (srow >= job.k ? srow - job.k : srow), 1u);
  }
#endif

  if (FWD) {
    if constexpr (LINSRC) {
      if (job.t_image) {  // (uniform) the row leaves as the image of the swizzled limb: one bulk copy, issued by one thread
        ntt_fwd_smem_mids<LOGN, AR, true>(sm, M, twbase, tid, job.src_same_mod ? q : mods[srow].q,
                                          PRE == PRE_GALOIS_REDUCE ? job.galois_einv : 0u);
        ntt_fwd_last<LOGN, AR, 0, true>(sm, M, twbase, q, ar_aux<AR>(q), tid, job.raw_reduce != 0);  // raw doubles: the consumer multiplies them as they are
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the bulk copy
        __syncthreads();
        if (tid == 0) {
          u64 *g = job.dst + (size_t)inst * job.dst_is + (size_t)drow * D::N;
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                       ::"l"(g), "r"((u32)__cvta_generic_to_shared(sm)), "r"((u32)D::SMEM) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          if (job.done) {  // chained launch: the tail rows of this instance wait for the L rows of their modulus
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            __threadfence();
            atomicAdd(job.done + inst * job.k + modidx, 1u);
          } else {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory is read out before the CTA retires
          }
        }
        return;
      }
      ntt_fwd_smem<LOGN, AR, true>(sm, M, twbase, tid, mods[srow].q, PRE == PRE_GALOIS_REDUCE ? job.galois_einv : 0u);
    } else {
      ntt_fwd_smem<LOGN, AR>(sm, M, twbase, tid);
    }
  }

  if (MUL) {
    const ulonglong2 *mp = reinterpret_cast<const ulonglong2 *>(job.mul + (size_t)inst * job.mul_is + (size_t)mrow * n) + eoff;
    for (int e2 = tid; e2 < D::N / 2; e2 += D::T) {
      ulonglong2 m = mp[e2];
      ulonglong2 *p = reinterpret_cast<ulonglong2 *>(&sm[swz_pair(tid, e2)]);
      ulonglong2 v = *p;
      if constexpr (AR == AR_F64) {  // six FP64 instructions per product instead of a 128-bit Barrett reduction
        const double qd = (double)q;
        const u64 qb = bits_of(qd);
        v.x = f64_to_canon(f64_of(mul_tw<AR_F64>(ar_from_canon<AR_F64>(v.x), ar_from_canon<AR_F64>(m.x), M.qinv_bits, q, qb)), qd);
        v.y = f64_to_canon(f64_of(mul_tw<AR_F64>(ar_from_canon<AR_F64>(v.y), ar_from_canon<AR_F64>(m.y), M.qinv_bits, q, qb)), qd);
      } else {
        v.x = mul_mod(v.x, m.x, q, M.mu_hi, M.mu_lo);
        v.y = mul_mod(v.y, m.y, q, M.mu_hi, M.mu_lo);
      }
      *p = v;
    }
    __syncthreads();
  }

  if (INV) ntt_inv_smem<LOGN, !TAIL, AR, PRE == PRE_KS_INNER || PRE == PRE_BEHZ_TENSOR>(sm, M, twbase, tid);

  // ---- store
  if (POST == POST_DECODE) {
    // BatchEncoder::decode (SealCiphertextFactory.cpp:151): gather by the index map, centre to signed
    long long *out = job.slots_out + (size_t)inst * D::N;
    const u64 half = q >> 1;
    for (int e = tid; e < D::N; e += D::T) {
      u64 v = sm[swz((int)job.index_map[e])];
      out[e] = v > half ? (long long)v - (long long)q : (long long)v;
    }
  } else {
    ModDownRow md;
    int wq = w;
    if (POST == POST_MODDOWN && job.flags) {
      // merged launch: rows 0,1 are the special-prime rows of components 0,1; every data row needs its component's
      // INTT_p(acc_L) for the ModDown.  Rows are dispatched in blockIdx order (the two special rows of an instance
      // first), so a data row only ever waits for CTAs that are already resident; it waits AFTER its own INTT.
      if (w < 2) {
        ulonglong2 *o = reinterpret_cast<ulonglong2 *>(const_cast<u64 *>(job.tl) + (size_t)inst * job.tl_is +
                                                       (size_t)(w * job.k + job.L) * n);
        for (int e2 = tid; e2 < D::N / 2; e2 += D::T) {
          ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(&sm[swz_pair(tid, e2)]);
          v.x = canon_inv<AR>(v.x, q, ar_aux<AR>(q)); v.y = canon_inv<AR>(v.y, q, ar_aux<AR>(q));
          o[e2] = v;
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) atomicExch(job.flags + inst * 2 + w, job.flag_serial);
#if ABC_KS_DISCARD_T
        if constexpr (PRE == PRE_KS_INNER) discard_consumed_T<D::N>(job, inst, srow, t_used_before, tid);
#endif
        return;
      }
      wq = w - 2;
      if (tid == 0) wait_word<false>(job.flags + inst * 2 + (wq >= job.nrows ? 1 : 0), job.flag_serial, job.fault);
      __syncthreads();
    }
    if (POST == POST_MODDOWN) md = moddown_row(job, n, inst, wq);
#if ABC_WHATIF & 8
    if (POST == POST_MODDOWN && PRE == PRE_KS_INNER) {  // what-if: the ModDown epilogue costs nothing (one store per thread keeps the row alive)
      job.dst[(size_t)inst * job.dst_is + (size_t)drow * D::N + tid] = sm[swz(tid)] ^ sm[swz(tid + 4096)];
      return;
    }
#endif
    if constexpr (POST == POST_MODDOWN && AR == AR_F64 && !TAIL) {
      moddown_store_f64<LOGN, D::T>(
          sm, M, md, job.base_einv, reinterpret_cast<ulonglong2 *>(job.dst + (size_t)inst * job.dst_is + (size_t)drow * D::N),
          job.dst2 ? reinterpret_cast<ulonglong2 *>(job.dst2 + (size_t)inst * job.dst_is + (size_t)drow * D::N) : nullptr,
          job.add ? reinterpret_cast<const ulonglong2 *>(job.add + (size_t)inst * job.add_is + (size_t)drow * D::N) : nullptr, tid);
    } else {
      for (int e2 = tid; e2 < D::N / 2; e2 += D::T) {
        ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(&sm[swz_pair(tid, e2)]);
        if (INV && !TAIL) { v.x = canon_inv<AR>(v.x, q, ar_aux<AR>(q)); v.y = canon_inv<AR>(v.y, q, ar_aux<AR>(q)); }  // a tail block stays in [0,2q) for the head pass
        limb_store_pair<POST>(job, M, md, n, inst, drow, mrow, eoff + e2, v);
      }
    }
  }
#if ABC_KS_DISCARD_T
  if constexpr (PRE == PRE_KS_INNER) discard_consumed_T<D::N>(job, inst, srow, t_used_before, tid);
#endif
  (void)t_used_before;
}

template <int LOGN, int PRE, bool FWD, bool MUL, bool INV, int POST, int AR, bool TAIL>
__global__ void __launch_bounds__(NttDims<LOGN>::T, NttDims<LOGN>::MINB) k_limb(LimbJob job,
                                                                               const ModInfo *__restrict__ mods) {
  int inst = blockIdx.y, w = TAIL ? (int)(blockIdx.x >> job.sub) : (int)blockIdx.x;
  if (POST == POST_MODDOWN && !TAIL && job.flags) {
    // merged special-row + ModDown launch: in ticket order the two special-prime rows of instance g + S come with the
    // data rows of instance g, so they have published INTT_p(acc_L) long before their own data rows ask for it (a data
    // row only ever waits for blocks with a smaller ticket: no deadlock)
    const int W = gridDim.x, Bn = gridDim.y, nd = W - 2, S = job.skew < Bn ? job.skew : Bn;
    const int b = (int)grid_ticket(job.ticket, job.ticket_base);
    if (b < 2 * S) { inst = b >> 1; w = b & 1; }
    else {
      const int b1 = b - 2 * S, full = (Bn - S) * W;
      if (b1 < full) { const int g = b1 / W, r = b1 - g * W; inst = r < 2 ? g + S : g; w = r; }
      else { const int b2 = b1 - full, g = b2 / nd; inst = Bn - S + g; w = 2 + (b2 - g * nd); }
    }
  }
  limb_body<LOGN, PRE, FWD, MUL, INV, POST, AR, TAIL>(job, mods, inst, w);
}

// ---- head passes of the two-pass transform (N >= 32768): 2^A coefficients per thread at stride N >> A,
// two adjacent columns per thread so every access is 16 bytes.  grid: (n / 2^A / 2 / 256, W, B)
template <int A, int PRE>
__global__ void __launch_bounds__(256) k_head_fwd(LimbJob job, const ModInfo *__restrict__ mods) {
  constexpr int R = 1 << A;
  const int w = blockIdx.y, inst = blockIdx.z, n = job.n;
  const int c2 = blockIdx.x * 256 + threadIdx.x;  // column pair index, < n / R / 2
  const int stride2 = n >> (A + 1);               // block stride in 16-byte units
  const ModInfo M = mods[job.rowmod[w]];
  const u64 q = M.q, q2 = 2 * q;
  const int drow = job.rowdst ? job.rowdst[w] : w;
  const int srow = job.rowsrc ? job.rowsrc[w] : drow;
  RngStream rs;
  if (PRE == PRE_TERNARY || PRE == PRE_CBD) rs = rng_stream(job.rng, job.domain, job.a0 + (u64)inst, job.b);
  u64 x[R], y[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    ulonglong2 v = limb_load_pair<PRE>(job, M, mods, n, inst, srow, c2 + r * stride2, rs);
    x[r] = v.x; y[r] = v.y;
  }
#pragma unroll
  for (int b = A - 1; b >= 0; --b) {  // stage s = A-1-b: 2^s groups, partner r ^ 2^b
    const int s = A - 1 - b;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (r & (1 << b)) continue;
      const ulonglong2 tw = __ldg(&M.tw[(1 << s) + (r >> (b + 1))]);
      bf_fwd<AR_SHOUP>(x[r], x[r | (1 << b)], tw, q, q2);
      bf_fwd<AR_SHOUP>(y[r], y[r | (1 << b)], tw, q, q2);
    }
  }
  ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(job.dst + (size_t)inst * job.dst_is + (size_t)drow * n) + c2;
#pragma unroll
  for (int r = 0; r < R; ++r) dst[r * stride2] = make_ulonglong2(x[r], y[r]);  // < 4q: the tail guards it
}

template <int A, int POST>
__global__ void __launch_bounds__(256) k_head_inv(LimbJob job, const ModInfo *__restrict__ mods) {
  constexpr int R = 1 << A;
  const int w = blockIdx.y, inst = blockIdx.z, n = job.n;
  const int c2 = blockIdx.x * 256 + threadIdx.x;
  const int stride2 = n >> (A + 1);
  const ModInfo M = mods[job.rowmod[w]];
  const u64 q = M.q, q2 = 2 * q;
  const int drow = job.rowdst ? job.rowdst[w] : w;
  const int srow = job.rowsrc ? job.rowsrc[w] : drow;
  const int arow = job.rowmul ? job.rowmul[w] : w;
  const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(job.src + (size_t)inst * job.src_is + (size_t)srow * n) + c2;
  u64 x[R], y[R];
#pragma unroll
  for (int r = 0; r < R; ++r) { ulonglong2 v = src[r * stride2]; x[r] = v.x; y[r] = v.y; }  // in [0,2q) from the tail
#pragma unroll
  for (int b = 0; b < A; ++b) {  // stage s = A-1-b, Gentleman-Sande order: small gaps (in block units) first
    const int s = A - 1 - b;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (r & (1 << b)) continue;
      const int r2 = r | (1 << b);
      if (s == 0) {  // last stage of the whole transform: fold N^-1
        const u64 ux = x[r], vx = x[r2], uy = y[r], vy = y[r2];
        x[r] = mul_shoup_lazy(ux + vx, M.ninv, M.ninv_s, q);
        x[r2] = mul_shoup_lazy(ux + q2 - vx, M.wl_ninv, M.wl_ninv_s, q);
        y[r] = mul_shoup_lazy(uy + vy, M.ninv, M.ninv_s, q);
        y[r2] = mul_shoup_lazy(uy + q2 - vy, M.wl_ninv, M.wl_ninv_s, q);
      } else {
        const ulonglong2 tw = __ldg(&M.itw[(1 << s) + (r >> (b + 1))]);
        bf_inv<AR_SHOUP>(x[r], x[r2], tw, q, q2);
        bf_inv<AR_SHOUP>(y[r], y[r2], tw, q, q2);
      }
    }
  }
  ModDownRow md;
  if (POST == POST_MODDOWN) md = moddown_row(job, n, inst, w);
#pragma unroll
  for (int r = 0; r < R; ++r)
    limb_store_pair<POST>(job, M, md, n, inst, drow, arow, c2 + r * stride2, make_ulonglong2(csub(x[r], q), csub(y[r], q)));
}

// ---- launchers, one translation unit per size (limb_12.cu, limb_13.cu, limb_14.cu, limb_big.cu)
// return a cudaError_t as int; W rows x B instances
template <int LOGN>
int limb_dispatch(int combo, int ar, const LimbJob &job, const ModInfo *mods, int W, int B, cudaStream_t stream);
// N = 2^(13+A), A in {2,3}: every combo as head/tail sequences (Shoup arithmetic; these sizes use 55-60-bit primes)
int limb_dispatch_big(int A, int combo, const LimbJob &job, const ModInfo *mods, int W, int B, cudaStream_t stream);

#ifdef ABC_LIMB_IMPL
template <int LOGN, int PRE, bool FWD, bool MUL, bool INV, int POST, int AR, bool TAIL>
static int limb_launch(const LimbJob &job, const ModInfo *mods, int W, int B, cudaStream_t stream) {
  typedef NttDims<LOGN> D;
  auto kern = k_limb<LOGN, PRE, FWD, MUL, INV, POST, AR, TAIL>;
  if (D::SMEM > 48 * 1024) {
    static bool done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!done[dev & 63]) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D::SMEM);
      if (e != cudaSuccess) return (int)e;
      done[dev & 63] = true;
    }
  }
  kern<<<dim3(W, B), D::T, D::SMEM, stream>>>(job, mods);
  return (int)cudaGetLastError();
}
#endif

#ifdef ABC_LIMB_IMPL_SIZE
template <int LOGN, int AR>
static int limb_dispatch_ar(int combo, const LimbJob &j, const ModInfo *m, int W, int B, cudaStream_t s) {
  switch (combo) {
    case LIMB_FWD: return limb_launch<LOGN, PRE_LOAD, true, false, false, POST_STORE, AR, false>(j, m, W, B, s);
    case LIMB_INV: return limb_launch<LOGN, PRE_LOAD, false, false, true, POST_STORE, AR, false>(j, m, W, B, s);
    case LIMB_REDUCE_FWD: return limb_launch<LOGN, PRE_REDUCE, true, false, false, POST_STORE, AR, false>(j, m, W, B, s);
    case LIMB_PLAINLIFT_FWD: return limb_launch<LOGN, PRE_PLAIN_LIFT, true, false, false, POST_STORE, AR, false>(j, m, W, B, s);
    case LIMB_TERNARY_FWD: return limb_launch<LOGN, PRE_TERNARY, true, false, false, POST_STORE, AR, false>(j, m, W, B, s);
    case LIMB_CBD_FWD: return limb_launch<LOGN, PRE_CBD, true, false, false, POST_STORE, AR, false>(j, m, W, B, s);
    case LIMB_ENCODE_INV: return limb_launch<LOGN, PRE_ENCODE, false, false, true, POST_STORE, AR, false>(j, m, W, B, s);
    case LIMB_FWD_DECODE: return limb_launch<LOGN, PRE_LOAD, true, false, false, POST_DECODE, AR, false>(j, m, W, B, s);
    case LIMB_FWD_MUL_INV: return limb_launch<LOGN, PRE_LOAD, true, true, true, POST_STORE, AR, false>(j, m, W, B, s);
    case LIMB_FWD_MUL_INV_ADD: return limb_launch<LOGN, PRE_LOAD, true, true, true, POST_ADD, AR, false>(j, m, W, B, s);
    case LIMB_MUL_INV: return limb_launch<LOGN, PRE_LOAD, false, true, true, POST_STORE, AR, false>(j, m, W, B, s);
    case LIMB_GALOIS_REDUCE_FWD: return limb_launch<LOGN, PRE_GALOIS_REDUCE, true, false, false, POST_STORE, AR, false>(j, m, W, B, s);
    case LIMB_INV_MODDOWN: return limb_launch<LOGN, PRE_LOAD, false, false, true, POST_MODDOWN, AR, false>(j, m, W, B, s);
    case LIMB_KSINNER_INV_MODDOWN:
      if constexpr (AR == AR_F64) return limb_launch<LOGN, PRE_KS_INNER, false, false, true, POST_MODDOWN, AR, false>(j, m, W, B, s);
      return (int)cudaErrorInvalidValue;
    case LIMB_BEHZTENSOR_INV:
      if constexpr (AR == AR_F64) return limb_launch<LOGN, PRE_BEHZ_TENSOR, false, false, true, POST_STORE, AR, false>(j, m, W, B, s);
      return (int)cudaErrorInvalidValue;
    default: return (int)cudaErrorInvalidValue;
  }
}
template <int LOGN>
int limb_dispatch(int combo, int ar, const LimbJob &job, const ModInfo *mods, int W, int B, cudaStream_t stream) {
  LimbJob j = job;
  j.n = 1 << LOGN; j.sub = 0;
  switch (ar) {
    case AR_SHOUP: return limb_dispatch_ar<LOGN, AR_SHOUP>(combo, j, mods, W, B, stream);
    case AR_FP: return limb_dispatch_ar<LOGN, AR_FP>(combo, j, mods, W, B, stream);
    case AR_FP_LAZY: return limb_dispatch_ar<LOGN, AR_FP_LAZY>(combo, j, mods, W, B, stream);
    case AR_F64: return limb_dispatch_ar<LOGN, AR_F64>(combo, j, mods, W, B, stream);
    default: return (int)cudaErrorInvalidValue;
  }
}
#endif
