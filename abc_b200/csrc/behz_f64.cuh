// behz_f64.cuh — the BEHZ product's base conversions and tensor product on the FP64 pipe, over an auxiliary base of
// sub-2^45 primes (exact-double class contexts only).
//
// Evaluator::bfv_multiply (SealCiphertext.cpp:104,122) lifts the operands from q to Bsk = B U {m_sk}, multiplies in both
// bases, and comes back with fast_floor + fastbconv_sk.  Which primes make up B and m_sk does not show in the result: every
// conversion is either exact (Shenoy-Kumaresan, the m~ Montgomery step) or has an error term that depends on the q
// residues only, so the auxiliary primes merely carry intermediate integers and need a large enough product (SEAL's own
// rule: more than 32 + bits(t) + bits(Q) bits).  SEAL takes 61-bit primes; k_behz_lift / k_behz_scale spend their time
// on 128-bit integer MACs for them (ncu: IMAD pipe 81-83 %) and their transforms run on the Shoup class.  Here the base is
// made of 44-bit primes: one more row per polynomial, but every transform runs on the exact-double class and every
// modular product below is six FP64 instructions.  The oracle (SEAL's base) pins the result bit for bit
// (tests/test_gpu_parity.py: size-3 product and mul+relin).
//
// Representatives: a value that feeds conversions into SEVERAL moduli (z_i, zb_j, alpha) is canonical, exactly as in
// the integer kernels; everything that stays inside one modulus may be centred until it is stored.
#pragma once
#include "limb.cuh"

#define BF_MAXQ 8    // data limbs
#define BF_MAXB 12   // |Bsk| of the sub-2^45 base
struct BehzF64 {
  int L, nB, nbsk;
  double q[BF_MAXQ], qinv[BF_MAXQ], lift_c[BF_MAXQ], scale_c[BF_MAXQ], B_mod_q[BF_MAXQ];
  u32 punct_q_mt[BF_MAXQ], neg_inv_q_mt;
  double b[BF_MAXB], binv[BF_MAXB], q_mod_b[BF_MAXB], inv_mt_b[BF_MAXB], t_mod_b[BF_MAXB], inv_q_b[BF_MAXB];
  double punct_q_b[BF_MAXB][BF_MAXQ];      // (Q/q_i) mod b_j
  double inv_punct_B[BF_MAXB], punct_B_msk[BF_MAXB], inv_B_msk, msk_half;
  double punct_B_q[BF_MAXQ][BF_MAXB];      // (B/b_j) mod q_i
};

// a * b mod m for integer-valued doubles, |a| < 2^51, 0 <= b < m < 2^45; centred result, |r| <= 0.6 m
__device__ __forceinline__ double bf_mul(double a, double b, double m, double minv) {
  const double ph = a * b;
  const double Q = rint_mul(ph, minv);
  const double pl = fma(a, b, -ph);
  return fma(-Q, m, ph) + pl;
}
__device__ __forceinline__ double bf_canon(double r, double m, double minv) {  // any |r| < 2^51 -> [0, m)
  return cadd_neg(reduce_f64(r, minv, m), m);
}
__device__ __forceinline__ double bf_mulc(double a, double b, double m, double minv) { return cadd_neg(bf_mul(a, b, m, minv), m); }
__device__ __forceinline__ double bf_in(u64 x) { return f64_of(ar_from_canon<AR_F64>(x)); }
__device__ __forceinline__ u64 bf_out(double r) { return f64_canon_bits(r); }

// ---- step 1: x * m~ -> FastBConv q -> Bsk U {m~} -> SmMRq (RNSTool::fastbconv_m_tilde + sm_mrq)
// X layout [inst][4][W = L + NBSK][N]: q rows (a copy of the input), then the Bsk rows.  grid: (N/128, polys, B)
template <int L, int NBSK>
__global__ void __launch_bounds__(128) k_behz_lift_f64(const u64 *__restrict__ a, const u64 *__restrict__ b,
                                                       u64 *__restrict__ X, const BehzF64 *__restrict__ F, int N, int copy_q = 1) {
  const int n = blockIdx.x * 128 + threadIdx.x, poly = blockIdx.y, inst = blockIdx.z;
  constexpr int W = L + NBSK;
  const u64 *src = (poly < 2 ? a : b) + ((size_t)inst * 2 + (poly & 1)) * L * N + n;
  u64 *dst = X + ((size_t)inst * 4 + poly) * W * N + n;
  double z[L];
  u32 xm = 0;
#pragma unroll
  for (int i = 0; i < L; ++i) {
    const u64 x = src[(size_t)i * N];
    if (copy_q) dst[(size_t)i * N] = x;   // (0: the forward transforms read the operand's own limbs)
    z[i] = bf_mulc(bf_in(x), F->lift_c[i], F->q[i], F->qinv[i]);
    xm += (u32)bf_out(z[i]) * F->punct_q_mt[i];
  }
  const double rc = (double)(int)(xm * F->neg_inv_q_mt);  // [-x * Q^-1] mod 2^32, centred
#pragma unroll
  for (int j = 0; j < NBSK; ++j) {
    const double m = F->b[j], mi = F->binv[j];
    double s = bf_mul(rc, F->q_mod_b[j], m, mi);
#pragma unroll
    for (int i = 0; i < L; ++i) s += bf_mul(z[i], F->punct_q_b[j][i], m, mi);
    // (x + Q*r) * m~^-1 mod b_j
    dst[(size_t)(L + j) * N] = bf_out(bf_mulc(reduce_f64(s, mi, m), F->inv_mt_b[j], m, mi));
  }
}

// ---- tensor product in q and Bsk, in place: (a0,a1,b0,b1) -> (a0*b0, a0*b1 + a1*b0, a1*b1).  grid: (N/256, W, B)
__global__ void __launch_bounds__(256) k_behz_tensor_f64(u64 *__restrict__ X, const ModInfo *__restrict__ mods,
                                                         const int *__restrict__ rowmod, int N, int W, int square) {
  const int n = blockIdx.x * 256 + threadIdx.x, w = blockIdx.y, inst = blockIdx.z;
  const ModInfo *Mp = mods + rowmod[w];
  const double m = (double)Mp->q, mi = f64_of(Mp->qinv_bits);
  u64 *p = X + ((size_t)inst * 4 * W + w) * N + n;
  const size_t ps = (size_t)W * N;
  const double a0 = bf_in(p[0]), a1 = bf_in(p[ps]);
  const double b0 = square ? a0 : bf_in(p[2 * ps]), b1 = square ? a1 : bf_in(p[3 * ps]);
  p[0] = bf_out(bf_mulc(a0, b0, m, mi));
  p[ps] = bf_out(bf_canon(bf_mul(a0, b1, m, mi) + bf_mul(a1, b0, m, mi), m, mi));
  p[2 * ps] = bf_out(bf_mulc(a1, b1, m, mi));
}

// ---- steps 6-8: * t, fast_floor (q U Bsk -> Bsk), fastbconv_sk (Bsk -> q).  grid: (N/128, 3, B)
template <int L, int NBSK>
__global__ void __launch_bounds__(128) k_behz_scale_f64(const u64 *__restrict__ X, u64 *__restrict__ dst,
                                                        const BehzF64 *__restrict__ F, int N, int polys_per_inst = 4) {
  constexpr int W = L + NBSK, NB = NBSK - 1;
  const int n = blockIdx.x * 128 + threadIdx.x, poly = blockIdx.y, inst = blockIdx.z;
  const u64 *src = X + ((size_t)inst * polys_per_inst + poly) * W * N + n;
  u64 *out = dst + ((size_t)inst * 3 + poly) * L * N + n;
  double z[L], zb[NB];
#pragma unroll
  for (int i = 0; i < L; ++i) z[i] = bf_mulc(bf_in(src[(size_t)i * N]), F->scale_c[i], F->q[i], F->qinv[i]);
  double ysk = 0.0;
#pragma unroll
  for (int j = 0; j < NBSK; ++j) {
    const double m = F->b[j], mi = F->binv[j];
    double s = bf_mul(bf_in(src[(size_t)(L + j) * N]), F->t_mod_b[j], m, mi);
#pragma unroll
    for (int i = 0; i < L; ++i) s -= bf_mul(z[i], F->punct_q_b[j][i], m, mi);
    const double y = bf_mul(reduce_f64(s, mi, m), F->inv_q_b[j], m, mi);   // centred, |y| <= 0.6 m
    if (j < NB) zb[j] = bf_mulc(y, F->inv_punct_B[j], m, mi);
    else ysk = y;
  }
  // Shenoy-Kumaresan: alpha = (FastBConv_{B->m_sk}(y) - y_sk) * B^-1 mod m_sk, centred
  const double msk = F->b[NB], mski = F->binv[NB];
  double s = -ysk;
#pragma unroll
  for (int j = 0; j < NB; ++j) s += bf_mul(zb[j], F->punct_B_msk[j], msk, mski);
  const double alpha = bf_mulc(reduce_f64(s, mski, msk), F->inv_B_msk, msk, mski);  // canonical
  const bool negative = alpha > F->msk_half;
  const double amag = negative ? msk - alpha : alpha;
#pragma unroll
  for (int i = 0; i < L; ++i) {
    const double q = F->q[i], qi = F->qinv[i];
    const double corr = bf_mul(amag, F->B_mod_q[i], q, qi);
    double c = negative ? corr : -corr;
#pragma unroll
    for (int j = 0; j < NB; ++j) c += bf_mul(zb[j], F->punct_B_q[i][j], q, qi);
    out[(size_t)i * N] = bf_out(bf_canon(c, q, qi));
  }
}
