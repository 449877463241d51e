// kernels.cuh — sm_100a kernels of the BFV hot path.  Each kernel names the SEAL 3.6.5 routine it
// replaces and the ABC call site (paths relative to /root/reference) that reaches it.
// Layouts: ciphertext [inst][poly][limb][N]; keys [J][comp][k][N] (NTT form); all u64, canonical residues.
// No tensor cores: this is 64-bit modular integer work; kernels are bound by the INT pipes (NTT, base
// conversion, key inner product) or by HBM (add/sub/negate/permute/moddown).
#pragma once
#include "limb.cuh"

// ------------------------------------------------------------------------------------------------
// add / sub / negate.  Evaluator::add_inplace / sub_inplace / negate_inplace
// (src/runtime/SealCiphertext.cpp:92,98,114,118,157,193).  HBM-bound: 24 B per coefficient.
// grid: (ceil(N/2/256), polys*L, B); 16-byte accesses.
template <int OP>  // 0 add, 1 sub, 2 negate
__global__ void __launch_bounds__(256) k_addsub(u64 *__restrict__ dst, const u64 *__restrict__ a,
                                                const u64 *__restrict__ b, const DevConst *__restrict__ C, int N,
                                                int L, long long istride, int i0, int nrows) {
  // grid.y = 2 * nrows: limbs i0 .. i0+nrows-1 of both polynomials (all L limbs unless the context is limb-sharded)
  const int e2 = blockIdx.x * 256 + threadIdx.x;
  if (e2 >= N / 2) return;
  const int limb = i0 + (int)(blockIdx.y % nrows), row = (int)(blockIdx.y / nrows) * L + limb;
  const u64 q = C->q[limb];
  const size_t off = (size_t)blockIdx.z * istride + (size_t)row * N;
  ulonglong2 x = reinterpret_cast<const ulonglong2 *>(a + off)[e2], r;
  if (OP == 2) {
    r.x = neg_mod(x.x, q); r.y = neg_mod(x.y, q);
  } else {
    ulonglong2 y = reinterpret_cast<const ulonglong2 *>(b + off)[e2];
    if (OP == 0) { r.x = add_mod(x.x, y.x, q); r.y = add_mod(x.y, y.y, q); }
    else { r.x = sub_mod(x.x, y.x, q); r.y = sub_mod(x.y, y.y, q); }
  }
  reinterpret_cast<ulonglong2 *>(dst + off)[e2] = r;
}

// ------------------------------------------------------------------------------------------------
// BEHZ step 1: x*m~ -> FastBConv q -> Bsk U {m~} -> SmMRq.  RNSTool::fastbconv_m_tilde + sm_mrq
// (Evaluator::bfv_multiply, reached from src/runtime/SealCiphertext.cpp:104,122).
// X layout [inst][4][W=2L+1][N]; this kernel fills all W rows (q rows are a copy of the input).
// grid: (N/128, 4, B).  INT-bound: L Shoup products + L*(L+2) 128-bit MACs per coefficient.
// LT > 0: limb count known at compile time (base conversion fully unrolled, residues in registers);
// LT = 0: generic limb count Lrt <= ABC_MAXL (loops not unrolled, residues in local memory).
template <int LT>
__global__ void __launch_bounds__(128) k_behz_lift(const u64 *__restrict__ a, const u64 *__restrict__ b,
                                                   u64 *__restrict__ X, const DevConst *__restrict__ C, int N, int Lrt,
                                                   int col0 = 0, int copy_q = 1) {
  // col0: first coefficient of this launch (limb-sharded contexts convert N / world coefficients per rank: the base
  // conversion is independent per coefficient); copy_q = 0: the q rows of X are filled elsewhere
  constexpr int CAP = LT ? LT : ABC_MAXL;
  const int L = LT ? LT : Lrt;
  const int n = col0 + blockIdx.x * blockDim.x + threadIdx.x, poly = blockIdx.y, inst = blockIdx.z;
  const int W = 2 * L + 1;
  const u64 *src = (poly < 2 ? a : b) + ((size_t)inst * 2 + (poly & 1)) * L * N + n;
  u64 *dst = X + ((size_t)inst * 4 + poly) * W * N + n;
  u64 z[CAP];
  u32 xm = 0;
#pragma unroll
  for (int i = 0; i < L; ++i) {
    u64 x = src[(size_t)i * N];
    if (copy_q) dst[(size_t)i * N] = x;
    z[i] = mul_shoup(x, C->lift_c[i], C->lift_c_s[i], C->q[i]);
    xm += (u32)z[i] * C->punct_q_mt[i];
  }
  const u32 r = xm * C->neg_inv_q_mt;  // [-x * Q^-1] mod 2^32
#pragma unroll
  for (int j = 0; j <= L; ++j) {
    const u64 pm = C->bsk[j], mh = C->bsk_mu_hi[j], ml = C->bsk_mu_lo[j];
    u64 lo = 0, hi = 0;
#pragma unroll
    for (int i = 0; i < L; ++i) mac128(lo, hi, z[i], C->punct_q_bsk[j][i]);
    const u64 xb = barrett128(lo, hi, pm, mh, ml);
    // centred r (m~ is a power of two, hence '>='), then (x + Q*r) * m~^-1 mod p_j
    const u64 rc = (r >= 0x80000000u) ? (u64)r + (pm - 0x100000000ULL) : (u64)r;
    lo = xb; hi = 0;
    mac128(lo, hi, rc, C->q_mod_bsk[j]);
    const u64 v = barrett128(lo, hi, pm, mh, ml);
    dst[(size_t)(L + j) * N] = mul_shoup(v, C->inv_mt_bsk[j], C->inv_mt_bsk_s[j], pm);
  }
}

// BEHZ tensor product in q and Bsk (Evaluator::bfv_multiply "behz_ciphertext_product"), in place on X:
// (X0,X1,X2,X3) = (a0,a1,b0,b1) -> (a0*b0, a0*b1+a1*b0, a1*b1).  grid: (N/256, W, B)
// square != 0: both operands are the same ciphertext, only (X0,X1) were lifted: (a0*a0, a0*a1+a1*a0, a1*a1).
__global__ void __launch_bounds__(256) k_behz_tensor(u64 *__restrict__ X, const ModInfo *__restrict__ mods,
                                                     const int *__restrict__ rowmod, int N, int W, int square) {
  const int n = blockIdx.x * 256 + threadIdx.x, w = blockIdx.y, inst = blockIdx.z;
  const ModInfo *Mp = mods + rowmod[w];
  const u64 q = Mp->q, mh = Mp->mu_hi, ml = Mp->mu_lo;
  u64 *p = X + ((size_t)inst * 4 * W + w) * N + n;
  const size_t ps = (size_t)W * N;
  const u64 a0 = p[0], a1 = p[ps], b0 = square ? a0 : p[2 * ps], b1 = square ? a1 : p[3 * ps];
  p[0] = mul_mod(a0, b0, q, mh, ml);
  u64 lo = a0 * b1, hi = __umul64hi(a0, b1);
  mac128(lo, hi, a1, b0);
  p[ps] = barrett128(lo, hi, q, mh, ml);
  p[2 * ps] = mul_mod(a1, b1, q, mh, ml);
}

// BEHZ steps 6-8: *t, fast_floor (q U Bsk -> Bsk), fastbconv_sk (Bsk -> q).  RNSTool::fast_floor + fastbconv_sk.
// X [inst][4][W][N] (polys 0..2, coefficient form) -> dst3 [inst][3][L][N].  grid: (N/128, 3, B)
template <int LT>
__global__ void __launch_bounds__(128) k_behz_scale(const u64 *__restrict__ X, u64 *__restrict__ dst,
                                                    const DevConst *__restrict__ C, int N, int Lrt, int col0 = 0) {
  constexpr int CAP = LT ? LT : ABC_MAXL;
  const int L = LT ? LT : Lrt;
  const int n = col0 + blockIdx.x * blockDim.x + threadIdx.x, poly = blockIdx.y, inst = blockIdx.z;
  const int W = 2 * L + 1;
  const u64 *src = X + ((size_t)inst * 4 + poly) * W * N + n;
  u64 *out = dst + ((size_t)inst * 3 + poly) * L * N + n;
  u64 z[CAP], zb[CAP];
#pragma unroll
  for (int i = 0; i < L; ++i) z[i] = mul_shoup(src[(size_t)i * N], C->scale_c[i], C->scale_c_s[i], C->q[i]);
  u64 ysk = 0;
#pragma unroll
  for (int j = 0; j <= L; ++j) {
    const u64 pm = C->bsk[j];
    u64 lo = 0, hi = 0;
#pragma unroll
    for (int i = 0; i < L; ++i) mac128(lo, hi, z[i], C->punct_q_bsk[j][i]);
    const u64 conv = barrett128(lo, hi, pm, C->bsk_mu_hi[j], C->bsk_mu_lo[j]);
    const u64 xb = mul_shoup(src[(size_t)(L + j) * N], C->t_mod_bsk[j], C->t_mod_bsk_s[j], pm);
    const u64 y = mul_shoup(sub_mod(xb, conv, pm), C->inv_q_bsk[j], C->inv_q_bsk_s[j], pm);
    if (j < L) zb[j] = mul_shoup(y, C->inv_punct_B[j], C->inv_punct_B_s[j], pm);
    else ysk = y;
  }
  // Shenoy-Kumaresan: alpha = (FastBConv_{B->m_sk}(y) - y_sk) * B^-1 mod m_sk, centred
  const u64 msk = C->bsk[L];
  u64 lo = 0, hi = 0;
#pragma unroll
  for (int j = 0; j < L; ++j) mac128(lo, hi, zb[j], C->punct_B_msk[j]);
  const u64 conv_sk = barrett128(lo, hi, msk, C->bsk_mu_hi[L], C->bsk_mu_lo[L]);
  const u64 alpha = mul_shoup(sub_mod(conv_sk, ysk, msk), C->inv_B_msk, C->inv_B_msk_s, msk);
  const bool negative = alpha > (msk >> 1);
  const u64 amag = negative ? msk - alpha : alpha;
#pragma unroll
  for (int i = 0; i < L; ++i) {
    const u64 q = C->q[i];
    lo = 0; hi = 0;
#pragma unroll
    for (int j = 0; j < L; ++j) mac128(lo, hi, zb[j], C->punct_B_q[i][j]);
    const u64 conv = barrett128(lo, hi, q, C->q_mu_hi[i], C->q_mu_lo[i]);
    const u64 corr = mul_shoup(barrett64(amag, q, C->q_mu_hi[i]), C->B_mod_q[i], C->B_mod_q_s[i], q);
    out[(size_t)i * N] = negative ? add_mod(conv, corr, q) : sub_mod(conv, corr, q);
  }
}

// ------------------------------------------------------------------------------------------------
// Key switching (Evaluator::switch_key_inplace; relinearize at SealCiphertext.cpp:105,123, rotate at :55,60).
// ModUp + NTT is the limb pipeline (LIMB_REDUCE_FWD) into T [inst][k][L][N].
// Inner product: acc[inst][comp][I][n] = sum_J T[inst][I][J][n] * key[J][comp][I][n] mod q_I.
// grid: (N/512, k, B).  128-bit lazy accumulation, one Barrett reduction per output.
template <int LT>
__global__ void __launch_bounds__(256) k_ks_inner(const u64 *__restrict__ T, const u64 *__restrict__ key,
                                                  u64 *__restrict__ acc, const ModInfo *__restrict__ mods, int N, int k,
                                                  int Lrt, const int *__restrict__ imap) {
  // grid.y indexes imap: the output moduli this rank computes (all k unless the context is limb-sharded)
  const int L = LT ? LT : Lrt;
  // two adjacent coefficients per thread (16-byte accesses), J loop unrolled
  const int n2 = blockIdx.x * 256 + threadIdx.x, I = imap[blockIdx.y], inst = blockIdx.z;
  const ModInfo *Mp = mods + I;
  const u64 q = Mp->q, mh = Mp->mu_hi, ml = Mp->mu_lo;
  const ulonglong2 *t = reinterpret_cast<const ulonglong2 *>(T + ((size_t)(inst * k + I) * L) * N) + n2;
  const ulonglong2 *kp = reinterpret_cast<const ulonglong2 *>(key + (size_t)I * N) + n2;
  const int rowv = N >> 1, keyv = k * rowv;  // row / key-component strides in 16-byte units
  u64 lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0};
#pragma unroll
  for (int J = 0; J < L; ++J) {
    const ulonglong2 tv = t[J * rowv];
    const ulonglong2 k0 = __ldg(kp + (2 * J) * keyv), k1 = __ldg(kp + (2 * J + 1) * keyv);
    mac128(lo[0], hi[0], tv.x, k0.x); mac128(lo[1], hi[1], tv.y, k0.y);
    mac128(lo[2], hi[2], tv.x, k1.x); mac128(lo[3], hi[3], tv.y, k1.y);
  }
  ulonglong2 *o = reinterpret_cast<ulonglong2 *>(acc + ((size_t)inst * 2 * k + I) * N) + n2;
  o[0] = make_ulonglong2(barrett128(lo[0], hi[0], q, mh, ml), barrett128(lo[1], hi[1], q, mh, ml));
  o[keyv] = make_ulonglong2(barrett128(lo[2], hi[2], q, mh, ml), barrett128(lo[3], hi[3], q, mh, ml));
}

// ModDown with rounding and accumulate (tail of switch_key_inplace): acc is in coefficient form.
// dst[inst][comp][i][n] = base + p^-1 * (acc_i - ([acc_L + p/2]_p mod q_i) + [p/2]_{q_i}) mod q_i
// base0/base1: polynomial added into component 0 / 1 (nullptr = zero).  grid: (N/256, 2, B)
__global__ void __launch_bounds__(256) k_ks_moddown(const u64 *__restrict__ acc, const u64 *__restrict__ base0,
                                                    long long base0_is, const u64 *__restrict__ base1,
                                                    long long base1_is, u64 *__restrict__ dst,
                                                    const DevConst *__restrict__ C, int N, int L, int k) {
  const int n = blockIdx.x * 256 + threadIdx.x, comp = blockIdx.y, inst = blockIdx.z;
  const u64 *ac = acc + ((size_t)inst * 2 + comp) * k * N + n;
  const u64 *base = comp == 0 ? base0 : base1;
  const long long bis = comp == 0 ? base0_is : base1_is;
  if (base) base += (size_t)inst * bis + n;
  u64 *o = dst + ((size_t)inst * 2 + comp) * L * N + n;
  const u64 tl = add_mod(ac[(size_t)L * N], C->p_half, C->p);
  for (int i = 0; i < L; ++i) {
    const u64 q = C->q[i];
    const u64 r = sub_mod(barrett64(tl, q, C->q_mu_hi[i]), C->p_half_mod_q[i], q);
    u64 v = mul_shoup(sub_mod(ac[(size_t)i * N], r, q), C->inv_p[i], C->inv_p_s[i], q);
    if (base) v = add_mod(v, base[(size_t)i * N], q);
    o[(size_t)i * N] = v;
  }
}

// Galois automorphism in coefficient form (GaloisTool::apply_galois), gather formulation so that writes
// are coalesced: out[j] = +-in[j * elt^-1 mod 2N].  c0 -> out0 [inst][L][N] (stride out0_is), c1 -> out1.
// grid: (N/256, 2L, B)
__global__ void __launch_bounds__(256) k_galois(const u64 *__restrict__ ct, u64 *__restrict__ out0, long long out0_is,
                                                u64 *__restrict__ out1, long long out1_is, u32 elt_inv,
                                                const DevConst *__restrict__ C, int N, int L) {
  const int j = blockIdx.x * 256 + threadIdx.x, row = blockIdx.y, inst = blockIdx.z;
  const int poly = row / L, i = row % L;
  const u64 q = C->q[i];
  const u32 raw = ((u32)j * elt_inv) & (2u * N - 1);
  const u64 v = ct[((size_t)inst * 2 * L + row) * N + (raw & (N - 1))];
  u64 *o = poly == 0 ? out0 + (size_t)inst * out0_is : out1 + (size_t)inst * out1_is;
  o[(size_t)i * N + j] = (raw >= (u32)N) ? neg_mod(v, q) : v;
}

// ------------------------------------------------------------------------------------------------
// Plain add/sub with Delta-scaling (util::multiply_add/sub_plain_with_scaling_variant;
// SealCiphertext.cpp:134,145,175,184).  plain: [inst or 1][N] mod t.  grid: (N/256, 1, B)
// floor(x / t) for 128-bit x via the Barrett quotient
__device__ __forceinline__ u64 div128_by(u64 lo, u64 hi, u64 t, u64 mu_hi, u64 mu_lo) {
  u64 carry = __umul64hi(lo, mu_lo);
  u64 t2lo = lo * mu_hi, t2hi = __umul64hi(lo, mu_hi);
  u64 t1 = t2lo + carry;
  u64 t3 = t2hi + (t1 < t2lo);
  u64 t4lo = hi * mu_lo, t4hi = __umul64hi(hi, mu_lo);
  u64 t5 = t1 + t4lo;
  u64 c2 = t4hi + (t5 < t1);
  u64 qhat = hi * mu_hi + t3 + c2;
  u64 r = lo - qhat * t;
  return qhat + (r >= t);
}
__device__ __forceinline__ u64 scaled_plain(u64 m, u64 fix, int i, const DevConst *__restrict__ C) {
  u64 lo = fix, hi = 0;
  mac128(lo, hi, m, C->delta[i]);
  return barrett128(lo, hi, C->q[i], C->q_mu_hi[i], C->q_mu_lo[i]);
}
template <int SUB>
__global__ void __launch_bounds__(256) k_plain_addsub(u64 *__restrict__ dst, const u64 *__restrict__ a,
                                                      const u64 *__restrict__ plain, long long plain_is,
                                                      const DevConst *__restrict__ C, int N, int L, int i0, int i1) {
  const int n = blockIdx.x * 256 + threadIdx.x, inst = blockIdx.z;
  const u64 m = plain[(size_t)inst * plain_is + n];
  u64 lo = C->t_half_up, hi = 0;
  mac128(lo, hi, m, C->q_mod_t);
  const u64 fix = div128_by(lo, hi, C->t, C->t_mu_hi, C->t_mu_lo);
  const size_t base = (size_t)inst * 2 * L * N + n;
  for (int i = i0; i < i1; ++i) {  // limbs [i0, i1): all of them unless the context is limb-sharded
    const u64 q = C->q[i], s = scaled_plain(m, fix, i, C);
    const u64 x = a[base + (size_t)i * N];
    dst[base + (size_t)i * N] = SUB ? sub_mod(x, s, q) : add_mod(x, s, q);
    if (dst != a) dst[base + (size_t)(L + i) * N] = a[base + (size_t)(L + i) * N];
  }
}

// ------------------------------------------------------------------------------------------------
// Encryption tail (Encryptor::encrypt_zero_internal + RNSTool::divide_and_round_q_last_inplace +
// multiply_add_plain_with_scaling_variant; SealCiphertextFactory.cpp:12).
// tmp [inst][2][k][N] = INTT(pk_j * NTT(u)) in coefficient form.  Adds e_j ~ CBD, divides by p with
// rounding, adds the scaled plaintext to component 0.  rnd [inst][3][N]: the words of the encryption's streams
// (DOM_ENC, nonce0 + inst, b = 0 (u), 1 (e_0), 2 (e_1)) from k_rng_fill.  grid: (N/256, 2, B)
__global__ void __launch_bounds__(256) k_enc_finish(const u64 *__restrict__ tmp, const u64 *__restrict__ plain,
                                                    long long plain_is, u64 *__restrict__ ct, const u64 *__restrict__ rnd,
                                                    const DevConst *__restrict__ C, int N, int L, int k) {
  const int n = blockIdx.x * 256 + threadIdx.x, comp = blockIdx.y, inst = blockIdx.z;
  const int e = cbd_of(rnd[((size_t)inst * 3 + 1 + comp) * N + n]);
  const u64 *tp = tmp + ((size_t)inst * 2 + comp) * k * N + n;
  u64 last = add_mod(tp[(size_t)L * N], small_to_mod(e, C->p), C->p);
  last = add_mod(last, C->p_half, C->p);
  u64 m = 0, fix = 0;
  if (comp == 0) {
    m = plain[(size_t)inst * plain_is + n];
    u64 lo = C->t_half_up, hi = 0;
    mac128(lo, hi, m, C->q_mod_t);
    fix = div128_by(lo, hi, C->t, C->t_mu_hi, C->t_mu_lo);
  }
  u64 *o = ct + ((size_t)inst * 2 + comp) * L * N + n;
  for (int i = 0; i < L; ++i) {
    const u64 q = C->q[i];
    const u64 d = add_mod(tp[(size_t)i * N], small_to_mod(e, q), q);
    const u64 r = sub_mod(barrett64(last, q, C->q_mu_hi[i]), C->p_half_mod_q[i], q);
    u64 v = mul_shoup(sub_mod(d, r, q), C->inv_p[i], C->inv_p_s[i], q);
    if (comp == 0) v = add_mod(v, scaled_plain(m, fix, i, C), q);
    o[(size_t)i * N] = v;
  }
}

// Decryption tail: RNSTool::decrypt_scale_and_round (Decryptor::bfv_decrypt; SealCiphertextFactory.cpp:150).
// x [inst][L][N] = c0 + c1*s (coefficient form) -> plain [inst][N] mod t.  grid: (N/128, 1, B)
template <int LT>
__global__ void __launch_bounds__(128) k_dec_finish(const u64 *__restrict__ x, u64 *__restrict__ plain,
                                                    const DevConst *__restrict__ C, int N, int Lrt) {
  const int L = LT ? LT : Lrt;
  const int n = blockIdx.x * 128 + threadIdx.x, inst = blockIdx.z;
  const u64 *src = x + (size_t)inst * L * N + n;
  const u64 t = C->t, g = C->gamma;
  u64 lt = 0, ht = 0, lg = 0, hg = 0;
#pragma unroll
  for (int i = 0; i < L; ++i) {
    const u64 z = mul_shoup(src[(size_t)i * N], C->dec_c[i], C->dec_c_s[i], C->q[i]);
    mac128(lt, ht, z, C->punct_t[i]);
    mac128(lg, hg, z, C->punct_g[i]);
  }
  const u64 yt = mul_shoup(barrett128(lt, ht, t, C->t_mu_hi, C->t_mu_lo), C->neg_inv_q_t, C->neg_inv_q_t_s, t);
  const u64 yg = mul_shoup(barrett128(lg, hg, g, C->g_mu_hi, C->g_mu_lo), C->neg_inv_q_g, C->neg_inv_q_g_s, g);
  u64 d;
  if (yg > C->gamma_half) d = add_mod(yt, barrett64(g - yg, t, C->t_mu_hi), t);
  else d = sub_mod(yt, barrett64(yg, t, C->t_mu_hi), t);
  plain[(size_t)inst * N + n] = d ? mul_shoup(d, C->inv_g_t, C->inv_g_t_s, t) : 0;
}

// ------------------------------------------------------------------------------------------------
// Key generation helpers (KeyGenerator; SealCiphertextFactory.cpp:89-93)
// uniform a_i mod q_i for all key-level limbs of one polynomial.  grid: (N/256, k)
// (SEAL sample_poly_uniform: rejection above the largest multiple of q; retry j of coefficient n is word n + j * 2^24)
__global__ void __launch_bounds__(256) k_sample_uniform(u64 *__restrict__ dst, RngKey key, u64 domain, u64 a,
                                                        u64 b_base, const DevConst *__restrict__ C, int N) {
  const int n = blockIdx.x * 256 + threadIdx.x, i = blockIdx.y;
  const u64 q = C->q[i];
  const RngStream rs = rng_stream(key, domain, a, ((b_base | (u64)i) << 2) | 0);
  const u64 max_random = ~0ULL, max_multiple = max_random - (max_random % q) - 1;
  u64 r;
  for (u64 attempt = 0;; ++attempt) {
    r = rng_word(rs, (u64)n | (attempt << 24));
    if (r < max_multiple) break;
  }
  dst[(size_t)i * N + n] = r % q;
}
// The words of `streams` sampler streams per instance, generated once: dst [inst][streams][N], stream s of instance inst =
// rng_stream(key, domain, a0 + inst, b0 + s); one ChaCha20 block (8 words) per thread.  grid: (N/8/128, streams, B)
__global__ void __launch_bounds__(128) k_rng_fill(u64 *__restrict__ dst, RngKey key, u64 domain, u64 a0, u64 b0, int N) {
  const int blk = blockIdx.x * 128 + threadIdx.x, s = blockIdx.y, inst = blockIdx.z;
  if (blk * 8 >= N) return;
  const RngStream rs = rng_stream(key, domain, a0 + (u64)inst, b0 + (u64)s);
  u64 w[8];
  chacha20_block(rs, (u32)blk, w);
  ulonglong2 *o = reinterpret_cast<ulonglong2 *>(dst + ((size_t)inst * gridDim.y + s) * N + (size_t)blk * 8);
#pragma unroll
  for (int i = 0; i < 4; ++i) o[i] = make_ulonglong2(w[2 * i], w[2 * i + 1]);
}
// c0 = -(a*s + e) [+ factor * newkey at limb J]  (encrypt_zero_symmetric + generate_one_kswitch_key)
// key block layout [2][k][N]: c0 rows then c1 = a rows.  grid: (N/256, k)
__global__ void __launch_bounds__(256) k_ksk_finish(u64 *__restrict__ keyblk, const u64 *__restrict__ e_ntt,
                                                    const u64 *__restrict__ sk, const u64 *__restrict__ newkey, int J,
                                                    const ModInfo *__restrict__ mods, const DevConst *__restrict__ C,
                                                    int N, int k) {
  const int n = blockIdx.x * 256 + threadIdx.x, i = blockIdx.y;
  const ModInfo *Mp = mods + i;
  const u64 q = Mp->q;
  const size_t o = (size_t)i * N + n;
  const u64 a = keyblk[(size_t)k * N + o];
  u64 v = neg_mod(add_mod(mul_mod(a, sk[o], q, Mp->mu_hi, Mp->mu_lo), e_ntt[o], q), q);
  if (newkey && i == J) v = add_mod(v, mul_mod(newkey[o], C->p_mod_q[i], q, Mp->mu_hi, Mp->mu_lo), q);
  keyblk[o] = v;
}
// new key material: relin (sk^2) or Galois (NTT-domain permutation, GaloisTool::apply_galois_ntt).  grid: (N/256, k)
__global__ void __launch_bounds__(256) k_newkey(u64 *__restrict__ nk, const u64 *__restrict__ sk, u32 galois_elt,
                                                const ModInfo *__restrict__ mods, int N, int logN) {
  const int n = blockIdx.x * 256 + threadIdx.x, i = blockIdx.y;
  const size_t o = (size_t)i * N;
  if (galois_elt == 0) {
    const ModInfo *Mp = mods + i;
    const u64 s = sk[o + n];
    nk[o + n] = mul_mod(s, s, Mp->q, Mp->mu_hi, Mp->mu_lo);
  } else {
    const u32 rev = __brev((u32)(n + N)) >> (32 - (logN + 1));
    const u32 idx = (u32)((((u64)galois_elt * rev) >> 1) & (u64)(N - 1));
    nk[o + n] = sk[o + (__brev(idx) >> (32 - logN))];
  }
}

// ------------------------------------------------------------------------------------------------
// BatchEncoder scatter / gather as separate kernels for N >= 32768 (below that they are fused into the limb pipeline)
// grid: (N/256, 1, B)
__global__ void __launch_bounds__(256) k_encode_scatter(const long long *__restrict__ slots, long long slots_is, int n_slots,
                                                        const u32 *__restrict__ index_map, u64 *__restrict__ plain, u64 t, int N) {
  const int e = blockIdx.x * 256 + threadIdx.x, inst = blockIdx.z;
  const long long v = slots[(size_t)inst * slots_is + (e < n_slots ? e : n_slots - 1)];
  plain[(size_t)inst * N + index_map[e]] = v < 0 ? t + (u64)v : (u64)v;
}
__global__ void __launch_bounds__(256) k_decode_gather(const u64 *__restrict__ vals, const u32 *__restrict__ index_map,
                                                       long long *__restrict__ out, u64 t, int N) {
  const int e = blockIdx.x * 256 + threadIdx.x, inst = blockIdx.z;
  const u64 v = vals[(size_t)inst * N + index_map[e]];
  out[(size_t)inst * N + e] = v > (t >> 1) ? (long long)v - (long long)t : (long long)v;
}

// ------------------------------------------------------------------------------------------------
// integer-pipe issue-rate microbenchmarks (the INT roofline denominator)
__global__ void __launch_bounds__(1024) k_peak_imad(u32 *out, int iters) {
  u32 a = threadIdx.x, b = blockIdx.x | 1, c0 = 1, c1 = 2, c2 = 3, c3 = 4, c4 = 5, c5 = 6, c6 = 7, c7 = 8;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      c0 = c0 * a + b; c1 = c1 * a + b; c2 = c2 * a + b; c3 = c3 * a + b;
      c4 = c4 * a + b; c5 = c5 * a + b; c6 = c6 * a + b; c7 = c7 * a + b;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7;
}
__global__ void __launch_bounds__(1024) k_peak_iadd(u32 *out, int iters) {
  u32 a = threadIdx.x, b = blockIdx.x | 1, c0 = 1, c1 = 2, c2 = 3, c3 = 4, c4 = 5, c5 = 6, c6 = 7, c7 = 8;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      c0 = (c0 + a) ^ b; c1 = (c1 + a) ^ b; c2 = (c2 + a) ^ b; c3 = (c3 + a) ^ b;
      c4 = (c4 + a) ^ b; c5 = (c5 + a) ^ b; c6 = (c6 + a) ^ b; c7 = (c7 + a) ^ b;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^ c7;
}
// pure-FP64 butterfly on exact integer-valued doubles (class 3 of the microbenchmark):
//   Q = rint(y*w/q) (DFMA + DADD with the 1.5*2^52 magic), p = y*w as an error-free product (DMUL + DFMA),
//   v = (p_hi - Q*q) + p_lo (DFMA exact because the result is < q, + DADD), x' = x + v, y' = x - v.
__device__ __forceinline__ void bf_fwd_f64(double &x, double &y, double w, double winv, double q) {
  const double Q = fma(y, winv, 6755399441055744.0) - 6755399441055744.0;
  const double ph = y * w;
  const double pl = fma(y, w, -ph);
  const double v = fma(-Q, q, ph) + pl;
  y = x - v;
  x = x + v;
}
__global__ void __launch_bounds__(1024) k_peak_butterfly_f64(double *out, int iters, double q, double w, double winv) {
  double x0 = threadIdx.x, y0 = blockIdx.x, x1 = x0 + 1, y1 = y0 + 2, x2 = x0 + 3, y2 = y0 + 4, x3 = x0 + 5, y3 = y0 + 6;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      bf_fwd_f64(x0, y0, w, winv, q); bf_fwd_f64(x1, y1, w, winv, q); bf_fwd_f64(x2, y2, w, winv, q); bf_fwd_f64(x3, y3, w, winv, q);
    }
    // range reset (8 ops per 32 butterflies), keeps the values exact integers below 2^44
    x0 = fma(-floor(x0 * 5.6843418860808015e-14), 17592186044416.0, x0); y0 = fma(-floor(y0 * 5.6843418860808015e-14), 17592186044416.0, y0);
    x1 = fma(-floor(x1 * 5.6843418860808015e-14), 17592186044416.0, x1); y1 = fma(-floor(y1 * 5.6843418860808015e-14), 17592186044416.0, y1);
    x2 = fma(-floor(x2 * 5.6843418860808015e-14), 17592186044416.0, x2); y2 = fma(-floor(y2 * 5.6843418860808015e-14), 17592186044416.0, y2);
    x3 = fma(-floor(x3 * 5.6843418860808015e-14), 17592186044416.0, x3); y3 = fma(-floor(y3 * 5.6843418860808015e-14), 17592186044416.0, y3);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + y0 + x1 + y1 + x2 + y2 + x3 + y3;
}

// register-resident NTT butterflies per second for one arithmetic class (the ceiling NTT kernels are quoted against)
template <int AR>
__global__ void __launch_bounds__(1024) k_peak_butterfly(u64 *out, int iters, u64 q, u64 w, u64 wc) {
  u64 x0 = threadIdx.x, y0 = blockIdx.x, x1 = x0 + 1, y1 = y0 + 2, x2 = x0 + 3, y2 = y0 + 4, x3 = x0 + 5, y3 = y0 + 6;
  const ulonglong2 tw = make_ulonglong2(w, wc);
  const u64 q2 = ar_aux<AR>(q);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      bf_fwd<AR>(x0, y0, tw, q, q2); bf_fwd<AR>(x1, y1, tw, q, q2); bf_fwd<AR>(x2, y2, tw, q, q2); bf_fwd<AR>(x3, y3, tw, q, q2);
    }
    if (AR == AR_FP_LAZY) {  // the unguarded class needs a range reset now and then (8 ops per 32 butterflies)
      const u64 m = 0xFFFFFFFFFFULL;
      x0 &= m; y0 &= m; x1 &= m; y1 &= m; x2 &= m; y2 &= m; x3 &= m; y3 &= m;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ y0 ^ x1 ^ y1 ^ x2 ^ y2 ^ x3 ^ y3;
}

// Ciphertext::is_transparent (SEAL 3.6.5 ciphertext.h: every polynomial after c0 is zero), per instance: nonzero[inst] is
// set when any word of c1 is non-zero.  ct [B][2][L][N]; limbs [l0, l1) of c1 are looked at (limb-sharded: the owned ones).
// grid: (chunks, B).
__global__ void __launch_bounds__(256) k_c1_nonzero(const u64 *__restrict__ ct, int N, int L, int l0, int l1, int *__restrict__ nonzero) {
  const int inst = blockIdx.y;
  const u64 *c1 = ct + ((size_t)inst * 2 + 1) * L * N + (size_t)l0 * N;
  const size_t words = (size_t)(l1 - l0) * N;
  u64 acc = 0;
  for (size_t w = (size_t)blockIdx.x * 256 + threadIdx.x; w < words; w += (size_t)gridDim.x * 256) acc |= c1[w];
  if (__syncthreads_or(acc != 0) && threadIdx.x == 0) nonzero[inst] = 1;
}

// Decryptor::invariant_noise_budget, the multi-precision part: x [B][L][N] = c0 + c1*s.  Per coefficient: CRT-compose
// t*x (RNSBase::compose_array: sum_i [t x_i (Q/q_i)^-1]_{q_i} * (Q/q_i) mod Q), centre it
// (poly_infty_norm_coeffmod), take its bit length; the block's maximum goes to bits[inst] (atomicMax).
// tab: L rows of (Q/q_i), then Q, then (Q+1)/2, L little-endian words each.  grid: (N/128, B).  Diagnostic path.
__global__ void __launch_bounds__(128) k_noise_bits(const u64 *__restrict__ x, const DevConst *__restrict__ C,
                                                    const u64 *__restrict__ tab, int N, int L, int *__restrict__ bits) {
  const int n = blockIdx.x * 128 + threadIdx.x, inst = blockIdx.y;
  u64 acc[ABC_MAXL + 1];
  for (int w = 0; w <= L; ++w) acc[w] = 0;
  for (int i = 0; i < L; ++i) {
    const u64 y = mul_shoup(x[((size_t)inst * L + i) * N + n], C->scale_c[i], C->scale_c_s[i], C->q[i]);
    const u64 *P = tab + (size_t)i * L;
    u64 carry = 0;
    for (int w = 0; w < L; ++w) {
      const u64 lo = y * P[w], hi = __umul64hi(y, P[w]);
      const u64 s1 = acc[w] + lo, s2 = s1 + carry;
      carry = hi + (s1 < lo) + (s2 < carry);
      acc[w] = s2;
    }
    acc[L] += carry;
  }
  const u64 *Q = tab + (size_t)L * L, *H = Q + L;
  auto ge = [&](const u64 *m) {  // acc >= m (m has L words)
    if (acc[L]) return true;
    for (int w = L - 1; w >= 0; --w) if (acc[w] != m[w]) return acc[w] > m[w];
    return true;
  };
  while (ge(Q)) {
    u64 borrow = 0;
    for (int w = 0; w < L; ++w) {
      const u64 a = acc[w], d = a - Q[w] - borrow;
      borrow = (a < Q[w]) || (a == Q[w] && borrow);
      acc[w] = d;
    }
    acc[L] -= borrow;
  }
  if (ge(H)) {  // centre: Q - acc
    u64 borrow = 0;
    for (int w = 0; w < L; ++w) {
      const u64 a = Q[w], d = a - acc[w] - borrow;
      borrow = (a < acc[w]) || (a == acc[w] && borrow);
      acc[w] = d;
    }
  }
  int b = 0;
  for (int w = L - 1; w >= 0 && !b; --w) if (acc[w]) b = 64 * w + (64 - __clzll((long long)acc[w]));
  b = __reduce_max_sync(0xffffffffu, b);
  if ((threadIdx.x & 31) == 0) atomicMax(bits + inst, b);
}
