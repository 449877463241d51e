// devconst.cuh — per-context constants shared by all kernels (device copy lives in ctx->dC).
#pragma once
#include "modarith.cuh"

#define ABC_MAXL 32  // max data limbs (base conversion is register-resident and fully unrolled up to 15)

// Per-context constants (device copy).  Index convention for `mods`: 0..k-1 key-level primes,
// k..k+nbsk-1 Bsk = (B_0..B_{nB-1}, m_sk), k+nbsk = plain modulus t, k+nbsk+1 = gamma.
struct DevConst {
  int N, logN, k, L, nB, nbsk;
  u64 q[ABC_MAXL + 1];                       // key-level primes (q[L] = special prime p)
  u64 q_mu_hi[ABC_MAXL + 1], q_mu_lo[ABC_MAXL + 1];
  u64 t, t_half_up, q_mod_t, t_mu_hi, t_mu_lo;
  u64 delta[ABC_MAXL];                       // floor(Q/t) mod q_i
  u64 p, p_half, p_mu_hi;
  u64 inv_p[ABC_MAXL], inv_p_s[ABC_MAXL], p_half_mod_q[ABC_MAXL], p_mod_q[ABC_MAXL];
  // decryption
  u64 gamma, gamma_half, g_mu_hi, g_mu_lo;
  u64 dec_c[ABC_MAXL], dec_c_s[ABC_MAXL];    // (t*gamma) * (Q/q_i)^-1 mod q_i
  u64 punct_t[ABC_MAXL], punct_g[ABC_MAXL];  // (Q/q_i) mod t, mod gamma
  u64 neg_inv_q_t, neg_inv_q_t_s, neg_inv_q_g, neg_inv_q_g_s, inv_g_t, inv_g_t_s;
  // BEHZ
  u64 bsk[ABC_MAXL + 1], bsk_mu_hi[ABC_MAXL + 1], bsk_mu_lo[ABC_MAXL + 1];
  u64 lift_c[ABC_MAXL], lift_c_s[ABC_MAXL];        // m~ * (Q/q_i)^-1 mod q_i
  u64 punct_q_bsk[ABC_MAXL + 1][ABC_MAXL];         // (Q/q_i) mod bsk_j
  u32 punct_q_mt[ABC_MAXL];                        // (Q/q_i) mod 2^32
  u32 neg_inv_q_mt;                                // -Q^-1 mod 2^32
  u64 q_mod_bsk[ABC_MAXL + 1];
  u64 inv_mt_bsk[ABC_MAXL + 1], inv_mt_bsk_s[ABC_MAXL + 1];
  u64 scale_c[ABC_MAXL], scale_c_s[ABC_MAXL];      // t * (Q/q_i)^-1 mod q_i
  u64 t_mod_bsk[ABC_MAXL + 1], t_mod_bsk_s[ABC_MAXL + 1];
  u64 inv_q_bsk[ABC_MAXL + 1], inv_q_bsk_s[ABC_MAXL + 1];
  u64 inv_punct_B[ABC_MAXL], inv_punct_B_s[ABC_MAXL];
  u64 punct_B_q[ABC_MAXL][ABC_MAXL];               // (B/B_j) mod q_i   [i][j]
  u64 punct_B_msk[ABC_MAXL];
  u64 inv_B_msk, inv_B_msk_s;
  u64 B_mod_q[ABC_MAXL], B_mod_q_s[ABC_MAXL];
};

