/*
 * bfv_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE ONLY; see bfv_oracle.h for scope and parity status).
 *
 * Restates the SEAL 3.6.5 BFV routines that /root/reference/src/runtime/SealCiphertext.cpp and
 * SealCiphertextFactory.cpp call.  Each function names the SEAL routine it follows and the ABC call
 * site that reaches it.  Everything is canonical-residue in/out; lazy ranges are internal.
 */
#include "bfv_oracle.h"
#include <malloc.h>
#include <stdlib.h>
#include <string.h>

typedef uint64_t u64;
typedef uint32_t u32;
typedef unsigned __int128 u128;

#define MAXK 64

/* ------------------------------------------------------------------ small modular helpers */
static inline u64 mulmod(u64 a, u64 b, u64 q) { return (u64)(((u128)a * b) % q); }
static inline u64 addmod(u64 a, u64 b, u64 q) { u64 s = a + b; return (s >= q) ? s - q : s; }
static inline u64 submod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }
static inline u64 negmod(u64 a, u64 q) { return a ? q - a : 0; }
static u64 powmod(u64 a, u64 e, u64 q) {
  u64 r = 1 % q; a %= q;
  while (e) { if (e & 1) r = mulmod(r, a, q); a = mulmod(a, a, q); e >>= 1; }
  return r;
}
/* modular inverse via extended Euclid (also for non-prime moduli such as 2^32 and 2N) */
static u64 invmod(u64 a, u64 m) {
  __int128 t = 0, nt = 1, r = (__int128)m, nr = (__int128)(a % m);
  while (nr) { __int128 qq = r / nr, tmp = t - qq * nt; t = nt; nt = tmp; tmp = r - qq * nr; r = nr; nr = tmp; }
  if (t < 0) t += m;
  return (u64)t;
}
static inline u64 shoup(u64 w, u64 q) { return (u64)(((u128)w << 64) / q); }
/* Harvey lazy product (SEAL multiply_uint_mod_lazy): result in [0,2q) for any 64-bit x */
static inline u64 mul_shoup_lazy(u64 x, u64 w, u64 ws, u64 q) {
  u64 h = (u64)(((u128)x * ws) >> 64);
  return x * w - h * q;
}

/* SEAL Modulus + barrett_reduce_128 / barrett_reduce_64 (util/uintarithsmallmod.h): const_ratio = floor(2^128/q) */
typedef struct { u64 q, mu_hi, mu_lo; } bmod;
static bmod bmod_init(u64 q) {
  bmod m; m.q = q;
  u128 all = ~(u128)0, r = all / q;
  if (all % q == q - 1) r += 1;
  m.mu_hi = (u64)(r >> 64); m.mu_lo = (u64)r;
  return m;
}
static inline u64 bred128(u128 x, const bmod *m) {
  u64 lo = (u64)x, hi = (u64)(x >> 64);
  u64 carry = (u64)(((u128)lo * m->mu_lo) >> 64);
  u128 t2 = (u128)lo * m->mu_hi;
  u64 t1 = (u64)t2 + carry;
  u64 t3 = (u64)(t2 >> 64) + (t1 < (u64)t2);
  u128 t4 = (u128)hi * m->mu_lo;
  u64 t5 = t1 + (u64)t4;
  u64 c2 = (u64)(t4 >> 64) + (t5 < t1);
  u64 qhat = hi * m->mu_hi + t3 + c2;
  u64 r = lo - qhat * m->q;
  return r >= m->q ? r - m->q : r;
}
static inline u64 bred64(u64 x, const bmod *m) {
  u64 r = x - (u64)(((u128)x * m->mu_hi) >> 64) * m->q;
  return r >= m->q ? r - m->q : r;
}
static inline u64 bmul(u64 a, u64 b, const bmod *m) { return bred128((u128)a * b, m); }

/* deterministic Miller-Rabin for 64-bit */
static int is_prime64(u64 n) {
  static const u64 sp[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
  if (n < 2) return 0;
  for (int i = 0; i < 12; i++) { if (n % sp[i] == 0) return n == sp[i]; }
  u64 d = n - 1; int s = 0;
  while (!(d & 1)) { d >>= 1; s++; }
  for (int i = 0; i < 12; i++) {
    u64 x = powmod(sp[i], d, n);
    if (x == 1 || x == n - 1) continue;
    int comp = 1;
    for (int r = 1; r < s; r++) { x = mulmod(x, x, n); if (x == n - 1) { comp = 0; break; } }
    if (comp) return 0;
  }
  return 1;
}

/* SEAL util::get_primes (numth.cpp): value = 2^bits - 2N + 1, step -2N, stop at 2^(bits-1) */
size_t obfv_get_primes(size_t N, int bits, size_t count, u64 *out) {
  u64 factor = 2 * (u64)N, value = ((u64)1 << bits) - factor + 1, lower = (u64)1 << (bits - 1);
  size_t n = 0;
  while (n < count && value > lower) {
    if (is_prime64(value)) out[n++] = value;
    value -= factor;
  }
  return n;
}

static u32 bitrev(u32 x, int bits) {
  u32 r = 0;
  for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
  return r;
}

/* ------------------------------------------------------------------ NTT tables (SEAL util/ntt.cpp) */
typedef struct {
  u64 q, psi;
  u64 *rp, *rps;   /* root_powers[bitrev(i)] = psi^i, and Shoup companions */
  u64 *irp, *irps; /* inverse of rp[j] at the same index j */
  u64 inv_n, inv_n_s;
} ntt_tab;

/* SEAL try_minimal_primitive_root: the smallest of all primitive 2N-th roots of unity */
static u64 minimal_primitive_root(u64 q, size_t N) {
  u64 two_n = 2 * (u64)N, g = 0;
  for (u64 x = 2;; x++) {
    g = powmod(x, (q - 1) / two_n, q);
    if (powmod(g, N, q) == q - 1) break;
  }
  u64 g2 = mulmod(g, g, q), cur = g, best = g;
  for (size_t i = 0; i < N; i++) { if (cur < best) best = cur; cur = mulmod(cur, g2, q); }
  return best;
}

static void ntt_tab_init(ntt_tab *T, u64 q, size_t N, int logN) {
  T->q = q;
  T->psi = minimal_primitive_root(q, N);
  T->rp = malloc(N * 8); T->rps = malloc(N * 8); T->irp = malloc(N * 8); T->irps = malloc(N * 8);
  u64 p = 1;
  for (size_t i = 0; i < N; i++) {
    u32 j = bitrev((u32)i, logN);
    T->rp[j] = p; T->rps[j] = shoup(p, q);
    p = mulmod(p, T->psi, q);
  }
  u64 ipsi = invmod(T->psi, q); p = 1;
  for (size_t i = 0; i < N; i++) {
    u32 j = bitrev((u32)i, logN);
    T->irp[j] = p; T->irps[j] = shoup(p, q);
    p = mulmod(p, ipsi, q);
  }
  T->inv_n = invmod((u64)N % q, q); T->inv_n_s = shoup(T->inv_n, q);
}
static void ntt_tab_free(ntt_tab *T) { free(T->rp); free(T->rps); free(T->irp); free(T->irps); }

/* SEAL ntt_negacyclic_harvey: Cooley-Tukey, natural in, bit-reversed out; Harvey lazy butterflies.
 * Input canonical, output canonical. */
static void ntt_fwd(const ntt_tab *T, u64 *x, size_t N) {
  const u64 q = T->q, q2 = 2 * q;
  for (size_t m = 1, gap = N >> 1; m < N; m <<= 1, gap >>= 1) {
    for (size_t i = 0; i < m; i++) {
      const u64 w = T->rp[m + i], ws = T->rps[m + i];
      u64 *a = x + 2 * i * gap, *b = a + gap;
      for (size_t j = 0; j < gap; j++) {
        u64 u = a[j]; if (u >= q2) u -= q2;
        u64 v = mul_shoup_lazy(b[j], w, ws, q);
        a[j] = u + v; b[j] = u + q2 - v;
      }
    }
  }
  for (size_t j = 0; j < N; j++) { u64 v = x[j]; if (v >= q2) v -= q2; if (v >= q) v -= q; x[j] = v; }
}
/* SEAL inverse_ntt_negacyclic_harvey: Gentleman-Sande, bit-reversed in, natural out, times N^-1 */
static void ntt_inv(const ntt_tab *T, u64 *x, size_t N) {
  const u64 q = T->q, q2 = 2 * q;
  for (size_t m = N >> 1, gap = 1; m >= 1; m >>= 1, gap <<= 1) {
    for (size_t i = 0; i < m; i++) {
      const u64 w = T->irp[m + i], ws = T->irps[m + i];
      u64 *a = x + 2 * i * gap, *b = a + gap;
      for (size_t j = 0; j < gap; j++) {
        u64 u = a[j], v = b[j];
        u64 s = u + v; if (s >= q2) s -= q2;
        a[j] = s; b[j] = mul_shoup_lazy(u + q2 - v, w, ws, q);
      }
    }
  }
  for (size_t j = 0; j < N; j++) {
    u64 v = mul_shoup_lazy(x[j], T->inv_n, T->inv_n_s, q);
    x[j] = v >= q ? v - q : v;
  }
}

/* ------------------------------------------------------------------ context */
struct obfv_ctx {
  size_t N; int logN; size_t k, L; u64 t;
  u64 q[MAXK]; ntt_tab nq[MAXK]; ntt_tab nt;
  bmod bq[MAXK], bbsk[MAXK], bt, bgamma;                  /* Barrett forms of the moduli */
  u32 *index_map;                                   /* BatchEncoder matrix_reps_index_map */
  /* plain scaling (SEAL ContextData: coeff_div_plain_modulus, coeff_modulus_mod_plain_modulus) */
  u64 q_mod_t, delta_mod_q[MAXK], t_half_up;
  /* key level: divide_and_round_q_last / key-switch ModDown */
  u64 p, p_half, inv_p_mod_q[MAXK], p_half_mod_q[MAXK], p_mod_q[MAXK];
  /* decryption (RNSTool: prod_t_gamma_mod_q, base_q_to_t_gamma_conv, neg_inv_q_mod_t_gamma, inv_gamma_mod_t) */
  u64 gamma, tgamma_mod_q[MAXK], inv_punct_q[MAXK], punct_q_mod_t[MAXK], punct_q_mod_gamma[MAXK];
  u64 neg_inv_q_mod_t, neg_inv_q_mod_gamma, inv_gamma_mod_t;
  /* BEHZ (RNSTool::initialize) */
  size_t nB, nbsk; u64 B[MAXK], msk, bsk[MAXK]; ntt_tab nbskt[MAXK];
  u64 mtilde, mtilde_mod_q[MAXK], punct_q_mod_bsk[MAXK][MAXK], punct_q_mod_mtilde[MAXK];
  u64 neg_inv_q_mod_mtilde, q_mod_bsk[MAXK], inv_mtilde_mod_bsk[MAXK], inv_q_mod_bsk[MAXK];
  u64 inv_punct_B[MAXK], punct_B_mod_q[MAXK][MAXK], punct_B_mod_msk[MAXK], inv_B_mod_msk, B_mod_q[MAXK];
  /* keys */
  u64 seed; u32 rng_key[8]; int rng_key_set; /* sampler key: expanded from seed unless obfv_set_rng_key was called */
  u64 *sk, *pk, *relin; u64 **galois; /* galois indexed by (elt-1)/2 */
};

static const u64 DEF_4096[] = {0xffffee001ULL, 0xffffc4001ULL, 0x1ffffe0001ULL};
static const u64 DEF_8192[] = {0x7fffffd8001ULL, 0x7fffffc8001ULL, 0xfffffffc001ULL, 0xffffff6c001ULL, 0xfffffebc001ULL};
static const u64 DEF_16384[] = {0xfffffffd8001ULL, 0xfffffffa0001ULL, 0xfffffff00001ULL, 0x1fffffff68001ULL,
                                0x1fffffff50001ULL, 0x1ffffffee8001ULL, 0x1ffffffea0001ULL, 0x1ffffffe88001ULL,
                                0x1ffffffe48001ULL};
static const u64 DEF_32768[] = {0x7fffffffe90001ULL, 0x7fffffffbf0001ULL, 0x7fffffffbd0001ULL, 0x7fffffffba0001ULL,
                                0x7fffffffaa0001ULL, 0x7fffffffa50001ULL, 0x7fffffff9f0001ULL, 0x7fffffff7e0001ULL,
                                0x7fffffff770001ULL, 0x7fffffff380001ULL, 0x7fffffff330001ULL, 0x7fffffff2d0001ULL,
                                0x7fffffff170001ULL, 0x7fffffff150001ULL, 0x7ffffffef00001ULL, 0xfffffffff70001ULL};

/* product of ps[0..n) except index skip (skip<0: none), reduced mod m */
static u64 prod_mod_except(const u64 *ps, size_t n, long skip, u64 m) {
  u64 r = 1 % m;
  for (size_t i = 0; i < n; i++) if ((long)i != skip) r = mulmod(r, ps[i] % m, m);
  return r;
}
/* bit length of prod(ps) — small multi-word product */
static int prod_bit_count(const u64 *ps, size_t n) {
  u64 w[MAXK + 2]; memset(w, 0, sizeof w); w[0] = 1; size_t len = 1;
  for (size_t i = 0; i < n; i++) {
    u64 carry = 0;
    for (size_t j = 0; j < len; j++) { u128 v = (u128)w[j] * ps[i] + carry; w[j] = (u64)v; carry = (u64)(v >> 64); }
    if (carry) w[len++] = carry;
  }
  int bits = 0; u64 top = w[len - 1];
  while (top) { bits++; top >>= 1; }
  return (int)(64 * (len - 1)) + bits;
}
static int bit_count64(u64 v) { int b = 0; while (v) { b++; v >>= 1; } return b; }

/* aux_bits != 0 (test hook, NOT SEAL behaviour): take the BEHZ auxiliary base B U {m_sk} as aux_count + 1 primes of aux_bits
   bits from get_primes (skipping the coefficient primes) instead of SEAL's 61-bit ones.  The product of bfv_multiply must not
   depend on that choice as long as the base satisfies SEAL's size rule; tests/test_oracle.py checks exactly this, which is
   what lets the CUDA path run the product over a sub-2^45 base (abc_b200/csrc/behz_f64.cuh).  gamma stays SEAL's. */
static int g_aux_bits = 0; static size_t g_aux_count = 0;
obfv_ctx *obfv_create(size_t N, const u64 *primes, size_t k, u64 t);
obfv_ctx *obfv_create_aux(size_t N, const u64 *primes, size_t k, u64 t, int aux_bits, size_t aux_count) {
  g_aux_bits = aux_bits; g_aux_count = aux_count;
  obfv_ctx *c = obfv_create(N, primes, k, t);
  g_aux_bits = 0; g_aux_count = 0;
  return c;
}
obfv_ctx *obfv_create(size_t N, const u64 *primes, size_t k, u64 t) {
  int logN = 0; while (((size_t)1 << logN) < N) logN++;
  if (((size_t)1 << logN) != N || N < 16 || N > 65536) return NULL;
  if (!primes) {
    switch (N) {
      case 4096: primes = DEF_4096; k = 3; break;
      case 8192: primes = DEF_8192; k = 5; break;
      case 16384: primes = DEF_16384; k = 9; break;
      case 32768: primes = DEF_32768; k = 16; break;
      default: return NULL; /* BFVDefault throws above 32768; 1024/2048 have no special prime */
    }
  }
  if (k < 2 || k >= MAXK - 2) return NULL;
  if (!t) { if (obfv_get_primes(N, 20, 1, &t) != 1) return NULL; } /* PlainModulus::Batching(N,20) */
  for (size_t i = 0; i < k; i++) if (!is_prime64(primes[i]) || (primes[i] - 1) % (2 * N)) return NULL;
  if (!is_prime64(t) || (t - 1) % (2 * N)) return NULL;

  /* keep large scratch buffers on the per-thread heaps: mmap/munmap per op serialises threads on the mm lock */
  mallopt(M_MMAP_THRESHOLD, 1 << 30);
  mallopt(M_TRIM_THRESHOLD, 1 << 30);
  obfv_ctx *c = calloc(1, sizeof *c);
  c->N = N; c->logN = logN; c->k = k; c->L = k - 1; c->t = t;
  const size_t L = c->L;
  for (size_t i = 0; i < k; i++) { c->q[i] = primes[i]; c->bq[i] = bmod_init(primes[i]); ntt_tab_init(&c->nq[i], primes[i], N, logN); }
  c->bt = bmod_init(t);
  ntt_tab_init(&c->nt, t, N, logN);

  /* BatchEncoder::populate_matrix_reps_index_map (batchencoder.cpp) */
  c->index_map = malloc(N * sizeof(u32));
  {
    u64 m = 2 * (u64)N, pos = 1; size_t row = N >> 1;
    for (size_t i = 0; i < row; i++) {
      c->index_map[i] = bitrev((u32)((pos - 1) >> 1), logN);
      c->index_map[row | i] = bitrev((u32)((m - pos - 1) >> 1), logN);
      pos = (pos * 3) & (m - 1);
    }
  }
  /* data-level Q = q_0..q_{L-1} */
  c->q_mod_t = prod_mod_except(c->q, L, -1, t);
  c->t_half_up = (t + 1) >> 1;
  for (size_t i = 0; i < L; i++) {
    /* floor(Q/t) mod q_i = -(Q mod t) * t^-1 mod q_i */
    c->delta_mod_q[i] = mulmod(negmod(c->q_mod_t % c->q[i], c->q[i]), invmod(t % c->q[i], c->q[i]), c->q[i]);
  }
  c->p = c->q[k - 1]; c->p_half = c->p >> 1;
  for (size_t i = 0; i < L; i++) {
    c->inv_p_mod_q[i] = invmod(c->p % c->q[i], c->q[i]);
    c->p_half_mod_q[i] = c->p_half % c->q[i];
    c->p_mod_q[i] = c->p % c->q[i];
  }
  /* RNSTool::initialize: base sizes and auxiliary primes */
  size_t nB = L;
  if (32 + bit_count64(t) + prod_bit_count(c->q, L) >= 61 * (int)L + 61) nB++;
  u64 aux[MAXK];
  if (obfv_get_primes(N, 61, nB + 2, aux) != nB + 2) { obfv_destroy(c); return NULL; }
  c->nB = nB; c->nbsk = nB + 1; c->msk = aux[0]; c->gamma = aux[1];
  for (size_t i = 0; i < nB; i++) { c->B[i] = aux[2 + i]; c->bsk[i] = aux[2 + i]; }
  if (g_aux_bits) {   /* test hook: see obfv_create_aux */
    u64 cand[MAXK], pick[MAXK]; size_t m = 0;
    size_t got = obfv_get_primes(N, g_aux_bits, g_aux_count + 1 + c->k + 2, cand);
    for (size_t a = 0; a < got && m < g_aux_count + 1; a++) {
      int used = cand[a] == t;
      for (size_t i = 0; i < c->k; i++) if (cand[a] == c->q[i]) used = 1;
      if (!used) pick[m++] = cand[a];
    }
    if (m != g_aux_count + 1) { obfv_destroy(c); return NULL; }
    nB = g_aux_count; c->nB = nB; c->nbsk = nB + 1; c->msk = pick[nB];
    for (size_t i = 0; i < nB; i++) { c->B[i] = pick[i]; c->bsk[i] = pick[i]; }
  }
  c->bsk[nB] = c->msk;
  for (size_t j = 0; j < c->nbsk; j++) { c->bbsk[j] = bmod_init(c->bsk[j]); ntt_tab_init(&c->nbskt[j], c->bsk[j], N, logN); }
  c->bgamma = bmod_init(c->gamma);
  c->mtilde = (u64)1 << 32;
  for (size_t i = 0; i < L; i++) {
    u64 qi = c->q[i];
    c->inv_punct_q[i] = invmod(prod_mod_except(c->q, L, (long)i, qi), qi);
    c->punct_q_mod_t[i] = prod_mod_except(c->q, L, (long)i, t);
    c->punct_q_mod_gamma[i] = prod_mod_except(c->q, L, (long)i, c->gamma);
    c->punct_q_mod_mtilde[i] = prod_mod_except(c->q, L, (long)i, c->mtilde);
    c->tgamma_mod_q[i] = mulmod(t % qi, c->gamma % qi, qi);
    c->mtilde_mod_q[i] = c->mtilde % qi;
    c->B_mod_q[i] = prod_mod_except(c->B, nB, -1, qi);
    for (size_t j = 0; j < c->nbsk; j++) c->punct_q_mod_bsk[j][i] = prod_mod_except(c->q, L, (long)i, c->bsk[j]);
    for (size_t j = 0; j < nB; j++) c->punct_B_mod_q[i][j] = prod_mod_except(c->B, nB, (long)j, qi);
  }
  c->neg_inv_q_mod_t = negmod(invmod(prod_mod_except(c->q, L, -1, t), t), t);
  c->neg_inv_q_mod_gamma = negmod(invmod(prod_mod_except(c->q, L, -1, c->gamma), c->gamma), c->gamma);
  c->inv_gamma_mod_t = invmod(c->gamma % t, t);
  c->neg_inv_q_mod_mtilde = negmod(invmod(prod_mod_except(c->q, L, -1, c->mtilde), c->mtilde), c->mtilde);
  for (size_t j = 0; j < c->nbsk; j++) {
    u64 pj = c->bsk[j];
    c->q_mod_bsk[j] = prod_mod_except(c->q, L, -1, pj);
    c->inv_q_mod_bsk[j] = invmod(c->q_mod_bsk[j], pj);
    c->inv_mtilde_mod_bsk[j] = invmod(c->mtilde % pj, pj);
  }
  for (size_t j = 0; j < nB; j++) {
    c->inv_punct_B[j] = invmod(prod_mod_except(c->B, nB, (long)j, c->B[j]), c->B[j]);
    c->punct_B_mod_msk[j] = prod_mod_except(c->B, nB, (long)j, c->msk);
  }
  c->inv_B_mod_msk = invmod(prod_mod_except(c->B, nB, -1, c->msk), c->msk);
  c->galois = calloc(N, sizeof(u64 *));
  return c;
}

void obfv_destroy(obfv_ctx *c) {
  if (!c) return;
  for (size_t i = 0; i < c->k; i++) ntt_tab_free(&c->nq[i]);
  for (size_t j = 0; j < c->nbsk; j++) ntt_tab_free(&c->nbskt[j]);
  ntt_tab_free(&c->nt);
  free(c->index_map); free(c->sk); free(c->pk); free(c->relin);
  if (c->galois) { for (size_t i = 0; i < c->N; i++) free(c->galois[i]); free(c->galois); }
  free(c);
}

size_t obfv_N(const obfv_ctx *c) { return c->N; }
size_t obfv_k(const obfv_ctx *c) { return c->k; }
size_t obfv_L(const obfv_ctx *c) { return c->L; }
u64 obfv_t(const obfv_ctx *c) { return c->t; }
void obfv_primes(const obfv_ctx *c, u64 *out) { memcpy(out, c->q, c->k * 8); }
size_t obfv_nbsk(const obfv_ctx *c) { return c->nbsk; }
void obfv_aux_primes(const obfv_ctx *c, u64 *msk, u64 *gamma, u64 *B) {
  *msk = c->msk; *gamma = c->gamma; memcpy(B, c->B, c->nB * 8);
}
u64 obfv_psi(const obfv_ctx *c, size_t i) { return c->nq[i].psi; }
u64 obfv_psi_t(const obfv_ctx *c) { return c->nt.psi; }

static const ntt_tab *tab_for(const obfv_ctx *c, size_t idx) {
  if (idx == (size_t)-1) return &c->nt;
  if (idx >= 1000) return &c->nbskt[idx - 1000];
  return &c->nq[idx];
}
void obfv_ntt_fwd(const obfv_ctx *c, size_t idx, u64 *limb) { ntt_fwd(tab_for(c, idx), limb, c->N); }
void obfv_ntt_inv(const obfv_ctx *c, size_t idx, u64 *limb) { ntt_inv(tab_for(c, idx), limb, c->N); }

/* ------------------------------------------------------------------ sampler (our own spec; SEAL's
 * Blake2xb/SHAKE stream is randomly seeded and cannot be matched, only the distributions are SEAL's:
 * sample_poly_ternary, sample_poly_cbd (SEAL 3.6 default noise), sample_poly_uniform — util/rlwe.cpp).
 * Randomness = the ChaCha20 key stream (20 rounds, RFC 8439 state layout) under a 256-bit key:
 *   state = "expand 32-byte k" | key[8] | block counter | nonce (domain | b << 4, a lo, a hi)
 *   word idx of stream (domain, a, b) = 32-bit words 2*(idx & 7), 2*(idx & 7) + 1 of block idx >> 3.
 * A 64-bit seed (tests) expands to the key (seed lo, seed hi, "abc-", "b200", 0, 0, 0, 0). */
typedef struct { u32 key[8]; u32 n0, n1, n2; u32 have; u32 ctr; u64 w[8]; } rng_stream_t;
static void rng_key_from_seed(u64 seed, u32 key[8]) {
  memset(key, 0, 32);
  key[0] = (u32)seed; key[1] = (u32)(seed >> 32); key[2] = 0x2d636261u; key[3] = 0x30303262u;
}
static rng_stream_t rng_stream(const u32 key[8], u64 domain, u64 a, u64 b) {
  rng_stream_t s;
  memcpy(s.key, key, 32);
  s.n0 = (u32)domain | ((u32)b << 4); s.n1 = (u32)a; s.n2 = (u32)(a >> 32);
  s.have = 0; s.ctr = 0;
  return s;
}
static inline u32 rotl32(u32 x, int r) { return (x << r) | (x >> (32 - r)); }
#define QR(a, b, c, d) \
  a += b; d ^= a; d = rotl32(d, 16); c += d; b ^= c; b = rotl32(b, 12); \
  a += b; d ^= a; d = rotl32(d, 8);  c += d; b ^= c; b = rotl32(b, 7);
static void chacha20_block(rng_stream_t *s, u32 counter) {
  u32 in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u}, x[16];
  memcpy(in + 4, s->key, 32);
  in[12] = counter; in[13] = s->n0; in[14] = s->n1; in[15] = s->n2;
  memcpy(x, in, sizeof x);
  for (int r = 0; r < 10; r++) {
    QR(x[0], x[4], x[8], x[12]) QR(x[1], x[5], x[9], x[13]) QR(x[2], x[6], x[10], x[14]) QR(x[3], x[7], x[11], x[15])
    QR(x[0], x[5], x[10], x[15]) QR(x[1], x[6], x[11], x[12]) QR(x[2], x[7], x[8], x[13]) QR(x[3], x[4], x[9], x[14])
  }
  for (int i = 0; i < 8; i++) s->w[i] = (u64)(x[2 * i] + in[2 * i]) | (u64)(x[2 * i + 1] + in[2 * i + 1]) << 32;
  s->ctr = counter; s->have = 1;
}
#undef QR
static u64 rng_word(rng_stream_t *s, u64 idx) {
  const u32 ctr = (u32)(idx >> 3);
  if (!s->have || s->ctr != ctr) chacha20_block(s, ctr);
  return s->w[idx & 7];
}
/* test hook: word idx of stream (domain, a, b) under the key a seed expands to; with key32 != NULL under that key */
u64 obfv_rng(u64 seed, const unsigned char *key32, u64 domain, u64 a, u64 b, u64 idx) {
  u32 key[8];
  if (key32) memcpy(key, key32, 32); else rng_key_from_seed(seed, key);
  rng_stream_t s = rng_stream(key, domain, a, b);
  return rng_word(&s, idx);
}

enum { DOM_SK = 1, DOM_PK = 2, DOM_KSK = 3, DOM_ENC = 4 };

/* ternary: {0,1,2} -> {-1,0,+1}; returns signed */
static inline int sample_ternary(rng_stream_t *h, u64 idx) {
  u64 r = rng_word(h, idx);
  return (int)(((r >> 32) * 3) >> 32) - 1;
}
/* centred binomial, 21 - 21 bits (SEAL sample_poly_cbd) */
static inline int sample_cbd(rng_stream_t *h, u64 idx) {
  u64 r = rng_word(h, idx);
  return __builtin_popcountll(r & 0x1fffffULL) - __builtin_popcountll((r >> 21) & 0x1fffffULL);
}
/* SEAL sample_poly_uniform: rejection above the largest multiple of q; retry j of coefficient idx is word idx + j * 2^24 */
static inline u64 sample_uniform(rng_stream_t *h, u64 idx, u64 q) {
  const u64 max_random = ~(u64)0, max_multiple = max_random - (max_random % q) - 1;
  for (u64 attempt = 0;; attempt++) {
    u64 r = rng_word(h, idx | (attempt << 24));
    if (r < max_multiple) return r % q;
  }
}
static inline u64 small_to_mod(int v, u64 q) { return v < 0 ? q - (u64)(-v) : (u64)v; }

/* ------------------------------------------------------------------ BatchEncoder (batchencoder.cpp)
 * reached from SealCiphertextFactory.cpp:127-132 (encode) and :151 (decode) */
void obfv_encode(const obfv_ctx *c, const int64_t *slots, u64 *plain) {
  for (size_t i = 0; i < c->N; i++) {
    int64_t v = slots[i];
    plain[c->index_map[i]] = v < 0 ? c->t + (u64)v : (u64)v;
  }
  ntt_inv(&c->nt, plain, c->N);
}
void obfv_decode(const obfv_ctx *c, const u64 *plain, int64_t *slots) {
  u64 *tmp = malloc(c->N * 8);
  memcpy(tmp, plain, c->N * 8);
  ntt_fwd(&c->nt, tmp, c->N);
  const u64 half = c->t >> 1;
  for (size_t i = 0; i < c->N; i++) {
    u64 v = tmp[c->index_map[i]];
    slots[i] = v > half ? (int64_t)v - (int64_t)c->t : (int64_t)v;
  }
  free(tmp);
}

/* ------------------------------------------------------------------ key generation (keygenerator.cpp, rlwe.cpp)
 * reached from SealCiphertextFactory.cpp:89-93 */

/* SEAL encrypt_zero_symmetric, NTT form, key level: c1 = a (uniform, taken as NTT form),
 * c0 = -(a*s + e).  out = [2][k][N].  a_stream/e_stream select the sampler streams. */
static void encrypt_zero_symmetric(const obfv_ctx *c, u64 dom, u64 a_id, u64 b_base, u64 *out) {
  const size_t N = c->N, k = c->k;
  u64 *c0 = out, *c1 = out + k * N;
  u64 *e = malloc(N * 8);
  rng_stream_t he = rng_stream(c->rng_key, dom, a_id, (b_base << 2) | 1);
  for (size_t i = 0; i < k; i++) {
    const u64 q = c->q[i];
    rng_stream_t ha = rng_stream(c->rng_key, dom, a_id, ((b_base | i) << 2) | 0);
    for (size_t j = 0; j < N; j++) c1[i * N + j] = sample_uniform(&ha, j, q);
    for (size_t j = 0; j < N; j++) e[j] = small_to_mod(sample_cbd(&he, j), q);
    ntt_fwd(&c->nq[i], e, N);
    for (size_t j = 0; j < N; j++) {
      u64 v = addmod(mulmod(c->sk[i * N + j], c1[i * N + j], q), e[j], q);
      c0[i * N + j] = negmod(v, q);
    }
  }
  free(e);
}

/* SEAL KeyGenerator::generate_one_kswitch_key: new_key [L][N] NTT form -> dst [L][2][k][N] */
static void gen_kswitch_key(const obfv_ctx *c, const u64 *new_key, u64 key_id, u64 *dst) {
  const size_t N = c->N, k = c->k, L = c->L;
  for (size_t J = 0; J < L; J++) {
    u64 *kj = dst + J * 2 * k * N;
    encrypt_zero_symmetric(c, DOM_KSK, key_id, (u64)J << 8, kj);
    const u64 q = c->q[J], factor = c->p % q;
    for (size_t j = 0; j < N; j++)
      kj[J * N + j] = addmod(kj[J * N + j], mulmod(new_key[J * N + j], factor, q), q);
  }
}

/* GaloisTool::generate_table_ntt + apply_galois_ntt (galois.cpp) */
static void apply_galois_ntt(const obfv_ctx *c, const u64 *in, u32 elt, u64 *out) {
  const size_t N = c->N;
  for (size_t i = 0; i < N; i++) {
    u32 rev = bitrev((u32)(i + N), c->logN + 1);
    u64 idx = (((u64)elt * rev) >> 1) & (N - 1);
    out[i] = in[bitrev((u32)idx, c->logN)];
  }
}

/* GaloisTool::get_elts_all */
size_t obfv_galois_elts(const obfv_ctx *c, u32 *out, size_t cap) {
  const u64 m = 2 * (u64)c->N;
  size_t n = 0;
  u64 pos = 3, neg = invmod(3, m);
  if (n < cap) out[n] = (u32)(m - 1);
  n++;
  for (int i = 0; i < c->logN - 1; i++) {
    if (n < cap) out[n] = (u32)pos;
    n++;
    pos = (pos * pos) & (m - 1);
    if (n < cap) out[n] = (u32)neg;
    n++;
    neg = (neg * neg) & (m - 1);
  }
  return n;
}

/* GaloisTool::get_elt_from_step */
u32 obfv_elt_from_step(const obfv_ctx *c, int step) {
  const u32 n = (u32)c->N, m = 2 * n;
  if (step == 0) return m - 1;
  u32 pos = (u32)(step < 0 ? -step : step);
  if (pos >= (n >> 1)) return 0; /* SEAL throws invalid_argument("step count too large") */
  u32 e = step < 0 ? (n >> 1) - pos : pos;
  u64 elt = 1;
  for (u32 i = 0; i < e; i++) elt = (elt * 3) & (m - 1);
  return (u32)elt;
}

static void keygen_impl(obfv_ctx *c, u64 seed, const u32 *elts, size_t ne) {
  const size_t N = c->N, k = c->k, L = c->L;
  c->seed = seed;
  if (!c->rng_key_set) rng_key_from_seed(seed, c->rng_key);
  free(c->sk); free(c->pk); free(c->relin);
  for (size_t i = 0; i < N; i++) { free(c->galois[i]); c->galois[i] = NULL; }
  /* secret key: ternary, NTT form at key level (KeyGenerator::generate_sk) */
  c->sk = malloc(k * N * 8);
  rng_stream_t hs = rng_stream(c->rng_key, DOM_SK, 0, 0);
  for (size_t i = 0; i < k; i++) {
    for (size_t j = 0; j < N; j++) c->sk[i * N + j] = small_to_mod(sample_ternary(&hs, j), c->q[i]);
    ntt_fwd(&c->nq[i], c->sk + i * N, N);
  }
  /* public key (generate_pk) */
  c->pk = malloc(2 * k * N * 8);
  encrypt_zero_symmetric(c, DOM_PK, 0, 0, c->pk);
  /* relinearisation key: new key = s^2 (create_relin_keys -> generate_kswitch_keys) */
  u64 *nk = malloc(k * N * 8);
  for (size_t i = 0; i < k; i++)
    for (size_t j = 0; j < N; j++) nk[i * N + j] = mulmod(c->sk[i * N + j], c->sk[i * N + j], c->q[i]);
  c->relin = malloc(L * 2 * k * N * 8);
  gen_kswitch_key(c, nk, 0, c->relin);
  /* Galois keys (create_galois_keys): the default element set, or the caller's list */
  for (size_t e = 0; e < ne; e++) {
    u32 elt = elts[e], idx = (elt - 1) >> 1;
    if (c->galois[idx]) continue;
    for (size_t i = 0; i < k; i++) apply_galois_ntt(c, c->sk + i * N, elt, nk + i * N);
    c->galois[idx] = malloc(L * 2 * k * N * 8);
    gen_kswitch_key(c, nk, elt, c->galois[idx]);
  }
  free(nk);
}
/* the sampler key itself (32 bytes) instead of a 64-bit seed: mirrors abc_set_rng_key */
void obfv_set_rng_key(obfv_ctx *c, const unsigned char *key32) { memcpy(c->rng_key, key32, 32); c->rng_key_set = 1; }
void obfv_keygen(obfv_ctx *c, u64 seed) {
  u32 elts[64]; size_t ne = obfv_galois_elts(c, elts, 64);
  keygen_impl(c, seed, elts, ne);
}
void obfv_keygen_select(obfv_ctx *c, u64 seed, const u32 *elts, size_t ne) { keygen_impl(c, seed, elts, ne); }
const u64 *obfv_secret_key(const obfv_ctx *c) { return c->sk; }
const u64 *obfv_public_key(const obfv_ctx *c) { return c->pk; }
const u64 *obfv_relin_key(const obfv_ctx *c) { return c->relin; }
const u64 *obfv_galois_key(const obfv_ctx *c, u32 elt) {
  if (!(elt & 1) || elt >= 2 * c->N) return NULL;
  return c->galois[(elt - 1) >> 1];
}

/* ------------------------------------------------------------------ plain scaling (scalingvariant.cpp)
 * multiply_add/sub_plain_with_scaling_variant; reached from SealCiphertext.cpp:134,145,175,184 and encryption */
static void scale_plain_addsub(const obfv_ctx *c, const u64 *plain, u64 *c0, int sub) {
  const size_t N = c->N, L = c->L;
  for (size_t j = 0; j < N; j++) {
    u128 num = (u128)plain[j] * c->q_mod_t + c->t_half_up;
    u64 fix = (u64)(num / c->t);
    for (size_t i = 0; i < L; i++) {
      const u64 q = c->q[i];
      u64 s = (u64)(((u128)plain[j] * c->delta_mod_q[i] + fix) % q);
      c0[i * N + j] = sub ? submod(c0[i * N + j], s, q) : addmod(c0[i * N + j], s, q);
    }
  }
}

/* ------------------------------------------------------------------ encryption (encryptor.cpp, rlwe.cpp)
 * Encryptor::encrypt -> encrypt_zero_asymmetric at key level, divide_and_round_q_last_inplace,
 * + scaled plaintext.  Reached from SealCiphertextFactory.cpp:12 */
void obfv_encrypt(const obfv_ctx *c, const u64 *plain, u64 nonce, u64 *ct) {
  const size_t N = c->N, k = c->k, L = c->L;
  u64 *u = malloc(k * N * 8), *tmp = malloc(2 * k * N * 8);
  rng_stream_t hu = rng_stream(c->rng_key, DOM_ENC, nonce, 0);
  for (size_t i = 0; i < k; i++) {
    for (size_t j = 0; j < N; j++) u[i * N + j] = small_to_mod(sample_ternary(&hu, j), c->q[i]);
    ntt_fwd(&c->nq[i], u + i * N, N);
  }
  for (size_t pidx = 0; pidx < 2; pidx++) {
    rng_stream_t he = rng_stream(c->rng_key, DOM_ENC, nonce, 1 + pidx);
    for (size_t i = 0; i < k; i++) {
      const u64 q = c->q[i];
      u64 *d = tmp + (pidx * k + i) * N;
      const u64 *pkp = c->pk + (pidx * k + i) * N;
      for (size_t j = 0; j < N; j++) d[j] = bmul(u[i * N + j], pkp[j], &c->bq[i]);
      ntt_inv(&c->nq[i], d, N);
      for (size_t j = 0; j < N; j++) d[j] = addmod(d[j], small_to_mod(sample_cbd(&he, j), q), q);
    }
    /* RNSTool::divide_and_round_q_last_inplace */
    u64 *last = tmp + (pidx * k + L) * N;
    for (size_t j = 0; j < N; j++) last[j] = addmod(last[j], c->p_half, c->p);
    for (size_t i = 0; i < L; i++) {
      const u64 q = c->q[i];
      u64 *d = tmp + (pidx * k + i) * N, *o = ct + (pidx * L + i) * N;
      for (size_t j = 0; j < N; j++) {
        u64 tl = submod(bred64(last[j], &c->bq[i]), c->p_half_mod_q[i], q);
        o[j] = bmul(submod(d[j], tl, q), c->inv_p_mod_q[i], &c->bq[i]);
      }
    }
  }
  scale_plain_addsub(c, plain, ct, 0);
  free(u); free(tmp);
}

/* FastBConv (BaseConverter::fast_convert_array) on one coefficient:
 * out = sum_i z_i * punct_i mod m, with z_i = x_i * inv_punct_i mod base_i supplied by the caller */
static inline u64 dot_mod(const u64 *z, const u64 *punct, size_t n, const bmod *m) {
  /* SEAL dot_product_mod: lazy 128-bit accumulation (products < 2^122, n < 64), one Barrett reduction */
  u128 acc = 0;
  for (size_t i = 0; i < n; i++) acc += (u128)z[i] * punct[i];
  return bred128(acc, m);
}

/* ------------------------------------------------------------------ decryption (decryptor.cpp, rns.cpp)
 * Decryptor::bfv_decrypt -> dot_product_ct_sk_array, RNSTool::decrypt_scale_and_round.
 * Reached from SealCiphertextFactory.cpp:150 */
void obfv_decrypt(const obfv_ctx *c, const u64 *ct, size_t size, u64 *plain) {
  const size_t N = c->N, L = c->L, k = c->k;
  u64 *x = malloc(L * N * 8), *tmp = malloc(N * 8), *spow = malloc(N * 8);
  for (size_t i = 0; i < L; i++) {
    const u64 q = c->q[i];
    memset(x + i * N, 0, N * 8);
    memcpy(spow, c->sk + i * N, N * 8);
    for (size_t pidx = 1; pidx < size; pidx++) {
      memcpy(tmp, ct + (pidx * L + i) * N, N * 8);
      ntt_fwd(&c->nq[i], tmp, N);
      for (size_t j = 0; j < N; j++) x[i * N + j] = addmod(x[i * N + j], bmul(tmp[j], spow[j], &c->bq[i]), q);
      if (pidx + 1 < size) for (size_t j = 0; j < N; j++) spow[j] = bmul(spow[j], c->sk[i * N + j], &c->bq[i]);
    }
    ntt_inv(&c->nq[i], x + i * N, N);
    for (size_t j = 0; j < N; j++) x[i * N + j] = addmod(x[i * N + j], ct[i * N + j], q);
  }
  (void)k;
  const u64 t = c->t, g = c->gamma, g_half = g >> 1;
  u64 z[MAXK];
  for (size_t j = 0; j < N; j++) {
    for (size_t i = 0; i < L; i++)
      z[i] = bmul(bmul(x[i * N + j], c->tgamma_mod_q[i], &c->bq[i]), c->inv_punct_q[i], &c->bq[i]);
    u64 yt = bmul(dot_mod(z, c->punct_q_mod_t, L, &c->bt), c->neg_inv_q_mod_t, &c->bt);
    u64 yg = bmul(dot_mod(z, c->punct_q_mod_gamma, L, &c->bgamma), c->neg_inv_q_mod_gamma, &c->bgamma);
    u64 d;
    if (yg > g_half) d = addmod(yt, bred64(g - yg, &c->bt), t);
    else d = submod(yt, bred64(yg, &c->bt), t);
    plain[j] = d ? bmul(d, c->inv_gamma_mod_t, &c->bt) : 0;
  }
  free(x); free(tmp); free(spow);
}

/* ------------------------------------------------------------------ noise budget (decryptor.cpp)
 * Decryptor::invariant_noise_budget, reached from SealCiphertext::noiseBits (SealCiphertext.cpp:80-83):
 * noise = dot_product_ct_sk_array(ct); noise *= t (mod q_i); CRT-compose (RNSBase::compose_array);
 * norm = poly_infty_norm_coeffmod(noise, Q); budget = max(0, bit_count(Q) - bit_count(norm) - 1). */
static int mp_bits(const u64 *w, size_t n) {
  for (size_t i = n; i-- > 0;)
    if (w[i]) { int b = 0; for (u64 v = w[i]; v; v >>= 1) b++; return (int)(64 * i) + b; }
  return 0;
}
static int mp_ge(const u64 *a, const u64 *b, size_t n) {
  for (size_t i = n; i-- > 0;) if (a[i] != b[i]) return a[i] > b[i];
  return 1;
}
static void mp_sub(u64 *a, const u64 *b, size_t n) { /* a -= b */
  u64 borrow = 0;
  for (size_t i = 0; i < n; i++) { u128 d = (u128)a[i] - b[i] - borrow; a[i] = (u64)d; borrow = (u64)(d >> 64) & 1; }
}
int obfv_noise_budget(const obfv_ctx *c, const u64 *ct, size_t size) {
  const size_t N = c->N, L = c->L;
  u64 *x = malloc(L * N * 8), *tmp = malloc(N * 8), *spow = malloc(N * 8);
  for (size_t i = 0; i < L; i++) {
    const u64 q = c->q[i];
    memset(x + i * N, 0, N * 8);
    memcpy(spow, c->sk + i * N, N * 8);
    for (size_t pidx = 1; pidx < size; pidx++) {
      memcpy(tmp, ct + (pidx * L + i) * N, N * 8);
      ntt_fwd(&c->nq[i], tmp, N);
      for (size_t j = 0; j < N; j++) x[i * N + j] = addmod(x[i * N + j], bmul(tmp[j], spow[j], &c->bq[i]), q);
      if (pidx + 1 < size) for (size_t j = 0; j < N; j++) spow[j] = bmul(spow[j], c->sk[i * N + j], &c->bq[i]);
    }
    ntt_inv(&c->nq[i], x + i * N, N);
    for (size_t j = 0; j < N; j++) x[i * N + j] = bmul(addmod(x[i * N + j], ct[i * N + j], q), bred64(c->t, &c->bq[i]), &c->bq[i]);
  }
  /* multi-precision constants: punctured products Q/q_i, Q, (Q+1)/2 (L words, little endian) */
  u64 P[MAXK][MAXK + 1], Q[MAXK + 1], H[MAXK + 1], acc[MAXK + 2];
  for (size_t i = 0; i <= L; i++) {
    u64 *w = i < L ? P[i] : Q;
    memset(w, 0, (MAXK + 1) * 8); w[0] = 1;
    for (size_t j = 0; j < L; j++) if (j != i) {
      u64 carry = 0;
      for (size_t a = 0; a < L; a++) { u128 t = (u128)w[a] * c->q[j] + carry; w[a] = (u64)t; carry = (u64)(t >> 64); }
    }
  }
  { u128 cy = 1; for (size_t a = 0; a <= L; a++) { cy += Q[a]; H[a] = (u64)cy; cy >>= 64; }
    for (size_t a = 0; a <= L; a++) H[a] = (H[a] >> 1) | (a < L ? H[a + 1] << 63 : 0); }
  int maxbits = 0;
  for (size_t j = 0; j < N; j++) {
    memset(acc, 0, sizeof acc);
    for (size_t i = 0; i < L; i++) {
      const u64 y = bmul(x[i * N + j], c->inv_punct_q[i], &c->bq[i]);
      u64 carry = 0;
      for (size_t a = 0; a < L; a++) { u128 t = (u128)y * P[i][a] + acc[a] + carry; acc[a] = (u64)t; carry = (u64)(t >> 64); }
      acc[L] += carry;
    }
    while (mp_ge(acc, Q, L + 1)) mp_sub(acc, Q, L + 1);
    if (mp_ge(acc, H, L + 1)) { u64 r[MAXK + 1]; memcpy(r, Q, (L + 1) * 8); mp_sub(r, acc, L + 1); memcpy(acc, r, (L + 1) * 8); }
    const int b = mp_bits(acc, L + 1);
    if (b > maxbits) maxbits = b;
  }
  free(x); free(tmp); free(spow);
  const int d = mp_bits(Q, L + 1) - maxbits - 1;
  return d > 0 ? d : 0;
}

/* ------------------------------------------------------------------ add / sub / negate (evaluator.cpp)
 * reached from SealCiphertext.cpp:92,98,114,118,157,193 */
void obfv_add(const obfv_ctx *c, const u64 *a, const u64 *b, u64 *out) {
  for (size_t p = 0; p < 2; p++) for (size_t i = 0; i < c->L; i++) {
    const u64 q = c->q[i]; const size_t o = (p * c->L + i) * c->N;
    for (size_t j = 0; j < c->N; j++) out[o + j] = addmod(a[o + j], b[o + j], q);
  }
}
void obfv_sub(const obfv_ctx *c, const u64 *a, const u64 *b, u64 *out) {
  for (size_t p = 0; p < 2; p++) for (size_t i = 0; i < c->L; i++) {
    const u64 q = c->q[i]; const size_t o = (p * c->L + i) * c->N;
    for (size_t j = 0; j < c->N; j++) out[o + j] = submod(a[o + j], b[o + j], q);
  }
}
void obfv_negate(const obfv_ctx *c, const u64 *a, u64 *out) {
  for (size_t p = 0; p < 2; p++) for (size_t i = 0; i < c->L; i++) {
    const u64 q = c->q[i]; const size_t o = (p * c->L + i) * c->N;
    for (size_t j = 0; j < c->N; j++) out[o + j] = negmod(a[o + j], q);
  }
}
void obfv_add_plain(const obfv_ctx *c, const u64 *a, const u64 *plain, u64 *out) {
  if (out != a) memcpy(out, a, 2 * c->L * c->N * 8);
  scale_plain_addsub(c, plain, out, 0);
}
void obfv_sub_plain(const obfv_ctx *c, const u64 *a, const u64 *plain, u64 *out) {
  if (out != a) memcpy(out, a, 2 * c->L * c->N * 8);
  scale_plain_addsub(c, plain, out, 1);
}

/* Evaluator::multiply_plain_normal: centred lift per limb, NTT, dyadic with both polys, INTT.
 * (SEAL's mono-nomial fast path yields the identical result.)  SealCiphertext.cpp:159,196 */
void obfv_multiply_plain(const obfv_ctx *c, const u64 *a, const u64 *plain, u64 *out) {
  const size_t N = c->N, L = c->L;
  u64 *pl = malloc(N * 8), *tmp = malloc(N * 8);
  for (size_t i = 0; i < L; i++) {
    const u64 q = c->q[i], inc = q - c->t;
    for (size_t j = 0; j < N; j++) pl[j] = plain[j] >= c->t_half_up ? plain[j] + inc : plain[j];
    ntt_fwd(&c->nq[i], pl, N);
    for (size_t p = 0; p < 2; p++) {
      memcpy(tmp, a + (p * L + i) * N, N * 8);
      ntt_fwd(&c->nq[i], tmp, N);
      for (size_t j = 0; j < N; j++) tmp[j] = bmul(tmp[j], pl[j], &c->bq[i]);
      ntt_inv(&c->nq[i], tmp, N);
      memcpy(out + (p * L + i) * N, tmp, N * 8);
    }
  }
  free(pl); free(tmp);
}

/* ------------------------------------------------------------------ BEHZ multiply (evaluator.cpp bfv_multiply, rns.cpp) */
/* fastbconv_m_tilde + sm_mrq on one polynomial (coefficient form in and out) */
void obfv_behz_lift(const obfv_ctx *c, const u64 *x, u64 *out) {
  const size_t N = c->N, L = c->L, nb = c->nbsk;
  const u64 mt = c->mtilde, mt_half = mt >> 1;
  u64 z[MAXK];
  for (size_t j = 0; j < N; j++) {
    for (size_t i = 0; i < L; i++)
      z[i] = bmul(bmul(x[i * N + j], c->mtilde_mod_q[i], &c->bq[i]), c->inv_punct_q[i], &c->bq[i]);
    u64 xm = 0;
    for (size_t i = 0; i < L; i++) xm += z[i] * c->punct_q_mod_mtilde[i];
    xm &= mt - 1;
    u64 r = (xm * c->neg_inv_q_mod_mtilde) & (mt - 1);
    for (size_t b = 0; b < nb; b++) {
      const u64 pm = c->bsk[b];
      u64 xb = dot_mod(z, c->punct_q_mod_bsk[b], L, &c->bbsk[b]);
      u64 rc = r >= mt_half ? r + (pm - mt) : r;
      u64 v = bred128((u128)rc * c->q_mod_bsk[b] + xb, &c->bbsk[b]);
      out[b * N + j] = bmul(v, c->inv_mtilde_mod_bsk[b], &c->bbsk[b]);
    }
  }
}
/* multiply by t, fast_floor, fastbconv_sk on one polynomial */
void obfv_behz_scale(const obfv_ctx *c, const u64 *in_q, const u64 *in_bsk, u64 *out) {
  const size_t N = c->N, L = c->L, nb = c->nbsk, nB = c->nB;
  const u64 t = c->t, msk = c->msk, msk_half = msk >> 1;
  u64 z[MAXK], y[MAXK], zb[MAXK];
  for (size_t j = 0; j < N; j++) {
    for (size_t i = 0; i < L; i++)
      z[i] = bmul(bmul(in_q[i * N + j], t % c->q[i], &c->bq[i]), c->inv_punct_q[i], &c->bq[i]);
    for (size_t b = 0; b < nb; b++) {
      const u64 pm = c->bsk[b];
      u64 conv = dot_mod(z, c->punct_q_mod_bsk[b], L, &c->bbsk[b]);
      u64 xb = bmul(in_bsk[b * N + j], t, &c->bbsk[b]);
      y[b] = bmul(submod(xb, conv, pm), c->inv_q_mod_bsk[b], &c->bbsk[b]);
    }
    for (size_t b = 0; b < nB; b++) zb[b] = bmul(y[b], c->inv_punct_B[b], &c->bbsk[b]);
    u64 conv_sk = dot_mod(zb, c->punct_B_mod_msk, nB, &c->bbsk[nB]);
    u64 alpha = bmul(submod(conv_sk, y[nB], msk), c->inv_B_mod_msk, &c->bbsk[nB]);
    for (size_t i = 0; i < L; i++) {
      const u64 q = c->q[i];
      u64 conv = dot_mod(zb, c->punct_B_mod_q[i], nB, &c->bq[i]);
      if (alpha > msk_half) out[i * N + j] = addmod(conv, bmul(bred64(msk - alpha, &c->bq[i]), c->B_mod_q[i], &c->bq[i]), q);
      else out[i * N + j] = submod(conv, bmul(bred64(alpha, &c->bq[i]), c->B_mod_q[i], &c->bq[i]), q);
    }
  }
}

/* Evaluator::bfv_multiply for size-2 x size-2 -> size-3.  SealCiphertext.cpp:104,122 */
void obfv_multiply(const obfv_ctx *c, const u64 *a, const u64 *b, u64 *out3) {
  const size_t N = c->N, L = c->L, nb = c->nbsk, W = L + nb;
  /* per input poly: [L q-limbs | nb bsk-limbs], NTT form */
  u64 *in = malloc(4 * W * N * 8), *prod = malloc(3 * W * N * 8);
  const u64 *src[4] = {a, a + L * N, b, b + L * N};
  for (size_t p = 0; p < 4; p++) {
    u64 *d = in + p * W * N;
    memcpy(d, src[p], L * N * 8);
    obfv_behz_lift(c, src[p], d + L * N);
    for (size_t i = 0; i < L; i++) ntt_fwd(&c->nq[i], d + i * N, N);
    for (size_t j = 0; j < nb; j++) ntt_fwd(&c->nbskt[j], d + (L + j) * N, N);
  }
  for (size_t w = 0; w < W; w++) {
    const bmod *bm = w < L ? &c->bq[w] : &c->bbsk[w - L];
    const u64 *a0 = in + (0 * W + w) * N, *a1 = in + (1 * W + w) * N;
    const u64 *b0 = in + (2 * W + w) * N, *b1 = in + (3 * W + w) * N;
    u64 *d0 = prod + (0 * W + w) * N, *d1 = prod + (1 * W + w) * N, *d2 = prod + (2 * W + w) * N;
    for (size_t j = 0; j < N; j++) {
      d0[j] = bmul(a0[j], b0[j], bm);
      d1[j] = bred128((u128)a0[j] * b1[j] + (u128)a1[j] * b0[j], bm);
      d2[j] = bmul(a1[j], b1[j], bm);
    }
    const ntt_tab *T = w < L ? &c->nq[w] : &c->nbskt[w - L];
    ntt_inv(T, d0, N); ntt_inv(T, d1, N); ntt_inv(T, d2, N);
  }
  for (size_t p = 0; p < 3; p++)
    obfv_behz_scale(c, prod + p * W * N, prod + (p * W + L) * N, out3 + p * L * N);
  free(in); free(prod);
}

/* ------------------------------------------------------------------ key switching (evaluator.cpp switch_key_inplace) */
void obfv_switch_key(const obfv_ctx *c, u64 *ct, const u64 *target, const u64 *key) {
  const size_t N = c->N, L = c->L, k = c->k;
  u64 *acc = malloc(2 * k * N * 8), *tn = malloc(N * 8);
  memset(acc, 0, 2 * k * N * 8);
  for (size_t I = 0; I < k; I++) { /* I == L is the special prime */
    const u64 q = c->q[I];
    for (size_t J = 0; J < L; J++) {
      if (c->q[J] <= q) memcpy(tn, target + J * N, N * 8);
      else for (size_t j = 0; j < N; j++) tn[j] = bred64(target[J * N + j], &c->bq[I]);
      ntt_fwd(&c->nq[I], tn, N);
      for (size_t comp = 0; comp < 2; comp++) {
        const u64 *kk = key + ((J * 2 + comp) * k + I) * N;
        u64 *ac = acc + (comp * k + I) * N;
        for (size_t j = 0; j < N; j++) ac[j] = addmod(ac[j], bmul(tn[j], kk[j], &c->bq[I]), q);
      }
    }
  }
  for (size_t comp = 0; comp < 2; comp++) {
    u64 *last = acc + (comp * k + L) * N;
    ntt_inv(&c->nq[L], last, N);
    for (size_t j = 0; j < N; j++) last[j] = addmod(last[j], c->p_half, c->p);
    for (size_t i = 0; i < L; i++) {
      const u64 q = c->q[i];
      u64 *ai = acc + (comp * k + i) * N, *o = ct + (comp * L + i) * N;
      ntt_inv(&c->nq[i], ai, N);
      for (size_t j = 0; j < N; j++) {
        u64 tl = submod(bred64(last[j], &c->bq[i]), c->p_half_mod_q[i], q);
        o[j] = addmod(o[j], bmul(submod(ai[j], tl, q), c->inv_p_mod_q[i], &c->bq[i]), q);
      }
    }
  }
  free(acc); free(tn);
}

/* Evaluator::relinearize_internal, size 3 -> 2 with relin key index 0.  SealCiphertext.cpp:105,123 */
void obfv_relinearize(const obfv_ctx *c, const u64 *ct3, u64 *out2) {
  const size_t sz = 2 * c->L * c->N;
  if (out2 != ct3) memcpy(out2, ct3, sz * 8);
  obfv_switch_key(c, out2, ct3 + sz, c->relin);
}

/* GaloisTool::apply_galois on every limb of one polynomial set ([polys*L] limbs handled by caller) */
static void apply_galois_limb(const u64 *in, u32 elt, u64 q, size_t N, int logN, u64 *out) {
  u64 raw = 0;
  for (size_t i = 0; i < N; i++) {
    size_t idx = raw & (N - 1);
    u64 v = in[i];
    if ((raw >> logN) & 1) v = negmod(v, q);
    out[idx] = v;
    raw += elt;
  }
}
/* Evaluator::apply_galois_inplace (BFV branch).  out may not alias a. */
void obfv_apply_galois(const obfv_ctx *c, const u64 *a, u32 elt, u64 *out) {
  const size_t N = c->N, L = c->L;
  u64 *tmp = malloc(L * N * 8);
  for (size_t i = 0; i < L; i++) {
    apply_galois_limb(a + i * N, elt, c->q[i], N, c->logN, out + i * N);
    apply_galois_limb(a + (L + i) * N, elt, c->q[i], N, c->logN, tmp + i * N);
  }
  memset(out + L * N, 0, L * N * 8);
  obfv_switch_key(c, out, tmp, obfv_galois_key(c, elt));
  free(tmp);
}

/* util::naf */
static int naf(int value, int *out) {
  int n = 0, sign = value < 0; if (sign) value = -value;
  for (int i = 0; value; i++) {
    int zi = (value & 1) ? 2 - (value & 3) : 0;
    value = (value - zi) >> 1;
    if (zi) out[n++] = (sign ? -zi : zi) * (1 << i);
  }
  return n;
}

/* Evaluator::rotate_internal.  SealCiphertext.cpp:55,60 */
static int rotate_internal(const obfv_ctx *c, u64 *ct, int steps, int *nks) {
  if (steps == 0) return 0;
  const size_t sz = 2 * c->L * c->N;
  u32 elt = obfv_elt_from_step(c, steps);
  if (!elt) return -1;
  if (obfv_galois_key(c, elt)) {
    if (ct) {
      u64 *tmp = malloc(sz * 8);
      obfv_apply_galois(c, ct, elt, tmp);
      memcpy(ct, tmp, sz * 8);
      free(tmp);
    }
    if (nks) (*nks)++;
    return 0;
  }
  int steps_naf[40], n = naf(steps, steps_naf);
  if (n == 1) return -2; /* SEAL: invalid_argument("Galois key not present") */
  for (int i = 0; i < n; i++) {
    int s = steps_naf[i], as = s < 0 ? -s : s;
    if ((size_t)as != (c->N >> 1)) { int r = rotate_internal(c, ct, s, nks); if (r) return r; }
  }
  return 0;
}
int obfv_rotate_rows(const obfv_ctx *c, const u64 *a, int steps, u64 *out) {
  if (out != a) memcpy(out, a, 2 * c->L * c->N * 8);
  return rotate_internal(c, out, steps, NULL);
}
int obfv_rotate_keyswitch_count(const obfv_ctx *c, int steps) {
  int n = 0;
  /* uses only which keys exist; needs keygen */
  if (rotate_internal(c, NULL, steps, &n)) return -1;
  return n;
}
