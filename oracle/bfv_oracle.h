/*
 * bfv_oracle.h — CPU ORACLE (TEST INFRASTRUCTURE ONLY — never shipped, never on the product path).
 *
 * A plain-C restatement of the BFV arithmetic that ABC's SEAL backend reaches through
 *   /root/reference/src/runtime/SealCiphertext.cpp:52-202        (17 seal::Evaluator calls)
 *   /root/reference/src/runtime/SealCiphertextFactory.cpp:9-24,72-152 (context, keys, encode/encrypt/decrypt)
 * The arithmetic itself lives in Microsoft SEAL 3.6.5 (pinned: /root/reference/Docker/Dockerfile:9,
 * /root/reference/CMakeLists.txt:57), which is an un-vendored third-party dependency ABSENT from
 * /root/reference and from this image.  This file restates SEAL 3.6.5's published algorithms
 * (native/src/seal/{evaluator,encryptor,decryptor,keygenerator,batchencoder}.cpp,
 *  util/{rns,ntt,galois,scalingvariant,rlwe,numth}.cpp) — see SURVEY.md Appendix A.
 *
 * PARITY STATUS: slot-level parity is pinned against every known-answer vector the reference's own
 * tests hold for this path (tests/golden/abc_kats.json, taken from
 * /root/reference/test/runtime/SealCiphertextFactoryTest.cpp and RuntimeVisitorTest.cpp).
 * COEFFICIENT-level parity with real SEAL is "parity unpinned": the reference holds no ciphertext
 * fixtures, and SEAL cannot be built or run here.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.
 */
#ifndef BFV_ORACLE_H
#define BFV_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct obfv_ctx obfv_ctx;

/* primes==NULL -> SEAL CoeffModulus::BFVDefault(N) (k is ignored); t==0 -> PlainModulus::Batching(N,20).
 * Returns NULL on invalid parameters. */
obfv_ctx *obfv_create(size_t N, const uint64_t *primes, size_t k, uint64_t t);
/* test hook (not SEAL behaviour): BEHZ auxiliary base of aux_count + 1 primes of aux_bits bits, see bfv_oracle.c */
obfv_ctx *obfv_create_aux(size_t N, const uint64_t *primes, size_t k, uint64_t t, int aux_bits, size_t aux_count);
void obfv_destroy(obfv_ctx *c);

size_t obfv_N(const obfv_ctx *c);
size_t obfv_k(const obfv_ctx *c);          /* key-level prime count */
size_t obfv_L(const obfv_ctx *c);          /* data-level limb count = k-1 */
uint64_t obfv_t(const obfv_ctx *c);
void obfv_primes(const obfv_ctx *c, uint64_t *out_k);
size_t obfv_nbsk(const obfv_ctx *c);       /* |B|+1 */
void obfv_aux_primes(const obfv_ctx *c, uint64_t *msk, uint64_t *gamma, uint64_t *B /* |B| */);
uint64_t obfv_psi(const obfv_ctx *c, size_t prime_index); /* minimal primitive 2N-th root mod q_i */
uint64_t obfv_psi_t(const obfv_ctx *c);

/* SEAL rule get_primes(N, bits, count): descending scan from 2^bits-2N+1 in steps of 2N. returns #found */
size_t obfv_get_primes(size_t N, int bits, size_t count, uint64_t *out);

/* negacyclic NTT on one limb under key-level prime idx (idx==(size_t)-1: plain modulus t;
 * idx >= 1000: Bsk prime idx-1000). Canonical [0,q) in and out. */
void obfv_ntt_fwd(const obfv_ctx *c, size_t idx, uint64_t *limb);
void obfv_ntt_inv(const obfv_ctx *c, size_t idx, uint64_t *limb);

/* keys: deterministic from seed (our own counter-based sampler; see oracle/README.md). */
void obfv_keygen(obfv_ctx *c, uint64_t seed);
void obfv_keygen_select(obfv_ctx *c, uint64_t seed, const uint32_t *galois_elts, size_t n); /* explicit Galois set */
const uint64_t *obfv_secret_key(const obfv_ctx *c);            /* [k][N] NTT form */
const uint64_t *obfv_public_key(const obfv_ctx *c);            /* [2][k][N] NTT form */
const uint64_t *obfv_relin_key(const obfv_ctx *c);             /* [L][2][k][N] NTT form */
const uint64_t *obfv_galois_key(const obfv_ctx *c, uint32_t galois_elt); /* same layout, or NULL */
size_t obfv_galois_elts(const obfv_ctx *c, uint32_t *out, size_t cap);   /* default element set, in SEAL order */
uint32_t obfv_elt_from_step(const obfv_ctx *c, int step);

/* BatchEncoder; slots has N entries. */
void obfv_encode(const obfv_ctx *c, const int64_t *slots, uint64_t *plain);
void obfv_decode(const obfv_ctx *c, const uint64_t *plain, int64_t *slots);

/* ciphertexts: [size][L][N] coefficient form, canonical residues. */
void obfv_encrypt(const obfv_ctx *c, const uint64_t *plain, uint64_t nonce, uint64_t *ct2);
void obfv_decrypt(const obfv_ctx *c, const uint64_t *ct, size_t size, uint64_t *plain);
/* Decryptor::invariant_noise_budget (SealCiphertext::noiseBits, SealCiphertext.cpp:80-83) */
int obfv_noise_budget(const obfv_ctx *c, const uint64_t *ct, size_t size);
void obfv_add(const obfv_ctx *c, const uint64_t *a, const uint64_t *b, uint64_t *out);
void obfv_sub(const obfv_ctx *c, const uint64_t *a, const uint64_t *b, uint64_t *out);
void obfv_negate(const obfv_ctx *c, const uint64_t *a, uint64_t *out);
void obfv_add_plain(const obfv_ctx *c, const uint64_t *a, const uint64_t *plain, uint64_t *out);
void obfv_sub_plain(const obfv_ctx *c, const uint64_t *a, const uint64_t *plain, uint64_t *out);
void obfv_multiply_plain(const obfv_ctx *c, const uint64_t *a, const uint64_t *plain, uint64_t *out);
void obfv_multiply(const obfv_ctx *c, const uint64_t *a, const uint64_t *b, uint64_t *out3);
void obfv_relinearize(const obfv_ctx *c, const uint64_t *ct3, uint64_t *out2);
void obfv_apply_galois(const obfv_ctx *c, const uint64_t *a, uint32_t galois_elt, uint64_t *out);
/* returns 0 ok, -1 if |steps| >= N/2 (SEAL throws std::invalid_argument) */
int obfv_rotate_rows(const obfv_ctx *c, const uint64_t *a, int steps, uint64_t *out);
/* number of key switches rotate_rows(steps) performs (NAF weight when no direct key) */
int obfv_rotate_keyswitch_count(const obfv_ctx *c, int steps);

/* intermediate probes for kernel-level parity */
/* BEHZ step 1 on one polynomial: in [L][N] -> out_bsk [nbsk][N] (coefficient form, after SmMRq) */
void obfv_behz_lift(const obfv_ctx *c, const uint64_t *poly_q, uint64_t *out_bsk);
/* BEHZ steps 6-8 on one polynomial: in_q [L][N], in_bsk [nbsk][N] (coefficient form) -> out [L][N] */
void obfv_behz_scale(const obfv_ctx *c, const uint64_t *in_q, const uint64_t *in_bsk, uint64_t *out);
/* switch_key_inplace: ct2 [2][L][N] += keyswitch(target [L][N]) using key [L][2][k][N] */
void obfv_switch_key(const obfv_ctx *c, uint64_t *ct2, const uint64_t *target, const uint64_t *key);

/* sampler (shared spec with the CUDA library, documented in DESIGN.md) */
uint64_t obfv_rng(uint64_t seed, const unsigned char *key32, uint64_t domain, uint64_t a, uint64_t b, uint64_t idx);
/* the sampler's 32-byte ChaCha20 key itself instead of the expansion of a 64-bit seed (mirror of abc_set_rng_key) */
void obfv_set_rng_key(obfv_ctx *c, const unsigned char *key32);

#ifdef __cplusplus
}
#endif
#endif
