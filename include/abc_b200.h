/*
 * abc_b200.h — C ABI of libabc_b200.so, the B200 (sm_100a) BFV ciphertext backend for MarbleHE/ABC.
 *
 * This is the drop-in boundary for the path ABC reaches through
 *   include/ast_opt/runtime/AbstractCiphertextFactory.h:19-49  and
 *   include/ast_opt/runtime/AbstractCiphertext.h:27-98
 * (paths relative to /root/reference).  Each entry point names the reference interface it replaces.
 * Plain pointers and sizes only; no C++ or torch types.  Every call returns an abc_status; on
 * failure abc_last_error() holds a message (the C++ wrapper rethrows it as std::runtime_error,
 * the reference's error convention: src/runtime/SealCiphertext.cpp:40,137,242).
 *
 * Execution model: a context owns ONE device and ONE CUDA stream.  Ciphertext ops only enqueue
 * kernels and return; abc_decrypt_decode, the export calls and abc_sync are the sync points.
 * There is no CPU fallback: without a usable sm_100-class device abc_ctx_create fails.
 *
 * Batch: a context is created for `batch` independent program instances.  Every abc_ct handle
 * holds `batch` ciphertexts, laid out [batch][poly][limb][N] (u64, coefficient form, canonical
 * residues — per instance this is exactly seal::Ciphertext's layout), and every op runs on all
 * instances in one launch sequence with shared keys.
 */
#ifndef ABC_B200_H
#define ABC_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef int abc_status; /* 0 = ok */
enum { ABC_OK = 0, ABC_ERR_PARAM = 1, ABC_ERR_CUDA = 2, ABC_ERR_STATE = 3, ABC_ERR_UNSUPPORTED = 4 };

typedef struct abc_ctx abc_ctx;
typedef struct abc_ct abc_ct;

typedef struct abc_params {
  uint32_t poly_degree;     /* N: 4096..65536, power of two */
  uint32_t n_primes;        /* k, incl. the special prime (last); 0 -> SEAL CoeffModulus::BFVDefault(N) */
  const uint64_t *primes;   /* k primes = 1 (mod 2N), each < 2^60; ignored when n_primes == 0 */
  uint64_t plain_modulus;   /* t; 0 -> SEAL PlainModulus::Batching(N, 20) */
  int32_t device;           /* CUDA device ordinal */
  uint32_t batch;           /* independent instances per handle (>= 1) */
  uint64_t seed;            /* 0 = the sampler's 256-bit key is drawn from the OS generator (getrandom), like the
                             * reference's randomly seeded SEAL PRNG.  A non-zero seed is EXPANDED to the key, which
                             * makes the keys reproducible by anyone who knows those 64 bits: tests, and the same keys
                             * on every GPU of a job when abc_set_rng_key is not used.  Encryption randomness is
                             * additionally salted per context from the OS (see abc_set_encrypt_nonce).  The sampler
                             * is the ChaCha20 key stream. */
} abc_params;

/* --- context: replaces SealCiphertextFactory::setupSealContext (src/runtime/SealCiphertextFactory.cpp:72-100)
 * up to, not including, key generation. */
abc_status abc_ctx_create(const abc_params *params, abc_ctx **out);
void abc_ctx_destroy(abc_ctx *ctx);
const char *abc_last_error(const abc_ctx *ctx); /* ctx may be NULL: error of the failed abc_ctx_create */
abc_status abc_sync(abc_ctx *ctx);
/* Host staging for the end-to-end path: slots travel fastest from / to page-locked memory (and only then
 * asynchronously).  abc_host_alloc returns page-locked memory, abc_host_register page-locks memory the caller already
 * owns (e.g. the std::vector a factory keeps its batch inputs in). */
abc_status abc_host_alloc(abc_ctx *ctx, size_t bytes, void **out);
void abc_host_free(void *p);
abc_status abc_host_register(abc_ctx *ctx, void *p, size_t bytes);
void abc_host_unregister(void *p);
/* Device-side fault: a key-switch grid waits (bounded, about a second) for rows of the same grid; if such a wait ever
 * gives up (GPU preempted or time-sliced for that long) the results are invalid.  Every synchronising call then fails,
 * and so does every later call that computes on, decrypts or exports ciphertexts, until abc_clear_fault acknowledges it
 * (ciphertexts produced in between must be recomputed). */
int abc_faulted(const abc_ctx *ctx);
abc_status abc_clear_fault(abc_ctx *ctx);

/* parameter queries (SealCiphertextFactory::getCiphertextSlotSize, include/ast_opt/runtime/SealCiphertextFactory.h:90) */
uint32_t abc_poly_degree(const abc_ctx *ctx);
uint32_t abc_n_primes(const abc_ctx *ctx);        /* k */
uint32_t abc_n_limbs(const abc_ctx *ctx);         /* L = k-1 */
uint32_t abc_batch(const abc_ctx *ctx);
uint64_t abc_plain_modulus(const abc_ctx *ctx);
abc_status abc_get_primes(const abc_ctx *ctx, uint64_t *out_k);
/* auxiliary BEHZ bases chosen by SEAL's rule: out = [m_sk, gamma, B_0..B_{nB-1}]; returns nB+2 via *count */
abc_status abc_get_aux_primes(const abc_ctx *ctx, uint64_t *out, uint32_t *count);

/* --- keys: replaces KeyGenerator usage at src/runtime/SealCiphertextFactory.cpp:89-93
 * (secret key, public key, relinearisation key, the default Galois key set). Runs on the device. */
abc_status abc_keygen(abc_ctx *ctx);
/* same, with an explicit list of Galois elements instead of the default set (KeyGenerator::create_galois_keys(elts)):
 * at N = 65536 with 31 primes one key-switching key is 975 MiB, so only the rotations a program needs are generated */
abc_status abc_keygen_select(abc_ctx *ctx, const uint32_t *galois_elts, size_t n);
/* GaloisTool::get_elt_from_step: 3^step mod 2N (step < 0: N/2 - |step|), 2N-1 for step 0; 0 if |step| >= N/2 */
uint32_t abc_galois_elt_from_step(const abc_ctx *ctx, int step);
enum { ABC_KEY_SECRET = 0, ABC_KEY_PUBLIC = 1, ABC_KEY_RELIN = 2, ABC_KEY_GALOIS = 3 };
/* words: secret k*N, public 2*k*N, relin/galois L*2*k*N (layout [J][component][limb][N], NTT form) */
size_t abc_key_words(const abc_ctx *ctx, int kind);
abc_status abc_key_export(abc_ctx *ctx, int kind, uint32_t galois_elt, uint64_t *host, size_t words);
abc_status abc_key_import(abc_ctx *ctx, int kind, uint32_t galois_elt, const uint64_t *host, size_t words);
int abc_has_galois_key(const abc_ctx *ctx, uint32_t galois_elt);

/* --- ciphertext handles (SealCiphertext ctor/copy/clone: src/runtime/SealCiphertext.cpp:10-34,71-78) */
abc_status abc_ct_alloc(abc_ctx *ctx, abc_ct **out);            /* uninitialised size-2 ciphertexts */
void abc_ct_free(abc_ct *ct);                                   /* stream-ordered; safe right after enqueue */
abc_status abc_ct_clone(abc_ctx *ctx, const abc_ct *src, abc_ct **out);   /* O(1): clones share the device buffer, copy-on-write */
int abc_ct_shared(const abc_ct *ct);                            /* handles sharing this ciphertext's buffer (>= 1) */
int abc_ct_deferred(const abc_ct *ct);                          /* 1 while the handle is a deferred rotation (see abc_rotate_rows) */
size_t abc_ct_words(const abc_ctx *ctx);                        /* batch*2*L*N */
abc_status abc_ct_export(abc_ctx *ctx, const abc_ct *ct, uint64_t *host, size_t words);
abc_status abc_ct_import(abc_ctx *ctx, abc_ct *ct, const uint64_t *host, size_t words);

/* --- encode+encrypt / decrypt+decode
 * replaces createCiphertext (src/runtime/SealCiphertextFactory.cpp:9-24: expandVector pad-with-last :102-115,
 * BatchEncoder::encode :127-132, Encryptor::encrypt :12) and decryptCiphertext (:146-152).
 * slots: n values per instance (1 <= n <= N); instance b reads slots[b*n .. b*n+n) unless `broadcast`,
 * in which case all instances read slots[0..n).  Values are padded to N with the last one. */
abc_status abc_encode_encrypt(abc_ctx *ctx, const int64_t *slots, size_t n, int broadcast, abc_ct **out);
abc_status abc_decrypt_decode(abc_ctx *ctx, const abc_ct *ct, int64_t *out_slots /* batch*N */);
/* The same, enqueued: out_slots (pinned host memory if the copy is to be asynchronous) is valid once abc_decrypt_wait or
 * abc_sync has returned.  The result leaves on a separate D2H stream from one of two device buffers, so the copy of one
 * decryptCiphertext overlaps the kernels enqueued after it (a third decryption in flight waits, on the device, for the
 * first one's copy; the host buffers are the caller's to keep apart). */
abc_status abc_decrypt_decode_async(abc_ctx *ctx, const abc_ct *ct, int64_t *out_slots /* batch*N */);
abc_status abc_decrypt_wait(abc_ctx *ctx);
/* SealCiphertext::noiseBits = Decryptor::invariant_noise_budget (src/runtime/SealCiphertext.cpp:80-83): remaining
 * noise budget in bits, one value per instance of the batch; 0 means decryption is no longer guaranteed. */
abc_status abc_noise_budget(abc_ctx *ctx, const abc_ct *ct, int32_t *out_bits_per_instance);
/* Ciphertext::is_transparent per instance (1 = every word of c1 is zero).  SEAL's Evaluator throws std::logic_error
 * ("result ciphertext is transparent") on such results (SEAL_THROW_ON_TRANSPARENT_CIPHERTEXT, on by default; reached by
 * `x --- x` since src/runtime/RuntimeVisitor.cpp:436 clones every variable read).  Synchronises the stream; on a
 * limb-sharded context the verdict covers the limbs this rank owns. */
abc_status abc_is_transparent(abc_ctx *ctx, const abc_ct *ct, int32_t *out_flags_per_instance);
/* Encryption randomness: instance b of the next encryption draws (u, e0, e1) from stream salt + nonce*batch + b; the
 * nonce counts encryptions, the salt is drawn from the OS generator when the context is created (two contexts that share
 * a key seed, e.g. one factory per GPU, never repeat randomness).  This call sets the counter AND zeroes the salt: the
 * encryptions that follow are reproducible — for parity tests against the oracle only. */
abc_status abc_set_encrypt_nonce(abc_ctx *ctx, uint64_t nonce);
/* The sampler's 256-bit key, supplied by the caller (32 bytes of its own entropy) instead of the OS-drawn one or the
 * expansion of a 64-bit `seed`: what gives every GPU of a job the same keys at full key strength.  Call before
 * abc_keygen; it also keys the encryptions that follow.  Keys, u and the error polynomials are the ChaCha20 key stream
 * under this key (one stream per key / encryption / component), mapped to SEAL's distributions (uniform mod q_i with
 * rejection, ternary, centred binomial 21 - 21). */
abc_status abc_set_rng_key(abc_ctx *ctx, const uint8_t *key32);

/* --- ciphertext-ciphertext ops.  dst may alias a (the *Inplace variants of the reference).
 * add/sub: Evaluator::add/sub          (src/runtime/SealCiphertext.cpp:90-100,113-119)
 * negate:  Evaluator::negate            (:157,193)
 * mul_relin: multiply + relinearize_inplace (:102-107,121-124)
 * rotate_rows: Evaluator::rotate_rows   (:52-61); |steps| >= N/2 is an error; NAF path when no direct key */
abc_status abc_add(abc_ctx *ctx, abc_ct *dst, const abc_ct *a, const abc_ct *b);
abc_status abc_sub(abc_ctx *ctx, abc_ct *dst, const abc_ct *a, const abc_ct *b);
abc_status abc_negate(abc_ctx *ctx, abc_ct *dst, const abc_ct *a);
abc_status abc_mul_relin(abc_ctx *ctx, abc_ct *dst, const abc_ct *a, const abc_ct *b);
/* rotate_rows may DEFER its last key switch: the handle then stands for "Galois step e of buffer X" until something
 * needs the data.  abc_add with a deferred operand runs that key switch with the other operand accumulated in its
 * ModDown; every other consumer (and abc_sync does not count) materialises it first.  Results are bit-identical to
 * the eager order.  ABC_EAGER_ROTATE=1 in the environment disables deferral; limb-sharded contexts never defer. */
abc_status abc_rotate_rows(abc_ctx *ctx, abc_ct *dst, const abc_ct *a, int steps);
/* dst = add(rotate_rows(a, steps), addend) in one pass: the addend is accumulated in the ModDown of the last key
 * switch (bit-identical to the two calls).  This is what `acc = acc +++ r` with `r = rotate(x, k)` costs when the
 * wrappers defer the rotation (RuntimeVisitor.cpp:69,157).  dst may alias a and/or addend. */
abc_status abc_rotate_rows_add(abc_ctx *ctx, abc_ct *dst, const abc_ct *a, int steps, const abc_ct *addend);

/* --- ciphertext-plaintext ops (src/runtime/SealCiphertext.cpp:130-202). slots/n/broadcast as above.
 * The reference's all-(-1) negate fast path for multiplyPlain lives in the C++ wrapper. */
abc_status abc_add_plain(abc_ctx *ctx, abc_ct *dst, const abc_ct *a, const int64_t *slots, size_t n, int broadcast);
abc_status abc_sub_plain(abc_ctx *ctx, abc_ct *dst, const abc_ct *a, const int64_t *slots, size_t n, int broadcast);
abc_status abc_mul_plain(abc_ctx *ctx, abc_ct *dst, const abc_ct *a, const int64_t *slots, size_t n, int broadcast);

/* --- device-resident plaintext operands (so a benchmark loop need not re-upload slots) */
typedef struct abc_pt abc_pt;
abc_status abc_pt_encode(abc_ctx *ctx, const int64_t *slots, size_t n, int broadcast, abc_pt **out);
void abc_pt_free(abc_pt *pt);
abc_status abc_add_plain_pt(abc_ctx *ctx, abc_ct *dst, const abc_ct *a, const abc_pt *pt);
abc_status abc_sub_plain_pt(abc_ctx *ctx, abc_ct *dst, const abc_ct *a, const abc_pt *pt);
abc_status abc_mul_plain_pt(abc_ctx *ctx, abc_ct *dst, const abc_ct *a, const abc_pt *pt);
abc_status abc_encrypt_pt(abc_ctx *ctx, const abc_pt *pt, abc_ct **out);

/* --- kernel-level probes used by the parity tests and the profiler (one limb-row = N words)
 * mod_index: 0..k-1 key-level primes, k..k+nbsk-1 the Bsk primes (B..., m_sk), k+nbsk the plain modulus t */
abc_status abc_probe_ntt(abc_ctx *ctx, int inverse, uint32_t mod_index, uint64_t *host_rows, size_t n_rows);
/* BEHZ multiply without relinearisation: out3 gets batch*3*L*N words */
abc_status abc_probe_multiply(abc_ctx *ctx, const abc_ct *a, const abc_ct *b, uint64_t *host_out3, size_t words);
/* device-time of `iters` launches of the limb NTT kernel over n_rows resident rows (microbenchmark) */
abc_status abc_bench_ntt(abc_ctx *ctx, int inverse, uint32_t mod_index, size_t n_rows, int iters, float *ms);

/* --- limb sharding across the GPUs of one box (BASELINE.json configs[4]; one process + one context per GPU).
 * After abc_comm_init every ciphertext is LIMB-SHARDED: rank r owns data limbs [lo_r, hi_r) (contiguous, sizes differ
 * by at most one) of both polynomials of every instance; the other limbs of a handle's buffer are not maintained.
 * add / sub / negate / plain ops touch owned limbs only (no communication).  rotate_rows all-gathers c1 (NCCL over
 * NVLink) in front of the key-switch ModUp and computes only the output moduli it owns, plus the special prime,
 * which every rank computes so that ModDown needs no second collective.  mul_relin all-gathers its operands, computes
 * the BEHZ product on every rank and shards only the relinearisation.  decrypt_decode all-gathers first.  All ranks
 * must issue the same ops in the same order with the same seed (same keys).  Results are bit-identical to world = 1. */
#define ABC_COMM_ID_BYTES 128
abc_status abc_comm_unique_id(abc_ctx *ctx, uint8_t *out128);   /* ncclGetUniqueId; call on one rank, share the bytes */
abc_status abc_comm_init(abc_ctx *ctx, int rank, int world, const uint8_t *id128);   /* 1 <= world <= L */
/* payload this rank has received through all-gathers so far, and the NCCL collectives it issued for them */
abc_status abc_comm_stats(const abc_ctx *ctx, uint64_t *bytes_received, uint64_t *nccl_calls);
int abc_comm_rank(const abc_ctx *ctx);
int abc_comm_world(const abc_ctx *ctx);
abc_status abc_owned_limbs(const abc_ctx *ctx, uint32_t *lo, uint32_t *hi);
abc_status abc_ct_allgather(abc_ctx *ctx, abc_ct *ct);          /* make every limb of ct valid on every rank */

/* --- timing on the context's stream (CUDA events; torch.cuda.Event cannot see this stream) */
abc_status abc_timer_start(abc_ctx *ctx);
abc_status abc_timer_stop(abc_ctx *ctx, float *ms);            /* synchronises */
/* writes `bytes` of scratch to evict L2 between timed iterations */
abc_status abc_flush_l2(abc_ctx *ctx, size_t bytes);
/* kernels launched by this context so far */
uint64_t abc_launch_count(const abc_ctx *ctx);
/* key switches run so far (relinearisations and Galois steps): what deferred rotations, rotate-and-add fusion and the
 * rotation-prefix cache save shows here */
uint64_t abc_key_switch_count(const abc_ctx *ctx);
/* per-kernel-family device time: enable, then read {name, launches, total ms} rows as a JSON string */
abc_status abc_profile_enable(abc_ctx *ctx, int on);
const char *abc_profile_json(abc_ctx *ctx);

/* integer-pipe issue-rate microbenchmark (IMAD / IADD3-class ops per second, whole GPU) */
abc_status abc_measure_int_peak(abc_ctx *ctx, double *imad_per_s, double *iadd_per_s);
/* register-resident 64-bit NTT butterflies per second, whole GPU (no memory traffic): the ceiling NTT kernels are
 * quoted against.  arith_class: 0 Shoup, 1 FP64-assisted, 2 FP64-assisted without range guards (see csrc/ntt.cuh) */
abc_status abc_measure_butterfly_peak(abc_ctx *ctx, int arith_class, double *butterflies_per_s);
/* arithmetic class the context uses for its key-level primes */
int abc_ntt_arith_class(const abc_ctx *ctx);

/* --- Microsoft SEAL 3.6 binary streams (seal::Serialization::Save/Load; SEAL is the un-vendored dependency behind
 * src/runtime/SealCiphertext*.cpp, pinned at 3.6.5 by Docker/Dockerfile:9).  This is how "the same SEAL-serialised
 * input ciphertexts and keys" reach the device: seal::Ciphertext (one per batch instance), SecretKey, PublicKey,
 * RelinKeys, GaloisKeys, EncryptionParameters.  compr_mode none / zlib / zstd on load and save (zstd through
 * libzstd.so.1 when it can be dlopen'ed).  Loads validate magic, version, parms_id (BLAKE2b of the parameters), sizes,
 * NTT-form flag and coefficient ranges, like seal::is_valid_for.  Seeded (symmetric) ciphertexts are rejected.
 * save: *len receives the stream size; with buf == NULL or cap too small the call fails and only reports the size. */
enum { ABC_SEAL_COMPR_NONE = 0, ABC_SEAL_COMPR_ZLIB = 1, ABC_SEAL_COMPR_ZSTD = 2 };
abc_status abc_seal_parms_id(abc_ctx *ctx, int key_level, uint64_t out[4]);  /* EncryptionParameters::parms_id */
abc_status abc_seal_params_save(abc_ctx *ctx, int compr, uint8_t *buf, size_t cap, size_t *len);
/* EncryptionParameters stream -> abc_params (poly_degree, primes, plain_modulus; device/batch/seed untouched) */
abc_status abc_seal_params_parse(const uint8_t *bytes, size_t len, abc_params *out, uint64_t *primes_out, size_t primes_cap);
abc_status abc_seal_ct_save(abc_ctx *ctx, const abc_ct *ct, uint32_t instance, int compr, uint8_t *buf, size_t cap, size_t *len);
abc_status abc_seal_ct_load(abc_ctx *ctx, abc_ct *ct, uint32_t instance, const uint8_t *bytes, size_t len);
abc_status abc_seal_key_save(abc_ctx *ctx, int kind, int compr, uint8_t *buf, size_t cap, size_t *len);
abc_status abc_seal_key_load(abc_ctx *ctx, int kind, const uint8_t *bytes, size_t len);  /* GALOIS: every key in the stream */
/* raw coefficients of ONE instance, [2][L][N] (what a seal::Ciphertext's DynArray holds) */
abc_status abc_ct_export_instance(abc_ctx *ctx, const abc_ct *ct, uint32_t instance, uint64_t *host, size_t words);
abc_status abc_ct_import_instance(abc_ctx *ctx, abc_ct *ct, uint32_t instance, const uint64_t *host, size_t words);
/* Galois elements with a key on the device: *n receives the count; out may be NULL to query it */
abc_status abc_galois_elts(const abc_ctx *ctx, uint32_t *out, size_t cap, size_t *n);

#ifdef __cplusplus
}
#endif
#endif
