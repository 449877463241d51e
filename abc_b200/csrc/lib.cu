// lib.cu — context, handles and op orchestration behind the C ABI of include/abc_b200.h.
// Host code only sequences kernels on the context's stream; all arithmetic on ciphertexts, keys and
// plaintexts happens in the sm_100a kernels of kernels.cuh.  There is no CPU fallback.
#include "../../include/abc_b200.h"

#include <dlfcn.h>
// NCCL is loaded with dlopen on first use (limb sharding only), so the library builds and runs without NCCL installed: the
// few types of its C API this file needs are declared here (values as in nccl.h 2.x: ncclSuccess = 0, ncclUint64 = 5,
// ncclUniqueId = 128 opaque bytes)
typedef struct ncclComm *ncclComm_t;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclUint64 = 5 } ncclDataType_t;
typedef struct { char internal[128]; } ncclUniqueId;
#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges are no-ops unless a profiler is attached
#include <sys/random.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "hostmath.hpp"
#include "kernels.cuh"
#include "ksfused.cuh"
#include "kschain.cuh"
#include "ksred.cuh"
#include "ks14.cuh"
#include "behz_f64.cuh"

namespace {
thread_local std::string g_create_error;

struct ProfRec { const char *name; cudaEvent_t a, b; };

enum { DOM_SK = 1, DOM_PK = 2, DOM_KSK = 3, DOM_ENC = 4 };
}  // namespace

struct CtBuf;
struct abc_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  // host slots travel on their own stream into one of two staging buffers, so the copy of the next operand overlaps the
  // kernels of the previous one (createCiphertext(x); createCiphertext(y): y's H2D runs under x's encryption)
  cudaStream_t copy_stream = nullptr;
  long long *h2d_buf[2] = {nullptr, nullptr};
  size_t h2d_words[2] = {0, 0};
  cudaEvent_t h2d_copied[2] = {nullptr, nullptr}, h2d_consumed[2] = {nullptr, nullptr};
  unsigned h2d_next = 0;
  // decrypted slots leave on their own stream from one of two device buffers: the D2H of one decryptCiphertext runs under
  // the kernels of whatever is enqueued next (abc_decrypt_decode_async + abc_decrypt_wait)
  cudaStream_t d2h_stream = nullptr;
  long long *d2h_buf[2] = {nullptr, nullptr};
  size_t d2h_words[2] = {0, 0};
  cudaEvent_t d2h_ready[2] = {nullptr, nullptr}, d2h_done[2] = {nullptr, nullptr};
  unsigned d2h_next = 0;
  int N = 0, logN = 0, k = 0, L = 0, nB = 0, nbsk = 0, B = 1, W = 0;
  u64 t = 0, seed = 0, enc_nonce = 0, gamma = 0, msk = 0;
  // encryption randomness: stream id = enc_salt + nonce * batch + instance.  The salt is drawn from the OS per context,
  // so two contexts that share a key seed (one factory per GPU) never reuse (u, e0, e1); abc_set_encrypt_nonce makes the
  // stream reproducible (salt 0) for parity tests.
  u64 enc_salt = 0;
  // the sampler's 256-bit ChaCha20 key (modarith.cuh): 32 bytes from the OS generator, or expanded from an explicit seed
  // (tests; the same keys on every GPU of a job), or set by abc_set_rng_key
  RngKey rng_key{};
  std::vector<u64> primes, bsk;
  DevConst hC;
  DevConst *dC = nullptr;
  ModInfo *d_mods = nullptr;
  std::vector<ModInfo> hmods;
  int prefetch_ahead = 0;   // CTAs resident at once (2 per SM): a CTA L2-prefetches the row of the CTA that replaces it
  int ar_q = 0, ar_t = 0, force_ar = -1;  // NTT arithmetic class of the key-level primes / of t (ntt.cuh)
  int ks_skew = 8;                               // ABC_KS_SKEW: special-prime rows run this many instances ahead
  bool no_square = false;                        // ABC_NO_SQUARE: multiply(x, x) takes the general path
  bool lazy_rotate = true;                       // ABC_EAGER_ROTATE: rotate_rows runs its last key switch immediately
  // Intermediate results of multi-step (NAF) rotations, by (source buffer, Galois element): rotate(x, 63) and
  // rotate(x, -65) both start with the step -1 of x, and rotate(x, -1) itself is that buffer (a 3x3 stencil needs 8 key
  // switches instead of 12).  An entry holds a reference on both buffers, so neither can be written in place or
  // reused while it is cached (copy-on-write); entries whose source nobody else holds are dropped.  ABC_ROT_CACHE=n
  // entries (default 4, 0 = off).
  struct RotCacheEntry { CtBuf *src; u32 elt; CtBuf *res; uint64_t stamp; };
  std::vector<RotCacheEntry> rot_cache;
  uint64_t rot_stamp = 0, rot_cache_hits = 0, key_switches = 0;
  int rot_cache_max = 4;
  bool ks_unmerged = false, ks_unfused = false;  // ABC_KS_UNMERGED / ABC_KS_UNFUSED: A/B switches for the key-switch tail
  bool ks_no_discard = false;                    // ABC_KS_NO_DISCARD: leave the consumed ModUp rows to L2's write-back
  bool ks_no_image = false;                      // ABC_KS_NO_IMAGE: ModUp rows stored element by element instead of as bulk-copied images
  int ks_one_launch = -1;                        // ABC_KS_ONE_LAUNCH=0/1: force the single-launch key switch off / on (-1: by size)
  int ks1_threads = 1024;                        // ABC_KS1_THREADS: CTA size of the single-launch key switch at N = 8192
  int ks1_skew = 40;                             // ABC_KS1_SKEW: single-launch key switch, special units run this far ahead
  int idx_t = 0;
  std::vector<void *> owned;  // device allocations freed at destroy
  // released ciphertext buffers, ready for reuse on c->stream (salloc / sfree); at most blk_cache_limit bytes are parked
  std::vector<std::pair<size_t, u64 *>> blk_free;
  std::map<void *, size_t> blk_live;
  size_t blk_free_words = 0, blk_cache_limit = (size_t)32 << 30;
  u32 *d_index_map = nullptr;
  int *rm_ct = nullptr;     // [3L]  w % L
  int *rm_key = nullptr;    // [2k]  w % k
  // BEHZ product on the FP64 pipe over a sub-2^45 auxiliary base (behz_f64.cuh); exact-double class contexts, L <= 8
  bool behz_f64 = false;                       // ABC_BEHZ_F64=0 keeps SEAL's 61-bit base and the integer kernels
  bool behz_fused = true;                      // ABC_BEHZ_FUSED=0: separate tensor launch, canonical rows between the launches
  int nbsk2 = 0, W2 = 0, idx_b2 = 0;           // |Bsk'|, rows per polynomial, index of its first modulus in d_mods
  std::vector<u64> bsk2;                       // B' primes, then m_sk'
  BehzF64 *dF = nullptr;
  int *rm_behz2 = nullptr;                     // [4 W2] modulus of X row
  int *rm_behz = nullptr;   // [4W]  q / Bsk modulus of X row
  int *rm_behz_q = nullptr, *rd_behz_q = nullptr;  // [4L]    q rows of X: modulus, row
  int *rm_behz_b = nullptr, *rd_behz_b = nullptr;  // [4nbsk] Bsk rows of X: modulus, row
  int *rm_modup = nullptr;  // [kL]  I
  int *rs_modup = nullptr;  // [kL]  J
  int *rm_t = nullptr;      // [1]   idx_t
  int *rs_c1 = nullptr;     // [L]   L + w
  int *rm_special = nullptr, *rd_special = nullptr;  // [2] special-prime rows of an accumulator block
  int *rs_accq = nullptr;   // [2L]  comp*k + i: data rows of an accumulator block
  // limb sharding (multi-GPU, one process per GPU): this rank owns data limbs [own_lo, own_hi) of every ciphertext and
  // computes the key-switch output moduli own U {special}; world == 1 owns everything.  Maps below follow the own set.
  int rank = 0, world = 1, own_lo = 0, own_hi = 0;
  ncclComm_t comm = nullptr;
  cudaStream_t comm_stream = nullptr;               // all-gathers that overlap compute (ModUp of locally owned source limbs)
  cudaEvent_t comm_ready = nullptr, comm_done = nullptr;
  int *rm_up_own = nullptr, *rd_up_own = nullptr, *rs_up_own = nullptr, n_up_own = 0;   // ModUp rows whose source limb this rank owns
  int *rm_up_oth = nullptr, *rd_up_oth = nullptr, *rs_up_oth = nullptr, n_up_oth = 0;   // ... and the rows that wait for the all-gather
  bool shard_overlap = true, shard_cols = true;     // ABC_SHARD_OVERLAP=0 / ABC_SHARD_COLS=0: A/B switches
  uint64_t gathered_bytes = 0, gather_calls = 0;    // payload this rank received through all-gathers / NCCL calls issued
  int ks_nI = 0;                                    // |own| + 1
  int *ks_I = nullptr;                              // [ks_nI] output moduli of the key switch (own..., special)
  int *rm_modup_s = nullptr, *rd_modup_s = nullptr, *rs_modup_s = nullptr;   // [ks_nI * L]
  int *rm_md = nullptr, *rd_md = nullptr, *rs_md = nullptr;                  // [2 * nown] ModDown rows
  int *rm_own = nullptr, *rd_own = nullptr;                                  // [2 * nown] own rows of a ciphertext
  int *rm_mdm = nullptr, *rd_mdm = nullptr, *rs_mdm = nullptr;               // [2 + 2 * nown] merged special + ModDown rows
  uint2 *ks_sched = nullptr; int ks_sched_n = 0;                                // chained key switch: block schedule (kschain.cu)
  u32 *ks_done = nullptr; u32 ks_chain_serial = 0;                            // ... [B][k] ModUp rows stored so far (L per launch)
  int ks_chain = 1, ks_chain_skew = 16;                                       // ABC_KS_CHAIN=0/1, ABC_KS_CHAIN_SKEW
  u32 *ks_fault_h = nullptr, *ks_fault_d = nullptr;                           // host-mapped: a dependency wait of a key-switch grid gave up
  u32 *ks_flags = nullptr; u32 ks_serial = 0;                                 // [B][2] ready flags of the merged launch
  // accumulating key switch (ksred.cu): accumulator ring [ring][k][2][N] doubles, [B][k] done / freed counters
  // ABC_KS_RED=1 selects it; measured 9 % slower than the chained grid at N = 8192 (the bulk reductions wait on the L2 atomic units)
  int ks_red = 0, ksr_ring = 0; double *ksr_acc = nullptr; u32 *ksr_done = nullptr, *ksr_freed = nullptr; u32 ksr_serial = 0;
  // split key switch at N = 16384 (ks14.cu): schedule of half-rows, [B][k][2] done + [B][2k][2] exchange + [B][2][2] special flags
  int ks_persist = 0, n_sms = 148;   // ABC_KS_PERSIST=1: the chained key switch on a persistent grid (2 CTAs per SM take tickets in a loop)
  int ks_split_maxb = 8;   // ABC_KS_SPLIT_MAXB: N = 8192 contexts of at most this many instances use the split rows too (latency)
  int ks14 = 1; uint2 *ks14_sched = nullptr; int ks14_sched_n = 0; u32 *ks14_done = nullptr, *ks14_xflag = nullptr, *ks14_flags = nullptr;
  u32 ks14_serial = 0;
  u32 *ks14_xflag_behz = nullptr; u32 ks14_behz_serial = 0;   // [B][3 W2][2] exchange flags of the BEHZ inverse half-rows
  u32 *ks_ticket = nullptr; u32 ks_ticket_total = 0;                          // start-order tickets of the dependency-ordered grids (limb.cuh grid_ticket)
  // handles (abc_ct / abc_pt) keep their context alive: abc_ctx_destroy with live handles only marks it, the last
  // abc_ct_free / abc_pt_free completes the destruction (a ciphertext may outlive its factory, as with SealCiphertext)
  long live_handles = 0; bool zombie = false;
  bool sync_each = getenv("ABC_SYNC_EACH") != nullptr;
  bool faulted = false;                                                       // sticky: a dependency wait timed out; every later call fails until abc_clear_fault
  int *rs_zero = nullptr;   // [2k]  0
  u64 *d_sk = nullptr, *d_pk = nullptr, *d_relin = nullptr;
  u64 *d_noise_tab = nullptr; int q_bits = 0;   // abc_noise_budget: (Q/q_i), Q, (Q+1)/2 as L-word integers; bit_count(Q)
  std::map<u32, u64 *> galois;
  std::map<const u64 *, double *> key_f64;  // exact-double copies of key-switch keys (single-launch key switch), made on first use
  bool have_keys = false;
  std::string err;
  uint64_t launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool prof = false;
  std::vector<ProfRec> prof_recs;
  std::string prof_json;
  size_t flush_bytes = 0;
  void *flush_buf = nullptr;
  // grow-only scratch slots reused by every op (one stream: an op's scratch is dead before the next op starts).
  // Allocating these per op from the stream-ordered pool fragmented it (135 MB / 335 MB / 170 MB blocks) and cost ms.
  u64 *sc_ptr[32] = {nullptr};
  size_t sc_words[32] = {0};
};
// A ciphertext handle.  The device buffer is shared between clones (copy-on-write): RuntimeVisitor clones on every
// variable read (RuntimeVisitor.cpp:436), so clone is O(1) and an op gives its destination a private buffer first.
// A buffer may also be a DEFERRED rotation (d == nullptr): its value is the last Galois step of a rotate_rows applied
// to `src`; an add consumes it fused (the addend is accumulated in that key switch's ModDown), anything else resolves it.
// After a fused add the rotation itself was never stored; a handle that still stands for it becomes the DIFFERENCE
// sum - other of two live buffers (exact: canonical modular subtraction undoes the addition), one HBM-bound kernel if
// anybody ever asks, nothing if the handle is simply dropped (the usual fate of `r` in `acc = acc +++ r`).
struct CtBuf {
  u64 *d = nullptr; int refs = 1;
  CtBuf *src = nullptr; u32 elt = 0;       // deferred rotation: value = Galois step `elt` of src
  CtBuf *sum = nullptr, *other = nullptr;  // deferred difference: value = sum - other
};
struct abc_ct { abc_ctx *ctx; CtBuf *b; };
extern "C" { static void rot_cache_clear(abc_ctx *c); }   // defined with the handle layer below
struct abc_pt { abc_ctx *ctx; u64 *d; int broadcast; };

namespace {

#define CK(call)                                                                                      \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess) {                                                                          \
      c->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                    \
      return ABC_ERR_CUDA;                                                                            \
    }                                                                                                 \
  } while (0)
#define TRY(call)                                                                                     \
  do {                                                                                                \
    abc_status s_ = (call);                                                                           \
    if (s_ != ABC_OK) return s_;                                                                      \
  } while (0)

abc_status fail(abc_ctx *c, abc_status s, const std::string &msg) { c->err = msg; return s; }
// after a stream synchronisation: did a key-switch grid give up waiting for one of its own producers?
const char *kFaultMsg = "key switch: a dependency wait inside the grid timed out; ciphertexts computed since the last "
                        "successful synchronisation are invalid (context poisoned until abc_clear_fault)";
// after a stream synchronisation: did a key-switch grid give up waiting for one of its own producers?  Sticky: once seen,
// every call that computes on or exports ciphertexts fails until the caller acknowledges it with abc_clear_fault.
abc_status check_ks_fault(abc_ctx *c) {
  if (c->ks_fault_h && *reinterpret_cast<volatile u32 *>(c->ks_fault_h)) c->faulted = true;
  return c->faulted ? fail(c, ABC_ERR_CUDA, kFaultMsg) : ABC_OK;
}
#define CHECK_POISON(c) do { if ((c)->faulted) return fail((c), ABC_ERR_CUDA, kFaultMsg); } while (0)

// NVTX range around one ABI call ("abc_mul_relin", "abc_rotate_rows", ...): what nsys / ncu --nvtx group kernels by
// ABC_HOST_TRACE=<ms>: every ABI call, launch and stream-ordered allocation whose HOST side took longer than that is
// reported on stderr (the device-side work is asynchronous: a slow one is a blocking driver call or a descheduled thread)
static double host_trace_ms() {
  static const double thr = [] { const char *e = getenv("ABC_HOST_TRACE"); return e ? atof(e) : -1.0; }();
  return thr;
}
struct NvtxOp {
  const char *name; std::chrono::steady_clock::time_point t0;
  explicit NvtxOp(const char *n) : name(n) { nvtxRangePushA(n); if (host_trace_ms() >= 0) t0 = std::chrono::steady_clock::now(); }
  ~NvtxOp() {
    nvtxRangePop();
    if (host_trace_ms() >= 0) {
      const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      if (ms > host_trace_ms()) fprintf(stderr, "[abc_b200 host] %-28s %9.3f ms\n", name, ms);
    }
  }
};
struct Launch {
  abc_ctx *c; const char *name; cudaEvent_t a = nullptr, b = nullptr;
  NvtxOp range;
  Launch(abc_ctx *c_, const char *n) : c(c_), name(n), range(n) {
    c->launches++;
    if (c->prof) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, c->stream); }
  }
  ~Launch() {
    if (c->prof) { cudaEventRecord(b, c->stream); c->prof_recs.push_back({name, a, b}); }
    if (c->sync_each) {   // ABC_SYNC_EACH=1 (debugging): name the launch a device-side error belongs to
      const cudaError_t e = cudaStreamSynchronize(c->stream);
      if (e != cudaSuccess) fprintf(stderr, "[abc_b200] %s: %s\n", name, cudaGetErrorString(e));
    }
  }
};

// Ciphertext buffers come in two or three sizes per context and are allocated and dropped by every op, so released
// blocks are kept on a per-context free list and handed out again without a driver call.  cudaMallocAsync, although
// stream-ordered, was measured to BLOCK the host for 1.5 ... 216 ms now and then (first allocations after a
// synchronisation, `ABC_HOST_TRACE`), which is what made end-to-end steps stall on a device that was waiting for work.
// Reuse is safe under the rule the stream-ordered free already relied on: a buffer is released only after everything
// that touches it has been enqueued on (or joined into) c->stream, and the next user is enqueued on c->stream.
static bool release_parked(abc_ctx *c) {
  rot_cache_clear(c);   // cached rotation results are memory we can do without
  if (c->blk_free.empty()) return false;
  for (auto &b : c->blk_free) cudaFreeAsync(b.second, c->stream);
  c->blk_free.clear(); c->blk_free_words = 0;
  return true;
}
abc_status salloc(abc_ctx *c, u64 **p, size_t words) {
  for (size_t i = c->blk_free.size(); i-- > 0;)
    if (c->blk_free[i].first == words) {
      *p = c->blk_free[i].second;
      c->blk_free.erase(c->blk_free.begin() + (long)i);
      c->blk_free_words -= words;
      c->blk_live[*p] = words;
      return ABC_OK;
    }
  NvtxOp nvtx_("cudaMallocAsync");
  cudaError_t e = cudaMallocAsync((void **)p, words * sizeof(u64), c->stream);
  if (e != cudaSuccess && release_parked(c)) {   // out of memory with blocks of other sizes parked: give them back, retry
    cudaGetLastError();
    e = cudaMallocAsync((void **)p, words * sizeof(u64), c->stream);
  }
  CK(e);
  c->blk_live[*p] = words;
  return ABC_OK;
}
void sfree(abc_ctx *c, void *p) {
  if (!p) return;
  auto it = c->blk_live.find(p);
  if (it == c->blk_live.end()) { cudaFreeAsync(p, c->stream); return; }   // not a salloc block (small per-call arrays)
  const size_t words = it->second;
  c->blk_live.erase(it);
  if (c->blk_free.size() < 24 && (c->blk_free_words + words) * sizeof(u64) <= c->blk_cache_limit) {
    c->blk_free.emplace_back(words, (u64 *)p);
    c->blk_free_words += words;
  } else {
    cudaFreeAsync(p, c->stream);
  }
}
enum { SC_T = 0, SC_ACC, SC_X, SC_OUT3, SC_U, SC_TMP, SC_DECX, SC_DECP, SC_P, SC_NK, SC_ENTT, SC_COMM_SEND, SC_COMM_ALL, SC_Y, SC_SX, SC_XCH, SC_XI, SC_RND, SC_NSLOTS };
static_assert(SC_NSLOTS <= 32, "abc_ctx::sc_ptr / sc_words hold 32 slots");
abc_status scratch(abc_ctx *c, int slot, u64 **p, size_t words) {
  if (c->sc_words[slot] < words) {
    if (c->sc_ptr[slot]) cudaFreeAsync(c->sc_ptr[slot], c->stream);
    c->sc_ptr[slot] = nullptr; c->sc_words[slot] = 0;
    cudaError_t e = cudaMallocAsync((void **)&c->sc_ptr[slot], words * sizeof(u64), c->stream);
    if (e != cudaSuccess && release_parked(c)) {
      cudaGetLastError();
      e = cudaMallocAsync((void **)&c->sc_ptr[slot], words * sizeof(u64), c->stream);
    }
    CK(e);
    c->sc_words[slot] = words;
  }
  *p = c->sc_ptr[slot];
  return ABC_OK;
}

template <typename T> abc_status upload(abc_ctx *c, T **dst, const std::vector<T> &v) {
  CK(cudaMalloc((void **)dst, v.size() * sizeof(T)));
  c->owned.push_back(*dst);
  CK(cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return ABC_OK;
}

// ---- limb-pipeline launcher (kernels live in limb_12/13/14.cu)
abc_status launch_limb(abc_ctx *c, int combo, int ar, const LimbJob &job, int W, int B, const char *name) {
  if (c->force_ar >= 0 && c->force_ar < ar) ar = c->force_ar;
  LimbJob jj = job;
  jj.prefetch_ahead = c->prefetch_ahead;
  Launch l(c, name);
  int e;
  switch (c->logN) {
    case 12: e = limb_dispatch<12>(combo, ar, jj, c->d_mods, W, B, c->stream); break;
    case 13: e = limb_dispatch<13>(combo, ar, jj, c->d_mods, W, B, c->stream); break;
    case 14: e = limb_dispatch<14>(combo, ar, jj, c->d_mods, W, B, c->stream); break;
    case 15: case 16: e = limb_dispatch_big(c->logN - 13, combo, job, c->d_mods, W, B, c->stream); break;
    default: return fail(c, ABC_ERR_UNSUPPORTED, "poly_degree not supported (4096 .. 65536)");
  }
  if (e != 0) { c->err = std::string(name) + ": " + cudaGetErrorString((cudaError_t)e); return ABC_ERR_CUDA; }
  return ABC_OK;
}
LimbJob blank_job() { LimbJob j; memset(&j, 0, sizeof j); return j; }
// explicit 64-bit seed -> sampler key (part of the sampler definition: DESIGN.md "Sampler")
RngKey rng_key_from_seed(u64 seed) {
  RngKey key{};
  key.k[0] = (u32)seed; key.k[1] = (u32)(seed >> 32);
  key.k[2] = 0x2d636261u; key.k[3] = 0x30303262u;   // "abc-b200"
  return key;
}

#define DISPATCH_L(c, EXPR)                                                        \
  switch ((c)->L) {                                                                \
    case 1: { constexpr int LL = 1; EXPR; } break;                                 \
    case 2: { constexpr int LL = 2; EXPR; } break;                                 \
    case 3: { constexpr int LL = 3; EXPR; } break;                                 \
    case 4: { constexpr int LL = 4; EXPR; } break;                                 \
    case 5: { constexpr int LL = 5; EXPR; } break;                                 \
    case 6: { constexpr int LL = 6; EXPR; } break;                                 \
    case 7: { constexpr int LL = 7; EXPR; } break;                                 \
    case 8: { constexpr int LL = 8; EXPR; } break;                                 \
    case 9: { constexpr int LL = 9; EXPR; } break;                                 \
    case 10: { constexpr int LL = 10; EXPR; } break;                               \
    case 11: { constexpr int LL = 11; EXPR; } break;                               \
    case 12: { constexpr int LL = 12; EXPR; } break;                               \
    case 13: { constexpr int LL = 13; EXPR; } break;                               \
    case 14: { constexpr int LL = 14; EXPR; } break;                               \
    case 15: { constexpr int LL = 15; EXPR; } break;                               \
    default: { constexpr int LL = 0; EXPR; } break;                                 \
  }

// ---- context construction --------------------------------------------------------------------
void fill_mod(ModInfo &m, u64 q, int N, int logN, std::vector<ulonglong2> &tw, std::vector<ulonglong2> &itw,
              std::vector<ulonglong2> &twf, std::vector<ulonglong2> &itwf, std::vector<ulonglong2> &twd,
              std::vector<ulonglong2> &itwd, std::vector<ulonglong2> &twp, std::vector<ulonglong2> &itwp) {
  using hm::mulmod; using hm::invmod; using hm::shoup; using hm::prod_mod; using hm::barrett_ratio; using hm::bit_reverse; using hm::minimal_2nth_root; using hm::get_primes; using hm::bits_of; using hm::prod_bits;
  m.q = q;
  barrett_ratio(q, m.mu_hi, m.mu_lo);
  m.ninv = invmod((u64)N % q, q);
  m.ninv_s = shoup(m.ninv, q);
  const u64 psi = minimal_2nth_root(q, (u64)N), ipsi = invmod(psi, q);
  tw.assign(N, make_ulonglong2(0, 0));
  itw.assign(N, make_ulonglong2(0, 0));
  u64 pw = 1, ipw = 1;
  for (int i = 0; i < N; ++i) {
    const uint32_t j = bit_reverse((uint32_t)i, logN);
    tw[j] = make_ulonglong2(pw, shoup(pw, q));
    itw[j] = make_ulonglong2(ipw, shoup(ipw, q));
    pw = mulmod(pw, psi, q);
    ipw = mulmod(ipw, ipsi, q);
  }
  m.wl_ninv = mulmod(itw[1].x, m.ninv, q);
  m.wl_ninv_s = shoup(m.wl_ninv, q);
  // FP64-assisted class: companion = double(w/q) (correctly rounded: both operands are exact doubles)
  auto dbits = [q](u64 w) { double d = (double)w / (double)q; u64 b; memcpy(&b, &d, 8); return b; };
  m.ar_class = AR_SHOUP;
  // exact-double class: q < 0.97 * 2^45 (the inverse transform's range plan keeps 7 stages between two reductions)
  // (range plans in ntt.cuh; AR_FP_LAZY: ABC_FORCE_AR=2).  Wide primes (up to 2^49: SEAL's N = 16384 defaults) run the
  // exact-double class with its extra reductions on the generic stage plan, i.e. at N = 16384; ABC_F64_WIDE=0 keeps them on AR_FP.
  const bool wide_ok = ABC_F64_FRND && !(getenv("ABC_F64_WIDE") && atoi(getenv("ABC_F64_WIDE")) == 0);
  if ((q >> 49) == 0) m.ar_class = (q < ABC_F64_NARROW_MAX || (wide_ok && logN == 14)) ? AR_F64 : AR_FP;
  twf.clear(); itwf.clear();
  m.ninv_f = m.wl_ninv_f = m.qinv_bits = 0;
  if (m.ar_class != AR_SHOUP) {
    twf.resize(N); itwf.resize(N);
    for (int j = 0; j < N; ++j) {
      twf[j] = make_ulonglong2(tw[j].x, dbits(tw[j].x));
      itwf[j] = make_ulonglong2(itw[j].x, dbits(itw[j].x));
    }
    m.ninv_f = dbits(m.ninv); m.wl_ninv_f = dbits(m.wl_ninv);
    double qi = 1.0 / (double)q; memcpy(&m.qinv_bits, &qi, 8);
  }
  twd.clear(); itwd.clear(); twp.clear(); itwp.clear();
  m.ninv_d = m.wl_ninv_d = 0;
  if (m.ar_class == AR_F64) {
    auto ibits = [](u64 w) { double d = (double)w; u64 b; memcpy(&b, &d, 8); return b; };  // exact: w < 2^45
    const int np = N < 1024 ? N : 1024;               // strided passes cover stages 0..8 at most
    twp.resize(np); itwp.resize(np);
    for (int j = 0; j < np; ++j) {
      twp[j] = make_ulonglong2(ibits(tw[j].x), dbits(tw[j].x));
      itwp[j] = make_ulonglong2(ibits(itw[j].x), dbits(itw[j].x));
    }
    // 8-byte entries (two per ulonglong2 slot), the last two stages lane-contiguous for the contiguous pass (ntt.cuh tw_get)
    twd.assign(N / 2, make_ulonglong2(0, 0)); itwd.assign(N / 2, make_ulonglong2(0, 0));
    u64 *fw = reinterpret_cast<u64 *>(twd.data()), *iw = reinterpret_cast<u64 *>(itwd.data());
    for (int j = 1; j < N; ++j) {
      int s = 0;
      while ((2 << s) <= j) ++s;                      // stage of index j: 2^s <= j < 2^(s+1)
      int dst = j;
      if (s >= logN - 2) {
        const int per = 1 << (s - (logN - 3)), i = j - (1 << s);   // 2 or 4 twiddles per thread of the contiguous pass
        dst = (1 << s) + (i % per) * (N / 8) + i / per;
      }
      fw[dst] = ibits(tw[j].x); iw[dst] = ibits(itw[j].x);
    }
    m.ninv_d = ibits(m.ninv); m.wl_ninv_d = ibits(m.wl_ninv);
  }
}

abc_status build_shard_maps(abc_ctx *c);

abc_status build_tables(abc_ctx *c) {
  using hm::mulmod; using hm::invmod; using hm::shoup; using hm::prod_mod; using hm::barrett_ratio; using hm::bit_reverse; using hm::minimal_2nth_root; using hm::get_primes; using hm::bits_of; using hm::prod_bits;
  const int N = c->N, logN = c->logN, k = c->k, L = c->L;
  const u64 t = c->t;
  std::vector<u64> Q(c->primes.begin(), c->primes.begin() + L);  // data-level base
  // auxiliary bases (RNSTool::initialize)
  int nB = L;
  if (32 + bits_of(t) + prod_bits(Q) >= 61 * L + 61) nB++;
  if (nB != L) return fail(c, ABC_ERR_UNSUPPORTED, "parameter set needs |B| = L+1 (not supported)");
  std::vector<u64> aux = get_primes((u64)N, 61, (size_t)nB + 2);
  c->msk = aux[0]; c->gamma = aux[1];
  std::vector<u64> Bv(aux.begin() + 2, aux.end());
  c->bsk = Bv; c->bsk.push_back(c->msk);
  c->nB = nB; c->nbsk = nB + 1; c->W = L + c->nbsk;
  c->idx_t = k + c->nbsk;

  // ---- second auxiliary base for the FP64 BEHZ path: 44-bit primes, product above 32 + bits(t) + bits(Q) + 8 bits
  c->behz_f64 = false; c->bsk2.clear();
  {
    bool all_small = L <= BF_MAXQ;
    {
      const bool wide_ok = ABC_F64_FRND && !(getenv("ABC_F64_WIDE") && atoi(getenv("ABC_F64_WIDE")) == 0);
      for (int i = 0; i < k; ++i)
        all_small = all_small && (c->primes[i] < ABC_F64_NARROW_MAX || (wide_ok && logN == 14 && (c->primes[i] >> 49) == 0));
    }
    const char *e = getenv("ABC_BEHZ_F64");
    if (all_small && !(e && atoi(e) == 0) && logN <= 14) {
      const int need_bits = 32 + bits_of(t) + prod_bits(Q) + 8;
      const int cnt = (need_bits + 42) / 43;                       // a 44-bit prime carries more than 43 bits
      if (cnt <= BF_MAXB) {
        std::vector<u64> cand = get_primes((u64)N, 44, (size_t)cnt + k + 3);
        for (u64 p44 : cand) {
          bool used = p44 == t || p44 == c->gamma;
          for (int i = 0; i < k; ++i) used = used || p44 == c->primes[i];
          if (!used && (int)c->bsk2.size() < cnt) c->bsk2.push_back(p44);
        }
        if ((int)c->bsk2.size() == cnt) {   // B' = bsk2[0 .. cnt-2], m_sk' = bsk2[cnt-1]
          c->behz_f64 = true; c->nbsk2 = cnt; c->W2 = L + cnt; c->idx_b2 = k + c->nbsk + 1;
        }
      }
    }
  }

  // ---- per-modulus tables
  const int nmods = k + c->nbsk + 1 + (c->behz_f64 ? c->nbsk2 : 0);
  std::vector<ModInfo> mods(nmods);
  std::vector<ulonglong2> tw, itw, twf, itwf, twd, itwd, twp, itwp;
  for (int i = 0; i < nmods; ++i) {
    const u64 q = i < k ? c->primes[i] : (i < k + c->nbsk ? c->bsk[i - k] : (i == k + c->nbsk ? t : c->bsk2[i - k - c->nbsk - 1]));
    fill_mod(mods[i], q, N, logN, tw, itw, twf, itwf, twd, itwd, twp, itwp);
    ulonglong2 *d_tw = nullptr, *d_itw = nullptr, *d_twf = nullptr, *d_itwf = nullptr, *d_twd = nullptr, *d_itwd = nullptr;
    ulonglong2 *d_twp = nullptr, *d_itwp = nullptr;
    if (!twd.empty()) { TRY(upload(c, &d_twd, twd)); TRY(upload(c, &d_itwd, itwd)); TRY(upload(c, &d_twp, twp)); TRY(upload(c, &d_itwp, itwp)); }
    mods[i].twd = d_twd; mods[i].itwd = d_itwd; mods[i].twp = d_twp; mods[i].itwp = d_itwp;
    TRY(upload(c, &d_tw, tw));
    TRY(upload(c, &d_itw, itw));
    if (!twf.empty()) { TRY(upload(c, &d_twf, twf)); TRY(upload(c, &d_itwf, itwf)); }
    mods[i].tw = d_tw; mods[i].itw = d_itw; mods[i].twf = d_twf; mods[i].itwf = d_itwf;
  }
  TRY(upload(c, &c->d_mods, mods));
  c->hmods = mods;
  c->ar_q = AR_F64;
  for (int i = 0; i < k; ++i) c->ar_q = std::min(c->ar_q, mods[i].ar_class);
  c->ar_t = mods[c->idx_t].ar_class;
  if (const char *e = getenv("ABC_FORCE_AR")) c->force_ar = atoi(e);
  c->ks_unmerged = getenv("ABC_KS_UNMERGED") != nullptr;
  c->ks_unfused = getenv("ABC_KS_UNFUSED") != nullptr;
  if (const char *e = getenv("ABC_KS_ONE_LAUNCH")) c->ks_one_launch = atoi(e) ? 1 : 0;
  if (getenv("ABC_KS_TWO_LAUNCH")) c->ks_one_launch = 0;
  c->ks_no_image = getenv("ABC_KS_NO_IMAGE") != nullptr;
  c->ks_no_discard = getenv("ABC_KS_NO_DISCARD") != nullptr;
  if (const char *e = getenv("ABC_KS_CHAIN")) c->ks_chain = atoi(e);
  if (const char *e = getenv("ABC_KS_RED")) c->ks_red = atoi(e);
  if (const char *e = getenv("ABC_KS14_SPLIT")) c->ks14 = atoi(e);
  if (const char *e = getenv("ABC_KS_PERSIST")) c->ks_persist = atoi(e);
  if (const char *e = getenv("ABC_BLOCK_CACHE_MIB")) c->blk_cache_limit = (size_t)atoll(e) << 20;   // 0: every release goes to the driver
  cudaDeviceGetAttribute(&c->n_sms, cudaDevAttrMultiProcessorCount, c->device);
  if (const char *e = getenv("ABC_KS_SPLIT_MAXB")) c->ks_split_maxb = atoi(e);
  if (const char *e = getenv("ABC_BEHZ_FUSED")) c->behz_fused = atoi(e) != 0;
  if (const char *e = getenv("ABC_KS_CHAIN_SKEW")) c->ks_chain_skew = atoi(e) < 1 ? 1 : atoi(e);
  if (const char *e = getenv("ABC_KS1_THREADS")) c->ks1_threads = atoi(e);
  if (const char *e = getenv("ABC_KS1_SKEW")) c->ks1_skew = atoi(e) < 0 ? 0 : atoi(e);
  c->lazy_rotate = getenv("ABC_EAGER_ROTATE") == nullptr;
  if (const char *e = getenv("ABC_ROT_CACHE")) c->rot_cache_max = atoi(e);
  c->no_square = getenv("ABC_NO_SQUARE") != nullptr;
  if (const char *e = getenv("ABC_KS_SKEW")) c->ks_skew = atoi(e) < 0 ? 0 : atoi(e);
  {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    c->prefetch_ahead = 0;  // > 0: a CTA also L2-prefetches the source row of the CTA this many blocks ahead (measured: no gain)
    if (const char *e = getenv("ABC_PREFETCH_AHEAD")) c->prefetch_ahead = atoi(e);
  }

  // ---- constants
  DevConst &C = c->hC;
  memset(&C, 0, sizeof C);
  C.N = N; C.logN = logN; C.k = k; C.L = L; C.nB = nB; C.nbsk = c->nbsk;
  for (int i = 0; i < k; ++i) { C.q[i] = c->primes[i]; barrett_ratio(C.q[i], C.q_mu_hi[i], C.q_mu_lo[i]); }
  C.t = t; C.t_half_up = (t + 1) >> 1; C.q_mod_t = prod_mod(Q, t);
  barrett_ratio(t, C.t_mu_hi, C.t_mu_lo);
  const u64 p = c->primes[k - 1];
  C.p = p; C.p_half = p >> 1; C.p_mu_hi = C.q_mu_hi[k - 1];
  C.gamma = c->gamma; C.gamma_half = c->gamma >> 1;
  barrett_ratio(c->gamma, C.g_mu_hi, C.g_mu_lo);
  const u64 mt = 1ull << 32;
  for (int i = 0; i < L; ++i) {
    const u64 qi = Q[i];
    const u64 inv_punct = invmod(prod_mod(Q, qi, i), qi);
    C.delta[i] = mulmod((qi - C.q_mod_t % qi) % qi, invmod(t % qi, qi), qi);
    C.inv_p[i] = invmod(p % qi, qi); C.inv_p_s[i] = shoup(C.inv_p[i], qi);
    C.p_half_mod_q[i] = C.p_half % qi; C.p_mod_q[i] = p % qi;
    C.dec_c[i] = mulmod(mulmod(t % qi, c->gamma % qi, qi), inv_punct, qi); C.dec_c_s[i] = shoup(C.dec_c[i], qi);
    C.punct_t[i] = prod_mod(Q, t, i); C.punct_g[i] = prod_mod(Q, c->gamma, i);
    C.lift_c[i] = mulmod(mt % qi, inv_punct, qi); C.lift_c_s[i] = shoup(C.lift_c[i], qi);
    C.punct_q_mt[i] = (u32)prod_mod(Q, mt, i);
    C.scale_c[i] = mulmod(t % qi, inv_punct, qi); C.scale_c_s[i] = shoup(C.scale_c[i], qi);
    C.B_mod_q[i] = prod_mod(Bv, qi); C.B_mod_q_s[i] = shoup(C.B_mod_q[i], qi);
    for (int j = 0; j < nB; ++j) C.punct_B_q[i][j] = prod_mod(Bv, qi, j);
  }
  C.p_mod_q[L] = 0;
  C.neg_inv_q_t = (t - invmod(prod_mod(Q, t), t)) % t; C.neg_inv_q_t_s = shoup(C.neg_inv_q_t, t);
  C.neg_inv_q_g = (c->gamma - invmod(prod_mod(Q, c->gamma), c->gamma)) % c->gamma;
  C.neg_inv_q_g_s = shoup(C.neg_inv_q_g, c->gamma);
  C.inv_g_t = invmod(c->gamma % t, t); C.inv_g_t_s = shoup(C.inv_g_t, t);
  C.neg_inv_q_mt = (u32)(mt - invmod(prod_mod(Q, mt), mt));
  for (int j = 0; j < c->nbsk; ++j) {
    const u64 pj = c->bsk[j];
    C.bsk[j] = pj; barrett_ratio(pj, C.bsk_mu_hi[j], C.bsk_mu_lo[j]);
    for (int i = 0; i < L; ++i) C.punct_q_bsk[j][i] = prod_mod(Q, pj, i);
    C.q_mod_bsk[j] = prod_mod(Q, pj);
    C.inv_mt_bsk[j] = invmod(mt % pj, pj); C.inv_mt_bsk_s[j] = shoup(C.inv_mt_bsk[j], pj);
    C.t_mod_bsk[j] = t % pj; C.t_mod_bsk_s[j] = shoup(C.t_mod_bsk[j], pj);
    C.inv_q_bsk[j] = invmod(C.q_mod_bsk[j], pj); C.inv_q_bsk_s[j] = shoup(C.inv_q_bsk[j], pj);
  }
  for (int j = 0; j < nB; ++j) {
    C.inv_punct_B[j] = invmod(prod_mod(Bv, Bv[j], j), Bv[j]); C.inv_punct_B_s[j] = shoup(C.inv_punct_B[j], Bv[j]);
    C.punct_B_msk[j] = prod_mod(Bv, c->msk, j);
  }
  C.inv_B_msk = invmod(prod_mod(Bv, c->msk), c->msk); C.inv_B_msk_s = shoup(C.inv_B_msk, c->msk);
  CK(cudaMalloc((void **)&c->dC, sizeof(DevConst)));
  c->owned.push_back(c->dC);
  CK(cudaMemcpy(c->dC, &C, sizeof(DevConst), cudaMemcpyHostToDevice));
  if (c->behz_f64) {   // the same constants as above for the sub-2^45 base, as exact doubles
    BehzF64 F;
    memset(&F, 0, sizeof F);
    const int nb2 = c->nbsk2 - 1;
    std::vector<u64> B2(c->bsk2.begin(), c->bsk2.begin() + nb2);
    const u64 msk2 = c->bsk2[nb2];
    F.L = L; F.nB = nb2; F.nbsk = c->nbsk2;
    for (int i = 0; i < L; ++i) {
      const u64 qi = Q[i];
      F.q[i] = (double)qi; F.qinv[i] = 1.0 / (double)qi;
      F.lift_c[i] = (double)C.lift_c[i]; F.scale_c[i] = (double)C.scale_c[i];
      F.punct_q_mt[i] = C.punct_q_mt[i];
      F.B_mod_q[i] = (double)prod_mod(B2, qi);
      for (int j = 0; j < nb2; ++j) F.punct_B_q[i][j] = (double)prod_mod(B2, qi, j);
    }
    F.neg_inv_q_mt = C.neg_inv_q_mt;
    for (int j = 0; j < c->nbsk2; ++j) {
      const u64 pj = c->bsk2[j];
      F.b[j] = (double)pj; F.binv[j] = 1.0 / (double)pj;
      for (int i = 0; i < L; ++i) F.punct_q_b[j][i] = (double)prod_mod(Q, pj, i);
      const u64 qmb = prod_mod(Q, pj);
      F.q_mod_b[j] = (double)qmb;
      F.inv_mt_b[j] = (double)invmod(mt % pj, pj);
      F.t_mod_b[j] = (double)(t % pj);
      F.inv_q_b[j] = (double)invmod(qmb, pj);
    }
    for (int j = 0; j < nb2; ++j) {
      F.inv_punct_B[j] = (double)invmod(prod_mod(B2, B2[j], j), B2[j]);
      F.punct_B_msk[j] = (double)prod_mod(B2, msk2, j);
    }
    F.inv_B_msk = (double)invmod(prod_mod(B2, msk2), msk2);
    F.msk_half = (double)(msk2 >> 1);
    CK(cudaMalloc((void **)&c->dF, sizeof(BehzF64)));
    c->owned.push_back(c->dF);
    CK(cudaMemcpy(c->dF, &F, sizeof(BehzF64), cudaMemcpyHostToDevice));
    std::vector<int> rm(4 * c->W2);
    for (int w = 0; w < 4 * c->W2; ++w) { const int r = w % c->W2; rm[w] = r < L ? r : c->idx_b2 + (r - L); }
    TRY(upload(c, &c->rm_behz2, rm));
    if (logN == 14) {
      const size_t nf = (size_t)c->B * 3 * c->W2 * 2;
      CK(cudaMalloc((void **)&c->ks14_xflag_behz, nf * sizeof(u32)));
      c->owned.push_back(c->ks14_xflag_behz);
      CK(cudaMemset(c->ks14_xflag_behz, 0, nf * sizeof(u32)));
    }
  }

  // ---- BatchEncoder index map (populate_matrix_reps_index_map)
  std::vector<u32> imap(N);
  {
    const u64 m = 2ull * N; u64 pos = 1; const int row = N >> 1;
    for (int i = 0; i < row; ++i) {
      imap[i] = bit_reverse((uint32_t)((pos - 1) >> 1), logN);
      imap[row | i] = bit_reverse((uint32_t)((m - pos - 1) >> 1), logN);
      pos = (pos * 3) & (m - 1);
    }
  }
  TRY(upload(c, &c->d_index_map, imap));

  // ---- row maps
  const int W = c->W;
  std::vector<int> v;
  v.resize(3 * L); for (int w = 0; w < 3 * L; ++w) v[w] = w % L;
  TRY(upload(c, &c->rm_ct, v));
  v.resize(2 * k); for (int w = 0; w < 2 * k; ++w) v[w] = w % k;
  TRY(upload(c, &c->rm_key, v));
  v.resize(4 * W); for (int w = 0; w < 4 * W; ++w) { int r = w % W; v[w] = r < L ? r : k + (r - L); }
  TRY(upload(c, &c->rm_behz, v));
  {
    std::vector<int> mq, dq, mb, db;
    for (int p = 0; p < 4; ++p) {
      for (int i = 0; i < L; ++i) { mq.push_back(i); dq.push_back(p * W + i); }
      for (int j = 0; j < c->nbsk; ++j) { mb.push_back(k + j); db.push_back(p * W + L + j); }
    }
    TRY(upload(c, &c->rm_behz_q, mq)); TRY(upload(c, &c->rd_behz_q, dq));
    TRY(upload(c, &c->rm_behz_b, mb)); TRY(upload(c, &c->rd_behz_b, db));
  }
  v.resize(k * L); for (int w = 0; w < k * L; ++w) v[w] = w / L;
  TRY(upload(c, &c->rm_modup, v));
  for (int w = 0; w < k * L; ++w) v[w] = w % L;
  TRY(upload(c, &c->rs_modup, v));
  v.assign(1, c->idx_t);
  TRY(upload(c, &c->rm_t, v));
  v.resize(L); for (int w = 0; w < L; ++w) v[w] = L + w;
  TRY(upload(c, &c->rs_c1, v));
  v.assign(2 * k, 0);
  TRY(upload(c, &c->rs_zero, v));
  v = {L, L};
  TRY(upload(c, &c->rm_special, v));
  v = {L, k + L};
  TRY(upload(c, &c->rd_special, v));
  v.resize(2 * L); for (int w = 0; w < 2 * L; ++w) v[w] = (w / L) * k + (w % L);
  TRY(upload(c, &c->rs_accq, v));
  c->own_lo = 0; c->own_hi = L;
  TRY(build_shard_maps(c));
  return ABC_OK;
}

// maps that depend on the set of limbs this rank owns (rebuilt by abc_comm_init)
abc_status build_shard_maps(abc_ctx *c) {
  const int L = c->L, k = c->k, lo = c->own_lo, hi = c->own_hi, nown = hi - lo;
  std::vector<int> I;
  for (int i = lo; i < hi; ++i) I.push_back(i);
  I.push_back(L);  // the special prime's accumulation is computed on every rank (ModDown needs it everywhere)
  c->ks_nI = (int)I.size();
  TRY(upload(c, &c->ks_I, I));
  std::vector<int> m, d, sr;
  for (int I_ : I) for (int J = 0; J < L; ++J) { m.push_back(I_); d.push_back(I_ * L + J); sr.push_back(J); }
  TRY(upload(c, &c->rm_modup_s, m)); TRY(upload(c, &c->rd_modup_s, d)); TRY(upload(c, &c->rs_modup_s, sr));
  {  // the same rows split by where their SOURCE limb lives: owned here (can start before the all-gather lands) or not
    std::vector<int> mo, dob, so, mt, dt, st;
    for (size_t w = 0; w < m.size(); ++w) {
      const bool mine = sr[w] >= lo && sr[w] < hi;
      (mine ? mo : mt).push_back(m[w]); (mine ? dob : dt).push_back(d[w]); (mine ? so : st).push_back(sr[w]);
    }
    c->n_up_own = (int)mo.size(); c->n_up_oth = (int)mt.size();
    if (c->n_up_own) { TRY(upload(c, &c->rm_up_own, mo)); TRY(upload(c, &c->rd_up_own, dob)); TRY(upload(c, &c->rs_up_own, so)); }
    if (c->n_up_oth) { TRY(upload(c, &c->rm_up_oth, mt)); TRY(upload(c, &c->rd_up_oth, dt)); TRY(upload(c, &c->rs_up_oth, st)); }
  }
  m.clear(); d.clear(); sr.clear();
  std::vector<int> od;
  for (int comp = 0; comp < 2; ++comp) for (int i = lo; i < hi; ++i) {
    m.push_back(i); d.push_back(comp * L + i); sr.push_back(comp * k + i); od.push_back(comp * L + i);
  }
  if (nown > 0) {
    TRY(upload(c, &c->rm_md, m)); TRY(upload(c, &c->rd_md, d)); TRY(upload(c, &c->rs_md, sr));
    TRY(upload(c, &c->rm_own, m)); TRY(upload(c, &c->rd_own, od));
    std::vector<int> mm = {L, L}, dm = {L, k + L}, sm_ = {L, k + L};
    mm.insert(mm.end(), m.begin(), m.end()); dm.insert(dm.end(), d.begin(), d.end()); sm_.insert(sm_.end(), sr.begin(), sr.end());
    TRY(upload(c, &c->rm_mdm, mm)); TRY(upload(c, &c->rd_mdm, dm)); TRY(upload(c, &c->rs_mdm, sm_));
  }
  if (nown > 0 && c->ks_nI * L < 256 && 2 * k < 256 && k * L < 256) {
    // chained key switch: ModUp rows of instance g, the two special-prime tail rows of instance g - (S1 - S2) and the
    // data tail rows of instance g - S1, for g = 0 .. B + S1 - 1 (every wait points at an earlier block).  Every entry
    // carries its row's modulus, destination row and source row (the row maps of the two-launch sequence, resolved here)
    // instances in flight: 16 at N <= 8192 (20 MiB of ModUp rows); fewer when an instance's block is larger (N = 16384,
    // k = 9: 9.4 MiB each) so that what is in flight stays L2-resident
    const size_t t_bytes = (size_t)c->ks_nI * L * c->N * 8;
    const int s1_fit = (int)std::max<size_t>(2, (size_t)(48u << 20) / std::max<size_t>(t_bytes, 1));
    const int Bn = c->B, S1 = getenv("ABC_KS_CHAIN_SKEW") ? c->ks_chain_skew : std::min(c->ks_chain_skew, s1_fit), S2 = std::min(c->ks_skew, S1 - 1) < 0 ? 0 : std::min(c->ks_skew, S1 - 1);
    std::vector<uint2> sch;
    auto up_row = [&](int g, int w) {      // w = idx(I) * L + J: T row I * L + J from target limb J, modulus I
      const int Iv = I[w / L], J = w % L;
      sch.push_back(make_uint2((u32)g, (u32)w | (u32)Iv << 8 | (u32)(Iv * L + J) << 16 | (u32)J << 24));
    };
    auto tail_row = [&](int g, int w) {    // w < 2: special-prime row of component w; else data row (comp, i)
      const int comp = w < 2 ? w : (w - 2) / nown, Iv = w < 2 ? L : lo + (w - 2) % nown;
      const int drow = w < 2 ? (w == 0 ? L : k + L) : comp * L + Iv;
      sch.push_back(make_uint2(1u << 31 | (u32)g, (u32)w | (u32)Iv << 8 | (u32)drow << 16 | (u32)(comp * k + Iv) << 24));
    };
    for (int g = 0; g < Bn + S1; ++g) {
      if (g < Bn) for (int w = 0; w < c->ks_nI * L; ++w) up_row(g, w);
      const int gs = g - (S1 - S2), gd = g - S1;
      if (gs >= 0 && gs < Bn) for (int w = 0; w < 2; ++w) tail_row(gs, w);
      if (gd >= 0 && gd < Bn) for (int w = 2; w < 2 + 2 * nown; ++w) tail_row(gd, w);
    }
    TRY(upload(c, &c->ks_sched, sch));
    c->ks_sched_n = (int)sch.size();
    CK(cudaMalloc((void **)&c->ks_done, (size_t)2 * c->B * c->k * sizeof(u32)));   // [B][k] rows stored + [B][k] rows consumed
    c->owned.push_back(c->ks_done);
    CK(cudaMemset(c->ks_done, 0, (size_t)2 * c->B * c->k * sizeof(u32)));
    c->ks_chain_serial = 0;
    // accumulating key switch: a ring of 2 * S1 instances' accumulators (a slot is reused once its tail rows, scheduled S1
    // instances after its ModUp rows, are done: every wait points at a smaller ticket)
    if (c->ksr_acc) { cudaStreamSynchronize(c->stream); cudaFree(c->ksr_acc); c->ksr_acc = nullptr; }
    if (c->ks_red) {
    c->ksr_ring = std::min(Bn, 2 * S1);
    const size_t accw = (size_t)c->ksr_ring * c->k * 2 * c->N;
    CK(cudaMalloc((void **)&c->ksr_acc, accw * sizeof(double)));
    CK(cudaMemset(c->ksr_acc, 0, accw * sizeof(double)));
    CK(cudaMalloc((void **)&c->ksr_done, (size_t)2 * c->B * c->k * sizeof(u32)));
    c->owned.push_back(c->ksr_done);
    CK(cudaMemset(c->ksr_done, 0, (size_t)2 * c->B * c->k * sizeof(u32)));
    c->ksr_freed = c->ksr_done + (size_t)c->B * c->k;
    c->ksr_serial = 0;
    }
  }
  // N = 8192: measured per batch (tools/ks_time.py, us per rotateRows, split vs chained): 1-2: 28.7 vs 32.8; 3-4: 34-37 vs 33;
  // 6: 37 vs 44; 8: 37 vs 47 -> split rows for B <= 2 and 5 <= B <= ks_split_maxb
  const bool split13 = c->logN == 13 && c->ks_split_maxb > 0 &&
                       (c->B <= 2 || (c->B >= 5 && c->B <= c->ks_split_maxb) || getenv("ABC_KS_SPLIT_FORCE") != nullptr);
  if (nown > 0 && ((c->logN == 14 && c->ks14) || split13) && c->ks_nI * L < 256 && 2 * k < 256) {
    // split key switch (ks14.cu): the chained schedule with every row as two half-rows holding adjacent tickets
    const int Bn = c->B;
    const size_t t_bytes = (size_t)c->ks_nI * L * c->N * 8;
    const int S1 = std::max(2, std::min(c->ks_chain_skew, (int)((size_t)(48u << 20) / std::max<size_t>(t_bytes, 1))));
    const int S2 = std::max(0, std::min(c->ks_skew, S1 - 1));
    std::vector<uint2> sch;
    auto up_row = [&](int g, int w) {
      const int Iv = I[w / L], J = w % L;
      for (u32 h = 0; h < 2; ++h)
        sch.push_back(make_uint2(h << 30 | (u32)g, (u32)w | (u32)Iv << 8 | (u32)(Iv * L + J) << 16 | (u32)J << 24));
    };
    auto tail_row = [&](int g, int w) {
      const int comp = w < 2 ? w : (w - 2) / nown, Iv = w < 2 ? L : lo + (w - 2) % nown;
      const int drow = w < 2 ? (w == 0 ? L : k + L) : comp * L + Iv;
      for (u32 h = 0; h < 2; ++h)
        sch.push_back(make_uint2(1u << 31 | h << 30 | (u32)g, (u32)w | (u32)Iv << 8 | (u32)drow << 16 | (u32)(comp * k + Iv) << 24));
    };
    for (int g = 0; g < Bn + S1; ++g) {
      if (g < Bn) for (int w = 0; w < c->ks_nI * L; ++w) up_row(g, w);
      const int gs = g - (S1 - S2), gd = g - S1;
      if (gs >= 0 && gs < Bn) for (int w = 0; w < 2; ++w) tail_row(gs, w);
      if (gd >= 0 && gd < Bn) for (int w = 2; w < 2 + 2 * nown; ++w) tail_row(gd, w);
    }
    TRY(upload(c, &c->ks14_sched, sch));
    c->ks14_sched_n = (int)sch.size();
    const size_t nflag = (size_t)c->B * (2 * k + 4 * k + 4);
    CK(cudaMalloc((void **)&c->ks14_done, nflag * sizeof(u32)));
    c->owned.push_back(c->ks14_done);
    CK(cudaMemset(c->ks14_done, 0, nflag * sizeof(u32)));
    c->ks14_xflag = c->ks14_done + (size_t)c->B * 2 * k;
    c->ks14_flags = c->ks14_xflag + (size_t)c->B * 4 * k;
    c->ks14_serial = 0;
  }
  if (!c->ks_fault_h) {
    CK(cudaHostAlloc((void **)&c->ks_fault_h, sizeof(u32), cudaHostAllocMapped));
    *c->ks_fault_h = 0;
    CK(cudaHostGetDevicePointer((void **)&c->ks_fault_d, c->ks_fault_h, 0));
  }
  if (!c->ks_flags) {
    CK(cudaMalloc((void **)&c->ks_flags, ((size_t)c->B * 2 + 1) * sizeof(u32)));
    c->owned.push_back(c->ks_flags);
    CK(cudaMemset(c->ks_flags, 0, ((size_t)c->B * 2 + 1) * sizeof(u32)));
    c->ks_ticket = c->ks_flags + (size_t)c->B * 2;
    c->ks_ticket_total = 0;
  }
  return ABC_OK;
}

// ---- NCCL, loaded on first use so single-GPU users need no NCCL at all (in a torch process this resolves to the
// libnccl.so.2 torch already loaded)
struct NcclApi {
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi *nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h) {
      api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
      api.GroupStart = (decltype(api.GroupStart))dlsym(h, "ncclGroupStart");
      api.GroupEnd = (decltype(api.GroupEnd))dlsym(h, "ncclGroupEnd");
      api.Broadcast = (decltype(api.Broadcast))dlsym(h, "ncclBroadcast");
      api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather");
      api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
      if (api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.GroupStart && api.GroupEnd && api.Broadcast && api.AllGather) api.h = h;
    }
  }
  return api.h ? &api : nullptr;
}
#define NCK(call)                                                                                     \
  do {                                                                                                \
    ncclResult_t r_ = (call);                                                                         \
    if (r_ != ncclSuccess) {                                                                          \
      c->err = std::string(#call) + ": " + (nccl_api()->GetErrorString ? nccl_api()->GetErrorString(r_) : "NCCL error"); \
      return ABC_ERR_CUDA;                                                                            \
    }                                                                                                 \
  } while (0)

void limb_range(int L, int world, int rank, int *lo, int *hi) {
  const int base = L / world, extra = L % world;
  *lo = rank * base + std::min(rank, extra);
  *hi = *lo + base + (rank < extra ? 1 : 0);
}

// all-gather of the limb-sharded polynomials `poly_mask` (bit 0: c0, bit 1: c1) of ciphertext block d, in place: the
// exchange step of the limb-sharded key switch (NCCL over NVLink).  Owned limb ranges are contiguous and ordered by rank,
// so with an even split (L % world == 0) a polynomial is ONE in-place ncclAllGather; an uneven split falls back to one
// broadcast per rank.
abc_status allgather_limbs(abc_ctx *c, u64 *d, int poly_mask, cudaStream_t stream = nullptr) {
  if (c->world == 1) return ABC_OK;
  if (!stream) stream = c->stream;
  NcclApi *n = nccl_api();
  const size_t N = c->N, L = c->L;
  const bool even = c->L % c->world == 0;
  c->launches++;
  NCK(n->GroupStart());
  for (int inst = 0; inst < c->B; ++inst)
    for (int p = 0; p < 2; ++p) {
      if (!(poly_mask & (1 << p))) continue;
      u64 *poly = d + ((size_t)inst * 2 + p) * L * N;
      if (even) {
        const size_t cnt = (L / c->world) * N;
        NCK(n->AllGather(poly + (size_t)c->rank * cnt, poly, cnt, ncclUint64, c->comm, stream));
        c->gather_calls++;
      } else {
        for (int r = 0; r < c->world; ++r) {
          int lo, hi;
          limb_range(c->L, c->world, r, &lo, &hi);
          if (hi == lo) continue;
          NCK(n->Broadcast(poly + (size_t)lo * N, poly + (size_t)lo * N, (size_t)(hi - lo) * N, ncclUint64, r, c->comm, stream));
          c->gather_calls++;
        }
      }
      c->gathered_bytes += (uint64_t)(L - (c->own_hi - c->own_lo)) * N * 8;
    }
  NCK(n->GroupEnd());
  return ABC_OK;
}
// all-gather by COLUMNS: `rows` rows of N words starting at `base` (row stride N, per-instance stride inst_stride), of
// which this rank has filled coefficients [rank * N / world, (rank + 1) * N / world) — the exchange after a
// coefficient-sharded base conversion.  The slices are packed into one contiguous block, exchanged with ONE
// ncclAllGather, and the other ranks' slices are scattered back into the rows (two copy kernels at HBM speed instead of
// one small collective per row: 125 collectives per multiply cost as much as the conversions saved).
// (a block of rows = `groups` groups of `rows` consecutive rows, group stride group_stride: e.g. the Bsk rows of the 2 or 4
// operand polynomials of the BEHZ block; blockIdx.z = inst * groups + group)
__global__ void k_cols_pack(const u64 *__restrict__ base, u64 *__restrict__ send, int rows, int N, int ncols, int col0,
                            size_t inst_stride, int groups, size_t group_stride) {
  const int r = blockIdx.y, inst = blockIdx.z / groups, g = blockIdx.z % groups;
  const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(base + (size_t)inst * inst_stride + (size_t)g * group_stride + (size_t)r * N + col0);
  ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(send + ((size_t)blockIdx.z * rows + r) * ncols);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ncols / 2; i += gridDim.x * blockDim.x) dst[i] = src[i];
}
__global__ void k_cols_unpack(u64 *__restrict__ base, const u64 *__restrict__ all, int rows, int N, int ncols, int world,
                              int rank, size_t inst_stride, int groups, size_t group_stride) {
  const int r = blockIdx.y, inst = blockIdx.z / groups, g = blockIdx.z % groups;
  for (int w = 0; w < world; ++w) {
    if (w == rank) continue;
    const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(all + (((size_t)w * gridDim.z + blockIdx.z) * rows + r) * ncols);
    ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(base + (size_t)inst * inst_stride + (size_t)g * group_stride + (size_t)r * N + (size_t)w * ncols);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ncols / 2; i += gridDim.x * blockDim.x) dst[i] = src[i];
  }
}
abc_status allgather_columns(abc_ctx *c, u64 *base, size_t rows, size_t inst_stride, int groups = 1, size_t group_stride = 0) {
  if (c->world == 1) return ABC_OK;
  NcclApi *n = nccl_api();
  const size_t N = c->N, cnt = N / c->world, per_rank = (size_t)c->B * groups * rows * cnt;
  u64 *send = nullptr, *all = nullptr;
  TRY(scratch(c, SC_COMM_SEND, &send, per_rank));
  TRY(scratch(c, SC_COMM_ALL, &all, per_rank * c->world));
  const dim3 grid((unsigned)std::max<size_t>(1, std::min<size_t>(8, cnt / 2 / 256)), (unsigned)rows, (unsigned)(c->B * groups));
  {
    Launch l(c, "shard_cols_pack");
    k_cols_pack<<<grid, 256, 0, c->stream>>>(base, send, (int)rows, (int)N, (int)cnt, (int)(c->rank * cnt), inst_stride, groups, group_stride);
    CK(cudaGetLastError());
  }
  c->launches++;
  NCK(n->AllGather(send, all, per_rank, ncclUint64, c->comm, c->stream));
  c->gather_calls++;
  c->gathered_bytes += (uint64_t)per_rank * (c->world - 1) * 8;
  {
    Launch l(c, "shard_cols_unpack");
    k_cols_unpack<<<grid, 256, 0, c->stream>>>(base, all, (int)rows, (int)N, (int)cnt, c->world, c->rank, inst_stride, groups, group_stride);
    CK(cudaGetLastError());
  }
  return ABC_OK;
}

// ---- op building blocks ------------------------------------------------------------------------
__global__ void k_key_to_f64(const u64 *__restrict__ in, double *__restrict__ out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = (double)in[i];
}
void drop_key_f64(abc_ctx *c) {   // keys are about to change
  rot_cache_clear(c);             // cached rotation prefixes were computed under the old keys
  if (c->key_f64.empty()) return;
  cudaStreamSynchronize(c->stream);
  for (auto &kv : c->key_f64) cudaFree(kv.second);
  c->key_f64.clear();
}
// the key as exact doubles (every residue < 2^45), same layout; converted once per key, dropped when keys change
abc_status key_as_f64(abc_ctx *c, const u64 *key, const double **out) {
  auto it = c->key_f64.find(key);
  if (it == c->key_f64.end()) {
    const size_t n = (size_t)c->L * 2 * c->k * c->N;
    double *d = nullptr;
    CK(cudaMalloc((void **)&d, n * 8));
    Launch l(c, "key_to_f64");
    k_key_to_f64<<<592, 256, 0, c->stream>>>(key, d, n);
    CK(cudaGetLastError());
    it = c->key_f64.emplace(key, d).first;
  }
  *out = it->second;
  return ABC_OK;
}

size_t ct_words1(const abc_ctx *c) { return (size_t)2 * c->L * c->N; }

// Evaluator::switch_key_inplace: dst[inst][2][L][N] = sigma(base0, base1) + KeySwitch(sigma(target)), where sigma is the
// Galois automorphism with inverse element einv (0: identity).  dst must not alias target/base when einv != 0.
//   1. ModUp: every target limb J reduced mod every key-level prime I (+ automorphism gather), forward NTT   [limb pipeline]
//   2. inner product with the key over J, 128-bit lazy accumulation                                          [k_ks_inner]
//   3. INTT of the two special-prime rows                                                                    [limb pipeline]
//   4. INTT of the 2L data rows fused with ModDown (rounded division by p) and the base accumulate            [limb pipeline]
abc_status keyswitch(abc_ctx *c, const u64 *target, long long target_is, const u64 *key, const u64 *base0,
                     long long base0_is, const u64 *base1, long long base1_is, u32 einv, u64 *dst,
                     const u64 *addend = nullptr, u64 *dst_plain = nullptr, u64 *gather_ct = nullptr) {
  // gather_ct (limb-sharded contexts): the ciphertext block whose c1 is `target`; its limbs are all-gathered in front of
  // ModUp.  With overlap on, the exchange runs on the communication stream while the ModUp rows whose source limb is
  // local are already transforming; the rows fed by remote limbs follow once it has landed.
  ++c->key_switches;
  const int N = c->N, L = c->L, k = c->k, B = c->B;
  u64 *T = nullptr, *acc = nullptr;
  // exact-double class: the whole key switch can be one launch (ksfused.cu), nothing but INTT_p(acc_L) goes through HBM.
  // N = 4096: two CTAs per SM, 20 % faster than ModUp launch + tail launch (measured, op_microbench).  N = 8192: its
  // accumulators leave room for one CTA per SM only, which makes it a tie (-1 .. +4 % depending on the batch) at a
  // third of the HBM traffic; the two-launch sequence stays the default there, ABC_KS_ONE_LAUNCH=1 selects this one.
  const bool one_launch = c->ks_one_launch >= 0 ? c->ks_one_launch == 1 : c->logN <= 12;
  if (one_launch && c->logN <= 13 && (c->own_hi - c->own_lo) > 0 && abc_ntt_arith_class(c) == AR_F64 && !c->ks_unmerged &&
      !c->ks_unfused) {
    if (gather_ct && c->world > 1) TRY(allgather_limbs(c, gather_ct, 2));
    TRY(scratch(c, SC_ACC, &acc, (size_t)B * 2 * N));
    KsJob kj;
    memset(&kj, 0, sizeof kj);
    const double *keyd = nullptr;
    TRY(key_as_f64(c, key, &keyd));
    kj.target = target; kj.target_is = target_is; kj.key = keyd;
    kj.dst = dst; kj.dst_is = (long long)2 * L * N; kj.dst2 = dst_plain; kj.add = addend; kj.add_is = (long long)2 * L * N;
    kj.base0 = base0; kj.base0_is = base0_is; kj.base1 = base1; kj.base1_is = base1_is; kj.einv = einv;
    kj.tl = acc; kj.tl_is = (long long)2 * N;
    kj.flags = c->ks_flags; kj.serial = ++c->ks_serial; kj.skew = c->ks1_skew; kj.fault = c->ks_fault_d;
    kj.ticket = c->ks_ticket; kj.ticket_base = c->ks_ticket_total; c->ks_ticket_total += (u32)(B * c->ks_nI);
    kj.C = c->dC; kj.Iset = c->ks_I; kj.nI = c->ks_nI; kj.L = L; kj.k = k; kj.B = B; kj.threads = c->ks1_threads;
    Launch l(c, "ks_fused");
    const int e = ks_fused_launch(c->logN, kj, c->d_mods, c->stream);
    if (e != 0) { c->err = std::string("ks_fused: ") + cudaGetErrorString((cudaError_t)e); return ABC_ERR_CUDA; }
    return ABC_OK;
  }
  const bool merged = c->logN <= 14 && (c->own_hi - c->own_lo) > 0 && !c->ks_unmerged;
  // exact-double class: the inner product runs in the load of the INTT + ModDown launch (no accumulator round trip)
  const bool fused = merged && abc_ntt_arith_class(c) == AR_F64 && !c->ks_unfused;
  const int t_image = fused && c->logN <= 14 && !c->ks_no_image ? 1 : 0;  // T rows as bulk-stored images of the swizzled limb
  const bool chain = t_image && c->ks_chain && c->ks_sched && c->logN <= 14 && c->force_ar < 0;
  const bool red = chain && c->ks_red && c->ksr_acc && c->logN <= 13;   // no ModUp block at all (ksred.cu)
  if (chain && c->ks14_sched && (c->logN == 14 || c->logN == 13) && c->ks_one_launch != 1 && !c->ks_red) {
    // rows of half a limb (ks14.cu): N = 16384 (two CTAs per SM instead of one), N = 8192 at small batch (latency)
    if (gather_ct && c->world > 1) TRY(allgather_limbs(c, gather_ct, 2));
    Ks14 kq;
    memset(&kq, 0, sizeof kq);
    const double *keyd = nullptr;
    TRY(key_as_f64(c, key, &keyd));
    u64 *sx = nullptr, *xch = nullptr;
    TRY(scratch(c, SC_T, &T, (size_t)B * k * L * N));
    TRY(scratch(c, SC_ACC, &acc, (size_t)B * 2 * k * N));
    TRY(scratch(c, SC_SX, &sx, (size_t)B * L * N));
    TRY(scratch(c, SC_XCH, &xch, (size_t)B * 2 * k * N));
    kq.target = target; kq.target_is = target_is; kq.sx = reinterpret_cast<double *>(sx); kq.key = keyd;
    kq.T = reinterpret_cast<double *>(T); kq.xch = reinterpret_cast<double *>(xch);
    kq.tl = acc; kq.tl_is = (long long)2 * k * N;
    kq.dst = dst; kq.dst_is = (long long)2 * L * N; kq.dst2 = dst_plain; kq.add = addend; kq.add_is = (long long)2 * L * N;
    kq.base0 = base0; kq.base0_is = base0_is; kq.base1 = base1; kq.base1_is = base1_is; kq.einv = einv;
    kq.sched = c->ks14_sched; kq.n_blocks = c->ks14_sched_n;
    kq.ticket = c->ks_ticket; kq.ticket_base = c->ks_ticket_total; c->ks_ticket_total += (u32)kq.n_blocks;
    kq.serial = ++c->ks14_serial;
    kq.done = c->ks14_done; kq.done_target = (u32)L * kq.serial;
    kq.xflag = c->ks14_xflag; kq.flags = c->ks14_flags; kq.fault = c->ks_fault_d;
    kq.C = c->dC; kq.L = L; kq.k = k; kq.B = B;
    {
      Launch l(c, "ks14_prep");
      const int e = ks14_prep_launch(c->logN, kq, c->d_mods, c->stream);
      if (e != 0) { c->err = std::string("ks14_prep: ") + cudaGetErrorString((cudaError_t)e); return ABC_ERR_CUDA; }
    }
    Launch l(c, einv ? "ks14" : "ks14_relin");
    const int e = ks14_launch(c->logN, kq, c->d_mods, c->stream);
    if (e != 0) { c->err = std::string("ks14: ") + cudaGetErrorString((cudaError_t)e); return ABC_ERR_CUDA; }
    return ABC_OK;
  }
  if (!red) TRY(scratch(c, SC_T, &T, (size_t)B * k * L * N));
  TRY(scratch(c, SC_ACC, &acc, (size_t)B * 2 * k * N));
  LimbJob j = blank_job();
  j.dst = T; j.dst_is = (long long)k * L * N; j.src = target; j.src_is = target_is;
  j.rowmod = c->rm_modup_s; j.rowdst = c->rd_modup_s; j.rowsrc = c->rs_modup_s; j.galois_einv = einv;
  j.t_image = t_image;
  LimbJob jup = j;
  if (gather_ct && c->world > 1) {
    const int combo = einv ? LIMB_GALOIS_REDUCE_FWD : LIMB_REDUCE_FWD;
    if (c->shard_overlap && c->comm_stream && c->n_up_own > 0 && c->n_up_oth > 0 && !chain) {
      CK(cudaEventRecord(c->comm_ready, c->stream));                  // the producer of target has been enqueued
      CK(cudaStreamWaitEvent(c->comm_stream, c->comm_ready, 0));
      TRY(allgather_limbs(c, gather_ct, 2, c->comm_stream));
      CK(cudaEventRecord(c->comm_done, c->comm_stream));
      LimbJob jo = j;
      jo.rowmod = c->rm_up_own; jo.rowdst = c->rd_up_own; jo.rowsrc = c->rs_up_own;
      TRY(launch_limb(c, combo, c->ar_q, jo, c->n_up_own, B, "ks_modup_ntt_local"));
      CK(cudaStreamWaitEvent(c->stream, c->comm_done, 0));
      jo.rowmod = c->rm_up_oth; jo.rowdst = c->rd_up_oth; jo.rowsrc = c->rs_up_oth;
      TRY(launch_limb(c, combo, c->ar_q, jo, c->n_up_oth, B, "ks_modup_ntt_remote"));
    } else {
      TRY(allgather_limbs(c, gather_ct, 2));
      if (!chain) TRY(launch_limb(c, combo, c->ar_q, j, c->ks_nI * L, B, "ks_modup_ntt"));
    }
  } else if (!chain) TRY(launch_limb(c, einv ? LIMB_GALOIS_REDUCE_FWD : LIMB_REDUCE_FWD, c->ar_q, j, c->ks_nI * L, B, "ks_modup_ntt"));
  if (!fused) {
    Launch l(c, "ks_inner");
    DISPATCH_L(c, (k_ks_inner<LL><<<dim3(N / 512, c->ks_nI, B), 256, 0, c->stream>>>(T, key, acc, c->d_mods, N, k, c->L, c->ks_I)));
    CK(cudaGetLastError());
  }
  if (!merged) {
    j = blank_job();
    j.dst = acc; j.src = acc; j.dst_is = j.src_is = (long long)2 * k * N;
    j.rowmod = c->rm_special; j.rowdst = c->rd_special;
    TRY(launch_limb(c, LIMB_INV, c->ar_q, j, 2, B, "ks_intt_special"));
  }
  j = blank_job();
  const int nown = c->own_hi - c->own_lo;
  j.src = acc; j.src_is = (long long)2 * k * N; j.rowsrc = c->rs_md; j.rowdst = c->rd_md; j.rowmod = c->rm_md;
  j.dst = dst; j.dst_is = (long long)2 * L * N;
  j.C = c->dC; j.tl = acc; j.tl_is = (long long)2 * k * N; j.L = L; j.k = k; j.i0 = c->own_lo; j.nrows = nown;
  j.base0 = base0; j.base0_is = base0_is; j.base1 = base1; j.base1_is = base1_is; j.base_einv = einv;
  j.add = addend; j.add_is = (long long)2 * L * N;  // a whole ciphertext accumulated into the result (rotate + add)
  j.dst2 = dst_plain;                                // ... and the result without it (same layout as dst)
  if (merged) {  // one launch: the two special-prime rows INTT and publish, the data rows INTT, wait, ModDown
    j.rowsrc = c->rs_mdm; j.rowdst = c->rd_mdm; j.rowmod = c->rm_mdm;
    j.flags = c->ks_flags; j.flag_serial = ++c->ks_serial; j.skew = c->ks_skew; j.fault = c->ks_fault_d;
    j.ticket = c->ks_ticket; j.ticket_base = c->ks_ticket_total;
    if (fused) {
      j.src = T; j.src_is = (long long)k * L * N; j.mul = key; j.t_image = t_image;
      if (t_image) {  // raw-double T rows are multiplied with the exact-double copy of the key
        const double *keyd = nullptr;
        TRY(key_as_f64(c, key, &keyd));
        j.mul = reinterpret_cast<const u64 *>(keyd);
      }
      if (red) {   // ModUp rows accumulate into the output accumulators: no T (ksred.cu)
        KsRed kr;
        memset(&kr, 0, sizeof kr);
        const double *keyd = nullptr;
        TRY(key_as_f64(c, key, &keyd));
        kr.target = target; kr.target_is = target_is; kr.key = keyd;
        kr.acc = c->ksr_acc; kr.ring = c->ksr_ring; kr.tl = acc; kr.tl_is = (long long)2 * k * N;
        kr.dst = dst; kr.dst_is = (long long)2 * L * N; kr.dst2 = dst_plain; kr.add = addend; kr.add_is = (long long)2 * L * N;
        kr.base0 = base0; kr.base0_is = base0_is; kr.base1 = base1; kr.base1_is = base1_is; kr.einv = einv;
        kr.sched = c->ks_sched; kr.n_blocks = c->ks_sched_n;
        kr.ticket = c->ks_ticket; kr.ticket_base = c->ks_ticket_total; c->ks_ticket_total += (u32)kr.n_blocks;
        ++c->ksr_serial;
        kr.done = c->ksr_done; kr.done_target = (u32)L * c->ksr_serial;
        kr.freed = c->ksr_freed; kr.freed_target = 2u * c->ksr_serial;
        kr.flags = c->ks_flags; kr.flag_serial = j.flag_serial; kr.fault = c->ks_fault_d;
        kr.C = c->dC; kr.L = L; kr.k = k; kr.B = B;
        Launch l(c, einv ? "ks_red" : "ks_red_relin");
        const int e = ks_red_launch(c->logN, kr, c->d_mods, c->stream);
        if (e != 0) { c->err = std::string("ks_red: ") + cudaGetErrorString((cudaError_t)e); return ABC_ERR_CUDA; }
        return ABC_OK;
      }
      if (chain) {
        KsChain ch;
        ch.up = jup; ch.tail = j; ch.sched = c->ks_sched; ch.n_blocks = c->ks_sched_n;
        ch.up.n = ch.tail.n = N; ch.up.sub = ch.tail.sub = 0; ch.up.prefetch_ahead = 0; ch.tail.prefetch_ahead = c->prefetch_ahead;
        ch.up.k = k; ch.up.L = L;
        ch.up.done = ch.tail.done = c->ks_done;
        ch.tail.t_used = c->ks_no_discard ? nullptr : c->ks_done + (size_t)c->B * k;
        ch.tail.done_target = ch.up.done_target = (u32)L * ++c->ks_chain_serial;
        ch.ticket = c->ks_ticket; ch.ticket_base = c->ks_ticket_total; c->ks_ticket_total += (u32)ch.n_blocks;
        Launch l(c, einv ? "ks_chain" : "ks_chain_relin");
        const int pgrid = c->ks_persist ? std::min(ch.n_blocks, 2 * c->n_sms) : 0;
        if (pgrid > 0 && c->logN <= 13) {
          ch.up.persist = ch.tail.persist = 1;
          c->ks_ticket_total += (u32)pgrid;   // every CTA takes one ticket past the end
          const int e = ks_chain_launch_persistent(c->logN, ch, c->d_mods, c->stream, pgrid);
          if (e != 0) { c->err = std::string("ks_chain (persistent): ") + cudaGetErrorString((cudaError_t)e); return ABC_ERR_CUDA; }
          return ABC_OK;
        }
        const int e = ks_chain_launch(c->logN, ch, c->d_mods, c->stream);
        if (e != 0) { c->err = std::string("ks_chain: ") + cudaGetErrorString((cudaError_t)e); return ABC_ERR_CUDA; }
        return ABC_OK;
      }
      c->ks_ticket_total += (u32)((2 + 2 * nown) * B);
      TRY(launch_limb(c, LIMB_KSINNER_INV_MODDOWN, c->ar_q, j, 2 + 2 * nown, B, "ks_inner_intt_moddown"));
    } else {
      c->ks_ticket_total += (u32)((2 + 2 * nown) * B);
      TRY(launch_limb(c, LIMB_INV_MODDOWN, c->ar_q, j, 2 + 2 * nown, B, "ks_intt_moddown"));
    }
  } else if (nown > 0) {
    TRY(launch_limb(c, LIMB_INV_MODDOWN, c->ar_q, j, 2 * nown, B, "ks_intt_moddown"));
  }
  return ABC_OK;
}

// Evaluator::bfv_multiply (size 2 x size 2 -> size 3): out3 [B][3][L][N]
#define DISPATCH_BF(c, EXPR)                                                                         \
  switch ((c)->L * 16 + (c)->nbsk2) {                                                               \
    case 2 * 16 + 4: { constexpr int LL = 2, NK = 4; EXPR; } break;                                 \
    case 2 * 16 + 5: { constexpr int LL = 2, NK = 5; EXPR; } break;                                 \
    case 3 * 16 + 5: { constexpr int LL = 3, NK = 5; EXPR; } break;                                 \
    case 4 * 16 + 6: { constexpr int LL = 4, NK = 6; EXPR; } break;                                 \
    case 4 * 16 + 7: { constexpr int LL = 4, NK = 7; EXPR; } break;                                 \
    case 8 * 16 + 11: { constexpr int LL = 8, NK = 11; EXPR; } break;                               \
    default: return fail(c, ABC_ERR_UNSUPPORTED, "FP64 BEHZ: no kernel for this (L, |Bsk|)");       \
  }
bool behz_f64_has_kernel(const abc_ctx *c) {
  const int key = c->L * 16 + c->nbsk2;
  return key == 2 * 16 + 4 || key == 2 * 16 + 5 || key == 3 * 16 + 5 || key == 4 * 16 + 6 || key == 4 * 16 + 7 || key == 8 * 16 + 11;
}
// the same product over the sub-2^45 auxiliary base, every transform on the exact-double class (behz_f64.cuh)
abc_status behz_multiply_f64(abc_ctx *c, const u64 *a, const u64 *b, u64 *out3) {
  const int N = c->N, B = c->B, W = c->W2;
  u64 *X = nullptr;
  TRY(scratch(c, SC_X, &X, (size_t)B * 4 * W * N));
  const bool square = a == b && !c->no_square;
  const int np = square ? 2 : 4;
  {
    Launch l(c, "behz_lift");
    DISPATCH_BF(c, (k_behz_lift_f64<LL, NK><<<dim3(N / 128, np, B), 128, 0, c->stream>>>(a, b, X, c->dF, N, c->behz_fused ? 0 : 1)));
    CK(cudaGetLastError());
  }
  LimbJob j = blank_job();
  j.dst = X; j.src = X; j.dst_is = j.src_is = (long long)4 * W * N; j.rowmod = c->rm_behz2;
  if (c->behz_fused) {
    // forward rows like the key switch's ModUp rows: bulk copy in, conversion in the first pass, raw-double image out
    // (reduced to |x| <= 0.5 q: the images feed products); then the tensor product in the LOAD of the inverse transforms
    // (no tensor launch, no round trip of the products through HBM)
    u64 *Y = nullptr;
    TRY(scratch(c, SC_Y, &Y, (size_t)B * 3 * W * N));
    if (c->logN == 14 && c->ks14 && c->ks14_xflag_behz) {   // N = 16384: the block's transforms on rows of half a limb (ks14.cu)
      u64 *XI = nullptr, *xch = nullptr;
      TRY(scratch(c, SC_XI, &XI, (size_t)B * 4 * W * N));
      TRY(scratch(c, SC_XCH, &xch, (size_t)B * std::max(3 * W, 2 * c->k) * N));
      Behz14 bz;
      memset(&bz, 0, sizeof bz);
      bz.X = X; bz.X_is = (long long)4 * W * N; bz.XI = XI; bz.XI_is = (long long)4 * W * N; bz.a = a; bz.b = b;
      bz.Y = Y; bz.Y_is = (long long)3 * W * N; bz.xch = reinterpret_cast<double *>(xch);
      bz.xflag = c->ks14_xflag_behz; bz.serial = ++c->ks14_behz_serial; bz.fault = c->ks_fault_d;
      bz.rowmod = c->rm_behz2; bz.W = W; bz.L = c->L; bz.np = np; bz.square = square ? 1 : 0; bz.B = B;
      {
        Launch l(c, "behz14_ntt");
        const int e = behz14_fwd_launch(bz, c->d_mods, c->stream);
        if (e != 0) { c->err = std::string("behz14_ntt: ") + cudaGetErrorString((cudaError_t)e); return ABC_ERR_CUDA; }
      }
      {
        bz.ticket = c->ks_ticket; bz.ticket_base = c->ks_ticket_total; c->ks_ticket_total += (u32)(B * 3 * W * 2);
        Launch l(c, "behz14_tensor_intt");
        const int e = behz14_inv_launch(bz, c->d_mods, c->stream);
        if (e != 0) { c->err = std::string("behz14_tensor_intt: ") + cudaGetErrorString((cudaError_t)e); return ABC_ERR_CUDA; }
      }
      Launch l(c, "behz_scale");
      DISPATCH_BF(c, (k_behz_scale_f64<LL, NK><<<dim3(N / 128, 3, B), 128, 0, c->stream>>>(Y, out3, c->dF, N, 3)));
      CK(cudaGetLastError());
      return ABC_OK;
    }
    j.t_image = 1; j.src_same_mod = 1; j.raw_reduce = 1;
    j.bz_W = W; j.bz_a = a; j.bz_b = b; j.L = c->L;
    TRY(launch_limb(c, LIMB_REDUCE_FWD, AR_F64, j, np * W, B, "behz_ntt"));
    LimbJob jt = blank_job();
    jt.src = X; jt.src_is = (long long)4 * W * N; jt.dst = Y; jt.dst_is = (long long)3 * W * N; jt.rowmod = c->rm_behz2;
    jt.bz_W = W; jt.bz_square = square ? 1 : 0;
    TRY(launch_limb(c, LIMB_BEHZTENSOR_INV, AR_F64, jt, 3 * W, B, "behz_tensor_intt"));
    Launch l(c, "behz_scale");
    DISPATCH_BF(c, (k_behz_scale_f64<LL, NK><<<dim3(N / 128, 3, B), 128, 0, c->stream>>>(Y, out3, c->dF, N, 3)));
    CK(cudaGetLastError());
    return ABC_OK;
  }
  TRY(launch_limb(c, LIMB_FWD, AR_F64, j, np * W, B, "behz_ntt"));
  {
    Launch l(c, "behz_tensor");
    k_behz_tensor_f64<<<dim3(N / 256, W, B), 256, 0, c->stream>>>(X, c->d_mods, c->rm_behz2, N, W, square ? 1 : 0);
    CK(cudaGetLastError());
  }
  TRY(launch_limb(c, LIMB_INV, AR_F64, j, 3 * W, B, "behz_intt"));
  {
    Launch l(c, "behz_scale");
    DISPATCH_BF(c, (k_behz_scale_f64<LL, NK><<<dim3(N / 128, 3, B), 128, 0, c->stream>>>(X, out3, c->dF, N)));
    CK(cudaGetLastError());
  }
  return ABC_OK;
}

abc_status behz_multiply(abc_ctx *c, const u64 *a, const u64 *b, u64 *out3) {
  if (c->behz_f64 && c->world == 1 && c->force_ar < 0 && behz_f64_has_kernel(c)) return behz_multiply_f64(c, a, b, out3);
  const int N = c->N, L = c->L, B = c->B, W = c->W;
  u64 *X = nullptr;
  TRY(scratch(c, SC_X, &X, (size_t)B * 4 * W * N));
  // multiply(x, x) (e.g. `d *** d` of the distance programs): the second operand's lift and forward transforms would
  // repeat the first's, so only polys 0,1 are prepared; the products are the same modular expressions, bit for bit
  const bool square = a == b && !c->no_square;
  const int np = square ? 2 : 4;
  // Limb-sharded contexts: the base conversions (O(L^2) modular products per coefficient, half of the product's time at
  // L = 30) are independent per coefficient, so every rank converts N / world coefficients and the ranks exchange the
  // converted columns (in-place all-gathers of row slices); the transforms in between run on whole rows on every rank.
  const bool cols = c->world > 1 && c->shard_cols && (N / c->world) % 128 == 0;
  const int ncols = cols ? N / c->world : N, col0 = cols ? c->rank * ncols : 0;
  // threads per block of the conversion kernels: a column slice of a single ciphertext is few coefficients (8192 at world
  // = 8) of heavy, latency-bound work each, so small blocks spread them over all SMs (192 blocks of 128 threads left most idle)
  const int cvt = (size_t)ncols * B * 3 >= (size_t)148 * 2 * 128 * 4 ? 128 : 32;
  {
    Launch l(c, "behz_lift");
    DISPATCH_L(c, (k_behz_lift<LL><<<dim3(ncols / cvt, np, B), cvt, 0, c->stream>>>(a, b, X, c->dC, N, c->L, col0, cols ? 0 : 1)));
    CK(cudaGetLastError());
  }
  if (cols) {
    for (int inst = 0; inst < B; ++inst)
      for (int p = 0; p < np; ++p) {   // q rows of X = the operand polynomials (every rank holds them whole)
        const u64 *srcp = (p < 2 ? a : b) + ((size_t)inst * 2 + (p & 1)) * L * N;
        CK(cudaMemcpyAsync(X + ((size_t)inst * 4 + p) * W * N, srcp, (size_t)L * N * 8, cudaMemcpyDeviceToDevice, c->stream));
      }
    // Bsk rows of the np operand polynomials: every rank's column slice to everybody, one exchange
    TRY(allgather_columns(c, X + (size_t)L * N, (size_t)c->nbsk, (size_t)4 * W * N, np, (size_t)W * N));
  }
  LimbJob j = blank_job();
  j.dst = X; j.src = X; j.dst_is = j.src_is = (long long)4 * W * N; j.rowmod = c->rm_behz;
  LimbJob jq = j, jb = j;
  jq.rowmod = c->rm_behz_q; jq.rowdst = c->rd_behz_q;
  jb.rowmod = c->rm_behz_b; jb.rowdst = c->rd_behz_b;
  if (c->ar_q == AR_SHOUP) {
    TRY(launch_limb(c, LIMB_FWD, AR_SHOUP, j, np * W, B, "behz_ntt"));
  } else {  // q rows on the FP64-assisted class, the 61-bit Bsk rows on the Shoup class
    TRY(launch_limb(c, LIMB_FWD, c->ar_q, jq, np * L, B, "behz_ntt_q"));
    TRY(launch_limb(c, LIMB_FWD, AR_SHOUP, jb, np * c->nbsk, B, "behz_ntt_bsk"));
  }
  {
    Launch l(c, "behz_tensor");
    k_behz_tensor<<<dim3(N / 256, W, B), 256, 0, c->stream>>>(X, c->d_mods, c->rm_behz, N, W, square ? 1 : 0);
    CK(cudaGetLastError());
  }
  if (c->ar_q == AR_SHOUP) {
    TRY(launch_limb(c, LIMB_INV, AR_SHOUP, j, 3 * W, B, "behz_intt"));
  } else {
    TRY(launch_limb(c, LIMB_INV, c->ar_q, jq, 3 * L, B, "behz_intt_q"));
    TRY(launch_limb(c, LIMB_INV, AR_SHOUP, jb, 3 * c->nbsk, B, "behz_intt_bsk"));
  }
  {
    Launch l(c, "behz_scale");
    DISPATCH_L(c, (k_behz_scale<LL><<<dim3(ncols / cvt, 3, B), cvt, 0, c->stream>>>(X, out3, c->dC, N, c->L, col0)));
    CK(cudaGetLastError());
  }
  // c2 is needed whole by every rank (ModUp of the relinearisation); c0, c1 only on their owners, but as whole rows
  if (cols) TRY(allgather_columns(c, out3, (size_t)3 * L, (size_t)3 * L * N));
  return ABC_OK;
}

// one Galois automorphism + key switch (Evaluator::apply_galois_inplace): dst = (sigma(c0), 0) + KeySwitch(sigma(c1)).
// The coefficient permutation is never materialised: it is a gather in the ModUp load and in the ModDown base read.
abc_status apply_galois(abc_ctx *c, const u64 *src, u64 *dst, u32 elt, const u64 *addend = nullptr, u64 *dst_plain = nullptr) {
  auto it = c->galois.find(elt);
  if (it == c->galois.end()) return fail(c, ABC_ERR_STATE, "Galois key not present");
  const long long LN = (long long)c->L * c->N;
  const u32 elt_inv = (u32)hm::invmod(elt, 2ull * c->N);
  // limb-sharded: ModUp needs every limb of c1 on every rank (all-gathered inside keyswitch, overlapped with the local rows)
  return keyswitch(c, src + LN, 2 * LN, it->second, src, 2 * LN, nullptr, 0, elt_inv, dst, addend, dst_plain,
                   c->world > 1 ? const_cast<u64 *>(src) : nullptr);
}

u32 elt_from_step(const abc_ctx *c, int step) {
  const u32 n = (u32)c->N, m = 2 * n;
  const u32 pos = (u32)(step < 0 ? -step : step);
  const u32 e = step < 0 ? (n >> 1) - pos : pos;
  u64 elt = 1;
  for (u32 i = 0; i < e; ++i) elt = (elt * 3) & (m - 1);
  return (u32)elt;
}

// Evaluator::rotate_internal flattened into the list of Galois elements it applies, in order:
// a direct key when one exists, else one step per non-zero NAF digit (least significant first, |digit| = N/2 skipped).
abc_status rotation_plan(abc_ctx *c, int steps, std::vector<u32> &plan) {
  if (steps == 0) return ABC_OK;
  const u32 elt = elt_from_step(c, steps);
  if (c->galois.count(elt)) { plan.push_back(elt); return ABC_OK; }
  std::vector<int> naf;  // util::naf
  {
    int v = steps < 0 ? -steps : steps; const bool neg = steps < 0;
    for (int i = 0; v; ++i) {
      int zi = (v & 1) ? 2 - (v & 3) : 0;
      v = (v - zi) >> 1;
      if (zi) naf.push_back((neg ? -zi : zi) * (1 << i));
    }
  }
  if (naf.size() == 1) return fail(c, ABC_ERR_STATE, "Galois key not present");
  for (int s : naf) {
    if ((s < 0 ? -s : s) == (c->N >> 1)) continue;
    TRY(rotation_plan(c, s, plan));
  }
  return ABC_OK;
}

abc_status encode_device(abc_ctx *c, const int64_t *slots, size_t n, int broadcast, u64 **plain_out) {
  const int N = c->N, Bp = broadcast ? 1 : c->B;
  if (!slots || n == 0) return fail(c, ABC_ERR_PARAM, "Cannot encode an empty vector.");
  if (n > (size_t)N)
    return fail(c, ABC_ERR_PARAM, "Cannot encode " + std::to_string(n) + " elements in a ciphertext of size " +
                                      std::to_string(N) + ". ");
  long long *d_slots = nullptr;
  u64 *plain = nullptr;
  const size_t words = (size_t)Bp * n;
  const unsigned hb = c->h2d_next++ & 1u;
  if (c->h2d_words[hb] < words) {  // grow-only staging buffer (rare: first use, or a larger operand)
    CK(cudaStreamSynchronize(c->stream)); CK(cudaStreamSynchronize(c->copy_stream));
    if (c->h2d_buf[hb]) CK(cudaFree(c->h2d_buf[hb]));
    c->h2d_buf[hb] = nullptr; c->h2d_words[hb] = 0;
    CK(cudaMalloc((void **)&c->h2d_buf[hb], words * sizeof(long long)));
    c->h2d_words[hb] = words;
  }
  d_slots = c->h2d_buf[hb];
  CK(cudaStreamWaitEvent(c->copy_stream, c->h2d_consumed[hb], 0));  // the encode kernel that last read this buffer is done
  CK(cudaMemcpyAsync(d_slots, slots, words * sizeof(long long), cudaMemcpyHostToDevice, c->copy_stream));
  CK(cudaEventRecord(c->h2d_copied[hb], c->copy_stream));
  CK(cudaStreamWaitEvent(c->stream, c->h2d_copied[hb], 0));
  TRY(salloc(c, &plain, (size_t)Bp * N));
  LimbJob j = blank_job();
  j.dst = plain; j.dst_is = N; j.rowmod = c->rm_t;
  j.slots_in = d_slots; j.slots_is = (long long)n; j.n_slots = (int)n; j.index_map = c->d_index_map;
  if (c->logN >= 15) {  // two-pass sizes: scatter kernel, then INTT mod t in place
    {
      Launch l(c, "encode_scatter");
      k_encode_scatter<<<dim3(N / 256, 1, Bp), 256, 0, c->stream>>>(d_slots, (long long)n, (int)n, c->d_index_map, plain, c->t, N);
      CK(cudaGetLastError());
    }
    j.src = plain; j.src_is = N;
    TRY(launch_limb(c, LIMB_INV, c->ar_t, j, 1, Bp, "encode_intt"));
  } else {
    TRY(launch_limb(c, LIMB_ENCODE_INV, c->ar_t, j, 1, Bp, "encode_intt"));
  }
  CK(cudaEventRecord(c->h2d_consumed[hb], c->stream));
  *plain_out = plain;
  return ABC_OK;
}

abc_status encrypt_device(abc_ctx *c, const u64 *plain, int broadcast, u64 *ct) {
  if (!c->have_keys) return fail(c, ABC_ERR_STATE, "keys not generated");
  const int N = c->N, L = c->L, k = c->k, B = c->B;
  const u64 nonce0 = c->enc_salt + c->enc_nonce * (u64)B;
  c->enc_nonce++;
  u64 *u = nullptr, *tmp = nullptr, *rnd = nullptr;
  TRY(scratch(c, SC_U, &u, (size_t)B * k * N));
  TRY(scratch(c, SC_TMP, &tmp, (size_t)B * 2 * k * N));
  TRY(scratch(c, SC_RND, &rnd, (size_t)B * 3 * N));
  {  // the three streams of every instance (u, e_0, e_1), one ChaCha20 block per thread; the k limb rows of u then read them
    Launch l(c, "enc_rng_fill");
    k_rng_fill<<<dim3((N / 8 + 127) / 128, 3, B), 128, 0, c->stream>>>(rnd, c->rng_key, DOM_ENC, nonce0, 0, N);
    CK(cudaGetLastError());
  }
  LimbJob j = blank_job();
  j.dst = u; j.dst_is = (long long)k * N; j.rowmod = c->rm_key;
  j.rng = c->rng_key; j.domain = DOM_ENC; j.a0 = nonce0; j.b = 0; j.rnd = rnd; j.rnd_is = (long long)3 * N;
  TRY(launch_limb(c, LIMB_TERNARY_FWD, c->ar_q, j, k, B, "enc_sample_u_ntt"));
  j = blank_job();
  j.src = u; j.src_is = (long long)k * N; j.rowsrc = c->rm_key;
  j.mul = c->d_pk; j.mul_is = 0;
  j.dst = tmp; j.dst_is = (long long)2 * k * N; j.rowmod = c->rm_key;
  TRY(launch_limb(c, LIMB_MUL_INV, c->ar_q, j, 2 * k, B, "enc_mul_pk_intt"));
  {
    Launch l(c, "enc_finish");
    k_enc_finish<<<dim3(N / 256, 2, B), 256, 0, c->stream>>>(tmp, plain, broadcast ? 0 : N, ct, rnd, c->dC, N, L, k);
    CK(cudaGetLastError());
  }
  return ABC_OK;
}

abc_status mul_plain_device(abc_ctx *c, u64 *dst, const u64 *a, const u64 *plain, int broadcast) {
  const int N = c->N, L = c->L, B = c->B, Bp = broadcast ? 1 : B;
  u64 *P = nullptr;
  TRY(scratch(c, SC_P, &P, (size_t)Bp * L * N));
  LimbJob j = blank_job();
  j.t = c->t; j.t_half_up = (c->t + 1) >> 1;
  j.src = plain; j.src_is = N; j.rowsrc = c->rs_zero; j.dst = P; j.dst_is = (long long)L * N; j.rowmod = c->rm_ct;
  TRY(launch_limb(c, LIMB_PLAINLIFT_FWD, c->ar_q, j, L, Bp, "plain_lift_ntt"));
  j = blank_job();
  const int nown = c->own_hi - c->own_lo;   // only the limbs this rank owns
  j.src = a; j.dst = dst; j.src_is = j.dst_is = (long long)2 * L * N; j.rowmod = c->rm_own; j.rowdst = c->rd_own;
  j.mul = P; j.mul_is = broadcast ? 0 : (long long)L * N; j.rowmul = c->rm_own;
  if (nown > 0) TRY(launch_limb(c, LIMB_FWD_MUL_INV, c->ar_q, j, 2 * nown, B, "ct_mul_plain"));
  return ABC_OK;
}

template <int SUB> abc_status plain_addsub_device(abc_ctx *c, u64 *dst, const u64 *a, const u64 *plain, int broadcast) {
  Launch l(c, SUB ? "plain_sub" : "plain_add");
  k_plain_addsub<SUB><<<dim3(c->N / 256, 1, c->B), 256, 0, c->stream>>>(dst, a, plain, broadcast ? 0 : c->N, c->dC, c->N,
                                                                      c->L, c->own_lo, c->own_hi);
  CK(cudaGetLastError());
  return ABC_OK;
}

// one key-switching-key block [2][k][N]: c1 = a uniform, c0 = -(a s + e) (+ p * newkey at limb J)
abc_status gen_key_block(abc_ctx *c, u64 *blk, u64 dom, u64 a_id, u64 b_base, const u64 *newkey, int J, u64 *e_ntt) {
  const int N = c->N, k = c->k;
  {
    Launch l(c, "keygen_uniform");
    k_sample_uniform<<<dim3(N / 256, k), 256, 0, c->stream>>>(blk + (size_t)k * N, c->rng_key, dom, a_id, b_base, c->dC, N);
    CK(cudaGetLastError());
  }
  LimbJob j = blank_job();
  j.dst = e_ntt; j.dst_is = 0; j.rowmod = c->rm_key;
  j.rng = c->rng_key; j.domain = dom; j.a0 = a_id; j.b = (b_base << 2) | 1;
  TRY(launch_limb(c, LIMB_CBD_FWD, c->ar_q, j, k, 1, "keygen_noise_ntt"));
  {
    Launch l(c, "keygen_finish");
    k_ksk_finish<<<dim3(N / 256, k), 256, 0, c->stream>>>(blk, e_ntt, c->d_sk, newkey, J, c->d_mods, c->dC, N, k);
    CK(cudaGetLastError());
  }
  return ABC_OK;
}
abc_status gen_kswitch_key(abc_ctx *c, u64 *key, u64 key_id, u32 galois_elt, u64 *nk, u64 *e_ntt) {
  const int N = c->N, k = c->k, L = c->L;
  {
    Launch l(c, "keygen_newkey");
    k_newkey<<<dim3(N / 256, k), 256, 0, c->stream>>>(nk, c->d_sk, galois_elt, c->d_mods, N, c->logN);
    CK(cudaGetLastError());
  }
  for (int J = 0; J < L; ++J)
    TRY(gen_key_block(c, key + (size_t)J * 2 * k * N, DOM_KSK, key_id, (u64)J << 8, nk, J, e_ntt));
  return ABC_OK;
}

std::vector<u32> galois_elts_all(const abc_ctx *c) {
  const u64 m = 2ull * c->N;
  std::vector<u32> v;
  v.push_back((u32)(m - 1));
  u64 pos = 3, neg = hm::invmod(3, m);
  for (int i = 0; i < c->logN - 1; ++i) {
    v.push_back((u32)pos); pos = (pos * pos) & (m - 1);
    v.push_back((u32)neg); neg = (neg * neg) & (m - 1);
  }
  return v;
}

bool valid_ct(const abc_ctx *c, const abc_ct *x) { return x && x->ctx == c && x->b; }

}  // namespace

// =================================================================================================
extern "C" {

abc_status abc_ctx_create(const abc_params *p, abc_ctx **out) {
  if (out) *out = nullptr;
  if (!p || !out) { g_create_error = "null argument"; return ABC_ERR_PARAM; }
  abc_ctx *c = new abc_ctx();
  auto bail = [&](abc_status s, const std::string &m) {
    g_create_error = m.empty() ? c->err : m;
    abc_ctx_destroy(c);
    return s;
  };
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return bail(ABC_ERR_CUDA, "no CUDA device: libabc_b200 has no CPU fallback");
  if (p->device < 0 || p->device >= ndev) return bail(ABC_ERR_PARAM, "invalid device ordinal");
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, p->device);
  if (prop.major < 10) return bail(ABC_ERR_CUDA, "device is not sm_100-class: kernels are built for sm_100a only");
  c->device = p->device;
  if (cudaSetDevice(c->device) != cudaSuccess) return bail(ABC_ERR_CUDA, "cudaSetDevice failed");
  const u64 N = p->poly_degree;
  int logN = 0;
  while ((1ull << logN) < N) ++logN;
  if ((1ull << logN) != N || logN < 12 || logN > 16) return bail(ABC_ERR_PARAM, "poly_degree must be a power of two in [4096, 65536]");
  c->N = (int)N; c->logN = logN; c->B = p->batch ? (int)p->batch : 1; c->seed = p->seed;
  {  // seed 0 = not reproducible: the sampler key and the encryption salt come from the OS generator
    u64 r[5] = {0, 0, 0, 0, 0};
    if (getrandom(r, sizeof r, 0) != (ssize_t)sizeof r) return bail(ABC_ERR_STATE, "getrandom failed: no entropy for keys / encryption randomness");
    c->enc_salt = r[4];
    if (c->seed == 0) memcpy(c->rng_key.k, r, 32);
    else c->rng_key = rng_key_from_seed(c->seed);
  }
  try {
    if (p->n_primes == 0) c->primes = hm::bfv_default_primes(N);
    else c->primes.assign(p->primes, p->primes + p->n_primes);
    c->t = p->plain_modulus ? p->plain_modulus : hm::get_primes(N, 20, 1)[0];
    c->k = (int)c->primes.size(); c->L = c->k - 1;
    if (c->k < 2 || c->L > ABC_MAXL - 1) return bail(ABC_ERR_PARAM, "need 2..32 coefficient primes (incl. the special prime)");
    for (u64 q : c->primes)
      if (!hm::is_prime(q) || (q - 1) % (2 * N) || q >> 60) return bail(ABC_ERR_PARAM, "coefficient primes must be < 2^60 and = 1 mod 2N");
    for (size_t i = 0; i < c->primes.size(); ++i)
      for (size_t j = 0; j < i; ++j)
        if (c->primes[i] == c->primes[j]) return bail(ABC_ERR_PARAM, "coefficient primes must be distinct");
    if (!hm::is_prime(c->t) || (c->t - 1) % (2 * N)) return bail(ABC_ERR_PARAM, "plain_modulus must be a prime = 1 mod 2N (batching)");
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(ABC_ERR_CUDA, "stream creation failed");
    if (cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(ABC_ERR_CUDA, "stream creation failed");
    if (cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(ABC_ERR_CUDA, "stream creation failed");
    for (int i = 0; i < 2; ++i) {
      cudaEventCreateWithFlags(&c->h2d_copied[i], cudaEventDisableTiming);
      cudaEventCreateWithFlags(&c->h2d_consumed[i], cudaEventDisableTiming);
      cudaEventCreateWithFlags(&c->d2h_ready[i], cudaEventDisableTiming);
      cudaEventCreateWithFlags(&c->d2h_done[i], cudaEventDisableTiming);
    }
    cudaEventCreate(&c->ev0); cudaEventCreate(&c->ev1);
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, c->device) == cudaSuccess) {
      unsigned long long thr = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    abc_status s = build_tables(c);
    if (s != ABC_OK) return bail(s, "");
  } catch (const std::exception &e) {
    return bail(ABC_ERR_PARAM, e.what());
  }
  *out = c;
  return ABC_OK;
}

void abc_ctx_destroy(abc_ctx *c) {
  if (!c) return;
  if (c->live_handles > 0) { c->zombie = true; return; }   // completed by the last handle's free
  cudaSetDevice(c->device);
  rot_cache_clear(c);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->comm_stream) { cudaStreamSynchronize(c->comm_stream); cudaStreamDestroy(c->comm_stream); }
  if (c->comm_ready) cudaEventDestroy(c->comm_ready);
  if (c->comm_done) cudaEventDestroy(c->comm_done);
  if (c->comm && nccl_api()) nccl_api()->CommDestroy(c->comm);
  for (auto &kv : c->galois) cudaFree(kv.second);
  for (auto &kv : c->key_f64) cudaFree(kv.second);
  if (c->ks_fault_h) cudaFreeHost(c->ks_fault_h);
  if (c->ksr_acc) cudaFree(c->ksr_acc);
  cudaFree(c->d_sk); cudaFree(c->d_pk); cudaFree(c->d_relin);
  for (void *p : c->owned) cudaFree(p);
  if (c->flush_buf) cudaFree(c->flush_buf);
  for (u64 *p : c->sc_ptr) if (p) cudaFree(p);
  for (auto &b : c->blk_free) cudaFree(b.second);
  for (auto &r : c->prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
  if (c->d2h_stream) { cudaStreamSynchronize(c->d2h_stream); cudaStreamDestroy(c->d2h_stream); }
  for (int i = 0; i < 2; ++i) {
    if (c->d2h_buf[i]) cudaFree(c->d2h_buf[i]);
    if (c->d2h_ready[i]) cudaEventDestroy(c->d2h_ready[i]);
    if (c->d2h_done[i]) cudaEventDestroy(c->d2h_done[i]);
  }
  for (int i = 0; i < 2; ++i) {
    if (c->h2d_buf[i]) cudaFree(c->h2d_buf[i]);
    if (c->h2d_copied[i]) cudaEventDestroy(c->h2d_copied[i]);
    if (c->h2d_consumed[i]) cudaEventDestroy(c->h2d_consumed[i]);
  }
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

const char *abc_last_error(const abc_ctx *c) { return c ? c->err.c_str() : g_create_error.c_str(); }
abc_status abc_sync(abc_ctx *c) {
  CK(cudaStreamSynchronize(c->stream));
  if (c->d2h_stream) CK(cudaStreamSynchronize(c->d2h_stream));
  return check_ks_fault(c);
}
abc_status abc_host_alloc(abc_ctx *c, size_t bytes, void **out) {
  CK(cudaSetDevice(c->device));
  CK(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
  return ABC_OK;
}
void abc_host_free(void *p) { if (p) cudaFreeHost(p); }
abc_status abc_host_register(abc_ctx *c, void *p, size_t bytes) {
  CK(cudaSetDevice(c->device));
  CK(cudaHostRegister(p, bytes, cudaHostRegisterDefault));
  return ABC_OK;
}
void abc_host_unregister(void *p) { if (p) cudaHostUnregister(p); }
int abc_faulted(const abc_ctx *c) { return c->faulted ? 1 : 0; }
abc_status abc_clear_fault(abc_ctx *c) {
  CK(cudaStreamSynchronize(c->stream));
  if (c->ks_fault_h) *c->ks_fault_h = 0;
  c->faulted = false;
  return ABC_OK;
}
uint32_t abc_poly_degree(const abc_ctx *c) { return (uint32_t)c->N; }
uint32_t abc_n_primes(const abc_ctx *c) { return (uint32_t)c->k; }
uint32_t abc_n_limbs(const abc_ctx *c) { return (uint32_t)c->L; }
uint32_t abc_batch(const abc_ctx *c) { return (uint32_t)c->B; }
uint64_t abc_plain_modulus(const abc_ctx *c) { return c->t; }
abc_status abc_get_primes(const abc_ctx *c, uint64_t *out) {
  for (int i = 0; i < c->k; ++i) out[i] = c->primes[i];
  return ABC_OK;
}
abc_status abc_get_aux_primes(const abc_ctx *c, uint64_t *out, uint32_t *count) {
  out[0] = c->msk; out[1] = c->gamma;
  for (int j = 0; j < c->nB; ++j) out[2 + j] = c->bsk[j];
  if (count) *count = (uint32_t)c->nB + 2;
  return ABC_OK;
}

// ---- keys
static abc_status keygen_impl(abc_ctx *c, const std::vector<u32> &elts) {
  const int N = c->N, k = c->k, L = c->L;
  CK(cudaSetDevice(c->device));
  const size_t kw = (size_t)L * 2 * k * N;
  drop_key_f64(c);
  if (!c->d_sk) CK(cudaMalloc((void **)&c->d_sk, (size_t)k * N * 8));
  if (!c->d_pk) CK(cudaMalloc((void **)&c->d_pk, (size_t)2 * k * N * 8));
  if (!c->d_relin) CK(cudaMalloc((void **)&c->d_relin, kw * 8));
  LimbJob j = blank_job();
  j.dst = c->d_sk; j.rowmod = c->rm_key; j.rng = c->rng_key; j.domain = DOM_SK;
  TRY(launch_limb(c, LIMB_TERNARY_FWD, c->ar_q, j, k, 1, "keygen_sk_ntt"));
  u64 *nk = nullptr, *e_ntt = nullptr;
  TRY(scratch(c, SC_NK, &nk, (size_t)k * N));
  TRY(scratch(c, SC_ENTT, &e_ntt, (size_t)k * N));
  TRY(gen_key_block(c, c->d_pk, DOM_PK, 0, 0, nullptr, -1, e_ntt));
  TRY(gen_kswitch_key(c, c->d_relin, 0, 0, nk, e_ntt));
  for (u32 elt : elts) {
    if (!(elt & 1) || elt >= 2u * c->N) return fail(c, ABC_ERR_PARAM, "invalid Galois element");
    if (c->galois.count(elt)) continue;
    u64 *key = nullptr;
    CK(cudaMalloc((void **)&key, kw * 8));
    c->galois[elt] = key;
    TRY(gen_kswitch_key(c, key, elt, elt, nk, e_ntt));
  }
  c->have_keys = true;
  return ABC_OK;
}
abc_status abc_keygen(abc_ctx *c) { NvtxOp nvtx_("abc_keygen"); return keygen_impl(c, galois_elts_all(c)); }
abc_status abc_keygen_select(abc_ctx *c, const uint32_t *galois_elts, size_t n) {
  return keygen_impl(c, std::vector<u32>(galois_elts, galois_elts + n));
}
uint32_t abc_galois_elt_from_step(const abc_ctx *c, int step) {
  const int as = step < 0 ? -step : step;
  if (step == 0) return 2u * c->N - 1;
  if (as >= (c->N >> 1)) return 0;
  return elt_from_step(c, step);
}
size_t abc_key_words(const abc_ctx *c, int kind) {
  const size_t kn = (size_t)c->k * c->N;
  switch (kind) {
    case ABC_KEY_SECRET: return kn;
    case ABC_KEY_PUBLIC: return 2 * kn;
    case ABC_KEY_RELIN: case ABC_KEY_GALOIS: return (size_t)c->L * 2 * kn;
    default: return 0;
  }
}
static u64 **key_slot(abc_ctx *c, int kind, uint32_t elt, bool create) {
  switch (kind) {
    case ABC_KEY_SECRET: return &c->d_sk;
    case ABC_KEY_PUBLIC: return &c->d_pk;
    case ABC_KEY_RELIN: return &c->d_relin;
    case ABC_KEY_GALOIS: {
      if (!(elt & 1) || elt >= 2u * c->N) return nullptr;
      auto it = c->galois.find(elt);
      if (it != c->galois.end()) return &it->second;
      if (!create) return nullptr;
      c->galois[elt] = nullptr;
      return &c->galois[elt];
    }
    default: return nullptr;
  }
}
abc_status abc_key_export(abc_ctx *c, int kind, uint32_t elt, uint64_t *host, size_t words) {
  u64 **slot = key_slot(c, kind, elt, false);
  if (!slot || !*slot) return fail(c, ABC_ERR_STATE, "key not present");
  if (words != abc_key_words(c, kind)) return fail(c, ABC_ERR_PARAM, "key size mismatch");
  CK(cudaMemcpyAsync(host, *slot, words * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return ABC_OK;
}
abc_status abc_key_import(abc_ctx *c, int kind, uint32_t elt, const uint64_t *host, size_t words) {
  CK(cudaSetDevice(c->device));
  u64 **slot = key_slot(c, kind, elt, true);
  if (!slot) return fail(c, ABC_ERR_PARAM, "invalid key kind / Galois element");
  if (words != abc_key_words(c, kind)) return fail(c, ABC_ERR_PARAM, "key size mismatch");
  drop_key_f64(c);
  if (!*slot) CK(cudaMalloc((void **)slot, words * 8));
  CK(cudaMemcpyAsync(*slot, host, words * 8, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (c->d_sk && c->d_pk && c->d_relin) c->have_keys = true;
  return ABC_OK;
}
int abc_has_galois_key(const abc_ctx *c, uint32_t elt) { return c->galois.count(elt) ? 1 : 0; }
abc_status abc_galois_elts(const abc_ctx *c, uint32_t *out, size_t cap, size_t *n) {
  if (n) *n = c->galois.size();
  if (!out) return ABC_OK;
  if (cap < c->galois.size()) return ABC_ERR_PARAM;
  size_t i = 0;
  for (auto &kv : c->galois) out[i++] = kv.first;
  return ABC_OK;
}
void abc_set_error(abc_ctx *c, const char *msg) {  // for the other translation units of the library (sealio.cu)
  if (c) c->err = msg; else g_create_error = msg;
}

// ---- handles
abc_status abc_ct_alloc(abc_ctx *c, abc_ct **out) {
  CK(cudaSetDevice(c->device));
  u64 *d = nullptr;
  TRY(salloc(c, &d, abc_ct_words(c)));
  CtBuf *b = new CtBuf; b->d = d;
  *out = new abc_ct{c, b};
  ++c->live_handles;
  return ABC_OK;
}
// drop one reference to a buffer (stream-ordered free when it was the last one)
static void buf_unref(abc_ctx *c, CtBuf *b) {
  if (!b || --b->refs > 0) return;
  if (b->d) sfree(c, b->d);
  buf_unref(c, b->src); buf_unref(c, b->sum); buf_unref(c, b->other);
  delete b;
}
// point the handle at a buffer it alone owns
static void ct_adopt(abc_ct *ct, u64 *d) {
  buf_unref(ct->ctx, ct->b);
  ct->b = new CtBuf; ct->b->d = d;
}
static void ct_share(abc_ct *dst, CtBuf *b) {
  if (dst->b == b) return;
  ++b->refs;
  buf_unref(dst->ctx, dst->b);
  dst->b = b;
}
// ---- rotation cache (abc_ctx::rot_cache)
static void rot_cache_drop(abc_ctx *c, size_t i) {
  abc_ctx::RotCacheEntry e = c->rot_cache[i];
  c->rot_cache.erase(c->rot_cache.begin() + (long)i);
  buf_unref(c, e.res);
  buf_unref(c, e.src);
}
// entries whose source buffer only the cache still holds can never hit again
static void rot_cache_purge(abc_ctx *c) {
  for (size_t i = c->rot_cache.size(); i-- > 0;)
    if (c->rot_cache[i].src->refs == 1) rot_cache_drop(c, i);
}
static void rot_cache_clear(abc_ctx *c) {
  while (!c->rot_cache.empty()) rot_cache_drop(c, c->rot_cache.size() - 1);
}
static CtBuf *rot_cache_find(abc_ctx *c, const CtBuf *src, u32 elt) {
  for (auto &e : c->rot_cache)
    if (e.src == src && e.elt == elt) { e.stamp = ++c->rot_stamp; ++c->rot_cache_hits; return e.res; }
  return nullptr;
}
static void rot_cache_put(abc_ctx *c, CtBuf *src, u32 elt, CtBuf *res) {
  if (c->rot_cache_max <= 0) return;
  if ((int)c->rot_cache.size() >= c->rot_cache_max) {   // least recently used goes
    size_t lru = 0;
    for (size_t i = 1; i < c->rot_cache.size(); ++i) if (c->rot_cache[i].stamp < c->rot_cache[lru].stamp) lru = i;
    rot_cache_drop(c, lru);
  }
  ++src->refs; ++res->refs;
  c->rot_cache.push_back({src, elt, res, ++c->rot_stamp});
}

// Materialise a deferred rotation: every clone sharing the buffer sees the result.
static abc_status buf_resolve(abc_ctx *c, CtBuf *b) {
  if (b->d) return ABC_OK;
  u64 *d = nullptr;
  TRY(salloc(c, &d, abc_ct_words(c)));
  if (b->sum) {
    Launch l(c, "sub");
    const int nown = c->own_hi - c->own_lo;
    k_addsub<1><<<dim3((c->N / 2 + 255) / 256, 2 * nown, c->B), 256, 0, c->stream>>>(d, b->sum->d, b->other->d, c->dC, c->N, c->L,
                                                                                      2ll * c->L * c->N, c->own_lo, nown);
    if (cudaGetLastError() != cudaSuccess) { sfree(c, d); return fail(c, ABC_ERR_CUDA, "sub launch failed"); }
    b->d = d;
    buf_unref(c, b->sum); buf_unref(c, b->other); b->sum = b->other = nullptr;
    return ABC_OK;
  }
  abc_status st = apply_galois(c, b->src->d, d, b->elt);
  if (st != ABC_OK) { sfree(c, d); return st; }
  b->d = d;
  buf_unref(c, b->src); b->src = nullptr;
  return ABC_OK;
}
static abc_status ct_resolve(abc_ctx *c, const abc_ct *ct) { return buf_resolve(c, ct->b); }
// Copy-on-write: call after the operands' pointers are captured and before dst is written.  Every op overwrites the
// whole of dst, so a shared (or deferred) buffer is simply swapped for a fresh one; the old one lives on in the clones.
static abc_status ct_make_private(abc_ctx *c, abc_ct *dst) {
  if (dst->b->refs == 1 && dst->b->d) return ABC_OK;
  u64 *d = nullptr;
  TRY(salloc(c, &d, abc_ct_words(c)));
  ct_adopt(dst, d);
  return ABC_OK;
}
static void handle_released(abc_ctx *c) {
  if (--c->live_handles == 0 && c->zombie) abc_ctx_destroy(c);
}
void abc_ct_free(abc_ct *ct) {
  if (!ct) return;
  abc_ctx *c = ct->ctx;
  buf_unref(c, ct->b);
  delete ct;
  rot_cache_purge(c);
  handle_released(c);
}
size_t abc_ct_words(const abc_ctx *c) { return (size_t)c->B * ct_words1(c); }
abc_status abc_ct_clone(abc_ctx *c, const abc_ct *src, abc_ct **out) {
  if (!valid_ct(c, src)) return fail(c, ABC_ERR_PARAM, "invalid ciphertext handle");
  ++src->b->refs;
  *out = new abc_ct{c, src->b};
  ++c->live_handles;
  return ABC_OK;
}
int abc_ct_shared(const abc_ct *ct) { return ct && ct->b ? ct->b->refs : 0; }
int abc_ct_deferred(const abc_ct *ct) { return ct && ct->b && !ct->b->d ? 1 : 0; }
abc_status abc_ct_export(abc_ctx *c, const abc_ct *ct, uint64_t *host, size_t words) {
  if (!valid_ct(c, ct) || words != abc_ct_words(c)) return fail(c, ABC_ERR_PARAM, "invalid ciphertext handle or size");
  CHECK_POISON(c);
  TRY(ct_resolve(c, ct));
  CK(cudaMemcpyAsync(host, ct->b->d, words * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return check_ks_fault(c);
}
// one instance of the batch (the unit of a SEAL stream); the other instances of the handle keep their content
abc_status abc_ct_export_instance(abc_ctx *c, const abc_ct *ct, uint32_t inst, uint64_t *host, size_t words) {
  if (!valid_ct(c, ct) || words != ct_words1(c) || inst >= (uint32_t)c->B)
    return fail(c, ABC_ERR_PARAM, "invalid ciphertext handle, instance or size");
  TRY(ct_resolve(c, ct));
  CK(cudaMemcpyAsync(host, ct->b->d + (size_t)inst * words, words * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return check_ks_fault(c);
}
abc_status abc_ct_import_instance(abc_ctx *c, abc_ct *ct, uint32_t inst, const uint64_t *host, size_t words) {
  if (!valid_ct(c, ct) || words != ct_words1(c) || inst >= (uint32_t)c->B)
    return fail(c, ABC_ERR_PARAM, "invalid ciphertext handle, instance or size");
  TRY(ct_resolve(c, ct));
  if (ct->b->refs > 1) {  // copy-on-write with content: only one instance is overwritten
    u64 *d = nullptr;
    TRY(salloc(c, &d, abc_ct_words(c)));
    CK(cudaMemcpyAsync(d, ct->b->d, abc_ct_words(c) * 8, cudaMemcpyDeviceToDevice, c->stream));
    ct_adopt(ct, d);
  }
  CK(cudaMemcpyAsync(ct->b->d + (size_t)inst * words, host, words * 8, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return ABC_OK;
}
abc_status abc_ct_import(abc_ctx *c, abc_ct *ct, const uint64_t *host, size_t words) {
  if (!valid_ct(c, ct) || words != abc_ct_words(c)) return fail(c, ABC_ERR_PARAM, "invalid ciphertext handle or size");
  TRY(ct_make_private(c, ct));
  CK(cudaMemcpyAsync(ct->b->d, host, words * 8, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return ABC_OK;
}

// ---- encode / encrypt / decrypt
abc_status abc_pt_encode(abc_ctx *c, const int64_t *slots, size_t n, int broadcast, abc_pt **out) {
  CK(cudaSetDevice(c->device));
  u64 *plain = nullptr;
  TRY(encode_device(c, slots, n, broadcast, &plain));
  *out = new abc_pt{c, plain, broadcast};
  ++c->live_handles;
  return ABC_OK;
}
void abc_pt_free(abc_pt *pt) {
  if (!pt) return;
  abc_ctx *c = pt->ctx;
  sfree(c, pt->d);
  delete pt;
  handle_released(c);
}
abc_status abc_encrypt_pt(abc_ctx *c, const abc_pt *pt, abc_ct **out) {
  if (!pt || pt->ctx != c) return fail(c, ABC_ERR_PARAM, "invalid plaintext handle");
  TRY(abc_ct_alloc(c, out));
  abc_status s = encrypt_device(c, pt->d, pt->broadcast, (*out)->b->d);
  if (s != ABC_OK) { abc_ct_free(*out); *out = nullptr; }
  return s;
}
abc_status abc_encode_encrypt(abc_ctx *c, const int64_t *slots, size_t n, int broadcast, abc_ct **out) {
  NvtxOp nvtx_("abc_encode_encrypt");
  abc_pt *pt = nullptr;
  TRY(abc_pt_encode(c, slots, n, broadcast, &pt));
  abc_status s = abc_encrypt_pt(c, pt, out);
  abc_pt_free(pt);
  return s;
}
// The sampler key from the caller (32 bytes of its own entropy): what gives every GPU of a job the same keys without
// falling back to a 64-bit seed.  Takes effect for the next abc_keygen and for the encryptions that follow.
abc_status abc_set_rng_key(abc_ctx *c, const uint8_t *key32) {
  if (!key32) return fail(c, ABC_ERR_PARAM, "null key");
  memcpy(c->rng_key.k, key32, 32);
  return ABC_OK;
}
abc_status abc_set_encrypt_nonce(abc_ctx *c, uint64_t nonce) { c->enc_nonce = nonce; c->enc_salt = 0; return ABC_OK; }

// c0 + c1 * s mod q in coefficient form (Decryptor::dot_product_ct_sk_array), x [B][L][N] in the SC_DECX scratch slot
static abc_status dot_ct_sk(abc_ctx *c, const abc_ct *ct, u64 **x_out) {
  if (!valid_ct(c, ct)) return fail(c, ABC_ERR_PARAM, "invalid ciphertext handle");
  CHECK_POISON(c);
  if (!c->d_sk) return fail(c, ABC_ERR_STATE, "secret key not present");
  CK(cudaSetDevice(c->device));
  TRY(ct_resolve(c, ct));
  const int N = c->N, L = c->L, B = c->B;
  u64 *x = nullptr;
  TRY(scratch(c, SC_DECX, &x, (size_t)B * L * N));
  TRY(allgather_limbs(c, ct->b->d, 3));  // limb-sharded: scale-and-round needs every limb
  LimbJob j = blank_job();
  j.src = ct->b->d; j.src_is = (long long)2 * L * N; j.rowsrc = c->rs_c1;
  j.mul = c->d_sk; j.mul_is = 0;
  j.add = ct->b->d; j.add_is = (long long)2 * L * N;
  j.dst = x; j.dst_is = (long long)L * N; j.rowmod = c->rm_ct;
  TRY(launch_limb(c, LIMB_FWD_MUL_INV_ADD, c->ar_q, j, L, B, "dec_c1s_plus_c0"));
  *x_out = x;
  return ABC_OK;
}

// Decryptor::invariant_noise_budget (SealCiphertext::noiseBits, SealCiphertext.cpp:80-83), per instance:
// bit_count(Q) - bit_count(|| t * (c0 + c1*s) mod Q ||_inf, centred) - 1, floored at 0.
abc_status abc_noise_budget(abc_ctx *c, const abc_ct *ct, int32_t *out_bits) {
  NvtxOp nvtx_("abc_noise_budget");
  u64 *x = nullptr;
  TRY(dot_ct_sk(c, ct, &x));
  const int N = c->N, L = c->L, B = c->B;
  if (!c->d_noise_tab) {  // multi-precision constants: (Q/q_i) for every i, Q, (Q+1)/2; L words each, little endian
    std::vector<u64> tab((size_t)(L + 2) * L, 0);
    auto mul_small = [&](u64 *w, u64 m) { u64 carry = 0; for (int i = 0; i < L; ++i) { unsigned __int128 t = (unsigned __int128)w[i] * m + carry; w[i] = (u64)t; carry = (u64)(t >> 64); } };
    for (int i = 0; i <= L; ++i) {  // row L: the full product Q
      u64 *w = &tab[(size_t)i * L];
      w[0] = 1;
      for (int j = 0; j < L; ++j) if (j != i) mul_small(w, c->primes[j]);
    }
    u64 *Q = &tab[(size_t)L * L], *H = Q + L;
    u64 carry = 1;                                   // (Q + 1) / 2: Q is odd
    for (int i = 0; i < L; ++i) { u64 v = Q[i] + carry; carry = v < carry ? 1 : 0; H[i] = v; }
    for (int i = 0; i < L; ++i) H[i] = (H[i] >> 1) | (i + 1 < L ? H[i + 1] << 63 : carry << 63);
    TRY(upload(c, &c->d_noise_tab, tab));
    int qb = 0;
    for (int i = L - 1; i >= 0 && !qb; --i) if (Q[i]) qb = 64 * i + hm::bits_of(Q[i]);
    c->q_bits = qb;
  }
  int *d_bits = nullptr;
  CK(cudaMallocAsync((void **)&d_bits, (size_t)B * sizeof(int), c->stream));
  CK(cudaMemsetAsync(d_bits, 0, (size_t)B * sizeof(int), c->stream));
  {
    Launch l(c, "noise_norm");
    k_noise_bits<<<dim3(N / 128, B), 128, 0, c->stream>>>(x, c->dC, c->d_noise_tab, N, L, d_bits);
    CK(cudaGetLastError());
  }
  std::vector<int> bits(B);
  CK(cudaMemcpyAsync(bits.data(), d_bits, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  sfree(c, d_bits);
  for (int i = 0; i < B; ++i) { const int d = c->q_bits - bits[i] - 1; out_bits[i] = d > 0 ? d : 0; }
  return check_ks_fault(c);
}

// Ciphertext::is_transparent per instance (SEAL's Evaluator throws std::logic_error "result ciphertext is transparent" on
// such results when built with SEAL_THROW_ON_TRANSPARENT_CIPHERTEXT, its default).  Synchronises: the C++
// drop-in calls it after every op only when the factory was asked to mirror that behaviour.
abc_status abc_is_transparent(abc_ctx *c, const abc_ct *ct, int32_t *out_flags) {
  NvtxOp nvtx_("abc_is_transparent");
  if (!valid_ct(c, ct)) return fail(c, ABC_ERR_PARAM, "invalid ciphertext handle");
  CHECK_POISON(c);
  CK(cudaSetDevice(c->device));
  TRY(ct_resolve(c, ct));
  const int N = c->N, L = c->L, B = c->B;
  const int l0 = c->own_lo, l1 = c->own_hi;   // limb-sharded: this rank's verdict covers the limbs it owns
  int *d_nz = nullptr;
  CK(cudaMallocAsync((void **)&d_nz, (size_t)B * sizeof(int), c->stream));
  CK(cudaMemsetAsync(d_nz, 0, (size_t)B * sizeof(int), c->stream));
  {
    Launch l(c, "c1_nonzero");
    k_c1_nonzero<<<dim3(std::max(1, std::min(64, (l1 - l0) * N / 2048)), B), 256, 0, c->stream>>>(ct->b->d, N, L, l0, l1, d_nz);
    CK(cudaGetLastError());
  }
  std::vector<int> nz(B);
  CK(cudaMemcpyAsync(nz.data(), d_nz, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  sfree(c, d_nz);
  for (int i = 0; i < B; ++i) out_flags[i] = nz[i] ? 0 : 1;
  return check_ks_fault(c);
}

// Decryptor::decrypt + BatchEncoder::decode, enqueued: the slots land in out_slots (pinned host memory for a truly
// asynchronous copy) once abc_decrypt_wait / abc_sync returns.  The device-side result sits in one of two buffers and
// leaves on the context's D2H stream, so the copy overlaps the kernels of the ops enqueued after this call.
abc_status abc_decrypt_decode_async(abc_ctx *c, const abc_ct *ct, int64_t *out_slots) {
  NvtxOp nvtx_("abc_decrypt_decode_async");
  u64 *x = nullptr, *plain = nullptr;
  TRY(dot_ct_sk(c, ct, &x));
  const int N = c->N, B = c->B;
  TRY(scratch(c, SC_DECP, &plain, (size_t)B * N));
  const unsigned hb = c->d2h_next++ & 1u;
  const size_t words = (size_t)B * N;
  if (c->d2h_words[hb] < words) {   // grow-only (first use)
    CK(cudaStreamSynchronize(c->stream)); CK(cudaStreamSynchronize(c->d2h_stream));
    if (c->d2h_buf[hb]) CK(cudaFree(c->d2h_buf[hb]));
    c->d2h_buf[hb] = nullptr; c->d2h_words[hb] = 0;
    CK(cudaMalloc((void **)&c->d2h_buf[hb], words * sizeof(long long)));
    c->d2h_words[hb] = words;
    CK(cudaEventRecord(c->d2h_done[hb], c->d2h_stream));
  }
  long long *d_out = c->d2h_buf[hb];
  CK(cudaStreamWaitEvent(c->stream, c->d2h_done[hb], 0));   // the copy that last read this buffer has finished
  LimbJob j = blank_job();
  {
    Launch l(c, "dec_scale_round");
    DISPATCH_L(c, (k_dec_finish<LL><<<dim3(N / 128, 1, B), 128, 0, c->stream>>>(x, plain, c->dC, N, c->L)));
    CK(cudaGetLastError());
  }
  j = blank_job();
  j.src = plain; j.src_is = N; j.rowmod = c->rm_t; j.slots_out = d_out; j.index_map = c->d_index_map;
  if (c->logN >= 15) {  // two-pass sizes: NTT mod t in place, then gather kernel
    j.dst = plain; j.dst_is = N;
    TRY(launch_limb(c, LIMB_FWD, c->ar_t, j, 1, B, "decode_ntt"));
    Launch l(c, "decode_gather");
    k_decode_gather<<<dim3(N / 256, 1, B), 256, 0, c->stream>>>(plain, c->d_index_map, d_out, c->t, N);
    CK(cudaGetLastError());
  } else {
    TRY(launch_limb(c, LIMB_FWD_DECODE, c->ar_t, j, 1, B, "decode_ntt"));
  }
  CK(cudaEventRecord(c->d2h_ready[hb], c->stream));
  CK(cudaStreamWaitEvent(c->d2h_stream, c->d2h_ready[hb], 0));
  CK(cudaMemcpyAsync(out_slots, d_out, words * sizeof(long long), cudaMemcpyDeviceToHost, c->d2h_stream));
  CK(cudaEventRecord(c->d2h_done[hb], c->d2h_stream));
  return ABC_OK;
}
abc_status abc_decrypt_wait(abc_ctx *c) {
  CK(cudaStreamSynchronize(c->d2h_stream));
  return check_ks_fault(c);
}
abc_status abc_decrypt_decode(abc_ctx *c, const abc_ct *ct, int64_t *out_slots) {
  TRY(abc_decrypt_decode_async(c, ct, out_slots));
  return abc_decrypt_wait(c);
}

// ---- ciphertext ops
static abc_status fused_rotate_add(abc_ctx *c, abc_ct *dst, const abc_ct *rot, const abc_ct *other);
static abc_status addsub(abc_ctx *c, abc_ct *dst, const abc_ct *a, const abc_ct *b, int op) {
  NvtxOp nvtx_(op == 0 ? "abc_add" : op == 1 ? "abc_sub" : "abc_negate");
  if (!valid_ct(c, dst) || !valid_ct(c, a) || (op != 2 && !valid_ct(c, b)))
    return fail(c, ABC_ERR_PARAM, "invalid ciphertext handle");
  CHECK_POISON(c);
  if (op == 0) {  // add with a deferred rotation: accumulate in its key switch
    if (!a->b->d && !b->b->d) TRY(ct_resolve(c, a));
    if (!b->b->d) return fused_rotate_add(c, dst, b, a);
    if (!a->b->d) return fused_rotate_add(c, dst, a, b);
  }
  TRY(ct_resolve(c, a));
  if (b) TRY(ct_resolve(c, b));
  const int N = c->N, L = c->L;
  Launch l(c, op == 0 ? "add" : op == 1 ? "sub" : "negate");
  const int i0 = c->own_lo, nown = c->own_hi - c->own_lo;
  if (nown == 0) return ABC_OK;
  dim3 grid((N / 2 + 255) / 256, 2 * nown, c->B);
  const long long is = 2ll * L * N;
  const u64 *pa = a->b->d, *pb = b ? b->b->d : nullptr;
  TRY(ct_make_private(c, dst));
  if (op == 0) k_addsub<0><<<grid, 256, 0, c->stream>>>(dst->b->d, pa, pb, c->dC, N, L, is, i0, nown);
  else if (op == 1) k_addsub<1><<<grid, 256, 0, c->stream>>>(dst->b->d, pa, pb, c->dC, N, L, is, i0, nown);
  else k_addsub<2><<<grid, 256, 0, c->stream>>>(dst->b->d, pa, nullptr, c->dC, N, L, is, i0, nown);
  CK(cudaGetLastError());
  return ABC_OK;
}
// dst = rot + other where rot is a deferred rotation (other is resolved).  One key switch writes the sum.  If rot has
// holders besides dst it turns into the deferred difference sum - other (see CtBuf): never a second key switch.
static abc_status fused_rotate_add(abc_ctx *c, abc_ct *dst, const abc_ct *rot, const abc_ct *other) {
  CtBuf *rb = rot->b, *ob = other->b;
  if (rb->sum) { TRY(buf_resolve(c, rb)); return addsub(c, dst, rot, other, 0); }  // already a difference: plain add
  const bool keep = !(dst->b == rb && rb->refs == 1);
  u64 *sum = nullptr;
  TRY(salloc(c, &sum, abc_ct_words(c)));
  abc_status st = apply_galois(c, rb->src->d, sum, rb->elt, ob->d, nullptr);
  if (st != ABC_OK) { sfree(c, sum); return st; }
  CtBuf *sb = new CtBuf; sb->d = sum;
  if (keep) {
    buf_unref(c, rb->src); rb->src = nullptr;
    rb->sum = sb; ++sb->refs;
    rb->other = ob; ++ob->refs;
  }
  buf_unref(c, dst->b);
  dst->b = sb;
  return ABC_OK;
}
abc_status abc_add(abc_ctx *c, abc_ct *dst, const abc_ct *a, const abc_ct *b) { return addsub(c, dst, a, b, 0); }
abc_status abc_sub(abc_ctx *c, abc_ct *dst, const abc_ct *a, const abc_ct *b) { return addsub(c, dst, a, b, 1); }
abc_status abc_negate(abc_ctx *c, abc_ct *dst, const abc_ct *a) { return addsub(c, dst, a, nullptr, 2); }

abc_status abc_mul_relin(abc_ctx *c, abc_ct *dst, const abc_ct *a, const abc_ct *b) {
  NvtxOp nvtx_("abc_mul_relin");
  if (!valid_ct(c, dst) || !valid_ct(c, a) || !valid_ct(c, b)) return fail(c, ABC_ERR_PARAM, "invalid ciphertext handle");
  if (!c->d_relin) return fail(c, ABC_ERR_STATE, "relinearisation key not present");
  CHECK_POISON(c);
  const size_t LN = (size_t)c->L * c->N;
  u64 *out3 = nullptr;
  TRY(scratch(c, SC_OUT3, &out3, (size_t)c->B * 3 * LN));
  // limb-sharded: BEHZ base conversion needs every limb of both operands; the product is computed on every rank
  // (replicated), the relinearisation key switch only for the output moduli this rank owns
  TRY(ct_resolve(c, a)); TRY(ct_resolve(c, b));
  TRY(allgather_limbs(c, a->b->d, 3));
  if (b->b != a->b) TRY(allgather_limbs(c, b->b->d, 3));
  TRY(behz_multiply(c, a->b->d, b->b->d, out3));
  TRY(ct_make_private(c, dst));  // the operands are consumed into out3 by now
  TRY(keyswitch(c, out3 + 2 * LN, 3ll * LN, c->d_relin, out3, 3ll * LN, out3 + LN, 3ll * LN, 0, dst->b->d));
  return ABC_OK;
}

// dst = rotate_rows(a, steps) [+ addend].  The addend is accumulated in the ModDown of the last key switch, which
// gives exactly add(rotate_rows(a), addend): modular addition of canonical residues is associative.
static abc_status rotate_impl(abc_ctx *c, abc_ct *dst, const abc_ct *a, int steps, const abc_ct *addend) {
  NvtxOp nvtx_(addend ? "abc_rotate_rows_add" : "abc_rotate_rows");
  if (!valid_ct(c, dst) || !valid_ct(c, a) || (addend && !valid_ct(c, addend)))
    return fail(c, ABC_ERR_PARAM, "invalid ciphertext handle");
  CHECK_POISON(c);
  const int as = steps < 0 ? -steps : steps;
  if (as >= (c->N >> 1)) return fail(c, ABC_ERR_PARAM, "step count too large");
  std::vector<u32> plan;
  TRY(rotation_plan(c, steps, plan));
  if (plan.empty()) {  // rotation by 0: dst becomes another clone of a
    if (addend) return addsub(c, dst, a, addend, 0);
    ct_share(dst, a->b);
    return ABC_OK;
  }
  TRY(ct_resolve(c, a));
  if (addend) TRY(ct_resolve(c, addend));
  // The last Galois step is deferred (see CtBuf) unless an addend is given or the context is limb-sharded; the steps
  // before it run now.  Each key switch writes a fresh buffer (its gathers read the previous one).
  const bool defer = !addend && c->lazy_rotate && c->world == 1;
  const bool cached = c->rot_cache_max > 0 && c->world == 1;
  CtBuf *cur = a->b;
  ++cur->refs;
  for (size_t pi = 0; pi < plan.size(); ++pi) {
    const bool last = pi + 1 == plan.size();
    if (cached && !(last && addend))
      if (CtBuf *hit = rot_cache_find(c, cur, plan[pi])) {   // this step of this buffer exists already
        ++hit->refs;
        buf_unref(c, cur);
        cur = hit;
        continue;
      }
    if (last && defer) {
      CtBuf *nb = new CtBuf; nb->src = cur; nb->elt = plan[pi];  // takes over the reference held on cur
      buf_unref(c, dst->b);
      dst->b = nb;
      return ABC_OK;
    }
    u64 *out = nullptr;
    abc_status s = salloc(c, &out, abc_ct_words(c));
    if (s == ABC_OK) {
      s = apply_galois(c, cur->d, out, plan[pi], (addend && last) ? addend->b->d : nullptr);
      if (s != ABC_OK) sfree(c, out);
    }
    if (s != ABC_OK) { buf_unref(c, cur); return s; }
    CtBuf *nb = new CtBuf; nb->d = out;
    if (cached && !last) rot_cache_put(c, cur, plan[pi], nb);   // the shared prefix of the NAF chains of one source
    buf_unref(c, cur);
    cur = nb;
  }
  buf_unref(c, dst->b);  // released stream-ordered: after the kernels that read it
  dst->b = cur;
  return ABC_OK;
}
abc_status abc_rotate_rows(abc_ctx *c, abc_ct *dst, const abc_ct *a, int steps) { return rotate_impl(c, dst, a, steps, nullptr); }
abc_status abc_rotate_rows_add(abc_ctx *c, abc_ct *dst, const abc_ct *a, int steps, const abc_ct *addend) {
  return rotate_impl(c, dst, a, steps, addend);
}

// ---- plaintext operands
abc_status abc_add_plain_pt(abc_ctx *c, abc_ct *dst, const abc_ct *a, const abc_pt *pt) {
  if (!valid_ct(c, dst) || !valid_ct(c, a) || !pt || pt->ctx != c) return fail(c, ABC_ERR_PARAM, "invalid handle");
  CHECK_POISON(c);
  TRY(ct_resolve(c, a));
  const u64 *pa = a->b->d;
  TRY(ct_make_private(c, dst));
  return plain_addsub_device<0>(c, dst->b->d, pa, pt->d, pt->broadcast);
}
abc_status abc_sub_plain_pt(abc_ctx *c, abc_ct *dst, const abc_ct *a, const abc_pt *pt) {
  if (!valid_ct(c, dst) || !valid_ct(c, a) || !pt || pt->ctx != c) return fail(c, ABC_ERR_PARAM, "invalid handle");
  CHECK_POISON(c);
  TRY(ct_resolve(c, a));
  const u64 *pa = a->b->d;
  TRY(ct_make_private(c, dst));
  return plain_addsub_device<1>(c, dst->b->d, pa, pt->d, pt->broadcast);
}
abc_status abc_mul_plain_pt(abc_ctx *c, abc_ct *dst, const abc_ct *a, const abc_pt *pt) {
  if (!valid_ct(c, dst) || !valid_ct(c, a) || !pt || pt->ctx != c) return fail(c, ABC_ERR_PARAM, "invalid handle");
  CHECK_POISON(c);
  TRY(ct_resolve(c, a));
  const u64 *pa = a->b->d;
  TRY(ct_make_private(c, dst));
  return mul_plain_device(c, dst->b->d, pa, pt->d, pt->broadcast);
}
#define PLAIN_OP(NAME, PTFN)                                                                                      \
  abc_status NAME(abc_ctx *c, abc_ct *dst, const abc_ct *a, const int64_t *slots, size_t n, int broadcast) {      \
    abc_pt *pt = nullptr;                                                                                         \
    TRY(abc_pt_encode(c, slots, n, broadcast, &pt));                                                              \
    abc_status s = PTFN(c, dst, a, pt);                                                                           \
    abc_pt_free(pt);                                                                                              \
    return s;                                                                                                     \
  }
PLAIN_OP(abc_add_plain, abc_add_plain_pt)
PLAIN_OP(abc_sub_plain, abc_sub_plain_pt)
PLAIN_OP(abc_mul_plain, abc_mul_plain_pt)

// ---- probes
abc_status abc_probe_ntt(abc_ctx *c, int inverse, uint32_t mod_index, uint64_t *host_rows, size_t n_rows) {
  CK(cudaSetDevice(c->device));
  if ((int)mod_index > c->idx_t || n_rows == 0 || n_rows > 65535) return fail(c, ABC_ERR_PARAM, "invalid probe arguments");
  const int N = c->N;
  u64 *d = nullptr; int *rm = nullptr;
  TRY(salloc(c, &d, n_rows * N));
  CK(cudaMallocAsync((void **)&rm, n_rows * sizeof(int), c->stream));
  std::vector<int> v(n_rows, (int)mod_index);
  CK(cudaMemcpyAsync(rm, v.data(), n_rows * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(d, host_rows, n_rows * N * 8, cudaMemcpyHostToDevice, c->stream));
  LimbJob j = blank_job();
  j.dst = d; j.src = d; j.rowmod = rm;
  if (inverse) TRY(launch_limb(c, LIMB_INV, c->hmods[mod_index].ar_class, j, (int)n_rows, 1, "probe_intt"));
  else TRY(launch_limb(c, LIMB_FWD, c->hmods[mod_index].ar_class, j, (int)n_rows, 1, "probe_ntt"));
  CK(cudaMemcpyAsync(host_rows, d, n_rows * N * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  sfree(c, d); sfree(c, rm);
  return ABC_OK;
}
abc_status abc_bench_ntt(abc_ctx *c, int inverse, uint32_t mod_index, size_t n_rows, int iters, float *ms) {
  CK(cudaSetDevice(c->device));
  if ((int)mod_index > c->idx_t || n_rows == 0) return fail(c, ABC_ERR_PARAM, "invalid arguments");
  const int N = c->N;
  const int W = n_rows > 32768 ? 32768 : (int)n_rows, By = (int)((n_rows + W - 1) / W);
  u64 *d = nullptr; int *rm = nullptr;
  TRY(salloc(c, &d, (size_t)W * By * N));
  CK(cudaMemsetAsync(d, 0, (size_t)W * By * N * 8, c->stream));
  CK(cudaMallocAsync((void **)&rm, W * sizeof(int), c->stream));
  std::vector<int> v(W, (int)mod_index);
  CK(cudaMemcpyAsync(rm, v.data(), W * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  LimbJob j = blank_job();
  j.dst = d; j.src = d; j.rowmod = rm; j.dst_is = j.src_is = (long long)W * N;
  for (int it = -1; it < iters; ++it) {
    if (it == 0) CK(cudaEventRecord(c->ev0, c->stream));
    if (inverse) TRY(launch_limb(c, LIMB_INV, c->hmods[mod_index].ar_class, j, W, By, "bench_intt"));
    else TRY(launch_limb(c, LIMB_FWD, c->hmods[mod_index].ar_class, j, W, By, "bench_ntt"));
  }
  CK(cudaEventRecord(c->ev1, c->stream));
  CK(cudaEventSynchronize(c->ev1));
  CK(cudaEventElapsedTime(ms, c->ev0, c->ev1));
  sfree(c, d); sfree(c, rm);
  return ABC_OK;
}
abc_status abc_probe_multiply(abc_ctx *c, const abc_ct *a, const abc_ct *b, uint64_t *host_out3, size_t words) {
  if (!valid_ct(c, a) || !valid_ct(c, b)) return fail(c, ABC_ERR_PARAM, "invalid ciphertext handle");
  const size_t need = (size_t)c->B * 3 * c->L * c->N;
  if (words != need) return fail(c, ABC_ERR_PARAM, "size mismatch");
  u64 *out3 = nullptr;
  TRY(salloc(c, &out3, need));
  TRY(ct_resolve(c, a)); TRY(ct_resolve(c, b));
  TRY(behz_multiply(c, a->b->d, b->b->d, out3));
  CK(cudaMemcpyAsync(host_out3, out3, need * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  sfree(c, out3);
  return ABC_OK;
}

// ---- limb sharding across GPUs (one process per GPU)
abc_status abc_comm_unique_id(abc_ctx *c, uint8_t *out128) {
  NcclApi *n = nccl_api();
  if (!n) return fail(c, ABC_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded");
  static_assert(sizeof(ncclUniqueId) == ABC_COMM_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  NCK(n->GetUniqueId(&id));
  memcpy(out128, &id, sizeof id);
  return ABC_OK;
}
abc_status abc_comm_init(abc_ctx *c, int rank, int world, const uint8_t *id128) {
  NcclApi *n = nccl_api();
  if (!n) return fail(c, ABC_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded");
  if (world < 1 || rank < 0 || rank >= world || world > c->L) return fail(c, ABC_ERR_PARAM, "need 1 <= world <= L and 0 <= rank < world");
  if (c->comm) return fail(c, ABC_ERR_STATE, "communicator already initialised");
  CK(cudaSetDevice(c->device));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  NCK(n->CommInitRank(&c->comm, world, id, rank));
  c->rank = rank; c->world = world;
  if (!c->comm_stream) {
    CK(cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->comm_ready, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->comm_done, cudaEventDisableTiming));
  }
  if (const char *e = getenv("ABC_SHARD_OVERLAP")) c->shard_overlap = atoi(e) != 0;
  if (const char *e = getenv("ABC_SHARD_COLS")) c->shard_cols = atoi(e) != 0;
  limb_range(c->L, world, rank, &c->own_lo, &c->own_hi);
  // the ranks jointly hold ONE ciphertext per handle, so an encryption must draw the same (u, e0, e1) on every rank:
  // rank 0's encryption salt replaces the per-context one (unless abc_set_encrypt_nonce already zeroed it everywhere)
  if (world > 1) {
    u64 *d_salt = nullptr;
    CK(cudaMalloc((void **)&d_salt, sizeof(u64)));
    CK(cudaMemcpyAsync(d_salt, &c->enc_salt, sizeof(u64), cudaMemcpyHostToDevice, c->stream));
    NCK(n->Broadcast(d_salt, d_salt, 1, ncclUint64, 0, c->comm, c->stream));
    CK(cudaMemcpyAsync(&c->enc_salt, d_salt, sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    cudaFree(d_salt);
  }
  CK(cudaStreamSynchronize(c->stream));
  return build_shard_maps(c);
}
abc_status abc_comm_stats(const abc_ctx *c, uint64_t *bytes_received, uint64_t *nccl_calls) {
  if (bytes_received) *bytes_received = c->gathered_bytes;
  if (nccl_calls) *nccl_calls = c->gather_calls;
  return ABC_OK;
}
int abc_comm_rank(const abc_ctx *c) { return c->rank; }
int abc_comm_world(const abc_ctx *c) { return c->world; }
abc_status abc_owned_limbs(const abc_ctx *c, uint32_t *lo, uint32_t *hi) {
  *lo = (uint32_t)c->own_lo; *hi = (uint32_t)c->own_hi;
  return ABC_OK;
}
abc_status abc_ct_allgather(abc_ctx *c, abc_ct *ct) {
  if (!valid_ct(c, ct)) return fail(c, ABC_ERR_PARAM, "invalid ciphertext handle");
  TRY(ct_resolve(c, ct));
  return allgather_limbs(c, ct->b->d, 3);
}

// ---- timing / profiling
abc_status abc_timer_start(abc_ctx *c) { CK(cudaEventRecord(c->ev0, c->stream)); return ABC_OK; }
abc_status abc_timer_stop(abc_ctx *c, float *ms) {
  CK(cudaEventRecord(c->ev1, c->stream));
  CK(cudaEventSynchronize(c->ev1));
  CK(cudaEventElapsedTime(ms, c->ev0, c->ev1));
  return ABC_OK;
}
abc_status abc_flush_l2(abc_ctx *c, size_t bytes) {
  if (bytes > c->flush_bytes) {
    if (c->flush_buf) cudaFree(c->flush_buf);
    c->flush_buf = nullptr; c->flush_bytes = 0;
    CK(cudaMalloc(&c->flush_buf, bytes));
    c->flush_bytes = bytes;
  }
  CK(cudaMemsetAsync(c->flush_buf, 0x5a, bytes, c->stream));
  return ABC_OK;
}
uint64_t abc_launch_count(const abc_ctx *c) { return c->launches; }
uint64_t abc_key_switch_count(const abc_ctx *c) { return c->key_switches; }
abc_status abc_profile_enable(abc_ctx *c, int on) {
  CK(cudaStreamSynchronize(c->stream));
  for (auto &r : c->prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  c->prof_recs.clear();
  c->prof = on != 0;
  return ABC_OK;
}
const char *abc_profile_json(abc_ctx *c) {
  cudaStreamSynchronize(c->stream);
  std::map<std::string, std::pair<long, double>> agg;
  std::vector<std::string> order;
  for (auto &r : c->prof_recs) {
    float ms = 0;
    cudaEventElapsedTime(&ms, r.a, r.b);
    if (!agg.count(r.name)) order.push_back(r.name);
    agg[r.name].first++; agg[r.name].second += ms;
  }
  std::string s = "[";
  for (size_t i = 0; i < order.size(); ++i) {
    char buf[256];
    snprintf(buf, sizeof buf, "%s{\"kernel\": \"%s\", \"launches\": %ld, \"ms\": %.6f}", i ? ", " : "", order[i].c_str(),
             agg[order[i]].first, agg[order[i]].second);
    s += buf;
  }
  s += "]";
  c->prof_json = s;
  return c->prof_json.c_str();
}

abc_status abc_measure_int_peak(abc_ctx *c, double *imad_per_s, double *iadd_per_s) {
  CK(cudaSetDevice(c->device));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
  const int grid = sms * 2, iters = 2048;
  u32 *out = nullptr;
  CK(cudaMallocAsync((void **)&out, (size_t)grid * 1024 * 8, c->stream));
  const double ops = (double)grid * 1024 * iters * 16 * 8;
  float ms = 0;
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaEventRecord(c->ev0, c->stream));
    k_peak_imad<<<grid, 1024, 0, c->stream>>>(out, iters);
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaEventSynchronize(c->ev1));
    CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  }
  if (imad_per_s) *imad_per_s = ops / (ms * 1e-3);
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaEventRecord(c->ev0, c->stream));
    k_peak_iadd<<<grid, 1024, 0, c->stream>>>(out, iters);
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaEventSynchronize(c->ev1));
    CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  }
  if (iadd_per_s) *iadd_per_s = 2.0 * ops / (ms * 1e-3);  // one IADD + one LOP3 per step
  sfree(c, out);
  return ABC_OK;
}

int abc_ntt_arith_class(const abc_ctx *c) { return c->force_ar >= 0 && c->force_ar < c->ar_q ? c->force_ar : c->ar_q; }

abc_status abc_measure_butterfly_peak(abc_ctx *c, int arith_class, double *butterflies_per_s) {
  CK(cudaSetDevice(c->device));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
  const int grid = sms * 2, iters = 512;
  u64 *out = nullptr;
  TRY(salloc(c, &out, (size_t)grid * 1024));
  const u64 q = c->primes[0], w = q / 3;
  double wd = (double)w / (double)q;
  u64 wc_fp; memcpy(&wc_fp, &wd, 8);
  const double n = (double)grid * 1024 * iters * 8 * 4;
  float ms = 0;
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaEventRecord(c->ev0, c->stream));
    if (arith_class == AR_SHOUP) k_peak_butterfly<AR_SHOUP><<<grid, 1024, 0, c->stream>>>(out, iters, q, w, hm::shoup(w, q));
    else if (arith_class == AR_FP) k_peak_butterfly<AR_FP><<<grid, 1024, 0, c->stream>>>(out, iters, q, w, wc_fp);
    else if (arith_class == 3) k_peak_butterfly_f64<<<grid, 1024, 0, c->stream>>>((double *)out, iters, (double)q, (double)w, wd);
    else k_peak_butterfly<AR_FP_LAZY><<<grid, 1024, 0, c->stream>>>(out, iters, q, w, wc_fp);
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaEventSynchronize(c->ev1));
    CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  }
  if (butterflies_per_s) *butterflies_per_s = n / (ms * 1e-3);
  sfree(c, out);
  return ABC_OK;
}

}  // extern "C"
