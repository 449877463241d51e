// Two-pass limb pipeline for N = 2^(13+A), A in {2, 3} (N = 32768, 65536): head kernels (first/last A stages in
// registers, straight from global memory, carrying the pre-/post-ops) + k_limb<13> in TAIL mode on the 2^A
// contiguous 8192-coefficient blocks of every limb.  Shoup arithmetic: these sizes use 55-60-bit primes.
#define ABC_LIMB_IMPL
#include "limb.cuh"

namespace {
template <int A, int PRE> int head_fwd(const LimbJob &j, const ModInfo *m, int W, int B, cudaStream_t s) {
  k_head_fwd<A, PRE><<<dim3((j.n >> (A + 1)) / 256, W, B), 256, 0, s>>>(j, m);
  return (int)cudaGetLastError();
}
template <int A, int POST> int head_inv(const LimbJob &j, const ModInfo *m, int W, int B, cudaStream_t s) {
  k_head_inv<A, POST><<<dim3((j.n >> (A + 1)) / 256, W, B), 256, 0, s>>>(j, m);
  return (int)cudaGetLastError();
}
template <bool FWD, bool MUL, bool INV> int tail(const LimbJob &j, const ModInfo *m, int W, int B, cudaStream_t s) {
  return limb_launch<13, PRE_LOAD, FWD, MUL, INV, POST_STORE, AR_SHOUP, true>(j, m, W << j.sub, B, s);
}
// same rows, reading what was just written to dst
LimbJob in_place_on_dst(const LimbJob &j) {
  LimbJob t = j;
  t.src = j.dst; t.src_is = j.dst_is; t.rowsrc = nullptr;
  return t;
}
// same rows, writing back over src (the inverse tail clobbers its source)
LimbJob in_place_on_src(const LimbJob &j) {
  LimbJob t = j;
  t.dst = const_cast<u64 *>(j.src); t.dst_is = j.src_is;
  t.rowdst = j.rowsrc ? j.rowsrc : j.rowdst; t.rowsrc = nullptr;
  return t;
}

template <int A> int dispatch(int combo, const LimbJob &j, const ModInfo *m, int W, int B, cudaStream_t s) {
  int e = 0;
#define STEP(x) do { e = (x); if (e) return e; } while (0)
  switch (combo) {
    case LIMB_FWD: STEP((head_fwd<A, PRE_LOAD>(j, m, W, B, s))); return tail<true, false, false>(in_place_on_dst(j), m, W, B, s);
    case LIMB_REDUCE_FWD: STEP((head_fwd<A, PRE_REDUCE>(j, m, W, B, s))); return tail<true, false, false>(in_place_on_dst(j), m, W, B, s);
    case LIMB_GALOIS_REDUCE_FWD: STEP((head_fwd<A, PRE_GALOIS_REDUCE>(j, m, W, B, s))); return tail<true, false, false>(in_place_on_dst(j), m, W, B, s);
    case LIMB_PLAINLIFT_FWD: STEP((head_fwd<A, PRE_PLAIN_LIFT>(j, m, W, B, s))); return tail<true, false, false>(in_place_on_dst(j), m, W, B, s);
    case LIMB_TERNARY_FWD: STEP((head_fwd<A, PRE_TERNARY>(j, m, W, B, s))); return tail<true, false, false>(in_place_on_dst(j), m, W, B, s);
    case LIMB_CBD_FWD: STEP((head_fwd<A, PRE_CBD>(j, m, W, B, s))); return tail<true, false, false>(in_place_on_dst(j), m, W, B, s);
    case LIMB_INV: STEP((tail<false, false, true>(in_place_on_src(j), m, W, B, s))); return head_inv<A, POST_STORE>(j, m, W, B, s);
    case LIMB_INV_MODDOWN: STEP((tail<false, false, true>(in_place_on_src(j), m, W, B, s))); return head_inv<A, POST_MODDOWN>(j, m, W, B, s);
    case LIMB_FWD_MUL_INV:
    case LIMB_FWD_MUL_INV_ADD: {
      STEP((head_fwd<A, PRE_LOAD>(j, m, W, B, s)));
      const LimbJob d = in_place_on_dst(j);
      STEP((tail<true, true, true>(d, m, W, B, s)));
      return combo == LIMB_FWD_MUL_INV ? head_inv<A, POST_STORE>(d, m, W, B, s) : head_inv<A, POST_ADD>(d, m, W, B, s);
    }
    case LIMB_MUL_INV: STEP((tail<false, true, true>(j, m, W, B, s))); return head_inv<A, POST_STORE>(in_place_on_dst(j), m, W, B, s);
    default: return (int)cudaErrorNotSupported;  // encode / decode: scatter / gather kernels + LIMB_INV / LIMB_FWD
  }
#undef STEP
}
}  // namespace

int limb_dispatch_big(int A, int combo, const LimbJob &job, const ModInfo *mods, int W, int B, cudaStream_t stream) {
  LimbJob j = job;
  j.n = 8192 << A; j.sub = A;
  if (A == 2) return dispatch<2>(combo, j, mods, W, B, stream);
  if (A == 3) return dispatch<3>(combo, j, mods, W, B, stream);
  return (int)cudaErrorInvalidValue;
}
