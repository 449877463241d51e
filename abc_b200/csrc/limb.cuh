// limb.cuh — the limb pipeline kernel: one CTA = one RNS limb staged whole in shared memory:
//   load (with a pre-op) -> [forward NTT] -> [pointwise * mul row] -> [inverse NTT] -> store (with a post-op)
// Every NTT of the BFV path runs through this kernel, fused with the elementwise work either side of it
// (ModUp reduction, plaintext lift, sampling, BatchEncoder scatter/gather, dyadic product, c0 accumulate).
#pragma once
#include "devconst.cuh"
#include "ntt.cuh"

enum { PRE_LOAD = 0, PRE_REDUCE = 1, PRE_PLAIN_LIFT = 2, PRE_TERNARY = 3, PRE_CBD = 4, PRE_ENCODE = 5, PRE_GALOIS_REDUCE = 6 };
enum { POST_STORE = 0, POST_ADD = 1, POST_DECODE = 2, POST_MODDOWN = 3 };

struct LimbJob {
  u64 *dst; const u64 *src; const u64 *mul; const u64 *add;
  long long dst_is, src_is, mul_is, add_is;  // per-instance strides in words (0 = shared by all instances)
  const int *rowmod;                          // [W] modulus index of row w
  const int *rowdst;                          // [W] destination row, or nullptr = w
  const int *rowsrc;                          // [W] source row, or nullptr = destination row
  const int *rowmul;                          // [W] row of `mul`/`add`, or nullptr = w
  // sampler (PRE_TERNARY / PRE_CBD): stream = stream_key(seed, domain, a0 + inst, b)
  u64 seed, domain, a0, b;
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
  int L, k;
};

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
  LIMB_NCOMBOS = 13
};

__device__ __forceinline__ u64 small_to_mod(int v, u64 q) { return v < 0 ? q - (u64)(-v) : (u64)v; }
__device__ __forceinline__ int sample_ternary(u64 h, u64 idx) {
  u64 r = mix64(h ^ idx);
  return (int)(((r >> 32) * 3) >> 32) - 1;
}
__device__ __forceinline__ int sample_cbd(u64 h, u64 idx) {
  u64 r = mix64(h ^ idx);
  return __popcll(r & 0x1fffffULL) - __popcll((r >> 21) & 0x1fffffULL);
}

template <int LOGN, int PRE, bool FWD, bool MUL, bool INV, int POST, int AR>
__global__ void __launch_bounds__(NttDims<LOGN>::T, NttDims<LOGN>::MINB) k_limb(LimbJob job,
                                                                               const ModInfo *__restrict__ mods) {
  typedef NttDims<LOGN> D;
  extern __shared__ __align__(16) u64 sm[];
  const int tid = threadIdx.x, w = blockIdx.x, inst = blockIdx.y;
  const ModInfo M = mods[job.rowmod[w]];
  const u64 q = M.q;
  const int drow = job.rowdst ? job.rowdst[w] : w;
  const int srow = job.rowsrc ? job.rowsrc[w] : drow;

  // ---- load
  if (PRE == PRE_TERNARY || PRE == PRE_CBD) {
    const u64 h = stream_key(job.seed, job.domain, job.a0 + (u64)inst, job.b);
    for (int e = tid; e < D::N; e += D::T) {
      int v = (PRE == PRE_TERNARY) ? sample_ternary(h, (u64)e) : sample_cbd(h, (u64)e);
      sm[swz(e)] = small_to_mod(v, q);
    }
  } else if (PRE == PRE_ENCODE) {
    // BatchEncoder::encode (SealCiphertextFactory.cpp:102-132): pad with the last value, scatter by the index map
    const long long *sl = job.slots_in + (size_t)inst * job.slots_is;
    for (int e = tid; e < D::N; e += D::T) {
      long long v = sl[e < job.n_slots ? e : job.n_slots - 1];
      sm[swz((int)job.index_map[e])] = v < 0 ? q + (u64)v : (u64)v;
    }
  } else if (PRE == PRE_GALOIS_REDUCE) {
    // GaloisTool::apply_galois as a gather (negation is modulo the SOURCE limb's prime), then the ModUp reduction
    const u64 *src = job.src + (size_t)inst * job.src_is + (size_t)srow * D::N;
    const u64 qs = mods[srow].q;
    const u32 einv = job.galois_einv, m2 = 2u * D::N - 1;
    for (int e2 = tid; e2 < D::N / 2; e2 += D::T) {
      const u32 r0 = ((u32)(2 * e2) * einv) & m2, r1 = (r0 + einv) & m2;
      ulonglong2 v;
      v.x = src[r0 & (D::N - 1)]; v.y = src[r1 & (D::N - 1)];
      if (r0 >= (u32)D::N) v.x = neg_mod(v.x, qs);
      if (r1 >= (u32)D::N) v.y = neg_mod(v.y, qs);
      v.x = barrett64(v.x, q, M.mu_hi); v.y = barrett64(v.y, q, M.mu_hi);
      *reinterpret_cast<ulonglong2 *>(&sm[swz(2 * e2)]) = v;
    }
  } else {
    const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(job.src + (size_t)inst * job.src_is + (size_t)srow * D::N);
    for (int e2 = tid; e2 < D::N / 2; e2 += D::T) {
      ulonglong2 v = src[e2];
      if (PRE == PRE_REDUCE) { v.x = barrett64(v.x, q, M.mu_hi); v.y = barrett64(v.y, q, M.mu_hi); }
      if (PRE == PRE_PLAIN_LIFT) {
        // multiply_plain_normal: centred lift of a mod-t coefficient into [0,q)
        const u64 th = job.t_half_up, inc = q - job.t;
        v.x = v.x >= th ? v.x + inc : v.x;
        v.y = v.y >= th ? v.y + inc : v.y;
      }
      *reinterpret_cast<ulonglong2 *>(&sm[swz(2 * e2)]) = v;
    }
  }
  __syncthreads();

  if (FWD) ntt_fwd_smem<LOGN, AR>(sm, M, 1u, tid);

  if (MUL) {
    const int mrow = job.rowmul ? job.rowmul[w] : w;
    const ulonglong2 *mp = reinterpret_cast<const ulonglong2 *>(job.mul + (size_t)inst * job.mul_is + (size_t)mrow * D::N);
    for (int e2 = tid; e2 < D::N / 2; e2 += D::T) {
      ulonglong2 m = mp[e2];
      ulonglong2 *p = reinterpret_cast<ulonglong2 *>(&sm[swz(2 * e2)]);
      ulonglong2 v = *p;
      v.x = mul_mod(v.x, m.x, q, M.mu_hi, M.mu_lo);
      v.y = mul_mod(v.y, m.y, q, M.mu_hi, M.mu_lo);
      *p = v;
    }
    __syncthreads();
  }

  if (INV) ntt_inv_smem<LOGN, true, AR>(sm, M, 1u, tid);

  // ---- store
  if (POST == POST_DECODE) {
    // BatchEncoder::decode (SealCiphertextFactory.cpp:151): gather by the index map, centre to signed
    long long *out = job.slots_out + (size_t)inst * D::N;
    const u64 half = q >> 1;
    for (int e = tid; e < D::N; e += D::T) {
      u64 v = sm[swz((int)job.index_map[e])];
      out[e] = v > half ? (long long)v - (long long)q : (long long)v;
    }
  } else if (POST == POST_MODDOWN) {
    // tail of switch_key_inplace for row (comp, i): dst = base + p^-1 * (acc_i - ([acc_L + p/2]_p mod q_i) + [p/2]_{q_i})
    const DevConst *C = job.C;
    const int comp = w / job.L, i = w - comp * job.L;
    const u64 p = C->p, p_half = C->p_half, phm = C->p_half_mod_q[i], ip = C->inv_p[i], ips = C->inv_p_s[i];
    const ulonglong2 *tl = reinterpret_cast<const ulonglong2 *>(job.tl + (size_t)inst * job.tl_is + (size_t)(comp * job.k + job.L) * D::N);
    const u64 *base = comp == 0 ? job.base0 : job.base1;
    if (base) base += (size_t)inst * (comp == 0 ? job.base0_is : job.base1_is) + (size_t)i * D::N;
    ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(job.dst + (size_t)inst * job.dst_is + (size_t)drow * D::N);
    const u32 einv = job.base_einv, m2 = 2u * D::N - 1;
    for (int e2 = tid; e2 < D::N / 2; e2 += D::T) {
      ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(&sm[swz(2 * e2)]);
      const ulonglong2 t = tl[e2];
      const u64 rx = sub_mod(barrett64(add_mod(t.x, p_half, p), q, M.mu_hi), phm, q);
      const u64 ry = sub_mod(barrett64(add_mod(t.y, p_half, p), q, M.mu_hi), phm, q);
      v.x = mul_shoup(sub_mod(csub(v.x, q), rx, q), ip, ips, q);
      v.y = mul_shoup(sub_mod(csub(v.y, q), ry, q), ip, ips, q);
      if (base) {
        ulonglong2 b;
        if (einv) {
          const u32 r0 = ((u32)(2 * e2) * einv) & m2, r1 = (r0 + einv) & m2;
          b.x = base[r0 & (D::N - 1)]; b.y = base[r1 & (D::N - 1)];
          if (r0 >= (u32)D::N) b.x = neg_mod(b.x, q);
          if (r1 >= (u32)D::N) b.y = neg_mod(b.y, q);
        } else {
          b = reinterpret_cast<const ulonglong2 *>(base)[e2];
        }
        v.x = add_mod(v.x, b.x, q); v.y = add_mod(v.y, b.y, q);
      }
      dst[e2] = v;
    }
  } else {
    ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(job.dst + (size_t)inst * job.dst_is + (size_t)drow * D::N);
    const ulonglong2 *ad = nullptr;
    if (POST == POST_ADD) {
      const int arow = job.rowmul ? job.rowmul[w] : w;
      ad = reinterpret_cast<const ulonglong2 *>(job.add + (size_t)inst * job.add_is + (size_t)arow * D::N);
    }
    for (int e2 = tid; e2 < D::N / 2; e2 += D::T) {
      ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(&sm[swz(2 * e2)]);
      if (INV) { v.x = csub(v.x, q); v.y = csub(v.y, q); }
      if (POST == POST_ADD) {
        ulonglong2 a = ad[e2];
        v.x = add_mod(v.x, a.x, q); v.y = add_mod(v.y, a.y, q);
      }
      dst[e2] = v;
    }
  }
}

// ---- launcher, instantiated once per LOGN in its own translation unit (limb_12.cu, limb_13.cu, limb_14.cu)
// returns a cudaError_t as int; W rows x B instances
template <int LOGN>
int limb_dispatch(int combo, int ar, const LimbJob &job, const ModInfo *mods, int W, int B, cudaStream_t stream);

#ifdef ABC_LIMB_IMPL
template <int LOGN, int PRE, bool FWD, bool MUL, bool INV, int POST, int AR>
static int limb_launch(const LimbJob &job, const ModInfo *mods, int W, int B, cudaStream_t stream) {
  typedef NttDims<LOGN> D;
  auto kern = k_limb<LOGN, PRE, FWD, MUL, INV, POST, AR>;
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
template <int LOGN, int AR>
static int limb_dispatch_ar(int combo, const LimbJob &j, const ModInfo *m, int W, int B, cudaStream_t s) {
  switch (combo) {
    case LIMB_FWD: return limb_launch<LOGN, PRE_LOAD, true, false, false, POST_STORE, AR>(j, m, W, B, s);
    case LIMB_INV: return limb_launch<LOGN, PRE_LOAD, false, false, true, POST_STORE, AR>(j, m, W, B, s);
    case LIMB_REDUCE_FWD: return limb_launch<LOGN, PRE_REDUCE, true, false, false, POST_STORE, AR>(j, m, W, B, s);
    case LIMB_PLAINLIFT_FWD: return limb_launch<LOGN, PRE_PLAIN_LIFT, true, false, false, POST_STORE, AR>(j, m, W, B, s);
    case LIMB_TERNARY_FWD: return limb_launch<LOGN, PRE_TERNARY, true, false, false, POST_STORE, AR>(j, m, W, B, s);
    case LIMB_CBD_FWD: return limb_launch<LOGN, PRE_CBD, true, false, false, POST_STORE, AR>(j, m, W, B, s);
    case LIMB_ENCODE_INV: return limb_launch<LOGN, PRE_ENCODE, false, false, true, POST_STORE, AR>(j, m, W, B, s);
    case LIMB_FWD_DECODE: return limb_launch<LOGN, PRE_LOAD, true, false, false, POST_DECODE, AR>(j, m, W, B, s);
    case LIMB_FWD_MUL_INV: return limb_launch<LOGN, PRE_LOAD, true, true, true, POST_STORE, AR>(j, m, W, B, s);
    case LIMB_FWD_MUL_INV_ADD: return limb_launch<LOGN, PRE_LOAD, true, true, true, POST_ADD, AR>(j, m, W, B, s);
    case LIMB_MUL_INV: return limb_launch<LOGN, PRE_LOAD, false, true, true, POST_STORE, AR>(j, m, W, B, s);
    case LIMB_GALOIS_REDUCE_FWD: return limb_launch<LOGN, PRE_GALOIS_REDUCE, true, false, false, POST_STORE, AR>(j, m, W, B, s);
    case LIMB_INV_MODDOWN: return limb_launch<LOGN, PRE_LOAD, false, false, true, POST_MODDOWN, AR>(j, m, W, B, s);
    default: return (int)cudaErrorInvalidValue;
  }
}
template <int LOGN>
int limb_dispatch(int combo, int ar, const LimbJob &job, const ModInfo *mods, int W, int B, cudaStream_t stream) {
  switch (ar) {
    case AR_SHOUP: return limb_dispatch_ar<LOGN, AR_SHOUP>(combo, job, mods, W, B, stream);
    case AR_FP: return limb_dispatch_ar<LOGN, AR_FP>(combo, job, mods, W, B, stream);
    case AR_FP_LAZY: return limb_dispatch_ar<LOGN, AR_FP_LAZY>(combo, job, mods, W, B, stream);
    default: return (int)cudaErrorInvalidValue;
  }
}
#endif
