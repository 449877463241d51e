// ksfused.cu — the whole key switch (Evaluator::switch_key_inplace of SEAL 3.6.5, reached from
// SealCiphertext.cpp:55 rotate_rows and :104-105 multiply + relinearize) as ONE launch on the exact-double class
// (every key-level prime < 2^45: SEAL's N = 4096 / 8192 defaults).
//
// Unit of work = (instance, output modulus I): one CTA keeps the two accumulators  acc_c = sum_J NTT_I(x_J) * key[J][c][I]
// (c = 0,1) in shared memory while it runs the L forward transforms of ModUp one after another, then inverse-transforms
// both and finishes with ModDown.  The ModUp block T[inst][I][J] that the two-launch path writes to and re-reads from HBM
// (k*L rows per instance: 2/3 of the key switch's traffic) never exists.
//
//   for J in 0..L-1:  bulk copy of target limb J (cp.async.bulk, issued while the previous J is still in its last pass)
//                     strided passes, first one gathering through the Galois map (no Barrett: exact doubles)
//                     contiguous pass in registers -> acc_c += x * key[J][c][I]   (thread-private shared-memory slots)
//   for c in 0,1:     INTT (contiguous pass straight from the accumulator), then
//     I == L (special prime p): publish INTT_p(acc_L[c]) through L2 + a flag
//     I <  L:                   wait for that flag, ModDown (+ sigma(c0) / base, + addend), store
//
// The special-prime unit of instance g + skew is dispatched with the data units of instance g (linear block order), so
// a data unit only ever waits for a block with a smaller ticket (the order blocks start in, limb.cuh grid_ticket): no
// deadlock whatever the dispatch order, and normally no waiting at all.
// Shared memory: N words of transform buffer + 2N words of accumulators (192 KiB at N = 8192: one 1024-thread CTA per
// SM; 96 KiB at N = 4096: two 512-thread CTAs).
#include "ksfused.cuh"

#ifndef ABC_KS_REG_EPI
#define ABC_KS_REG_EPI 0  /* ModDown fed from the registers of the last inverse pass (0: through shared memory) */
#endif

namespace {

__device__ __forceinline__ void mbar_wait(u32 mb, u32 parity) {
  u32 ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(mb), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_issue(u32 dst, const u64 *src, u32 bytes, u32 mb) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(mb) : "memory");
}

// ---- tails fed from the registers of the last inverse pass (e = coefficient index, xb = centred double, |x| < q)
struct PublishSpecial {   // special-prime unit: INTT_p(acc_L[c]) as canonical residues, for the data units of the instance
  u64 *tl; u64 q, aux;
  __device__ __forceinline__ void operator()(int e, u64 xb) const { tl[e] = canon_inv<AR_F64>(xb, q, aux); }
};
template <int LOGN> struct ModDownStore {  // data unit: ModDown (+ base through the automorphism, + addend), store
  ModDownF64 f;
  const u64 *tl, *base, *add;
  u64 *out, *out2;
  u32 einv; u64 q;
  __device__ __forceinline__ void operator()(int e, u64 xb) const {
    constexpr u32 N = 1u << LOGN;
    const u64 t = __ldcg(tl + e);  // written by another CTA of this launch: L2, not L1
    u64 b = 0;
    if (base) {
      if (einv) {
        const u32 r0 = ((u32)e * einv) & (2u * N - 1);
        b = base[r0 & (N - 1)];
        if (r0 >= N) b = neg_mod(b, q);
      } else {
        b = base[e];
      }
    }
    double r = moddown_one_f64(f64_of(xb), t, base != nullptr, b, f);
    if (add) {
      if (out2) out2[e] = f64_canon_bits(r);
      r = add_canon_f64(r, add[e], f.qd);
    }
    out[e] = f64_canon_bits(r);
  }
};

template <int LOGN, int TT>
__global__ void __launch_bounds__(NttDims<LOGN, TT>::T, NttDims<LOGN, TT>::MINB) k_ks_fused(KsJob job,
                                                                                          const ModInfo *__restrict__ mods) {
  typedef NttDims<LOGN, TT> D;
  typedef NttPlan<LOGN> P;
  constexpr int AR = AR_F64, T = D::T, N = D::N, G = NttLast<LOGN, TT>::GROUPS;
  static_assert(D::IT == G, "groups of 8 coefficients per thread and pass");
  extern __shared__ __align__(16) u64 sm[];
  double *accs = reinterpret_cast<double *>(sm + N);  // [2][G][8][T]: slot (c, g, r) of thread tid, thread-private
  __shared__ __align__(8) u64 mbar;
  __shared__ unsigned rd_count;
  const int tid = threadIdx.x;
  const int L = job.L, k = job.k;

  // ---- unit of this block: special units run `skew` instances ahead of their data units
  int inst, unit;  // unit 0 = special prime, 1.. = data modulus Iset[unit - 1]
  {
    const int nI = job.nI, nd = nI - 1, Bn = job.B, S = job.skew < Bn ? job.skew : Bn;
    const int b = (int)grid_ticket(job.ticket, job.ticket_base);
    if (b < S) { inst = b; unit = 0; }
    else {
      const int b1 = b - S, full = (Bn - S) * nI;
      if (b1 < full) { const int g = b1 / nI, r = b1 - g * nI; inst = r == 0 ? g + S : g; unit = r; }
      else { const int b2 = b1 - full, g = b2 / nd; inst = Bn - S + g; unit = 1 + (b2 - g * nd); }
    }
  }
  const int I = unit == 0 ? L : job.Iset[unit - 1];
  const ModInfo M = mods[I];
  const u64 q = M.q, aux = ar_aux<AR>(q);
  const double qinv = f64_of(M.qinv_bits), qd = f64_of(aux);
  const u32 mb = (u32)__cvta_generic_to_shared(&mbar), smaddr = (u32)__cvta_generic_to_shared(sm);
  const u64 *tgt = job.target + (size_t)inst * job.target_is;

  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
    rd_count = 0;
    bulk_issue(smaddr, tgt, (u32)D::SMEM, mb);
  }
  __syncthreads();  // the barrier object is initialised before anyone polls it

  u64 x[G][8];
  for (int J = 0; J < L; ++J) {
    mbar_wait(mb, (u32)(J & 1));
    ntt_fwd_smem_mids<LOGN, AR, true, TT>(sm, M, 1u, tid, mods[J].q, job.einv);
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(&sm[swz(8 * (tid + g * T) + 2 * i)]);
        x[g][2 * i] = v.x; x[g][2 * i + 1] = v.y;
      }
    // the buffer is dead once every warp has taken its coefficients: the last warp to get here starts the copy of the
    // next target limb, which then lands while the CTA is still busy with this pass and the key products
    __syncwarp();
    if ((tid & 31) == 0) {
      __threadfence_block();
      const unsigned old = atomicAdd(&rd_count, 1u);
      if (old == (unsigned)((J + 1) * (T / 32) - 1) && J + 1 < L) bulk_issue(smaddr, tgt + (size_t)(J + 1) * N, (u32)D::SMEM, mb);
    }
    // key[J][0][I] is requested before the butterflies of this pass and key[J][1][I] while component 0 is multiplied:
    // the L2 latency of the key rows hides behind arithmetic instead of sitting in front of it
    const double2 *kp0 = reinterpret_cast<const double2 *>(job.key + (((size_t)J * 2) * k + I) * N) + 4 * tid;
    const double2 *kp1 = kp0 + (size_t)k * (N / 2);
    double2 kv[G][4];
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int i = 0; i < 4; ++i) kv[g][i] = __ldg(kp0 + g * 4 * T + i);
#pragma unroll
    for (int g = 0; g < G; ++g) ntt_fwd_last_math<LOGN, AR, TT>(x[g], M.twd, 1u, q, aux, tid + g * T, qinv);
    // acc_c += x * key[J][c][I]
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
      for (int g = 0; g < G; ++g) {
        double *a = accs + (size_t)((c * G + g) * 8) * T + tid;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const double2 k2 = kv[g][i];
          if (c == 0) kv[g][i] = __ldg(kp1 + g * 4 * T + i);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int r = 2 * i + h;
            const double kd = h ? k2.y : k2.x;
            const double v = f64_of(mul_tw<AR>(x[g][r], bits_of(kd), bits_of(qinv), q, aux));
            if (J == 0) a[r * T] = v; else a[r * T] += v;
          }
        }
      }
    }
  }

  // ---- inverse transforms + tail, one component at a time.  What the tails read is requested now: the addend rows go
  // to L2, and sigma(c0) — a gather with 8 useful bytes per 32-byte sector from global memory — is staged whole in
  // component 0's accumulator slots (dead after the first inverse pass) by one bulk copy, then gathered from there.
  const bool stage_base = unit != 0 && job.einv != 0 && job.base0 != nullptr;
  if (unit != 0 && job.add) {
    const char *pa = reinterpret_cast<const char *>(job.add + (size_t)inst * job.add_is + (size_t)I * N);
    for (int line = tid; line < 2 * (N * 8 / 128); line += T) {
      const int c = line / (N * 8 / 128), l = line - c * (N * 8 / 128);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(pa + (size_t)c * L * N * 8 + (size_t)l * 128));
    }
  }
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const double *a = accs + (size_t)((c * G + g) * 8) * T + tid;
#pragma unroll
      for (int r = 0; r < 8; ++r) x[g][r] = bits_of(reduce_f64(a[r * T], qinv, qd));
    }
    if (c == 0 && stage_base) {  // uniform over the CTA: every warp has taken its component-0 sums
      __syncwarp();
      if ((tid & 31) == 0) {
        __threadfence_block();
        const unsigned old = atomicAdd(&rd_count, 1u);
        if (old == (unsigned)((L + 1) * (T / 32) - 1))
          bulk_issue((u32)__cvta_generic_to_shared(accs), job.base0 + (size_t)inst * job.base0_is + (size_t)I * N, (u32)D::SMEM, mb);
      }
    }
#pragma unroll
    for (int g = 0; g < G; ++g) ntt_inv_first_math<LOGN, AR, TT>(x[g], M.itwd, 1u, q, aux, tid + g * T, qinv);
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        ulonglong2 v; v.x = x[g][2 * i]; v.y = x[g][2 * i + 1];
        *reinterpret_cast<ulonglong2 *>(&sm[swz(8 * (tid + g * T) + 2 * i)]) = v;
      }
#if ABC_KS_REG_EPI
    pass_sync<LOGN - P::R0 - P::R1 - P::R2, T>(tid);
    if constexpr (P::R2 > 0) {
      ntt_inv_mid<LOGN, P::R0 + P::R1, P::R2, false, true, AR, TT>(sm, M, 1u, q, aux, tid);
      pass_sync<LOGN - P::R0 - P::R1, T>(tid);
    }
    ntt_inv_mid<LOGN, P::R0, P::R1, false, P::R2 == 0, AR, TT>(sm, M, 1u, q, aux, tid);
    pass_sync<LOGN - P::R0, T>(tid);
    // the last pass hands its outputs (coefficients tid + r * N/8, in registers) straight to the tail
    u64 *tlp = job.tl + (size_t)inst * job.tl_is + (size_t)c * N;
    if (unit == 0) {
      ntt_inv_mid<LOGN, 0, P::R0, true, true, AR, TT>(sm, M, 1u, q, aux, tid, PublishSpecial{tlp, q, aux});
      __threadfence();
      __syncthreads();
      if (tid == 0) atomicExch(job.flags + inst * 2 + c, job.serial);
    } else {
      if ((tid & 31) == 0) wait_word<false>(job.flags + inst * 2 + c, job.serial, job.fault);
      __syncwarp();
      const DevConst *C = job.C;
      ModDownRow md;
      md.p = C->p; md.p_half = C->p_half; md.phm = C->p_half_mod_q[I]; md.ip = C->inv_p[I]; md.ips = C->inv_p_s[I];
      md.tl = nullptr;
      md.base = c == 0 ? job.base0 : job.base1;
      if (md.base) md.base += (size_t)inst * (c == 0 ? job.base0_is : job.base1_is) + (size_t)I * N;
      const size_t drow = (size_t)c * L + I;
      ModDownStore<LOGN> epi;
      epi.f = moddown_f64(md, M);
      epi.tl = tlp; epi.base = md.base; epi.einv = job.einv; epi.q = q;
      epi.out = job.dst + (size_t)inst * job.dst_is + drow * N;
      epi.out2 = job.dst2 ? job.dst2 + (size_t)inst * job.dst_is + drow * N : nullptr;
      epi.add = job.add ? job.add + (size_t)inst * job.add_is + drow * N : nullptr;
      ntt_inv_mid<LOGN, 0, P::R0, true, true, AR, TT>(sm, M, 1u, q, aux, tid, epi);
      __syncthreads();  // this component's reads of the buffer finish before the next one overwrites it
    }
#else
    ntt_inv_smem_mids<LOGN, true, AR, TT>(sm, M, 1u, tid);
    ulonglong2 *tlp = reinterpret_cast<ulonglong2 *>(job.tl + (size_t)inst * job.tl_is + (size_t)c * N);
    if (unit == 0) {
      for (int e2 = tid; e2 < N / 2; e2 += T) {
        ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(&sm[swz_pair(tid, e2)]);
        v.x = canon_inv<AR>(v.x, q, aux); v.y = canon_inv<AR>(v.y, q, aux);
        tlp[e2] = v;
      }
      __threadfence();
      __syncthreads();
      if (tid == 0) atomicExch(job.flags + inst * 2 + c, job.serial);
    } else {
      if (tid == 0) wait_word<false>(job.flags + inst * 2 + c, job.serial, job.fault);
      __syncthreads();
      const DevConst *C = job.C;
      ModDownRow md;
      md.p = C->p; md.p_half = C->p_half; md.phm = C->p_half_mod_q[I]; md.ip = C->inv_p[I]; md.ips = C->inv_p_s[I];
      md.tl = tlp;
      md.base = c == 0 ? job.base0 : job.base1;
      if (md.base) md.base += (size_t)inst * (c == 0 ? job.base0_is : job.base1_is) + (size_t)I * N;
      if (c == 0 && stage_base) { mbar_wait(mb, (u32)(L & 1)); md.base = reinterpret_cast<const u64 *>(accs); }
      const size_t drow = (size_t)c * L + I;
      moddown_store_f64<LOGN, T>(
          sm, M, md, job.einv, reinterpret_cast<ulonglong2 *>(job.dst + (size_t)inst * job.dst_is + drow * N),
          job.dst2 ? reinterpret_cast<ulonglong2 *>(job.dst2 + (size_t)inst * job.dst_is + drow * N) : nullptr,
          job.add ? reinterpret_cast<const ulonglong2 *>(job.add + (size_t)inst * job.add_is + drow * N) : nullptr, tid);
      __syncthreads();  // the epilogue's reads of the buffer finish before the next component overwrites it
    }
#endif
  }
}

template <int LOGN, int TT>
int launch(const KsJob &job, const ModInfo *mods, cudaStream_t stream) {
  typedef NttDims<LOGN, TT> D;
  auto kern = k_ks_fused<LOGN, TT>;
  const size_t smem = 3 * D::SMEM;
  static bool done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!done[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    done[dev & 63] = true;
  }
  kern<<<(unsigned)(job.B * job.nI), D::T, smem, stream>>>(job, mods);
  return (int)cudaGetLastError();
}

}  // namespace

int ks_fused_launch(int logN, const KsJob &job, const ModInfo *mods, cudaStream_t stream) {
  switch (logN) {
    case 12: return launch<12, 0>(job, mods, stream);
    case 13: return job.threads == 512 ? launch<13, 512>(job, mods, stream) : launch<13, 1024>(job, mods, stream);
    default: return (int)cudaErrorInvalidValue;
  }
}
