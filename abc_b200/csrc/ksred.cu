// ksred.cu — the key switch of the exact-double class (Evaluator::switch_key_inplace of SEAL 3.6.5, reached from
// SealCiphertext.cpp:55 rotate_rows and :104-105 multiply + relinearize) as ONE dependency-ordered grid in which the ModUp
// block T[I][J] = NTT_I(x_J) never exists: every ModUp row multiplies its transform by the two key rows while it is still
// in registers and ADDS the products into the accumulators acc[c][I] = sum_J T[I][J] * key[J][c][I] with a bulk reduction
// (cp.reduce.async.bulk .add.f64, shared -> global, performed by the L2 atomic units: 4.7 TB/s measured, tools/ubench.cu).
//
// Why: kschain.cu's tail rows read L rows of T and L rows of key per output row (512 KiB at N = 8192, L = 4) through
// the SM's 68 B/clk L2 port — 7.7 k cycles per row against 3.6 k cycles of FP64 work: the inner product was port-bound
// (timing what-if: 18 % of the key switch).  Here a ModUp row reads its two key rows (128 KiB), and a tail row receives
// ONE finished accumulator row by a bulk copy, so it is load -> INTT -> ModDown like any other limb-pipeline row.
// Every term is an exact integer below 0.6 q and a sum below 2.4 q, so the floating-point additions are exact and the
// order in which the L2 performs them does not show: results are bit-identical to the other key-switch paths.
//
// Rows (host-built schedule, abc_ctx::ksr_sched; a block's position is the ticket it takes when it starts, limb.cuh):
//   ModUp row (inst, I, J):  bulk copy of target limb J -> [Galois gather, convert] -> forward NTT mod q_I (no Barrett:
//                            exact doubles) -> in registers: P_c = T * key[J][c][I] -> shared -> bulk reduce into
//                            acc[inst % ring][I][c], c = 0, 1 -> done[inst][I] += 1
//   tail row (inst, I, c):   wait done[inst][I] == L -> bulk copy of acc row -> zero it for the slot's next user ->
//                            INTT -> freed[inst][I] += 1 -> special prime: publish INTT_p(acc_L) + flag;
//                            data prime: wait flag, ModDown (+ sigma(c0) / base, + addend), store
// The accumulators live in a ring of `ring` instances (L2-resident); a ModUp row of instance g waits for the tail rows
// of instance g - ring before its first reduction.  Every wait points at a smaller ticket: no deadlock.
#define ABC_LIMB_IMPL
#include "ksred.cuh"

#define KSR_ZBYTES 4096

namespace {

__device__ __forceinline__ void mbar_wait0(u32 mb) {
  u32 ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(mb) : "memory");
  } while (!ok);
}

// ---- ModUp row: target limb J -> NTT mod q_I -> both key products -> accumulators
template <int LOGN, bool GAL>
__device__ __forceinline__ void ksred_up(const KsRed &ks, const ModInfo *__restrict__ mods, int inst, int I, int J, u64 *sm,
                                         u64 *mbar) {
  typedef NttDims<LOGN> D;
  typedef NttLast<LOGN> P;
  constexpr int AR = AR_F64, G = P::GROUPS;
  constexpr bool P16 = UsePlan16<LOGN, AR, 0>::value;
  const int tid = threadIdx.x;
  const ModInfo M = mods[I];
  const u64 q = M.q, aux = ar_aux<AR>(q);
  const double qinv = f64_of(M.qinv_bits);
  bulk_row_to_smem(sm, ks.target + (size_t)inst * ks.target_is + (size_t)J * D::N, (u32)D::SMEM, mbar, tid);
  ntt_fwd_smem_mids<LOGN, AR, true>(sm, M, 1u, tid, mods[J].q, GAL ? ks.einv : 0u);

  // contiguous pass in registers; T = x stays there for both products
  u64 x[G][8];
  int prow[G];
  const double2 *k0[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const int vt = P16 ? p16_block8(tid, g) : tid + g * D::T;
    prow[g] = swz_row8(vt);
    k0[g] = reinterpret_cast<const double2 *>(ks.key + (((size_t)J * 2) * ks.k + I) * D::N) + 4 * vt;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(&sm[prow[g] ^ (2 * i)]);
      x[g][2 * i] = v.x; x[g][2 * i + 1] = v.y;
    }
  }
#pragma unroll
  for (int g = 0; g < G; ++g)
    ntt_fwd_last_math<LOGN, AR, 0, P16 ? 0 : P::NSH>(x[g], M.twd, 1u, q, aux, P16 ? p16_block8(tid, g) : tid + g * D::T, qinv);

  // P_0 = T * key[J][0][I] into the thread's own shared-memory slots (the image of the swizzled limb)
#pragma unroll
  for (int g = 0; g < G; ++g)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double2 kv = __ldg(k0[g] + i);
      ulonglong2 p;
      p.x = mul_tw<AR>(x[g][2 * i], bits_of(kv.x), M.qinv_bits, q, aux);
      p.y = mul_tw<AR>(x[g][2 * i + 1], bits_of(kv.y), M.qinv_bits, q, aux);
      *reinterpret_cast<ulonglong2 *>(&sm[prow[g] ^ (2 * i)]) = p;
    }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the bulk reduction
  __syncthreads();
  double *a0 = ks.acc + (((size_t)(inst % ks.ring) * ks.k + I) * 2) * D::N;
  const u32 smaddr = (u32)__cvta_generic_to_shared(sm);
  if (tid == 0) {
    // the accumulator slot's previous user (instance inst - ring) has taken and zeroed its rows
    if (inst >= ks.ring) wait_word<true>(ks.freed + (size_t)(inst - ks.ring) * ks.k + I, ks.freed_target, ks.fault);
    asm volatile("fence.proxy.async.global;" ::: "memory");
#if !(ABC_WHATIF & 32)
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;"
                 ::"l"(a0), "r"(smaddr), "r"((u32)D::SMEM) : "memory");
#endif
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  // P_1 = T * key[J][1][I] in place in the registers while the reduction reads shared memory
#pragma unroll
  for (int g = 0; g < G; ++g)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double2 kv = __ldg(k0[g] + (size_t)ks.k * (D::N / 2) + i);
      x[g][2 * i] = mul_tw<AR>(x[g][2 * i], bits_of(kv.x), M.qinv_bits, q, aux);
      x[g][2 * i + 1] = mul_tw<AR>(x[g][2 * i + 1], bits_of(kv.y), M.qinv_bits, q, aux);
    }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // P_0 has been read out
  __syncthreads();
#pragma unroll
  for (int g = 0; g < G; ++g)
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<ulonglong2 *>(&sm[prow[g] ^ (2 * i)]) = make_ulonglong2(x[g][2 * i], x[g][2 * i + 1]);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
#if !(ABC_WHATIF & 32)
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;"
                 ::"l"(a0 + D::N), "r"(smaddr), "r"((u32)D::SMEM) : "memory");
#endif
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // both reductions performed
    __threadfence();
    atomicAdd(ks.done + (size_t)inst * ks.k + I, 1u);
  }
}

// ---- tail row: accumulator row (inst, I, comp) -> INTT -> publish (special prime) / ModDown + store (data prime)
template <int LOGN>
__device__ __forceinline__ void ksred_tail(const KsRed &ks, const ModInfo *__restrict__ mods, int inst, int I, int comp, u64 *sm,
                                           u64 *mbar, const double *zeros) {
  typedef NttDims<LOGN> D;
  constexpr int AR = AR_F64;
  const int tid = threadIdx.x;
  const ModInfo M = mods[I];
  const u64 q = M.q, aux = ar_aux<AR>(q);
  double *arow = ks.acc + (((size_t)(inst % ks.ring) * ks.k + I) * 2 + comp) * D::N;
  const u32 mb = (u32)__cvta_generic_to_shared(mbar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    wait_word<true>(ks.done + (size_t)inst * ks.k + I, ks.done_target, ks.fault);   // the L ModUp rows of this modulus
    asm volatile("fence.proxy.async.global;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"((u32)D::SMEM) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((u32)__cvta_generic_to_shared(sm)), "l"(arow), "r"((u32)D::SMEM), "r"(mb) : "memory");
  }
  __syncthreads();
  mbar_wait0(mb);
  // the row is ours: leave zeros behind for the slot's next user (instance inst + ring).  Bulk stores from a small zeroed
  // buffer, issued by one thread and awaited after the transform: no store instructions, no fence in front of the INTT
  // (plain stores + __threadfence() by every thread cost 6 % of the key switch: the fence sits on the critical path).
#if !(ABC_WHATIF & 64)
  if (tid == 0) {
    const u32 za = (u32)__cvta_generic_to_shared(zeros);
#pragma unroll 1
    for (int off = 0; off < (int)D::SMEM; off += KSR_ZBYTES)
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                   ::"l"(reinterpret_cast<char *>(arow) + off), "r"(za), "r"((u32)KSR_ZBYTES) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
#endif
  ntt_inv_smem<LOGN, true, AR, true>(sm, M, 1u, tid);   // sums of L centred products, |x| <= 2.4 q: inside the range plan
  if (tid == 0) {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __threadfence();
    atomicAdd(ks.freed + (size_t)inst * ks.k + I, 1u);
  }

  constexpr int N = D::N;
  ulonglong2 *tlp = reinterpret_cast<ulonglong2 *>(ks.tl + (size_t)inst * ks.tl_is + (size_t)comp * N);
  if (I == ks.L) {   // special prime: INTT_p(acc_L[comp]) as canonical residues for the data rows of the instance
    for (int e2 = tid; e2 < N / 2; e2 += D::T) {
      ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(&sm[swz_pair(tid, e2)]);
      v.x = canon_inv<AR>(v.x, q, aux); v.y = canon_inv<AR>(v.y, q, aux);
      tlp[e2] = v;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) atomicExch(ks.flags + inst * 2 + comp, ks.flag_serial);
    return;
  }
  if (tid == 0) wait_word<false>(ks.flags + inst * 2 + comp, ks.flag_serial, ks.fault);
  __syncthreads();
  const DevConst *C = ks.C;
  ModDownRow md;
  md.p = C->p; md.p_half = C->p_half; md.phm = C->p_half_mod_q[I]; md.ip = C->inv_p[I]; md.ips = C->inv_p_s[I];
  md.tl = tlp;
  md.base = comp == 0 ? ks.base0 : ks.base1;
  if (md.base) md.base += (size_t)inst * (comp == 0 ? ks.base0_is : ks.base1_is) + (size_t)I * N;
  const size_t drow = (size_t)comp * ks.L + I;
  moddown_store_f64<LOGN, D::T>(
      sm, M, md, ks.einv, reinterpret_cast<ulonglong2 *>(ks.dst + (size_t)inst * ks.dst_is + drow * N),
      ks.dst2 ? reinterpret_cast<ulonglong2 *>(ks.dst2 + (size_t)inst * ks.dst_is + drow * N) : nullptr,
      ks.add ? reinterpret_cast<const ulonglong2 *>(ks.add + (size_t)inst * ks.add_is + drow * N) : nullptr, tid);
}

template <int LOGN, bool GAL>
__global__ void __launch_bounds__(NttDims<LOGN>::T, NttDims<LOGN>::MINB) k_ks_red(KsRed ks, const ModInfo *__restrict__ mods) {
  extern __shared__ __align__(128) u64 sm[];
  __shared__ __align__(8) u64 mbar;
  __shared__ __align__(128) double zeros[KSR_ZBYTES / 8];
  for (int i = threadIdx.x; i < KSR_ZBYTES / 8; i += NttDims<LOGN>::T) zeros[i] = 0.0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // (the ticket's barrier below orders it before any bulk store)
  // schedule entry: x = role << 31 | inst, y = w | modulus << 8 | drow << 16 | srow << 24
  const uint2 s = __ldg(ks.sched + grid_ticket(ks.ticket, ks.ticket_base));
  const int inst = (int)(s.x & 0x7fffffffu), I = (int)((s.y >> 8) & 0xff);
  if ((s.x >> 31) == 0) ksred_up<LOGN, GAL>(ks, mods, inst, I, (int)(s.y >> 24), sm, &mbar);
  else ksred_tail<LOGN>(ks, mods, inst, I, (int)(s.y & 0xff) < 2 ? (int)(s.y & 0xff) : ((int)(s.y >> 24) >= ks.k ? 1 : 0), sm, &mbar, zeros);
}

template <int LOGN, bool GAL>
int launch(const KsRed &ks, const ModInfo *mods, cudaStream_t stream) {
  typedef NttDims<LOGN> D;
  auto kern = k_ks_red<LOGN, GAL>;
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
  kern<<<(unsigned)ks.n_blocks, D::T, D::SMEM, stream>>>(ks, mods);
  return (int)cudaGetLastError();
}

}  // namespace

int ks_red_launch(int logN, const KsRed &ks, const ModInfo *mods, cudaStream_t stream) {
  const bool gal = ks.einv != 0;
  switch (logN) {
    case 12: return gal ? launch<12, true>(ks, mods, stream) : launch<12, false>(ks, mods, stream);
    case 13: return gal ? launch<13, true>(ks, mods, stream) : launch<13, false>(ks, mods, stream);
    default: return (int)cudaErrorInvalidValue;
  }
}
