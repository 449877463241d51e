// kschain.cu — the two-launch key switch of the exact-double class (ModUp + NTT rows, then inner product + INTT + ModDown
// rows; limb.cuh) as ONE grid in dependency order.
//
// Why: as two launches the ModUp block T (k*L rows per instance, 1.25 MiB at N = 8192) of the whole batch goes out to HBM
// and comes back; the tail rows then spend more time waiting for T and the key than transforming.  Chained, the ModUp
// rows of instance g + S1 are dispatched together with the tail rows of instance g (linear block order), so only about
// S1 instances' worth of T is alive at any time (L2-resident), and an SM holds a transform-bound ModUp CTA next to a
// load-bound tail CTA instead of two of the same kind.  A tail row waits (acquire on a counter its L ModUp rows bump
// after their bulk stores complete) only for blocks that took a smaller ticket (limb.cuh grid_ticket: the schedule is
// indexed by the order in which blocks actually start, not by blockIdx), so the grid cannot deadlock.
// Schedule (host-built, abc_ctx::ks_sched): entry.x = role << 31 | inst, entry.y = row | modulus << 8 | drow << 16 | srow << 24.
#define ABC_LIMB_IMPL
#include "kschain.cuh"

namespace {

template <int LOGN, bool GAL>
__global__ void __launch_bounds__(NttDims<LOGN>::T, NttDims<LOGN>::MINB) k_ks_chain(KsChain ch, const ModInfo *__restrict__ mods) {
  // schedule entry: x = role << 31 | inst, y = w | modulus << 8 | drow << 16 | srow << 24
  const uint2 s = __ldg(ch.sched + grid_ticket(ch.ticket, ch.ticket_base));
  const int inst = (int)(s.x & 0x7fffffffu), w = (int)(s.y & 0xff);
  const RowIds ids{(int)((s.y >> 8) & 0xff), (int)((s.y >> 16) & 0xff), (int)(s.y >> 24)};
  if ((s.x >> 31) == 0)
    limb_body<LOGN, GAL ? PRE_GALOIS_REDUCE : PRE_REDUCE, true, false, false, POST_STORE, AR_F64, false>(ch.up, mods, inst, w, ids);
  else
    limb_body<LOGN, PRE_KS_INNER, false, false, true, POST_MODDOWN, AR_F64, false>(ch.tail, mods, inst, w, ids);
}

// The same rows on a PERSISTENT grid (ABC_KS_PERSIST=1): 2 CTAs per SM take tickets in a loop; the ticket and schedule
// entry of a CTA's NEXT row are fetched by a second warp while the current row is being processed, so the atomic + load
// round trips and the CTA launch / retire of every row leave the critical path.
// (one row as a call: inlined into the loop, the job descriptors' fields get hoisted out of it and spill — 1.4 KiB per thread;
// the descriptors sit in constant memory so that their fields are immediate-offset loads as kernel parameters are)
#define KSC_SLOTS 8
__constant__ KsChain c_chain[KSC_SLOTS];
template <int LOGN, bool GAL>
__device__ __noinline__ void ks_chain_row(int slot, const ModInfo *__restrict__ mods, uint2 s) {
  const KsChain *ch = &c_chain[slot];
  const int inst = (int)(s.x & 0x7fffffffu), w = (int)(s.y & 0xff);
  const RowIds ids{(int)((s.y >> 8) & 0xff), (int)((s.y >> 16) & 0xff), (int)(s.y >> 24)};
  if ((s.x >> 31) == 0)
    limb_body<LOGN, GAL ? PRE_GALOIS_REDUCE : PRE_REDUCE, true, false, false, POST_STORE, AR_F64, false>(ch->up, mods, inst, w, ids);
  else
    limb_body<LOGN, PRE_KS_INNER, false, false, true, POST_MODDOWN, AR_F64, false>(ch->tail, mods, inst, w, ids);
}
template <int LOGN, bool GAL>
__global__ void __launch_bounds__(NttDims<LOGN>::T, NttDims<LOGN>::MINB) k_ks_chain_p(int slot, const ModInfo *__restrict__ mods) {
  const KsChain &ch = c_chain[slot];
  __shared__ uint2 s_cur, s_next;
  __shared__ unsigned s_tk, s_ntk;
  const int tid = threadIdx.x;
  if (tid == 0) {
    const unsigned t = atomicAdd(ch.ticket, 1u) - ch.ticket_base;
    s_tk = t;
    if (t < (unsigned)ch.n_blocks) s_cur = __ldg(ch.sched + t);
  }
  __syncthreads();
  for (;;) {
    const unsigned t = s_tk;
    if (t >= (unsigned)ch.n_blocks) break;
    const uint2 s = s_cur;
    if (tid == 32) {   // another warp: the next row's ticket and descriptor, off the critical path
      const unsigned tn = atomicAdd(ch.ticket, 1u) - ch.ticket_base;
      s_ntk = tn;
      if (tn < (unsigned)ch.n_blocks) s_next = __ldg(ch.sched + tn);
    }
    ks_chain_row<LOGN, GAL>(slot, mods, s);
    __syncthreads();   // the row is done with shared memory; s_next / s_ntk are visible
    if (tid == 0) { s_tk = s_ntk; s_cur = s_next; }
    __syncthreads();
  }
}

template <int LOGN, bool GAL>
int launch_p(const KsChain &ch, const ModInfo *mods, cudaStream_t stream, int grid) {
  typedef NttDims<LOGN> D;
  auto kern = k_ks_chain_p<LOGN, GAL>;
  static bool done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!done[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D::SMEM);
    if (e != cudaSuccess) return (int)e;
    done[dev & 63] = true;
  }
  // one constant-memory slot per launch in flight on this device, round robin (launches of one context are stream-ordered;
  // KSC_SLOTS bounds how many key switches of DIFFERENT contexts may be enqueued on a device at the same time)
  static unsigned next_slot[64] = {0};
  const int slot = (int)(next_slot[dev & 63]++ % KSC_SLOTS);
  cudaError_t e = cudaMemcpyToSymbolAsync(c_chain, &ch, sizeof(KsChain), (size_t)slot * sizeof(KsChain), cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) return (int)e;
  kern<<<(unsigned)grid, D::T, D::SMEM, stream>>>(slot, mods);
  return (int)cudaGetLastError();
}

template <int LOGN, bool GAL>
int launch(const KsChain &ch, const ModInfo *mods, cudaStream_t stream) {
  typedef NttDims<LOGN> D;
  auto kern = k_ks_chain<LOGN, GAL>;
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
  kern<<<(unsigned)ch.n_blocks, D::T, D::SMEM, stream>>>(ch, mods);
  return (int)cudaGetLastError();
}

}  // namespace

// persistent variant: `grid` CTAs (2 per SM); consumes n_blocks + grid tickets
int ks_chain_launch_persistent(int logN, const KsChain &ch, const ModInfo *mods, cudaStream_t stream, int grid) {
  const bool gal = ch.up.galois_einv != 0;
  switch (logN) {
    case 12: return gal ? launch_p<12, true>(ch, mods, stream, grid) : launch_p<12, false>(ch, mods, stream, grid);
    case 13: return gal ? launch_p<13, true>(ch, mods, stream, grid) : launch_p<13, false>(ch, mods, stream, grid);
    default: return (int)cudaErrorInvalidValue;
  }
}
int ks_chain_launch(int logN, const KsChain &ch, const ModInfo *mods, cudaStream_t stream) {
  const bool gal = ch.up.galois_einv != 0;
  switch (logN) {
    case 12: return gal ? launch<12, true>(ch, mods, stream) : launch<12, false>(ch, mods, stream);
    case 13: return gal ? launch<13, true>(ch, mods, stream) : launch<13, false>(ch, mods, stream);
    case 14: return gal ? launch<14, true>(ch, mods, stream) : launch<14, false>(ch, mods, stream);
    default: return (int)cudaErrorInvalidValue;
  }
}
