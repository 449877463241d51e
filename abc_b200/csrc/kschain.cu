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

int ks_chain_launch(int logN, const KsChain &ch, const ModInfo *mods, cudaStream_t stream) {
  const bool gal = ch.up.galois_einv != 0;
  switch (logN) {
    case 12: return gal ? launch<12, true>(ch, mods, stream) : launch<12, false>(ch, mods, stream);
    case 13: return gal ? launch<13, true>(ch, mods, stream) : launch<13, false>(ch, mods, stream);
    case 14: return gal ? launch<14, true>(ch, mods, stream) : launch<14, false>(ch, mods, stream);
    default: return (int)cudaErrorInvalidValue;
  }
}
