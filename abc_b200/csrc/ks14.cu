// ks14.cu — the key switch at N = 16384 (the reference's default factory size, SealCiphertextFactory.h:13) on the
// exact-double class, as rows of HALF a limb: every 16384-point transform is stage 0 (the one butterfly stage that couples
// the two halves) plus two independent 8192-point blocks, and a block is one 512-thread CTA with 64 KiB of shared memory
// — the N = 8192 kernel's shape, two CTAs per SM, so one CTA's barriers and global-memory phases overlap the other's
// butterflies.  A whole 16384-point limb needs 128 KiB: one 1024-thread CTA per SM, every phase serialised (ncu: 45 % warps
// active, MIO throttle from the two shuffle stages of that size's plan); this grid runs the same work 1.3x faster.
//
//   prep launch   (inst, J):          sigma(target limb J) -> exact doubles, natural order (Galois gather + negation + conversion
//                                     ONCE per source limb instead of once per target modulus)
//   ModUp row     (inst, I, J, h):    b = sx[8192 ..) by bulk copy, a = sx[0 .. 8192) by coalesced loads; stage 0:
//                                     y = a +- w0 * b (block 0: +, block 1: -); 13 local forward stages with twiddle base
//                                     2 + h (ntt.cuh, block of a larger transform); raw-double image -> T[I][J][h]
//   tail row      (inst, I, c, h):    inner product over J on block h; 13 local inverse stages; the block's outputs go to
//                                     xch and a flag; the LAST stage pairs them with the partner block's (u +- v, folded with
//                                     N^-1) while they pass through the ModDown epilogue (special prime: publish)
// Partner rows hold adjacent tickets (a row waits for at most the next ticket, which is the next block to start, and
// otherwise only for smaller ones), so the waits cannot deadlock.
#define ABC_LIMB_IMPL
#include "ks14.cuh"

namespace {

// LG = log2 of the block: 13 for N = 16384 (block = 8192 coefficients, 512 threads, 64 KiB); 12 for N = 8192, where the
// split is used for SMALL batches only (B <= 8: twice as many, half as long rows fill more of the 148 SMs and the
// critical path ModUp row -> special tail row -> data tail row shortens: batch-1 rotateRows 43 -> about 27 us)
constexpr int AR = AR_F64;
#define KS14_DIMS typedef NttDims<LG> D; constexpr int NB = D::N, NGL = 2 * NB, NP = NB / 2 / D::T;

__device__ __forceinline__ void mbar_wait0(u32 mb) {
  u32 ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(mb) : "memory");
  } while (!ok);
}
__device__ __forceinline__ double mulc(double x, double w, const ModInfo &M, u64 q, u64 aux) {
  return f64_of(mul_tw<AR>(bits_of(x), bits_of(w), M.qinv_bits, q, aux));
}

// ---- sigma(target limb J) as doubles, once per source limb.  grid (L, B), 1024 threads, the whole limb in shared memory
template <int LG>
__global__ void __launch_bounds__(1024, 1) k_ks14_prep(Ks14 ks, const ModInfo *__restrict__ mods) {
  KS14_DIMS; (void)NP;
  extern __shared__ __align__(128) u64 sm[];
  __shared__ __align__(8) u64 mbar;
  const int J = blockIdx.x, inst = blockIdx.y, tid = threadIdx.x;
  bulk_row_to_smem(sm, ks.target + (size_t)inst * ks.target_is + (size_t)J * NGL, (u32)(NGL * 8), &mbar, tid);
  const u64 qs = mods[J].q;
  const u32 e1 = ks.einv ? ks.einv : 1u, m2 = 2u * NGL - 1;
  double2 *out = reinterpret_cast<double2 *>(ks.sx + ((size_t)inst * ks.L + J) * NGL);
#pragma unroll 4
  for (int e2 = tid; e2 < NGL / 2; e2 += 1024) {
    // GaloisTool::apply_galois as a gather: out[e] = +-in[e * einv mod 2N]; negation modulo the source prime
    const u32 r0 = ((u32)(2 * e2) * e1) & m2, r1 = (r0 + e1) & m2;
    u64 x = sm[r0 & (NGL - 1)], y = sm[r1 & (NGL - 1)];
    if (r0 >= (u32)NGL) x = neg_mod(x, qs);
    if (r1 >= (u32)NGL) y = neg_mod(y, qs);
    out[e2] = make_double2(f64_of(ar_from_canon<AR>(x)), f64_of(ar_from_canon<AR>(y)));
  }
}

// ---- ModUp half-row
template <int LG>
__device__ __forceinline__ void ks14_up(const Ks14 &ks, const ModInfo *__restrict__ mods, int inst, int I, int J, int h, u64 *sm,
                                        u64 *mbar) {
  KS14_DIMS;
  const int tid = threadIdx.x;
  const ModInfo M = mods[I];
  const u64 q = M.q, aux = ar_aux<AR>(q);
  const double qinv = f64_of(M.qinv_bits), qd = f64_of(aux);
  const u32 twbase = 2u + (u32)h;
  const double *srow = ks.sx + ((size_t)inst * ks.L + J) * NGL;
  const u32 mb = (u32)__cvta_generic_to_shared(mbar);
  if (tid == 0) {   // b = the upper half of the source row, as it lies
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"((u32)D::SMEM) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((u32)__cvta_generic_to_shared(sm)), "l"(srow + NB), "r"((u32)D::SMEM), "r"(mb) : "memory");
  }
  double2 a[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) a[i] = __ldcs(reinterpret_cast<const double2 *>(srow) + tid + i * D::T);
  const double w0 = f64_of(__ldg(reinterpret_cast<const u64 *>(M.twd) + 1));   // the single twiddle of stage 0
  const u64 qs = mods[J].q;
  const bool red_in = qs >= ABC_F64_NARROW_MAX && qs > q;   // a wide source prime above the target: reduce first (ntt.cuh range plan)
  __syncthreads();   // the barrier object is initialised before anyone polls it
  mbar_wait0(mb);
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    double2 b = *reinterpret_cast<const double2 *>(&sm[2 * (tid + i * D::T)]);
    if (red_in) {
      a[i].x = reduce_f64(a[i].x, qinv, qd); a[i].y = reduce_f64(a[i].y, qinv, qd);
      b.x = reduce_f64(b.x, qinv, qd); b.y = reduce_f64(b.y, qinv, qd);
    }
    const double px = mulc(b.x, w0, M, q, aux), py = mulc(b.y, w0, M, q, aux);
    a[i].x = h ? a[i].x - px : a[i].x + px;
    a[i].y = h ? a[i].y - py : a[i].y + py;
  }
  __syncthreads();   // every thread has taken its b's: the buffer becomes the swizzled block
#pragma unroll
  for (int i = 0; i < NP; ++i) *reinterpret_cast<double2 *>(&sm[swz_pair(tid, tid + i * D::T)]) = a[i];
  __syncthreads();
  ntt_fwd_smem_mids<LG, AR, false, 0, true>(sm, M, twbase, tid);
  ntt_fwd_last<LG, AR, 0, true>(sm, M, twbase, q, aux, tid);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    double *g = ks.T + (((size_t)inst * ks.k + I) * ks.L + J) * NGL + (size_t)h * NB;
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(g), "r"((u32)__cvta_generic_to_shared(sm)), "r"((u32)D::SMEM) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __threadfence();
    atomicAdd(ks.done + ((size_t)inst * ks.k + I) * 2 + h, 1u);
  }
}

// ---- tail half-row
template <int LG>
__device__ __forceinline__ void ks14_tail(const Ks14 &ks, const ModInfo *__restrict__ mods, int inst, int I, int comp, int h,
                                          u64 *sm) {
  KS14_DIMS;
  const int tid = threadIdx.x;
  const ModInfo M = mods[I];
  const u64 q = M.q, aux = ar_aux<AR>(q);
  const double qinv = f64_of(M.qinv_bits), qd = f64_of(aux);
  const u32 twbase = 2u + (u32)h;
  // inner product over J on this block: T rows as the raw-double images the ModUp rows left, the key's exact-double copy
  if (tid == 0) wait_word<true>(ks.done + ((size_t)inst * ks.k + I) * 2 + h, ks.done_target, ks.fault);
  __syncthreads();
  {
    const double2 *t = reinterpret_cast<const double2 *>(ks.T + (((size_t)inst * ks.k + I) * ks.L) * NGL + (size_t)h * NB);
    const double2 *kp = reinterpret_cast<const double2 *>(ks.key + ((size_t)comp * ks.k + I) * NGL + (size_t)h * NB);
    if (ks.L % 4 == 0) {   // software-pipelined, four source limbs at a time (SEAL's default here: L = 8)
      ks_inner_rows_f64<4, NB / 2 / D::T, D::T, false>(sm, t, kp, NGL / 2, ks.k * NGL, M, tid);
      for (int J0 = 4; J0 < ks.L; J0 += 4)
        ks_inner_rows_f64<4, NB / 2 / D::T, D::T, true>(sm, t + (size_t)J0 * (NGL / 2), kp + (size_t)J0 * ks.k * NGL, NGL / 2, ks.k * NGL, M, tid);
    } else {
      for (int e2 = tid; e2 < NB / 2; e2 += D::T)
        *reinterpret_cast<ulonglong2 *>(&sm[swz_pair(tid, e2)]) =
            ks_inner_pair_f64(t + swz2(e2), kp + e2, ks.L, NGL / 2, ks.k * NGL, M);
    }
  }
  __syncthreads();
  ntt_inv_first<LG, AR, true>(sm, M, twbase, q, aux, tid);
  ntt_inv_smem_mids<LG, false, AR>(sm, M, twbase, tid);   // 13 local stages, no N^-1 fold; ends with a CTA barrier
  // hand the block to the partner, take the partner's
  const size_t arow = (size_t)inst * 2 * ks.k + (size_t)comp * ks.k + I;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(ks.xch + arow * NGL + (size_t)h * NB), "r"((u32)__cvta_generic_to_shared(sm)), "r"((u32)D::SMEM) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __threadfence();
    atomicExch(ks.xflag + arow * 2 + h, ks.serial);
    wait_word<false>(ks.xflag + arow * 2 + (1 - h), ks.serial, ks.fault);
    if (I != ks.L) wait_word<false>(ks.flags + ((size_t)inst * 2 + comp) * 2 + h, ks.serial, ks.fault);   // data rows: INTT_p(acc_L) of this block's range
  }
  __syncthreads();
  const double2 *part = reinterpret_cast<const double2 *>(ks.xch + arow * NGL + (size_t)(1 - h) * NB);
  const double wlast = f64_of(h ? M.wl_ninv_d : M.ninv_d);
  const bool wide = f64_wide(q);
  // last stage of the whole transform: block 0 keeps (u + v) * N^-1, block 1 (u - v) * (w_last * N^-1); u = block 0's value
  auto last = [&](double mine, double theirs) {
    double z = mulc(h ? theirs - mine : mine + theirs, wlast, M, q, aux);
    if (wide) z = reduce_f64(z, qinv, qd);
    return z;
  };
  const int E0 = h * (NB / 2);   // this block's first pair in the whole limb
  if (I == ks.L) {   // special prime: publish INTT_p(acc_L[comp]) on this block's range
    ulonglong2 *tlp = reinterpret_cast<ulonglong2 *>(ks.tl + (size_t)inst * ks.tl_is + ((size_t)comp * ks.k + ks.L) * NGL) + E0;
    for (int e2 = tid; e2 < NB / 2; e2 += D::T) {
      const double2 m = *reinterpret_cast<const double2 *>(&sm[swz_pair(tid, e2)]), p = __ldcg(part + swz2(e2));
      tlp[e2] = make_ulonglong2(f64_to_canon(last(m.x, p.x), qd), f64_to_canon(last(m.y, p.y), qd));
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) atomicExch(ks.flags + ((size_t)inst * 2 + comp) * 2 + h, ks.serial);
    return;
  }
  // ModDown (+ sigma(base) through the automorphism, + addend) on this block's range
  const DevConst *C = ks.C;
  ModDownRow md;
  md.p = C->p; md.p_half = C->p_half; md.phm = C->p_half_mod_q[I]; md.ip = C->inv_p[I]; md.ips = C->inv_p_s[I];
  md.tl = reinterpret_cast<const ulonglong2 *>(ks.tl + (size_t)inst * ks.tl_is + ((size_t)comp * ks.k + ks.L) * NGL);
  md.base = comp == 0 ? ks.base0 : ks.base1;
  if (md.base) md.base += (size_t)inst * (comp == 0 ? ks.base0_is : ks.base1_is) + (size_t)I * NGL;
  const ModDownF64 f = moddown_f64(md, M);
  const size_t drow = (size_t)comp * ks.L + I;
  ulonglong2 *out = reinterpret_cast<ulonglong2 *>(ks.dst + (size_t)inst * ks.dst_is + drow * NGL);
  ulonglong2 *out2 = ks.dst2 ? reinterpret_cast<ulonglong2 *>(ks.dst2 + (size_t)inst * ks.dst_is + drow * NGL) : nullptr;
  const ulonglong2 *addp = ks.add ? reinterpret_cast<const ulonglong2 *>(ks.add + (size_t)inst * ks.add_is + drow * NGL) : nullptr;
  const u32 einv = ks.einv, m2 = 2u * NGL - 1;
#pragma unroll 2
  for (int e2 = tid; e2 < NB / 2; e2 += D::T) {
    const int E2 = E0 + e2;
    const ulonglong2 t = __ldcg(md.tl + E2);
    ulonglong2 ad = make_ulonglong2(0, 0);
    if (addp) ad = addp[E2];
    const double2 m = *reinterpret_cast<const double2 *>(&sm[swz_pair(tid, e2)]), p = __ldcg(part + swz2(e2));
    ulonglong2 b = make_ulonglong2(0, 0);
    if (md.base) {
      if (einv) {
        const u32 r0 = ((u32)(2 * E2) * einv) & m2, r1 = (r0 + einv) & m2;
        b.x = md.base[r0 & (NGL - 1)]; b.y = md.base[r1 & (NGL - 1)];
        if (r0 >= (u32)NGL) b.x = neg_mod(b.x, q);
        if (r1 >= (u32)NGL) b.y = neg_mod(b.y, q);
      } else {
        b = reinterpret_cast<const ulonglong2 *>(md.base)[E2];
      }
    }
    u64 rx = moddown_finish_int(moddown_core_f64(last(m.x, p.x), t.x, f), md.base != nullptr, b.x, q);
    u64 ry = moddown_finish_int(moddown_core_f64(last(m.y, p.y), t.y, f), md.base != nullptr, b.y, q);
    if (addp) {
      if (out2) out2[E2] = make_ulonglong2(rx, ry);
      rx = add_mod(rx, ad.x, q); ry = add_mod(ry, ad.y, q);
    }
    out[E2] = make_ulonglong2(rx, ry);
  }
}

template <int LG>
__global__ void __launch_bounds__(NttDims<LG>::T, NttDims<LG>::MINB) k_ks14(Ks14 ks, const ModInfo *__restrict__ mods) {
  extern __shared__ __align__(128) u64 sm[];
  __shared__ __align__(8) u64 mbar;
  const uint2 s = __ldg(ks.sched + grid_ticket(ks.ticket, ks.ticket_base));
  const int inst = (int)(s.x & 0x3fffffffu), h = (int)((s.x >> 30) & 1u), I = (int)((s.y >> 8) & 0xff);
  if ((s.x >> 31) == 0) ks14_up<LG>(ks, mods, inst, I, (int)(s.y >> 24), h, sm, &mbar);
  else ks14_tail<LG>(ks, mods, inst, I, (int)(s.y >> 24) >= ks.k ? 1 : 0, h, sm);
}



// ---------------------------------------------------------------- BEHZ block at N = 16384 (LG = 13 only)
// forward half-row: grid = B * np * W * 2 blocks, block index -> (inst, w, h) by ticket (no cross-row dependencies, but the
// same launch shape as the inverse rows)
__global__ void __launch_bounds__(NttDims<13>::T, NttDims<13>::MINB) k_behz14_fwd(Behz14 bz, const ModInfo *__restrict__ mods) {
  constexpr int LG = 13;
  KS14_DIMS;
  extern __shared__ __align__(128) u64 sm[];
  __shared__ __align__(8) u64 mbar;
  const int tid = threadIdx.x;
  const unsigned blk = blockIdx.x;
  const int h = (int)(blk & 1u), w = (int)((blk >> 1) % (unsigned)(bz.np * bz.W)), inst = (int)((blk >> 1) / (unsigned)(bz.np * bz.W));
  const int p = w / bz.W, r = w - p * bz.W;
  const ModInfo M = mods[bz.rowmod[w]];
  const u64 q = M.q, aux = ar_aux<AR>(q);
  const u32 twbase = 2u + (u32)h;
  const u64 *src = r < bz.L ? (p < 2 ? bz.a : bz.b) + (((size_t)inst * 2 + (p & 1)) * bz.L + r) * NGL
                            : bz.X + (size_t)inst * bz.X_is + (size_t)w * NGL;
  const u32 mb = (u32)__cvta_generic_to_shared(&mbar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"((u32)D::SMEM) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((u32)__cvta_generic_to_shared(sm)), "l"(src + NB), "r"((u32)D::SMEM), "r"(mb) : "memory");
  }
  ulonglong2 a[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) a[i] = __ldcs(reinterpret_cast<const ulonglong2 *>(src) + tid + i * D::T);
  const double w0 = f64_of(__ldg(reinterpret_cast<const u64 *>(M.twd) + 1));
  __syncthreads();
  mbar_wait0(mb);
  double2 y[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const ulonglong2 b = *reinterpret_cast<const ulonglong2 *>(&sm[2 * (tid + i * D::T)]);
    const double ax = f64_of(ar_from_canon<AR>(a[i].x)), ay = f64_of(ar_from_canon<AR>(a[i].y));
    const double px = mulc(f64_of(ar_from_canon<AR>(b.x)), w0, M, q, aux), py = mulc(f64_of(ar_from_canon<AR>(b.y)), w0, M, q, aux);
    y[i] = make_double2(h ? ax - px : ax + px, h ? ay - py : ay + py);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NP; ++i) *reinterpret_cast<double2 *>(&sm[swz_pair(tid, tid + i * D::T)]) = y[i];
  __syncthreads();
  ntt_fwd_smem_mids<LG, AR, false, 0, true>(sm, M, twbase, tid);
  ntt_fwd_last<LG, AR, 0, true>(sm, M, twbase, q, aux, tid, true);   // raw, reduced to |x| <= 0.5 q: the images feed products
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    u64 *g = bz.XI + (size_t)inst * bz.XI_is + (size_t)w * NGL + (size_t)h * NB;
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(g), "r"((u32)__cvta_generic_to_shared(sm)), "r"((u32)D::SMEM) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

// inverse half-row with the tensor product in its load
__global__ void __launch_bounds__(NttDims<13>::T, NttDims<13>::MINB) k_behz14_inv(Behz14 bz, const ModInfo *__restrict__ mods) {
  constexpr int LG = 13;
  KS14_DIMS; (void)NP;
  extern __shared__ __align__(128) u64 sm[];
  const int tid = threadIdx.x;
  const unsigned blk = grid_ticket(bz.ticket, bz.ticket_base);   // partner blocks hold adjacent tickets
  const int h = (int)(blk & 1u), w = (int)((blk >> 1) % (unsigned)(3 * bz.W)), inst = (int)((blk >> 1) / (unsigned)(3 * bz.W));
  const int p = w / bz.W, r = w - p * bz.W;
  const ModInfo M = mods[bz.rowmod[w]];
  const u64 q = M.q, aux = ar_aux<AR>(q);
  const double qinv = f64_of(M.qinv_bits), qd = f64_of(aux);
  const u32 twbase = 2u + (u32)h;
  {
    // operand images: row (poly * W + r) of the image block, block h
    const double2 *a0 = reinterpret_cast<const double2 *>(bz.XI + (size_t)inst * bz.XI_is + (size_t)r * NGL + (size_t)h * NB);
    const size_t ps = (size_t)bz.W * (NGL / 2);
    const double2 *a1 = a0 + ps, *b0 = bz.square ? a0 : a0 + 2 * ps, *b1 = bz.square ? a1 : a0 + 3 * ps;
    const double2 *x = p == 2 ? a1 : a0, *y = p == 0 ? b0 : b1;
    for (int j = tid; j < NB / 2; j += D::T) {
      const double2 xv = __ldcs(x + j), yv = __ldcs(y + j);
      double2 o = make_double2(mulc(xv.x, yv.x, M, q, aux), mulc(xv.y, yv.y, M, q, aux));
      if (p == 1) {
        if (bz.square) { o.x += o.x; o.y += o.y; }
        else {
          const double2 uv = __ldcs(a1 + j), vv = __ldcs(b0 + j);
          o.x += mulc(uv.x, vv.x, M, q, aux); o.y += mulc(uv.y, vv.y, M, q, aux);
        }
        o.x = reduce_f64(o.x, qinv, qd); o.y = reduce_f64(o.y, qinv, qd);
      }
      *reinterpret_cast<double2 *>(&sm[2 * j]) = o;
    }
  }
  __syncthreads();
  ntt_inv_first<LG, AR, true>(sm, M, twbase, q, aux, tid);
  ntt_inv_smem_mids<LG, false, AR>(sm, M, twbase, tid);
  const size_t xrow = (size_t)inst * 3 * bz.W + w;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(bz.xch + xrow * NGL + (size_t)h * NB), "r"((u32)__cvta_generic_to_shared(sm)), "r"((u32)D::SMEM) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __threadfence();
    atomicExch(bz.xflag + xrow * 2 + h, bz.serial);
    wait_word<false>(bz.xflag + xrow * 2 + (1 - h), bz.serial, bz.fault);
  }
  __syncthreads();
  const double2 *part = reinterpret_cast<const double2 *>(bz.xch + xrow * NGL + (size_t)(1 - h) * NB);
  const double wlast = f64_of(h ? M.wl_ninv_d : M.ninv_d);
  const bool wide = f64_wide(q);
  ulonglong2 *out = reinterpret_cast<ulonglong2 *>(bz.Y + (size_t)inst * bz.Y_is + (size_t)w * NGL) + h * (NB / 2);
  for (int e2 = tid; e2 < NB / 2; e2 += D::T) {
    const double2 m = *reinterpret_cast<const double2 *>(&sm[swz_pair(tid, e2)]), pv = __ldcg(part + swz2(e2));
    double zx = mulc(h ? pv.x - m.x : m.x + pv.x, wlast, M, q, aux), zy = mulc(h ? pv.y - m.y : m.y + pv.y, wlast, M, q, aux);
    if (wide) { zx = reduce_f64(zx, qinv, qd); zy = reduce_f64(zy, qinv, qd); }
    out[e2] = make_ulonglong2(f64_to_canon(zx, qd), f64_to_canon(zy, qd));
  }
}

template <int LG> int prep_launch(const Ks14 &ks, const ModInfo *mods, cudaStream_t stream) {
  KS14_DIMS; (void)NP;
  static bool done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!done[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(k_ks14_prep<LG>, cudaFuncAttributeMaxDynamicSharedMemorySize, NGL * 8);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_ks14<LG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D::SMEM);
    if (e != cudaSuccess) return (int)e;
    done[dev & 63] = true;
  }
  k_ks14_prep<LG><<<dim3(ks.L, ks.B), 1024, NGL * 8, stream>>>(ks, mods);
  return (int)cudaGetLastError();
}
template <int LG> int main_launch(const Ks14 &ks, const ModInfo *mods, cudaStream_t stream) {
  k_ks14<LG><<<(unsigned)ks.n_blocks, NttDims<LG>::T, NttDims<LG>::SMEM, stream>>>(ks, mods);
  return (int)cudaGetLastError();
}

}  // namespace

// logN = log2 of the whole limb: 14 (blocks of 8192) or 13 (blocks of 4096)
int ks14_prep_launch(int logN, const Ks14 &ks, const ModInfo *mods, cudaStream_t stream) {
  return logN == 14 ? prep_launch<13>(ks, mods, stream) : logN == 13 ? prep_launch<12>(ks, mods, stream) : (int)cudaErrorInvalidValue;
}
int ks14_launch(int logN, const Ks14 &ks, const ModInfo *mods, cudaStream_t stream) {
  return logN == 14 ? main_launch<13>(ks, mods, stream) : logN == 13 ? main_launch<12>(ks, mods, stream) : (int)cudaErrorInvalidValue;
}

int behz14_fwd_launch(const Behz14 &bz, const ModInfo *mods, cudaStream_t stream) {
  static bool done[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!done[dev & 63]) {
    cudaError_t e = cudaFuncSetAttribute(k_behz14_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NttDims<13>::SMEM);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k_behz14_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NttDims<13>::SMEM);
    if (e != cudaSuccess) return (int)e;
    done[dev & 63] = true;
  }
  k_behz14_fwd<<<(unsigned)(bz.B * bz.np * bz.W * 2), NttDims<13>::T, NttDims<13>::SMEM, stream>>>(bz, mods);
  return (int)cudaGetLastError();
}
int behz14_inv_launch(const Behz14 &bz, const ModInfo *mods, cudaStream_t stream) {
  k_behz14_inv<<<(unsigned)(bz.B * 3 * bz.W * 2), NttDims<13>::T, NttDims<13>::SMEM, stream>>>(bz, mods);
  return (int)cudaGetLastError();
}
