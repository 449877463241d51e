// ksfused.cuh — job descriptor + launcher of the single-launch key switch (ksfused.cu)
#pragma once
#include "limb.cuh"

struct KsJob {
  const u64 *target; long long target_is;    // polynomial being switched: [L][N] per instance, coefficient form
  const double *key;                          // KSwitchKey [L][2][k][N], NTT form at key level, as exact doubles
  u64 *dst; long long dst_is;                 // result ciphertext [2][L][N]
  u64 *dst2;                                  // with `add`: also the result without the addend (layout of dst)
  const u64 *add; long long add_is;           // a whole ciphertext accumulated into the result (rotate + add)
  const u64 *base0, *base1; long long base0_is, base1_is;  // polynomial added into component 0 / 1 (nullptr = 0)
  u32 einv;                                   // automorphism applied to target and bases while reading (0: none)
  u64 *tl; long long tl_is;                   // [2][N] per instance: INTT_p(acc_L[c]), published by the special unit
  u32 *fault;                                 // host-mapped word raised when a dependency wait gives up (limb.cuh wait_word)
  u32 *flags; u32 serial; int skew;           // flags[inst][c] == serial when tl[inst][c] is ready
  u32 *ticket; u32 ticket_base;               // unit of a block = the ticket it takes when it starts (limb.cuh grid_ticket)
  const DevConst *C;
  const int *Iset; int nI;                    // output moduli: the data limbs this rank owns, then the special prime
  int L, k, B;
  int threads;                                // N = 8192: 1024 (default) or 512 threads per CTA, one CTA per SM either way
};

// returns a cudaError_t as int; logN in {12, 13}, every key-level prime < 2^45
int ks_fused_launch(int logN, const KsJob &job, const ModInfo *mods, cudaStream_t stream);
