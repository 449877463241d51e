// ksred.cuh — descriptor + launcher of the accumulating key switch (ksred.cu)
#pragma once
#include "limb.cuh"

struct KsRed {
  const u64 *target; long long target_is;     // polynomial being switched: [L][N] per instance, coefficient form
  const double *key;                          // KSwitchKey [L][2][k][N], NTT form at key level, as exact doubles
  double *acc; int ring;                      // accumulators [ring][k][2][N]: slot inst % ring, raw doubles, swizzled-image order
  u64 *tl; long long tl_is;                   // [2][N] per instance: INTT_p(acc_L[c]), published by the special rows
  u64 *dst; long long dst_is;                 // result ciphertext [2][L][N]
  u64 *dst2;                                  // with `add`: also the result without the addend (layout of dst)
  const u64 *add; long long add_is;           // a whole ciphertext accumulated into the result (rotate + add)
  const u64 *base0, *base1; long long base0_is, base1_is;  // polynomial added into component 0 / 1 (nullptr = 0)
  u32 einv;                                   // automorphism applied to target and bases while reading (0: none)
  const uint2 *sched; int n_blocks;           // rows in dependency order: x = role << 31 | inst, y = w | modulus << 8 | drow << 16 | srow << 24
  u32 *ticket; u32 ticket_base;               // schedule position of a block = the ticket it takes when it starts
  u32 *done; u32 done_target;                 // [B][k] ModUp rows whose two products are in the accumulators (L per launch)
  u32 *freed; u32 freed_target;               // [B][k] tail rows that have taken their accumulator row and zeroed it (2 per launch)
  u32 *flags; u32 flag_serial;                // [B][2] == serial when tl[inst][c] is ready
  u32 *fault;                                 // host-mapped word raised when a dependency wait gives up
  const DevConst *C;
  int L, k, B;
};

// returns a cudaError_t as int; logN in {12, 13}, every key-level prime < 0.97 * 2^45
int ks_red_launch(int logN, const KsRed &ks, const ModInfo *mods, cudaStream_t stream);
