// ks14.cuh — descriptor + launchers of the split key switch at N = 16384 (ks14.cu)
#pragma once
#include "limb.cuh"

struct Ks14 {
  const u64 *target; long long target_is;     // polynomial being switched: [L][N] per instance, coefficient form
  double *sx;                                 // [B][L][N]: sigma(target) as exact doubles, natural order (k_ks14_prep)
  const double *key;                          // KSwitchKey [L][2][k][N], NTT form at key level, as exact doubles
  double *T;                                  // ModUp block [B][k][L][N]: row (I, J) = two 8192-blocks, each the raw-double image of its swizzled block
  double *xch;                                // [B][2k][N]: block h of accumulator row (c*k + I) after its 13 local inverse stages (image)
  u64 *tl; long long tl_is;                   // accumulator block [2][k][N] per instance: rows (c*k + L) receive INTT_p(acc_L[c]) (canonical)
  u64 *dst; long long dst_is;                 // result ciphertext [2][L][N]
  u64 *dst2;                                  // with `add`: also the result without the addend
  const u64 *add; long long add_is;           // a whole ciphertext accumulated into the result (rotate + add)
  const u64 *base0, *base1; long long base0_is, base1_is;  // polynomial added into component 0 / 1 (nullptr = 0)
  u32 einv;                                   // automorphism applied to target and bases while reading (0: none)
  const uint2 *sched; int n_blocks;           // x = role << 31 | block << 30 | inst, y = w | modulus << 8 | drow << 16 | srow << 24
  u32 *ticket; u32 ticket_base;
  u32 *done; u32 done_target;                 // [B][k][2]: ModUp half-rows of (modulus, block) stored so far (L per launch)
  u32 *xflag; u32 serial;                     // [B][2k][2] == serial when xch[inst][row][block] is written
  u32 *flags;                                 // [B][2][2] == serial when tl[inst][c] block h is published
  u32 *fault;
  const DevConst *C;
  int L, k, B;
};

// returns a cudaError_t as int
// logN = log2 of the whole limb: 14 (rows of 8192-coefficient blocks) or 13 (4096-coefficient blocks, small batches)
int ks14_prep_launch(int logN, const Ks14 &ks, const ModInfo *mods, cudaStream_t stream);
int ks14_launch(int logN, const Ks14 &ks, const ModInfo *mods, cudaStream_t stream);

// ---- the BEHZ block's transforms at N = 16384 on the same half-limb rows (ks14.cu): forward rows (operand limb or lifted
// Bsk row -> stage 0 while loading -> 13 local stages -> raw-double image, reduced to 0.5 q), and inverse rows with the
// tensor product in their load (13 local inverse stages, partner exchange, last stage, canonical store)
struct Behz14 {
  const u64 *X; long long X_is;               // [inst][4][W][N]: the lifted Bsk rows (canonical), read only
  u64 *XI; long long XI_is;                   // [inst][4][W][N]: the transformed rows as raw-double images (two blocks per row); a
                                              // separate block: both half-rows read a whole source row before either stores
  const u64 *a, *b;                           // operand ciphertexts [inst][2][L][N] (rows r < L of polys 0,1 / 2,3 are read from here)
  u64 *Y; long long Y_is;                     // [inst][3][W][N]: the product in coefficient form (canonical)
  double *xch;                                // [B][3W][N]: partner exchange of the inverse rows
  u32 *xflag; u32 serial;                     // [B][3W][2]
  u32 *ticket; u32 ticket_base;
  u32 *fault;
  const int *rowmod;                          // [4W] modulus index of row w
  int W, L, np, square, B;
};
int behz14_fwd_launch(const Behz14 &bz, const ModInfo *mods, cudaStream_t stream);
int behz14_inv_launch(const Behz14 &bz, const ModInfo *mods, cudaStream_t stream);
