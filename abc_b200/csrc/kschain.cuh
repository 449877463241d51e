// kschain.cuh — descriptor + launcher of the chained key switch (kschain.cu)
#pragma once
#include "limb.cuh"

struct KsChain {
  LimbJob up;          // ModUp + NTT rows (t_image = 1, done = counters)
  LimbJob tail;        // inner product + INTT + ModDown rows (flags / done set)
  const uint2 *sched;  // [n_blocks] in dependency order: x = role << 31 | inst, y = row | modulus << 8 | drow << 16 | srow << 24
  int n_blocks;
  u32 *ticket; u32 ticket_base;   // schedule position of a block = the ticket it takes when it starts (limb.cuh grid_ticket)
};

// returns a cudaError_t as int; logN in {12, 13}, every key-level prime < 2^45
int ks_chain_launch(int logN, const KsChain &ch, const ModInfo *mods, cudaStream_t stream);
// persistent grid of `grid` CTAs taking tickets in a loop (logN in {12, 13}); consumes n_blocks + grid tickets
int ks_chain_launch_persistent(int logN, const KsChain &ch, const ModInfo *mods, cudaStream_t stream, int grid);
