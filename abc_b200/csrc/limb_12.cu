// limb pipeline kernels for N = 2^12 (one translation unit per size so the sizes compile in parallel)
#define ABC_LIMB_IMPL
#define ABC_LIMB_IMPL_SIZE
#include "limb.cuh"
template int limb_dispatch<12>(int, int, const LimbJob &, const ModInfo *, int, int, cudaStream_t);
