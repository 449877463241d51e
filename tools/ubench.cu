// ubench.cu — register-resident / memory-system microbenchmarks behind the design decisions of the key switch
// (DESIGN.md "what bounds each kernel").  Standalone: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu
//   bf_magic    exact-double butterfly as ntt.cuh has it: 8 FP64-pipe instructions (quotient via the 1.5*2^52 magic)
//   bf_frnd     the same butterfly with Q = rint(ph * qinv) as DMUL + FRND.F64: 7 FP64-pipe instructions + 1 FRND
//   frnd_only   FRND.F64 issue rate alone (which pipe? how wide?)
//   mul_magic / mul_frnd   modular product alone (6 vs 5 + FRND)
//   bulk_red    cp.reduce.async.bulk .add.f64 shared -> global, 64 KiB per operation, 4 CTAs per destination row
//   l2_read     LDG.128 streaming of an L2-resident buffer, bytes/clk/SM
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
#define MAGIC 6755399441055744.0

__device__ __forceinline__ double frnd(double x) { double r; asm("cvt.rni.f64.f64 %0, %1;" : "=d"(r) : "d"(x)); return r; }

template <int MODE> __device__ __forceinline__ double mulmod(double y, double w, double q, double qinv) {
  const double ph = y * w;
  const double Q = MODE == 0 ? fma(ph, qinv, MAGIC) - MAGIC : frnd(ph * qinv);
  const double pl = fma(y, w, -ph);
  return fma(-Q, q, ph) + pl;
}
template <int MODE> __device__ __forceinline__ void bf(double &x, double &y, double w, double q, double qinv) {
  const double v = mulmod<MODE>(y, w, q, qinv);
  y = x - v;
  x = x + v;
}
__device__ __forceinline__ double reset(double x) { return fma(-floor(x * 5.6843418860808015e-14), 17592186044416.0, x); }

template <int MODE>
__global__ void __launch_bounds__(1024) k_bf(double *out, int iters, double q, double w, double qinv) {
  double x0 = threadIdx.x, y0 = blockIdx.x, x1 = x0 + 1, y1 = y0 + 2, x2 = x0 + 3, y2 = y0 + 4, x3 = x0 + 5, y3 = y0 + 6;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) { bf<MODE>(x0, y0, w, q, qinv); bf<MODE>(x1, y1, w, q, qinv); bf<MODE>(x2, y2, w, q, qinv); bf<MODE>(x3, y3, w, q, qinv); }
    x0 = reset(x0); y0 = reset(y0); x1 = reset(x1); y1 = reset(y1); x2 = reset(x2); y2 = reset(y2); x3 = reset(x3); y3 = reset(y3);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + y0 + x1 + y1 + x2 + y2 + x3 + y3;
}
template <int MODE>
__global__ void __launch_bounds__(1024) k_mul(double *out, int iters, double q, double w, double qinv) {
  double x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = threadIdx.x + 3 * j + blockIdx.x;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = mulmod<MODE>(x[j], w, q, qinv);
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(1024) k_frnd(double *out, int iters, double a) {
  double x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = (threadIdx.x + 3 * j + blockIdx.x) * a;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = frnd(x[j]);
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// FRND next to independent DFMA work: does FRND take FP64-pipe slots?  per iteration 32 DFMA + NF FRND
template <int NF>
__global__ void __launch_bounds__(1024) k_mix(double *out, int iters, double a, double b) {
  double x[8], r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { x[j] = (threadIdx.x + 3 * j + blockIdx.x) * a; r[j] = x[j] * 1.5; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        x[j] = fma(x[j], a, b);
        if (u * 8 + j < NF) r[j] = frnd(r[j] + 0.0 * 0);
      }
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += x[j] + r[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- bulk reduce: every CTA owns 64 KiB of shared memory filled with doubles and adds it `reps` times into
// dst[(cta / 4 + rep) % ring] (4 CTAs share a destination row, like the L ModUp rows of one output modulus)
__global__ void __launch_bounds__(512) k_bulk_red(double *dst, int ring, int reps, double val) {
  extern __shared__ __align__(128) double sm[];
  for (int i = threadIdx.x; i < 8192; i += 512) sm[i] = val;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int r = 0; r < reps; ++r) {
      double *g = dst + (size_t)((blockIdx.x / 4 + r) % ring) * 8192;
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;"
                   ::"l"(g), "r"((unsigned)__cvta_generic_to_shared(sm)), "r"(65536u) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}
// the same traffic as plain bulk stores (what the ModUp rows do today)
__global__ void __launch_bounds__(512) k_bulk_st(double *dst, int ring, int reps, double val) {
  extern __shared__ __align__(128) double sm[];
  for (int i = threadIdx.x; i < 8192; i += 512) sm[i] = val;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int r = 0; r < reps; ++r) {
      double *g = dst + (size_t)((blockIdx.x + r) % ring) * 8192;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                   ::"l"(g), "r"((unsigned)__cvta_generic_to_shared(sm)), "r"(65536u) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}
// coalesced red.global.add.f64, 4 CTAs per destination row
__global__ void __launch_bounds__(512) k_red(double *dst, int ring, int reps, double val) {
  for (int r = 0; r < reps; ++r) {
    double *g = dst + (size_t)((blockIdx.x / 4 + r) % ring) * 8192;
    for (int i = threadIdx.x; i < 8192; i += 512) atomicAdd(g + i, val);
  }
}
__global__ void __launch_bounds__(512) k_l2_read(const double2 *src, size_t n16, int reps, double *out) {
  double s = 0;
  for (int r = 0; r < reps; ++r) {
    const double2 *p = src + ((size_t)blockIdx.x * 4096 + (size_t)r * 4096 * 37) % (n16 - 4096);
#pragma unroll
    for (int i = 0; i < 8; ++i) { const double2 v = __ldcg(p + threadIdx.x + i * 512); s += v.x + v.y; }
  }
  if (s == 12345.678) out[0] = s;
}

template <class F> float timeit(F f, int warm = 1, int reps = 3) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < warm; ++i) f();
  float best = 1e30f;
  for (int i = 0; i < reps; ++i) {
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  printf("{\"sms\": %d, \"clock_khz\": %d", sms, khz);
  const int grid = sms * 2, iters = 512;
  double *out;
  CK(cudaMalloc(&out, (size_t)grid * 1024 * 8));
  const double q = 8796092858369.0, w = 2932030952789.0, qinv = 1.0 / q;
  const double nbf = (double)grid * 1024 * iters * 32;
  float ms;
  ms = timeit([&] { k_bf<0><<<grid, 1024>>>(out, iters, q, w, qinv); });
  printf(", \"bf_magic_G_per_s\": %.1f", nbf / ms / 1e6);
  ms = timeit([&] { k_bf<1><<<grid, 1024>>>(out, iters, q, w, qinv); });
  printf(", \"bf_frnd_G_per_s\": %.1f", nbf / ms / 1e6);
  ms = timeit([&] { k_mul<0><<<grid, 1024>>>(out, iters, q, w, qinv); });
  printf(", \"mul_magic_G_per_s\": %.1f", nbf / ms / 1e6);
  ms = timeit([&] { k_mul<1><<<grid, 1024>>>(out, iters, q, w, qinv); });
  printf(", \"mul_frnd_G_per_s\": %.1f", nbf / ms / 1e6);
  ms = timeit([&] { k_frnd<<<grid, 1024>>>(out, iters, 1.000001); });
  printf(", \"frnd_only_G_per_s\": %.1f", nbf / ms / 1e6);
  ms = timeit([&] { k_mix<0><<<grid, 1024>>>(out, iters, 1.0000001, 0.5); });
  printf(", \"dfma32_frnd0_Gdfma_per_s\": %.1f", nbf / ms / 1e6);
  ms = timeit([&] { k_mix<4><<<grid, 1024>>>(out, iters, 1.0000001, 0.5); });
  printf(", \"dfma32_frnd4_Gdfma_per_s\": %.1f", nbf / ms / 1e6);
  ms = timeit([&] { k_mix<8><<<grid, 1024>>>(out, iters, 1.0000001, 0.5); });
  printf(", \"dfma32_frnd8_Gdfma_per_s\": %.1f", nbf / ms / 1e6);

  // bulk reduce / store / red: ring of 160 rows of 64 KiB (10 MiB: L2-resident), 2 CTAs per SM
  const int ring = 160, reps = 64;
  double *dst;
  CK(cudaMalloc(&dst, (size_t)ring * 65536));
  CK(cudaMemset(dst, 0, (size_t)ring * 65536));
  CK(cudaFuncSetAttribute(k_bulk_red, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(k_bulk_st, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  const double bytes = (double)grid * reps * 65536;
  k_bulk_red<<<grid, 512, 65536>>>(dst, ring, 1, 3.0);
  CK(cudaDeviceSynchronize());
  {
    std::vector<double> h((size_t)ring * 8192);
    CK(cudaMemcpy(h.data(), dst, h.size() * 8, cudaMemcpyDeviceToHost));
    double sum = 0; for (double v : h) sum += v;
    printf(", \"bulk_red_check\": %s", sum == 3.0 * grid * 8192 ? "true" : "false");
  }
  ms = timeit([&] { k_bulk_red<<<grid, 512, 65536>>>(dst, ring, reps, 1.0); });
  printf(", \"bulk_red_GBps\": %.0f", bytes / ms / 1e6);
  ms = timeit([&] { k_bulk_st<<<grid, 512, 65536>>>(dst, ring, reps, 1.0); });
  printf(", \"bulk_store_GBps\": %.0f", bytes / ms / 1e6);
  ms = timeit([&] { k_red<<<grid, 512>>>(dst, ring, reps, 1.0); });
  printf(", \"red_f64_GBps\": %.0f", bytes / ms / 1e6);
  // L2 read: 40 MiB buffer
  const size_t n16 = (size_t)40 << 16;
  double2 *src;
  CK(cudaMalloc(&src, n16 * 16));
  CK(cudaMemset(src, 0, n16 * 16));
  ms = timeit([&] { k_l2_read<<<grid, 512>>>(src, n16, 256, out); }, 2, 3);
  printf(", \"l2_read_GBps\": %.0f", (double)grid * 256 * 65536 / ms / 1e6);
  printf("}\n");
  return 0;
}
