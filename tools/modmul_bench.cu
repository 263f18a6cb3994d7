// Microbenchmark: what does the sm_100a integer path sustain for the 255-bit Montgomery product and the XYZZ mixed addition?
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I taiga_b200/csrc tools/modmul_bench.cu -o tools/modmul_bench
// Prints, per (variant, independent chains per thread, warps per SM): G modmul/s and issue cycles per warp-modmul per SM sub-partition.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define TB_NOINLINE_MUL 0
#include "curve.cuh"
using namespace tb;

template <int ILP, int MODE>
__global__ void __launch_bounds__(128) k_mul(Fq* io, int iters) {
  extern __shared__ uint4 dummy[];
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  Fq x[ILP], y = io[(t * 7 + 1) & 1023];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = io[(t + i * 131) & 1023];
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = Fq::mul_body(x[i], y);
    } else if (MODE == 1) {   // out-of-line pair product
#ifdef __CUDA_ARCH__
#pragma unroll
      for (int i = 0; i < ILP; i += 2) { Fq::Pair p = Fq::mul2_call(x[i], y, x[(i + 1) % ILP], y); x[i] = p.a; x[(i + 1) % ILP] = p.b; }
#endif
    } else {                  // add / sub chain (alu pipe only)
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = (x[i] + y) - x[(i + 1) % ILP];
    }
  }
  Fq acc = x[0];
#pragma unroll
  for (int i = 1; i < ILP; ++i) acc = acc + x[i];
  if (acc.l[0] == 0x12345678u && acc.l[7] == 0x1u) io[t & 1023] = acc;
}

template <int NOINL>
__global__ void __launch_bounds__(128) k_madd(Aff<Fq>* pts, Xyzz<Fq>* out, int iters) {
  extern __shared__ uint4 dummy[];
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  Xyzz<Fq> acc = Xyzz<Fq>::inf();
  for (int it = 0; it < iters; ++it) acc.add_affine(pts[(t * 31 + it) & 1023]);
  if (acc.X.l[0] == 0x12345678u) out[t & 1023] = acc;
}

__global__ void k_inv(Fq* io, int iters, long long* cyc) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  Fq x = io[t & 1023];
  long long c0 = clock64();
  for (int it = 0; it < iters; ++it) { x = x.inv(); x.l[0] ^= (uint32_t)it; x.l[7] &= 0x3fffffffu; }
  long long c1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = (c1 - c0) / iters;
  if (x.l[0] == 0x12345678u && x.l[7] == 0x1u) io[t & 1023] = x;
}

struct Res { double gops; double cyc; };
template <class L> Res timeit(L launch, double ops, int sms, double mhz) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  launch(); cudaDeviceSynchronize();
  cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  Res r; r.gops = ops / (ms * 1e-3) / 1e9;
  r.cyc = (ms * 1e-3 * mhz * 1e6) * sms * 4 / (ops / 32.0);
  return r;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount; int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double mhz = clk / 1000.0;
  printf("device %s, %d SMs, %.0f MHz nominal\n", p.name, sms, mhz);
  Fq* io; cudaMalloc(&io, 1024 * sizeof(Fq));
  { Fq h[1024]; for (int i = 0; i < 1024; ++i) { for (int j = 0; j < 8; ++j) h[i].l[j] = 0x9e3779b9u * (i * 8 + j + 1); h[i].l[7] &= 0x3fffffffu; } cudaMemcpy(io, h, sizeof(h), cudaMemcpyHostToDevice); }
  Aff<Fq>* pts; cudaMalloc(&pts, 1024 * sizeof(Aff<Fq>)); cudaMemcpy(pts, io, 512 * sizeof(Aff<Fq>), cudaMemcpyDeviceToDevice); cudaMemcpy(pts + 512, io, 512 * sizeof(Aff<Fq>), cudaMemcpyDeviceToDevice);
  Xyzz<Fq>* out; cudaMalloc(&out, 1024 * sizeof(Xyzz<Fq>));
  { long long* dc; cudaMalloc(&dc, 8); long long hc = 0;
    for (int threads : {1, 32, 128}) for (int blocks : {1, sms, sms * 4}) {
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      k_inv<<<blocks, threads>>>(io, 20, dc); cudaDeviceSynchronize();
      cudaEventRecord(a); k_inv<<<blocks, threads>>>(io, 20, dc); cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b); cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost);
      printf("inv(): %4d threads x %4d blocks: %lld cycles per inversion (thread 0), %.3f us per inversion step, %.2f M inv/s\n", threads, blocks, hc, ms * 1e3 / 20,
             (double)threads * blocks * 20 / (ms * 1e-3) / 1e6);
    } }
  const int iters = 2000;
  printf("%-22s %4s %9s %12s %10s\n", "variant", "ilp", "warps/SM", "Gmodmul/s", "cyc/wmul/SMSP");
  for (int wps : {4, 8, 12, 16, 24, 32}) {          // warps per SM; 128-thread CTAs, limited through dynamic shared memory
    int ctas = wps / 4; size_t smem = (200 * 1024) / ctas - 2048; if (smem > 200 * 1024) smem = 200 * 1024;
#define RUN(NAME, KERN, ILPV, OPS)                                                                              \
    { cudaFuncSetAttribute(KERN, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);                       \
      int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, KERN, 128, smem);                        \
      Res r = timeit([&] { KERN<<<sms * ctas, 128, smem>>>(io, iters); }, (double)sms * ctas * 128 * iters * (OPS), sms, mhz); \
      printf("%-22s %4d %4d(%d) %12.1f %10.1f\n", NAME, ILPV, wps, occ * 4, r.gops, r.cyc); }
    RUN("mul inline", (k_mul<1, 0>), 1, 1) RUN("mul inline", (k_mul<2, 0>), 2, 2) RUN("mul inline", (k_mul<4, 0>), 4, 4)
    RUN("mul2_call", (k_mul<2, 1>), 2, 2) RUN("mul2_call", (k_mul<4, 1>), 4, 4)
    RUN("add+sub (x2 ops)", (k_mul<2, 2>), 2, 2) RUN("add+sub (x2 ops)", (k_mul<4, 2>), 4, 4)
    { auto K = k_madd<0>; cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, K, 128, smem);
      Res r = timeit([&] { K<<<sms * ctas, 128, smem>>>(pts, out, iters); }, (double)sms * ctas * 128 * iters, sms, mhz);
      printf("%-22s %4d %4d(%d) %12.2f G madd/s %8.1f cyc/wmadd/SMSP\n", "xyzz madd (10 mul)", 1, wps, occ * 4, r.gops, r.cyc); }
  }
  return 0;
}
