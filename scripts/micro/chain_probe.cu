// Dependent-chain latencies of the fp64 building blocks of the solver step on sm_100a (one warp, clock64):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -I code-adaptive-prob-ode-solvers_b200/csrc
//        scripts/micro/chain_probe.cu -o scripts/micro/chain_probe && scripts/micro/chain_probe
#include <cstdio>
#include "pn_math.cuh"
using namespace pn;
__global__ void probe(double* out, long long* cyc, double x0, double a, double b) {
  double x = x0 + threadIdx.x * 1e-9;
  long long t0, t1;
  // 1. dependent DFMA chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 512; ++i) x = fma(x, a, b);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  // 2. dependent DMUL chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 512; ++i) x = x * a;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[1] = t1 - t0;
  // 3. dependent dsqrt_raw chain
  x = fabs(x) + 2.0;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) x = dsqrt_raw(x) + 2.0;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[2] = t1 - t0;
  // 4. dependent rcp_raw chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) x = rcp_raw(x) + 2.0;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[3] = t1 - t0;
  // 5. make_reflector chain (alpha <- beta)
  double al = x, s2 = 0.5;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) {
    Reflector r = make_reflector(al, s2);
    al = r.g + r.v0;
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[4] = t1 - t0;
  // 6. det_log + det_exp chain
  x = fabs(al) + 1.5;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 32; ++i) x = det_exp(0.3 * det_log(x)) + 1.5;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[5] = t1 - t0;
  // 7. shared-memory round trip (STS -> LDS dependent)
  __shared__ double sm[64];
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) {
    sm[threadIdx.x] = x;
    __syncwarp();
    x = sm[(threadIdx.x + 1) & 31] + 1.0;
    __syncwarp();
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[6] = t1 - t0;
  // 8. warp butterfly sum (5 x shfl + dadd)
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 32; ++i) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) x = x + __shfl_xor_sync(0xffffffffu, x, off);
    x = x * 1e-3;
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[7] = t1 - t0;
  out[threadIdx.x] = x;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 32 * 8); cudaMallocManaged(&cyc, 8 * 8);
  for (int rep = 0; rep < 2; ++rep) { probe<<<1, 32>>>(out, cyc, 1.0, 0.999, 1e-3); cudaDeviceSynchronize(); }
  printf("dependent DFMA            %6.1f cycles\n", cyc[0] / 512.0);
  printf("dependent DMUL            %6.1f cycles\n", cyc[1] / 512.0);
  printf("dsqrt_raw (+1 dadd)       %6.1f cycles\n", cyc[2] / 64.0);
  printf("rcp_raw (+1 dadd)         %6.1f cycles\n", cyc[3] / 64.0);
  printf("make_reflector (+1 dadd)  %6.1f cycles\n", cyc[4] / 64.0);
  printf("det_log + det_exp (+2)    %6.1f cycles\n", cyc[5] / 32.0);
  printf("STS -> syncwarp -> LDS -> syncwarp (+1 dadd) %6.1f cycles\n", cyc[6] / 64.0);
  printf("warp butterfly sum of a double (+1 dmul)     %6.1f cycles\n", cyc[7] / 32.0);
  return 0;
}
