// dmma_probe.cu -- what the FP64 tensor path (DMMA.8x8x4) does on this GPU:
//   (1) bit pattern: is D = A*B + C the ascending-k fma chain per element?
//   (2) throughput: independent accumulator chains per warp, vs the DFMA-chain peak.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// one warp: C(8x8) = A(8x4) B(4x8) + C0
__global__ void k_exact(const double* A, const double* B, const double* C0, double* C) {
  const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  double a = A[g * 4 + t], b = B[t * 8 + g];
  double c0 = C0[g * 8 + 2 * t], c1 = C0[g * 8 + 2 * t + 1];
  dmma884(c0, c1, a, b);
  C[g * 8 + 2 * t] = c0;
  C[g * 8 + 2 * t + 1] = c1;
}

template <int ACC>
__global__ void __launch_bounds__(256) k_tput(double* out, int iters) {
  double c[ACC][2];
#pragma unroll
  for (int i = 0; i < ACC; ++i) c[i][0] = c[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < ACC; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ACC; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_dfma(double* out, double a, double b, int iters) {
  double x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = (threadIdx.x + i) * 1e-3;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = fma(x[i], a, b);
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  if (s == 123456.789) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA fed from shared memory: warp tile 32x32 (4x4 mma tiles), k chunk 4 per iteration, operands
// re-read from shared memory every iteration (the realistic inner loop of a blocked GEMM)
__global__ void __launch_bounds__(256) k_smem(double* out, int iters) {
  __shared__ double As[64 * 36], Bs[64 * 36];  // [k][m] padded
  for (int i = threadIdx.x; i < 64 * 36; i += 256) { As[i] = 1.0 + i * 1e-9; Bs[i] = 1.0 - i * 1e-9; }
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3, w = threadIdx.x >> 5;
  double c[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j][0] = c[i][j][1] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const int k0 = ((kk + w) & 15) * 4;
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[(k0 + t) * 36 + i * 8 + g];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[(k0 + t) * 36 + j * 8 + g];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(c[i][j][0], c[i][j][1], a[i], b[j]);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += c[i][j][0] + c[i][j][1];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static float time_ms(void (*launch)(), int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

static double* g_out;
static int g_ctas, g_iters = 2048;
template <int ACC> static void l_tput() { k_tput<ACC><<<g_ctas, 256>>>(g_out, g_iters); }
static void l_dfma() { k_dfma<<<g_ctas, 256>>>(g_out, 0.999999, 1e-9, g_iters); }
static void l_smem() { k_smem<<<g_ctas, 256>>>(g_out, g_iters / 16); }

int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  printf("device %s SMs %d\n", prop.name, prop.multiProcessorCount);
  // (1) exactness
  double hA[32], hB[32], hC0[64], hC[64];
  srand(1);
  auto rnd = []() { return (rand() / (double)RAND_MAX - 0.5) * exp2((rand() % 40) - 20); };
  int mism_asc = 0, mism_desc = 0, mism_pair = 0, trials = 2000;
  double *dA, *dB, *dC0, *dC;
  cudaMalloc(&dA, 32 * 8); cudaMalloc(&dB, 32 * 8); cudaMalloc(&dC0, 64 * 8); cudaMalloc(&dC, 64 * 8);
  for (int tr = 0; tr < trials; ++tr) {
    for (int i = 0; i < 32; ++i) { hA[i] = rnd(); hB[i] = rnd(); }
    for (int i = 0; i < 64; ++i) hC0[i] = rnd();
    cudaMemcpy(dA, hA, 32 * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB, 32 * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dC0, hC0, 64 * 8, cudaMemcpyHostToDevice);
    k_exact<<<1, 32>>>(dA, dB, dC0, dC);
    cudaMemcpy(hC, dC, 64 * 8, cudaMemcpyDeviceToHost);
    for (int i = 0; i < 8; ++i)
      for (int j = 0; j < 8; ++j) {
        double asc = hC0[i * 8 + j], desc = hC0[i * 8 + j];
        for (int k = 0; k < 4; ++k) asc = fma(hA[i * 4 + k], hB[k * 8 + j], asc);
        for (int k = 3; k >= 0; --k) desc = fma(hA[i * 4 + k], hB[k * 8 + j], desc);
        double p01 = fma(hA[i * 4 + 1], hB[8 + j], hA[i * 4] * hB[j]);
        double p23 = fma(hA[i * 4 + 3], hB[24 + j], hA[i * 4 + 2] * hB[16 + j]);
        double pair = (p01 + p23) + hC0[i * 8 + j];
        if (asc != hC[i * 8 + j]) ++mism_asc;
        if (desc != hC[i * 8 + j]) ++mism_desc;
        if (pair != hC[i * 8 + j]) ++mism_pair;
      }
  }
  printf("exactness over %d elements: mismatches vs ascending-k fma chain %d, descending %d, pairwise %d\n",
         trials * 64, mism_asc, mism_desc, mism_pair);
  // (2) throughput
  g_ctas = prop.multiProcessorCount * 8;
  cudaMalloc(&g_out, (size_t)g_ctas * 256 * 8);
  {
    float ms = time_ms(l_dfma, 5);
    printf("DFMA chain: %.3f ms  %.2f TFLOP/s\n", ms, 2.0 * 16 * 8 * g_iters * 256.0 * g_ctas / (ms * 1e-3) / 1e12);
  }
  auto rep = [&](const char* name, void (*l)(), int acc) {
    float ms = time_ms(l, 5);
    double flops = 512.0 * 4 * acc * (double)g_iters * 8 * g_ctas;  // 512 flop per warp-level mma
    printf("%s: %.3f ms  %.2f TFLOP/s\n", name, ms, flops / (ms * 1e-3) / 1e12);
  };
  rep("DMMA 1 acc/warp ", l_tput<1>, 1);
  rep("DMMA 2 acc/warp ", l_tput<2>, 2);
  rep("DMMA 4 acc/warp ", l_tput<4>, 4);
  rep("DMMA 8 acc/warp ", l_tput<8>, 8);
  rep("DMMA 16 acc/warp", l_tput<16>, 16);
  for (int ctas_per_sm = 1; ctas_per_sm <= 2; ++ctas_per_sm) {
    g_ctas = prop.multiProcessorCount * ctas_per_sm;
    float ms = time_ms(l_smem, 5);
    double flops = 512.0 * 16 * 16 * (double)(g_iters / 16) * 8 * g_ctas;
    printf("DMMA smem-fed 32x32 warp tile, %d CTA(s)/SM x 8 warps: %.3f ms  %.2f TFLOP/s\n", ctas_per_sm, ms, flops / (ms * 1e-3) / 1e12);
  }
  return 0;
}
