// Microbenchmarks: DFMA issue rate vs operand pattern and dependent-chain latency on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; } } while (0)

template <int MODE, int NACC>
__global__ void __launch_bounds__(128) k(double* out, const double* in, int iters) {
  double x[NACC], y[NACC], z[NACC];
  for (int i = 0; i < NACC; ++i) { x[i] = in[i] + threadIdx.x; y[i] = in[i + 32] * 1e-3; z[i] = in[i + 64]; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < NACC; ++i) {
        if (MODE == 0) x[i] = fma(x[i], y[0], z[0]);            // 2 operands shared by all (reuse cache)
        if (MODE == 1) x[i] = fma(x[i], y[i], z[(i + 1) % NACC]); // 3 distinct registers, no reuse
        if (MODE == 2) x[i] = fma(y[0], y[i], x[i]);              // 1 shared operand (Householder v_i)
        if (MODE == 3) x[i] = fma(y[i], y[i], x[i]);              // squares
        if (MODE == 4) x[i] = x[i] * y[i];                        // DMUL 2 distinct
      }
    }
  }
  double s = 0;
  for (int i = 0; i < NACC; ++i) s += x[i];
  if (s == 1.2345) out[0] = s;
}

template <int MODE, int NACC>
int run(const char* name, double* out, double* in, int warps_per_sm) {
  int iters = 2048;
  int threads = 128, ctas = 148 * (warps_per_sm * 32 / threads);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k<MODE, NACC><<<ctas, threads>>>(out, in, iters);
    cudaEventRecord(e1);
    CHECK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  double inst = (double)iters * 8 * NACC * threads * ctas;   // thread-level DFMA
  double per_smsp_cyc = best * 1e-3 * 1.965e9;               // cycles (at max clock)
  double warp_inst_per_smsp = inst / 32.0 / (148 * 4);
  printf("%-28s NACC=%2d warps/SM=%2d  %.2f TFLOP/s  cycles/warp-inst/SMSP=%.2f\n", name, NACC, warps_per_sm,
         2 * inst / (best * 1e-3) / 1e12, per_smsp_cyc / warp_inst_per_smsp);
  return 0;
}

int main() {
  double *out, *in;
  CHECK(cudaMalloc(&out, 1024)); CHECK(cudaMalloc(&in, 1024));
  double h[128]; for (int i = 0; i < 128; ++i) h[i] = 0.5 + i * 1e-3;
  CHECK(cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice));
  for (int w : {4, 8, 16}) {
    run<0, 16>("shared a,b (reuse)", out, in, w);
    run<1, 16>("3 distinct regs", out, in, w);
    run<2, 16>("1 shared operand", out, in, w);
    run<3, 16>("square accumulate", out, in, w);
    run<4, 16>("DMUL 2 distinct", out, in, w);
  }
  // latency: 1 accumulator, 1 warp per SMSP
  run<1, 1>("dependent chain (latency)", out, in, 4);
  run<1, 2>("2 chains", out, in, 4);
  run<1, 4>("4 chains", out, in, 4);
  run<1, 8>("8 chains", out, in, 4);
  run<1, 4>("4 chains, 2 warps/SMSP", out, in, 8);
  run<1, 2>("2 chains, 2 warps/SMSP", out, in, 8);
  run<1, 1>("1 chain, 2 warps/SMSP", out, in, 8);
  return 0;
}
