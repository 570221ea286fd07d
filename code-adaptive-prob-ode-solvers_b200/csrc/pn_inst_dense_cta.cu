// Instances of the CTA-per-IVP dense kernel (large runtime dimension, blocked QR + DMMA): Brusselator,
// dense EKF0 / EKF1, nu = 4 (BASELINE config 5), fixed-point smoother and filter.
#include "pn_registry.h"
namespace pn {
PN_REGISTER_DENSE_CTA(BrusselatorRt, 4, 1, 16, 1);
PN_REGISTER_DENSE_CTA(BrusselatorRt, 4, 1, 16, 2);
PN_REGISTER_DENSE_CTA(BrusselatorRt, 4, 0, 16, 1);
PN_REGISTER_DENSE_CTA(BrusselatorRt, 2, 1, 16, 1);
}  // namespace pn

// ---- test hooks (not part of the public header): the blocked primitives on caller-provided DEVICE
// matrices, one CTA, so that tests can compare each of them with oracle/pn_blocked.c bit for bit -------
namespace pn {
namespace cta {
template <int NB>
__global__ void __launch_bounds__(T, 1) selftest_qr_kernel(double* M, int ld, int rows, int cols, int ncols, int shape, int ntop) {
  extern __shared__ double smem[];
  QrSmem<NB> qs;
  qs.layout(rows + NB);
  qr_blocked<NB>(M, ld, rows, cols, ncols, shape, ntop, qs);
}
__global__ void __launch_bounds__(T, 1) selftest_gemm_kernel(double* C, int ldc, const double* A, int lda, int a_kmajor, const double* B,
                                                             int ldb, int M, int N, int K, const double* C0, int ldc0, int neg) {
  if (neg)
    gemm<true>(C, ldc, A, lda, a_kmajor != 0, B, ldb, M, N, K, C0, ldc0);
  else
    gemm<false>(C, ldc, A, lda, a_kmajor != 0, B, ldb, M, N, K, C0, ldc0);
}
__global__ void __launch_bounds__(T, 1) selftest_trsm_kernel(const double* R, int ldr, const double* B, int ldb, double* X, int ldx, int n, int c) {
  solve_upper_blocked(R, ldr, B, ldb, X, ldx, n, c);
}
}  // namespace cta
}  // namespace pn

extern "C" {
int pn_b200_selftest_qr(double* M, int ld, int rows, int cols, int ncols, int shape, int ntop) {
  using namespace pn::cta;
  const size_t smem = smem_doubles<16>(rows) * sizeof(double);
  auto kern = selftest_qr_kernel<16>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -4;
  kern<<<1, T, smem>>>(M, ld, rows, cols, ncols, shape, ntop);
  return cudaDeviceSynchronize() == cudaSuccess ? 0 : -4;
}
int pn_b200_selftest_gemm(double* C, int ldc, const double* A, int lda, int a_kmajor, const double* B, int ldb, int M, int N, int K,
                          const double* C0, int ldc0, int neg) {
  using namespace pn::cta;
  selftest_gemm_kernel<<<1, T, GEMM_SMEM * sizeof(double)>>>(C, ldc, A, lda, a_kmajor, B, ldb, M, N, K, C0, ldc0, neg);
  return cudaDeviceSynchronize() == cudaSuccess ? 0 : -4;
}
int pn_b200_selftest_trsm(const double* R, int ldr, const double* B, int ldb, double* X, int ldx, int n, int c) {
  using namespace pn::cta;
  selftest_trsm_kernel<<<1, T, GEMM_SMEM * sizeof(double)>>>(R, ldr, B, ldb, X, ldx, n, c);
  return cudaDeviceSynchronize() == cudaSuccess ? 0 : -4;
}
}
