/*
 * pn_linalg.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Small dense linear algebra and deterministic elementary functions for the
 * oracle.  Everything here fixes an OPERATION ORDER that the thread-per-IVP CUDA
 * kernels reproduce (skipping structural zeros only, which is exact), so that
 * GPU results can be compared bit for bit:
 *   - inner products accumulate in ascending index order with explicit fma();
 *   - Householder QR (R only) processes columns left to right; a column whose
 *     sub-diagonal is exactly zero is left untouched (no reflection);
 *   - one sqrt and one reciprocal per Householder column, both IEEE-rounded.
 *
 * What the reference uses instead: jnp.linalg.qr / LAPACK geqrf inside
 * probdiffeq's sqrt utilities (called from every step of
 * ivpsolve.solve_adaptive_save_at, src/odecheckpts/ivpsolvers.py:71-77).  Only
 * R^T R is observable, so any QR variant is admissible.
 */
#include <math.h>
#include <string.h>

#include "pn_internal.h"

/* ------------------------------------------------------------------------- */
/* Householder QR, R only.  M is rows x cols row-major (ld = cols), rows>=1.    */
/* On exit rows 0..min(rows,cols)-1 hold R (upper triangular), everything      */
/* below the diagonal is zero.                                                 */
/* ------------------------------------------------------------------------- */
void pn_qr_r(double *M, int rows, int cols) { pn_qr_r_partial(M, rows, cols, cols); }

/* The same, but only the first `ncols` columns are triangularised (all columns receive the
 * reflectors).  The fixed-point predict only needs R_Y and R_12 of the 2N x 2N block matrix in
 * triangular form; the lower-right block is used as a (non-triangular) square-root factor. */
void pn_qr_r_partial(double *M, int rows, int cols, int ncols) {
  int kmax = rows < cols ? rows : cols;
  if (ncols < kmax) kmax = ncols;
  for (int j = 0; j < kmax; ++j) {
    double sigma2 = 0.0;
    for (int i = j + 1; i < rows; ++i) sigma2 = fma(M[i * cols + j], M[i * cols + j], sigma2);
    if (!(sigma2 > 0.0)) continue; /* already triangular in this column (or NaN) */
    double alpha = M[j * cols + j];
    double norm = sqrt(fma(alpha, alpha, sigma2));
    double v0 = (alpha >= 0.0) ? (alpha + norm) : (alpha - norm);
    double beta = (alpha >= 0.0) ? -norm : norm;
    double g = 1.0 / (norm * (fabs(alpha) + norm)); /* = 2 / (v^T v) */
    for (int c = j + 1; c < cols; ++c) {
      /* w = v^T M[:, c]: the sub-diagonal part first (it does not depend on the reflector's norm,
       * so a GPU can overlap it with the sqrt / reciprocal chain), the v0 term last */
      double w = 0.0;
      for (int i = j + 1; i < rows; ++i) w = fma(M[i * cols + j], M[i * cols + c], w);
      w = fma(v0, M[j * cols + c], w);
      double f = w * g;
      M[j * cols + c] = fma(-f, v0, M[j * cols + c]);
      for (int i = j + 1; i < rows; ++i) M[i * cols + c] = fma(-f, M[i * cols + j], M[i * cols + c]);
    }
    M[j * cols + j] = beta;
    for (int i = j + 1; i < rows; ++i) M[i * cols + j] = 0.0;
  }
}

/* C[r x c] = A[r x k] * B[k x c], ascending-k fma accumulation starting from the k=0 product */
void pn_matmul(const double *A, const double *B, double *C, int r, int k, int c) {
  for (int i = 0; i < r; ++i)
    for (int j = 0; j < c; ++j) {
      double acc = A[i * k] * B[j];
      for (int l = 1; l < k; ++l) acc = fma(A[i * k + l], B[l * c + j], acc);
      C[i * c + j] = acc;
    }
}

/* Solve R X = B for X (R: n x n upper triangular, B: n x c), back substitution.
 * X[i] = (B[i] - sum_{k>i} R[i][k] X[k]) * (1/R[i][i]); k ascending. */
void pn_solve_upper(const double *R, const double *B, double *X, int n, int c) {
  for (int i = n - 1; i >= 0; --i) {
    double inv = 1.0 / R[i * n + i];
    for (int j = 0; j < c; ++j) {
      double acc = B[i * c + j];
      for (int k = i + 1; k < n; ++k) acc = fma(-R[i * n + k], X[k * c + j], acc);
      X[i * c + j] = acc * inv;
    }
  }
}

/* Solve R^T X = B for X (R upper triangular => R^T lower), forward substitution. */
void pn_solve_upper_transposed(const double *R, const double *B, double *X, int n, int c) {
  for (int i = 0; i < n; ++i) {
    double inv = 1.0 / R[i * n + i];
    for (int j = 0; j < c; ++j) {
      double acc = B[i * c + j];
      for (int k = 0; k < i; ++k) acc = fma(-R[k * n + i], X[k * c + j], acc);
      X[i * c + j] = acc * inv;
    }
  }
}

/* ------------------------------------------------------------------------- */
/* Deterministic log / exp / pow: IEEE basic operations only, fixed order.     */
/* Accuracy ~2 ulp, which is what the PI controller needs                      */
/* (control_proportional_integral: (1/e)^(0.3/n) * (e_prev/e)^(0.4/n)).         */
/* ------------------------------------------------------------------------- */
static const double PN_LN2_HI = 6.93147180369123816490e-01; /* 0x3fe62e42fee00000 */
static const double PN_LN2_LO = 1.90821492927058770002e-10; /* 0x3dea39ef35793c76 */
static const double PN_INV_LN2 = 1.44269504088896338700e+00;
static const double PN_SQRT_HALF = 7.07106781186547524401e-01;

double pn_det_log(double x) {
  /* x finite, > 0 */
  int k;
  double m = frexp(x, &k); /* m in [0.5, 1) */
  if (m < PN_SQRT_HALF) {
    m = m * 2.0;
    k -= 1;
  }
  double s = (m - 1.0) * (1.0 / (m + 1.0));
  double z = s * s;
  /* log(m) = 2 s (1 + z/3 + z^2/5 + ... + z^11/23), |s| <= 0.1716 */
  double P = 1.0 / 23.0;
  P = fma(P, z, 1.0 / 21.0);
  P = fma(P, z, 1.0 / 19.0);
  P = fma(P, z, 1.0 / 17.0);
  P = fma(P, z, 1.0 / 15.0);
  P = fma(P, z, 1.0 / 13.0);
  P = fma(P, z, 1.0 / 11.0);
  P = fma(P, z, 1.0 / 9.0);
  P = fma(P, z, 1.0 / 7.0);
  P = fma(P, z, 1.0 / 5.0);
  P = fma(P, z, 1.0 / 3.0);
  P = fma(P, z, 1.0);
  double lm = (2.0 * s) * P;
  double kd = (double)k;
  return fma(kd, PN_LN2_HI, fma(kd, PN_LN2_LO, lm));
}

double pn_det_exp(double y) {
  /* |y| < 700 */
  double kd = floor(fma(y, PN_INV_LN2, 0.5));
  double r = fma(-kd, PN_LN2_HI, y);
  r = fma(-kd, PN_LN2_LO, r);
  /* exp(r), |r| <= 0.3466: Taylor to degree 14 */
  double P = 1.0 / 87178291200.0;
  P = fma(P, r, 1.0 / 6227020800.0);
  P = fma(P, r, 1.0 / 479001600.0);
  P = fma(P, r, 1.0 / 39916800.0);
  P = fma(P, r, 1.0 / 3628800.0);
  P = fma(P, r, 1.0 / 362880.0);
  P = fma(P, r, 1.0 / 40320.0);
  P = fma(P, r, 1.0 / 5040.0);
  P = fma(P, r, 1.0 / 720.0);
  P = fma(P, r, 1.0 / 120.0);
  P = fma(P, r, 1.0 / 24.0);
  P = fma(P, r, 1.0 / 6.0);
  P = fma(P, r, 0.5);
  P = fma(P, r, 1.0);
  P = fma(P, r, 1.0);
  return ldexp(P, (int)kd);
}

double pn_det_pow(double x, double y) {
  if (x != x) return x;
  if (x == 0.0) return 0.0;
  if (x > 1.79769313486231570815e+308) return x; /* +inf */
  if (x < 2.2250738585072014e-308) x = 2.2250738585072014e-308; /* flush subnormals */
  return pn_det_exp(y * pn_det_log(x));
}
