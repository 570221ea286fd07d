/*
 * pn_blocked.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Blocked (panel / compact-WY) Householder QR and blocked triangular solve for the dense
 * factorisation with a large state dimension D = (nu+1) d (BASELINE config 5: Brusselator, dense
 * sqrt-EKF1; reference call sites: impl.select("dense") experiments/1_van_der_pol/vdp.py:61,
 * problem src/odecheckpts/ivps.py:124-156, driver experiments/4_brusselator/run.py:51-61).
 *
 * These routines compute the same R factors / solutions as pn_qr_r / pn_solve_upper in pn_linalg.c
 * (any Householder QR is admissible: only R^T R is observable), but in the
 * OPERATION ORDER of the CTA-per-IVP CUDA kernel (csrc/pn_dense_cta_kernel.cuh), whose block
 * products run on the FP64 tensor path (DMMA.8x8x4).  Measured on B200: one DMMA is, per output
 * element, exactly the ascending-k chain fma(a3,b3, fma(a2,b2, fma(a1,b1, fma(a0,b0,c)))) -- so a
 * blocked product is an ascending-k fma chain per element and can be restated here bit for bit.
 * The order that is fixed here and reproduced by the kernel:
 *   - panels of `nb` columns; the panel's ACTIVE rows (the rows that can be non-zero in its columns,
 *     by the structure of the stacked matrix) are gathered in ascending order: gathered index r;
 *   - inside a panel, column jj: ONE pass forms the inner products of column jj (rows r > jj) with
 *     every panel column; each product is summed in "CTA order": thread tau of PN_BLK_THREADS owns
 *     the rows r = tau (mod PN_BLK_THREADS) and chains them with fma in ascending order from 0, the
 *     32 lanes of a warp are then combined by a butterfly (xor 16, 8, 4, 2, 1), and the warps are
 *     added in ascending order; the pivot-row term fma(v0, ., .) comes last;
 *   - T of the compact WY form Q = I - V T V^T by the forward recurrence
 *     T[0:j, j] = -g_j T[0:j, 0:j] (V^T v_j), inner index ascending;
 *   - trailing columns: W = V^T C (ascending gathered row), Y = T^T W (ascending index),
 *     C <- C - V Y (ascending panel column, chain starting from C).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "pn_internal.h"

static double cta_reduce(double *part) {
  /* part[PN_BLK_THREADS]: butterfly inside each group of 32, then the groups in ascending order */
  double nxt[PN_BLK_THREADS];
  for (int off = 16; off >= 1; off >>= 1) {
    for (int l = 0; l < PN_BLK_THREADS; ++l) nxt[l] = part[l] + part[l ^ off];
    memcpy(part, nxt, sizeof(nxt));
  }
  double r = part[0];
  for (int w = 1; w < PN_BLK_THREADS / 32; ++w) r = r + part[w * 32];
  return r;
}

void pn_qr_blocked(double *M, int ld, int rows, int cols, int ncols, int shape, int ntop, int nb) {
  int kmax = rows < cols ? rows : cols;
  if (ncols < kmax) kmax = ncols;
  int hmax = rows + nb;
  double *P = (double *)malloc(sizeof(double) * (size_t)hmax * nb);
  int *rowmap = (int *)malloc(sizeof(int) * (size_t)hmax);
  double *T = (double *)calloc((size_t)nb * nb, sizeof(double));
  double *v0 = (double *)calloc(nb, sizeof(double)), *beta = (double *)calloc(nb, sizeof(double));
  double *gg = (double *)calloc(nb, sizeof(double));
  double *tot = (double *)calloc(nb, sizeof(double)), *S = (double *)calloc(nb, sizeof(double));
  for (int j0 = 0; j0 < kmax; j0 += nb) {
    const int w = (kmax - j0 < nb) ? (kmax - j0) : nb;
    int h = 0;
    if (shape == PN_QR_TOPTRI_BOTFULL) {
      for (int r = 0; r < w; ++r) rowmap[h++] = j0 + r;
      for (int r = ntop; r < rows; ++r) rowmap[h++] = r;
    } else if (shape == PN_QR_TOPFULL_BOTTRI) {
      int end = ntop + j0 + w;
      if (end > rows) end = rows;
      for (int r = j0; r < end; ++r) rowmap[h++] = r;
    } else {
      for (int r = j0; r < rows; ++r) rowmap[h++] = r;
    }
    for (int r = 0; r < h; ++r)
      for (int c = 0; c < nb; ++c) P[r * nb + c] = (c < w) ? M[(size_t)rowmap[r] * ld + j0 + c] : 0.0;
    memset(T, 0, sizeof(double) * (size_t)nb * nb);
    for (int jj = 0; jj < w; ++jj) {
      /* inner products of column jj with every panel column over the rows r > jj, CTA order */
      for (int c = 0; c < w; ++c) {
        double part[PN_BLK_THREADS];
        for (int tau = 0; tau < PN_BLK_THREADS; ++tau) {
          double acc = 0.0;
          for (int r = tau; r < h; r += PN_BLK_THREADS)
            if (r > jj) acc = fma(P[r * nb + jj], P[r * nb + c], acc);
          part[tau] = acc;
        }
        tot[c] = cta_reduce(part);
      }
      const double sigma2 = tot[jj];
      const double alpha = P[jj * nb + jj];
      const int on = sigma2 > 0.0;
      const double norm = sqrt(fma(alpha, alpha, sigma2));
      const double sn = (alpha >= 0.0) ? norm : -norm;
      const double ginv = 1.0 / (norm * (fabs(alpha) + norm));
      const double vv = on ? (alpha + sn) : 0.0, g = on ? ginv : 0.0, bt = on ? -sn : alpha;
      v0[jj] = vv;
      gg[jj] = g;
      beta[jj] = bt;
      for (int c = jj + 1; c < w; ++c) {
        const double wd = fma(vv, P[jj * nb + c], tot[c]);
        const double f = wd * g;
        P[jj * nb + c] = fma(-f, vv, P[jj * nb + c]);
        for (int r = jj + 1; r < h; ++r) P[r * nb + c] = fma(-f, P[r * nb + jj], P[r * nb + c]);
      }
      for (int b = 0; b < jj; ++b) S[b] = fma(vv, P[jj * nb + b], tot[b]);
      for (int i = 0; i < jj; ++i) {
        double acc = 0.0;
        for (int k = i; k < jj; ++k) acc = fma(T[i * nb + k], S[k], acc);
        T[i * nb + jj] = (-g) * acc;
      }
      T[jj * nb + jj] = g;
      P[jj * nb + jj] = vv;
    }
    /* R entries of the panel back to M; V = P with the entries above the pivots cleared */
    for (int r = 0; r < h; ++r)
      for (int c = 0; c < w; ++c) {
        double val = 0.0;
        if (r < w && r < c) val = P[r * nb + c];
        if (r == c) val = beta[c];
        M[(size_t)rowmap[r] * ld + j0 + c] = val;
        if (r < w && r < c) P[r * nb + c] = 0.0;
      }
    /* trailing columns */
#pragma omp parallel for schedule(static) if ((long)h * (cols - j0 - w) > 20000)
    for (int c = j0 + w; c < cols; ++c) {
      double W0[64], Y[64];
      for (int a = 0; a < w; ++a) {
        double acc = 0.0;
        for (int r = 0; r < h; ++r) acc = fma(P[r * nb + a], M[(size_t)rowmap[r] * ld + c], acc);
        W0[a] = acc;
      }
      for (int a = 0; a < w; ++a) {
        double acc = 0.0;
        for (int i = 0; i <= a; ++i) acc = fma(T[i * nb + a], W0[i], acc);
        Y[a] = acc;
      }
      for (int r = 0; r < h; ++r) {
        double acc = M[(size_t)rowmap[r] * ld + c];
        for (int a = 0; a < w; ++a) acc = fma(-P[r * nb + a], Y[a], acc);
        M[(size_t)rowmap[r] * ld + c] = acc;
      }
    }
  }
  free(P); free(rowmap); free(T); free(v0); free(beta); free(gg); free(tot); free(S);
}

/* R X = B (R n x n upper, B n x c), blocked back substitution: block rows of `nb` aligned at
 * multiples of nb, last block first; inside a block row the contributions of the rows BELOW the
 * block come first (ascending k: this is the tensor-core product), then the rows of the block. */
void pn_solve_upper_blocked(const double *R, int ldr, const double *B, int ldb, double *X, int ldx,
                            int n, int c, int nb) {
  int nblk = (n + nb - 1) / nb;
  for (int bi = nblk - 1; bi >= 0; --bi) {
    const int i0 = bi * nb, i1 = (i0 + nb < n) ? i0 + nb : n;
#pragma omp parallel for schedule(static) if ((long)(n - i1) * c > 20000)
    for (int j = 0; j < c; ++j) {
      double acc[64];
      for (int i = i0; i < i1; ++i) {
        double a = B[(size_t)i * ldb + j];
        for (int k = i1; k < n; ++k) a = fma(-R[(size_t)i * ldr + k], X[(size_t)k * ldx + j], a);
        acc[i - i0] = a;
      }
      for (int i = i1 - 1; i >= i0; --i) {
        double a = acc[i - i0];
        for (int k = i + 1; k < i1; ++k) a = fma(-R[(size_t)i * ldr + k], X[(size_t)k * ldx + j], a);
        X[(size_t)i * ldx + j] = a * (1.0 / R[(size_t)i * ldr + i]);
      }
    }
  }
}

/* C = (C0 ? C0 : 0) -/+ A B, row-major, per element the ascending-k fma chain from the initial value:
 * the contract of the kernel's tensor-core products (one DMMA.8x8x4 = four chained fma's per element). */
void pn_gemm_chain(double *Cm, const double *A, const double *B, int M, int N, int K, const double *C0, int neg) {
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      double acc = C0 ? C0[(size_t)i * N + j] : 0.0;
      for (int k = 0; k < K; ++k) acc = fma(neg ? -A[(size_t)i * K + k] : A[(size_t)i * K + k], B[(size_t)k * N + j], acc);
      Cm[(size_t)i * N + j] = acc;
    }
}
