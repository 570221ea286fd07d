/* pn_internal.h -- CPU ORACLE internals (test infrastructure, NOT product code). */
#ifndef PN_INTERNAL_H
#define PN_INTERNAL_H

#include "pn_oracle.h"

#define PN_MAX_N 12 /* nu + 1 */

void pn_qr_r(double *M, int rows, int cols);
void pn_qr_r_partial(double *M, int rows, int cols, int ncols);
void pn_matmul(const double *A, const double *B, double *C, int r, int k, int c);
void pn_solve_upper(const double *R, const double *B, double *X, int n, int c);
void pn_solve_upper_transposed(const double *R, const double *B, double *X, int n, int c);

/* blocked variants in the operation order of the CTA-per-IVP CUDA kernel (pn_blocked.c) */
#define PN_BLK_THREADS 256 /* threads of the kernel's CTA: fixes the order of its reductions */
#define PN_TRSM_BLOCK 64   /* block rows of the kernel's blocked back substitution (cta::TRSM_BLOCK) */
enum { PN_QR_FULL = 0, PN_QR_TOPTRI_BOTFULL = 1, PN_QR_TOPFULL_BOTTRI = 2 };
void pn_qr_blocked(double *M, int ld, int rows, int cols, int ncols, int shape, int ntop, int nb);
void pn_solve_upper_blocked(const double *R, int ldr, const double *B, int ldb, double *X, int ldx,
                            int n, int c, int nb);

void pn_gemm_chain(double *Cm, const double *A, const double *B, int M, int N, int K, const double *C0, int neg);

int pn_problem_has_jacobian(int problem);

#endif
