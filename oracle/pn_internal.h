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

int pn_problem_has_jacobian(int problem);

#endif
