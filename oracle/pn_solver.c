/*
 * pn_solver.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * The adaptive probabilistic IVP solver loop of probdiffeq as driven by
 * src/odecheckpts/ivpsolvers.py:14-91 (solve) and the experiment scripts
 * (experiments/1_van_der_pol/vdp.py:61-91, experiments/4_brusselator/run.py:51-61,
 * 82-90,119-129, experiments/5_vs_interpolation/measure.py:44-68).  Restated from
 * SURVEY.md Appendix A:
 *   A.1 prior constants            -> pn_oracle_prior
 *   A.3 one attempted step         -> attempt_step
 *   A.4 conditional algebra        -> marginalise, merge
 *   A.5 checkpoints                -> pn_oracle_solve_save_at
 *   A.6 final smoothing sweep      -> pn_oracle_solve_save_at (tail)
 *   A.7 other outer loops          -> pn_oracle_solve_save_every_step / _fixed_grid
 *
 * One engine serves every factorisation: F factor sets, each with an N x N
 * square-root factor, owning C mean columns and observing r residual rows:
 *   isotropic : F=1, N=n,   C=d, r=1 (h = e_q shared by all columns)
 *   blockdiag : F=d, N=n,   C=1, r=1
 *   dense d=1 : F=1, N=n,   C=1, r=1 (h general: EKF1)
 *   dense d>1 : F=1, N=n*d, C=1, r=d (H general d x D, derivative-major index i*d+j)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "pn_internal.h"

/* ------------------------------------------------------------------------- */
/* prior constants (App. A.1)                                                 */
/* ------------------------------------------------------------------------- */
void pn_oracle_prior(int nu, double *a1, double *lq) {
  int n = nu + 1;
  /* flipped Pascal: A1[i][j] = binom(nu-i, nu-j), j >= i */
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double v = 0.0;
      if (j >= i) {
        int a = nu - i, b = nu - j; /* binom(a, b) */
        long double c = 1.0L;
        for (int k = 1; k <= b; ++k) c = c * (long double)(a - b + k) / (long double)k;
        v = (double)c;
      }
      a1[i * n + j] = v;
    }
  /* flipped Hilbert Q1[i][j] = 1/(2 nu - i - j + 1); Cholesky in extended precision, rounded once */
  long double Q[PN_MAX_N * PN_MAX_N], Lc[PN_MAX_N * PN_MAX_N];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      Q[i * n + j] = 1.0L / (long double)(2 * nu - i - j + 1);
      Lc[i * n + j] = 0.0L;
    }
  for (int j = 0; j < n; ++j) {
    long double s = Q[j * n + j];
    for (int k = 0; k < j; ++k) s -= Lc[j * n + k] * Lc[j * n + k];
    long double djj = sqrtl(s);
    Lc[j * n + j] = djj;
    for (int i = j + 1; i < n; ++i) {
      long double t = Q[i * n + j];
      for (int k = 0; k < j; ++k) t -= Lc[i * n + k] * Lc[j * n + k];
      Lc[i * n + j] = t / djj;
    }
  }
  for (int i = 0; i < n * n; ++i) lq[i] = (double)Lc[i];
}

/* ------------------------------------------------------------------------- */
/* engine                                                                     */
/* ------------------------------------------------------------------------- */
typedef struct {
  pn_oracle_config cfg;
  int n, d, q, nu;
  int F, N, C, Ctot, r;
  int dense;         /* r > 1 path */
  int blk;           /* > 0: blocked QR / back substitution with panels of blk columns (dense only) */
  double *A, *LQ;    /* N x N (Kronecker-expanded for dense d>1) */
  double fact[PN_MAX_N], invfact[PN_MAX_N];
  double inv_sqrt_d, inv_sqrt_C;
  /* workspaces */
  double *p, *pinv;              /* N */
  double *m_p, *m_ext_p, *m_ext; /* N*Ctot */
  double *z, *err, *fbuf, *ubuf; /* d, d, d, q*d */
  double *h;                     /* r*N (kron: N; dense: d*N) */
  double *jac;                   /* d*q*d */
  double *L_p, *AL, *M2, *Gp, *Lamp, *X, *T, *Mq; /* N*N (M2: 2N*2N) */
  double *L_ext;                 /* F*N*N */
  double *gain;                  /* N*r */
  double *hL, *W, *Y, *Rm, *Rs;  /* r*N, N*r, ... */
  double *sig;                   /* F */
  double zz0, invS0, mle_x2;     /* solver_mle: sum of squared residuals, 1/S of the last attempt; z^T S^-1 z / d */
  double *keep1, *keep2, *gnew;  /* N*Ctot scratch */
  double *idG, *idL, *idg;       /* identity conditional (I, 0, 0) */
} engine;

typedef struct {
  double t;
  double *mean;  /* N*Ctot */
  double *chol;  /* F*N*N */
  double *G, *g, *Lam; /* backward conditional */
  double *sigma; /* F : output scale carried by the state */
} pstate;

static double *dalloc(size_t k) { return (double *)calloc(k > 0 ? k : 1, sizeof(double)); }

static int engine_init(engine *E, const pn_oracle_config *cfg) {
  memset(E, 0, sizeof(*E));
  E->cfg = *cfg;
  E->nu = cfg->nu;
  E->n = cfg->nu + 1;
  E->d = cfg->d;
  E->q = cfg->ode_order;
  if (E->n > PN_MAX_N || E->q < 1 || E->q > cfg->nu || E->d < 1) return -1;
  int n = E->n, d = E->d;
  if (cfg->factorisation == PN_FACT_ISOTROPIC) {
    if (cfg->correction != PN_CORR_TS0) return -2; /* EKF1 breaks the Kronecker structure */
    E->F = 1; E->N = n; E->C = d; E->Ctot = d; E->r = 1;
  } else if (cfg->factorisation == PN_FACT_BLOCKDIAG) {
    if (cfg->correction != PN_CORR_TS0) return -2;
    E->F = d; E->N = n; E->C = 1; E->Ctot = d; E->r = 1;
  } else if (cfg->factorisation == PN_FACT_DENSE) {
    E->F = 1; E->N = n * d; E->C = 1; E->Ctot = 1; E->r = d;
  } else {
    return -3;
  }
  if (cfg->correction == PN_CORR_TS1 && !pn_problem_has_jacobian(cfg->problem)) return -4;
  /* solver_mle: one factor set with a scalar innovation variance (isotropic, dense with d = 1) */
  if (cfg->calibration == PN_CALIB_MLE && (E->F != 1 || E->r != 1)) return -6;
  E->dense = (E->r > 1);
  E->blk = (E->dense && cfg->dense_block > 0) ? cfg->dense_block : 0;
  if (E->blk > 64) return -5;
  int N = E->N;
  double a1[PN_MAX_N * PN_MAX_N], lq[PN_MAX_N * PN_MAX_N];
  pn_oracle_prior(cfg->nu, a1, lq);
  E->A = dalloc((size_t)N * N);
  E->LQ = dalloc((size_t)N * N);
  if (!E->dense) {
    memcpy(E->A, a1, sizeof(double) * n * n);
    memcpy(E->LQ, lq, sizeof(double) * n * n);
  } else {
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j)
        for (int l = 0; l < d; ++l) {
          E->A[(i * d + l) * N + (j * d + l)] = a1[i * n + j];
          E->LQ[(i * d + l) * N + (j * d + l)] = lq[i * n + j];
        }
  }
  double f = 1.0;
  for (int k = 0; k < n; ++k) {
    if (k > 0) f *= (double)k;
    E->fact[k] = f;
    E->invfact[k] = 1.0 / f;
  }
  E->inv_sqrt_d = 1.0 / sqrt((double)d);
  E->inv_sqrt_C = 1.0 / sqrt((double)E->C);
  size_t NN = (size_t)N * N, NC = (size_t)N * E->Ctot;
  E->p = dalloc(N); E->pinv = dalloc(N);
  E->m_p = dalloc(NC); E->m_ext_p = dalloc(NC); E->m_ext = dalloc(NC);
  E->z = dalloc(d); E->err = dalloc(d); E->fbuf = dalloc(d); E->ubuf = dalloc((size_t)E->q * d);
  E->h = dalloc((size_t)E->r * N);
  E->jac = dalloc((size_t)d * E->q * d);
  E->L_p = dalloc(NN); E->AL = dalloc(NN); E->M2 = dalloc(4 * NN); E->Gp = dalloc(NN);
  E->Lamp = dalloc(NN); E->X = dalloc(NN); E->T = dalloc(NN); E->Mq = dalloc(NN);
  E->L_ext = dalloc((size_t)E->F * NN);
  E->gain = dalloc((size_t)N * E->r);
  E->hL = dalloc((size_t)E->r * N); E->W = dalloc((size_t)N * E->r); E->Y = dalloc((size_t)N * E->r);
  E->Rm = dalloc((size_t)N * E->r); E->Rs = dalloc((size_t)N * E->r);
  E->sig = dalloc(E->F);
  E->keep1 = dalloc(NC); E->keep2 = dalloc(NC); E->gnew = dalloc(NC);
  E->idG = dalloc(NN); E->idL = dalloc(NN); E->idg = dalloc(NC);
  for (int i = 0; i < N; ++i) E->idG[(size_t)i * N + i] = 1.0;
  return 0;
}

static void engine_free(engine *E) {
  free(E->A); free(E->LQ); free(E->p); free(E->pinv); free(E->m_p); free(E->m_ext_p);
  free(E->m_ext); free(E->z); free(E->err); free(E->fbuf); free(E->ubuf); free(E->h); free(E->jac);
  free(E->L_p); free(E->AL); free(E->M2); free(E->Gp); free(E->Lamp); free(E->X); free(E->T);
  free(E->Mq); free(E->L_ext); free(E->gain); free(E->hL); free(E->W); free(E->Y); free(E->Rm);
  free(E->Rs); free(E->sig); free(E->keep1); free(E->keep2); free(E->gnew); free(E->idG); free(E->idL); free(E->idg);
}

static void pstate_alloc(const engine *E, pstate *S) {
  size_t NN = (size_t)E->N * E->N, NC = (size_t)E->N * E->Ctot;
  S->t = 0.0;
  S->mean = dalloc(NC);
  S->chol = dalloc((size_t)E->F * NN);
  S->G = dalloc((size_t)E->F * NN);
  S->g = dalloc(NC);
  S->Lam = dalloc((size_t)E->F * NN);
  S->sigma = dalloc(E->F);
}
static void pstate_free(pstate *S) {
  free(S->mean); free(S->chol); free(S->G); free(S->g); free(S->Lam); free(S->sigma);
}
static void pstate_copy(const engine *E, pstate *dst, const pstate *src) {
  size_t NN = (size_t)E->N * E->N, NC = (size_t)E->N * E->Ctot;
  dst->t = src->t;
  memcpy(dst->mean, src->mean, sizeof(double) * NC);
  memcpy(dst->chol, src->chol, sizeof(double) * E->F * NN);
  memcpy(dst->G, src->G, sizeof(double) * E->F * NN);
  memcpy(dst->g, src->g, sizeof(double) * NC);
  memcpy(dst->Lam, src->Lam, sizeof(double) * E->F * NN);
  memcpy(dst->sigma, src->sigma, sizeof(double) * E->F);
}
/* backward conditional <- identity (I, 0, 0)  (App. A.2) */
static void bw_identity(const engine *E, pstate *S) {
  size_t NN = (size_t)E->N * E->N, NC = (size_t)E->N * E->Ctot;
  memset(S->G, 0, sizeof(double) * E->F * NN);
  memset(S->g, 0, sizeof(double) * NC);
  memset(S->Lam, 0, sizeof(double) * E->F * NN);
  for (int f = 0; f < E->F; ++f)
    for (int i = 0; i < E->N; ++i) S->G[f * NN + (size_t)i * E->N + i] = 1.0;
}

/* p_i(dt) = |dt|^(nu-i+1/2)/(nu-i)!,  pinv_i = 1/p_i  (App. A.1) */
static void precondition(const engine *E, double dt, double *p, double *pinv) {
  double adt = fabs(dt);
  double sq = sqrt(adt);
  double isq = 1.0 / sq, idt = 1.0 / adt;
  double dtp = 1.0, idtp = 1.0;
  double pn[PN_MAX_N], pinvn[PN_MAX_N];
  for (int k = 0; k <= E->nu; ++k) {
    int i = E->nu - k;
    pn[i] = (sq * dtp) * E->invfact[k];
    pinvn[i] = (isq * idtp) * E->fact[k];
    dtp *= adt;
    idtp *= idt;
  }
  if (!E->dense) {
    for (int i = 0; i < E->n; ++i) { p[i] = pn[i]; pinv[i] = pinvn[i]; }
  } else {
    for (int i = 0; i < E->n; ++i)
      for (int l = 0; l < E->d; ++l) { p[i * E->d + l] = pn[i]; pinv[i * E->d + l] = pinvn[i]; }
  }
}

/* sum of squares of v[0..k) in the configured order (see pn_oracle_config.reduction_group) */
static double sum_squares(const engine *E, const double *v, int k) {
  int G = E->cfg.reduction_group;
  if (G <= 1) {
    double acc = 0.0;
    for (int c = 0; c < k; ++c) acc = fma(v[c], v[c], acc);
    return acc;
  }
  /* lane l accumulates the entries l, l+G, l+2G, ... in order (fma chain from 0); lanes are then
   * summed by a butterfly inside each group of 32 and the group sums are added in order */
  double lanes[4096], next[4096];
  for (int l = 0; l < G; ++l) {
    double acc = 0.0;
    for (int c = l; c < k; c += G) acc = fma(v[c], v[c], acc);
    lanes[l] = acc;
  }
  int W = G < 32 ? G : 32;
  for (int off = W / 2; off >= 1; off >>= 1) {
    for (int l = 0; l < G; ++l) next[l] = lanes[l] + lanes[l ^ off];
    for (int l = 0; l < G; ++l) lanes[l] = next[l];
  }
  double r = lanes[0];
  for (int w = 1; w < G / W; ++w) r = r + lanes[w * W];
  return r;
}

/* QR (R only) of a rows x cols matrix with leading dimension cols; only the first ncols columns are
 * triangularised.  `shape` / `ntop` name the structural zeros of the stacked matrix: the unblocked
 * routine ignores them (zeros are exact no-ops in its full loops), the blocked one uses them to pick
 * the active rows of a panel. */
static void eng_qr(const engine *E, double *M, int rows, int cols, int ncols, int shape, int ntop) {
  if (E->blk > 0)
    pn_qr_blocked(M, cols, rows, cols, ncols, shape, ntop, E->blk);
  else
    pn_qr_r_partial(M, rows, cols, ncols);
}

/* predicted mean: m_p = pinv*m, m_ext_p = A m_p, m_ext = p*m_ext_p */
static void predict_mean(engine *E, const double *mean) {
  int N = E->N, Ct = E->Ctot;
  for (int i = 0; i < N; ++i)
    for (int c = 0; c < Ct; ++c) E->m_p[i * Ct + c] = E->pinv[i] * mean[i * Ct + c];
  pn_matmul(E->A, E->m_p, E->m_ext_p, N, N, Ct);
  for (int i = 0; i < N; ++i)
    for (int c = 0; c < Ct; ++c) E->m_ext[i * Ct + c] = E->p[i] * E->m_ext_p[i * Ct + c];
}

/* A.4 merge: run=(G1,g1,Lam1) older, new=(G2,g2,Lam2); result overwrites (Go,go,Lo).
 * Operates on ONE factor set; g has leading dimension Ct, columns [c0, c0+C). */
static void merge_one(engine *E, const double *G1, const double *g1, const double *L1,
                      const double *G2, const double *g2, const double *L2, double *Go, double *go,
                      double *Lo, int c0, int C) {
  int N = E->N, Ct = E->Ctot;
  double *Gtmp = E->X, *T = E->T, *M = E->M2;
  pn_matmul(G1, G2, Gtmp, N, N, N);
  /* g = G1 g2 + g1 */
  for (int i = 0; i < N; ++i)
    for (int c = c0; c < c0 + C; ++c) {
      double acc = g1[i * Ct + c];
      for (int k = 0; k < N; ++k) acc = fma(G1[i * N + k], g2[k * Ct + c], acc);
      E->m_p[i * Ct + c] = acc; /* m_p is free at this point (callers consumed it) */
    }
  pn_matmul(G1, L2, T, N, N, N);
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      M[i * N + j] = T[j * N + i];
      M[(N + i) * N + j] = L1[j * N + i];
    }
  eng_qr(E, M, 2 * N, N, N, PN_QR_TOPFULL_BOTTRI, N);
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) Lo[i * N + j] = (j <= i) ? M[j * N + i] : 0.0;
  memcpy(Go, Gtmp, sizeof(double) * N * N);
  for (int i = 0; i < N; ++i)
    for (int c = c0; c < c0 + C; ++c) go[i * Ct + c] = E->m_p[i * Ct + c];
}

/* A.4 marginalise one factor set: (m,L) through (G,g,Lam) -> (mo, Lo) */
static void marginalise_one(engine *E, const double *m, const double *L, const double *G,
                            const double *g, const double *Lam, double *mo, double *Lo, int c0,
                            int C) {
  int N = E->N, Ct = E->Ctot;
  double *T = E->T, *M = E->M2;
  double *tmp = E->m_ext_p; /* scratch N*Ct */
  for (int i = 0; i < N; ++i)
    for (int c = c0; c < c0 + C; ++c) {
      double acc = g[i * Ct + c];
      for (int k = 0; k < N; ++k) acc = fma(G[i * N + k], m[k * Ct + c], acc);
      tmp[i * Ct + c] = acc;
    }
  pn_matmul(G, L, T, N, N, N);
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      M[i * N + j] = T[j * N + i];
      M[(N + i) * N + j] = Lam[j * N + i];
    }
  eng_qr(E, M, 2 * N, N, N, PN_QR_TOPFULL_BOTTRI, N);
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) Lo[i * N + j] = (j <= i) ? M[j * N + i] : 0.0;
  for (int i = 0; i < N; ++i)
    for (int c = c0; c < c0 + C; ++c) mo[i * Ct + c] = tmp[i * Ct + c];
}

/* Covariance prediction for one factor set (App. A.3 "predict covariance").
 * Requires E->p/pinv, E->m_p, E->m_ext_p for the set's columns.
 * L: N x N factor at the start; sigma: process-noise scale.
 * Writes L_ext (N x N).  If with_bw: forms the new conditional and merges it with the running
 * one (runG==NULL means the running conditional is the identity), writing (Go, go, Lo). */
static void predict_cov_one(engine *E, const double *L, double sigma, double *L_ext, int with_bw,
                            const double *runG, const double *rung, const double *runL, double *Go,
                            double *go, double *Lo, int c0, int C) {
  int N = E->N, Ct = E->Ctot;
  double *L_p = E->L_p, *AL = E->AL, *M = E->M2;
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) L_p[i * N + j] = E->pinv[i] * L[i * N + j];
  pn_matmul(E->A, L_p, AL, N, N, N);
  if (!with_bw) {
    /* filter: R of [ (sigma LQ)^T ; (A L_p)^T ]  (2N x N) */
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < N; ++j) {
        M[i * N + j] = sigma * E->LQ[j * N + i];
        M[(N + i) * N + j] = AL[j * N + i];
      }
    eng_qr(E, M, 2 * N, N, N, PN_QR_TOPTRI_BOTFULL, N);
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < N; ++j) L_ext[i * N + j] = (j <= i) ? E->p[i] * M[j * N + i] : 0.0;
    return;
  }
  /* fixed-point: R of [[ (sigma LQ)^T, 0 ], [ (A L_p)^T, L_p^T ]]  (2N x 2N) */
  int W2 = 2 * N;
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      M[i * W2 + j] = sigma * E->LQ[j * N + i];
      M[i * W2 + N + j] = 0.0;
      M[(N + i) * W2 + j] = AL[j * N + i];
      M[(N + i) * W2 + N + j] = L_p[j * N + i];
    }
  /* only the first N columns are triangularised: R_Y, R_12 are final after that, and the
   * lower-right block B (N x N, full) satisfies B^T B = R_XY^T R_XY, i.e. B^T is a valid
   * (non-triangular) square-root factor of the backward noise; the merge below re-triangularises */
  eng_qr(E, M, W2, W2, N, PN_QR_TOPTRI_BOTFULL, N);
  double *RY = E->Mq, *R12 = E->Gp, *X = E->X;
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      RY[i * N + j] = M[i * W2 + j];
      R12[i * N + j] = M[i * W2 + N + j];
    }
  if (E->blk > 0)
    pn_solve_upper_blocked(RY, N, R12, N, X, N, N, N, PN_TRSM_BLOCK);
  else
    pn_solve_upper(RY, R12, X, N, N); /* X = RY^{-1} R12 ; G_p = X^T */
  /* g_p = m_p - G_p m_ext_p ; un-precondition */
  double *Gn = E->AL; /* AL no longer needed */
  double *Ln = E->Lamp;
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      Gn[i * N + j] = (E->p[i] * X[j * N + i]) * E->pinv[j];
      Ln[i * N + j] = E->p[i] * M[(N + j) * W2 + N + i];
    }
  double *gnew = E->gnew;
  for (int i = 0; i < N; ++i)
    for (int c = c0; c < c0 + C; ++c) {
      double acc = E->m_p[i * Ct + c];
      for (int k = 0; k < N; ++k) acc = fma(-X[k * N + i], E->m_ext_p[k * Ct + c], acc);
      gnew[i * Ct + c] = E->p[i] * acc;
    }
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) L_ext[i * N + j] = (j <= i) ? E->p[i] * M[j * W2 + i] : 0.0;
  if (runG == NULL) {
    /* running conditional = identity: still merge, so that Lam comes out triangular */
    double *Gi = E->idG, *Li = E->idL, *gi = E->idg;
    merge_one(E, Gi, gi, Li, Gn, gnew, Ln, Go, go, Lo, c0, C);
  } else {
    /* merge_one uses X, T, M2, m_p as scratch: Gn lives in AL, Ln in Lamp, gnew in its own buffer;
     * m_p is overwritten -> callers restore it per factor set. */
    merge_one(E, runG, rung, runL, Gn, gnew, Ln, Go, go, Lo, c0, C);
  }
}

/* Build the linearisation at the predicted mean: fills E->z (residual), E->h. */
static void linearise(engine *E, double t, const double *params) {
  int n = E->n, d = E->d, q = E->q, N = E->N;
  /* u, u', ... at the predicted mean: rows 0..q-1 of m_ext (flat index i*d + l in every layout) */
  for (int k = 0; k < q; ++k)
    for (int l = 0; l < d; ++l) E->ubuf[k * d + l] = E->m_ext[k * d + l];
  pn_oracle_vf(E->cfg.problem, d, E->ubuf, t, params, E->fbuf);
  for (int l = 0; l < d; ++l) E->z[l] = E->m_ext[q * d + l] - E->fbuf[l];
  if (!E->dense) {
    for (int i = 0; i < n; ++i) E->h[i] = 0.0;
    E->h[q] = 1.0;
    if (E->cfg.correction == PN_CORR_TS1) { /* d == 1 */
      pn_oracle_jac(E->cfg.problem, d, E->ubuf, t, params, E->jac);
      for (int k = 0; k < q; ++k) E->h[k] = -E->jac[k];
    }
  } else {
    memset(E->h, 0, sizeof(double) * (size_t)d * N);
    for (int l = 0; l < d; ++l) E->h[l * N + q * d + l] = 1.0;
    if (E->cfg.correction == PN_CORR_TS1) {
      pn_oracle_jac(E->cfg.problem, d, E->ubuf, t, params, E->jac);
      for (int l = 0; l < d; ++l)
        for (int k = 0; k < q; ++k)
          for (int l2 = 0; l2 < d; ++l2) E->h[l * N + k * d + l2] = -E->jac[l * (q * d) + k * d + l2];
    }
  }
}

/* Local calibration + error estimate from the process noise only (App. A.3).
 * Fills E->sig[f] = sigma_hat_f and E->err[c] (unscaled error per ODE dimension). */
static void calibrate_and_estimate(engine *E, double dt) {
  int N = E->N, d = E->d;
  double adt = fabs(dt);
  if (!E->dense) {
    /* w_j = sum_i (h_i p_i) LQ[i][j];  s = ||w|| */
    double s2 = 0.0;
    for (int j = 0; j < N; ++j) {
      double acc = 0.0;
      for (int i = 0; i < N; ++i) acc = fma(E->h[i] * E->p[i], E->LQ[i * N + j], acc);
      s2 = fma(acc, acc, s2);
    }
    double s = sqrt(s2);
    for (int f = 0; f < E->F; ++f) {
      int c0 = f * E->C;
      double zz = sum_squares(E, E->z + c0, E->C);
      if (f == 0) E->zz0 = zz;
      double sigma_hat = (sqrt(zz) * (1.0 / s)) * E->inv_sqrt_C;
      E->sig[f] = sigma_hat;
      double er = (adt * sigma_hat) * s;
      for (int c = c0; c < c0 + E->C; ++c) E->err[c] = er;
    }
  } else {
    /* S_Q = (H p LQ)(H p LQ)^T via QR of the D x d transposed product */
    double *Rs = E->Rs; /* N x d */
    for (int j = 0; j < N; ++j)
      for (int l = 0; l < d; ++l) {
        double acc = 0.0;
        for (int i = 0; i < N; ++i) acc = fma(E->h[l * N + i] * E->p[i], E->LQ[i * N + j], acc);
        Rs[j * d + l] = acc;
      }
    /* rows beyond block q are exactly zero (LQ is lower triangular, H has no entries beyond block q) */
    eng_qr(E, Rs, (E->q + 1) * d, d, d, PN_QR_FULL, 0);
    /* y = R^{-T} z */
    double *y = E->fbuf;
    pn_solve_upper_transposed(Rs, E->z, y, d, 1);
    double yy = 0.0;
    for (int l = 0; l < d; ++l) yy = fma(y[l], y[l], yy);
    double sigma_hat = sqrt(yy) * E->inv_sqrt_d;
    E->sig[0] = sigma_hat;
    for (int l = 0; l < d; ++l) {
      double cc = 0.0;
      for (int i = 0; i <= l; ++i) cc = fma(Rs[i * d + l], Rs[i * d + l], cc);
      E->err[l] = (adt * sigma_hat) * sqrt(cc);
    }
  }
}

/* Correction (App. A.3 "correct"): L_ext (one set) -> L_new; gain in E->gain. */
static void correct_cov_one(engine *E, const double *L_ext, double *L_new) {
  int N = E->N, r = E->r;
  double *M = E->M2;
  if (!E->dense) {
    double *hL = E->hL;
    double S = 0.0;
    for (int j = 0; j < N; ++j) {
      double acc = 0.0;
      for (int i = 0; i < N; ++i) acc = fma(E->h[i], L_ext[i * N + j], acc);
      hL[j] = acc;
      S = fma(acc, acc, S);
    }
    double invS = 1.0 / S;
    E->invS0 = invS;
    for (int i = 0; i < N; ++i) {
      double acc = 0.0;
      for (int j = 0; j < N; ++j) acc = fma(L_ext[i * N + j], hL[j], acc);
      E->gain[i] = acc * invS;
    }
    for (int j = 0; j < N; ++j)
      for (int i = 0; i < N; ++i) M[j * N + i] = fma(-hL[j], E->gain[i], L_ext[i * N + j]);
  } else {
    double *HL = E->hL; /* r x N */
    pn_matmul(E->h, L_ext, HL, r, N, N);
    double *Rm = E->Rm; /* N x r */
    for (int j = 0; j < N; ++j)
      for (int l = 0; l < r; ++l) Rm[j * r + l] = HL[l * N + j];
    eng_qr(E, Rm, N, r, r, PN_QR_FULL, 0); /* R_marg in the top r x r */
    /* W = L_ext HL^T (N x r);  gain^T = (R^T R)^{-1} W^T */
    double *Wt = E->W; /* r x N : W^T */
    for (int l = 0; l < r; ++l)
      for (int i = 0; i < N; ++i) {
        double acc = 0.0;
        for (int j = 0; j < N; ++j) acc = fma(L_ext[i * N + j], HL[l * N + j], acc);
        Wt[l * N + i] = acc;
      }
    double *Y = E->Y; /* r x N */
    pn_solve_upper_transposed(Rm, Wt, Y, r, N);
    double *Gt = E->gain; /* r x N : gain^T */
    pn_solve_upper(Rm, Y, Gt, r, N);
    for (int j = 0; j < N; ++j)
      for (int i = 0; i < N; ++i) {
        double acc = L_ext[i * N + j];
        for (int l = 0; l < r; ++l) acc = fma(-HL[l * N + j], Gt[l * N + i], acc);
        M[j * N + i] = acc;
      }
  }
  eng_qr(E, M, N, N, N, PN_QR_FULL, 0);
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) L_new[i * N + j] = (j <= i) ? M[j * N + i] : 0.0;
}

static void correct_mean_one(engine *E, double *m_new, int c0, int C) {
  int N = E->N, Ct = E->Ctot, r = E->r;
  if (!E->dense) {
    for (int i = 0; i < N; ++i)
      for (int c = c0; c < c0 + C; ++c) m_new[i * Ct + c] = fma(-E->gain[i], E->z[c], E->m_ext[i * Ct + c]);
  } else {
    for (int i = 0; i < N; ++i) {
      double acc = E->m_ext[i];
      for (int l = 0; l < r; ++l) acc = fma(-E->gain[l * N + i], E->z[l], acc);
      m_new[i] = acc;
    }
  }
}

typedef struct {
  double e_prev; /* error norm of the previously accepted step */
} pi_state;

/* control_proportional_integral().apply  (App. A.3 last lines; ivpsolvers.py:52) */
/* ln of an error norm as the controller uses it: subnormals are flushed, 0 maps to ln(DBL_TRUE_MIN) */
static double pi_log(double e) {
  if (e == 0.0) return -745.0;
  return pn_det_log(e < 2.2250738585072014e-308 ? 2.2250738585072014e-308 : e);
}

static double pi_factor(const engine *E, double e, double e_prev) {
  /* safety (1/e)^n1 (e_prev/e)^n2 = safety exp(n2 ln e_prev - (n1 + n2) ln e): one log and one exp per
   * attempt (a GPU lane keeps ln e_prev from the step that was accepted) */
  double nn = (double)(E->nu + 1);
  double n1 = E->cfg.power_integral / nn, n2 = E->cfg.power_proportional / nn;
  double fac;
  if (e != e) {
    fac = e;
  } else if (e == 0.0) {
    fac = E->cfg.factor_max;
  } else {
    fac = E->cfg.safety * pn_det_exp(fma(n2, pi_log(e_prev), -((n1 + n2) * pi_log(e))));
  }
  fac = (fac < E->cfg.factor_max) ? fac : E->cfg.factor_max;
  fac = (fac > E->cfg.factor_min) ? fac : E->cfg.factor_min;
  return fac;
}

/* One attempted step from S with step dt (App. A.3).  Writes the proposal into P.
 * atol/rtol passed explicitly (per-member tolerances in the ensemble entry). */
static void attempt_step(engine *E, const double *params, const pstate *S, double dt, double e_prev,
                         double output_scale, double atol, double rtol, pstate *P,
                         pn_oracle_attempt_info *info) {
  int N = E->N, Ct = E->Ctot, d = E->d;
  size_t NN = (size_t)N * N;
  int fixedpoint = (E->cfg.strategy == PN_STRATEGY_FIXEDPOINT);
  precondition(E, dt, E->p, E->pinv);
  predict_mean(E, S->mean);
  linearise(E, S->t + dt, params);
  calibrate_and_estimate(E, dt);
  /* keep copies of what later phases overwrite */
  double *m_ext_keep = E->keep1, *m_p_keep = E->keep2;
  memcpy(m_ext_keep, E->m_ext, sizeof(double) * (size_t)N * Ct);
  memcpy(m_p_keep, E->m_p, sizeof(double) * (size_t)N * Ct);
  for (int f = 0; f < E->F; ++f) {
    double sigma = (E->cfg.calibration == PN_CALIB_DYNAMIC) ? E->sig[f] : output_scale;
    P->sigma[f] = sigma;
    int c0 = f * E->C;
    memcpy(E->m_p, m_p_keep, sizeof(double) * (size_t)N * Ct); /* merge_one clobbers m_p */
    predict_cov_one(E, S->chol + f * NN, sigma, E->L_ext + f * NN, fixedpoint, S->G + f * NN, S->g,
                    S->Lam + f * NN, P->G + f * NN, P->g, P->Lam + f * NN, c0, E->C);
  }
  memcpy(E->m_ext, m_ext_keep, sizeof(double) * (size_t)N * Ct);
  for (int f = 0; f < E->F; ++f) {
    correct_cov_one(E, E->L_ext + f * NN, P->chol + f * NN);
    correct_mean_one(E, P->mean, f * E->C, E->C);
  }
  P->t = S->t + dt;
  /* solver_mle (probdiffeq's running-mean calibration, no call site in the reference, [P?]): the whitened
   * squared residual of this step under the innovation covariance S, per ODE dimension */
  E->mle_x2 = (E->zz0 * E->invS0) * (1.0 / (double)E->C);
  /* scaled error norm: uses the PROPOSED u only (App. A.3, quirk confirmed against the golden) */
  for (int l = 0; l < d; ++l) E->fbuf[l] = E->err[l] * (1.0 / fma(rtol, fabs(P->mean[l]), atol));
  double acc = sum_squares(E, E->fbuf, d);
  double e = sqrt(acc) * E->inv_sqrt_d;
  info->error_norm = e;
  info->dt_proposed = pi_factor(E, e, e_prev) * dt;
  info->sigma = P->sigma[0];
  info->sigma_hat = E->sig[0];
}

/* prediction only (no correction) from S over dt with scale sigma[F]: used by interpolation.
 * with_bw: fixed-point backward model; run==NULL -> running conditional is the identity. */
static void extrapolate(engine *E, const pstate *S, double dt, const double *sigma, int with_bw,
                        const pstate *run, pstate *P) {
  int N = E->N, Ct = E->Ctot;
  size_t NN = (size_t)N * N;
  precondition(E, dt, E->p, E->pinv);
  predict_mean(E, S->mean);
  double *m_ext_keep = E->keep1, *m_p_keep = E->keep2;
  memcpy(m_ext_keep, E->m_ext, sizeof(double) * (size_t)N * Ct);
  memcpy(m_p_keep, E->m_p, sizeof(double) * (size_t)N * Ct);
  for (int f = 0; f < E->F; ++f) {
    memcpy(E->m_p, m_p_keep, sizeof(double) * (size_t)N * Ct);
    predict_cov_one(E, S->chol + f * NN, sigma[f], P->chol + f * NN, with_bw,
                    run ? run->G + f * NN : NULL, run ? run->g : NULL, run ? run->Lam + f * NN : NULL,
                    P->G + f * NN, P->g, P->Lam + f * NN, f * E->C, E->C);
    P->sigma[f] = sigma[f];
  }
  memcpy(P->mean, m_ext_keep, sizeof(double) * (size_t)N * Ct);
  P->t = S->t + dt;
}

static void marginalise(engine *E, const pstate *rv, const pstate *cond, pstate *out) {
  size_t NN = (size_t)E->N * E->N;
  for (int f = 0; f < E->F; ++f)
    marginalise_one(E, rv->mean, rv->chol + f * NN, cond->G + f * NN, cond->g, cond->Lam + f * NN,
                    out->mean, out->chol + f * NN, f * E->C, E->C);
}

static void initial_state(engine *E, const double *u0, const double *params, double t0,
                          double output_scale0, pstate *S) {
  size_t NN = (size_t)E->N * E->N;
  pn_oracle_taylor_init(E->cfg.problem, E->d, E->nu, E->q, u0, t0, params, S->mean);
  memset(S->chol, 0, sizeof(double) * E->F * NN);
  bw_identity(E, S);
  for (int f = 0; f < E->F; ++f) S->sigma[f] = output_scale0;
  S->t = t0;
}

static void marginal_std(const engine *E, const double *chol, double *std_out) {
  int N = E->N, d = E->d;
  size_t NN = (size_t)N * N;
  for (int l = 0; l < d; ++l) {
    const double *L;
    int row;
    if (E->dense) { L = chol; row = l; }
    else if (E->F == 1) { L = chol; row = 0; }
    else { L = chol + (size_t)l * NN; row = 0; }
    double acc = 0.0;
    for (int j = 0; j < N; ++j) acc = fma(L[row * N + j], L[row * N + j], acc);
    std_out[l] = sqrt(acc);
  }
}

static int has_nan(const double *x, size_t k) {
  for (size_t i = 0; i < k; ++i)
    if (x[i] != x[i]) return 1;
  return 0;
}

#define PN_TIME_EPS (10.0 * 2.220446049250313e-16)

/* Rejection loop + accept (the two inner loops of App. A.5 / SURVEY 3.2). Returns status. */
typedef struct {
  pstate step_from, interp_from, proposed;
  double dt, e_prev;
  double mle_ss; /* solver_mle: sum of mle_x2 over the accepted steps */
  int64_t n_accepted, n_rejected, n_attempts;
} adaptive_state;

static int adaptive_step(engine *E, const double *params, adaptive_state *A, double output_scale0,
                         double atol, double rtol, double *e_out) {
  pn_oracle_attempt_info info;
  for (;;) {
    if (E->cfg.max_attempts > 0 && A->n_attempts >= E->cfg.max_attempts) return PN_STATUS_MAX_ATTEMPTS;
    attempt_step(E, params, &A->step_from, A->dt, A->e_prev, output_scale0, atol, rtol, &A->proposed,
                 &info);
    A->n_attempts += 1;
    if (info.error_norm != info.error_norm) return PN_STATUS_NAN;
    A->dt = info.dt_proposed;
    if (info.error_norm <= 1.0) {
      A->e_prev = info.error_norm;
      if (e_out) *e_out = info.error_norm;
      /* interp_from <- step_from ; step_from <- proposed */
      pstate tmp = A->interp_from;
      A->interp_from = A->step_from;
      A->step_from = A->proposed;
      A->proposed = tmp;
      A->n_accepted += 1;
      A->mle_ss = A->mle_ss + E->mle_x2;
      return PN_STATUS_OK;
    }
    A->n_rejected += 1;
  }
}

static void adaptive_alloc(engine *E, adaptive_state *A) {
  pstate_alloc(E, &A->step_from);
  pstate_alloc(E, &A->interp_from);
  pstate_alloc(E, &A->proposed);
  A->n_accepted = A->n_rejected = A->n_attempts = 0;
  A->mle_ss = 0.0;
}
static void adaptive_free(adaptive_state *A) {
  pstate_free(&A->step_from);
  pstate_free(&A->interp_from);
  pstate_free(&A->proposed);
}

/* ------------------------------------------------------------------------- */
/* public: single attempted step                                              */
/* ------------------------------------------------------------------------- */
int pn_oracle_attempt_step(const pn_oracle_config *cfg, const double *params, double t, double dt,
                           double e_prev, double output_scale, const double *mean,
                           const double *chol, const double *bw_G, const double *bw_g,
                           const double *bw_Lam, double *mean_out, double *chol_out,
                           double *bw_G_out, double *bw_g_out, double *bw_Lam_out,
                           pn_oracle_attempt_info *info) {
  engine E;
  int rc = engine_init(&E, cfg);
  if (rc) return rc;
  size_t NN = (size_t)E.N * E.N, NC = (size_t)E.N * E.Ctot;
  pstate S, P;
  pstate_alloc(&E, &S);
  pstate_alloc(&E, &P);
  S.t = t;
  memcpy(S.mean, mean, sizeof(double) * NC);
  memcpy(S.chol, chol, sizeof(double) * E.F * NN);
  if (bw_G) {
    memcpy(S.G, bw_G, sizeof(double) * E.F * NN);
    memcpy(S.g, bw_g, sizeof(double) * NC);
    memcpy(S.Lam, bw_Lam, sizeof(double) * E.F * NN);
  } else {
    bw_identity(&E, &S);
  }
  attempt_step(&E, params, &S, dt, e_prev, output_scale, cfg->atol, cfg->rtol, &P, info);
  memcpy(mean_out, P.mean, sizeof(double) * NC);
  memcpy(chol_out, P.chol, sizeof(double) * E.F * NN);
  if (bw_G_out && cfg->strategy == PN_STRATEGY_FIXEDPOINT) {
    memcpy(bw_G_out, P.G, sizeof(double) * E.F * NN);
    memcpy(bw_g_out, P.g, sizeof(double) * NC);
    memcpy(bw_Lam_out, P.Lam, sizeof(double) * E.F * NN);
  }
  pstate_free(&S);
  pstate_free(&P);
  engine_free(&E);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* public: solve_adaptive_save_at + backward marginalisation                  */
/* ------------------------------------------------------------------------- */
/* optional outputs of the checkpoint solver beyond (u, u_std, marginals) */
typedef struct {
  const double *lml_data; /* [K,d] observations of u at the checkpoints              */
  const double *lml_std;  /* [K] observation noise standard deviations               */
  double *lml_out;        /* scalar                                                  */
  double *cond_out;       /* nullable [K, F*N*N + N*Ctot + F*N*N]: (G, g, Lam) of the
                             conditional checkpoint k -> k-1 (entry 0 unused)         */
  double *scale_out;      /* nullable [K, F] output scale carried by each checkpoint  */
} solve_extras;

/* stats.log_marginal_likelihood(u, standard_deviation=, posterior=) (src/odecheckpts/train_util.py:22-24,
 * experiments/old/6_learn_ode/learn.py:112-114): a Kalman filter that runs BACKWARDS over the
 * checkpoint Markov sequence.  Start from the terminal marginal; at checkpoint k observe
 * y_k = (0-th derivative) + N(0, std_k^2) in square-root form (QR of [[std, 0], [L^T e_0, L^T]]),
 * add log N(y_k; m_0, s^2) to a RUNNING MEAN over the data points (probdiffeq's estimator keeps
 * (rv, num_data, mean logpdf)), condition on y_k, and move to checkpoint k-1 through the stored
 * backward conditional.  Kronecker factorisations and dense with d == 1. */
#define PN_HALF_LOG_2PI 0.91893853320467274178
static double lml_sweep(engine *E, pstate *emit, int64_t K, const double *data, const double *std) {
  int N = E->N, Ct = E->Ctot, C = E->C, d = E->d, M1 = E->N + 1;
  size_t NN = (size_t)N * N;
  if (E->dense) return NAN;
  pstate rv, nxt;
  pstate_alloc(E, &rv);
  pstate_alloc(E, &nxt);
  pstate_copy(E, &rv, &emit[K - 1]);
  double *M = E->M2; /* (N+1)^2 <= 4 N^2 */
  double mean_lp = 0.0, ndata = 0.0;
  for (int64_t k = K - 1; k >= 0; --k) {
    double lp = 0.0;
    for (int f = 0; f < E->F; ++f) {
      double *L = rv.chol + f * NN;
      for (int j = 0; j < M1 * M1; ++j) M[j] = 0.0;
      M[0] = std[k];
      for (int i = 0; i < N; ++i) {
        M[(1 + i) * M1] = L[i]; /* (L^T e_0)_i = L[0][i] */
        for (int j = 0; j < N; ++j) M[(1 + i) * M1 + 1 + j] = L[j * N + i];
      }
      pn_qr_r(M, M1, M1);
      double s = M[0], inv_s = 1.0 / s;
      double log_s = pn_det_log(fabs(s));
      for (int c = f * C; c < f * C + C; ++c) {
        double z = rv.mean[c] - data[k * d + c];
        double w = z * inv_s;
        lp = fma(-0.5 * w, w, lp);
        for (int i = 0; i < N; ++i) {
          double gi = M[1 + i] * inv_s;
          rv.mean[i * Ct + c] = fma(-gi, z, rv.mean[i * Ct + c]);
        }
      }
      lp = fma(-(double)C, log_s + PN_HALF_LOG_2PI, lp);
      for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) L[i * N + j] = (j <= i) ? M[(1 + j) * M1 + 1 + i] : 0.0;
    }
    mean_lp = fma(mean_lp, ndata, lp) * (1.0 / (ndata + 1.0));
    ndata += 1.0;
    if (k == 0) break;
    marginalise(E, &rv, &emit[k], &nxt);
    pstate t2 = rv; rv = nxt; nxt = t2;
  }
  pstate_free(&rv);
  pstate_free(&nxt);
  return mean_lp;
}

static int solve_save_at_impl(engine *E, const double *u0, const double *params,
                              const double *save_at, int64_t K, double output_scale0, double atol,
                              double rtol, double *u, double *u_std, double *marg_mean,
                              double *marg_chol, int64_t *n_accepted, int64_t *n_rejected,
                              int32_t *status, double *filt_u, const solve_extras *X) {
  int N = E->N, Ct = E->Ctot, d = E->d;
  size_t NN = (size_t)N * N, NC = (size_t)N * Ct, FNN = (size_t)E->F * NN;
  int fixedpoint = (E->cfg.strategy == PN_STRATEGY_FIXEDPOINT);
  adaptive_state A;
  adaptive_alloc(E, &A);
  initial_state(E, u0, params, save_at[0], output_scale0, &A.step_from);
  pstate_copy(E, &A.interp_from, &A.step_from);
  A.dt = E->cfg.dt0;
  A.e_prev = 1.0;
  /* emitted marginals and conditionals, K entries (entry 0 = initial state) */
  pstate *emit = (pstate *)calloc((size_t)K, sizeof(pstate));
  for (int64_t k = 0; k < K; ++k) pstate_alloc(E, &emit[k]);
  pstate_copy(E, &emit[0], &A.step_from);
  pstate tmp_t, tmp_1;
  pstate_alloc(E, &tmp_t);
  pstate_alloc(E, &tmp_1);
  int st = PN_STATUS_OK;
  if (n_accepted) n_accepted[0] = 0;
  int64_t k_done = 0;
  for (int64_t k = 1; k < K && st == PN_STATUS_OK; ++k) {
    double t_next = save_at[k];
    while (A.step_from.t + PN_TIME_EPS < t_next) {
      st = adaptive_step(E, params, &A, output_scale0, atol, rtol, NULL);
      if (st != PN_STATUS_OK) break;
    }
    if (st != PN_STATUS_OK) break;
    if (A.step_from.t > t_next + PN_TIME_EPS) {
      /* overshoot: interpolate (App. A.5) */
      const pstate *S0 = &A.interp_from, *S1 = &A.step_from;
      if (fixedpoint) {
        extrapolate(E, S0, t_next - S0->t, S1->sigma, 1, S0, &tmp_t);  /* (m_t, L_t, bw_t) */
        tmp_t.t = t_next;
        extrapolate(E, &tmp_t, S1->t - t_next, S1->sigma, 1, NULL, &tmp_1); /* bw_1t from identity */
        /* rv_c = marginalise((m1,L1), bw_1t) */
        marginalise(E, S1, &tmp_1, &emit[k]);
        emit[k].t = t_next;
        memcpy(emit[k].G, tmp_t.G, sizeof(double) * FNN);
        memcpy(emit[k].g, tmp_t.g, sizeof(double) * NC);
        memcpy(emit[k].Lam, tmp_t.Lam, sizeof(double) * FNN);
        memcpy(emit[k].sigma, S1->sigma, sizeof(double) * E->F);
        /* step_from <- (t1, m1, L1, bw_1t); interp_from <- (t_c, m_t, L_t, identity) */
        memcpy(A.step_from.G, tmp_1.G, sizeof(double) * FNN);
        memcpy(A.step_from.g, tmp_1.g, sizeof(double) * NC);
        memcpy(A.step_from.Lam, tmp_1.Lam, sizeof(double) * FNN);
        pstate_copy(E, &A.interp_from, &tmp_t);
        bw_identity(E, &A.interp_from);
      } else {
        extrapolate(E, S0, t_next - S0->t, S1->sigma, 0, NULL, &tmp_t);
        tmp_t.t = t_next;
        bw_identity(E, &tmp_t);
        pstate_copy(E, &emit[k], &tmp_t);
        pstate_copy(E, &A.interp_from, &tmp_t);
      }
    } else {
      /* exact hit */
      pstate_copy(E, &emit[k], &A.step_from);
      emit[k].t = t_next;
      bw_identity(E, &A.step_from);
      pstate_copy(E, &A.interp_from, &A.step_from);
    }
    if (n_accepted) n_accepted[k] = A.n_accepted;
    if (E->cfg.calibration == PN_CALIB_MLE) /* solution.output_scale: the running MLE at emission */
      for (int f = 0; f < E->F; ++f)
        emit[k].sigma[f] = output_scale0 * ((A.n_accepted > 0) ? sqrt(A.mle_ss * (1.0 / (double)A.n_accepted)) : 1.0);
    k_done = k;
  }
  if (st == PN_STATUS_OK && has_nan(emit[K - 1].mean, NC)) st = PN_STATUS_NAN;
  /* outputs */
  if (st == PN_STATUS_OK) {
    if (filt_u)
      for (int64_t k = 0; k < K; ++k)
        for (int l = 0; l < d; ++l) filt_u[k * d + l] = emit[k].mean[l];
    /* A.6 backward sweep (fixed-point); filter: emitted marginals are the result */
    pstate rv, rv_prev;
    pstate_alloc(E, &rv);
    pstate_alloc(E, &rv_prev);
    pstate_copy(E, &rv, &emit[K - 1]);
    /* solver_mle: the final quasi-MLE rescales the posterior covariances */
    double mle_sc = 1.0;
    if (E->cfg.calibration == PN_CALIB_MLE && A.n_accepted > 0) mle_sc = sqrt(A.mle_ss * (1.0 / (double)A.n_accepted));
    for (int64_t k = K - 1; k >= 0; --k) {
      for (int l = 0; l < d; ++l) u[k * d + l] = rv.mean[l];
      marginal_std(E, rv.chol, u_std + k * d);
      for (int l = 0; l < d; ++l) u_std[k * d + l] = mle_sc * u_std[k * d + l];
      if (marg_mean) memcpy(marg_mean + (size_t)k * NC, rv.mean, sizeof(double) * NC);
      if (marg_chol)
        for (size_t e = 0; e < FNN; ++e) marg_chol[(size_t)k * FNN + e] = mle_sc * rv.chol[e];
      if (k == 0) break;
      if (fixedpoint) {
        marginalise(E, &rv, &emit[k], &rv_prev);
        pstate t2 = rv; rv = rv_prev; rv_prev = t2;
      } else {
        pstate_copy(E, &rv, &emit[k - 1]);
      }
    }
    pstate_free(&rv);
    pstate_free(&rv_prev);
    if (X && X->cond_out) {
      size_t stride = 2 * FNN + NC;
      for (int64_t k = 0; k < K; ++k) {
        memcpy(X->cond_out + k * stride, emit[k].G, sizeof(double) * FNN);
        memcpy(X->cond_out + k * stride + FNN, emit[k].g, sizeof(double) * NC);
        memcpy(X->cond_out + k * stride + FNN + NC, emit[k].Lam, sizeof(double) * FNN);
      }
    }
    if (X && X->scale_out)
      for (int64_t k = 0; k < K; ++k) memcpy(X->scale_out + k * E->F, emit[k].sigma, sizeof(double) * E->F);
    if (X && X->lml_out) *X->lml_out = fixedpoint ? lml_sweep(E, emit, K, X->lml_data, X->lml_std) : NAN;
  } else {
    if (X && X->lml_out) *X->lml_out = NAN;
    for (int64_t k = 0; k < K * d; ++k) { u[k] = NAN; u_std[k] = NAN; }
    if (n_accepted) for (int64_t k = k_done + 1; k < K; ++k) n_accepted[k] = A.n_accepted;
  }
  if (n_rejected) *n_rejected = A.n_rejected;
  if (status) *status = st;
  for (int64_t k = 0; k < K; ++k) pstate_free(&emit[k]);
  free(emit);
  pstate_free(&tmp_t);
  pstate_free(&tmp_1);
  adaptive_free(&A);
  return 0;
}

int pn_oracle_solve_save_at(const pn_oracle_config *cfg, const double *u0, const double *params,
                            const double *save_at, int64_t K, double output_scale0, double *u,
                            double *u_std, double *marg_mean, double *marg_chol,
                            int64_t *n_accepted, int64_t *n_rejected, int32_t *status,
                            double *filt_u) {
  engine E;
  int rc = engine_init(&E, cfg);
  if (rc) return rc;
  rc = solve_save_at_impl(&E, u0, params, save_at, K, output_scale0, cfg->atol, cfg->rtol, u, u_std,
                          marg_mean, marg_chol, n_accepted, n_rejected, status, filt_u, NULL);
  engine_free(&E);
  return rc;
}

int pn_oracle_solve_save_at_lml(const pn_oracle_config *cfg, const double *u0, const double *params,
                                const double *save_at, int64_t K, double output_scale0,
                                const double *data, const double *obs_std, double *u, double *u_std,
                                double *marg_mean, double *marg_chol, double *cond_out,
                                double *scale_out, double *lml, int32_t *status) {
  engine E;
  int rc = engine_init(&E, cfg);
  if (rc) return rc;
  solve_extras X = {data, obs_std, lml, cond_out, scale_out};
  rc = solve_save_at_impl(&E, u0, params, save_at, K, output_scale0, cfg->atol, cfg->rtol, u, u_std,
                          marg_mean, marg_chol, NULL, NULL, status, NULL, &X);
  engine_free(&E);
  return rc;
}

int pn_oracle_solve_save_at_batch(const pn_oracle_config *cfg, int64_t B, const double *u0,
                                  const double *params, const double *tol, const double *save_at,
                                  int64_t K, const double *output_scale0, double *u, double *u_std,
                                  int64_t *n_accepted, int64_t *n_rejected, int32_t *status,
                                  int num_threads) {
  int d = cfg->d, q = cfg->ode_order, P = cfg->num_params;
  int rc_all = 0;
#ifdef _OPENMP
  if (num_threads <= 0) num_threads = omp_get_max_threads();
#else
  (void)num_threads;
#endif
#pragma omp parallel num_threads(num_threads)
  {
    engine E;
    int rc = engine_init(&E, cfg);
    if (rc) {
#pragma omp critical
      rc_all = rc;
    } else {
#pragma omp for schedule(dynamic, 1)
      for (int64_t b = 0; b < B; ++b) {
        double atol = tol ? tol[2 * b] : cfg->atol, rtol = tol ? tol[2 * b + 1] : cfg->rtol;
        double os0 = output_scale0 ? output_scale0[b] : 1.0;
        solve_save_at_impl(&E, u0 + (size_t)b * q * d, params + (size_t)b * P, save_at, K, os0, atol,
                           rtol, u + (size_t)b * K * d, u_std + (size_t)b * K * d, NULL, NULL,
                           n_accepted + (size_t)b * K, n_rejected + b, status + b, NULL, NULL);
      }
      engine_free(&E);
    }
  }
  return rc_all;
}

/* ------------------------------------------------------------------------- */
/* public: solve_adaptive_save_every_step (vdp.py:77-79)                      */
/* ------------------------------------------------------------------------- */
int64_t pn_oracle_solve_save_every_step(const pn_oracle_config *cfg, const double *u0,
                                        const double *params, double t0, double t1,
                                        double output_scale0, int64_t max_grid, double *grid,
                                        double *u, double *u_std, int64_t *n_rejected) {
  engine E;
  int rc = engine_init(&E, cfg);
  if (rc) return rc;
  int d = E.d;
  adaptive_state A;
  adaptive_alloc(&E, &A);
  initial_state(&E, u0, params, t0, output_scale0, &A.step_from);
  pstate_copy(&E, &A.interp_from, &A.step_from);
  A.dt = cfg->dt0;
  A.e_prev = 1.0;
  int64_t cnt = 0;
  grid[0] = t0;
  for (int l = 0; l < d; ++l) u[l] = A.step_from.mean[l];
  if (u_std) marginal_std(&E, A.step_from.chol, u_std);
  cnt = 1;
  pstate tmp;
  pstate_alloc(&E, &tmp);
  int64_t ret = 0;
  while (A.step_from.t + PN_TIME_EPS < t1) {
    int st = adaptive_step(&E, params, &A, output_scale0, cfg->atol, cfg->rtol, NULL);
    if (st != PN_STATUS_OK) { ret = -2; break; }
    if (cnt >= max_grid) { ret = -1; break; }
    const pstate *rec = &A.step_from;
    if (A.step_from.t > t1 + PN_TIME_EPS) {
      /* final point: interpolate at t1 from interp_from (extrapolation for the filter;
       * for the fixed-point strategy the marginal informed by the overshooting step) */
      const pstate *S0 = &A.interp_from, *S1 = &A.step_from;
      if (cfg->strategy == PN_STRATEGY_FIXEDPOINT) {
        pstate tmp1;
        pstate_alloc(&E, &tmp1);
        extrapolate(&E, S0, t1 - S0->t, S1->sigma, 1, S0, &tmp);
        tmp.t = t1;
        extrapolate(&E, &tmp, S1->t - t1, S1->sigma, 1, NULL, &tmp1);
        marginalise(&E, S1, &tmp1, &tmp);
        pstate_free(&tmp1);
      } else {
        extrapolate(&E, S0, t1 - S0->t, S1->sigma, 0, NULL, &tmp);
      }
      tmp.t = t1;
      rec = &tmp;
      grid[cnt] = t1;
    } else {
      grid[cnt] = A.step_from.t;
    }
    for (int l = 0; l < d; ++l) u[cnt * d + l] = rec->mean[l];
    if (u_std) marginal_std(&E, rec->chol, u_std + cnt * d);
    cnt += 1;
    if (rec == &tmp) break;
  }
  if (n_rejected) *n_rejected = A.n_rejected;
  pstate_free(&tmp);
  adaptive_free(&A);
  engine_free(&E);
  return ret < 0 ? ret : cnt;
}

/* ------------------------------------------------------------------------- */
/* public: solve_fixed_grid (vdp.py:88-91)                                    */
/* ------------------------------------------------------------------------- */
int pn_oracle_solve_fixed_grid(const pn_oracle_config *cfg, const double *u0, const double *params,
                               const double *grid, int64_t G, double output_scale0, double *u,
                               double *u_std, double *err_norms) {
  engine E;
  int rc = engine_init(&E, cfg);
  if (rc) return rc;
  int d = E.d;
  pstate S, P;
  pstate_alloc(&E, &S);
  pstate_alloc(&E, &P);
  initial_state(&E, u0, params, grid[0], output_scale0, &S);
  for (int l = 0; l < d; ++l) u[l] = S.mean[l];
  if (u_std) marginal_std(&E, S.chol, u_std);
  if (err_norms) err_norms[0] = 0.0;
  pn_oracle_attempt_info info;
  for (int64_t k = 1; k < G; ++k) {
    double dt = grid[k] - grid[k - 1];
    attempt_step(&E, params, &S, dt, 1.0, output_scale0, cfg->atol, cfg->rtol, &P, &info);
    P.t = grid[k];
    pstate t2 = S; S = P; P = t2;
    for (int l = 0; l < d; ++l) u[k * d + l] = S.mean[l];
    if (u_std) marginal_std(&E, S.chol, u_std + k * d);
    if (err_norms) err_norms[k] = info.error_norm;
  }
  pstate_free(&S);
  pstate_free(&P);
  engine_free(&E);
  return 0;
}
