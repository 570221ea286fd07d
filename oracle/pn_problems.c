/*
 * pn_problems.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Vector fields, Jacobians and Taylor-mode initialisation of the reference's
 * IVP zoo (src/odecheckpts/ivps.py).  The reference obtains Jacobians and the
 * Taylor coefficients by JAX autodiff (probdiffeq `taylor.odejet_padded_scan`,
 * src/odecheckpts/ivpsolvers.py:63-67; `correction_ts1`,
 * experiments/1_van_der_pol/vdp.py:64).  Here they are analytic Jacobians and
 * truncated power-series ("jet") recurrences.
 *
 * Expression order is part of the contract with the CUDA functors
 * (code-adaptive-prob-ode-solvers_b200/csrc/pn_problems.cuh mirrors it).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "pn_internal.h"

/* ------------------------------------------------------------------------- */
/* plain vector fields                                                        */
/* ------------------------------------------------------------------------- */
static inline double inv_pow32(double s) { return 1.0 / (s * sqrt(s)); } /* s^(-3/2) */

void pn_oracle_vf(int problem, int d, const double *u, double t, const double *params, double *f) {
  (void)t;
  switch (problem) {
    case PN_PROBLEM_LOGISTIC: { /* ivps.py:8-17, diffeqzoo: a u (1 - b u) */
      double a = params[0], b = params[1];
      f[0] = (a * u[0]) * fma(-b, u[0], 1.0);
      break;
    }
    case PN_PROBLEM_RIGID_BODY: { /* ivps.py:20-29, diffeqzoo: (a u1 u2, b u0 u2, c u0 u1) */
      f[0] = params[0] * (u[1] * u[2]);
      f[1] = params[1] * (u[0] * u[2]);
      f[2] = params[2] * (u[0] * u[1]);
      break;
    }
    case PN_PROBLEM_LOTKA_VOLTERRA: { /* diffeqzoo: (a u0 - b u0 u1, -c u1 + d u0 u1) */
      double uv = u[0] * u[1];
      f[0] = fma(-params[1], uv, params[0] * u[0]);
      f[1] = fma(params[3], uv, -(params[2] * u[1]));
      break;
    }
    case PN_PROBLEM_VAN_DER_POL: { /* ivps.py:159-167: mu (u' (1 - u^2) - u) */
      double mu = params[0];
      double y = u[0], yd = u[1];
      f[0] = mu * fma(yd, fma(-y, y, 1.0), -y);
      break;
    }
    case PN_PROBLEM_THREE_BODY: { /* ivps.py:32-41, diffeqzoo three_body_restricted */
      double mu = params[0], mp = 1.0 - mu;
      double x = u[0], y = u[1], xd = u[2], yd = u[3];
      double a = x + mu, b = x - mp;
      double p1 = inv_pow32(fma(y, y, a * a));
      double p2 = inv_pow32(fma(y, y, b * b));
      f[0] = fma(-mu, b * p2, fma(-mp, a * p1, fma(2.0, yd, x)));
      f[1] = fma(-mu, y * p2, fma(-mp, y * p1, fma(-2.0, xd, y)));
      break;
    }
    case PN_PROBLEM_PLEIADES: { /* ivps.py:59-99; u = (x[7], y[7]); masses 1..7 */
      const double *x = u, *y = u + 7;
      for (int i = 0; i < 7; ++i) {
        double ax = 0.0, ay = 0.0;
        for (int j = 0; j < 7; ++j) {
          if (j == i) continue; /* nan_to_num(0/0) = 0, ivps.py:95-96 */
          double dx = x[j] - x[i], dy = y[j] - y[i];
          double p = inv_pow32(fma(dy, dy, dx * dx));
          double mj = (double)(j + 1);
          ax = fma(mj, p * dx, ax);
          ay = fma(mj, p * dy, ay);
        }
        f[i] = ax;
        f[7 + i] = ay;
      }
      break;
    }
    case PN_PROBLEM_BRUSSELATOR: { /* ivps.py:124-156; u = (u[N], v[N]); params (alpha) */
      int N = d / 2;
      double c = params[0] * (double)((N + 1) * (N + 1));
      const double *uu = u, *vv = u + N;
      for (int i = 0; i < N; ++i) {
        double ul = (i == 0) ? 1.0 : uu[i - 1], ur = (i == N - 1) ? 1.0 : uu[i + 1];
        double vl = (i == 0) ? 3.0 : vv[i - 1], vr = (i == N - 1) ? 3.0 : vv[i + 1];
        double uuv = (uu[i] * uu[i]) * vv[i];
        double lap_u = fma(-2.0, uu[i], ul + ur);
        double lap_v = fma(-2.0, vv[i], vl + vr);
        f[i] = fma(c, lap_u, fma(-4.0, uu[i], 1.0 + uuv));
        f[N + i] = fma(c, lap_v, fma(3.0, uu[i], -uuv));
      }
      break;
    }
    default:
      for (int i = 0; i < d; ++i) f[i] = NAN;
  }
}

int pn_problem_has_jacobian(int problem) {
  switch (problem) {
    case PN_PROBLEM_LOGISTIC:
    case PN_PROBLEM_RIGID_BODY:
    case PN_PROBLEM_LOTKA_VOLTERRA:
    case PN_PROBLEM_VAN_DER_POL:
    case PN_PROBLEM_THREE_BODY:
    case PN_PROBLEM_BRUSSELATOR:
      return 1;
    default:
      return 0;
  }
}

/* jac: d x (q*d) row-major; column k*d + l = d f_i / d u^{(k)}_l */
void pn_oracle_jac(int problem, int d, const double *u, double t, const double *params, double *jac) {
  (void)t;
  switch (problem) {
    case PN_PROBLEM_LOGISTIC: {
      double a = params[0], b = params[1];
      jac[0] = a * fma(-2.0 * b, u[0], 1.0);
      break;
    }
    case PN_PROBLEM_RIGID_BODY: {
      double a = params[0], b = params[1], c = params[2];
      jac[0] = 0.0;        jac[1] = a * u[2];   jac[2] = a * u[1];
      jac[3] = b * u[2];   jac[4] = 0.0;        jac[5] = b * u[0];
      jac[6] = c * u[1];   jac[7] = c * u[0];   jac[8] = 0.0;
      break;
    }
    case PN_PROBLEM_LOTKA_VOLTERRA: {
      double a = params[0], b = params[1], c = params[2], dd = params[3];
      jac[0] = fma(-b, u[1], a);  jac[1] = -(b * u[0]);
      jac[2] = dd * u[1];         jac[3] = fma(dd, u[0], -c);
      break;
    }
    case PN_PROBLEM_VAN_DER_POL: {
      double mu = params[0];
      double y = u[0], yd = u[1];
      jac[0] = mu * fma(-2.0 * y, yd, -1.0); /* d/du  */
      jac[1] = mu * fma(-y, y, 1.0);         /* d/du' */
      break;
    }
    case PN_PROBLEM_THREE_BODY: {
      /* f0 = x + 2 y' - mp a p1 - mu b p2, f1 = y - 2 x' - mp y p1 - mu y p2,
       * p = s^(-3/2), s1 = a^2 + y^2, s2 = b^2 + y^2; dp/ds = -1.5 p / s */
      double mu = params[0], mp = 1.0 - mu;
      double x = u[0], y = u[1];
      double a = x + mu, b = x - mp;
      double s1 = fma(y, y, a * a), s2 = fma(y, y, b * b);
      double p1 = inv_pow32(s1), p2 = inv_pow32(s2);
      double q1 = (-3.0 * p1) * (1.0 / s1), q2 = (-3.0 * p2) * (1.0 / s2); /* dp/d(a or y) = q * (a or y) */
      /* row 0: d/dx, d/dy, d/dx', d/dy' */
      jac[0] = 1.0 - fma(mu, fma(b * b, q2, p2), mp * fma(a * a, q1, p1));
      jac[1] = -fma(mu, (b * y) * q2, mp * ((a * y) * q1));
      jac[2] = 0.0;
      jac[3] = 2.0;
      jac[4] = jac[1];
      jac[5] = 1.0 - fma(mu, fma(y * y, q2, p2), mp * fma(y * y, q1, p1));
      jac[6] = -2.0;
      jac[7] = 0.0;
      break;
    }
    case PN_PROBLEM_BRUSSELATOR: {
      int N = d / 2;
      double c = params[0] * (double)((N + 1) * (N + 1));
      memset(jac, 0, sizeof(double) * (size_t)d * (size_t)d);
      const double *uu = u, *vv = u + N;
      for (int i = 0; i < N; ++i) {
        double two_uv = (2.0 * uu[i]) * vv[i];
        double u2 = uu[i] * uu[i];
        jac[i * d + i] = fma(-2.0, c, two_uv - 4.0);
        jac[i * d + N + i] = u2;
        jac[(N + i) * d + i] = 3.0 - two_uv;
        jac[(N + i) * d + N + i] = fma(-2.0, c, -u2);
        if (i > 0) {
          jac[i * d + i - 1] = c;
          jac[(N + i) * d + N + i - 1] = c;
        }
        if (i < N - 1) {
          jac[i * d + i + 1] = c;
          jac[(N + i) * d + N + i + 1] = c;
        }
      }
      break;
    }
    default:
      break;
  }
}

/* ------------------------------------------------------------------------- */
/* truncated power series ("jets"), normalised coefficients c_k = u^(k)/k!     */
/* ------------------------------------------------------------------------- */
static void jet_mul(const double *a, const double *b, double *out, int n) {
  for (int k = 0; k < n; ++k) {
    double acc = a[0] * b[k];
    for (int j = 1; j <= k; ++j) acc = fma(a[j], b[k - j], acc);
    out[k] = acc;
  }
}

/* out = s^(-3/2) */
static void jet_inv_pow32(const double *s, double *out, int n) {
  out[0] = inv_pow32(s[0]);
  for (int k = 1; k < n; ++k) {
    double acc = 0.0;
    for (int j = 1; j <= k; ++j) {
      double coef = -1.5 * (double)j - (double)(k - j);
      acc = fma(coef * s[j], out[k - j], acc);
    }
    out[k] = acc / ((double)k * s[0]);
  }
}

/* Evaluate the vector field on jets.  U: q*d jets of length n (U[(k*d+l)*n + i]), F: d jets. */
static void vf_jet(int problem, int d, int n, const double *U, const double *params, double *F,
                   double *work) {
  switch (problem) {
    case PN_PROBLEM_LOGISTIC: {
      double a = params[0], b = params[1];
      double *au = work, *w = work + n;
      for (int k = 0; k < n; ++k) {
        au[k] = a * U[k];
        w[k] = fma(-b, U[k], (k == 0) ? 1.0 : 0.0);
      }
      jet_mul(au, w, F, n);
      break;
    }
    case PN_PROBLEM_RIGID_BODY: {
      double *pr = work;
      jet_mul(U + 1 * n, U + 2 * n, pr, n);
      for (int k = 0; k < n; ++k) F[0 * n + k] = params[0] * pr[k];
      jet_mul(U + 0 * n, U + 2 * n, pr, n);
      for (int k = 0; k < n; ++k) F[1 * n + k] = params[1] * pr[k];
      jet_mul(U + 0 * n, U + 1 * n, pr, n);
      for (int k = 0; k < n; ++k) F[2 * n + k] = params[2] * pr[k];
      break;
    }
    case PN_PROBLEM_LOTKA_VOLTERRA: {
      double *uv = work;
      jet_mul(U, U + n, uv, n);
      for (int k = 0; k < n; ++k) {
        F[k] = fma(-params[1], uv[k], params[0] * U[k]);
        F[n + k] = fma(params[3], uv[k], -(params[2] * U[n + k]));
      }
      break;
    }
    case PN_PROBLEM_VAN_DER_POL: {
      double mu = params[0];
      const double *y = U, *yd = U + n;
      double *yy = work, *w = work + n, *pr = work + 2 * n;
      jet_mul(y, y, yy, n);
      /* w = 1 - y^2 with the k=0 coefficient formed as fma(-y0, y0, 1) like the plain vf */
      w[0] = fma(-y[0], y[0], 1.0);
      for (int k = 1; k < n; ++k) w[k] = -yy[k];
      jet_mul(yd, w, pr, n);
      /* k = 0 must equal fma(yd, w0, -y): jet_mul's k=0 term is yd0*w0 (a product), so redo it */
      F[0] = mu * fma(yd[0], w[0], -y[0]);
      for (int k = 1; k < n; ++k) F[k] = mu * (pr[k] - y[k]);
      break;
    }
    case PN_PROBLEM_THREE_BODY: {
      double mu = params[0], mp = 1.0 - mu;
      const double *x = U, *y = U + n, *xd = U + 2 * n, *yd = U + 3 * n;
      double *a = work, *b = work + n, *s1 = work + 2 * n, *s2 = work + 3 * n, *p1 = work + 4 * n,
             *p2 = work + 5 * n, *t1 = work + 6 * n, *t2 = work + 7 * n, *yy = work + 8 * n;
      for (int k = 0; k < n; ++k) {
        a[k] = (k == 0) ? x[0] + mu : x[k];
        b[k] = (k == 0) ? x[0] - mp : x[k];
      }
      jet_mul(y, y, yy, n);
      jet_mul(a, a, t1, n);
      jet_mul(b, b, t2, n);
      /* k=0: fma(y,y,a*a) as in the plain vf; higher coefficients: sum of the two squares */
      s1[0] = fma(y[0], y[0], a[0] * a[0]);
      s2[0] = fma(y[0], y[0], b[0] * b[0]);
      for (int k = 1; k < n; ++k) {
        s1[k] = yy[k] + t1[k];
        s2[k] = yy[k] + t2[k];
      }
      jet_inv_pow32(s1, p1, n);
      jet_inv_pow32(s2, p2, n);
      jet_mul(a, p1, t1, n); /* a p1 */
      jet_mul(b, p2, t2, n); /* b p2 */
      for (int k = 0; k < n; ++k) F[k] = fma(-mu, t2[k], fma(-mp, t1[k], fma(2.0, yd[k], x[k])));
      jet_mul(y, p1, t1, n);
      jet_mul(y, p2, t2, n);
      for (int k = 0; k < n; ++k) F[n + k] = fma(-mu, t2[k], fma(-mp, t1[k], fma(-2.0, xd[k], y[k])));
      break;
    }
    case PN_PROBLEM_PLEIADES: {
      const double *x = U, *y = U + 7 * n;
      double *dx = work, *dy = work + n, *s = work + 2 * n, *p = work + 3 * n, *t1 = work + 4 * n,
             *t2 = work + 5 * n;
      for (int i = 0; i < 7; ++i) {
        double *ax = F + i * n, *ay = F + (7 + i) * n;
        for (int k = 0; k < n; ++k) ax[k] = ay[k] = 0.0;
        for (int j = 0; j < 7; ++j) {
          if (j == i) continue;
          double mj = (double)(j + 1);
          for (int k = 0; k < n; ++k) {
            dx[k] = x[j * n + k] - x[i * n + k];
            dy[k] = y[j * n + k] - y[i * n + k];
          }
          jet_mul(dx, dx, t1, n);
          jet_mul(dy, dy, t2, n);
          s[0] = fma(dy[0], dy[0], dx[0] * dx[0]);
          for (int k = 1; k < n; ++k) s[k] = t2[k] + t1[k];
          jet_inv_pow32(s, p, n);
          jet_mul(p, dx, t1, n);
          jet_mul(p, dy, t2, n);
          for (int k = 0; k < n; ++k) {
            ax[k] = fma(mj, t1[k], ax[k]);
            ay[k] = fma(mj, t2[k], ay[k]);
          }
        }
      }
      break;
    }
    case PN_PROBLEM_BRUSSELATOR: {
      int N = d / 2;
      double c = params[0] * (double)((N + 1) * (N + 1));
      double *u2 = work, *uuv = work + n;
      for (int i = 0; i < N; ++i) {
        const double *ui = U + i * n, *vi = U + (N + i) * n;
        jet_mul(ui, ui, u2, n);
        jet_mul(u2, vi, uuv, n);
        for (int k = 0; k < n; ++k) {
          double padu = (k == 0) ? 1.0 : 0.0, padv = (k == 0) ? 3.0 : 0.0;
          double ul = (i == 0) ? padu : U[(i - 1) * n + k];
          double ur = (i == N - 1) ? padu : U[(i + 1) * n + k];
          double vl = (i == 0) ? padv : U[(N + i - 1) * n + k];
          double vr = (i == N - 1) ? padv : U[(N + i + 1) * n + k];
          double lap_u = fma(-2.0, ui[k], ul + ur);
          double lap_v = fma(-2.0, vi[k], vl + vr);
          F[i * n + k] = fma(c, lap_u, fma(-4.0, ui[k], padu + uuv[k]));
          F[(N + i) * n + k] = fma(c, lap_v, fma(3.0, ui[k], -uuv[k]));
        }
      }
      break;
    }
    default:
      break;
  }
}

/* taylor.odejet_padded_scan / odejet_unroll replacement (ivpsolvers.py:63-67, run.py:64).
 * tcoeffs[k*d + l] = u_l^{(k)}(t0), k = 0..nu. */
void pn_oracle_taylor_init(int problem, int d, int nu, int q, const double *u0, double t0,
                           const double *params, double *tcoeffs) {
  (void)t0;
  int n = nu + 1;
  /* C: d jets of u (normalised); U: q*d jets handed to the vector field; F: d jets */
  double *C = (double *)calloc((size_t)d * n, sizeof(double));
  double *U = (double *)calloc((size_t)q * d * n, sizeof(double));
  double *F = (double *)calloc((size_t)d * n, sizeof(double));
  double *work = (double *)calloc((size_t)16 * n, sizeof(double));
  for (int l = 0; l < d; ++l) {
    C[l * n + 0] = u0[l];
    if (q == 2 && n > 1) C[l * n + 1] = u0[d + l];
  }
  for (int k = 0; k + q <= nu; ++k) {
    /* assemble the argument jets from what is known so far */
    for (int l = 0; l < d; ++l) {
      for (int i = 0; i < n; ++i) U[l * n + i] = C[l * n + i];
      if (q == 2)
        for (int i = 0; i < n; ++i)
          U[(d + l) * n + i] = (i + 1 < n) ? (double)(i + 1) * C[l * n + i + 1] : 0.0;
    }
    vf_jet(problem, d, n, U, params, F, work);
    for (int l = 0; l < d; ++l) {
      if (q == 1)
        C[l * n + k + 1] = F[l * n + k] / (double)(k + 1);
      else
        C[l * n + k + 2] = F[l * n + k] / (double)((k + 2) * (k + 1));
    }
  }
  double fact = 1.0;
  for (int k = 0; k <= nu; ++k) {
    if (k > 0) fact *= (double)k;
    for (int l = 0; l < d; ++l) tcoeffs[k * d + l] = fact * C[l * n + k];
  }
  free(C);
  free(U);
  free(F);
  free(work);
}
