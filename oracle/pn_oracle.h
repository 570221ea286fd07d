/*
 * pn_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C99, fp64 restatement of the adaptive probabilistic IVP solver loop
 * that pnkraemer/code-adaptive-prob-ode-solvers (`odecheckpts`) runs through the
 * third-party package `probdiffeq` (unpinned in the reference's
 * pyproject.toml:8-11; API generation ~0.5-0.6, identified by the calls at
 * src/odecheckpts/ivpsolvers.py:33,42-53,65-81).  probdiffeq's sources are not
 * under /root/reference and neither it nor jax is installable here, so this file
 * restates its published algorithm and is PINNED against
 * the golden artefacts the reference commits (experiments/ ... .npy, extracted
 * to tests/golden/reference_goldens.npz by tests/golden/make_golden.py):
 *   - Brusselator accepted-step counts 610 (N=4) and 3294 (N=16) and the 200
 *     smoothed checkpoint means (experiments/4_brusselator/run.py:119-138),
 *   - three-body accepted-step counts 448 / 2570 / 14469
 *     (experiments/5_vs_interpolation/measure.py:44-68,191-192),
 *   - rigid-body grid lengths and checkpoint RMSEs
 *     (experiments/2_workprec_simple/run_simple.py:38-80,181-215),
 *   - the stiff Van-der-Pol adaptive grid + filter solution
 *     (experiments/1_van_der_pol/vdp.py:61-80).
 * See tests/test_oracle_golden.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product (the CUDA library under
 * code-adaptive-prob-ode-solvers_b200/csrc) never links or calls it.
 *
 * Arithmetic contract (what makes GPU-vs-oracle comparisons bit-exact for the
 * thread-per-IVP kernels): IEEE-754 binary64, round-to-nearest; every fused
 * multiply-add is an explicit fma() call and the file is compiled with
 * -ffp-contract=off; sums run in ascending index order; transcendental
 * functions (the two pow() calls of the PI controller) use the explicit
 * polynomial kernels pn_det_log / pn_det_exp below instead of libm.
 */
#ifndef PN_ORACLE_H
#define PN_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* problem ids: src/odecheckpts/ivps.py */
enum {
  PN_PROBLEM_LOGISTIC = 0,      /* ivps.py:8-17   d=1  q=1  params (a, b)        */
  PN_PROBLEM_RIGID_BODY = 1,    /* ivps.py:20-29  d=3  q=1  params (a, b, c)     */
  PN_PROBLEM_THREE_BODY = 2,    /* ivps.py:32-41  d=2  q=2  params (mu)          */
  PN_PROBLEM_PLEIADES = 3,      /* ivps.py:59-99  d=14 q=2  no params            */
  PN_PROBLEM_BRUSSELATOR = 4,   /* ivps.py:124-156 d=2N q=1 params (alpha)       */
  PN_PROBLEM_VAN_DER_POL = 5,   /* ivps.py:159-167 d=1 q=2  params (mu)          */
  PN_PROBLEM_LOTKA_VOLTERRA = 6 /* diffeqzoo default; d=2 q=1 params (a,b,c,d)   */
};

enum { PN_FACT_ISOTROPIC = 0, PN_FACT_BLOCKDIAG = 1, PN_FACT_DENSE = 2 };
enum { PN_CORR_TS0 = 0, PN_CORR_TS1 = 1 };
enum { PN_STRATEGY_FILTER = 0, PN_STRATEGY_FIXEDPOINT = 1 };
enum { PN_CALIB_NONE = 0, PN_CALIB_DYNAMIC = 1, PN_CALIB_MLE = 2 /* solver_mle: running quasi-MLE, see pn_solver.c */ };
enum { PN_STATUS_OK = 0, PN_STATUS_NAN = 1, PN_STATUS_MAX_ATTEMPTS = 2 };

/* Mirrors the solver construction of src/odecheckpts/ivpsolvers.py:14-53. */
typedef struct {
  int32_t problem;       /* PN_PROBLEM_*                                          */
  int32_t d;             /* ODE dimension                                         */
  int32_t nu;            /* prior_ibm(num_derivatives=nu)        ivpsolvers.py:42 */
  int32_t ode_order;     /* correction_ts0(ode_order=q)          ivpsolvers.py:37 */
  int32_t factorisation; /* impl.select(...)                     ivpsolvers.py:33 */
  int32_t correction;    /* ts0 | ts1                                             */
  int32_t strategy;      /* filter | fixedpoint                  ivpsolvers.py:43 */
  int32_t calibration;   /* solver | solver_dynamic              ivpsolvers.py:45 */
  double atol, rtol;     /* ivpsolve.adaptive(atol=, rtol=)      ivpsolvers.py:53 */
  double dt0;
  /* control_proportional_integral() defaults: 0.95, 0.2, 10.0, 0.3, 0.4 */
  double safety, factor_min, factor_max, power_integral, power_proportional;
  int64_t max_attempts;  /* per member; <=0: unlimited                            */
  int32_t num_params;
  /* Summation order of the two norms over the ODE dimensions (||z|| for the isotropic calibration
   * and the scaled error norm).  <= 1: ascending-index fma chain (what a thread-per-IVP kernel
   * does).  G > 1 (power of two): lane l sums the squares of entries l, l+G, ... in order, then a
   * butterfly (xor 2^k) inside each group of 32 lanes and the group sums added in order -- the order
   * the lane-per-dimension (G <= 32) and CTA-per-IVP (G = 128) CUDA kernels use.  Any order is a
   * faithful restatement; this knob only exists so that comparisons can be bit-exact. */
  int32_t reduction_group;
  /* Dense factorisation with d > 1 only.  0: the unblocked column-by-column Householder QR / back
   * substitution of pn_linalg.c (what the warp-per-IVP CUDA kernels reproduce, D <= 40).  nb > 0:
   * panel-blocked compact-WY QR with panels of nb columns and blocked back substitution in the
   * operation order of the CTA-per-IVP tensor-core kernel (pn_blocked.c).  Both are Householder QRs
   * of the same matrices: results agree to rounding (tests/test_oracle_blocked.py), and the blocked
   * order exists so that the large-D GPU kernel can be compared bit for bit. */
  int32_t dense_block;
} pn_oracle_config;

/* ---- deterministic elementary functions (shared contract with the kernel) ---- */
double pn_det_log(double x);
double pn_det_exp(double y);
double pn_det_pow(double x, double y); /* x >= 0, y > 0 */

/* ---- prior constants ---- */
/* a1: n*n row-major flipped Pascal; lq: n*n row-major lower Cholesky of flipped Hilbert */
void pn_oracle_prior(int nu, double *a1, double *lq);

/* ---- vector fields, Jacobians, Taylor-mode initialisation ---- */
/* u: q*d (u, u', ...), f: d */
void pn_oracle_vf(int problem, int d, const double *u, double t, const double *params, double *f);
/* jac: d x (q*d) row-major: d f_i / d u^{(k)}_l at column k*d+l */
void pn_oracle_jac(int problem, int d, const double *u, double t, const double *params, double *jac);
/* tcoeffs: (nu+1)*d unnormalised derivatives u^{(k)}(t0), row k */
void pn_oracle_taylor_init(int problem, int d, int nu, int q, const double *u0, double t0,
                           const double *params, double *tcoeffs);

/* ---- a single attempted step from a given state (unit-test entry point) ----
 * Layouts ("kron" engine: isotropic, blockdiag, dense with d==1):
 *   mean[n*d] row-major (derivative, dimension); chol[F*n*n] with F=1 (iso, dense d=1)
 *   or F=d (blockdiag), each n*n row-major lower triangular;
 *   backward conditional G[F*n*n], g[n*d], Lam[F*n*n].
 * Dense engine (dense with d>1): mean[D], chol[D*D], G[D*D], g[D], Lam[D*D],
 *   D=n*d, derivative-major index i*d+j.
 * Outputs are the PROPOSED state (whether or not it is accepted). */
typedef struct {
  double error_norm;   /* scaled error e (accept iff <= 1)                        */
  double dt_proposed;  /* PI proposal for the next attempt                        */
  double sigma;        /* output scale used for this step's process noise         */
  double sigma_hat;    /* local calibration (first factor set / dense scalar)     */
} pn_oracle_attempt_info;

int pn_oracle_attempt_step(const pn_oracle_config *cfg, const double *params, double t, double dt,
                           double e_prev, double output_scale, const double *mean,
                           const double *chol, const double *bw_G, const double *bw_g,
                           const double *bw_Lam, double *mean_out, double *chol_out,
                           double *bw_G_out, double *bw_g_out, double *bw_Lam_out,
                           pn_oracle_attempt_info *info);

/* ---- outer loops ---- */
/* solve_adaptive_save_at + backward marginalisation (ivpsolvers.py:71-89).
 * u, u_std: [K,d]; marg_mean: nullable [K,n,d]; marg_chol: nullable
 * [K,F,n,n] (kron) / [K,D,D] (dense); n_accepted: [K] cumulative accepted steps
 * when checkpoint k was emitted (n_accepted[0]=0); n_rejected, status: scalars.
 * filt_u: nullable [K,d] un-smoothed (filtering) means at the checkpoints. */
int pn_oracle_solve_save_at(const pn_oracle_config *cfg, const double *u0, const double *params,
                            const double *save_at, int64_t K, double output_scale0, double *u,
                            double *u_std, double *marg_mean, double *marg_chol,
                            int64_t *n_accepted, int64_t *n_rejected, int32_t *status,
                            double *filt_u);

/* The same solve plus stats.log_marginal_likelihood of observations at the checkpoints
 * (src/odecheckpts/train_util.py:22-24): data [K,d], obs_std [K]; *lml = running mean over the K
 * data points of log p(y_k | y_{k+1..K-1}) (probdiffeq's reverse Kalman filter estimator), NaN for the
 * filter strategy and for dense with d > 1.  Optional: cond_out [K, F*N*N + N*Ctot + F*N*N] the
 * backward conditionals (G, g, Lam) checkpoint k -> k-1, scale_out [K,F] the output scale carried by
 * each checkpoint.  PARITY UNPINNED: probdiffeq's sources are absent and the reference commits no
 * likelihood values; tests check it against a brute-force joint Gaussian instead. */
int pn_oracle_solve_save_at_lml(const pn_oracle_config *cfg, const double *u0, const double *params,
                                const double *save_at, int64_t K, double output_scale0,
                                const double *data, const double *obs_std, double *u, double *u_std,
                                double *marg_mean, double *marg_chol, double *cond_out,
                                double *scale_out, double *lml, int32_t *status);

/* Ensemble: members are independent; OpenMP over members (num_threads<=0: all cores).
 * u0: [B,q,d]; params: [B,P]; tol: nullable [B,2] per-member (atol, rtol);
 * outputs member-major: u [B,K,d], u_std [B,K,d], n_accepted [B,K], n_rejected [B], status [B]. */
int pn_oracle_solve_save_at_batch(const pn_oracle_config *cfg, int64_t B, const double *u0,
                                  const double *params, const double *tol, const double *save_at,
                                  int64_t K, const double *output_scale0, double *u, double *u_std,
                                  int64_t *n_accepted, int64_t *n_rejected, int32_t *status,
                                  int num_threads);

/* solve_adaptive_save_every_step (vdp.py:77-79): records every accepted state;
 * the last grid point is t1 exactly (interpolated).  Returns the number of grid
 * points written (<= max_grid), or -1 if max_grid was too small.
 * grid: [max_grid]; u: [max_grid,d]; err_norms: nullable [max_grid] accepted error norms. */
int64_t pn_oracle_solve_save_every_step(const pn_oracle_config *cfg, const double *u0,
                                        const double *params, double t0, double t1,
                                        double output_scale0, int64_t max_grid, double *grid,
                                        double *u, double *u_std, int64_t *n_rejected);

/* solve_fixed_grid (vdp.py:88-91): the same step on a given grid, no error control.
 * u: [G,d]; err_norms: nullable [G] (err_norms[0]=0) the scaled error each step WOULD have had. */
int pn_oracle_solve_fixed_grid(const pn_oracle_config *cfg, const double *u0, const double *params,
                               const double *grid, int64_t G, double output_scale0, double *u,
                               double *u_std, double *err_norms);

#ifdef __cplusplus
}
#endif
#endif
