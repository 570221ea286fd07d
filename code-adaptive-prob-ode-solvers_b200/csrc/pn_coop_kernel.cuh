// pn_coop_kernel.cuh -- cooperative solver kernel for SMALL ensembles of scalar ODEs (sm_100a, fp64):
// n = nu+1 lanes per IVP, lane c owns column c of the stacked matrices / row c of the factors.
//
// Same path as pn_scalar_kernel.cuh (ivpsolve.solve_adaptive_save_at, src/odecheckpts/ivpsolvers.py:71-77;
// dense EKF1 on the stiff Van der Pol problem, experiments/1_van_der_pol/vdp.py:61-66) for d = 1, but for
// the latency-bound regime: when an ensemble is much smaller than the GPU's resident lanes (BASELINE
// config 1: a single IVP; config 2 strong-scaled over 8 GPUs: 8,192 members per GPU), the thread-per-IVP
// kernel leaves every scheduler with one lone warp that executes ~2.4k instructions per attempted step
// one after the other (4.8 us per attempt).  Here the n columns of every Householder QR, the rows of the
// products of the conditional algebra and the n right-hand sides of the triangular solve are spread
// over n lanes, so one attempted step costs every lane a quarter of the instructions and the step's
// time approaches its serial chain (12 reflectors: norm -> sqrt -> reciprocal -> inner products).
// Lanes exchange through a few hundred bytes of shared memory per IVP (reflector broadcast, factor
// gathers) with __syncwarp; floor(32 / n) IVPs share a warp and the warp runs the same straight-line
// "uber step" as the thread-per-IVP kernel (attempt / checkpoint prediction A / B).
//
// Every matrix element is produced by ONE lane with exactly the fma chain of pn_scalar_kernel.cuh and
// oracle/pn_solver.c, so results are bit-identical to both; the workspace slots have the thread-per-IVP
// layout, so the smoothing, sampling and likelihood kernels are shared.
#pragma once
#include "pn_scalar_kernel.cuh"

namespace pn {

template <int N>
struct CoopLayout {
  static constexpr int G = 32 / N;          // IVPs per warp
  static constexpr int XL = N * N;          // gathered factor rows (L_p, then R_Y)
  static constexpr int XV = 2 * (2 * N + 2);  // reflector broadcast, double buffered: v[2N], v0, g
  static constexpr int XC = 2 * N * N + N;  // new conditional: Gn, Ln rows + gn
  static constexpr int XM = N + 2;          // state mean gather + ticket / flags
  static constexpr int PER_GROUP = XL + XV + XC + XM;
  // parked state per LANE ([element][thread]): mean, factor row, conditional rows (G, g, Lam), pending (mean, row)
  static constexpr int STATE = 1 + N + (N + 1 + N) + (1 + N);
};

template <class Prob, int NU, int STRAT, int THREADS>
__global__ void __launch_bounds__(THREADS, 3) pn_coop_kernel(const __grid_constant__ SolveArgs a) {
  static_assert(Prob::D == 1, "cooperative kernel: scalar ODEs (dense / isotropic factorisation with d = 1)");
  constexpr int N = NU + 1, Q = Prob::Q, P = (Prob::P > 0 ? Prob::P : 1);
  constexpr bool FIX = (STRAT == 1);
  using Lay = Layout<N, 1>;
  using CL = CoopLayout<N>;
  constexpr int G = CL::G;
  constexpr int SLOT = FIX ? Lay::SLOT_FIX : Lay::SLOT_FILT;
  constexpr int OFF_G = 0, OFF_g = N * N, OFF_LAM = N * N + N;
  constexpr double TIME_EPS = 10.0 * 2.220446049250313e-16;
  constexpr int WARPS = THREADS / 32;

  extern __shared__ double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = lane / N;                 // group inside the warp (>= G: idle lanes)
  const int c = lane - grp * N;             // owned column / row
  const bool real = grp < G;
  double* xg = smem + ((size_t)warp * G + (real ? grp : 0)) * CL::PER_GROUP;  // idle lanes alias group 0 (never write)
  double* xL = xg;
  double* xV = xL + CL::XL;
  double* xC = xV + CL::XV;
  double* xM = xC + CL::XC;
  double* st = smem + (size_t)WARPS * G * CL::PER_GROUP;  // [STATE][THREADS]
#define ST(e) st[(e) * THREADS + tid]
  // state elements
  constexpr int S_M = 0, S_L = 1, S_G = 1 + N, S_g = 1 + 2 * N, S_LAM = 2 + 2 * N, S_PM = 2 + 3 * N, S_PL = 3 + 3 * N;

  const double* LQ = a.lq;
  const double inv_sqrt_d = 1.0;

  bool have = false, exhausted = false;
  long long b = 0;
  double par[P];
  double atol = a.atol, rtol = a.rtol, sigma0 = 1.0;
  double t = 0.0, dt_next = a.dt0, le_prev = 0.0, sigma_state = 1.0, pend_t = 0.0, pend_sigma = 1.0;
  int mode = MODE_STEP;
  long long k_next = 1, n_acc = 0, n_rej = 0, n_att = 0;
#pragma unroll
  for (int i = 0; i < P; ++i) par[i] = 0.0;
#pragma unroll
  for (int e = 0; e < CL::STATE; ++e) ST(e) = 0.0;

  for (;;) {
    // ---- fetch a member per idle group (convergent) ------------------------------------------------------
    if (real && c == 0 && !have && !exhausted) {
      const unsigned long long tk = atomicAdd(a.ticket, 1ULL);
      xM[N] = __longlong_as_double((long long)tk);
    }
    __syncwarp();
    if (real && !have && !exhausted) {
      const unsigned long long tk = (unsigned long long)__double_as_longlong(xM[N]);
      if (tk < (unsigned long long)a.B) {
        b = a.order ? a.order[tk] : (long long)tk;
        have = true;
        double u0[Q];
#pragma unroll
        for (int i = 0; i < Q; ++i) u0[i] = a.u0[b * Q + i];
#pragma unroll
        for (int i = 0; i < P; ++i) par[i] = (i < a.num_params) ? a.params[b * a.num_params + i] : 0.0;
        atol = a.tol ? a.tol[2 * b] : a.atol;
        rtol = a.tol ? a.tol[2 * b + 1] : a.rtol;
        sigma0 = a.sigma0 ? a.sigma0[b] : 1.0;
        double tc[N][1];
        taylor_init<Prob, NU>(u0, par, tc);
        double mine = tc[0][0];
#pragma unroll
        for (int i = 1; i < N; ++i) mine = (c == i) ? tc[i][0] : mine;
        ST(S_M) = mine;
#pragma unroll
        for (int j = 0; j < N; ++j) {
          ST(S_L + j) = 0.0;
          ST(S_G + j) = (j == c) ? 1.0 : 0.0;
          ST(S_LAM + j) = 0.0;
        }
        ST(S_g) = 0.0;
        t = a.save_at[0];
        dt_next = a.dt0;
        le_prev = 0.0;
        sigma_state = sigma0;
        mode = MODE_STEP;
        k_next = 1;
        n_acc = n_rej = n_att = 0;
        if (c == 0) {
          a.n_accepted[b * a.K] = 0;
          if (a.out_scale) a.out_scale[b * a.K] = sigma0;
          if (a.flags & FLAG_RECORD) {
            a.traj_t[b] = t;
            a.traj_u[b] = mine;
            a.traj_std[b] = 0.0;
          }
        }
        if (!FIX) {  // filter: slot 0 holds the initial marginal
          double* s0 = a.cond + (long long)b * a.K * SLOT;
          s0[c] = mine;
#pragma unroll
          for (int j = 0; j < N; ++j)
            if (j <= c) s0[N + Lay::tri(c, j)] = 0.0;
        }
      } else {
        exhausted = true;
      }
    }
    if (!__any_sync(0xffffffffu, have)) break;

    // ---- choose this iteration's prediction ---------------------------------------------------------------
    const double t_ck = a.save_at[k_next < a.K ? k_next : a.K - 1];
    double dt, sigma_given;
    if (mode == MODE_STEP) {
      dt = (a.flags & FLAG_FIXED_GRID) ? (t_ck - t) : dt_next;
      sigma_given = sigma0;
    } else if (mode == MODE_INTERP_A) {
      dt = t_ck - t;
      sigma_given = pend_sigma;
    } else {
      dt = pend_t - t;
      sigma_given = pend_sigma;
    }
    if (!have) dt = 1.0;  // idle lanes run the step on harmless operands

    // ==================== uber step (straight-line for the whole warp) =======================================
    double p[N], pinv[N];
    {
      const double adt = fabs(dt);
      const double sq = dsqrt(adt);
      const double isq = rcp(sq), idt = rcp(adt);
      double dtp = 1.0, idtp = 1.0;
#pragma unroll
      for (int k = 0; k <= NU; ++k) {
        const int i = NU - k;
        p[i] = (sq * dtp) * (1.0 / factorial(k));
        pinv[i] = (isq * idtp) * factorial(k);
        dtp *= adt;
        idtp *= idt;
      }
    }
    double p_c = p[0], pinv_c = pinv[0];
#pragma unroll
    for (int i = 1; i < N; ++i) {
      p_c = (c == i) ? p[i] : p_c;
      pinv_c = (c == i) ? pinv[i] : pinv_c;
    }
    // gather the state mean and the preconditioned factor rows
    double Lrow[N];
#pragma unroll
    for (int j = 0; j < N; ++j) Lrow[j] = ST(S_L + j);
    if (real) {
      xM[c] = ST(S_M);
#pragma unroll
      for (int j = 0; j < N; ++j) xL[c * N + j] = (j <= c) ? pinv_c * Lrow[j] : 0.0;  // L_p[c][j]
    }
    __syncwarp();
    double m_p[N], m_ext_p[N], m_ext[N];
#pragma unroll
    for (int i = 0; i < N; ++i) m_p[i] = pinv[i] * xM[i];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double acc = m_p[i];
#pragma unroll
      for (int j = i + 1; j < N; ++j) acc = fma(Binom<N>::at(i, j), m_p[j], acc);
      m_ext_p[i] = acc;
      m_ext[i] = p[i] * acc;
    }
    double z, h[Q + 1];
    {
      double uarg[Q], f[1];
#pragma unroll
      for (int k = 0; k < Q; ++k) uarg[k] = m_ext[k];
      Prob::vf(uarg, par, f);
      z = m_ext[Q] - f[0];
#pragma unroll
      for (int k = 0; k < Q; ++k) h[k] = 0.0;
      h[Q] = 1.0;
      if (Prob::HAS_JAC && a.correction == 1) {
        double J[Q];
        Prob::jac(uarg, par, J);
#pragma unroll
        for (int k = 0; k < Q; ++k) h[k] = -J[k];
      }
    }
    double err, sigma;
    {
      double s2 = 0.0;
#pragma unroll
      for (int j = 0; j <= Q; ++j) {
        double acc = 0.0;
#pragma unroll
        for (int i = j; i <= Q; ++i) acc = fma(h[i] * p[i], LQ[i * N + j], acc);
        s2 = fma(acc, acc, s2);
      }
      const double s = dsqrt(s2);
      const double zz = fma(z, z, 0.0);
      const double sigma_hat = (dsqrt(zz) * rcp(s)) * inv_sqrt_d;
      err = (fabs(dt) * sigma_hat) * s;
      sigma = (mode == MODE_STEP) ? ((a.calibration == 1) ? sigma_hat : sigma_given) : sigma_given;
    }
    // ---- predict: column c of BL = (A L_p)^T and of BR = L_p^T, block QR over the lanes --------------------
    double BL[N], BR[N], RY[N], R12[N];
#pragma unroll
    for (int r = 0; r < N; ++r) RY[r] = R12[r] = 0.0;
#pragma unroll
    for (int r = 0; r < N; ++r) {
      // BL[r][c] = (A L_p)[c][r] = sum_{k >= max(c, r)} A1[c][k] L_p[k][r]; A1[c][k] depends on the lane
      double acc = 0.0;
      bool first = true;
#pragma unroll
      for (int k = r; k < N; ++k) {
        // binom(c, k) for this lane's c (k >= c contributes)
        double bk = Binom<N>::at(0, k);
#pragma unroll
        for (int cc = 1; cc < N; ++cc) bk = (c == cc) ? Binom<N>::at(cc, k) : bk;
        const double x = xL[k * N + r];
        const bool use = (k >= c);
        const bool diag = (k == c);
        const double term_first = diag ? x : bk * x;  // (k0 == i) ? L_p[i][j] : binom * L_p[k0][j]
        const double term_next = fma(bk, x, acc);
        acc = use ? (first ? term_first : term_next) : acc;
        first = use ? false : first;
      }
      BL[r] = acc;
      BR[r] = (r <= c) ? xL[c * N + r] : 0.0;  // BR[r][c] = L_p[c][r]
    }
    // LQ entries of this lane: LQ[c][j] (row c)
    double lq_c[N];
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double v = LQ[j];
#pragma unroll
      for (int cc = 1; cc < N; ++cc) v = (c == cc) ? LQ[cc * N + j] : v;
      lq_c[j] = v;
    }
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double* buf = xV + (j & 1) * (2 * N + 2);
      if (real && c == j) {
        double sigma2 = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) sigma2 = fma(BL[i], BL[i], sigma2);
        const Reflector rf = make_reflector(sigma * lq_c[j], sigma2);
        RY[j] = rf.beta;
#pragma unroll
        for (int i = 0; i < N; ++i) buf[i] = BL[i];
        buf[2 * N] = rf.v0;
        buf[2 * N + 1] = rf.g;
      }
      __syncwarp();
      double vj[N];
#pragma unroll
      for (int i = 0; i < N; ++i) vj[i] = buf[i];
      const double v0 = buf[2 * N], gg = buf[2 * N + 1];
      {  // left block, column c > j
        const double top = sigma * lq_c[j];
        double w = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) w = fma(vj[i], BL[i], w);
        w = fma(v0, top, w);
        const double f = w * gg;
        const bool on = c > j;
        RY[j] = on ? fma(-f, v0, top) : RY[j];
#pragma unroll
        for (int i = 0; i < N; ++i) BL[i] = on ? fma(-f, vj[i], BL[i]) : BL[i];
      }
      if (FIX) {  // right block, column c
        double w = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
          const double wn = fma(vj[i], BR[i], w);
          w = (j == 0 && i > c) ? w : wn;  // still structurally zero
        }
        const double f = w * gg;
        R12[j] = fma(-f, v0, 0.0);
#pragma unroll
        for (int i = 0; i < N; ++i) BR[i] = fma(-f, vj[i], BR[i]);
      }
    }
    // L_ext row c: L_ext[c][j] = p[c] RY[j][c]; gather R_Y for the substitution and the correction
    double Lext[N];
#pragma unroll
    for (int j = 0; j < N; ++j) Lext[j] = (j <= c) ? p_c * RY[j] : 0.0;
    __syncwarp();  // everybody is done with xL (L_p)
    if (real) {
#pragma unroll
      for (int r = 0; r < N; ++r) xL[r * N + c] = (r <= c) ? RY[r] : 0.0;  // R_Y[r][c]
    }
    __syncwarp();
    double Gn[N], Ln[N], gn = 0.0;
    if (FIX) {
      double X[N];
#pragma unroll
      for (int i = N - 1; i >= 0; --i) {
        const double inv = rcp(xL[i * N + i]);
        double acc = R12[i];
#pragma unroll
        for (int k = i + 1; k < N; ++k) acc = fma(-xL[i * N + k], X[k], acc);
        X[i] = acc * inv;
      }
      double m_p_c = m_p[0];
#pragma unroll
      for (int i = 1; i < N; ++i) m_p_c = (c == i) ? m_p[i] : m_p_c;
      double acc = m_p_c;
#pragma unroll
      for (int k = 0; k < N; ++k) acc = fma(-X[k], m_ext_p[k], acc);
      gn = p_c * acc;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        Gn[j] = (p_c * X[j]) * pinv[j];
        Ln[j] = p_c * BR[j];
      }
      if (real) {
#pragma unroll
        for (int j = 0; j < N; ++j) {
          xC[c * N + j] = Gn[j];
          xC[N * N + c * N + j] = Ln[j];
        }
        xC[2 * N * N + c] = gn;
      }
    }
    // ---- correction (replicated scalars, own row / column) ----------------------------------------------------
    double hL[Q + 1], gain_c, m_new_c, e_norm;
    double Mc[Q + 1];
    {
      double S = 0.0;
#pragma unroll
      for (int j = 0; j <= Q; ++j) {
        double acc = 0.0;
#pragma unroll
        for (int i = j; i <= Q; ++i) acc = fma(h[i], p[i] * xL[j * N + i], acc);  // L_ext[i][j] = p[i] R_Y[j][i]
        hL[j] = acc;
        S = fma(acc, acc, S);
      }
      const double invS = rcp(S);
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j <= Q; ++j) acc = (j <= c) ? fma(Lext[j], hL[j], acc) : acc;
      gain_c = acc * invS;
      const double gain0 = ((p[0] * xL[0]) * hL[0]) * invS;  // fma(L_ext[0][0], hL[0], 0) * invS
      double m_ext_c = m_ext[0];
#pragma unroll
      for (int i = 1; i < N; ++i) m_ext_c = (c == i) ? m_ext[i] : m_ext_c;
      m_new_c = fma(-gain_c, z, m_ext_c);
      const double m_new0 = fma(-gain0, z, m_ext[0]);
      const double ratio = err * rcp(fma(rtol, fabs(m_new0), atol));
      e_norm = dsqrt(fma(ratio, ratio, 0.0)) * inv_sqrt_d;
    }
    double fac, le_now;
    {
      le_now = det_log(e_norm < 2.2250738585072014e-308 ? 2.2250738585072014e-308 : e_norm);
      le_now = (e_norm == 0.0) ? -745.0 : le_now;
      fac = a.safety * det_exp(fma(a.pow_p, le_prev, -((a.pow_i + a.pow_p) * le_now)));
      fac = (e_norm == 0.0) ? a.factor_max : fac;
      fac = (e_norm != e_norm) ? e_norm : fac;
      fac = (fac < a.factor_max) ? fac : a.factor_max;
      fac = (fac > a.factor_min) ? fac : a.factor_min;
    }
    // corrected factor: column c of Mc (rows 0..Q), Q reflectors
    double Lnew[N];
    {
#pragma unroll
      for (int j = 0; j <= Q; ++j) Mc[j] = fma(-hL[j], gain_c, (j <= c) ? Lext[j] : 0.0);
#pragma unroll
      for (int c0 = 0; c0 < Q; ++c0) {
        double* buf = xV + ((N + c0) & 1) * (2 * N + 2);
        __syncwarp();
        if (real && c == c0) {
          double sigma2 = 0.0;
#pragma unroll
          for (int i = c0 + 1; i <= Q; ++i) sigma2 = fma(Mc[i], Mc[i], sigma2);
          const Reflector rf = make_reflector(Mc[c0], sigma2);
#pragma unroll
          for (int i = c0 + 1; i <= Q; ++i) buf[i] = Mc[i];
          buf[2 * N] = rf.v0;
          buf[2 * N + 1] = rf.g;
          Mc[c0] = rf.beta;
        }
        __syncwarp();
        const double v0 = buf[2 * N], gg = buf[2 * N + 1];
        double w = 0.0;
#pragma unroll
        for (int i = c0 + 1; i <= Q; ++i) w = fma(buf[i], Mc[i], w);
        w = fma(v0, Mc[c0], w);
        const double f = w * gg;
        const bool on = c > c0;
        Mc[c0] = on ? fma(-f, v0, Mc[c0]) : Mc[c0];
#pragma unroll
        for (int i = c0 + 1; i <= Q; ++i) Mc[i] = on ? fma(-f, buf[i], Mc[i]) : Mc[i];
      }
#pragma unroll
      for (int j = 0; j < N; ++j) Lnew[j] = (j <= Q) ? ((j <= c) ? Mc[j] : 0.0) : Lext[j];
    }
    // ---- merge with the running conditional (A.4): row c of the products, column c of the stacked matrix ----
    double Gm[N], gm = 0.0, Lm[N];
    if (FIX) {
      __syncwarp();  // xC (Gn, Ln, gn) complete
      double G1[N];
#pragma unroll
      for (int k = 0; k < N; ++k) G1[k] = ST(S_G + k);
#pragma unroll
      for (int j = 0; j < N; ++j) {
        double acc = G1[0] * xC[j];
#pragma unroll
        for (int k = 1; k < N; ++k) acc = fma(G1[k], xC[k * N + j], acc);
        Gm[j] = acc;
      }
      {
        double acc = ST(S_g);
#pragma unroll
        for (int k = 0; k < N; ++k) acc = fma(G1[k], xC[2 * N * N + k], acc);
        gm = acc;
      }
      double Mt[N], Mb[N];  // column c: Mt[r] = T[c][r], Mb[r] = Lam_run[c][r] (r <= c)
#pragma unroll
      for (int j = 0; j < N; ++j) {
        double acc = G1[0] * xC[N * N + j];
#pragma unroll
        for (int k = 1; k < N; ++k) acc = fma(G1[k], xC[N * N + k * N + j], acc);
        Mt[j] = acc;
        Mb[j] = (j <= c) ? ST(S_LAM + j) : 0.0;
      }
#pragma unroll
      for (int j = 0; j < N; ++j) {
        double* buf = xV + ((N + Q + j) & 1) * (2 * N + 2);
        __syncwarp();
        if (real && c == j) {
          double sigma2 = 0.0;
#pragma unroll
          for (int i = j + 1; i < N; ++i) sigma2 = fma(Mt[i], Mt[i], sigma2);
#pragma unroll
          for (int i = 0; i <= j; ++i) sigma2 = fma(Mb[i], Mb[i], sigma2);
          const Reflector rf = make_reflector(Mt[j], sigma2);
#pragma unroll
          for (int i = 0; i < N; ++i) {
            buf[i] = Mt[i];
            buf[N + i] = Mb[i];
          }
          buf[2 * N] = rf.v0;
          buf[2 * N + 1] = rf.g;
          Mt[j] = rf.beta;
        }
        __syncwarp();
        const double v0 = buf[2 * N], gg = buf[2 * N + 1];
        double w = 0.0;
#pragma unroll
        for (int i = j + 1; i < N; ++i) w = fma(buf[i], Mt[i], w);
#pragma unroll
        for (int i = 0; i <= j; ++i) w = fma(buf[N + i], Mb[i], w);
        w = fma(v0, Mt[j], w);
        const double f = w * gg;
        const bool on = c > j;
        Mt[j] = on ? fma(-f, v0, Mt[j]) : Mt[j];
#pragma unroll
        for (int i = j + 1; i < N; ++i) Mt[i] = on ? fma(-f, buf[i], Mt[i]) : Mt[i];
#pragma unroll
        for (int i = 0; i <= j; ++i) Mb[i] = on ? fma(-f, buf[N + i], Mb[i]) : Mb[i];
      }
#pragma unroll
      for (int j = 0; j < N; ++j) Lm[j] = (j <= c) ? Mt[j] : 0.0;  // Lm[c][j] = R[j][c]
    }
    __syncwarp();  // exchange buffers are free again before the next iteration writes them

    // ==================== per-group bookkeeping (may diverge between groups) =====================================
    if (have) {
      double* base = a.cond + (long long)b * a.K * SLOT;
      auto store_cond_rows = [&](double* dst, const double* Grow, double gval, const double* Lamrow) {
#pragma unroll
        for (int j = 0; j < N; ++j) dst[OFF_G + c * N + j] = Grow[j];
        dst[OFF_g + c] = gval;
#pragma unroll
        for (int j = 0; j < N; ++j)
          if (j <= c) dst[OFF_LAM + Lay::tri(c, j)] = Lamrow[j];
      };
      auto store_marg_row = [&](double* dst, double mval, const double* Lr) {
        dst[c] = mval;
#pragma unroll
        for (int j = 0; j < N; ++j)
          if (j <= c) dst[N + Lay::tri(c, j)] = Lr[j];
      };
      auto bw_reset = [&]() {
#pragma unroll
        for (int j = 0; j < N; ++j) {
          ST(S_G + j) = (j == c) ? 1.0 : 0.0;
          ST(S_LAM + j) = 0.0;
        }
        ST(S_g) = 0.0;
      };
      auto bw_commit = [&]() {
#pragma unroll
        for (int j = 0; j < N; ++j) {
          ST(S_G + j) = Gm[j];
          ST(S_LAM + j) = Lm[j];
        }
        ST(S_g) = gm;
      };
      auto ck_time = [&](long long k) { return a.save_at[k < a.K ? k : a.K - 1]; };
      auto resolve_hits = [&](bool& fin) {
        while (k_next < a.K && !(t + TIME_EPS < ck_time(k_next))) {
          double* slot = base + (long long)k_next * SLOT;
          double Lr[N], Gr[N], Lamr[N];
#pragma unroll
          for (int j = 0; j < N; ++j) {
            Lr[j] = ST(S_L + j);
            Gr[j] = ST(S_G + j);
            Lamr[j] = ST(S_LAM + j);
          }
          if (FIX) {
            store_cond_rows(slot, Gr, ST(S_g), Lamr);
            if (k_next == a.K - 1) {
              double Gi[N], Li[N];
#pragma unroll
              for (int j = 0; j < N; ++j) {
                Gi[j] = (j == c) ? 1.0 : 0.0;
                Li[j] = 0.0;
              }
              store_cond_rows(base, Gi, 0.0, Li);
              store_marg_row(base + Lay::BW, ST(S_M), Lr);
            }
            bw_reset();
          } else {
            store_marg_row(slot, ST(S_M), Lr);
          }
          if (c == 0) {
            a.n_accepted[b * a.K + k_next] = n_acc;
            if (a.out_scale) a.out_scale[b * a.K + k_next] = sigma_state;
          }
          k_next += 1;
        }
        if (k_next >= a.K) fin = true;
      };
      auto after_checkpoint = [&](bool& fin) {
        if (k_next < a.K && pend_t > ck_time(k_next) + TIME_EPS) {
          mode = MODE_INTERP_A;
        } else {
          t = pend_t;
          sigma_state = pend_sigma;
          ST(S_M) = ST(S_PM);
#pragma unroll
          for (int j = 0; j < N; ++j) ST(S_L + j) = ST(S_PL + j);
          if (FIX) bw_commit();
          mode = MODE_STEP;
          resolve_hits(fin);
        }
      };
      bool finished = false;
      int status = 0;
      const bool fixed_grid = (a.flags & FLAG_FIXED_GRID) != 0;
      if (mode == MODE_STEP) {
        n_att += 1;
        if (e_norm != e_norm && !fixed_grid) {
          finished = true;
          status = 1;
        } else {
          dt_next = fac * dt;
          if (e_norm <= 1.0 || fixed_grid) {
            if (!fixed_grid) le_prev = le_now;
            n_acc += 1;
            const double t1 = fixed_grid ? t_ck : (t + dt);
            const bool overshoot = (k_next < a.K) && (t1 > t_ck + TIME_EPS);
            if (overshoot) {
              pend_t = t1;
              pend_sigma = sigma;
              ST(S_PM) = m_new_c;
#pragma unroll
              for (int j = 0; j < N; ++j) ST(S_PL + j) = Lnew[j];
              mode = MODE_INTERP_A;
            } else {
              t = t1;
              sigma_state = sigma;
              ST(S_M) = m_new_c;
#pragma unroll
              for (int j = 0; j < N; ++j) ST(S_L + j) = Lnew[j];
              if (FIX) bw_commit();
              if ((a.flags & FLAG_RECORD) && n_acc < a.traj_cap && c == 0) {
                a.traj_t[n_acc * a.B + b] = t1;
                a.traj_u[n_acc * a.B + b] = m_new_c;
                a.traj_std[n_acc * a.B + b] = dsqrt(fma(Lnew[0], Lnew[0], 0.0));
              }
              resolve_hits(finished);
            }
          } else {
            n_rej += 1;
          }
          if (!finished && mode == MODE_STEP && a.max_attempts > 0 && n_att >= a.max_attempts) {
            finished = true;
            status = 2;
          }
        }
      } else if (mode == MODE_INTERP_A) {
        double* slot = base + (long long)k_next * SLOT;
        double m_ext_c = m_ext[0];
#pragma unroll
        for (int i = 1; i < N; ++i) m_ext_c = (c == i) ? m_ext[i] : m_ext_c;
        if (FIX) {
          store_cond_rows(slot, Gm, gm, Lm);
          bw_reset();
        } else {
          store_marg_row(slot, m_ext_c, Lext);
          if ((a.flags & FLAG_RECORD) && n_acc < a.traj_cap && c == 0) {
            a.traj_t[n_acc * a.B + b] = t_ck;
            a.traj_u[n_acc * a.B + b] = m_ext_c;
            a.traj_std[n_acc * a.B + b] = dsqrt(fma(Lext[0], Lext[0], 0.0));
          }
        }
        t = t_ck;
        ST(S_M) = m_ext_c;
#pragma unroll
        for (int j = 0; j < N; ++j) ST(S_L + j) = Lext[j];
        if (c == 0) {
          a.n_accepted[b * a.K + k_next] = n_acc;
          if (a.out_scale) a.out_scale[b * a.K + k_next] = sigma_given;
        }
        if (FIX) {
          mode = MODE_INTERP_B;
        } else {
          k_next += 1;
          after_checkpoint(finished);
        }
      } else {
        if (k_next == a.K - 1) {
          store_cond_rows(base, Gm, gm, Lm);
          double Pl[N];
#pragma unroll
          for (int j = 0; j < N; ++j) Pl[j] = ST(S_PL + j);
          store_marg_row(base + Lay::BW, ST(S_PM), Pl);
        }
        k_next += 1;
        after_checkpoint(finished);
      }
      if (finished) {
        if (c == 0) {
          a.n_rejected[b] = n_rej;
          a.status[b] = status;
          if (status != 0)
            for (long long kk = k_next; kk < a.K; ++kk) a.n_accepted[b * a.K + kk] = n_acc;
          if (a.flags & FLAG_RECORD) a.traj_len[b] = (n_acc + 1 < a.traj_cap) ? (n_acc + 1) : a.traj_cap;
        }
        have = false;
      }
    }
  }
#undef ST
}

}  // namespace pn
