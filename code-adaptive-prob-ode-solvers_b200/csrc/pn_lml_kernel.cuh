// pn_lml_kernel.cuh -- log marginal likelihood of observations at the checkpoints (sm_100a, fp64).
//
// Reference: stats.log_marginal_likelihood(u, standard_deviation=, posterior=)
// (src/odecheckpts/train_util.py:22-24, experiments/old/6_learn_ode/learn.py:112-114): a Kalman
// filter that runs BACKWARDS over the checkpoint Markov sequence the fixed-point smoother left in
// the workspace.  At checkpoint k: observe y_k = (0-th derivative) + N(0, std_k^2) in square-root
// form (QR of [[std, 0], [L^T e_0, L^T]]), take log N(y_k; m_0, s^2), condition on y_k, move to
// checkpoint k-1 through the stored backward conditional.  probdiffeq's estimator keeps a running
// MEAN of the K log-densities; so does the reduction below.
//
// Two kernels: the sweep runs one thread per (member, owned dimension) -- the same virtual members
// as pn_smooth_kernel -- and leaves whitened residuals and log|s| per checkpoint in scratch; the
// reduction adds them per member in the CPU oracle's order (oracle/pn_solver.c: lml_sweep), so the
// result is bit-identical to it.  Thread-per-IVP and lane-per-dimension families.
#pragma once
#include "pn_smooth_kernel.cuh"

namespace pn {

struct LmlArgs {
  long long B, K;
  int dv;       // virtual members per IVP (1, or d for the lane-per-dimension kernels)
  int D;        // mean columns per virtual member
  int per_dim;  // 1: every virtual member has its own factor (blockdiag)
  const double* cond;     // [B*dv][K][SLOT_FIX]
  const int32_t* status;  // [B]
  const double* data;     // [B][K][d]
  const double* obs_std;  // [B][K]
  double* w;     // scratch [B*dv][K][D]: whitened residuals (m_0 - y) / s
  double* logs;  // scratch [B*dv][K]: log|s|
  double* lml;   // [B]
};

template <int N, int D>
__global__ void __launch_bounds__(128) pn_lml_sweep_kernel(const LmlArgs a) {
  using Lay = Layout<N, D>;
  constexpr int SLOT = Lay::SLOT_FIX;
  constexpr int M1 = N + 1;
  const long long vb = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (vb >= a.B * a.dv) return;
  const long long b = vb / a.dv;
  const int cv = (int)(vb - b * a.dv);
  const int dtot = D * a.dv;
  double m[N][D], L[N][N];
  {
    const double* src = a.cond + vb * a.K * SLOT + Lay::BW;
#pragma unroll
    for (int i = 0; i < N; ++i) {
#pragma unroll
      for (int c = 0; c < D; ++c) m[i][c] = src[i * D + c];
#pragma unroll
      for (int j = 0; j < N; ++j) L[i][j] = (j <= i) ? src[N * D + Lay::tri(i, j)] : 0.0;
    }
    marginalise_from_global<N, D>(m, L, a.cond + vb * a.K * SLOT, 1);
  }
  for (long long k = a.K - 1; k >= 0; --k) {
    // observation update: R = qr([[std, 0], [L^T e_0, L^T]])
    double M[M1][M1];
#pragma unroll
    for (int i = 0; i < M1; ++i)
#pragma unroll
      for (int j = 0; j < M1; ++j) M[i][j] = 0.0;
    M[0][0] = a.obs_std[b * a.K + k];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      M[1 + i][0] = (i == 0) ? L[0][0] : 0.0;
#pragma unroll
      for (int j = i; j < N; ++j) M[1 + i][1 + j] = L[j][i];
    }
#pragma unroll
    for (int j = 0; j < M1; ++j) {
      double sigma2 = 0.0;
#pragma unroll
      for (int i = j + 1; i < M1; ++i) sigma2 = fma(M[i][j], M[i][j], sigma2);
      const Reflector rf = make_reflector(M[j][j], sigma2);
#pragma unroll
      for (int c = j + 1; c < M1; ++c) {
        double w = 0.0;
#pragma unroll
        for (int i = j + 1; i < M1; ++i) w = fma(M[i][j], M[i][c], w);
        w = fma(rf.v0, M[j][c], w);
        const double f = w * rf.g;
        M[j][c] = fma(-f, rf.v0, M[j][c]);
#pragma unroll
        for (int i = j + 1; i < M1; ++i) M[i][c] = fma(-f, M[i][j], M[i][c]);
      }
      M[j][j] = rf.beta;
#pragma unroll
      for (int i = j + 1; i < M1; ++i) M[i][j] = 0.0;
    }
    const double s = M[0][0], inv_s = rcp(s);
    a.logs[vb * a.K + k] = det_log(fabs(s));
#pragma unroll
    for (int c = 0; c < D; ++c) {
      const double z = m[0][c] - a.data[(b * a.K + k) * dtot + cv * D + c];
      a.w[(vb * a.K + k) * D + c] = z * inv_s;
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const double gi = M[0][1 + i] * inv_s;
        m[i][c] = fma(-gi, z, m[i][c]);
      }
    }
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) L[i][j] = (j <= i) ? M[1 + j][1 + i] : 0.0;
    if (k == 0) break;
    marginalise_from_global<N, D>(m, L, a.cond + (vb * a.K + k) * SLOT, 1);
  }
}

template <int UNUSED = 0>  // a template only so that the header can be included from every instance TU
__global__ void __launch_bounds__(128) pn_lml_reduce_kernel(const LmlArgs a) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= a.B) return;
  constexpr double HALF_LOG_2PI = 0.91893853320467274178;
  double mean_lp = 0.0, nd = 0.0;
  for (long long k = a.K - 1; k >= 0; --k) {
    double lp = 0.0;
    for (int cv = 0; cv < a.dv; ++cv) {
      const long long vb = b * a.dv + cv;
      for (int c = 0; c < a.D; ++c) {
        const double wv = a.w[(vb * a.K + k) * a.D + c];
        lp = fma(-0.5 * wv, wv, lp);
      }
      if (a.per_dim) lp = fma(-(double)a.D, a.logs[vb * a.K + k] + HALF_LOG_2PI, lp);
    }
    if (!a.per_dim) lp = fma(-(double)(a.dv * a.D), a.logs[(b * a.dv) * a.K + k] + HALF_LOG_2PI, lp);
    mean_lp = fma(mean_lp, nd, lp) * rcp(nd + 1.0);
    nd += 1.0;
  }
  a.lml[b] = (a.status[b] == 0) ? mean_lp : __longlong_as_double(0x7ff8000000000000LL);
}

}  // namespace pn

namespace pn {

// ---- CTA-per-IVP isotropic family: one thread per (member, column) ------------------------------------------------
// Same backward Kalman filter; the n x n factor recursion is shared by the d columns of a member (every thread
// replicates it, as the solver kernel does) and each thread carries its own mean column.  Scratch layout and the
// reduction are those of the lane-per-dimension isotropic kernels (dv = d virtual members with one column each),
// so pn_lml_reduce_kernel adds the terms in the oracle's order.
struct WideLmlArgs {
  long long B, K;
  int d;
  const double* cond;     // [B][K][wslot]
  const double* data;     // [B][K][d]
  const double* obs_std;  // [B][K]
  double* w;     // [B*d][K]
  double* logs;  // [B*d][K] (column 0 of every member is the one that is read)
};

template <int N>
__global__ void __launch_bounds__(128) pn_wide_lml_sweep_kernel(const WideLmlArgs a) {
  using Lay = Layout<N, 1>;
  constexpr int WSLOT = Lay::BW + Lay::NT, M1 = N + 1;
  const long long vb = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (vb >= a.B * a.d) return;
  const long long b = vb / a.d;
  const int c = (int)(vb - b * a.d);
  const long long wslot = (long long)WSLOT + 2LL * N * a.d;
  const double* base = a.cond + (size_t)b * a.K * wslot;
  double m[N], L[N][N], mdummy[N][1];
#pragma unroll
  for (int i = 0; i < N; ++i) mdummy[i][0] = 0.0;
  auto mean_through = [&](const double* slot) {  // m <- G m + g[:, c]
    double mo[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double acc = slot[WSLOT + (size_t)i * a.d + c];
#pragma unroll
      for (int k = 0; k < N; ++k) acc = fma(slot[i * N + k], m[k], acc);
      mo[i] = acc;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) m[i] = mo[i];
  };
  {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      m[i] = base[WSLOT + (size_t)N * a.d + (size_t)i * a.d + c];
#pragma unroll
      for (int j = 0; j < N; ++j) L[i][j] = (j <= i) ? base[Lay::BW + Lay::tri(i, j)] : 0.0;
    }
    mean_through(base);
    marginalise_from_global<N, 1>(mdummy, L, base, 1);
  }
  for (long long k = a.K - 1; k >= 0; --k) {
    double M[M1][M1];
#pragma unroll
    for (int i = 0; i < M1; ++i)
#pragma unroll
      for (int j = 0; j < M1; ++j) M[i][j] = 0.0;
    M[0][0] = a.obs_std[b * a.K + k];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      M[1 + i][0] = (i == 0) ? L[0][0] : 0.0;
#pragma unroll
      for (int j = i; j < N; ++j) M[1 + i][1 + j] = L[j][i];
    }
#pragma unroll
    for (int j = 0; j < M1; ++j) {
      double sigma2 = 0.0;
#pragma unroll
      for (int i = j + 1; i < M1; ++i) sigma2 = fma(M[i][j], M[i][j], sigma2);
      const Reflector rf = make_reflector(M[j][j], sigma2);
#pragma unroll
      for (int cc = j + 1; cc < M1; ++cc) {
        double w = 0.0;
#pragma unroll
        for (int i = j + 1; i < M1; ++i) w = fma(M[i][j], M[i][cc], w);
        w = fma(rf.v0, M[j][cc], w);
        const double f = w * rf.g;
        M[j][cc] = fma(-f, rf.v0, M[j][cc]);
#pragma unroll
        for (int i = j + 1; i < M1; ++i) M[i][cc] = fma(-f, M[i][j], M[i][cc]);
      }
      M[j][j] = rf.beta;
#pragma unroll
      for (int i = j + 1; i < M1; ++i) M[i][j] = 0.0;
    }
    const double sv = M[0][0], inv_s = rcp(sv);
    a.logs[vb * a.K + k] = det_log(fabs(sv));
    {
      const double z = m[0] - a.data[(b * a.K + k) * a.d + c];
      a.w[vb * a.K + k] = z * inv_s;
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const double gi = M[0][1 + i] * inv_s;
        m[i] = fma(-gi, z, m[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) L[i][j] = (j <= i) ? M[1 + j][1 + i] : 0.0;
    if (k == 0) break;
    const double* slot = base + (size_t)k * wslot;
    mean_through(slot);
    marginalise_from_global<N, 1>(mdummy, L, slot, 1);
  }
}

}  // namespace pn
