// pn_smooth_kernel.cuh -- backward marginalisation over the K checkpoints (sm_100a, fp64).
//
// Reference: stats.markov_select_terminal + stats.markov_marginals(reverse=True)
// (src/odecheckpts/ivpsolvers.py:80-81) followed by the QOI selection
// (impl.hidden_model.qoi_from_sample, ivpsolvers.py:89); SURVEY.md App. A.4/A.6:
//   rv_{K-1} = terminal marginal;  rv_{k-1} = ( G_k m_k + g_k , R(QR([ (G_k L_k)^T ; Lam_k^T ]))^T ).
// One thread per member, K-1 sequential marginalisations, all lanes convergent; the
// conditionals are read member-minor ([checkpoint][element][member]) so every load coalesces.
// Filter strategy: the stored marginals are the result; this kernel only gathers them.
#pragma once
#include "pn_scalar_kernel.cuh"

namespace pn {

struct SmoothArgs {
  long long B, K;
  int dv;              // lanes ("virtual members") per IVP: 1, or d for the lane-per-dimension kernels
  int chol_per_dim;    // marg_chol layout: 1 -> [B][K][d][N][N] (blockdiag), 0 -> [B][K][N][N]
  int wide_d;          // wide (CTA-per-IVP) kernels: runtime ODE dimension, else 0
  double* wide_mean;   // wide kernels: per-member mean scratch [B][3][n][d]; dense CTA kernels: per-CTA scratch
  long long wide_ctas; // dense CTA kernels: CTAs the scratch has room for
  const double* cond;  // [B*dv][K][SLOT]
  const double* mle_scale;  // nullable [B]: solver_mle's final factor on the standard deviations / factors
  const int32_t* status;
  double* u;           // [B][K][D]
  double* u_std;       // [B][K][D]
  double* marg_mean;   // nullable [B][K][N][D]
  double* marg_chol;   // nullable [B][K][N][N] (lower triangular, dense storage)
};

// (m, L) <- marginalise((m, L), (G, g, Lam)); cond read from global memory (stride B)
template <int N, int D>
PN_DEV void marginalise_from_global(double (&m)[N][D], double (&L)[N][N], const double* c, long long B) {
  using Lay = Layout<N, D>;
  constexpr int OFF_G = 0, OFF_g = N * N, OFF_LAM = N * N + N * D;
  double G[N][N];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) G[i][j] = c[(long long)(OFF_G + i * N + j) * B];
  double mo[N][D];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int cc = 0; cc < D; ++cc) {
      double acc = c[(long long)(OFF_g + i * D + cc) * B];
#pragma unroll
      for (int k = 0; k < N; ++k) acc = fma(G[i][k], m[k][cc], acc);
      mo[i][cc] = acc;
    }
  double Mt[N][N], Mb[N][N];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double acc = G[i][j] * L[j][j];
#pragma unroll
      for (int k = j + 1; k < N; ++k) acc = fma(G[i][k], L[k][j], acc);
      Mt[j][i] = acc;
    }
#pragma unroll
  for (int j = 0; j < N; ++j)
#pragma unroll
    for (int i = 0; i <= j; ++i) Mb[i][j] = c[(long long)(OFF_LAM + Lay::tri(j, i)) * B];
#pragma unroll
  for (int j = 0; j < N; ++j) {
    double sigma2 = 0.0;
#pragma unroll
    for (int i = j + 1; i < N; ++i) sigma2 = fma(Mt[i][j], Mt[i][j], sigma2);
#pragma unroll
    for (int i = 0; i <= j; ++i) sigma2 = fma(Mb[i][j], Mb[i][j], sigma2);
    Reflector rf = make_reflector(Mt[j][j], sigma2);
#pragma unroll
    for (int cc = j + 1; cc < N; ++cc) {
      double w = 0.0;
#pragma unroll
      for (int i = j + 1; i < N; ++i) w = fma(Mt[i][j], Mt[i][cc], w);
#pragma unroll
      for (int i = 0; i <= j; ++i) w = fma(Mb[i][j], Mb[i][cc], w);
      w = fma(rf.v0, Mt[j][cc], w);
      double f = w * rf.g;
      Mt[j][cc] = fma(-f, rf.v0, Mt[j][cc]);
#pragma unroll
      for (int i = j + 1; i < N; ++i) Mt[i][cc] = fma(-f, Mt[i][j], Mt[i][cc]);
#pragma unroll
      for (int i = 0; i <= j; ++i) Mb[i][cc] = fma(-f, Mb[i][j], Mb[i][cc]);
    }
    Mt[j][j] = rf.beta;
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int cc = 0; cc < D; ++cc) m[i][cc] = mo[i][cc];
#pragma unroll
    for (int j = 0; j <= i; ++j) L[i][j] = Mt[j][i];
  }
}

template <int N, int D, int STRAT>
__global__ void __launch_bounds__(128) pn_smooth_kernel(const SmoothArgs a) {
  using Lay = Layout<N, D>;
  constexpr bool FIX = (STRAT == 1);
  constexpr int SLOT = FIX ? Lay::SLOT_FIX : Lay::SLOT_FILT;
  // one thread per (member, owned dimension): vb = b * dv + cv; workspace stride = B * dv
  const long long vb = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long VB = a.B * a.dv;
  if (vb >= VB) return;
  const long long b = vb / a.dv;
  const int cv = (int)(vb - b * a.dv);
  const int dtot = D * a.dv;
  const bool ok = (a.status[b] == 0);
  double m[N][D], L[N][N];
  auto load_marg = [&](const double* src) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
#pragma unroll
      for (int c = 0; c < D; ++c) m[i][c] = src[i * D + c];
#pragma unroll
      for (int j = 0; j <= i; ++j) L[i][j] = src[N * D + Lay::tri(i, j)];
    }
  };
  if (FIX) {
    // terminal marginal = marginalise((m1, L1), bw_1t), both stored in slot 0
    load_marg(a.cond + vb * a.K * SLOT + Lay::BW);
    marginalise_from_global<N, D>(m, L, a.cond + vb * a.K * SLOT, 1);
  }
  const double nanv = __longlong_as_double(0x7ff8000000000000LL);
  const double sc = a.mle_scale ? a.mle_scale[b] : 1.0;
  for (long long k = a.K - 1; k >= 0; --k) {
    if (!FIX) load_marg(a.cond + (vb * a.K + k) * SLOT);
    const double sd = sc * dsqrt(fma(L[0][0], L[0][0], 0.0));
#pragma unroll
    for (int c = 0; c < D; ++c) {
      a.u[(b * a.K + k) * dtot + cv * D + c] = ok ? m[0][c] : nanv;
      a.u_std[(b * a.K + k) * dtot + cv * D + c] = ok ? sd : nanv;
    }
    if (a.marg_mean) {
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int c = 0; c < D; ++c) a.marg_mean[((b * a.K + k) * N + i) * dtot + cv * D + c] = ok ? m[i][c] : nanv;
    }
    if (a.marg_chol && (a.chol_per_dim || cv == 0)) {
      const long long blk = a.chol_per_dim ? ((b * a.K + k) * a.dv + cv) : (b * a.K + k);
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) a.marg_chol[(blk * N + i) * N + j] = ok ? ((j <= i) ? sc * L[i][j] : 0.0) : nanv;
    }
    if (k == 0) break;
    if (FIX) marginalise_from_global<N, D>(m, L, a.cond + (vb * a.K + k) * SLOT, 1);
  }
}

// ---- wide (CTA-per-IVP) variant: n x n factor per thread (replicated), n x d means in global ---
struct WideSmoothArgs {
  long long B, K;
  int d;
  const double* cond;   // [B][K][wslot]
  double* mean;         // [B][3][n][d] scratch: the first [n][d] block is reused as the running mean
  const int32_t* status;
  double* u;            // [B][K][d]
  double* u_std;        // [B][K][d]
  double* marg_mean;    // nullable [B][K][n][d]
  double* marg_chol;    // nullable [B][K][n][n]
};

template <int N, int STRAT, int THREADS>
__global__ void __launch_bounds__(THREADS) pn_wide_smooth_kernel(const WideSmoothArgs a) {
  using Lay = Layout<N, 1>;
  constexpr bool FIX = (STRAT == 1);
  constexpr int WSLOT = Lay::BW + Lay::NT;
  constexpr int OFF_G = 0;
  const int tid = threadIdx.x, d = a.d;
  const long long b = blockIdx.x;
  if (b >= a.B) return;
  const long long wslot = (long long)WSLOT + 2LL * N * d;
  const double* base = a.cond + (size_t)b * a.K * wslot;
  double* rm = a.mean + (size_t)b * 3 * N * d;  // running mean [n][d]
  const bool ok = (a.status[b] == 0);
  const double nanv = __longlong_as_double(0x7ff8000000000000LL);
  double mdummy[N][1], L[N][N];
#pragma unroll
  for (int i = 0; i < N; ++i) mdummy[i][0] = 0.0;
  auto load_L1 = [&](const double* slot) {
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j <= i; ++j) L[i][j] = slot[Lay::BW + Lay::tri(i, j)];
  };
  auto marginalise_mean = [&](const double* slot) {  // rm[:, c] <- G rm[:, c] + g[:, c]
    double G[N][N];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) G[i][j] = slot[OFF_G + i * N + j];
    const double* g = slot + WSLOT;
    for (int c = tid; c < d; c += THREADS) {
      double mi[N], mo[N];
#pragma unroll
      for (int i = 0; i < N; ++i) mi[i] = rm[(size_t)i * d + c];
#pragma unroll
      for (int i = 0; i < N; ++i) {
        double acc = g[(size_t)i * d + c];
#pragma unroll
        for (int k = 0; k < N; ++k) acc = fma(G[i][k], mi[k], acc);
        mo[i] = acc;
      }
#pragma unroll
      for (int i = 0; i < N; ++i) rm[(size_t)i * d + c] = mo[i];
    }
  };
  if (FIX) {
    for (int e = tid; e < N * d; e += THREADS) rm[e] = base[WSLOT + (size_t)N * d + e];
    load_L1(base);
    __syncthreads();
    marginalise_mean(base);
    marginalise_from_global<N, 1>(mdummy, L, base, 1);
    __syncthreads();
  }
  for (long long k = a.K - 1; k >= 0; --k) {
    const double* slot = base + (size_t)k * wslot;
    if (!FIX) {
      for (int e = tid; e < N * d; e += THREADS) rm[e] = slot[WSLOT + (size_t)N * d + e];
      load_L1(slot);
      __syncthreads();
    }
    const double sd = dsqrt(fma(L[0][0], L[0][0], 0.0));
    for (int c = tid; c < d; c += THREADS) {
      a.u[(b * a.K + k) * d + c] = ok ? rm[c] : nanv;
      a.u_std[(b * a.K + k) * d + c] = ok ? sd : nanv;
    }
    if (a.marg_mean)
      for (int e = tid; e < N * d; e += THREADS) a.marg_mean[(b * a.K + k) * N * d + e] = ok ? rm[e] : nanv;
    if (a.marg_chol && tid == 0) {
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) a.marg_chol[((b * a.K + k) * N + i) * N + j] = ok ? ((j <= i) ? L[i][j] : 0.0) : nanv;
    }
    __syncthreads();
    if (k == 0) break;
    if (FIX) {
      marginalise_mean(slot);
      marginalise_from_global<N, 1>(mdummy, L, slot, 1);
      __syncthreads();
    }
  }
}

}  // namespace pn
