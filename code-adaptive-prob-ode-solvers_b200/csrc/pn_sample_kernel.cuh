// pn_sample_kernel.cuh -- joint samples from the checkpoint Markov sequence (sm_100a, fp64).
//
// Reference: stats.markov_sample(key, posterior, shape=(S,), reverse=True)
// (experiments/5_vs_interpolation/measure.py:69-77): draw the terminal state, then walk the K-1
// backward conditionals  x_{k-1} = G_k x_k + g_k + Lam_k xi  that the fixed-point smoother collapsed
// at the checkpoints.  The conditionals are read from the workspace of a finished solve
// ([member][K][slot], thread-per-IVP and lane-per-dimension families).  One thread per
// (member, owned dimension, sample); all loads of a thread are contiguous.
// Random numbers: Philox4x32-10 (counter = thread / checkpoint / draw, key = seed) + Box-Muller.
// jax.random cannot be reproduced bit for bit, so parity for this path is distributional.
#pragma once
#include "pn_scalar_kernel.cuh"

namespace pn {

struct SampleArgs {
  long long B, K, S;
  int dv;
  int d;  // ODE dimension (the CTA-per-IVP and dense families size their slots from it at run time)
  unsigned long long seed;
  const double* cond;     // [B*dv][K][SLOT]
  const int32_t* status;  // [B]
  double* samples;        // [B][S][K][d]
};

PN_DEV void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0;
    const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c2;
    const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0;
    const unsigned n1 = (unsigned)p1;
    const unsigned n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1;
    const unsigned n3 = (unsigned)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// two standard normals from one Philox block
PN_DEV void normal_pair(unsigned long long seed, unsigned long long tid, unsigned step, unsigned draw, double& z0, double& z1) {
  unsigned r[4];
  philox4x32_10((unsigned)tid, (unsigned)(tid >> 32), step, draw, (unsigned)seed, (unsigned)(seed >> 32), r);
  const double u0 = ((double)(((unsigned long long)r[0] << 21) ^ (r[1] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);  // (0, 1)
  const double u1 = ((double)(((unsigned long long)r[2] << 21) ^ (r[3] >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
  const double rad = sqrt(-2.0 * log(u0));
  double sn, cs;
  sincospi(2.0 * u1, &sn, &cs);
  z0 = rad * cs;
  z1 = rad * sn;
}

template <int N, int D>
__global__ void __launch_bounds__(128) pn_sample_kernel(const SampleArgs a) {
  using Lay = Layout<N, D>;
  constexpr int SLOT = Lay::SLOT_FIX;
  constexpr int OFF_G = 0, OFF_g = N * N, OFF_LAM = N * N + N * D;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = a.B * a.dv * a.S;
  if (gid >= total) return;
  const long long vb = gid / a.S, s = gid - vb * a.S;
  const long long b = vb / a.dv;
  const int cv = (int)(vb - b * a.dv);
  const int dtot = D * a.dv;
  const bool ok = a.status[b] == 0;
  const double* base = a.cond + vb * a.K * SLOT;
  double x[N][D];
  unsigned draw = 0;
  auto gaussian_fill = [&](double (&xi)[N][D], unsigned step) {
    double buf[N * D + 1];
#pragma unroll
    for (int e = 0; e < N * D; e += 2) {
      double z0, z1;
      normal_pair(a.seed, (unsigned long long)gid, step, draw++, z0, z1);
      buf[e] = z0;
      buf[e + 1] = z1;
    }
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int c = 0; c < D; ++c) xi[i][c] = buf[i * D + c];
  };
  // terminal state x1 ~ N(m1, L1 L1^T) (slot 0 carries the accepted state behind the last checkpoint)
  {
    double xi[N][D];
    gaussian_fill(xi, 0xffffffffu);
    const double* mg = base + Lay::BW;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int c = 0; c < D; ++c) {
        double acc = mg[i * D + c];
#pragma unroll
        for (int j = 0; j <= i; ++j) acc = fma(mg[N * D + Lay::tri(i, j)], xi[j][c], acc);
        x[i][c] = acc;
      }
  }
  auto through = [&](const double* cnd, unsigned step) {  // x <- G x + g + Lam xi
    double xi[N][D], xn[N][D];
    gaussian_fill(xi, step);
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int c = 0; c < D; ++c) {
        double acc = cnd[OFF_g + i * D + c];
#pragma unroll
        for (int k = 0; k < N; ++k) acc = fma(cnd[OFF_G + i * N + k], x[k][c], acc);
#pragma unroll
        for (int j = 0; j <= i; ++j) acc = fma(cnd[OFF_LAM + Lay::tri(i, j)], xi[j][c], acc);
        xn[i][c] = acc;
      }
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int c = 0; c < D; ++c) x[i][c] = xn[i][c];
  };
  through(base, 0xfffffffeu);  // accepted state -> last checkpoint
  const double nanv = __longlong_as_double(0x7ff8000000000000LL);
  for (long long k = a.K - 1; k >= 0; --k) {
#pragma unroll
    for (int c = 0; c < D; ++c) a.samples[((b * a.S + s) * a.K + k) * dtot + cv * D + c] = ok ? x[0][c] : nanv;
    if (k == 0) break;
    through(base + k * SLOT, (unsigned)k);
  }
}

}  // namespace pn

namespace pn {

// ---- CTA-per-IVP isotropic family (runtime dimension d): one thread per (member, sample, column) ----------------
// The factor part of a wide slot -- [G | (n unused) | Lam | L1] -- is shared by all d columns of the member;
// the offsets g [n][d] and the terminal mean m1 [n][d] follow it (pn_scalar_kernel.cuh, WIDE = 1).  Columns are
// independent given the factors (the isotropic factorisation is a Kronecker product with I_d), so every thread
// draws its own n-vector per checkpoint.
struct WideSampleArgs {
  long long B, K, S;
  int d;
  unsigned long long seed;
  const double* cond;     // [B][K][wslot]
  const int32_t* status;  // [B]
  double* samples;        // [B][S][K][d]
};

template <int N>
__global__ void __launch_bounds__(128) pn_wide_sample_kernel(const WideSampleArgs a) {
  using Lay = Layout<N, 1>;
  constexpr int WSLOT = Lay::BW + Lay::NT;
  constexpr int OFF_G = 0, OFF_LAM = N * N + N;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = a.B * a.S * a.d;
  if (gid >= total) return;
  const int c = (int)(gid % a.d);
  const long long bs = gid / a.d, b = bs / a.S, s = bs - b * a.S;
  const long long wslot = (long long)WSLOT + 2LL * N * a.d;
  const double* base = a.cond + (size_t)b * a.K * wslot;
  const bool ok = a.status[b] == 0;
  double x[N];
  auto draw = [&](double (&xi)[N], unsigned step) {
    double buf[N + 1];
#pragma unroll
    for (int e = 0; e < N; e += 2) normal_pair(a.seed, (unsigned long long)gid, step, (unsigned)(e >> 1), buf[e], buf[e + 1]);
#pragma unroll
    for (int i = 0; i < N; ++i) xi[i] = buf[i];
  };
  {
    // terminal state: slot 0 carries the accepted state behind the last checkpoint (m1 [n][d], L1)
    double xi[N];
    draw(xi, 0xffffffffu);
    const double* m1 = base + WSLOT + (size_t)N * a.d;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double acc = m1[(size_t)i * a.d + c];
#pragma unroll
      for (int j = 0; j <= i; ++j) acc = fma(base[Lay::BW + Lay::tri(i, j)], xi[j], acc);
      x[i] = acc;
    }
  }
  auto through = [&](const double* slot, unsigned step) {  // x <- G x + g[:, c] + Lam xi
    double xi[N], xn[N];
    draw(xi, step);
    const double* g = slot + WSLOT;
#pragma unroll
    for (int i = 0; i < N; ++i) {
      double acc = g[(size_t)i * a.d + c];
#pragma unroll
      for (int k = 0; k < N; ++k) acc = fma(slot[OFF_G + i * N + k], x[k], acc);
#pragma unroll
      for (int j = 0; j <= i; ++j) acc = fma(slot[OFF_LAM + Lay::tri(i, j)], xi[j], acc);
      xn[i] = acc;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = xn[i];
  };
  through(base, 0xfffffffeu);  // accepted state -> last checkpoint
  const double nanv = __longlong_as_double(0x7ff8000000000000LL);
  for (long long k = a.K - 1; k >= 0; --k) {
    a.samples[((b * a.S + s) * a.K + k) * a.d + c] = ok ? x[0] : nanv;
    if (k == 0) break;
    through(base + (size_t)k * wslot, (unsigned)k);
  }
}

// ---- dense factorisation with d > 1 (all three dense families): one warp per (member, sample) --------------------
// Slot: [G | g | Lam] + [m | L] with D x D matrices, D = n d, derivative-major state index i d + l.  The
// warp-per-IVP and register-column kernels store G, Lam, L row-major; the CTA-per-IVP kernel stores their
// transposes (kernel-native storage, pn_dense_cta_kernel.cuh).  x lives in shared memory; a row-major matrix is
// applied row by row (lanes over the columns + butterfly), a transposed one column by column (lanes over the rows).
struct DenseSampleArgs {
  long long B, K, S;
  int Dn, d;
  int transposed;  // 1: the slot holds G^T, Lam^T, L^T
  unsigned long long seed;
  const double* cond;     // [B][K][2 (Dn^2 + Dn) + Dn^2 ... see slot above]
  const int32_t* status;
  double* samples;        // [B][S][K][d]
};

template <int WARPS>
__global__ void __launch_bounds__(32 * WARPS) pn_dense_sample_kernel(const DenseSampleArgs a) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long bs = (long long)blockIdx.x * WARPS + warp;
  if (bs >= a.B * a.S) return;
  const long long b = bs / a.S, s = bs - b * a.S;
  const int Dn = a.Dn;
  const size_t MAT = (size_t)Dn * Dn, BW = 2 * MAT + Dn, SLOT = BW + Dn + MAT;
  double* x = smem + (size_t)warp * 3 * Dn;
  double* xi = x + Dn;
  double* xn = xi + Dn;
  const double* base = a.cond + (size_t)b * a.K * SLOT;
  const bool ok = a.status[b] == 0;
  auto draw = [&](unsigned step) {
    for (int j = lane; j < Dn; j += 32) {
      double z0, z1;
      normal_pair(a.seed, (unsigned long long)bs, step, (unsigned)(j >> 1), z0, z1);
      xi[j] = (j & 1) ? z1 : z0;
    }
    __syncwarp();
  };
  // xn = off + A x (A full, optional) + T xi (T lower triangular), then x <- xn
  auto affine = [&](const double* off, const double* A, const double* T) {
    if (a.transposed) {
      for (int i = lane; i < Dn; i += 32) {
        double acc = off[i];
        if (A)
          for (int k = 0; k < Dn; ++k) acc = fma(A[(size_t)k * Dn + i], x[k], acc);
        for (int j = 0; j <= i; ++j) acc = fma(T[(size_t)j * Dn + i], xi[j], acc);
        xn[i] = acc;
      }
    } else {
      for (int i = 0; i < Dn; ++i) {
        double acc = 0.0;
        if (A)
          for (int k = lane; k < Dn; k += 32) acc = fma(A[(size_t)i * Dn + k], x[k], acc);
        for (int j = lane; j <= i; j += 32) acc = fma(T[(size_t)i * Dn + j], xi[j], acc);
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) xn[i] = off[i] + acc;
      }
    }
    __syncwarp();
    for (int i = lane; i < Dn; i += 32) x[i] = xn[i];
    __syncwarp();
  };
  draw(0xffffffffu);
  affine(base + BW, nullptr, base + BW + Dn);  // terminal state x1 ~ N(m1, L1 L1^T)
  draw(0xfffffffeu);
  affine(base + MAT, base, base + MAT + Dn);   // accepted state -> last checkpoint
  const double nanv = __longlong_as_double(0x7ff8000000000000LL);
  for (long long k = a.K - 1; k >= 0; --k) {
    for (int l = lane; l < a.d; l += 32) a.samples[((b * a.S + s) * a.K + k) * a.d + l] = ok ? x[l] : nanv;
    if (k == 0) break;
    const double* slot = base + (size_t)k * SLOT;
    draw((unsigned)k);
    affine(slot + MAT, slot, slot + MAT + Dn);
  }
}

}  // namespace pn
