// pn_registry.h -- table of compiled kernel instances (one per problem functor x nu x strategy).
#pragma once
#include <cuda_runtime.h>

#include "pn_scalar_kernel.cuh"
#include "pn_smooth_kernel.cuh"

namespace pn {

enum : int { FAMILY_SCALAR = 0 /* thread per IVP, n x n factors */ };

struct KernelEntry {
  int family, problem, nu, strategy;
  int N, D, Q, P;
  int slot_doubles;   // workspace doubles per (checkpoint, member)
  int smem_doubles;   // dynamic shared memory doubles per thread
  int threads;
  bool has_jac;
  const void* solve_func;
  cudaError_t (*launch_solve)(const SolveArgs&, int grid, size_t smem, cudaStream_t);
  cudaError_t (*launch_smooth)(const SmoothArgs&, cudaStream_t);
};

void register_kernel(const KernelEntry& e);
const KernelEntry* find_kernel(int family, int problem, int nu, int strategy);

template <class Prob, int NU, int STRAT, int THREADS>
struct ScalarInstance {
  static cudaError_t launch_solve(const SolveArgs& a, int grid, size_t smem, cudaStream_t s) {
    pn_scalar_kernel<Prob, NU, STRAT, THREADS><<<grid, THREADS, smem, s>>>(a);
    return cudaGetLastError();
  }
  static cudaError_t launch_smooth(const SmoothArgs& a, cudaStream_t s) {
    int grid = (int)((a.B + 127) / 128);
    pn_smooth_kernel<NU + 1, Prob::D, STRAT><<<grid, 128, 0, s>>>(a);
    return cudaGetLastError();
  }
  static KernelEntry entry() {
    using Lay = Layout<NU + 1, Prob::D>;
    KernelEntry e;
    e.family = FAMILY_SCALAR;
    e.problem = Prob::ID;
    e.nu = NU;
    e.strategy = STRAT;
    e.N = NU + 1;
    e.D = Prob::D;
    e.Q = Prob::Q;
    e.P = Prob::P;
    e.slot_doubles = (STRAT == 1) ? Lay::SLOT_FIX : Lay::SLOT_FILT;
    e.smem_doubles = ((STRAT == 1) ? Lay::BW : 0) + Lay::PEND;
    e.threads = THREADS;
    e.has_jac = Prob::HAS_JAC;
    e.solve_func = (const void*)&pn_scalar_kernel<Prob, NU, STRAT, THREADS>;
    e.launch_solve = &launch_solve;
    e.launch_smooth = &launch_smooth;
    return e;
  }
};

struct Registrar {
  explicit Registrar(const KernelEntry& e) { register_kernel(e); }
};

#define PN_CAT2(a, b) a##b
#define PN_CAT(a, b) PN_CAT2(a, b)
#define PN_REGISTER_SCALAR(Prob, NU, STRAT) \
  static ::pn::Registrar PN_CAT(pn_reg_, __COUNTER__)(::pn::ScalarInstance<::pn::Prob, NU, STRAT, 128>::entry())

}  // namespace pn
