"""B200-native drop-in for the hot path of pnkraemer/code-adaptive-prob-ode-solvers.

Mirrors the reference's package layout for that path:
  odecheckpts_b200.ivps          <- src/odecheckpts/ivps.py        (IVP zoo, as device functors)
  odecheckpts_b200.ivpsolvers    <- src/odecheckpts/ivpsolvers.py  (solve(...) -> solve_(u0, p))
  odecheckpts_b200.probdiffeq.*  <- the probdiffeq builder vocabulary the reference calls
                                    (ivpsolvers, ivpsolve, impl, taylor, stats)
  odecheckpts_b200.ensemble      <- new: ensemble sharding over the GPUs of one box
All numerics run in the CUDA library (libpn_b200.so) behind the C ABI of include/pn_b200.h.
"""

from . import _cabi, ensemble, ivps, ivpsolvers, probdiffeq  # noqa: F401

__all__ = ["ivps", "ivpsolvers", "probdiffeq", "ensemble"]
