"""``taylor.odejet_padded_scan`` / ``odejet_unroll`` (src/odecheckpts/ivpsolvers.py:65-67,
experiments/4_brusselator/run.py:64).

Taylor-mode initialisation runs INSIDE the solver kernel (one truncated-power-series recurrence
per member, csrc/pn_problems.cuh: taylor_init), so these functions only record what to
initialise from; nothing is computed on the host.
"""

from typing import NamedTuple


class TaylorCoefficients(NamedTuple):
    vf: object
    inits: tuple
    num: int  # number of additional derivatives, as in probdiffeq


def odejet_padded_scan(vf, inits, /, num):
    return TaylorCoefficients(vf, tuple(inits), int(num))


def odejet_unroll(vf, inits, /, num):
    return TaylorCoefficients(vf, tuple(inits), int(num))
