"""The slice of probdiffeq's builder API that the reference calls, backed by the CUDA library.

Reference call sites: src/odecheckpts/ivpsolvers.py:10-11,33,42-53,65-81;
experiments/1_van_der_pol/vdp.py:61-91; experiments/4_brusselator/run.py:51-61,82-90,119-129;
experiments/5_vs_interpolation/measure.py:44-68.
"""

from . import impl, ivpsolve, ivpsolvers, stats, taylor  # noqa: F401
