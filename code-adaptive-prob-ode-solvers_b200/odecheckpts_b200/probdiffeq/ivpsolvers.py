"""``probdiffeq.ivpsolvers``: prior, correction, strategy, solver (construction only).

Reference call sites: src/odecheckpts/ivpsolvers.py:35-50; experiments/1_van_der_pol/vdp.py:63-66;
experiments/5_vs_interpolation/measure.py:44-47.
"""

from typing import NamedTuple


class Prior(NamedTuple):
    num_derivatives: int


class Correction(NamedTuple):
    name: str  # "ts0" | "ts1"
    ode_order: int


class Strategy(NamedTuple):
    name: str  # "filter" | "fixedpoint" | "smoother"
    prior: Prior
    correction: Correction


class InitialCondition(NamedTuple):
    tcoeffs: object
    output_scale: object


class Solver(NamedTuple):
    strategy: Strategy
    calibration: str  # "none" | "dynamic" | "mle"

    def initial_condition(self, tcoeffs, output_scale=1.0):
        """solver.initial_condition (ivpsolvers.py:68): mean = Taylor coefficients, zero covariance,
        identity backward model; assembled on the device at solve time."""
        return InitialCondition(tcoeffs, output_scale)


def prior_ibm(num_derivatives):
    """nu-times integrated Wiener process."""
    nu = int(num_derivatives)
    if nu < 1:
        raise ValueError("num_derivatives must be >= 1")
    return Prior(nu)


def correction_ts0(ode_order=1):
    """EKF0: H = e_q^T."""
    return Correction("ts0", int(ode_order))


def correction_ts1(ode_order=1):
    """EKF1: H = E_q - J_f(m) E_{<q}; dense factorisation only."""
    return Correction("ts1", int(ode_order))


def strategy_filter(prior, correction):
    return Strategy("filter", prior, correction)


def strategy_fixedpoint(prior, correction):
    return Strategy("fixedpoint", prior, correction)


def strategy_smoother(prior, correction):
    """The textbook smoother (src/odecheckpts/ivpsolvers.py:112; experiments/4_brusselator/run.py:103): one
    backward conditional is kept PER ACCEPTED STEP -- O(#steps) memory, the comparator the fixed-point
    strategy replaces.  Served by ``ivpsolve.solve_adaptive_save_every_step`` (which stores the per-step
    conditionals in device memory) and ``stats.offgrid_marginals_searchsorted``."""
    return Strategy("smoother", prior, correction)


def solver(strategy):
    """Uncalibrated solver: the output scale stays at its initial value."""
    return Solver(strategy, "none")


def solver_dynamic(strategy):
    """Dynamic calibration: per-step quasi-MLE of the output scale."""
    return Solver(strategy, "dynamic")


def solver_mle(strategy):
    """Global quasi-MLE calibration: the solve runs with the initial output scale; the running mean of the
    whitened squared residuals z^T S^-1 z / d over the accepted steps is carried as ``solution.output_scale``
    and its final value rescales the posterior covariances (u_std, marginal factors).  No call site in the
    reference; restated from probdiffeq, thread-per-IVP kernels."""
    return Solver(strategy, "mle")
