"""Solution routines for initial value problems (mirror of src/odecheckpts/ivpsolvers.py).

``solve`` keeps the reference's signature and conventions (ivpsolvers.py:14-26,55-57,88-91):
``solve(method, vf, u0_like, save_at=..., dt0=..., atol=..., rtol=..., ode_order=1,
calibrate="dynamic") -> solve_(u0: tuple, p, output_scale=1.0) -> (means[K, d], aux)``.
Two extensions, both keyword-only: ``factorisation=`` / ``correction=`` (hard-coded to
"isotropic" / "ts0" in the reference, ivpsolvers.py:32,36-39), and ensembles -- ``u0`` arrays
may carry a leading ensemble axis ``[B, d]`` and parameters may be arrays ``[B]``; the result is
then ``[B, K, d]``.
"""

import numpy as np

from .probdiffeq import ivpsolve, ivpsolvers, stats, taylor


def solve(method: str, vf, u0_like, /, save_at, *, dt0, atol, rtol, ode_order=1, calibrate="dynamic",
          factorisation="isotropic", correction=None, return_marginals=True, tol=None, device=None):  # fmt: skip
    # Select a state-space model (explicitly, instead of the reference's global impl.select)
    num_derivatives = int(method[-1])
    kind = method[:3] if correction is None else correction
    if kind == "ts0":
        corr = ivpsolvers.correction_ts0(ode_order=ode_order)
    elif kind == "ts1" and factorisation == "dense":
        corr = ivpsolvers.correction_ts1(ode_order=ode_order)
    else:
        raise ValueError

    # Build a solver
    ibm = ivpsolvers.prior_ibm(num_derivatives=num_derivatives)
    strategy = ivpsolvers.strategy_fixedpoint(ibm, corr)
    if calibrate == "dynamic":
        solver = ivpsolvers.solver_dynamic(strategy)
    elif calibrate == "none":
        solver = ivpsolvers.solver(strategy)
    else:
        raise ValueError
    control = ivpsolve.control_proportional_integral()
    asolver = ivpsolve.adaptive(solver, atol=atol, rtol=rtol, control=control)

    def solve_(u0: tuple, p, output_scale=1.0):
        if not isinstance(u0, tuple):
            raise ValueError("Tuple expected.")

        def vf_wrapped(*y, t):
            return vf(*y, t=t, p=p)

        tcoeffs = taylor.odejet_padded_scan(vf_wrapped, u0, num=num_derivatives + 1 - ode_order)
        init = solver.initial_condition(tcoeffs, output_scale=output_scale)
        sol = ivpsolve.solve_adaptive_save_at(
            vf_wrapped, init, save_at=save_at, dt0=dt0, adaptive_solver=asolver,
            factorisation=factorisation, return_marginals=return_marginals, tol=tol, device=device,
        )  # fmt: skip
        # sol.u already is the backward-marginalised (smoothed) mean at every checkpoint, row 0 the
        # initial value and the last row the terminal marginal (ivpsolvers.py:80-89)
        aux = {"solution": sol, "u0_solve": sol.u}
        return sol.u, aux

    return solve_


def solve_via_interpolate(method: str, vf, u0_like, /, save_at, *, dt0, atol, rtol, device=None):
    """The reference's "textbook" comparator (src/odecheckpts/ivpsolvers.py:94-148), same construction:
    ``strategy_smoother`` + ``solve_adaptive_save_every_step`` on [save_at[0] - 1e-6, save_at[-1] + 1e-6], then
    ``stats.offgrid_marginals_searchsorted`` at ``save_at``.  Memory grows with the number of accepted steps
    (one backward conditional per step in device memory): this is the O(#steps) route the checkpoint solver
    ``solve`` replaces, kept so that the paper's memory / run-time comparison can be regenerated.  The two
    routes evaluate the same posterior; they agree to rounding (tests/test_gpu_api.py)."""
    small_value = 1e-6
    num_derivatives = int(method[-1])
    if method[:3] == "ts0":
        correction = ivpsolvers.correction_ts0()
    else:
        raise ValueError
    ibm = ivpsolvers.prior_ibm(num_derivatives=num_derivatives)
    strategy = ivpsolvers.strategy_smoother(ibm, correction)
    solver = ivpsolvers.solver_dynamic(strategy)
    control = ivpsolve.control_proportional_integral()
    asolver = ivpsolve.adaptive(solver, atol=atol, rtol=rtol, control=control)

    def solve_(u0: tuple, p, output_scale=1.0):
        if not isinstance(u0, tuple):
            raise ValueError("Tuple expected.")

        def vf_wrapped(*y, t):
            return vf(*y, t=t, p=p)

        tcoeffs = taylor.odejet_padded_scan(vf_wrapped, u0, num=num_derivatives)
        init = solver.initial_condition(tcoeffs, output_scale=output_scale)
        sol = ivpsolve.solve_adaptive_save_every_step(
            vf_wrapped, init,
            # small perturbation so that all save_at values are in the interior of the domain (ivpsolvers.py:136-140)
            t0=save_at[0] - small_value, t1=save_at[-1] + small_value, dt0=dt0, adaptive_solver=asolver,
            factorisation="isotropic", device=device,
        )  # fmt: skip
        dense, _ = stats.offgrid_marginals_searchsorted(ts=np.asarray(save_at, dtype=np.float64), solution=sol, solver=solver)
        return dense, {"solution": sol, "u0_solve": sol.u}

    return solve_


def asolve_scipy(method: str, vf, /, time_span, *, atol, rtol):
    """scipy reference solver (ivpsolvers.py:196-210); host-side, used for truth trajectories."""
    import scipy.integrate

    def solve_(u0: tuple, p):
        if not isinstance(u0, tuple):
            raise ValueError("Tuple expected.")

        def vf_scipy(t, y):
            return vf(y, t=t, p=p)

        (y0,) = u0
        solution = scipy.integrate.solve_ivp(vf_scipy, y0=np.asarray(y0), t_span=time_span, atol=atol, rtol=rtol, method=method)
        return solution.t, solution.y.T

    return solve_
