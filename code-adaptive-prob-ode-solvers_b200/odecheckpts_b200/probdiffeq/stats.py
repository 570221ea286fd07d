"""``probdiffeq.stats``: the two calls of src/odecheckpts/ivpsolvers.py:80-81.

The backward marginalisation over the K checkpoints already ran on the device
(csrc/pn_smooth_kernel.cuh) when the solution was computed; these functions re-expose its
result in the shapes the reference indexes.
"""

from typing import NamedTuple


class Normal(NamedTuple):
    mean: object      # [..., n, d]
    cholesky: object  # [..., n, n] lower triangular


class MarkovSeq(NamedTuple):
    init: Normal          # marginals at checkpoints 1..K-1 (init.mean[-1] = terminal)
    marginals_all: Normal  # smoothed marginals at checkpoints 0..K-1
    handle: object = None  # (descriptor, device workspace, device status) when the solve kept its conditionals
    context: object = None  # smoother solutions: what stats.offgrid_marginals_searchsorted needs to evaluate off the grid


def markov_select_terminal(posterior):
    if not isinstance(posterior, MarkovSeq):
        raise TypeError("expected the .posterior of a solve_adaptive_save_at solution")
    return posterior


def markov_marginals(markov_seq, *, reverse):
    """Marginals of the checkpoint Markov sequence at t_0 .. t_{K-2} (the terminal marginal is
    ``markov_seq.init.mean[-1]``), as ivpsolvers.py:80-86 concatenates them."""
    if not reverse:
        raise NotImplementedError("only the backward (reverse=True) factorisation is produced by the solver")
    allm = markov_seq.marginals_all
    return Normal(allm.mean[:-1], allm.cholesky[:-1])


def markov_sample(key, markov_seq, *, shape, reverse):
    """Joint samples from the checkpoint Markov sequence (experiments/5_vs_interpolation/measure.py:69-77).

    `key` is an integer seed (jax PRNG keys cannot be reproduced: Philox4x32-10 + Box-Muller on the
    device; the samples have the reference's distribution, not its bits).  Needs a solution computed
    with ``solve_adaptive_save_at(..., keep_conditionals=True)``.  Returns ``((qoi, samples), (init, None))``
    like probdiffeq: ``qoi`` [S, K-1, d] at t_0..t_{K-2} and ``init`` [S, d] at the terminal checkpoint;
    ``samples`` is None (only the quantity of interest is materialised)."""
    if not reverse:
        raise NotImplementedError("only reverse=True (backward factorisation) is produced by the solver")
    if markov_seq.handle is None:
        raise ValueError("the solution did not keep its backward conditionals: pass keep_conditionals=True to the solve")
    from .. import _cabi

    desc, workspace, status, as_numpy, batched = markov_seq.handle
    (num,) = shape if isinstance(shape, (tuple, list)) else (shape,)
    seed = int(key) if not hasattr(key, "__len__") else int(sum(int(x) << (32 * i) for i, x in enumerate(key)))
    out = _cabi.markov_sample_device(desc, workspace, status, seed, num)  # [B, S, K, d]
    if as_numpy:
        out = out.cpu().numpy()
    if not batched:
        out = out[0]
    return (out[..., :-1, :], None), (out[..., -1, :], None)


def log_marginal_likelihood(u, /, *, standard_deviation, posterior):
    """Log marginal likelihood of observations `u` [K, d] of the ODE solution at the checkpoints, observed
    with noise `standard_deviation` [K] (src/odecheckpts/train_util.py:22-24; forward value only -- the
    reference differentiates it with jax, which is out of scope here).  Ensembles: `u` may be [B, K, d]
    and `standard_deviation` [B, K]; the result is then [B].  Needs ``keep_conditionals=True``.

    Like probdiffeq's reverse Kalman-filter estimator this is the running MEAN over the K data points of
    log p(u_k | u_{k+1}, ..., u_{K-1}), i.e. the joint log density divided by K."""
    if not isinstance(posterior, MarkovSeq):
        raise TypeError("expected the .posterior of a solve_adaptive_save_at solution")
    if posterior.handle is None:
        raise ValueError("the solution did not keep its backward conditionals: pass keep_conditionals=True to the solve")
    import numpy as np

    from .. import _cabi

    desc, workspace, status, as_numpy, batched = posterior.handle
    K = desc.num_save_at
    if np.ndim(u) < 2:
        raise ValueError("u must have shape (K, d) (or (B, K, d) for an ensemble)")
    if tuple(np.shape(standard_deviation))[-1:] != (K,) or tuple(np.shape(u))[-2] != K:
        raise ValueError("u and standard_deviation need one entry per checkpoint")
    out = _cabi.log_marginal_likelihood_device(desc, workspace, status, u, standard_deviation)
    if as_numpy:
        out = out.cpu().numpy()
    return out if batched else out[0]


def offgrid_marginals_searchsorted(*, ts, solution, solver):
    """Marginals of a SMOOTHER solution at off-grid times ``ts`` (src/odecheckpts/ivpsolvers.py:117,144):
    for t in (t_i, t_{i+1}) the filter marginal at t_i is extrapolated to t, the backward conditional
    t_{i+1} -> t comes from extrapolating on to t_{i+1}, and the smoothed marginal at t_{i+1} is pulled back
    through it.  On the device this is the checkpoint machinery of the solver kernel with the accepted grid
    AND ``ts`` as checkpoints: grid points are hit exactly (one un-merged conditional per step, exactly the
    smoother's posterior), the off-grid times are interpolated inside their step, and the backward sweep
    marginalises through all of them.  Returns ``(u [len(ts), d], Normal marginals)`` like probdiffeq."""
    post = solution.posterior
    if not isinstance(post, MarkovSeq) or post.context is None:
        raise TypeError("expected a solution of solve_adaptive_save_every_step with strategy_smoother")
    import numpy as np

    from . import ivpsolve

    ctx = post.context
    if solver is not ctx["solver"]:
        raise ValueError("offgrid_marginals_searchsorted: `solver` is not the solver the solution was computed with")
    ts = np.asarray(ts, dtype=np.float64)
    grid = ctx["grid"]
    if ts.ndim != 1 or ts.min() <= grid[0] or ts.max() >= grid[-1]:
        raise ValueError("ts must lie strictly inside the solution's time interval")
    union = np.union1d(grid, ts)
    sol = ivpsolve._solve_on_checkpoints(ctx, union)
    idx = np.searchsorted(union, ts)
    marg = Normal(sol.marginals.mean[idx], sol.marginals.cholesky[idx])
    return sol.u[idx], marg
