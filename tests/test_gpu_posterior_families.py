"""Posterior sampling and the log marginal likelihood for the kernel families beyond thread-per-IVP /
lane-per-dimension: the CTA-per-IVP isotropic kernel (Brusselator posteriors, experiments/4_brusselator) and
the three dense families with d > 1.  Sampling parity is distributional (stats.markov_sample,
experiments/5_vs_interpolation/measure.py:69-77: jax.random bits cannot be reproduced); the likelihood
(src/odecheckpts/train_util.py:22-24) is compared with the oracle bit for bit."""

import numpy as np
import pytest

import problems_util as pu

pytestmark = pytest.mark.gpu


def _solve(problem, d, nu, q, fact, corr, u0, params, save_at, tol, B=1):
    import torch

    from odecheckpts_b200 import _cabi

    K, P = len(save_at), len(params)
    desc = _cabi.Desc(_cabi.PROBLEM_IDS[problem], d, nu, q, _cabi.FACTORISATIONS[fact], _cabi.CORRECTIONS[corr], 1, 1,
                      tol, tol, 0.01, 0.95, 0.2, 10.0, 0.3, 0.4, B, K, 0, P, 0, 0)  # fmt: skip
    dev = torch.device("cuda:0")
    T = lambda x: torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64, device=dev)  # noqa: E731
    u0_b = np.tile(np.asarray(u0, dtype=float)[None], (B, 1, 1))
    par = T(np.tile(np.asarray(params, dtype=float), (B, 1))) if P else None
    out = _cabi.solve_device(desc, T(u0_b), par, None, T(save_at), None)
    assert int((out["status"] != 0).sum()) == 0
    return desc, out


SAMPLE_CASES = [
    # problem, d, nu, q, factorisation, correction, u0, params, save_at, tol  -> kernel family
    ("brusselator", 64, 4, 1, "isotropic", "ts0", pu.brusselator_u0(32), (0.02,), np.linspace(0, 1.0, 6), 1e-5),  # CTA per IVP
    ("rigid_body", 3, 4, 1, "dense", "ts1", pu.rigid_body_u0(), pu.RIGID_BODY_PARAMS, np.linspace(0, 5.0, 6), 1e-4),  # register columns
    ("brusselator", 8, 4, 1, "dense", "ts1", pu.brusselator_u0(4), (0.02,), np.linspace(0, 1.0, 5), 1e-4),  # D = 40: shared memory
    ("brusselator", 16, 4, 1, "dense", "ts1", pu.brusselator_u0(8), (0.02,), np.linspace(0, 0.5, 4), 1e-4),  # D = 80: dense CTA (transposed slots)
]


@pytest.mark.parametrize("case", SAMPLE_CASES, ids=["wide", "dense_rows", "dense_smem", "dense_cta"])
def test_posterior_samples_match_the_smoothed_marginals(case):
    from odecheckpts_b200 import _cabi

    problem, d, nu, q, fact, corr, u0, params, save_at, tol = case
    desc, out = _solve(problem, d, nu, q, fact, corr, u0, params, save_at, tol, B=2)
    S = 6000
    smp = _cabi.markov_sample_device(desc, out["_workspace"], out["status"], 7, S).cpu().numpy()  # [B, S, K, d]
    assert smp.shape == (2, S, len(save_at), d) and np.isfinite(smp).all()
    u, u_std = out["u"].cpu().numpy(), out["u_std"].cpu().numpy()
    for b in range(2):
        mean, std = smp[b].mean(axis=0), smp[b].std(axis=0)
        se = u_std[b] / np.sqrt(S)
        assert np.all(np.abs(mean - u[b]) <= 5 * se + 1e-12), np.abs(mean - u[b]).max()
        big = u_std[b] > 1e-10
        np.testing.assert_allclose(std[big], u_std[b][big], rtol=0.07)
    # one trajectory per draw: neighbouring checkpoints are correlated
    k = len(save_at) // 2
    c = np.corrcoef(smp[0, :, k, 0], smp[0, :, k + 1, 0])[0, 1]
    assert abs(c) > 0.1
    # the two members are identical IVPs with different random streams
    assert not np.array_equal(smp[0], smp[1])
    a16 = _cabi.markov_sample_device(desc, out["_workspace"], out["status"], 9, 16).cpu().numpy()
    b16 = _cabi.markov_sample_device(desc, out["_workspace"], out["status"], 9, 16).cpu().numpy()
    np.testing.assert_array_equal(a16, b16)  # same key, same draws


def test_wide_family_log_marginal_likelihood_bitwise_vs_oracle(oracle):
    from odecheckpts_b200 import _cabi

    N = 32
    d, K = 2 * N, 5
    save_at = np.linspace(0.0, 1.0, K)
    u0 = pu.brusselator_u0(N)
    desc, out = _solve("brusselator", d, 4, 1, "isotropic", "ts0", u0, (0.02,), save_at, 1e-5, B=2)
    rng = np.random.default_rng(0)
    std = 0.05 + 0.1 * rng.random((2, K))
    data = out["u"].cpu().numpy() + 0.1 * rng.standard_normal((2, K, d))
    lml = _cabi.log_marginal_likelihood_device(desc, out["_workspace"], out["status"], data, std).cpu().numpy()
    cfg = oracle.make_config("brusselator", d, 4, 1, atol=1e-5, rtol=1e-5, dt0=0.01, num_params=1, reduction_group=128)
    for b in range(2):
        ora = oracle.solve_save_at_lml(cfg, u0, [0.02], save_at, data[b], std[b])
        assert ora["status"] == 0
        np.testing.assert_array_equal(out["u"][b].cpu().numpy(), ora["u"])
        assert lml[b] == ora["lml"], (b, lml[b], ora["lml"])
    assert lml[0] != lml[1]


def test_brusselator_textbook_smoother_on_the_cta_per_ivp_kernel(oracle):
    """experiments/4_brusselator/run.py:51-117 at N = 32 (d = 64, the CTA-per-IVP isotropic kernel): the baseline
    step count (solve_adaptive_terminal_values), the textbook smoother (strategy_smoother +
    solve_adaptive_save_every_step: one backward conditional per accepted step) and the checkpoint solver
    (strategy_fixedpoint + solve_adaptive_save_at) -- same posterior by two routes, the paper's comparison."""
    from odecheckpts_b200 import ivps
    from odecheckpts_b200.probdiffeq import impl, ivpsolve, taylor
    from odecheckpts_b200.probdiffeq import ivpsolvers as pdi

    N, tol, t1 = 32, 1e-6, 2.0
    vf, u0, (t0, _), params = ivps.brusselator(N=N)
    impl.impl.select("isotropic", ode_shape=(2 * N,))
    ctrl = ivpsolve.control_proportional_integral()
    ibm, ts0 = pdi.prior_ibm(num_derivatives=4), pdi.correction_ts0(ode_order=1)
    f = lambda *y, t=None: vf(*y, t=t, p=params)  # noqa: E731
    tcoeffs = taylor.odejet_unroll(lambda *y: vf(*y, t=t0, p=params), u0, num=4)

    def adaptive(strategy):
        solver = pdi.solver_dynamic(strategy)
        return solver.initial_condition(tcoeffs, 1.0), ivpsolve.adaptive(solver, atol=tol, rtol=tol, control=ctrl)

    init, fixedpoint = adaptive(pdi.strategy_fixedpoint(ibm, ts0))
    base = ivpsolve.solve_adaptive_terminal_values(f, init, t0=t0, t1=t1, dt0=0.01, adaptive_solver=fixedpoint)
    n_steps = int(base.num_steps)
    assert n_steps > 100
    init_s, smoother = adaptive(pdi.strategy_smoother(ibm, ts0))
    text = ivpsolve.solve_adaptive_save_every_step(f, init_s, t0=t0, t1=t1, dt0=0.01, adaptive_solver=smoother)
    grid = np.asarray(text.t)
    assert len(grid) == n_steps + 1 and grid[0] == t0 and grid[-1] == t1 and text.u.shape == (n_steps + 1, 2 * N)
    # the checkpoint solver on a coarse grid evaluates the same posterior with O(K) memory
    save_at = np.linspace(t0, t1, 9)
    ckpt = ivpsolve.solve_adaptive_save_at(f, init, save_at=save_at, dt0=0.01, adaptive_solver=fixedpoint)
    assert int(ckpt.num_steps[-1]) == n_steps
    union = np.union1d(grid, save_at)
    both = ivpsolve._solve_on_checkpoints(text.posterior.context, union)
    np.testing.assert_allclose(both.u[np.searchsorted(union, save_at)], ckpt.u, rtol=1e-9, atol=1e-12)
    np.testing.assert_array_equal(both.u[np.searchsorted(union, grid)][1:-1].shape, text.u[1:-1].shape)
    # bit-identical to the oracle run with the accepted grid as checkpoints
    cfg = oracle.make_config("brusselator", 2 * N, 4, 1, atol=tol, rtol=tol, dt0=0.01, num_params=1, reduction_group=128)
    ora = oracle.solve_save_at(cfg, np.asarray(u0[0])[None], list(params), grid)
    assert ora["status"] == 0 and ora["n_accepted"][-1] == n_steps
    np.testing.assert_array_equal(np.asarray(text.u), ora["u"])
    np.testing.assert_array_equal(np.asarray(text.u_std), ora["u_std"])
