"""CPU: the panel-blocked dense engine of the oracle (the operation order of the CTA-per-IVP tensor-core
kernel, oracle/pn_blocked.c) against the unblocked column-by-column engine that is pinned to the
reference's goldens.  Both are Householder QRs of the same stacked matrices, so they agree to rounding."""

import numpy as np
import pytest


def _u0(N):
    return np.concatenate([np.sin(2 * np.pi * np.linspace(0, 1, N)) + 1, 3 * np.ones(N)])[None, :]


@pytest.mark.parametrize("N,corr,strat,nb", [(4, "ts1", "fixedpoint", 16), (4, "ts0", "filter", 8), (8, "ts1", "fixedpoint", 16),
                                             (3, "ts1", "fixedpoint", 16)])  # fmt: skip
def test_blocked_engine_matches_unblocked_single_steps(oracle, N, corr, strat, nb):
    d, nu = 2 * N, 4
    D = (nu + 1) * d
    kw = dict(factorisation="dense", correction=corr, strategy=strat, atol=1e-6, rtol=1e-6, dt0=0.01, num_params=1)
    rng = np.random.default_rng(N)
    mean = oracle.taylor_init("brusselator", _u0(N), nu, [0.02]).reshape(D)
    A = rng.standard_normal((D, D)) * 1e-3
    chol = np.linalg.cholesky(A @ A.T + 1e-8 * np.eye(D))
    G = np.eye(D) + 0.01 * rng.standard_normal((D, D))
    g = 0.01 * rng.standard_normal(D)
    B = rng.standard_normal((D, D)) * 1e-3
    Lam = np.linalg.cholesky(B @ B.T + 1e-8 * np.eye(D))
    a = oracle.attempt_step(oracle.make_config("brusselator", d, nu, 1, **kw), [0.02], 0.0, 0.02, 1.0, 1.0, mean, chol, (G, g, Lam))
    b = oracle.attempt_step(oracle.make_config("brusselator", d, nu, 1, dense_block=nb, **kw), [0.02], 0.0, 0.02, 1.0, 1.0, mean, chol, (G, g, Lam))
    for key in ("error_norm", "dt_proposed", "sigma"):
        assert abs(a[key] - b[key]) <= 1e-10 * abs(a[key])
    np.testing.assert_allclose(b["mean"], a["mean"], rtol=1e-10, atol=1e-13)
    cov = lambda L: L[0] @ L[0].T  # noqa: E731
    np.testing.assert_allclose(cov(b["chol"]), cov(a["chol"]), rtol=1e-8, atol=1e-16)
    if strat == "fixedpoint":
        np.testing.assert_allclose(b["G"], a["G"], rtol=1e-8, atol=1e-11)
        np.testing.assert_allclose(b["g"], a["g"], rtol=1e-8, atol=1e-11)
        np.testing.assert_allclose(cov(b["Lam"]), cov(a["Lam"]), rtol=1e-7, atol=1e-16)


@pytest.mark.parametrize("N,nb", [(4, 16), (8, 16), (8, 32)])
def test_blocked_engine_full_solve_matches_unblocked(oracle, N, nb):
    """Adaptive checkpoint solve, dense EKF1 + fixed-point smoother: same accepted / rejected counts and
    smoothed means / standard deviations within the north-star tolerance."""
    d, nu, K = 2 * N, 4, 6
    kw = dict(factorisation="dense", correction="ts1", strategy="fixedpoint", atol=1e-5, rtol=1e-5, dt0=0.01, num_params=1)
    save_at = np.linspace(0.0, 1.0, K)
    a = oracle.solve_save_at(oracle.make_config("brusselator", d, nu, 1, **kw), _u0(N), [0.02], save_at)
    b = oracle.solve_save_at(oracle.make_config("brusselator", d, nu, 1, dense_block=nb, reduction_group=256, **kw), _u0(N), [0.02], save_at)
    assert a["status"] == 0 and b["status"] == 0
    np.testing.assert_array_equal(a["n_accepted"], b["n_accepted"])
    assert a["n_rejected"] == b["n_rejected"]
    np.testing.assert_allclose(b["u"], a["u"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(b["u_std"], a["u_std"], rtol=1e-6, atol=1e-14)
