"""GPU checks that are about trust rather than parity: the lock-free time-sliced scheduler under stress
(every launch bit-identical to the run-to-completion launch) and a calibration sanity check of the
reported standard deviations against an independent high-accuracy solution (scipy DOP853)."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cabi():
    import torch

    assert torch.cuda.is_available(), "these tests need a CUDA device"
    from odecheckpts_b200 import _cabi

    _cabi.lib()
    return _cabi


@pytest.mark.parametrize("B", [38000, 40960])
def test_time_sliced_scheduler_stress(cabi, monkeypatch, B):
    """Ensembles just above the resident lanes (37,888), small quanta (many hand-overs per member), repeated
    launches: means, standard deviations and step counts must equal the unsliced launch bit for bit."""
    import torch

    import bench

    dev = torch.device("cuda:0")
    K = 16
    save_at = torch.linspace(bench.T0, 2.0, K, dtype=torch.float64, device=dev)
    u0, par = bench.ensemble_inputs(0, B)
    u0, par = torch.as_tensor(u0, device=dev), torch.as_tensor(par, device=dev)
    desc = cabi.Desc(5, 1, 4, 2, 2, 1, 1, 1, 1e-5, 1e-5, 0.01, 0.95, 0.2, 10.0, 0.3, 0.4, B, K, 0, 1, 0, 0)

    def run():
        out = cabi.solve_device(desc, u0, par, None, save_at, None)
        torch.cuda.synchronize()
        return {k: out[k].clone() for k in ("u", "u_std", "n_accepted", "n_rejected", "status")}

    monkeypatch.setenv("PN_B200_NO_SLICE", "1")
    ref = run()
    assert int((ref["status"] != 0).sum()) == 0
    monkeypatch.delenv("PN_B200_NO_SLICE")
    for quantum in ("64", "256", None):
        if quantum is None:
            monkeypatch.delenv("PN_B200_SLICE_QUANTUM", raising=False)
        else:
            monkeypatch.setenv("PN_B200_SLICE_QUANTUM", quantum)
        for rep in range(3):
            got = run()
            for key, val in ref.items():
                assert torch.equal(got[key], val), (B, quantum, rep, key)


def test_reported_standard_deviations_are_calibrated(cabi):
    """|u - truth| / u_std = O(1) at the checkpoints: the marginal standard deviation the solver reports is a
    usable error estimate (dynamic calibration), neither wildly over- nor under-confident."""
    from scipy.integrate import solve_ivp

    a, b, c = -2.0, 1.25, -0.5  # rigid body (diffeqzoo defaults)
    save_at = np.linspace(0.0, 10.0, 21)
    y0 = np.array([1.0, 0.0, 0.9])
    truth = solve_ivp(lambda t, y: [a * y[1] * y[2], b * y[0] * y[2], c * y[0] * y[1]], (0, 10), y0, method="DOP853",
                      t_eval=save_at, rtol=1e-13, atol=1e-13).y.T  # fmt: skip
    for fact, corr, nu in (("isotropic", "ts0", 4), ("dense", "ts1", 4), ("blockdiag", "ts0", 4)):
        for tol in (1e-4, 1e-7):
            desc = cabi.Desc(1, 3, nu, 1, cabi.FACTORISATIONS[fact], cabi.CORRECTIONS[corr], 1, 1, tol, tol, 0.1,
                             0.95, 0.2, 10.0, 0.3, 0.4, 1, len(save_at), 0, 3, 0, 0)  # fmt: skip
            out = cabi.solve_host(desc, y0[None, None], np.array([[a, b, c]]), None, save_at, None)
            assert out["status"][0] == 0
            err = np.abs(out["u"][0, 1:] - truth[1:])
            ratio = err / out["u_std"][0, 1:]
            assert np.isfinite(ratio).all()
            assert 1e-3 < np.median(ratio) < 10.0, (fact, tol, np.median(ratio))
            assert ratio.max() < 100.0, (fact, tol, ratio.max())
            assert err.max() < 300 * tol


@pytest.mark.parametrize("problem,d,nu,q,P,params,u0,fact,corr,t1", [
    ("van_der_pol", 1, 4, 2, 1, (5.0,), np.array([[2.0], [0.0]]), "dense", "ts1", 3.0),
    ("logistic", 1, 3, 1, 2, (1.0, 1.0), np.array([[0.1]]), "isotropic", "ts0", 2.5),
    ("rigid_body", 3, 4, 1, 3, (-2.0, 1.25, -0.5), np.array([[1.0, 0.0, 0.9]]), "isotropic", "ts0", 10.0),
])  # fmt: skip
def test_solver_mle_matches_the_oracle_and_rescales_the_uncalibrated_posterior(cabi, problem, d, nu, q, P, params, u0, fact, corr, t1):
    """ivpsolvers.solver_mle (no call site in the reference, restated from probdiffeq): the solve runs
    with the initial output scale, the running quasi-MLE is reported per checkpoint and its final value scales the
    posterior standard deviations.  GPU vs oracle bit for bit; means equal the uncalibrated solver's."""
    from oracle import pn_oracle

    K, tol, B = 9, 1e-5, 3
    save_at = np.linspace(0.0, t1, K)
    rng = np.random.default_rng(1)
    u0b = u0[None] * (1.0 + 0.01 * rng.standard_normal((B, 1, 1)))
    par = np.tile(params, (B, 1))
    mk = lambda calib: cabi.Desc(cabi.PROBLEM_IDS[problem], d, nu, q, cabi.FACTORISATIONS[fact], cabi.CORRECTIONS[corr], 1,  # noqa: E731
                                 cabi.CALIBRATIONS[calib], tol, tol, 0.01, 0.95, 0.2, 10.0, 0.3, 0.4, B, K, 0, P, 0, 0)
    mle = cabi.solve_host(mk("mle"), u0b, par, None, save_at, None, full=True)
    none = cabi.solve_host(mk("none"), u0b, par, None, save_at, None, full=True)
    assert (mle["status"] == 0).all()
    np.testing.assert_array_equal(mle["u"], none["u"])
    np.testing.assert_array_equal(mle["n_accepted"], none["n_accepted"])
    scale = mle["output_scale"][:, -1]
    assert (scale > 0).all() and np.isfinite(scale).all()
    np.testing.assert_array_equal(mle["u_std"], scale[:, None, None] * none["u_std"])
    cfg = pn_oracle.make_config(problem, d, nu, q, factorisation=fact, correction=corr, calibration="mle", atol=tol, rtol=tol,
                                dt0=0.01, num_params=P)  # fmt: skip
    for b in range(B):
        ora = pn_oracle.solve_save_at_lml(cfg, u0b[b], par[b], save_at, np.zeros((K, d)), np.ones(K))
        np.testing.assert_array_equal(mle["u"][b], ora["u"])
        np.testing.assert_array_equal(mle["u_std"][b], ora["u_std"])
        np.testing.assert_array_equal(mle["output_scale"][b, 1:], ora["output_scale"][1:, 0])
        np.testing.assert_array_equal(mle["marg_chol"][b].reshape(K, -1), ora["marg_chol"].reshape(K, -1))
