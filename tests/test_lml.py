"""stats.log_marginal_likelihood at the checkpoints (src/odecheckpts/train_util.py:22-24).

probdiffeq's sources are absent and the reference commits no likelihood values, so the oracle's
restatement is pinned against a brute-force computation instead: the joint Gaussian of the 0-th
derivative at all K checkpoints, built from the backward conditionals with numpy.  The GPU path is
then compared with the oracle bit for bit."""

import numpy as np
import pytest

import problems_util as pu

CASES = [
    # problem, d, nu, q, P, params, u0, t1, factorisation, correction, reduction_group
    ("logistic", 1, 3, 1, 2, (1.0, 1.0), np.array([[0.1]]), 2.5, "isotropic", "ts0", 0),
    ("van_der_pol", 1, 4, 2, 1, (10.0,), np.array([[2.0], [0.0]]), 3.0, "dense", "ts1", 0),
    ("rigid_body", 3, 4, 1, 3, pu.RIGID_BODY_PARAMS, np.array([[1.0, 0.0, 0.9]]), 5.0, "isotropic", "ts0", 0),
    ("rigid_body", 3, 2, 1, 3, pu.RIGID_BODY_PARAMS, np.array([[1.0, 0.0, 0.9]]), 5.0, "blockdiag", "ts0", 4),
    ("pleiades", 14, 3, 2, 0, (), pu.pleiades_u0(), 1.0, "isotropic", "ts0", 16),
    ("pleiades", 14, 3, 2, 0, (), pu.pleiades_u0(), 1.0, "blockdiag", "ts0", 16),
]
IDS = [f"{c[0]}-{c[8]}" for c in CASES]


def _brute_force_joint_logpdf(res, data, std, F, N, Ct):
    """log N(data; H mu, H Sigma H^T + diag(std^2)) of the checkpoint Markov chain
    x_{K-1} ~ N(m, L L^T), x_{k-1} = G_k x_k + g_k + Lam_k xi, per factor set and mean column."""
    K = len(data)
    C_ = Ct // F
    total = 0.0
    for f in range(F):
        G, Lam = res["cond_G"][:, f], res["cond_Lam"][:, f]
        maps = {K - 1: {K - 1: res["marg_chol"][K - 1, f]}}  # state k = mu_k + sum_j maps[k][j] eps_j
        for k in range(K - 1, 0, -1):
            maps[k - 1] = {j: G[k] @ M for j, M in maps[k].items()}
            maps[k - 1][k - 1] = Lam[k]
        T = np.zeros((K, K * N))
        for k in range(K):
            for j, M in maps[k].items():
                T[k, j * N:(j + 1) * N] = M[0]
        S = T @ T.T + np.diag(np.asarray(std) ** 2)
        _, logdet = np.linalg.slogdet(S)
        for c in range(f * C_, f * C_ + C_):
            mu = np.zeros((K, N))
            mu[K - 1] = res["marg_mean"][K - 1][:, c]
            for k in range(K - 1, 0, -1):
                mu[k - 1] = G[k] @ mu[k] + res["cond_g"][k][:, c]
            r = data[:, c] - mu[:, 0]
            total += -0.5 * r @ np.linalg.solve(S, r) - 0.5 * logdet - 0.5 * K * np.log(2 * np.pi)
    return total


def _setup(oracle, case, K=6, seed=0):
    prob, d, nu, q, P, params, u0, t1, fact, corr, group = case
    save_at = np.linspace(0, t1, K)
    cfg = oracle.make_config(prob, d, nu, q, factorisation=fact, correction=corr, atol=1e-4, rtol=1e-4, dt0=0.1,
                             num_params=P, reduction_group=group)  # fmt: skip
    plain = oracle.solve_save_at(cfg, u0, params, save_at, full=True)
    rng = np.random.default_rng(seed)
    std = 0.05 + 0.1 * rng.random(K)
    data = plain["u"] + 0.1 * rng.standard_normal((K, d))
    return cfg, save_at, plain, data, std


@pytest.mark.parametrize("case", CASES[:4], ids=IDS[:4])
def test_oracle_lml_equals_the_joint_gaussian_log_density(oracle, case):
    cfg, save_at, plain, data, std = _setup(oracle, case)
    res = oracle.solve_save_at_lml(cfg, case[6], case[5], save_at, data, std)
    assert res["status"] == 0
    np.testing.assert_array_equal(res["u"], plain["u"])  # the likelihood sweep does not disturb the solve
    np.testing.assert_array_equal(res["marg_chol"], plain["marg_chol"])
    F, N, Ct = oracle._engine_dims(cfg)
    joint = _brute_force_joint_logpdf(res, data, std, F, N, Ct)
    K = len(save_at)
    np.testing.assert_allclose(res["lml"] * K, joint, rtol=1e-12)  # running mean over K data points
    # more noise than signal mismatch -> flatter likelihood; data far away -> much smaller likelihood
    far = oracle.solve_save_at_lml(cfg, case[6], case[5], save_at, data + 5.0, std)
    assert far["lml"] < res["lml"] - 100.0


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_gpu_lml_bitwise_vs_oracle(oracle, case):
    import torch

    from odecheckpts_b200 import _cabi

    prob, d, nu, q, P, params, u0, t1, fact, corr, group = case
    cfg, save_at, plain, data, std = _setup(oracle, case)
    K, B = len(save_at), 3
    rng = np.random.default_rng(3)
    u0_b = u0[None] * (1.0 + 1e-3 * rng.standard_normal((B,) + u0.shape))
    u0_b[0] = u0
    data_b = data[None] + 0.05 * rng.standard_normal((B, K, d))
    data_b[0] = data
    std_b = np.tile(std, (B, 1)) * (1.0 + 0.2 * rng.random((B, K)))
    std_b[0] = std
    desc = _cabi.Desc(_cabi.PROBLEM_IDS[prob], d, nu, q, _cabi.FACTORISATIONS[fact], _cabi.CORRECTIONS[corr], 1, 1,
                      1e-4, 1e-4, 0.1, 0.95, 0.2, 10.0, 0.3, 0.4, B, K, 0, P, 0, 0)  # fmt: skip
    dev = torch.device("cuda:0")
    T = lambda x: torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64, device=dev)  # noqa: E731
    par = T(np.tile(np.asarray(params, dtype=float), (B, 1))) if P else None
    out = _cabi.solve_device(desc, T(u0_b), par, None, T(save_at), None)
    lml = _cabi.log_marginal_likelihood_device(desc, out["_workspace"], out["status"], data_b, std_b).cpu().numpy()
    for b in range(B):
        ora = oracle.solve_save_at_lml(cfg, u0_b[b], params, save_at, data_b[b], std_b[b])
        assert ora["status"] == 0
        np.testing.assert_array_equal(out["u"][b].cpu().numpy(), ora["u"])
        assert lml[b] == ora["lml"], (b, lml[b], ora["lml"])


@pytest.mark.gpu
def test_learn_ode_style_likelihood_through_the_builder_api(oracle):
    # experiments/old/6_learn_ode/learn.py:83-114: dense EKF1 + fixed-point, uncalibrated solver,
    # likelihood of noisy observations of the solution as a function of the initial value
    from odecheckpts_b200 import ivps
    from odecheckpts_b200.probdiffeq import impl, ivpsolve, ivpsolvers, stats, taylor

    vf, (y0, dy0), _ = ivps.van_der_pol(mu=10.0)
    impl.impl.select("dense", ode_shape=(1,))
    ibm = ivpsolvers.prior_ibm(num_derivatives=4)
    ts1 = ivpsolvers.correction_ts1(ode_order=2)
    solver = ivpsolvers.solver(ivpsolvers.strategy_fixedpoint(ibm, ts1))
    ctrl = ivpsolve.control_proportional_integral()
    asolver = ivpsolve.adaptive(solver, atol=1e-4, rtol=1e-4, control=ctrl)
    save_at = np.linspace(0.0, 3.0, 12)

    def solve(init):
        tcoeffs = taylor.odejet_padded_scan(lambda *y: vf(*y, t=0.0), init, num=3)
        ic = solver.initial_condition(tcoeffs, np.ones(()))
        return ivpsolve.solve_adaptive_save_at(vf, ic, save_at=save_at, dt0=0.1, adaptive_solver=asolver, keep_conditionals=True)

    truth = solve((y0, dy0))
    std = np.full(len(save_at), 0.05)
    data = truth.u + std[:, None] * np.random.default_rng(0).standard_normal(truth.u.shape)
    lml_true = stats.log_marginal_likelihood(data, standard_deviation=std, posterior=truth.posterior)
    lml_off = stats.log_marginal_likelihood(data, standard_deviation=std, posterior=solve((y0 + 0.3, dy0)).posterior)
    assert np.isfinite(lml_true) and lml_off < lml_true - 1.0
    # against the oracle (uncalibrated solver = calibration "none")
    cfg = oracle.make_config("van_der_pol", 1, 4, 2, factorisation="dense", correction="ts1", calibration="none",
                             atol=1e-4, rtol=1e-4, dt0=0.1, num_params=1)  # fmt: skip
    ora = oracle.solve_save_at_lml(cfg, np.array([y0, dy0]).reshape(2, 1), (10.0,), save_at, data, std)
    assert lml_true == ora["lml"]
    with pytest.raises(ValueError):
        plain = ivpsolve.solve_adaptive_save_at(vf, solver.initial_condition(
            taylor.odejet_padded_scan(lambda *y: vf(*y, t=0.0), (y0, dy0), num=3), np.ones(())),
            save_at=save_at, dt0=0.1, adaptive_solver=asolver)  # fmt: skip
        stats.log_marginal_likelihood(data, standard_deviation=std, posterior=plain.posterior)


def test_oracle_output_scale_per_checkpoint(oracle):
    # solution.output_scale: entry 0 is the initial scale; the uncalibrated solver carries it unchanged,
    # the dynamic solver carries the calibrated scale of the step that reached / crossed each checkpoint
    save_at = np.linspace(0.0, 3.0, 8)
    u0, params = np.array([[2.0], [0.0]]), (10.0,)
    for calib in ("none", "dynamic"):
        cfg = oracle.make_config("van_der_pol", 1, 4, 2, factorisation="dense", correction="ts1", calibration=calib,
                                 atol=1e-5, rtol=1e-5, dt0=0.1, num_params=1)  # fmt: skip
        res = oracle.solve_save_at_lml(cfg, u0, params, save_at, np.zeros((8, 1)), np.ones(8), output_scale0=2.5)
        scale = res["output_scale"][:, 0]
        assert scale[0] == 2.5
        if calib == "none":
            assert (scale == 2.5).all()
        else:
            assert (scale[1:] > 0).all() and len(np.unique(scale)) == 8
    # blockdiag: one scale per dimension
    cfg = oracle.make_config("rigid_body", 3, 3, 1, factorisation="blockdiag", atol=1e-6, rtol=1e-6, dt0=0.1, num_params=3)
    res = oracle.solve_save_at_lml(cfg, np.array([[1.0, 0.0, 0.9]]), pu.RIGID_BODY_PARAMS, save_at, np.zeros((8, 3)), np.ones(8))
    assert res["output_scale"].shape == (8, 3) and (np.abs(np.diff(res["output_scale"][1:], axis=1)) > 0).any()
