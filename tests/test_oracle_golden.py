"""Pin the CPU oracle against the reference's committed golden artefacts (CPU only).

The reference asserts none of these itself; they are the
`.npy` outputs of its experiment scripts, extracted by tests/golden/make_golden.py.
"""

import numpy as np
import pytest
import scipy.integrate

import problems_util as pu


def _truth(oracle, problem, y0, params, ts, q=1):
    def f(t, y):
        if q == 1:
            return oracle.vf(problem, y.reshape(1, -1), params)
        return np.concatenate([y[len(y) // 2 :], oracle.vf(problem, y.reshape(2, -1), params)])

    s = scipy.integrate.solve_ivp(f, (ts[0], ts[-1]), y0, t_eval=ts, method="DOP853", atol=1e-13, rtol=1e-13)
    return s.y.T


# experiments/4_brusselator/run.py:51-61,119-138: isotropic EKF0 nu=4 tol=1e-8 dynamic fixed-point
@pytest.mark.parametrize("N,exact", [(4, True), (16, True), (8, True), (32, False)])
def test_brusselator_step_counts_and_checkpoint_means(oracle, goldens, N, exact):
    cfg = oracle.make_config("brusselator", 2 * N, 4, 1, atol=1e-8, rtol=1e-8, dt0=0.01, num_params=1)
    out = oracle.solve_save_at(cfg, pu.brusselator_u0(N), [1.0 / 50.0], np.linspace(0.0, 10.0, 200))
    assert out["status"] == 0
    idx = list(goldens["brusselator_N"]).index(N)
    want = int(goldens["brusselator_num_steps_checkpoint"][idx])
    assert want == int(goldens["brusselator_num_steps_terminal"][idx])  # no step clipping
    got = int(out["n_accepted"][-1])
    if exact:
        assert got == want
    else:
        assert abs(got - want) <= 0.02 * want
    ys = goldens[f"brusselator_ys_N{N}"]
    # N = 8: same step sequence, means within 2e-9 (the reference's LAPACK QR rounds differently)
    np.testing.assert_allclose(out["u"], ys, rtol=0, atol=(5e-9 if N == 8 else 1e-9) if exact else 5e-8)


# experiments/5_vs_interpolation/measure.py:44-68,191-192: iso EKF0 o2 nu=4 UNCALIBRATED
def test_three_body_step_counts(oracle, goldens):
    for tol, want in zip(goldens["threebody_tols"], goldens["threebody_num_steps"]):
        cfg = oracle.make_config(
            "three_body", 2, 4, 2, calibration="none", atol=tol, rtol=tol, dt0=0.01, num_params=1
        )
        out = oracle.solve_save_at(cfg, pu.three_body_u0(), [pu.THREE_BODY_MU], np.linspace(0, pu.THREE_BODY_T, 50))
        assert out["status"] == 0
        assert int(out["n_accepted"][-1]) == int(want)


# experiments/2_workprec_simple/run_simple.py:58-80,200: grid length of the smoother variant
@pytest.mark.parametrize("nu,key", [(2, "rigid_interp_nu2"), (4, "rigid_interp_nu4")])
def test_rigid_body_grid_lengths(oracle, goldens, nu, key):
    tols = goldens[key + "_list_of_args"]
    want = goldens[key + "_length_of_longest_vector"]
    got = []
    for tol in tols:
        t = tol * 100  # run_simple.py:60-64
        cfg = oracle.make_config("rigid_body", 3, nu, 1, strategy="filter", atol=1e-3 * t, rtol=t, dt0=50.0, num_params=3)
        out = oracle.solve_save_every_step(cfg, pu.rigid_body_u0(), pu.RIGID_BODY_PARAMS, -1e-6, 50.0 + 1e-6)
        got.append(len(out["t"]))
    got = np.asarray(got, dtype=float)
    # dt0 = 50 makes the first steps chaotic at the ulp level: within 4 %
    np.testing.assert_allclose(got, want, rtol=0.04)
    if nu == 2:
        assert got[-1] == want[-1] == 4158


# experiments/2_workprec_simple/run_simple.py:38-56,169-178: RMSE of the smoothed checkpoint means
@pytest.mark.parametrize("nu,key", [(2, "rigid_loop_nu2"), (4, "rigid_loop_nu4")])
def test_rigid_body_checkpoint_rmse(oracle, goldens, nu, key):
    xs = goldens["rigid_checkpoints"]
    ref = _truth(oracle, "rigid_body", pu.rigid_body_u0()[0], pu.RIGID_BODY_PARAMS, xs)
    for tol, want in zip(goldens[key + "_list_of_args"], goldens[key + "_precision"]):
        t = tol * 100
        cfg = oracle.make_config("rigid_body", 3, nu, 1, atol=1e-3 * t, rtol=t, dt0=50.0, num_params=3)
        out = oracle.solve_save_at(cfg, pu.rigid_body_u0(), pu.RIGID_BODY_PARAMS, xs)
        rmse = np.linalg.norm(out["u"] - ref) / np.sqrt(ref.size)
        # loose tolerances + dt0=50 are chaotic (first step rejected many times); tight ones are not
        rel = 0.15 if tol > 2e-8 else 5e-3
        assert abs(rmse / want - 1.0) < rel, (nu, tol, rmse, want)


# experiments/3_workprec_harder/run_harder.py:42-60: Pleiades, isotropic EKF0, ode_order=2, 50 checkpoints
@pytest.mark.parametrize("nu,key,ntol", [(3, "pleiades_nu3", 5), (5, "pleiades_nu5", 5), (8, "pleiades_nu8", 3)])
def test_pleiades_checkpoint_rmse(oracle, goldens, nu, key, ntol):
    xs = goldens["pleiades_checkpoints"]
    y0 = pu.pleiades_u0()
    ref = _truth(oracle, "pleiades", y0.ravel(), [], xs, q=2)[:, :14]
    for tol, want in list(zip(goldens[key + "_list_of_args"], goldens[key + "_precision"]))[:ntol]:
        t = tol * 10  # run_harder.py:45-47
        cfg = oracle.make_config("pleiades", 14, nu, 2, atol=1e-3 * t, rtol=t, dt0=0.1)
        out = oracle.solve_save_at(cfg, y0, [], xs)
        assert out["status"] == 0
        rmse = np.linalg.norm(out["u"] - ref) / np.sqrt(ref.size)
        assert abs(rmse / want - 1.0) < 1e-3, (nu, tol, rmse, want)


# experiments/1_van_der_pol/vdp.py:61-80: dense EKF1 o2 nu=4 filter dynamic tol=1e-3 dt0=0.01
def _vdp_cfg(oracle, strategy="filter", tol=1e-3):
    return oracle.make_config(
        "van_der_pol", 1, 4, 2, factorisation="dense", correction="ts1", strategy=strategy,
        atol=tol, rtol=tol, dt0=0.01, num_params=1,
    )  # fmt: skip


def test_vdp_adaptive_grid_prefix_and_count(oracle, goldens):
    grid, sol = goldens["vdp_grid"], goldens["vdp_solution"][:, 0]
    out = oracle.solve_save_every_step(_vdp_cfg(oracle), pu.van_der_pol_u0(), [1e3], 0.0, 6.3)
    t, u = out["t"], out["u"][:, 0]
    # ulp-level agreement until the stiff transient amplifies rounding differences
    np.testing.assert_allclose(t[:26], grid[:26], rtol=0, atol=2e-14)
    np.testing.assert_allclose(u[:26], sol[:26], rtol=0, atol=2e-14)
    assert t[-1] == 6.3 and grid[-1] == 6.3
    # chaotic afterwards: the reference's own count is only reproducible within ~1 %
    assert abs((len(t) - 1) - (len(grid) - 1)) <= 0.01 * (len(grid) - 1)


def test_vdp_teacher_forced_replay(oracle, goldens):
    # vdp.py:88-91 replays the adaptive grid with solve_fixed_grid
    grid, sol = goldens["vdp_grid"], goldens["vdp_solution"][:, 0]
    rep = oracle.solve_fixed_grid(_vdp_cfg(oracle), pu.van_der_pol_u0(), [1e3], grid)
    en = rep["error_norms"][1:-1]  # the last golden step is the interpolated end point
    assert (en <= 1.0).mean() > 0.999
    assert np.sort(en)[-2] < 1.0
    rel = np.abs(rep["u"][:, 0] - sol) / np.abs(sol)
    assert np.median(rel) < 1e-9
    assert (rel < 1e-6).mean() > 0.85
    assert np.abs(rep["u"][:, 0] - sol).max() < 1e-3  # worst point sits inside the relaxation jump


def test_taylor_coefficients_vdp(oracle):
    tc = oracle.taylor_init("van_der_pol", pu.van_der_pol_u0(), 4, [1e3])
    np.testing.assert_allclose(tc[:, 0], [2.0, 0.0, -2e3, 6e6, -1.7998e10], rtol=1e-15)


def test_smoothed_initial_value_reproduces_u0(oracle):
    # App. A.6: marginalising back to t0 reproduces u0
    cfg = oracle.make_config("brusselator", 8, 4, 1, atol=1e-8, rtol=1e-8, dt0=0.01, num_params=1)
    out = oracle.solve_save_at(cfg, pu.brusselator_u0(4), [0.02], np.linspace(0.0, 10.0, 200))
    np.testing.assert_allclose(out["u"][0], pu.brusselator_u0(4)[0], rtol=0, atol=5e-15)


def test_solver_mle_is_the_uncalibrated_solve_with_rescaled_covariances(oracle):
    """solver_mle restated (no call site in the reference): same means and steps as the uncalibrated
    solver; standard deviations multiplied by the final running quasi-MLE; the MLE equals the RMS of the per-step
    whitened residuals, which for dynamic calibration would be the per-step scales themselves."""
    import numpy as np

    save_at = np.linspace(0.0, 2.5, 6)
    kw = dict(atol=1e-5, rtol=1e-5, dt0=0.1, num_params=2)
    a = oracle.solve_save_at_lml(oracle.make_config("logistic", 1, 3, 1, calibration="mle", **kw), [[0.1]], [1.0, 1.0], save_at,
                                 np.zeros((6, 1)), np.ones(6))  # fmt: skip
    b = oracle.solve_save_at_lml(oracle.make_config("logistic", 1, 3, 1, calibration="none", **kw), [[0.1]], [1.0, 1.0], save_at,
                                 np.zeros((6, 1)), np.ones(6))  # fmt: skip
    np.testing.assert_array_equal(a["u"], b["u"])
    s = a["output_scale"][-1, 0]
    assert 0 < s < 10
    np.testing.assert_allclose(a["u_std"], s * b["u_std"], rtol=1e-15)
    # the running value moves from checkpoint to checkpoint and ends at the final scale
    assert len(set(np.round(a["output_scale"][1:, 0], 12))) > 1
