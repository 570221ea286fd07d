"""GPU parity tests of the CTA-per-IVP dense kernel (large D = (nu+1) d, blocked Householder QR and
products on the FP64 tensor path) against the oracle's blocked engine (oracle/pn_blocked.c).

DMMA.8x8x4 is, per output element, the ascending-k chain of four fma's (scripts/micro/dmma_probe.cu,
measured on B200), and every reduction of the kernel has a fixed order that the oracle restates, so
the comparison is BIT FOR BIT -- far inside the 1e-9 relative tolerance of the north star.  The
blocked oracle itself is checked against the unblocked, golden-pinned engine in tests/test_oracle_blocked.py.
"""

import ctypes as C

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


def _report(name, got, want):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, f"{name}: shape {got.shape} vs {want.shape}"
    same = (got == want) | (np.isnan(got.astype(float)) & np.isnan(want.astype(float)))
    if same.all():
        return
    bad = tuple(np.argwhere(~same)[0])
    with np.errstate(all="ignore"):
        rel = np.nanmax(np.abs(got.astype(float) - want) / np.maximum(np.abs(want), 1e-300))
    raise AssertionError(f"{name}: {(~same).sum()} of {same.size} entries differ, first at {bad}: gpu={got[bad]!r} "
                         f"oracle={want[bad]!r}; max rel diff {rel:.3e}")  # fmt: skip


def _dev(x):
    import torch

    return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64), device="cuda")


def _ptr(t):
    return C.c_void_p(t.data_ptr())


# ---- the blocked primitives, one at a time ------------------------------------------------------------


@pytest.mark.parametrize("M,N,K,kmajor,neg,init", [(64, 64, 64, 0, 0, 0), (80, 80, 80, 0, 0, 0), (80, 80, 80, 1, 0, 0), (16, 80, 80, 1, 0, 0),
                                                   (80, 80, 16, 0, 1, 1), (100, 37, 53, 0, 1, 1), (37, 100, 53, 1, 0, 0), (130, 130, 7, 0, 1, 1)])  # fmt: skip
def test_dmma_gemm_is_the_ascending_fma_chain(cabi, oracle, M, N, K, kmajor, neg, init):
    rng = np.random.default_rng(M * 1000 + N * 10 + K)
    A = rng.standard_normal((K, M) if kmajor else (M, K))
    B = rng.standard_normal((K, N))
    C0 = rng.standard_normal((M, N))
    dA, dB, dC0 = _dev(A), _dev(B), _dev(C0)
    import torch

    dC = torch.zeros((M, N), dtype=torch.float64, device="cuda")
    rc = cabi.lib().pn_b200_selftest_gemm(_ptr(dC), N, _ptr(dA), A.shape[1], kmajor, _ptr(dB), N, M, N, K,
                                          _ptr(dC0) if init else None, N, neg)  # fmt: skip
    assert rc == 0
    want = oracle.gemm_chain(A.T if kmajor else A, B, C0 if init else None, bool(neg))
    _report("C", dC.cpu().numpy(), want)


@pytest.mark.parametrize("rows,cols,ncols,shape,ntop", [
    (80, 80, 80, "full", 0), (80, 16, 16, "full", 0), (32, 16, 16, "full", 0), (100, 37, 37, "full", 0),
    (160, 160, 80, "toptri_botfull", 80), (160, 80, 80, "toptri_botfull", 80), (160, 80, 80, "topfull_bottri", 80),
    (60, 30, 30, "topfull_bottri", 30), (60, 60, 30, "toptri_botfull", 30), (320, 320, 160, "toptri_botfull", 160),
])  # fmt: skip
def test_blocked_qr_bitwise(cabi, oracle, rows, cols, ncols, shape, ntop):
    rng = np.random.default_rng(rows * 7 + cols)
    M = rng.standard_normal((rows, cols))
    if shape == "toptri_botfull":  # top block: upper triangular in its first ntop columns
        M[:ntop, :ntop] = np.triu(M[:ntop, :ntop])
        M[:ntop, ntop:] = 0.0 if cols > ntop else M[:ntop, ntop:]
    if shape == "topfull_bottri":  # bottom block: upper triangular
        M[ntop:, :] = np.triu(M[ntop:, :])
    want = oracle.qr_blocked(M, ncols, shape, ntop, 16)
    dM = _dev(M)
    assert cabi.lib().pn_b200_selftest_qr(_ptr(dM), cols, rows, cols, ncols, oracle.QR_SHAPES[shape], ntop) == 0
    got = dM.cpu().numpy()
    k = min(rows, cols, ncols)
    # R rows (upper triangle of the first k rows, all columns); entries below the diagonal are not cleared on the GPU
    mask = np.triu(np.ones((k, cols), dtype=bool))
    _report("R", got[:k][mask], want[:k][mask])
    if cols > ncols:  # partial QR: the rows below also carry the reflected trailing columns
        _report("trailing block", got[k:, ncols:], want[k:, ncols:])
    # and it is a QR of M: R^T R = M^T M on the triangularised columns
    R = np.triu(want[:k, :ncols])
    np.testing.assert_allclose(R.T @ R, M[:, :ncols].T @ M[:, :ncols], rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("n,c", [(80, 80), (64, 16), (130, 70), (30, 30)])
def test_blocked_back_substitution_bitwise(cabi, oracle, n, c):
    rng = np.random.default_rng(n + c)
    R = np.triu(rng.standard_normal((n, n))) + 4 * np.eye(n)
    B = rng.standard_normal((n, c))
    want = oracle.solve_upper_blocked(R, B, 64)
    np.testing.assert_allclose(R @ want, B, rtol=1e-8, atol=1e-8)
    import torch

    dR, dB = _dev(R), _dev(B)
    dX = torch.zeros((n, c), dtype=torch.float64, device="cuda")
    assert cabi.lib().pn_b200_selftest_trsm(_ptr(dR), n, _ptr(dB), c, _ptr(dX), c, n, c) == 0
    _report("X", dX.cpu().numpy(), want)


# ---- the solver ---------------------------------------------------------------------------------------


def _u0(N):
    return np.concatenate([np.sin(2 * np.pi * np.linspace(0, 1, N)) + 1, 3 * np.ones(N)])[None, :]


def _desc(cabi, N, nu, B, K, *, corr="ts1", strat="fixedpoint", calib="dynamic", atol=1e-5, rtol=1e-5, dt0=0.01, flags=0):
    return cabi.Desc(cabi.PROBLEM_IDS["brusselator"], 2 * N, nu, 1, cabi.FACTORISATIONS["dense"], cabi.CORRECTIONS[corr],
                     cabi.STRATEGIES[strat], cabi.CALIBRATIONS[calib], atol, rtol, dt0, 0.95, 0.2, 10.0, 0.3, 0.4,
                     B, K, 0, 1, flags, 0)  # fmt: skip


def _ocfg(oracle, N, nu, *, corr="ts1", strat="fixedpoint", calib="dynamic", atol=1e-5, rtol=1e-5, dt0=0.01):
    return oracle.make_config("brusselator", 2 * N, nu, 1, factorisation="dense", correction=corr, strategy=strat,
                              calibration=calib, atol=atol, rtol=rtol, dt0=dt0, num_params=1, reduction_group=256,
                              dense_block=16)  # fmt: skip


@pytest.mark.parametrize("N,corr", [(8, "ts1"), (8, "ts0"), (5, "ts1")])
def test_fixed_grid_filter_steps_bitwise(cabi, oracle, N, corr):
    """solve_fixed_grid (vdp.py:88-91) with the filter strategy: three steps, no error control."""
    grid = np.array([0.0, 0.01, 0.03, 0.06])
    kw = dict(corr=corr, strat="filter")
    desc = _desc(cabi, N, 4, 1, len(grid), flags=cabi.FLAG_FIXED_GRID, **kw)
    gpu = cabi.solve_host(desc, _u0(N)[None], np.array([[0.02]]), None, grid, None)
    ora = oracle.solve_fixed_grid(_ocfg(oracle, N, 4, **kw), _u0(N), [0.02], grid)
    assert gpu["status"][0] == 0
    _report("u", gpu["u"][0], ora["u"])
    _report("u_std", gpu["u_std"][0], ora["u_std"])


@pytest.mark.parametrize("N,nu,corr,strat,calib,K,t1,tol", [
    (8, 4, "ts1", "fixedpoint", "dynamic", 5, 0.5, 1e-4),
    (8, 4, "ts0", "fixedpoint", "none", 4, 0.3, 1e-4),
    (8, 4, "ts1", "filter", "dynamic", 4, 0.3, 1e-4),
    (8, 2, "ts1", "fixedpoint", "dynamic", 4, 0.3, 1e-3),
    (16, 4, "ts1", "fixedpoint", "dynamic", 4, 0.2, 1e-4),
    (32, 4, "ts1", "fixedpoint", "dynamic", 3, 0.05, 1e-3),
])  # fmt: skip
def test_adaptive_checkpoint_solve_bitwise(cabi, oracle, N, nu, corr, strat, calib, K, t1, tol):
    """solve_adaptive_save_at + backward marginalisation, dense EKF0 / EKF1 at D = (nu+1) 2N = 48 ... 320:
    smoothed means, standard deviations, full marginals and accepted / rejected counts."""
    save_at = np.linspace(0.0, t1, K)
    kw = dict(corr=corr, strat=strat, calib=calib, atol=tol, rtol=tol)
    desc = _desc(cabi, N, nu, 1, K, **kw)
    gpu = cabi.solve_host(desc, _u0(N)[None], np.array([[0.02]]), None, save_at, None, full=True)
    ora = oracle.solve_save_at(_ocfg(oracle, N, nu, **kw), _u0(N), [0.02], save_at, full=True)
    assert ora["status"] == 0 and ora["n_accepted"][-1] >= K - 1
    assert gpu["status"][0] == 0
    _report("n_accepted", gpu["n_accepted"][0], ora["n_accepted"])
    assert gpu["n_rejected"][0] == ora["n_rejected"]
    _report("u", gpu["u"][0], ora["u"])
    _report("u_std", gpu["u_std"][0], ora["u_std"])
    D = (nu + 1) * 2 * N
    _report("marg_mean", gpu["marg_mean"][0].reshape(K, D), ora["marg_mean"].reshape(K, D))
    _report("marg_chol", gpu["marg_chol"][0].reshape(K, D, D), ora["marg_chol"].reshape(K, D, D))


def test_dense_ekf1_brusselator_n64_three_steps(cabi, oracle):
    """D = 640 (Brusselator N = 64, BASELINE config 5): a short adaptive solve, bit for bit."""
    N, K = 64, 2
    save_at = np.array([0.0, 0.02])
    kw = dict(atol=1e-2, rtol=1e-2)
    desc = _desc(cabi, N, 4, 1, K, **kw)
    gpu = cabi.solve_host(desc, _u0(N)[None], np.array([[0.02]]), None, save_at, None)
    ora = oracle.solve_save_at(_ocfg(oracle, N, 4, **kw), _u0(N), [0.02], save_at)
    assert ora["status"] == 0 and gpu["status"][0] == 0
    _report("n_accepted", gpu["n_accepted"][0], ora["n_accepted"])
    assert gpu["n_rejected"][0] == ora["n_rejected"]
    _report("u", gpu["u"][0], ora["u"])
    _report("u_std", gpu["u_std"][0], ora["u_std"])


def test_ensemble_over_the_diffusion_parameter(cabi, oracle):
    """BASELINE config 5's ensemble: members differ in alpha; more members than one CTA each is needed to
    exercise the work queue, per-member tolerances and the longest-first launch order."""
    N, K, B = 8, 4, 5
    rng = np.random.default_rng(3)
    alpha = (1.0 / 50.0) * 10.0 ** rng.uniform(-0.5, 0.5, B)
    tol = np.stack([10.0 ** -rng.integers(3, 6, B), 10.0 ** -rng.integers(3, 6, B)], 1).astype(float)
    save_at = np.linspace(0.0, 0.3, K)
    desc = _desc(cabi, N, 4, B, K)
    u0 = np.tile(_u0(N)[None], (B, 1, 1))
    gpu = cabi.solve_host(desc, u0, alpha[:, None], tol, save_at, None)
    ora = oracle.solve_save_at_batch(_ocfg(oracle, N, 4), u0, alpha[:, None], save_at, tol=tol)
    assert (ora["status"] == 0).all()
    for key in ("status", "n_accepted", "n_rejected", "u", "u_std"):
        _report(key, gpu[key], ora[key])


def test_accuracy_against_a_reference_solution(cabi):
    """Independent of the oracle: the dense EKF1 solution of the Brusselator (N = 8) agrees with scipy's DOP853."""
    from scipy.integrate import solve_ivp

    N, K = 8, 6
    c = 0.02 * (N + 1) ** 2

    def f(t, y):
        u, v = y[:N], y[N:]
        up, vp = np.concatenate([[1.0], u, [1.0]]), np.concatenate([[3.0], v, [3.0]])
        return np.concatenate([1 + u * u * v - 4 * u + c * (up[:-2] - 2 * u + up[2:]), 3 * u - u * u * v + c * (vp[:-2] - 2 * v + vp[2:])])

    save_at = np.linspace(0.0, 1.0, K)
    truth = solve_ivp(f, (0, 1), _u0(N)[0], method="DOP853", t_eval=save_at, rtol=1e-12, atol=1e-12).y.T
    gpu = cabi.solve_host(_desc(cabi, N, 4, 1, K, atol=1e-7, rtol=1e-7), _u0(N)[None], np.array([[0.02]]), None, save_at, None)
    assert gpu["status"][0] == 0
    np.testing.assert_allclose(gpu["u"][0], truth, rtol=1e-5, atol=1e-6)
    # calibration sanity: the error is of the order of the reported standard deviation
    err = np.abs(gpu["u"][0][1:] - truth[1:])
    assert (err <= 50 * gpu["u_std"][0][1:] + 1e-12).all()
