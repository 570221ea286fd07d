"""Throughput of the other BASELINE configs (parity-test cases, not the bench line): SURVEY 8d C3/C4/C5.

    python scripts/config_throughput.py [--big] [--only SUBSTRING]      # on a B200
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "code-adaptive-prob-ode-solvers_b200"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from odecheckpts_b200 import _cabi  # noqa: E402

dev = torch.device("cuda:0")


def T(x):
    return torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64, device=dev)


ONLY = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None


def run(name, desc, u0, params, tol, save_at, reps=3):
    if ONLY is not None and ONLY not in name:
        return
    args = (T(u0), None if params is None else T(params), None if tol is None else T(tol), T(save_at), None)
    out = None
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        out = _cabi.solve_device(desc, *args, workspace=None if out is None else out["_workspace"])
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    acc = float(out["n_accepted"][:, -1].double().sum())
    rej = float(out["n_rejected"].double().sum())
    bad = int((out["status"] != 0).sum())
    rec = dict(config=name, members=int(desc.batch), ms=best, solves_per_s=desc.batch / best * 1e3,
               accepted_steps_per_s=acc / best * 1e3, attempts_per_s=(acc + rej) / best * 1e3, failed=bad)  # fmt: skip
    print(json.dumps(rec), flush=True)


def desc(problem, d, nu, q, B, K, fact, corr, atol, rtol, dt0, P, calib=1):
    return _cabi.Desc(_cabi.PROBLEM_IDS[problem], d, nu, q, _cabi.FACTORISATIONS[fact], _cabi.CORRECTIONS[corr], 1, calib,
                      atol, rtol, dt0, 0.95, 0.2, 10.0, 0.3, 0.4, B, K, 0, P, 0, 0)  # fmt: skip


# C3: rigid body, 2,048 initial conditions x 8 tolerances in ONE launch (per-member tol)
rng = np.random.default_rng(1)
n_ic = 2048
tols = 10.0 ** -np.arange(3, 11)
u0 = np.array([1.0, 0.0, 0.9]) + 0.05 * rng.standard_normal((n_ic, 3))
u0 = np.repeat(u0[:, None, :], len(tols), 0).reshape(-1, 1, 3)
t = np.tile(tols * 100, n_ic)
tol = np.stack([1e-3 * t, t], 1)
B = len(u0)
par = np.tile([-2.0, 1.25, -0.5], (B, 1))
xs = np.linspace(0, 50, 5)
for nu in (2, 4):
    run(f"C3 rigid body isotropic EKF0 nu={nu}, 16,384 = 2,048 ICs x 8 tolerances (1e-3..1e-10)",
        desc("rigid_body", 3, nu, 1, B, 5, "isotropic", "ts0", 1e-6, 1e-6, 50.0, 3), u0, par, tol, xs)  # fmt: skip
run("C3 rigid body dense EKF1 nu=4, same 16,384 members",
    desc("rigid_body", 3, 4, 1, B, 5, "dense", "ts1", 1e-6, 1e-6, 50.0, 3), u0, par, tol, xs, reps=2)  # fmt: skip

# C4: Pleiades, 2,048 x 8 tolerances, blockdiag and isotropic, nu = 3..5
rng = np.random.default_rng(2)
x = np.array([3.0, 3.0, -1.0, -3.0, 2.0, -2.0, 2.0, 3.0, -3.0, 2.0, 0.0, 0.0, -4.0, 4.0])
dx = np.array([0, 0, 0, 0, 0, 1.75, -1.5, 0, 0, 0, -1.25, 1.0, 0, 0.0])
pos = x + 0.01 * rng.standard_normal((n_ic, 14))
u0 = np.stack([pos, np.tile(dx, (n_ic, 1))], 1)
u0 = np.repeat(u0[:, None], len(tols), 0).reshape(-1, 2, 14)
t = np.tile(tols * 10, n_ic)
tol = np.stack([1e-3 * t, t], 1)
B = len(u0)
for fact in ("blockdiag", "isotropic"):
    for nu in (3, 5):
        run(f"C4 Pleiades {fact} EKF0 nu={nu}, 16,384 = 2,048 ICs x 8 tolerances",
            desc("pleiades", 14, nu, 2, B, 50, fact, "ts0", 1e-6, 1e-6, 0.1, 0), u0, None, tol, np.linspace(0, 3, 50), reps=2)  # fmt: skip

# C5: Brusselator ensemble over the diffusion parameter
rng = np.random.default_rng(3)
for N, B in ((16, 1024), (128, 296), (128, 1184), (512, 148)):
    alpha = (1.0 / 50.0) * 10.0 ** rng.uniform(-0.5, 0.5, B)
    y0 = np.concatenate([np.sin(2 * np.pi * np.linspace(0, 1, N)) + 1, 3 * np.ones(N)])
    if N == 512 and "--big" not in sys.argv:
        continue
    run(f"C5 Brusselator N={N} (d={2 * N}) isotropic EKF0 nu=4 tol=1e-8, {B} members over alpha, 200 checkpoints",
        desc("brusselator", 2 * N, 4, 1, B, 200, "isotropic", "ts0", 1e-8, 1e-8, 0.01, 1),
        np.tile(y0[None, None], (B, 1, 1)), alpha[:, None], None, np.linspace(0, 10, 200), reps=1)  # fmt: skip
