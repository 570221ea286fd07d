"""experiments/4_brusselator/run.py through the builder API on the GPU: for every grid size N the baseline step
count (solve_adaptive_terminal_values, run.py:82-90), the TEXTBOOK smoother (strategy_smoother +
solve_adaptive_save_every_step, run.py:98-117: one backward conditional per accepted step, while that fits the
reference's 4000 MB budget) and the CHECKPOINT solver (strategy_fixedpoint + solve_adaptive_save_at with 200
checkpoints, run.py:119-138) -- the paper's memory / run-time comparison (Fig. 4).

    python scripts/brusselator_run.py [max N = 128]        # on a B200
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "code-adaptive-prob-ode-solvers_b200"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import ctypes  # noqa: E402

from odecheckpts_b200 import _cabi, ivps  # noqa: E402
from odecheckpts_b200.probdiffeq import impl, ivpsolve, taylor  # noqa: E402
from odecheckpts_b200.probdiffeq import ivpsolvers as pdi  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "reference_goldens.npz"))
max_n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
num, tol = 4, 1e-8


def workspace_mb(N, K, strategy):
    """Device workspace of one solve with K checkpoints (the backward conditionals dominate it)."""
    desc = _cabi.Desc(4, 2 * N, num, 1, 0, 0, strategy, 1, tol, tol, 0.01, 0.95, 0.2, 10.0, 0.3, 0.4, 1, K, 0, 1, 0, 0)
    fn = _cabi.lib().pn_b200_workspace_bytes
    fn.restype = ctypes.c_size_t
    return fn(ctypes.byref(desc)) / 1024**2


def timed(fn):
    torch.cuda.synchronize()
    t = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, time.perf_counter() - t


for idx, N in enumerate(g["brusselator_N"]):
    N = int(N)
    if N > max_n:
        break
    vf, u0, (t0, t1), params = ivps.brusselator(N=N)
    impl.impl.select("isotropic", ode_shape=(2 * N,))
    ctrl = ivpsolve.control_proportional_integral()
    ibm, ts0 = pdi.prior_ibm(num_derivatives=num), pdi.correction_ts0(ode_order=1)
    f = lambda *y, t=None: vf(*y, t=t, p=params)  # noqa: E731
    tcoeffs = taylor.odejet_unroll(lambda *y: vf(*y, t=t0, p=params), u0, num=num)
    solver = pdi.solver_dynamic(pdi.strategy_fixedpoint(ibm, ts0))
    init = solver.initial_condition(tcoeffs, 1.0)
    asolver = ivpsolve.adaptive(solver, atol=tol, rtol=tol, control=ctrl)
    # the reference's memory model: three copies of the state per accepted step (run.py:70-77)
    n = num + 1
    size_init = 3 * 8 * (n * 2 * N + 3 * n * n + n * 2 * N + 1)
    base, t_base = timed(lambda: ivpsolve.solve_adaptive_terminal_values(f, init, t0=t0, t1=t1, dt0=0.01, adaptive_solver=asolver))
    steps = int(base.num_steps)
    mem_text = steps * size_init / 1024**2
    line = f"N={N:4d}: baseline {steps:8d} steps (golden {int(g['brusselator_num_steps_terminal'][idx]):8d}) in {t_base:7.3f}s, textbook memory {mem_text:9.0f} MB; "
    if mem_text < 4000:
        solver_s = pdi.solver_dynamic(pdi.strategy_smoother(ibm, ts0))
        asolver_s = ivpsolve.adaptive(solver_s, atol=tol, rtol=tol, control=ctrl)
        init_s = solver_s.initial_condition(tcoeffs, 1.0)
        text, t_text = timed(lambda: ivpsolve.solve_adaptive_save_every_step(f, init_s, t0=t0, t1=t1, dt0=0.01, adaptive_solver=asolver_s,
                                                                             max_steps=steps + 16))  # fmt: skip
        line += f"textbook smoother {t_text:7.3f}s, {workspace_mb(N, len(text.t), 1):8.1f} MB of conditionals on the device ({len(text.t)} grid points); "
    else:
        line += "textbook smoother skipped (over the reference's 4000 MB budget); "
    save_at = np.linspace(t0, t1, 200)
    ck, t_ck = timed(lambda: ivpsolve.solve_adaptive_save_at(f, init, save_at=save_at, dt0=0.01, adaptive_solver=asolver, keep_conditionals=True))
    peak_ck = workspace_mb(N, 200, 1)
    ok = int(np.max(ck.num_steps)) == int(g["brusselator_num_steps_checkpoint"][idx])
    line += f"checkpoint solver {t_ck:7.3f}s, {peak_ck:7.1f} MB of conditionals, steps {'==' if ok else '!='} golden"
    print(line, flush=True)
