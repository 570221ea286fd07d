"""experiments/1_van_der_pol/vdp.py on the GPU: adaptive baseline (save every step), the same grid replayed
with solve_fixed_grid, a uniform grid with as many points (must blow up), and the uniform grid of equal
accuracy (T / min step points).  Reference numbers (JAX CPU, second jit-compiled call):
vdp_runtime_adaptive.npy = 0.0227 s, vdp_runtime_fixed_accurate.npy = 6.50 s.

    python scripts/vdp_fig1.py            # on a B200
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "code-adaptive-prob-ode-solvers_b200"))
import numpy as np  # noqa: E402

from odecheckpts_b200 import ivps  # noqa: E402
from odecheckpts_b200.probdiffeq import impl, ivpsolve, ivpsolvers, taylor  # noqa: E402


def timed(fn, reps=3):
    fn()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t0)
    return out, best


vf, (u0, du0), (t0, t1) = ivps.van_der_pol(mu=1e3)
impl.impl.select("dense", ode_shape=(1,))
ibm = ivpsolvers.prior_ibm(num_derivatives=4)
solver = ivpsolvers.solver_dynamic(ivpsolvers.strategy_filter(ibm, ivpsolvers.correction_ts1(ode_order=2)))
init = solver.initial_condition(taylor.odejet_padded_scan(lambda *y: vf(*y, t=t0), [u0, du0], num=3), 1.0)
asolver = ivpsolve.adaptive(solver, atol=1e-3, rtol=1e-3, control=ivpsolve.control_proportional_integral())
base, t_base = timed(lambda: ivpsolve.solve_adaptive_save_every_step(vf, init, t0=t0, t1=t1, dt0=0.01, adaptive_solver=asolver))
steps = np.diff(base.t)
required = int((t1 - t0) / steps.min())
print(f"adaptive baseline: {len(base.t)} grid points (reference golden 2912), min step {steps.min():.3e}, max {steps.max():.3e}; {t_base * 1e3:.1f} ms")
replay, t_replay = timed(lambda: ivpsolve.solve_fixed_grid(vf, init, grid=base.t, solver=solver))
print(f"solve_fixed_grid on the adaptive grid: {t_replay * 1e3:.2f} ms = {t_replay / len(base.t) * 1e6:.2f} us/step  (reference 22.7 ms = 7.8 us/step)")
bad, _ = timed(lambda: ivpsolve.solve_fixed_grid(vf, init, grid=np.linspace(t0, t1, len(base.t)), solver=solver), reps=1)
print(f"uniform grid with as many points: NaN = {bool(np.isnan(bad.u).any())} (the reference asserts it blows up)")
acc, t_acc = timed(lambda: ivpsolve.solve_fixed_grid(vf, init, grid=np.linspace(t0, t1, required), solver=solver), reps=2)
err = float(np.abs(acc.u[-1, 0] - replay.u[-1, 0]))
print(f"uniform grid of equal accuracy: {required} points: {t_acc:.3f} s = {t_acc / required * 1e6:.2f} us/step  (reference 743,181 points, 6.50 s = 8.7 us/step); |u(t1) - adaptive| = {err:.2e}")
