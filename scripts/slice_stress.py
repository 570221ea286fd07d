"""Stress the time-sliced scheduler: many launches with small quanta on an ensemble just above the resident
lanes, every result compared bit for bit with the unsliced launch.   python scripts/slice_stress.py [reps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "code-adaptive-prob-ode-solvers_b200"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from odecheckpts_b200 import _cabi  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda:0")
K = bench.K_CHECKPOINTS
save_at = torch.linspace(bench.T0, bench.T1, K, dtype=torch.float64, device=dev)


def run(B, env):
    for k in ("PN_B200_NO_SLICE", "PN_B200_SLICE_QUANTUM"):
        os.environ.pop(k, None)
    os.environ.update(env)
    u0, par = bench.ensemble_inputs(0, B)
    desc = _cabi.Desc(5, 1, 4, 2, 2, 1, 1, 1, bench.TOL, bench.TOL, 0.01, 0.95, 0.2, 10.0, 0.3, 0.4, B, K, 0, 1, 0, 0)
    t0 = time.perf_counter()
    out = _cabi.solve_device(desc, torch.as_tensor(u0, device=dev), torch.as_tensor(par, device=dev), None, save_at, None)
    torch.cuda.synchronize()
    return out, time.perf_counter() - t0


for B in (38000, 40960, 65536):
    ref, _ = run(B, {"PN_B200_NO_SLICE": "1"})
    for env in ({"PN_B200_SLICE_QUANTUM": "256"}, {"PN_B200_SLICE_QUANTUM": "1024"},
                {"PN_B200_SLICE_QUANTUM": "64"}, {}):
        worst = 0.0
        for r in range(reps):
            out, dt = run(B, env)
            worst = max(worst, dt)
            for key in ("u", "u_std", "n_accepted", "n_rejected", "status"):
                assert torch.equal(out[key], ref[key]), (B, env, r, key)
        print(f"B={B} {env}: {reps} launches identical to the unsliced one, slowest {worst * 1e3:.0f} ms", flush=True)
print("OK")
