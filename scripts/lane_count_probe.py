"""How the time of one warp's pass over the headline workload depends on its number of active lanes (B200)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "code-adaptive-prob-ode-solvers_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
from odecheckpts_b200 import _cabi
K = 50
dev = torch.device("cuda:0")
for pair in ("1", "0"):
    os.environ["PN_B200_PAIR"] = pair
    for B in (1, 2, 8, 16, 17, 24, 32, 64, 128):
        u0 = np.tile(np.array([[[2.0], [0.0]]]), (B, 1, 1))  # identical members: no divergence of any kind
        rng = np.random.default_rng(0)
        u1 = np.stack([2.0 + 0.5 * rng.uniform(-1, 1, B), 0.5 * rng.uniform(-1, 1, B)], 1).reshape(B, 2, 1)
        desc = _cabi.Desc(5, 1, 4, 2, 2, 1, 1, 1, 1e-6, 1e-6, 0.01, 0.95, 0.2, 10.0, 0.3, 0.4, B, K, 0, 1, 0, 0)
        par = torch.full((B, 1), 1e3, dtype=torch.float64, device=dev)
        save = torch.linspace(0, 6.3, K, dtype=torch.float64, device=dev)
        res = []
        for u in (u0, u1):
            u_d = torch.as_tensor(u, device=dev)
            best, out = 1e30, None
            for it in range(3):
                torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(); out = _cabi.solve_device(desc, u_d, par, None, save, None, workspace=None if out is None else out["_workspace"]); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            att = (out["n_accepted"][:, -1] + out["n_rejected"]).max().item()
            res.append((best, att))
        print(f"PAIR={pair} members {B:4d}: identical members {res[0][0]:7.2f} ms ({res[0][1]} attempts)   randomised members {res[1][0]:7.2f} ms (max {res[1][1]} attempts)", flush=True)
