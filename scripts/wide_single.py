"""One Brusselator IVP on the CTA-per-IVP isotropic kernel (ncu target): python scripts/wide_single.py [N=64] [t1=10]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "code-adaptive-prob-ode-solvers_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
from odecheckpts_b200 import _cabi
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
t1 = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
d, K = 2 * N, 200
u0 = np.concatenate([np.sin(2 * np.pi * np.linspace(0, 1, N)) + 1, 3 * np.ones(N)])
desc = _cabi.Desc(4, d, 4, 1, 0, 0, 1, 1, 1e-8, 1e-8, 0.01, 0.95, 0.2, 10.0, 0.3, 0.4, 1, K, 0, 1, 0, 0)
dev = torch.device("cuda:0")
args = (torch.as_tensor(u0[None, None], device=dev).contiguous(), torch.full((1, 1), 0.02, dtype=torch.float64, device=dev), None,
        torch.linspace(0, t1, K, dtype=torch.float64, device=dev), None)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = _cabi.solve_device(desc, *args)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
acc = int(out["n_accepted"][0, -1]); rej = int(out["n_rejected"][0])
print(f"N={N} d={d}: {acc} accepted, {rej} rejected, {dt:.3f} s, {1e6 * dt / (acc + rej):.2f} us per attempted step")
