"""Phase statistics of the PAIR build (library built with -DPN_PIPE_STATS): cycles per iteration of filter lane 0."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "code-adaptive-prob-ode-solvers_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
from odecheckpts_b200 import _cabi
B, K = 1, 50
desc = _cabi.Desc(5, 1, 4, 2, 2, 1, 1, 1, 1e-6, 1e-6, 0.01, 0.95, 0.2, 10.0, 0.3, 0.4, B, K, 0, 1, 0, 0)
dev = torch.device("cuda:0")
out = _cabi.solve_device(desc, torch.tensor([[[2.0], [0.0]]], dtype=torch.float64, device=dev), torch.full((B, 1), 1e3, dtype=torch.float64, device=dev),
                         None, torch.linspace(0, 6.3, K, dtype=torch.float64, device=dev), None)
torch.cuda.synchronize()
w = out["_workspace"].view(torch.int64)[:32].cpu().numpy()
it = max(int(w[11]), 1)
print("iterations", it, "accepted", int(out["n_accepted"][0, -1]), "kernel cycles/iter %.0f" % (w[4] / it))
print("filter lane 0 phases (cycles/iter): tail+fetch+barriers %.0f | precond+mean+vf %.0f | calib+left QR+mailbox %.0f | correction+error norm %.0f | controller+cf %.0f | bookkeeping %.0f" % tuple(w[14:20] / it))
