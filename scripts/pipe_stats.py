"""Where the PIPE build of the CTA-per-IVP kernel waits (needs a library built with -DPN_PIPE_STATS):
cycles the main warps wait for X, cycles the backward warp waits for a job / for the op list."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "code-adaptive-prob-ode-solvers_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
from odecheckpts_b200 import _cabi
for N in (32, 128, 512):
    d, K = 2 * N, 20
    u0 = np.concatenate([np.sin(2 * np.pi * np.linspace(0, 1, N)) + 1, 3 * np.ones(N)])
    desc = _cabi.Desc(4, d, 4, 1, 0, 0, 1, 1, 1e-8, 1e-8, 0.01, 0.95, 0.2, 10.0, 0.3, 0.4, 1, K, 0, 1, 0, 0)
    dev = torch.device("cuda:0")
    t1 = 10.0 if N < 512 else 0.5
    out = _cabi.solve_device(desc, torch.as_tensor(u0[None, None], device=dev).contiguous(), torch.full((1, 1), 0.02, dtype=torch.float64, device=dev),
                             None, torch.linspace(0, t1, K, dtype=torch.float64, device=dev), None)
    torch.cuda.synchronize()
    w = out["_workspace"].view(torch.int64)[:16].cpu().numpy()
    w = out["_workspace"].view(torch.int64)[:32].cpu().numpy()
    it = max(int(w[11]), 1)
    print("   main warp 0 phases (cycles/iter): tail+fetch %.0f | precond+pass1 %.0f | calib+predict QR+publish %.0f | correction+pass2 %.0f | enorm+cf+controller %.0f | rest of bookkeeping %.0f | X wait+load %.0f | pass 3 (accepted) %.0f | park state %.0f | empty marker %.0f | commit op+record %.0f | resolve_hits %.0f" % tuple(w[14:26] / it))
    print(f"N={N}: iterations {it}, kernel cycles/iter {w[4]/it:.0f}, main waits for X {w[8]/it:.0f}, backward warp waits for job {w[9]/it:.0f} / for ops {w[10]/it:.0f}, job -> X {w[12]/it:.0f}, merge {w[13]/it:.0f} cycles per iteration")
