"""experiments/4_brusselator/run.py on the GPU: isotropic EKF0 nu=4, tol 1e-8, 200 checkpoints,
N = 2 ... 512; accepted-step counts vs the reference's goldens and wall time per solve."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "code-adaptive-prob-ode-solvers_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
from odecheckpts_b200 import _cabi

g = np.load(os.path.join(ROOT, "tests", "golden", "reference_goldens.npz"))
ref_runtime = [0.70, 0.70, 0.71, 0.82, 1.18, 2.34, 7.18, 35.0, 221.8]
maxN = int(sys.argv[1]) if len(sys.argv) > 1 else 256
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
K = 200
for idx, N in enumerate(g["brusselator_N"]):
    N = int(N)
    if N > maxN: break
    d = 2 * N
    u0 = np.concatenate([np.sin(2 * np.pi * np.linspace(0, 1, N)) + 1, 3 * np.ones(N)])
    desc = _cabi.Desc(4, d, 4, 1, 0, 0, 1, 1, 1e-8, 1e-8, 0.01, 0.95, 0.2, 10.0, 0.3, 0.4, B, K, 0, 1, 0, 0)
    dev = torch.device("cuda:0")
    u0_d = torch.as_tensor(np.tile(u0[None, None], (B, 1, 1)), device=dev).contiguous()
    par = torch.full((B, 1), 1.0 / 50.0, dtype=torch.float64, device=dev)
    save = torch.linspace(0, 10, K, dtype=torch.float64, device=dev)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = _cabi.solve_device(desc, u0_d, par, None, save, None)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    acc = int(out["n_accepted"][0, -1]); want = int(g["brusselator_num_steps_checkpoint"][idx])
    print(f"N={N:4d} d={d:5d} B={B}: accepted={acc:8d} golden={want:8d} ({100*(acc-want)/want:+.2f}%) rejected={int(out['n_rejected'][0]):6d} "
          f"status={int(out['status'][0])} gpu={dt:8.3f}s  reference(JAX CPU, incl. jit)={ref_runtime[idx]}s  steps/s={acc/dt:.3e}", flush=True)
