"""BASELINE config 5 as stated: Brusselator, DENSE sqrt-EKF1 factorisation + checkpointed (fixed-point) smoother,
ensemble over the diffusion parameter -- the CTA-per-IVP kernel with blocked QR on the FP64 tensor path.

    python scripts/dense_cta_throughput.py [--N 8 16 32 64] [--members 148] [--attempts 0] [--reps 2]      # on a B200

One JSON line per N: CUDA-event time of solver + smoothing kernels, attempted steps/s, and the algorithmic
fp64 flop rate (20.33 D^3 + (2n + 8d) D^2 + 4 D d^2 per attempt) against the measured DFMA peak.
--attempts caps the attempted steps per member (status MAX_ATTEMPTS) so that the large sizes stay short.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "code-adaptive-prob-ode-solvers_b200"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from odecheckpts_b200 import _cabi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, nargs="+", default=[8, 16, 32, 64])
ap.add_argument("--members", type=int, default=148)
ap.add_argument("--attempts", type=int, default=0)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--tol", type=float, default=1e-6)
ap.add_argument("--t1", type=float, default=1.0)
ap.add_argument("--K", type=int, default=11)
ap.add_argument("--corr", default="ts1")
args = ap.parse_args()
dev = torch.device("cuda:0")
peak = _cabi.measure_fp64_peak()
T = lambda x: torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64, device=dev)  # noqa: E731
for N in args.N:
    d, nu, B, K = 2 * N, 4, args.members, args.K
    n, D = nu + 1, (nu + 1) * 2 * N
    rng = np.random.default_rng(3)
    alpha = (1.0 / 50.0) * 10.0 ** rng.uniform(-0.5, 0.5, B)  # seed 3, SURVEY 8d C5
    y0 = np.concatenate([np.sin(2 * np.pi * np.linspace(0, 1, N)) + 1, 3 * np.ones(N)])
    desc = _cabi.Desc(_cabi.PROBLEM_IDS["brusselator"], d, nu, 1, _cabi.FACTORISATIONS["dense"], _cabi.CORRECTIONS[args.corr],
                      1, 1, args.tol, args.tol, 0.01, 0.95, 0.2, 10.0, 0.3, 0.4, B, K, args.attempts, 1, 0, 0)  # fmt: skip
    u0, par, save = T(np.tile(y0[None, None], (B, 1, 1))), T(alpha[:, None]), T(np.linspace(0, args.t1, K))
    out, best = None, 1e30
    for _ in range(args.reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = _cabi.solve_device(desc, u0, par, None, save, None, workspace=None if out is None else out["_workspace"])
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    acc = float(out["n_accepted"][:, -1].double().sum())
    rej = float(out["n_rejected"].double().sum())
    W = 20.33 * D**3 + (2 * n + 8 * d) * D**2 + 4 * D * d**2 + 20 * N + 8 * N
    flops = (acc + rej) * W
    print(json.dumps(dict(
        config=f"C5 Brusselator N={N} (d={d}, D={D}) DENSE EK{'F1' if args.corr == 'ts1' else 'F0'} nu=4 + fixed-point smoother, tol={args.tol:g}, "
               f"{B} members over alpha, {K} checkpoints on [0, {args.t1:g}]" + (f", at most {args.attempts} attempts per member" if args.attempts else ""),
        members=B, ms=best, solves_per_s=B / best * 1e3, attempts=acc + rej, attempts_per_s=(acc + rej) / best * 1e3,
        ms_per_attempt_per_cta=best * min(B, 148) / max(acc + rej, 1), flops_per_attempt=W,
        algorithmic_tflops=flops / best / 1e9, fp64_peak_tflops=peak, frac=flops / best / 1e9 / peak,
        status_ok=int((out["status"] == 0).sum()), status_max_attempts=int((out["status"] == 2).sum()),
        kernel=_cabi.kernel_info(desc))), flush=True)  # fmt: skip
