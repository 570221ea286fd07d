"""Quick GPU sanity + timing of the headline ensemble (not the bench)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "code-adaptive-prob-ode-solvers_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
from odecheckpts_b200 import _cabi

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
STRAT = int(sys.argv[2]) if len(sys.argv) > 2 else 1
NUQ = int(sys.argv[3]) if len(sys.argv) > 3 else 4
K = 50
rng = np.random.default_rng(0)
u0 = np.stack([2.0 + 0.5 * rng.uniform(-1, 1, B), 0.5 * rng.uniform(-1, 1, B)], 1).reshape(B, 2, 1)
desc = _cabi.Desc(5, 1, NUQ, 2, 2, 1, STRAT, 1, 1e-6, 1e-6, 0.01, 0.95, 0.2, 10.0, 0.3, 0.4, B, K, 0, 1, 0, 0)
print("kernel info", _cabi.kernel_info(desc))
print("fp64 peak TF", _cabi.measure_fp64_peak())
dev = torch.device("cuda:0")
u0_d = torch.as_tensor(u0, device=dev); par = torch.full((B, 1), 1e3, dtype=torch.float64, device=dev)
save = torch.linspace(0, 6.3, K, dtype=torch.float64, device=dev)
out = None
for it in range(int(os.environ.get("PN_QUICK_ITERS", "3"))):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); out = _cabi.solve_device(desc, u0_d, par, None, save, None, workspace=None if out is None else out["_workspace"]); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    acc = out["n_accepted"][:, -1].double(); rej = out["n_rejected"].double()
    att = (acc.sum() + rej.sum()).item()
    st = out["_workspace"][:16].cpu().numpy()
    if st[9] > 0:
        print(f"   slice: claims={st[9]:.3e} avg {st[8]/max(st[9],1):.0f} cyc; restores={st[11]:.3e} avg {st[10]/max(st[11],1):.0f} cyc; leaves={st[13]:.3e} avg {st[12]/max(st[13],1):.0f} cyc, parked={st[14]:.3e}")
    print(f"   stats: warp_iters={st[1]:.4e} lane_iters={st[2]:.4e} util={st[2]/(32*st[1]):.3f} interp_frac={st[3]/st[2]:.4f} max_warp_cycles={st[4]:.4e} -> {st[4]/ms/1e6:.3f} GHz-equivalent")
    print(f"iter {it}: {ms:.2f} ms  solves/s={B/ms*1e3:.0f}  attempts={att:.3e}  attempts/s={att/ms*1e3:.3e} acc mean={acc.mean().item():.1f} rej mean={rej.mean().item():.1f} status_bad={(out['status']!=0).sum().item()}")
print("fp64 peak TF", _cabi.measure_fp64_peak())
