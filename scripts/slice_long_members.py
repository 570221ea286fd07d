import os, sys, time
sys.path.insert(0, "/root/repo/code-adaptive-prob-ode-solvers_b200"); sys.path.insert(0, "/root/repo")
import numpy as np, torch, bench
from odecheckpts_b200 import _cabi
dev = torch.device("cuda:0")
B, K = 40960, 50
u0, par = bench.ensemble_inputs(0, B)
save_at = torch.linspace(bench.T0, bench.T1, K, dtype=torch.float64, device=dev)
desc = _cabi.Desc(5, 1, 4, 2, 2, 1, 1, 1, 1e-10, 1e-10, 0.01, 0.95, 0.2, 10.0, 0.3, 0.4, B, K, 0, 1, 0, 0)
res = {}
for name, env in (("unsliced", {"PN_B200_NO_SLICE": "1"}), ("sliced", {})):
    os.environ.pop("PN_B200_NO_SLICE", None); os.environ.update(env)
    for it in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = _cabi.solve_device(desc, torch.as_tensor(u0, device=dev), torch.as_tensor(par, device=dev), None, save_at, None)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    res[name] = out
    print(name, f"{dt*1e3:.0f} ms", "attempts/member", float(out["n_accepted"][:, -1].double().mean() + out["n_rejected"].double().mean()))
print("identical:", all(torch.equal(res["sliced"][k], res["unsliced"][k]) for k in ("u", "u_std", "n_accepted", "n_rejected", "status")))
