"""bench.py -- headline benchmark: Van der Pol ensemble of 65,536 randomised initial conditions.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--members B]

Workload (BASELINE.json configs[1]): stiff Van der Pol (mu=1e3), dense EKF1
(ode_order=2) + fixed-point smoother at 50 checkpoints on [0, 6.3], nu=4, dynamic calibration,
atol=rtol=1e-6, dt0=0.01; member b starts at u0 = 2 + 0.5 U(-1,1), u'0 = 0.5 U(-1,1) (seed 0).
A "step" is one pass of the hot path over the whole ensemble: the persistent solver kernel + the
smoothing sweep (N>1: one ensemble shard of 65,536 members per rank -- weak scaling -- followed by
one NCCL all-gather of the checkpoint results and step statistics).

One JSON line on stdout (rank 0).  `value` = IVP solves/s with inputs resident in HBM; `e2e` = the
same through the public Python API with HOST buffers (H2D + solve + D2H inside the timed region);
`roofline` = the solver kernel's algorithmic fp64 flop rate against the DFMA peak measured on this
GPU by a register-resident FMA-chain kernel (MEASURED_PEAKS.json carries no fp64 entry; the path is
fp64-CUDA-core bound, not HBM or tensor bound -- `roofline.hbm` shows the HBM side for context);
`cpu_baseline` = the CPU oracle (C port of the reference algorithm; jax/probdiffeq cannot be
installed here) on the host cores on a bounded sample of the same members.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "code-adaptive-prob-ode-solvers_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

MEMBERS = 65536
K_CHECKPOINTS = 50
T0, T1 = 0.0, 6.3
TOL = 1e-6
MU = 1e3
NU = 4
METRIC = "ivp_solves_per_s"
UNIT = "IVP-solves/s"

# algorithmic fp64 flops (SURVEY 8d / BASELINE.md 3): one attempted step of dense EKF1 +
# fixed-point, D = n*d = 5, d = 1, n = 5:  20.33 D^3 + (2n + 8d) D^2 + 4 D d^2 + F_f + F_J
N_, D_ = NU + 1, 1
W_ATTEMPT = 20.33 * (N_ * D_) ** 3 + (2 * N_ + 8 * D_) * (N_ * D_) ** 2 + 4 * (N_ * D_) * D_**2 + 8 + 8
W_CHECKPOINT = 1.3 * W_ATTEMPT                      # two extra predictions + one marginalisation
W_SWEEP_PER_K = 5.33 * N_**3 + 2 * N_**2 * D_       # one backward marginalisation
# DRAM traffic of one solver-kernel launch on the headline workload (ncu, profiles/r01_scalar_kernel_final_ncu.txt):
# 0.213 GB read + 1.286 GB written = the checkpoint conditionals of the fixed-point smoother
# (65,536 members x 49 checkpoints x 65 doubles = 1.67 GB, part of it still in L2 at kernel end).
NCU_DRAM_BYTES_PER_LAUNCH = 312.109056e6 + 1.402017e9  # read + write, profiles/r01_scalar_kernel_final_ncu.txt


def ensemble_inputs(first, count, stride=1):
    """Members first, first+stride, ... of the seeded global ensemble (seed 0)."""
    total = first + stride * count
    rng = np.random.default_rng(0)
    ab = rng.uniform(-1, 1, (total, 2))  # row b = member b, whatever the ensemble size
    idx = first + stride * np.arange(count)
    u0 = np.stack([2.0 + 0.5 * ab[idx, 0], 0.5 * ab[idx, 1]], 1).reshape(count, 2, 1)
    params = np.full((count, 1), MU)
    return np.ascontiguousarray(u0), params


def algorithmic_bytes(B):
    # inputs (q*d + P + 3 doubles) + outputs (K*2d doubles + (K+2) counters) per member
    return B * ((2 * 1 + 1 + 3) * 8 + K_CHECKPOINTS * 2 * 1 * 8 + (K_CHECKPOINTS + 2) * 8)


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)  # fmt: skip
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nm, val in zip(names, r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(max(smax)) if smax else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


def cpu_oracle_rate(sample_members, threads):
    """Times the CPU oracle (C port of the reference algorithm, OpenMP over members) on the first
    `sample_members` members of the same seeded ensemble.  Returns (solves/s, accepted steps/s, seconds)."""
    from oracle import pn_oracle

    cfg = pn_oracle.make_config("van_der_pol", 1, NU, 2, factorisation="dense", correction="ts1", strategy="fixedpoint",
                                calibration="dynamic", atol=TOL, rtol=TOL, dt0=0.01, num_params=1)  # fmt: skip
    u0, params = ensemble_inputs(0, sample_members)
    save_at = np.linspace(T0, T1, K_CHECKPOINTS)
    pn_oracle.solve_save_at_batch(cfg, u0[: max(threads, 1)], params[: max(threads, 1)], save_at, num_threads=threads)
    t0 = time.perf_counter()
    out = pn_oracle.solve_save_at_batch(cfg, u0, params, save_at, num_threads=threads)
    dt = time.perf_counter() - t0
    assert (out["status"] == 0).all()
    return sample_members / dt, float(out["n_accepted"][:, -1].sum()) / dt, dt


def run_reference(args):
    """--impl reference: the reference algorithm's CPU implementation on the host cores.  The
    reference itself (JAX + probdiffeq, jit + vmap) cannot be installed in this image (no wheels,
    no network), so this is the validated C port under oracle/ with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = int(min(MEMBERS, max(64, 24 * cores)))  # ~1 s of CPU work per step
    rates, steps_rates = [], []
    for _ in range(args.warmup):
        cpu_oracle_rate(sample, cores)
    t_all = time.perf_counter()
    for _ in range(args.steps):
        r, s, _ = cpu_oracle_rate(sample, cores)
        rates.append(r)
        steps_rates.append(s)
    ms = (time.perf_counter() - t_all) / args.steps * 1e3
    value = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "accepted_steps_per_s": float(np.mean(steps_rates)),
        "config": {"workload": f"van_der_pol mu=1e3 ensemble ({MEMBERS} members; CPU sample of {sample}), dense EKF1 + "
                               f"fixed-point smoother, nu={NU}, atol=rtol={TOL:g}, {K_CHECKPOINTS} checkpoints on [0, 6.3]"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"first {sample} members of the seeded ensemble per step, OpenMP over members"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference = C port of the reference's algorithm (oracle/); JAX/probdiffeq are not installable here",
    }  # fmt: skip
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--members", type=int, default=MEMBERS, help="members per GPU (default: the headline 65,536)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from odecheckpts_b200 import _cabi, ensemble, ivps, ivpsolvers

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the solver has no CPU fallback")
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _cabi.lib()

    B = args.members               # members per rank (weak scaling)
    B_total = B * world
    K = K_CHECKPOINTS
    # interleaved sharding of the global seeded ensemble: rank r owns members r, r+G, ...
    u0_h, par_h = ensemble_inputs(rank, B, world)
    save_h = np.linspace(T0, T1, K)
    desc = _cabi.Desc(_cabi.PROBLEM_IDS["van_der_pol"], 1, NU, 2, _cabi.FACTORISATIONS["dense"], _cabi.CORRECTIONS["ts1"],
                      _cabi.STRATEGIES["fixedpoint"], _cabi.CALIBRATIONS["dynamic"], TOL, TOL, 0.01,
                      0.95, 0.2, 10.0, 0.3, 0.4, B, K, 0, 1, 0, 0)  # fmt: skip
    info = _cabi.kernel_info(desc)
    u0_d = torch.as_tensor(u0_h, device=dev)
    par_d = torch.as_tensor(par_h, device=dev)
    save_d = torch.as_tensor(save_h, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    state = {"out": None}

    def step():
        flush.zero_()  # evict L2 between timed iterations
        ws = None if state["out"] is None else state["out"]["_workspace"]
        out = _cabi.solve_device(desc, u0_d, par_d, None, save_d, None, workspace=ws, out=state["out"])
        state["out"] = out
        if world > 1:
            local = {k: out[k] for k in ("u", "u_std", "n_accepted", "n_rejected", "status")}
            state["gathered"] = ensemble.all_gather_results(local, B_total)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fp64_peak = _cabi.measure_fp64_peak()
    _cabi.set_profiling(True)
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    kernel_ms, smooth_ms = [], []
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
        a, b = _cabi.last_timing()  # waits for this step's kernels (events on the launching stream)
        kernel_ms.append(a)
        smooth_ms.append(b)
    e1.record()
    barrier()
    elapsed_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    _cabi.set_profiling(False)
    if world > 1:
        tt = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed_ms = float(tt.item())
    out = state["out"]
    n_acc = out["n_accepted"][:, -1].double().sum()
    n_rej = out["n_rejected"].double().sum()
    bad = (out["status"] != 0).sum().double()
    stats = torch.stack([n_acc, n_rej, bad])
    if world > 1:
        dist.all_reduce(stats)
    acc_total, rej_total, bad_total = (float(x) for x in stats.tolist())
    ms_per_step = elapsed_ms / args.steps
    value = B_total / (ms_per_step * 1e-3)

    # ---- end-to-end through the public API with host buffers ------------------------------
    vf, (y0, dy0), _ = ivps.van_der_pol(mu=MU)
    solve = ivpsolvers.solve(f"ts0-{NU}", vf, y0, save_at=save_h, dt0=0.01, atol=TOL, rtol=TOL, ode_order=2,
                             factorisation="dense", correction="ts1", return_marginals=False, device=local_rank)  # fmt: skip
    pin = [torch.empty((B, 1), dtype=torch.float64, pin_memory=True) for _ in range(2)]  # pinned host inputs
    pin[0].copy_(torch.from_numpy(u0_h[:, 0, :]))
    pin[1].copy_(torch.from_numpy(u0_h[:, 1, :]))
    u0_pin = (pin[0].numpy(), pin[1].numpy())
    e2e_steps = max(2, min(args.steps, 3))
    res, aux = solve(u0_pin, ())  # warm-up: two calls, so that both generations of recycled host result
    res, aux = solve(u0_pin, ())  # buffers exist (the previous results are still alive during a call)
    barrier()
    t_e2e = time.perf_counter()
    for _ in range(e2e_steps):
        res, aux = solve(u0_pin, ())
        sol = aux["solution"]
    barrier()
    e2e_s = (time.perf_counter() - t_e2e) / e2e_steps
    if world > 1:
        tt = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    h2d = u0_h.nbytes + par_h.nbytes + save_h.nbytes
    d2h = res.nbytes + sol.u_std.nbytes + sol.num_steps.nbytes + sol.num_rejected.nbytes + sol.status.nbytes
    e2e_ok = bool(np.array_equal(res, out["u"].cpu().numpy()))

    if rank == 0:
        attempts_rank = (acc_total + rej_total) / world
        flops = attempts_rank * W_ATTEMPT + B * (K - 1) * W_CHECKPOINT
        k_ms = float(np.mean(kernel_ms))
        achieved = flops / (k_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_ach = algorithmic_bytes(B) / (k_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "accepted_steps_per_s": acc_total / (ms_per_step * 1e-3),
            "attempted_steps_per_s": (acc_total + rej_total) / (ms_per_step * 1e-3),
            "accepted_per_member": acc_total / B_total, "rejected_per_member": rej_total / B_total,
            "failed_members": int(bad_total),
            "config": {
                "workload": f"van_der_pol mu=1e3 ensemble, {B} members per GPU x {world} GPU(s) (seed 0: u0=2+0.5U, u'0=0.5U), "
                            f"dense EKF1 (ode_order=2) + fixed-point smoother, nu={NU}, dynamic calibration, "
                            f"atol=rtol={TOL:g}, dt0=0.01, {K} checkpoints on [0, 6.3]",
                "members_total": B_total, "parallelism": f"ensemble sharded x{world}, one all-gather of results" if world > 1 else "single GPU",
                "l2": "256 MiB device memset between timed steps (L2 flush)",
                "kernel": info,
                "scheduling": ("run to completion (PN_B200_NO_SLICE)" if os.environ.get("PN_B200_NO_SLICE")
                               else "time-sliced members: parked at quantum boundaries, most lagging ready member first"),
            },
            "gpu_launches": 2 * args.steps,
            "e2e": {"value": B_total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "api": "odecheckpts_b200.ivpsolvers.solve(...)(u0_host, p) -> pn_b200_solve_save_at_host",
                    "matches_device_path": e2e_ok},
            "roofline": {
                "bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                "traffic": NCU_DRAM_BYTES_PER_LAUNCH if B == MEMBERS else None,
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full, profiles/r01_scalar_kernel_final_ncu.txt",
                "kernel": "pn_scalar_kernel<VanDerPol,4,fixedpoint>", "kernel_ms": k_ms,
                "smooth_kernel_ms": float(np.mean(smooth_ms)),
                "flops_per_attempt": W_ATTEMPT, "attempts_per_launch": attempts_rank,
                "peak_source": "DFMA-chain microbenchmark in this run (pn_b200_measure_fp64_peak); MEASURED_PEAKS.json has no fp64 entry",
                "hbm": {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650"},
            },
            "clocks": clocks,
        }  # fmt: skip
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            sample = int(min(B, max(512, 640 * cores)))  # ~10-20 s of CPU work
            rate, srate, secs = cpu_oracle_rate(sample, cores)
            line["cpu_baseline"] = {
                "value": rate, "unit": UNIT, "cores": cores, "kind": "port", "accepted_steps_per_s": srate, "seconds": secs,
                "sample": f"first {sample} members of the same seeded ensemble, C oracle with OpenMP over members "
                          "(JAX/probdiffeq not installable here)",
            }  # fmt: skip
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
