"""Extract the reference's committed golden artefacts into small fixtures.

Run HERE (the container that mounts /root/reference, read-only):

    python tests/golden/make_golden.py

It reads `experiments/**/*.npy` of pnkraemer/code-adaptive-prob-ode-solvers and
writes `tests/golden/reference_goldens.npz` (plain float64/int64 arrays, no
pickles).  Nothing at test/bench time reads /root/reference; only this file does.

Several reference artefacts are pickled dicts that hold jax arrays
(`jnp.save(..., allow_pickle=True)`: experiments/4_brusselator/run.py:142-151,
experiments/2_workprec_simple/run_simple.py:133-136).  jax is not installed in
this image, so they are decoded with an Unpickler that maps jax's
`_reconstruct_array(fun, args, arr_state, aval_state)` onto the plain numpy
reconstruction it wraps.
"""

import io
import os
import pickle
import sys

import numpy as np

REF = os.environ.get("PN_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_goldens.npz")


def _reconstruct_array(fun, args, arr_state, aval_state=None):
    a = fun(*args)
    a.__setstate__(arr_state)
    return np.asarray(a)


# /root/reference is untrusted public content: the pickles are decoded with an ALLOW-LIST of exactly
# the globals a pickled numpy / jax array needs.  Anything else (os.system, builtins.eval, ...) raises.
_ALLOWED_GLOBALS = {
    ("numpy.core.multiarray", "_reconstruct"),
    ("numpy._core.multiarray", "_reconstruct"),
    ("numpy.core.multiarray", "scalar"),
    ("numpy._core.multiarray", "scalar"),
    ("numpy", "ndarray"),
    ("numpy", "dtype"),
}


class _NoJaxUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.split(".")[0] in ("jax", "jaxlib"):
            if name == "_reconstruct_array":
                return _reconstruct_array
            raise pickle.UnpicklingError(f"unexpected jax global {module}.{name}")
        if (module, name) in _ALLOWED_GLOBALS:
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"global {module}.{name} is not on the allow-list")


def load_pickled_npy(path):
    with open(path, "rb") as fh:
        version = np.lib.format.read_magic(fh)
        if version == (1, 0):
            np.lib.format.read_array_header_1_0(fh)
        else:
            np.lib.format.read_array_header_2_0(fh)
        payload = fh.read()
    obj = _NoJaxUnpickler(io.BytesIO(payload)).load()
    return obj.item() if isinstance(obj, np.ndarray) and obj.dtype == object else obj


def main():
    ex = os.path.join(REF, "experiments")
    out = {}

    # --- 1_van_der_pol: dense EKF1 (ode_order=2), nu=4, filter, dynamic, tol=1e-3,
    #     dt0=0.01, save-every-step (vdp.py:61-80,117-120)
    out["vdp_grid"] = np.load(os.path.join(ex, "1_van_der_pol", "vdp_baseline_grid.npy")).astype(np.float64)
    out["vdp_solution"] = np.load(os.path.join(ex, "1_van_der_pol", "vdp_baseline_solution.npy")).astype(np.float64)
    out["vdp_grid_replay"] = np.load(os.path.join(ex, "1_van_der_pol", "vdp_grid_adaptive.npy")).astype(np.float64)

    # --- 4_brusselator: isotropic EKF0 nu=4 tol=1e-8 dynamic fixed-point,
    #     200 checkpoints (run.py:51-61,119-138)
    chk = load_pickled_npy(os.path.join(ex, "4_brusselator", "data_checkpoint.npy"))
    txt = load_pickled_npy(os.path.join(ex, "4_brusselator", "data_textbook.npy"))
    out["brusselator_N"] = np.asarray(chk["N"], dtype=np.int64)
    out["brusselator_num_steps_checkpoint"] = np.asarray([int(x) for x in chk["num_steps"]], dtype=np.int64)
    out["brusselator_num_steps_terminal"] = np.asarray([int(x) for x in txt["num_steps"]], dtype=np.int64)
    for N, ts, ys in zip(chk["N"], chk["ts"], chk["ys"]):
        if int(N) <= 32:  # keep the fixture small: (200, 2N) doubles each
            out[f"brusselator_ts_N{int(N)}"] = np.asarray(ts, dtype=np.float64)
            out[f"brusselator_ys_N{int(N)}"] = np.asarray(ys, dtype=np.float64)

    # --- 2_workprec_simple: rigid body, isotropic EKF0 (run_simple.py:38-80,181-215)
    res = load_pickled_npy(os.path.join(ex, "2_workprec_simple", "data_results.npy"))
    for label, short in [
        ("TS0(2) (jit step) via probdiffeq", "rigid_interp_nu2"),
        ("TS0(4) (jit step) via probdiffeq", "rigid_interp_nu4"),
        ("TS0(2) (jit loop) via probdiffeq", "rigid_loop_nu2"),
        ("TS0(4) (jit loop) via probdiffeq", "rigid_loop_nu4"),
    ]:
        for key, val in res[label].items():
            arr = np.asarray(val, dtype=np.float64)
            out[f"{short}_{key}"] = arr
    out["rigid_truth_ts"] = np.load(os.path.join(ex, "2_workprec_simple", "data_ts.npy")).astype(np.float64)
    out["rigid_truth_ys"] = np.load(os.path.join(ex, "2_workprec_simple", "data_ys.npy")).astype(np.float64)
    out["rigid_checkpoints"] = np.load(os.path.join(ex, "2_workprec_simple", "data_checkpoints.npy")).astype(np.float64)

    # --- 3_workprec_harder: Pleiades (run_harder.py:42-60)
    res = load_pickled_npy(os.path.join(ex, "3_workprec_harder", "data_results.npy"))
    for label, short in [
        ("Prob(3) via probdiffeq", "pleiades_nu3"),
        ("Prob(5) via probdiffeq", "pleiades_nu5"),
        ("Prob(8) via probdiffeq", "pleiades_nu8"),
    ]:
        for key, val in res[label].items():
            out[f"{short}_{key}"] = np.asarray(val, dtype=np.float64)
    ts = np.load(os.path.join(ex, "3_workprec_harder", "data_ts.npy")).astype(np.float64)
    ys = np.load(os.path.join(ex, "3_workprec_harder", "data_ys.npy")).astype(np.float64)
    out["pleiades_truth_ts"] = ts
    out["pleiades_truth_ys"] = ys
    out["pleiades_checkpoints"] = np.load(os.path.join(ex, "3_workprec_harder", "data_checkpoints.npy")).astype(np.float64)

    # --- 5_vs_interpolation: three-body, isotropic EKF0 o2 nu=4 uncalibrated
    #     (measure.py:44-68,191-192): step counts at tol 1e-4/1e-7/1e-10
    res = load_pickled_npy(os.path.join(ex, "5_vs_interpolation", "data_results.npy"))
    steps = {}
    for row in res.values():
        steps[row["Tolerance"]] = int(row["No. steps"].replace(",", ""))
    out["threebody_tols"] = np.asarray([1e-4, 1e-7, 1e-10])
    out["threebody_num_steps"] = np.asarray(
        [steps["$10^{-4}$"], steps["$10^{-7}$"], steps["$10^{-10}$"]], dtype=np.int64
    )
    sol = np.load(os.path.join(ex, "5_vs_interpolation", "data_solution.npy"), allow_pickle=False)
    out["threebody_filter_solution"] = np.asarray(sol, dtype=np.float64)

    np.savez_compressed(OUT, **out)
    for k, v in out.items():
        print(f"{k:45s} {v.dtype} {v.shape}")
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    sys.exit(main())
