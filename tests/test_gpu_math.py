"""The kernels' branch-free reciprocal / square root must be IEEE-exact on normal operands (GPU)."""

import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_branch_free_rcp_and_sqrt_are_correctly_rounded():
    import torch

    from odecheckpts_b200 import _cabi

    lib = _cabi.lib()
    lib.pn_b200_selftest_math.restype = C.c_int
    lib.pn_b200_selftest_math.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    rng = np.random.default_rng(123)
    n = 1 << 22
    mant = rng.uniform(1.0, 2.0, n)
    expo = rng.integers(-600, 600, n)
    x = np.ldexp(mant, expo) * rng.choice([-1.0, 1.0], n)
    x[:8] = [1.0, 2.0, 3.0, 0.1, 1e-300, 1e300, 0.0, np.inf]
    y = rng.uniform(0.01, 0.2, n)
    xd = torch.as_tensor(x, device="cuda")
    yd = torch.as_tensor(y, device="cuda")
    r = torch.empty_like(xd)
    sq = torch.empty_like(xd)
    pw = torch.empty_like(xd)
    rc = lib.pn_b200_selftest_math(xd.data_ptr(), yd.data_ptr(), r.data_ptr(), sq.data_ptr(), pw.data_ptr(), n)
    assert rc == 0
    torch.cuda.synchronize()
    with np.errstate(divide="ignore"):
        np.testing.assert_array_equal(r.cpu().numpy(), 1.0 / x)
    np.testing.assert_array_equal(sq.cpu().numpy(), np.sqrt(np.abs(x)))
    # deterministic pow: a few ulp from libm, and identical to the oracle's C implementation
    from oracle import pn_oracle

    ax = np.abs(x[:4096])
    ax[6:8] = [0.5, 7.0]
    got = pw.cpu().numpy()[:4096]
    ol = pn_oracle.lib()
    want = np.array([ol.pn_det_pow(float(a), float(b)) for a, b in zip(ax, y[:4096])])
    sel = np.arange(4096) != 6
    sel[7] = False
    np.testing.assert_array_equal(got[sel], want[sel])
    np.testing.assert_allclose(got[sel], ax[sel] ** y[:4096][sel], rtol=1e-12)
