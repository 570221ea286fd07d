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


def test_reflector_scalars_equal_the_plain_square_root_then_reciprocal():
    """make_reflector takes its signs through integer operations and its negated 2 / v^T v from the reciprocal's last
    Newton step; the results must be the bits of the plain composition sqrt -> reciprocal -> negations, which is what
    the oracle's reflector computes with the IEEE operations (oracle/pn_linalg.c)."""
    import torch

    from odecheckpts_b200 import _cabi

    lib = _cabi.lib()
    lib.pn_b200_selftest_reflector.restype = C.c_int
    lib.pn_b200_selftest_reflector.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    rng = np.random.default_rng(321)
    n = 1 << 22
    alpha = np.ldexp(rng.uniform(1.0, 2.0, n), rng.integers(-120, 120, n)) * rng.choice([-1.0, 1.0], n)
    # sub-diagonal sums of squares from far below to far above alpha^2, and exactly zero (degenerate column)
    sigma2 = alpha * alpha * np.ldexp(rng.uniform(1.0, 2.0, n), rng.integers(-110, 60, n))
    sigma2[: n // 16] = 0.0
    alpha[: n // 64] = 0.0
    alpha[n // 64 : n // 32] = -0.0
    sigma2[n // 16 : n // 8] = np.ldexp(rng.uniform(1.0, 2.0, n // 16), rng.integers(-200, 200, n // 16))
    alpha[n // 16 : n // 12] = 0.0
    ad = torch.as_tensor(alpha, device="cuda")
    sd = torch.as_tensor(sigma2, device="cuda")
    out = torch.empty(n * 8, dtype=torch.float64, device="cuda")
    assert lib.pn_b200_selftest_reflector(ad.data_ptr(), sd.data_ptr(), out.data_ptr(), n) == 0
    torch.cuda.synchronize()
    o = out.cpu().numpy().reshape(n, 8)
    assert np.isfinite(o).all()
    # bit patterns, so that the sign of a zero counts as well
    np.testing.assert_array_equal(o[:, :4].view(np.int64), o[:, 4:].view(np.int64))
    # and against numpy's IEEE operations where no fused multiply-add is involved (alpha = 0: x = sigma2)
    z = (alpha == 0.0) & (sigma2 > 0)
    norm = np.sqrt(sigma2[z])
    np.testing.assert_array_equal(o[z, 0], norm)
    np.testing.assert_array_equal(o[z, 1], -norm)
    np.testing.assert_array_equal(o[z, 2], 1.0 / (norm * norm))
