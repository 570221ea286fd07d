"""``from probdiffeq.impl import impl; impl.select("isotropic", ode_shape=(d,))``

The reference relies on this piece of process-global state (src/odecheckpts/ivpsolvers.py:29-33,
experiments/1_van_der_pol/vdp.py:61).  It is mirrored for drop-in compatibility; every solve
routine also accepts an explicit ``factorisation=`` override, which is what the C ABI uses.
"""

from .._cabi import FACTORISATIONS


class _Impl:
    def __init__(self):
        self.name = None
        self.ode_shape = None

    def select(self, name, *, ode_shape):
        if name not in FACTORISATIONS:
            raise ValueError(f"unknown state-space factorisation {name!r}; choose from {sorted(FACTORISATIONS)}")
        self.name = name
        self.ode_shape = tuple(ode_shape)

    def selected(self):
        if self.name is None:
            raise RuntimeError("no state-space factorisation selected; call impl.select(...) first")
        return self.name


impl = _Impl()
select = impl.select
