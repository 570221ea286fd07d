// More prior orders for the small problems (the reference's method strings are "ts0-<nu>",
// src/odecheckpts/ivpsolvers.py:35) and nu = 8 for the Pleiades (Prob(8) of
// experiments/3_workprec_harder/run_harder.py:74-77).  nu = 8 keeps n = 9 factors per lane: the
// state no longer fits the register file (local-memory spills) and a CTA holds 64 threads.
#include "pn_registry.h"
PN_REGISTER_SCALAR(RigidBody, 3, 1);
PN_REGISTER_SCALAR(RigidBody, 5, 1);
PN_REGISTER_SCALAR(ThreeBody, 2, 1);
PN_REGISTER_SCALAR(ThreeBody, 5, 1);
PN_REGISTER_SCALAR(LotkaVolterra, 2, 1);
PN_REGISTER_SCALAR(LotkaVolterra, 3, 1);
PN_REGISTER_SCALAR(Logistic, 5, 1);
PN_REGISTER_SCALAR(Logistic, 5, 0);
