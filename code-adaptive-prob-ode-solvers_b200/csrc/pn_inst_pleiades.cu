// Pleiades instances (src/odecheckpts/ivps.py:59-99; experiments/3_workprec_harder/run_harder.py:42-60;
// BASELINE config 4): d = 14, ode_order = 2, lane-per-dimension kernels (16 lanes per IVP),
// isotropic (what the reference runs) and blockdiag (what BASELINE config 4 asks for), EKF0.
#include "pn_registry.h"
PN_REGISTER_GROUP(Pleiades, 3, 1, 16, 0);
PN_REGISTER_GROUP(Pleiades, 3, 1, 16, 1);
PN_REGISTER_GROUP(Pleiades, 4, 1, 16, 0);
PN_REGISTER_GROUP(Pleiades, 4, 1, 16, 1);
PN_REGISTER_GROUP(Pleiades, 5, 1, 16, 0);
PN_REGISTER_GROUP(Pleiades, 5, 1, 16, 1);
PN_REGISTER_GROUP(Pleiades, 3, 0, 16, 0);
PN_REGISTER_GROUP(Pleiades, 5, 0, 16, 0);
