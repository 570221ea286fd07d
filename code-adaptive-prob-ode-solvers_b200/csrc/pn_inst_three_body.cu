// Restricted three-body instances (src/odecheckpts/ivps.py:32-41;
// experiments/5_vs_interpolation/measure.py:44-68,163), isotropic EKF0, ode_order = 2.
#include "pn_registry.h"
PN_REGISTER_SCALAR(ThreeBody, 3, 0);
PN_REGISTER_SCALAR(ThreeBody, 3, 1);
PN_REGISTER_SCALAR(ThreeBody, 4, 0);
PN_REGISTER_SCALAR(ThreeBody, 4, 1);
