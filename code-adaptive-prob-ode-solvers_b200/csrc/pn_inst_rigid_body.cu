// Rigid-body instances (src/odecheckpts/ivps.py:20-29; experiments/2_workprec_simple/run_simple.py:38-56;
// BASELINE config 3), isotropic EKF0.
#include "pn_registry.h"
PN_REGISTER_SCALAR(RigidBody, 2, 0);
PN_REGISTER_SCALAR(RigidBody, 2, 1);
PN_REGISTER_SCALAR(RigidBody, 4, 0);
PN_REGISTER_SCALAR(RigidBody, 4, 1);
