// Van der Pol instances (src/odecheckpts/ivps.py:159-167; experiments/1_van_der_pol/vdp.py;
// BASELINE configs 1 and 2).  D = 1: serves dense (EKF0/EKF1) and isotropic (EKF0).
#include "pn_registry.h"
PN_REGISTER_SCALAR(VanDerPol, 2, 0);
PN_REGISTER_SCALAR(VanDerPol, 2, 1);
PN_REGISTER_SCALAR(VanDerPol, 3, 0);
PN_REGISTER_SCALAR(VanDerPol, 3, 1);
PN_REGISTER_SCALAR(VanDerPol, 4, 0);
PN_REGISTER_SCALAR(VanDerPol, 4, 1);
PN_REGISTER_SCALAR(VanDerPol, 5, 0);
PN_REGISTER_SCALAR(VanDerPol, 5, 1);
