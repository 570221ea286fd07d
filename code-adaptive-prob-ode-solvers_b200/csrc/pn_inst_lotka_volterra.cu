// Lotka-Volterra instances (diffeqzoo default problem), isotropic EKF0.
#include "pn_registry.h"
PN_REGISTER_SCALAR(LotkaVolterra, 4, 0);
PN_REGISTER_SCALAR(LotkaVolterra, 4, 1);
