// Logistic ODE instances (src/odecheckpts/ivps.py:8-17; tests/test_ivpsolvers.py:27-28).
#include "pn_registry.h"
PN_REGISTER_SCALAR(Logistic, 2, 0);
PN_REGISTER_SCALAR(Logistic, 2, 1);
PN_REGISTER_SCALAR(Logistic, 3, 0);
PN_REGISTER_SCALAR(Logistic, 3, 1);
PN_REGISTER_SCALAR(Logistic, 4, 0);
PN_REGISTER_SCALAR(Logistic, 4, 1);
