// blockdiag factorisation for the small problems (lane per dimension), EKF0.
#include "pn_registry.h"
PN_REGISTER_GROUP(RigidBody, 2, 1, 4, 1);
PN_REGISTER_GROUP(RigidBody, 4, 1, 4, 1);
PN_REGISTER_GROUP(ThreeBody, 4, 1, 2, 1);
PN_REGISTER_GROUP(LotkaVolterra, 4, 1, 2, 1);
