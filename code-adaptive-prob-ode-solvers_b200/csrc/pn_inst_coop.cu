// Cooperative (n lanes per IVP) instances for small ensembles of scalar ODEs: Van der Pol (BASELINE
// config 1: a single IVP; config 2 strong-scaled over 8 GPUs) and the logistic ODE of the reference's test.
#include "pn_registry.h"
PN_REGISTER_COOP(VanDerPol, 4, 0);
PN_REGISTER_COOP(VanDerPol, 4, 1);
PN_REGISTER_COOP(VanDerPol, 2, 1);
PN_REGISTER_COOP(Logistic, 2, 1);
PN_REGISTER_COOP(Logistic, 4, 1);
