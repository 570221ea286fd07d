// nu = 8 (n = 9): Pleiades isotropic (Prob(8), experiments/3_workprec_harder/run_harder.py:74-77) and the
// logistic ODE; 64-thread CTAs because the per-thread shared-memory state is 245 doubles.
#include "pn_registry.h"
PN_REGISTER_GROUP_T(Pleiades, 8, 1, 16, 0, 64);
PN_REGISTER_GROUP_T(Pleiades, 8, 1, 16, 1, 64);  // blockdiag (BASELINE config 4, n = 9)
PN_REGISTER_SCALAR_T(Logistic, 8, 1, 64);
