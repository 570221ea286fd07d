// Brusselator with a runtime grid size N (d = 2N up to 4096): CTA-per-IVP wide kernel, isotropic EKF0
// nu = 4 (experiments/4_brusselator/run.py:42-61 sweeps N = 2 ... 512; BASELINE config 5 goes to 1024).
#include "pn_registry.h"
PN_REGISTER_WIDE(BrusselatorWide, 4, 1);
PN_REGISTER_WIDE(BrusselatorWide, 4, 0);
