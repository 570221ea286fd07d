// Brusselator with a runtime grid size N (d = 2N up to 4096): CTA-per-IVP wide kernel, isotropic EKF0
// nu = 4 (experiments/4_brusselator/run.py:42-61 sweeps N = 2 ... 512; BASELINE config 5 goes to 1024).
#include "pn_registry.h"
PN_REGISTER_WIDE(BrusselatorWide, 4, 1);
PN_REGISTER_WIDE(BrusselatorWide, 4, 0);
// One warp per IVP: every thread of a CTA replicates the n x n factor arithmetic (it is the serial part of the
// step), so a 128-thread CTA issues it four times per attempted step.  For ensembles with several members per
// SM the 32-thread build issues it once and keeps up to eight members resident per SM instead of two.
PN_REGISTER_WIDE_T(BrusselatorWide, 4, 1, 32);
// At most one member per SM, fixed-point strategy: 128 main threads + a backward warp that owns the running
// conditional (pn_scalar_kernel.cuh, PIPE = 1): two fifths of the per-step dependency chain move off the critical path.
PN_REGISTER_WIDE_PIPE(BrusselatorWide, 4, 1);
