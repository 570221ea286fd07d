// Thread-per-IVP kernels with a backward lane per IVP (pn_scalar_kernel.cuh, PAIR = 1) for small ensembles of the
// problems the reference solves one at a time or in small batches: Van der Pol (BASELINE config 1; config 2
// strong-scaled over 8 GPUs), the logistic ODE of the reference's test, three-body (the mailbox of a d = 3 problem
// does not fit beside 128 members' state in one SM's shared memory).
#include "pn_registry.h"
PN_REGISTER_PAIR(VanDerPol, 4);
PN_REGISTER_PAIR(VanDerPol, 2);
PN_REGISTER_PAIR(Logistic, 4);
PN_REGISTER_PAIR(Logistic, 2);
PN_REGISTER_PAIR(ThreeBody, 4);
