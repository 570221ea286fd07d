// Brusselator instances with a small fixed grid (src/odecheckpts/ivps.py:124-156;
// experiments/4_brusselator/run.py:42-61: N = 2, 4, 8, 16 of the reference's sweep), isotropic EKF0
// nu = 4 with the fixed-point smoother, lane per dimension (GROUP = 2N lanes per IVP).
#include "pn_registry.h"
namespace pn {
using Brusselator2 = Brusselator<2>;
using Brusselator4 = Brusselator<4>;
using Brusselator8 = Brusselator<8>;
using Brusselator16 = Brusselator<16>;
}  // namespace pn
PN_REGISTER_GROUP(Brusselator2, 4, 1, 4, 0);
PN_REGISTER_GROUP(Brusselator4, 4, 1, 8, 0);
PN_REGISTER_GROUP(Brusselator8, 4, 1, 16, 0);
PN_REGISTER_GROUP(Brusselator16, 4, 1, 32, 0);
PN_REGISTER_GROUP(Brusselator4, 4, 1, 8, 1);
PN_REGISTER_GROUP(Brusselator4, 4, 0, 8, 0);
