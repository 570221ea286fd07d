// Dense-factorisation instances with D = (nu+1) d <= 16: two IVPs per warp, Householder columns
// in registers (pn_dense_rows_kernel.cuh).  BASELINE config 3's dense EKF1 on the rigid body.
#include "pn_registry.h"
namespace pn {
using Brusselator2r = Brusselator<2>;
}  // namespace pn
PN_REGISTER_DENSE_ROWS(RigidBody, 2, 1, 16, 1);
PN_REGISTER_DENSE_ROWS(RigidBody, 4, 1, 16, 2);  // two warps per CTA: four CTAs = eight warps per SM at D = 15
PN_REGISTER_DENSE_ROWS(RigidBody, 4, 0, 16, 2);
PN_REGISTER_DENSE_ROWS(LotkaVolterra, 4, 1, 16, 1);
PN_REGISTER_DENSE_ROWS(ThreeBody, 4, 1, 16, 1);
PN_REGISTER_DENSE_ROWS(Brusselator2r, 4, 1, 32, 1);  // D = 20: one IVP per warp
