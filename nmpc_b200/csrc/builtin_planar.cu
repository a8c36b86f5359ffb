/* nmpc_b200 -- DDP kernels for the planar quadrotor functor: n_x = 6, two inputs, so the control-limited backward pass
   runs the two-variable box QP (BoxQP<2>, DDPSolver.hpp:450-497).  tests/test_ddp_planar.py pins it to the reference's
   DDPSolver<6, 2>. */
#include <nmpc_b200/models/planar_quadrotor.h>

#include <nmpc_b200/engine/register.cuh>

NMPC_B200_REGISTER_DDP_MODEL("planar_quadrotor", nmpc_b200::models::PlanarQuadrotor<double>);
