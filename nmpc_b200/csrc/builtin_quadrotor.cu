/* nmpc_b200 -- DDP kernels for the quadrotor functor (n_x = 12, n_u = 4): fp32 (BASELINE.json configs[3]) and fp64. */
#include <nmpc_b200/models/quadrotor.h>

#include <nmpc_b200/engine/register.cuh>

NMPC_B200_REGISTER_DDP_MODEL("quadrotor", nmpc_b200::models::Quadrotor<float>);
NMPC_B200_REGISTER_DDP_MODEL("quadrotor_f64", nmpc_b200::models::Quadrotor<double>);
