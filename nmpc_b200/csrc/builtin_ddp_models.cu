/* nmpc_b200 -- DDP kernels instantiated for the problem functors shipped with the library. */
#include <nmpc_b200/models/bipedal.h>
#include <nmpc_b200/models/cartpole.h>

#include <nmpc_b200/engine/register.cuh>

NMPC_B200_REGISTER_DDP_MODEL("cartpole", nmpc_b200::models::CartPole<double>);
NMPC_B200_REGISTER_DDP_MODEL("bipedal", nmpc_b200::models::Bipedal<double>);
