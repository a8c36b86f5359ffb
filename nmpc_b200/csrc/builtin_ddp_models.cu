/* nmpc_b200 -- DDP kernels instantiated for the problem functors shipped with the library. */
#include <nmpc_b200/models/bipedal.h>
#include <nmpc_b200/models/cartpole.h>

#include <nmpc_b200/engine/register.cuh>

NMPC_B200_REGISTER_DDP_MODEL("cartpole", nmpc_b200::models::CartPole<double>);
NMPC_B200_REGISTER_DDP_MODEL("bipedal", nmpc_b200::models::Bipedal<double>);
// the cart-pole written for instruction latency (models/cartpole.h BRANCH_FREE), as the latency-bound kernels evaluate
// it: registered for nmpc_b200_model_eval only, so that tests can compare its values with "cartpole"
NMPC_B200_REGISTER_EVAL_ONLY("cartpole_branch_free", nmpc_b200::models::CartPole<double, true>);
