/* nmpc_b200 -- FMPC kernels instantiated for the problem functors shipped with the library. */
#include <nmpc_b200/models/cartpole.h>
#include <nmpc_b200/models/oscillator.h>

#include "register.cuh"

NMPC_B200_REGISTER_FMPC_MODEL("cartpole", nmpc_b200::models::CartPole<double>);
NMPC_B200_REGISTER_FMPC_MODEL("oscillator", nmpc_b200::models::Oscillator<double>);
