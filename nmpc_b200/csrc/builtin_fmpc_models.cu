/* nmpc_b200 -- FMPC kernels instantiated for the problem functors shipped with the library. */
#include <nmpc_b200/models/cartpole.h>
#include <nmpc_b200/models/oscillator.h>
#include <nmpc_b200/models/planar_quadrotor.h>

#include <nmpc_b200/engine/register.cuh>

NMPC_B200_REGISTER_FMPC_MODEL("cartpole", nmpc_b200::models::CartPole<double>);
NMPC_B200_REGISTER_FMPC_MODEL("oscillator", nmpc_b200::models::Oscillator<double>);
// two inputs: the pivoted-LDLT / FullPivLU gain solve of FmpcSolver.hpp:596-617
NMPC_B200_REGISTER_FMPC_MODEL("planar_quadrotor", nmpc_b200::models::PlanarQuadrotor<double>);
// time-varying inequality dimension (FmpcProblem<4, 1, Eigen::Dynamic>)
NMPC_B200_REGISTER_FMPC_MODEL("cartpole_windowed", nmpc_b200::models::CartPoleWindowed<double>);
