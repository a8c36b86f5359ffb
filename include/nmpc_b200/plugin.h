/* nmpc_b200 -- bringing YOUR problem to the engine: a user functor in its own shared library.
 *
 * The reference binds any std::shared_ptr<DDPProblem<S, I>> to a solver at run time through virtual calls
 * (isri-aist/NMPC nmpc_ddp/include/nmpc_ddp/DDPSolver.h:255, FmpcSolver.h:296).  A GPU kernel cannot call host
 * virtuals, so the engine's stage kernels are TEMPLATES over a trivially-copyable problem functor and a problem
 * reaches them by being compiled with them.  That does not require touching libnmpc_b200.so:
 *
 *   1. write the functor (same method names / argument order as DDPProblem, see include/nmpc_b200/models/cartpole.h:
 *      static constexpr int NX, NU[, NG], NUM_PARAMS; using Scalar; fromParams / defaultParams; dt(); stateEq;
 *      runningCost; terminalCost; calcStateEqDeriv; calcRunningCostDeriv; calcTerminalCostDeriv
 *      [; ineqConst; calcIneqConstDeriv] [; inputDim(t)] [; ineqDim(t)]), all NMPC_HD;
 *   2. in ONE .cu file:
 *          #include <nmpc_b200/plugin.h>
 *          #include "my_problem.h"
 *          NMPC_B200_REGISTER_DDP_MODEL("my_problem", MyProblem<double>);      // and / or ..._FMPC_MODEL
 *   3. nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 --expt-relaxed-constexpr -Xcompiler -fPIC -shared \
 *          -I<repo>/include my_problem.cu -o libmy_problem.so -L<repo>/nmpc_b200 -lnmpc_b200
 *   4. at run time: nmpc_b200_load_plugin("libmy_problem.so") (c_api.h), then nmpc_b200_ddp_create("my_problem", ...)
 *      or, through the C++ facade, nmpc_ddp::FunctorProblem<MyProblem<double>>("my_problem") + nmpc_ddp::DDPSolver.
 *
 * The registrar runs when the library is loaded and adds the functor's kernels (instantiated inside the plugin) to the
 * registry of libnmpc_b200.so.  tests/plugin/ holds a complete example (a torque-limited pendulum that is NOT part of
 * the library) and tests/test_plugin.py builds it, loads it and checks it against the oracle.
 */
#pragma once

#include <nmpc_b200/engine/register.cuh>
