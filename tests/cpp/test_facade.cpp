// Host C++ test of the template facades (include/nmpc_ddp, include/nmpc_fmpc) written the way the
// reference's own tests use the solvers (nmpc_ddp/tests/src/TestDDPCartPole.cpp:268-309, :609-649;
// nmpc_fmpc/tests/src/TestFmpcCartPole.cpp:315-330).  Compiled with g++ (no nvcc), links libnmpc_b200.so.
// Prints one "key value..." line per result; tests/test_cpp_facade.py compares them with the oracle.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <memory>

#include <nmpc_b200/models/cartpole.h>
#include <nmpc_b200/models/centroidal_motion.h>
#include <nmpc_b200/models/vertical_motion.h>
#include <nmpc_ddp/DDPSolver.h>
#include <nmpc_fmpc/FmpcSolver.h>

using CartPoleF = nmpc_b200::models::CartPole<double>;
using DDPProblemCartPole = nmpc_ddp::FunctorProblem<CartPoleF>;
using FmpcProblemCartPole = nmpc_fmpc::FunctorProblem<CartPoleF>;
using DDPProblemCentroidalMotion = nmpc_ddp::FunctorProblem<nmpc_b200::models::CentroidalMotion<double>>;
using DDPProblemVerticalMotion = nmpc_ddp::FunctorProblem<nmpc_b200::models::VerticalMotion<double>>;

// TestDDPCentroidalMotion.CheckDerivative (TestDDPCentroidalMotion.cpp:355-411) through the host-side virtuals, and
// DDPProblem::inputDim(t) of the two problems with a time-varying input dimension
static int checkCentroidal()
{
  auto problem = std::make_shared<DDPProblemCentroidalMotion>("centroidal_motion");
  double t = 0;
  DDPProblemCentroidalMotion::StateDimVector x;
  DDPProblemCentroidalMotion::InputDimVector u;
  for(int i = 0; i < 9; i++) x[i] = std::sin(1.0 + i); // any point: the test draws a random one
  for(int i = 0; i < 16; i++) u[i] = std::cos(2.0 + i);
  DDPProblemCentroidalMotion::StateStateDimMatrix fx_a, fx_n;
  DDPProblemCentroidalMotion::StateInputDimMatrix fu_a, fu_n;
  problem->calcStateEqDeriv(t, x, u, fx_a, fu_a);
  constexpr double deriv_eps = 1e-6;
  for(int i = 0; i < problem->stateDim(); i++)
  {
    auto xp = x, xm = x;
    xp[i] += deriv_eps, xm[i] -= deriv_eps;
    auto d = problem->stateEq(t, xp, u) - problem->stateEq(t, xm, u);
    for(int r = 0; r < 9; r++) fx_n(r, i) = d[r] / (2 * deriv_eps);
  }
  for(int i = 0; i < problem->inputDim(t); i++)
  {
    auto up = u, um = u;
    up[i] += deriv_eps, um[i] -= deriv_eps;
    auto d = problem->stateEq(t, x, up) - problem->stateEq(t, x, um);
    for(int r = 0; r < 9; r++) fu_n(r, i) = d[r] / (2 * deriv_eps);
  }
  double ex = std::sqrt((fx_a - fx_n).squaredNorm()), eu = std::sqrt((fu_a - fu_n).squaredNorm());
  std::printf("centroidal_deriv_err %.3e %.3e\n", ex, eu);
  auto vertical = std::make_shared<DDPProblemVerticalMotion>("vertical_motion");
  std::printf("input_dims %d %d %d %d %d %d\n", problem->inputDim(0.0), problem->inputDim(1.5), problem->inputDim(2.0),
              vertical->inputDim(0.0), vertical->inputDim(2.5), vertical->inputDim(4.7));
  return (ex < 1e-6 && eu < 1e-6) ? 0 : 1;
}

static int checkDerivative()
{
  // TestDDPCartPole.CheckDerivative (TestDDPCartPole.cpp:609-649), through the host-side virtuals
  auto ddp_problem = std::make_shared<DDPProblemCartPole>("cartpole");
  double t = 0;
  DDPProblemCartPole::StateDimVector x;
  x[0] = 1.0, x[1] = -2.0, x[2] = 3.0, x[3] = -4.0;
  DDPProblemCartPole::InputDimVector u;
  u[0] = 10.0;
  DDPProblemCartPole::StateStateDimMatrix fx_a, fx_n;
  DDPProblemCartPole::StateInputDimMatrix fu_a, fu_n;
  ddp_problem->calcStateEqDeriv(t, x, u, fx_a, fu_a);
  constexpr double deriv_eps = 1e-6;
  for(int i = 0; i < ddp_problem->stateDim(); i++)
  {
    auto xp = x, xm = x;
    xp[i] += deriv_eps, xm[i] -= deriv_eps;
    auto d = ddp_problem->stateEq(t, xp, u) - ddp_problem->stateEq(t, xm, u);
    for(int r = 0; r < 4; r++) fx_n(r, i) = d[r] / (2 * deriv_eps);
  }
  {
    auto up = u, um = u;
    up[0] += deriv_eps, um[0] -= deriv_eps;
    auto d = ddp_problem->stateEq(t, x, up) - ddp_problem->stateEq(t, x, um);
    for(int r = 0; r < 4; r++) fu_n(r, 0) = d[r] / (2 * deriv_eps);
  }
  double ex = std::sqrt((fx_a - fx_n).squaredNorm()), eu = std::sqrt((fu_a - fu_n).squaredNorm());
  std::printf("deriv_err %.3e %.3e\n", ex, eu);
  return (ex < 1e-6 && eu < 1e-6) ? 0 : 1;
}

int main(int argc, char ** argv)
{
  int rc = checkDerivative();
  rc |= checkCentroidal();
  if(argc > 1 && std::strcmp(argv[1], "--host-only") == 0)
  {
    // no GPU: creating a solver must fail loudly (no CPU fallback)
    try
    {
      auto problem = std::make_shared<DDPProblemCartPole>("cartpole");
      nmpc_ddp::DDPSolver<4, 1> solver(problem);
      std::vector<DDPProblemCartPole::InputDimVector> u(100, DDPProblemCartPole::InputDimVector::Zero());
      solver.solve(0.0, DDPProblemCartPole::StateDimVector::Zero(), u);
      std::printf("host_only solved\n");
    }
    catch(const std::runtime_error & e)
    {
      std::printf("host_only_error %s\n", e.what());
    }
    return rc;
  }

  // ---- DDP, as TestDDPCartPole sets it up but unconstrained, N = 100, max_iter = 10 ----
  auto ddp_problem = std::make_shared<DDPProblemCartPole>("cartpole");
  auto ddp_solver = std::make_shared<nmpc_ddp::DDPSolver<4, 1>>(ddp_problem);
  ddp_solver->config().horizon_steps = 100;
  ddp_solver->config().max_iter = 10;
  DDPProblemCartPole::StateDimVector current_x;
  current_x[0] = 0, current_x[1] = M_PI, current_x[2] = 0, current_x[3] = 0;
  std::vector<DDPProblemCartPole::InputDimVector> initial_u_list(100, DDPProblemCartPole::InputDimVector::Zero());
  bool converged = ddp_solver->solve(0.0, current_x, initial_u_list);
  std::printf("ddp_converged %d\n", converged ? 1 : 0);
  std::printf("ddp_trace_cost");
  for(const auto & tr : ddp_solver->traceDataList()) std::printf(" %.17g", tr.cost);
  std::printf("\nddp_u");
  for(const auto & u : ddp_solver->controlData().u_list) std::printf(" %.17g", u[0]);
  std::printf("\nddp_cost_sum %.17g\n", ddp_solver->controlData().cost_list.sum());
  std::printf("ddp_duration_ms %.4f %.4f %.4f %.4f\n", ddp_solver->computationDuration().solve,
              ddp_solver->computationDuration().derivative, ddp_solver->computationDuration().backward,
              ddp_solver->computationDuration().forward);
  ddp_solver->dumpTraceDataList("/tmp/nmpc_b200_TestDDPCartPoleTraceData.txt");

  // warm start with the previous u_list like mpcTimerCallback (TestDDPCartPole.cpp:392-395)
  bool again = ddp_solver->solve(0.0, current_x, ddp_solver->controlData().u_list);
  std::printf("ddp_warm_iters %d %d\n", ddp_solver->traceDataList().back().iter, again ? 1 : 0);

  // the test's MPC loop (TestDDPCartPole.cpp:313-343 + :388-396) for a batch of 2, 5 ticks on the device:
  // max_iter 3, input limits +-15 N, plant at sim_dt = 2 ms twice per 4 ms tick, applied input clamped
  {
    auto mpc_solver = std::make_shared<nmpc_ddp::DDPSolver<4, 1>>(ddp_problem, 2);
    mpc_solver->config().horizon_steps = 200;
    mpc_solver->config().max_iter = 3;
    mpc_solver->config().with_input_constraint = true;
    mpc_solver->setInputLimitsFunc([](double) {
      std::array<DDPProblemCartPole::InputDimVector, 2> limits;
      limits[0][0] = -15.0;
      limits[1][0] = 15.0;
      return limits;
    });
    nmpc_b200_mpc_config mpc{};
    mpc.n_ticks = 5, mpc.plant = 1, mpc.shift_inputs = 0, mpc.clamp_u0 = 1, mpc.n_substeps = 2;
    mpc.tick_dt = 0.004, mpc.sim_dt = 0.002;
    const double x0[8] = {0, M_PI, 0, 0, 0.5, 2.0, 0, 0};
    std::vector<double> u_init(2 * 200, 0.0), x_log(2 * 6 * 4), u_log(2 * 5);
    std::vector<int> iters(2 * 5);
    mpc_solver->runMpc(2, 0.0, x0, u_init.data(), 200, mpc, x_log.data(), u_log.data(), iters.data());
    std::printf("mpc_u");
    for(double v : u_log) std::printf(" %.17g", v);
    std::printf("\nmpc_x_final");
    for(int b = 0; b < 2; b++)
      for(int d = 0; d < 4; d++) std::printf(" %.17g", x_log[(b * 6 + 5) * 4 + d]);
    std::printf("\n");
  }

  // wrong initial_u_list length => std::invalid_argument (DDPSolver.hpp:41-45)
  try
  {
    initial_u_list.pop_back();
    ddp_solver->solve(0.0, current_x, initial_u_list);
    std::printf("ddp_invalid_argument none\n");
    rc = 1;
  }
  catch(const std::invalid_argument & e)
  {
    std::printf("ddp_invalid_argument %s\n", e.what());
  }

  // ---- FMPC, as TestFmpcCartPole sets it up: N = 100, max_iter = 5, Variable.reset(0,0,0,1,1) ----
  auto fmpc_problem = std::make_shared<FmpcProblemCartPole>("cartpole");
  auto fmpc_solver = std::make_shared<nmpc_fmpc::FmpcSolver<4, 1, 4>>(fmpc_problem);
  fmpc_solver->config().horizon_steps = 100;
  fmpc_solver->config().max_iter = 5;
  using Variable = nmpc_fmpc::FmpcSolver<4, 1, 4>::Variable;
  Variable variable(100);
  variable.reset(0.0, 0.0, 0.0, 1e0, 1e0);
  auto status = fmpc_solver->solve(0.0, current_x, variable);
  std::printf("fmpc_status %d\n", static_cast<int>(status));
  std::printf("fmpc_kkt");
  for(const auto & tr : fmpc_solver->traceDataList()) std::printf(" %.17g", tr.kkt_error);
  std::printf("\nfmpc_u");
  for(const auto & u : fmpc_solver->variable().u_list) std::printf(" %.17g", u[0]);
  std::printf("\nfmpc_K0");
  for(int d = 0; d < 4; d++) std::printf(" %.17g", fmpc_solver->coeffList().front().K(0, d));
  std::printf("\n");
  try
  {
    Variable bad(99);
    bad.reset(0, 0, 0, 1, 1);
    fmpc_solver->solve(0.0, current_x, bad);
    rc = 1;
  }
  catch(const std::invalid_argument & e)
  {
    std::printf("fmpc_invalid_argument %s\n", e.what());
  }
  try
  {
    variable.s_list[3][1] = -1.0;
    fmpc_solver->solve(0.0, current_x, variable);
    rc = 1;
  }
  catch(const std::runtime_error & e)
  {
    std::printf("fmpc_runtime_error %s\n", e.what());
  }
  return rc;
}
