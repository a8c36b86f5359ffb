// TEST INFRASTRUCTURE ONLY -- the reference's DDPSolver<2, Eigen::Dynamic> (time-varying input dimension) on the
// problem of nmpc_ddp/tests/src/TestDDPVerticalMotion.cpp:25-234, bodies restated in the test's own Eigen idioms,
// and the test's MPC loop (:236-330).  Pins the padded-dimension implementation of the oracle and of the device.
#include <array>
#include <cmath>
#include <cstring>
#include <functional>
#include <iostream>
#include <memory>
#include <vector>

#include <nmpc_ddp/DDPSolver.h>

namespace
{
class DDPProblemVerticalMotion : public nmpc_ddp::DDPProblem<2, Eigen::Dynamic>
{
public:
  explicit DDPProblemVerticalMotion(double dt) : DDPProblem(dt)
  {
    running_x << 1.0, 1e-3;
    terminal_x << 1.0, 1e-3;
  }
  using DDPProblem::inputDim;
  int inputDim(double t) const override
  {
    constexpr double epsilon_t = 1e-6;
    t += epsilon_t;
    if(2.0 < t && t < 3.0) return 2;
    if(4.5 < t && t < 5.0) return 0;
    return 1;
  }
  static double refPos(double t)
  {
    t += 1e-6;
    return t < 8.0 ? 1.0 : 0.0;
  }
  StateDimVector stateEq(double, const StateDimVector & x, const InputDimVector & u) const override
  {
    StateDimVector x_dot;
    x_dot << x[1], u.sum() / mass_ - g_;
    return x + dt_ * x_dot;
  }
  double runningCost(double t, const StateDimVector & x, const InputDimVector & u) const override
  {
    StateDimVector ref_x;
    ref_x << refPos(t), 0;
    double cost_x = 0.5 * running_x.dot((x - ref_x).cwiseAbs2());
    double cost_u = 0.5 * running_u * u.squaredNorm();
    return cost_x + cost_u;
  }
  double terminalCost(double t, const StateDimVector & x) const override
  {
    StateDimVector ref_x;
    ref_x << refPos(t), 0;
    return 0.5 * terminal_x.dot((x - ref_x).cwiseAbs2());
  }
  void calcStateEqDeriv(double, const StateDimVector &, const InputDimVector &, Eigen::Ref<StateStateDimMatrix> Fx,
                        Eigen::Ref<StateInputDimMatrix> Fu) const override
  {
    Fx << 0, 1, 0, 0;
    Fx *= dt_;
    Fx.diagonal().array() += 1.0;
    Fu.row(0).setZero();
    Fu.row(1).setConstant(1.0 / mass_);
    Fu *= dt_;
  }
  void calcStateEqDeriv(double t, const StateDimVector & x, const InputDimVector & u, Eigen::Ref<StateStateDimMatrix> Fx,
                        Eigen::Ref<StateInputDimMatrix> Fu, std::vector<StateStateDimMatrix> &,
                        std::vector<InputInputDimMatrix> &, std::vector<StateInputDimMatrix> &) const override
  {
    calcStateEqDeriv(t, x, u, Fx, Fu);
  }
  void calcRunningCostDeriv(double t, const StateDimVector & x, const InputDimVector & u, Eigen::Ref<StateDimVector> Lx,
                            Eigen::Ref<InputDimVector> Lu) const override
  {
    StateDimVector ref_x;
    ref_x << refPos(t), 0;
    Lx = running_x.cwiseProduct(x - ref_x);
    Lu = running_u * u;
  }
  void calcRunningCostDeriv(double t, const StateDimVector & x, const InputDimVector & u, Eigen::Ref<StateDimVector> Lx,
                            Eigen::Ref<InputDimVector> Lu, Eigen::Ref<StateStateDimMatrix> Lxx,
                            Eigen::Ref<InputInputDimMatrix> Luu, Eigen::Ref<StateInputDimMatrix> Lxu) const override
  {
    StateDimVector ref_x;
    ref_x << refPos(t), 0;
    Lx = running_x.cwiseProduct(x - ref_x);
    Lxx = running_x.asDiagonal();
    Lxu.setZero();
    Lu = running_u * u;
    Luu.setIdentity();
    Luu *= running_u;
  }
  void calcTerminalCostDeriv(double t, const StateDimVector & x, Eigen::Ref<StateDimVector> Vx) const override
  {
    StateDimVector ref_x;
    ref_x << refPos(t), 0;
    Vx = terminal_x.cwiseProduct(x - ref_x);
  }
  void calcTerminalCostDeriv(double t, const StateDimVector & x, Eigen::Ref<StateDimVector> Vx,
                             Eigen::Ref<StateStateDimMatrix> Vxx) const override
  {
    StateDimVector ref_x;
    ref_x << refPos(t), 0;
    Vx = terminal_x.cwiseProduct(x - ref_x);
    Vxx = terminal_x.asDiagonal();
  }

protected:
  static constexpr double g_ = 9.80665;
  StateDimVector running_x, terminal_x;
  double running_u = 1e-4;
  double mass_ = 1.0;
};
} // namespace

extern "C"
{
/** TestDDPVerticalMotion's loop (TestDDPVerticalMotion.cpp:236-330) for `n_ticks` ticks from (t0, x0).  Outputs per
    tick: x_log[tick][2] = current_x, u0_log[tick][2] = u_list[0] padded with zeros, dim_log[tick] = u_list[0].size(),
    iters_log[tick]; of the LAST solve: x_out[N+1][2], u_out[N][2] (padded).  max_iter: default (500) for the first
    solve, 3 afterwards, as in the test (:287). */
int ref_vertical_mpc(int horizon_steps, int with_constraint, int n_ticks, double t0, const double * x0, double * x_log,
                     double * u0_log, int * dim_log, int * iters_log, double * x_out, double * u_out)
{
  const double dt = 0.01;
  auto problem = std::make_shared<DDPProblemVerticalMotion>(dt);
  auto solver = std::make_shared<nmpc_ddp::DDPSolver<2, Eigen::Dynamic>>(problem);
  solver->setInputLimitsFunc([&](double t) -> std::array<Eigen::VectorXd, 2> {
    std::array<Eigen::VectorXd, 2> limits;
    int input_dim = problem->inputDim(t);
    limits[0].setConstant(input_dim, 0.0);
    limits[1].setConstant(input_dim, 30.0);
    return limits;
  });
  solver->config().print_level = 0;
  solver->config().with_input_constraint = with_constraint != 0;
  solver->config().horizon_steps = horizon_steps;
  solver->config().initial_lambda = 1e-6;

  double current_t = t0;
  DDPProblemVerticalMotion::StateDimVector current_x(x0[0], x0[1]);
  std::vector<DDPProblemVerticalMotion::InputDimVector> current_u_list;
  for(int i = 0; i < horizon_steps; i++)
    current_u_list.push_back(DDPProblemVerticalMotion::InputDimVector::Zero(problem->inputDim(current_t + i * dt)));

  std::streambuf * old = std::cout.rdbuf(nullptr);
  int rc = 0;
  try
  {
    for(int tick = 0; tick < n_ticks; tick++)
    {
      solver->solve(current_t, current_x, current_u_list);
      solver->config().max_iter = 3;
      const auto & cd = solver->controlData();
      x_log[2 * tick] = current_x[0], x_log[2 * tick + 1] = current_x[1];
      const auto & u0 = cd.u_list[0];
      u0_log[2 * tick] = u0.size() > 0 ? u0[0] : 0.0;
      u0_log[2 * tick + 1] = u0.size() > 1 ? u0[1] : 0.0;
      dim_log[tick] = (int)u0.size();
      iters_log[tick] = solver->traceDataList().back().iter;
      if(tick == n_ticks - 1)
      {
        for(int i = 0; i <= horizon_steps; i++) x_out[2 * i] = cd.x_list[i][0], x_out[2 * i + 1] = cd.x_list[i][1];
        for(int i = 0; i < horizon_steps; i++)
        {
          u_out[2 * i] = cd.u_list[i].size() > 0 ? cd.u_list[i][0] : 0.0;
          u_out[2 * i + 1] = cd.u_list[i].size() > 1 ? cd.u_list[i][1] : 0.0;
        }
      }
      current_x = cd.x_list[1];
      current_u_list = cd.u_list;
      current_u_list.erase(current_u_list.begin());
      double terminal_t = current_t + horizon_steps * dt;
      int terminal_input_dim = problem->inputDim(terminal_t);
      if(current_u_list.back().size() == terminal_input_dim)
        current_u_list.push_back(current_u_list.back());
      else
        current_u_list.push_back(DDPProblemVerticalMotion::InputDimVector::Zero(terminal_input_dim));
      current_t += dt;
    }
  }
  catch(const std::exception & e)
  {
    std::cout.rdbuf(old);
    std::cerr << "ref_vertical_mpc: " << e.what() << std::endl;
    return -1;
  }
  std::cout.rdbuf(old);
  return rc;
}
} // extern "C"
