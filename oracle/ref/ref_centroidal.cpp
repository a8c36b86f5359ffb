// TEST INFRASTRUCTURE ONLY -- the reference's DDPSolver<9, Eigen::Dynamic> (input dimension 16 or 0 along the horizon)
// on the problem of nmpc_ddp/tests/src/TestDDPCentroidalMotion.cpp:18-201 and the test's MPC loop (:238-353).  The
// problem bodies are written coefficient by coefficient in the evaluation order of the test's Eigen expressions (the
// Eigen stand-in of eigen_shim/ has no Matrix3Xd / cross / segment); the SOLVER is the reference's own header.
// Pins the padded-dimension implementation of the oracle and of the device for n_x = 9, n_u = 16.
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
struct StanceData
{
  int n = 0; // number of columns of vertices_mat / ridges_mat
  double vertices[16][3];
  double ridges[16][3];
};

/** makeStanceDataFromRect (:203-236) */
StanceData makeStanceDataFromRect(double min_x, double min_y, double max_x, double max_y)
{
  const double vertex_list[4][3] = {{min_x, min_y, 0.0}, {min_x, max_y, 0.0}, {max_x, max_y, 0.0}, {max_x, min_y, 0.0}};
  double ridge_list[4][3];
  for(int i = 0; i < 4; i++)
  {
    double theta = 2 * M_PI * (static_cast<double>(i) / 4);
    double v[3] = {0.5 * std::cos(theta), 0.5 * std::sin(theta), 1};
    double n = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    for(int k = 0; k < 3; k++) ridge_list[i][k] = v[k] / n;
  }
  StanceData s;
  s.n = 16;
  int col_idx = 0;
  for(int vi = 0; vi < 4; vi++)
    for(int ri = 0; ri < 4; ri++)
    {
      for(int k = 0; k < 3; k++) s.vertices[col_idx][k] = vertex_list[vi][k], s.ridges[col_idx][k] = ridge_list[ri][k];
      col_idx++;
    }
  return s;
}

/** ref_stance_func of TEST(TestDDPCentroidalMotion, SolveMpc) (:246-266) */
StanceData refStance(double t)
{
  constexpr double epsilon_t = 1e-6;
  t += epsilon_t;
  if(t < 1.4) return makeStanceDataFromRect(-0.1, -0.1, 0.1, 0.1);
  if(t < 1.6) return StanceData();
  return makeStanceDataFromRect(0.4, -0.1, 0.6, 0.1);
}

/** ref_pos_func (:267-279) */
std::array<double, 3> refPos(double t)
{
  constexpr double epsilon_t = 1e-6;
  t += epsilon_t;
  if(t < 1.5) return {0.0, 0.0, 1.0};
  return {0.5, 0.0, 1.0};
}

void cross(const double a[3], const double b[3], double out[3])
{
  out[0] = a[1] * b[2] - a[2] * b[1];
  out[1] = a[2] * b[0] - a[0] * b[2];
  out[2] = a[0] * b[1] - a[1] * b[0];
}

class DDPProblemCentroidalMotion : public nmpc_ddp::DDPProblem<9, Eigen::Dynamic>
{
public:
  explicit DDPProblemCentroidalMotion(double dt) : DDPProblem(dt) {}
  using DDPProblem::inputDim;
  int inputDim(double t) const override
  {
    return refStance(t).n;
  }
  StateDimVector stateEq(double t, const StateDimVector & x, const InputDimVector & u) const override
  {
    const StanceData s = refStance(t);
    double force[3] = {0, 0, 0}, am_dot[3] = {0, 0, 0};
    for(int i = 0; i < u.size(); i++)
    {
      double arm[3] = {s.vertices[i][0] - x[0], s.vertices[i][1] - x[1], s.vertices[i][2] - x[2]}, c[3];
      cross(arm, s.ridges[i], c);
      for(int k = 0; k < 3; k++)
      {
        force[k] += s.ridges[i][k] * u[i];
        am_dot[k] += u[i] * c[k];
      }
    }
    StateDimVector out;
    for(int k = 0; k < 3; k++)
    {
      out[k] = x[k] + dt_ * (x[3 + k] / mass_);
      out[3 + k] = x[3 + k] + dt_ * (force[k] - mass_ * (k == 2 ? g_ : 0.0));
      out[6 + k] = x[6 + k] + dt_ * am_dot[k];
    }
    return out;
  }
  double weightedSquares(const double w[9], double t, const StateDimVector & x) const
  {
    const auto rp = refPos(t);
    double s = 0;
    for(int k = 0; k < 9; k++)
    {
      double e = k < 3 ? x[k] - rp[k] : x[k];
      s += w[k] * (e * e);
    }
    return s;
  }
  double runningCost(double t, const StateDimVector & x, const InputDimVector & u) const override
  {
    double sq = 0;
    for(int i = 0; i < u.size(); i++) sq += u[i] * u[i];
    return 0.5 * weightedSquares(running_x, t, x) + 0.5 * running_u * sq;
  }
  double terminalCost(double t, const StateDimVector & x) const override
  {
    return 0.5 * weightedSquares(terminal_x, t, x);
  }
  void calcStateEqDeriv(double t, const StateDimVector & x, const InputDimVector & u, Eigen::Ref<StateStateDimMatrix> Fx,
                        Eigen::Ref<StateInputDimMatrix> Fu) const override
  {
    const StanceData s = refStance(t);
    double force[3] = {0, 0, 0};
    Fu.setZero();
    for(int i = 0; i < u.size(); i++)
    {
      double arm[3] = {s.vertices[i][0] - x[0], s.vertices[i][1] - x[1], s.vertices[i][2] - x[2]}, c[3];
      cross(arm, s.ridges[i], c);
      for(int k = 0; k < 3; k++)
      {
        force[k] += s.ridges[i][k] * u[i];
        Fu(3 + k, i) = s.ridges[i][k];
        Fu(6 + k, i) = c[k];
      }
    }
    Fu *= dt_;
    Fx.setZero();
    for(int k = 0; k < 3; k++) Fx(k, 3 + k) = 1 / mass_;
    Fx(6, 1) = -force[2], Fx(6, 2) = force[1];
    Fx(7, 0) = force[2], Fx(7, 2) = -force[0];
    Fx(8, 0) = -force[1], Fx(8, 1) = force[0];
    Fx *= dt_;
    for(int k = 0; k < 9; k++) Fx(k, k) += 1.0;
  }
  void calcStateEqDeriv(double, const StateDimVector &, const InputDimVector &, Eigen::Ref<StateStateDimMatrix>,
                        Eigen::Ref<StateInputDimMatrix>, std::vector<StateStateDimMatrix> &,
                        std::vector<InputInputDimMatrix> &, std::vector<StateInputDimMatrix> &) const override
  {
    throw std::runtime_error("Second-order derivatives of state equation are not implemented.");
  }
  void calcRunningCostDeriv(double t, const StateDimVector & x, const InputDimVector & u, Eigen::Ref<StateDimVector> Lx,
                            Eigen::Ref<InputDimVector> Lu) const override
  {
    const auto rp = refPos(t);
    for(int k = 0; k < 9; k++) Lx[k] = running_x[k] * (k < 3 ? x[k] - rp[k] : x[k]);
    for(int i = 0; i < u.size(); i++) Lu[i] = running_u * u[i];
  }
  void calcRunningCostDeriv(double t, const StateDimVector & x, const InputDimVector & u, Eigen::Ref<StateDimVector> Lx,
                            Eigen::Ref<InputDimVector> Lu, Eigen::Ref<StateStateDimMatrix> Lxx,
                            Eigen::Ref<InputInputDimMatrix> Luu, Eigen::Ref<StateInputDimMatrix> Lxu) const override
  {
    calcRunningCostDeriv(t, x, u, Lx, Lu);
    Lxx.setZero();
    for(int k = 0; k < 9; k++) Lxx(k, k) = running_x[k];
    Luu.setIdentity();
    Luu *= running_u;
    Lxu.setZero();
  }
  void calcTerminalCostDeriv(double t, const StateDimVector & x, Eigen::Ref<StateDimVector> Vx) const override
  {
    const auto rp = refPos(t);
    for(int k = 0; k < 9; k++) Vx[k] = terminal_x[k] * (k < 3 ? x[k] - rp[k] : x[k]);
  }
  void calcTerminalCostDeriv(double t, const StateDimVector & x, Eigen::Ref<StateDimVector> Vx,
                             Eigen::Ref<StateStateDimMatrix> Vxx) const override
  {
    calcTerminalCostDeriv(t, x, Vx);
    Vxx.setZero();
    for(int k = 0; k < 9; k++) Vxx(k, k) = terminal_x[k];
  }

protected:
  static constexpr double g_ = 9.80665;
  double running_x[9] = {1, 1, 1, 0, 0, 0, 1, 1, 1}; // CostWeight (:39-51)
  double running_u = 1e-6;
  double terminal_x[9] = {1, 1, 1, 0, 0, 0, 1, 1, 1};
  double mass_ = 100.0;
};
} // namespace

extern "C"
{
/** TestDDPCentroidalMotion's loop (:238-353) for `n_ticks` ticks.  first_max_iter: max_iter of the first solve (the
    test leaves the default, 500), 3 afterwards (:303).  Outputs per tick: x_log[tick][9] = current_x,
    u0_log[tick][16] = u_list[0] padded with zeros, dim_log[tick], iters_log[tick]; of the FIRST solve
    x_first[N+1][9], u_first[N][16], cost_first; of the LAST solve x_out / u_out (same shapes). */
int ref_centroidal_mpc(int horizon_steps, int first_max_iter, int n_ticks, double * x_log, double * u0_log,
                       int * dim_log, int * iters_log, double * x_first, double * u_first, double * cost_first,
                       double * x_out, double * u_out)
{
  const double dt = 0.03;
  auto problem = std::make_shared<DDPProblemCentroidalMotion>(dt);
  auto solver = std::make_shared<nmpc_ddp::DDPSolver<9, Eigen::Dynamic>>(problem);
  solver->config().print_level = 0;
  solver->config().horizon_steps = horizon_steps;
  solver->config().max_iter = first_max_iter;

  double current_t = 0;
  DDPProblemCentroidalMotion::StateDimVector current_x;
  current_x.setZero();
  current_x[2] = 1.0;
  std::vector<DDPProblemCentroidalMotion::InputDimVector> current_u_list;
  for(int i = 0; i < horizon_steps; i++)
    current_u_list.push_back(DDPProblemCentroidalMotion::InputDimVector::Zero(problem->inputDim(current_t + i * dt)));

  auto dump = [&](double * xo, double * uo)
  {
    const auto & cd = solver->controlData();
    for(int i = 0; i <= horizon_steps; i++)
      for(int k = 0; k < 9; k++) xo[9 * i + k] = cd.x_list[i][k];
    for(int i = 0; i < horizon_steps; i++)
      for(int k = 0; k < 16; k++) uo[16 * i + k] = k < cd.u_list[i].size() ? cd.u_list[i][k] : 0.0;
  };

  std::streambuf * old = std::cout.rdbuf(nullptr);
  try
  {
    for(int tick = 0; tick < n_ticks; tick++)
    {
      solver->solve(current_t, current_x, current_u_list);
      solver->config().max_iter = 3;
      const auto & cd = solver->controlData();
      for(int k = 0; k < 9; k++) x_log[9 * tick + k] = current_x[k];
      const auto & u0 = cd.u_list[0];
      for(int k = 0; k < 16; k++) u0_log[16 * tick + k] = k < u0.size() ? u0[k] : 0.0;
      dim_log[tick] = (int)u0.size();
      iters_log[tick] = solver->traceDataList().back().iter;
      if(tick == 0)
      {
        dump(x_first, u_first);
        *cost_first = cd.cost_list.sum();
      }
      if(tick == n_ticks - 1) dump(x_out, u_out);
      current_x = cd.x_list[1];
      current_u_list = cd.u_list;
      current_u_list.erase(current_u_list.begin());
      double terminal_t = current_t + horizon_steps * dt;
      int terminal_input_dim = problem->inputDim(terminal_t);
      if(current_u_list.back().size() == terminal_input_dim)
        current_u_list.push_back(current_u_list.back());
      else
        current_u_list.push_back(DDPProblemCentroidalMotion::InputDimVector::Zero(terminal_input_dim));
      current_t += dt;
    }
  }
  catch(const std::exception & e)
  {
    std::cout.rdbuf(old);
    std::cerr << "ref_centroidal_mpc: " << e.what() << std::endl;
    return -1;
  }
  std::cout.rdbuf(old);
  return 0;
}
} // extern "C"
