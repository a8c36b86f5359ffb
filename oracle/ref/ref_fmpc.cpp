// TEST INFRASTRUCTURE ONLY -- builds oracle/_ref/libnmpc_ref.so: the REFERENCE's own solver headers
// (straight from /root/reference, unmodified) compiled against oracle/ref/eigen_shim, with the problem
// classes of the reference's tests restated in the tests' own Eigen idioms (ref_models.h).  Used to pin
// oracle/ (the plain restatement) against the reference's real control flow, and to generate
// tests/golden/*.npz.  DDPSolver.hpp and FmpcSolver.hpp each define calcDuration() in an anonymous
// namespace, so the two solvers live in separate translation units.
#include <cstring>
#include <iostream>
#include <memory>
#include <string>

#include <nmpc_fmpc/FmpcSolver.h>

#include "ref_models.h"

extern "C"
{
typedef struct
{
  int horizon_steps, max_iter, check_nan, init_complementary_variable, update_barrier_eps, break_if_llt_fails,
      enable_line_search, merit_const_scale_from_lagrange_multipliers;
  double kkt_error_thre, initial_barrier_eps;
} ref_fmpc_config;
}

namespace
{
/** nmpc_fmpc/tests/src/TestFmpcCartPole.cpp:32-256: the same bodies plus ineqConst / calcIneqConstDeriv. */
class FmpcProblemCartPole : public CartPoleBodies<nmpc_fmpc::FmpcProblem<4, 1, 4>>
{
public:
  static constexpr bool kDynamicIneq = false;
  using CartPoleBodies<nmpc_fmpc::FmpcProblem<4, 1, 4>>::CartPoleBodies;
  IneqDimVector ineqConst(double, const StateDimVector & x, const InputDimVector & u) const override
  {
    constexpr double u_max = 15.0;
    constexpr double u_min = -1 * u_max;
    constexpr double x_max = 20.0;
    constexpr double x_min = -20.0;
    IneqDimVector g;
    g[0] = -1 * u[0] + u_min;
    g[1] = u[0] - u_max;
    g[2] = -1 * x[0] + x_min;
    g[3] = x[0] - x_max;
    return g;
  }
  void calcIneqConstDeriv(double,
                          const StateDimVector &,
                          const InputDimVector &,
                          Eigen::Ref<IneqStateDimMatrix> ineq_const_deriv_x,
                          Eigen::Ref<IneqInputDimMatrix> ineq_const_deriv_u) const override
  {
    ineq_const_deriv_x.setZero();
    ineq_const_deriv_x(2, 0) = -1;
    ineq_const_deriv_x(3, 0) = 1;
    ineq_const_deriv_u.setZero();
    ineq_const_deriv_u(0, 0) = -1;
    ineq_const_deriv_u(1, 0) = 1;
  }
};

/** nmpc_fmpc/tests/src/TestFmpcOscillator.cpp:18-135. */
class FmpcProblemOscillator : public nmpc_fmpc::FmpcProblem<2, 1, 3>
{
public:
  static constexpr bool kDynamicIneq = false;
  explicit FmpcProblemOscillator(const double * p) : FmpcProblem(p[0]) {}
  StateDimVector stateEq(double t, const StateDimVector & x, const InputDimVector & u) const override
  {
    return stateEq(t, x, u, dt_);
  }
  StateDimVector stateEq(double, const StateDimVector & x, const InputDimVector & u, double dt) const
  {
    StateDimVector x_dot;
    x_dot << (1.0 - std::pow(x[1], 2)) * x[0] - x[1] + u[0], x[0];
    return x + dt * x_dot;
  }
  double runningCost(double, const StateDimVector & x, const InputDimVector & u) const override
  {
    return 0.5 * (x.squaredNorm() + u.squaredNorm());
  }
  double terminalCost(double, const StateDimVector &) const override
  {
    return 0;
  }
  IneqDimVector ineqConst(double, const StateDimVector & x, const InputDimVector & u) const override
  {
    IneqDimVector g;
    g[0] = -1 * x[1] - 0.05;
    g[1] = -1 * u[0] - 1.0;
    g[2] = u[0] - 0.9;
    return g;
  }
  void calcStateEqDeriv(double,
                        const StateDimVector & x,
                        const InputDimVector &,
                        Eigen::Ref<StateStateDimMatrix> state_eq_deriv_x,
                        Eigen::Ref<StateInputDimMatrix> state_eq_deriv_u) const override
  {
    state_eq_deriv_x.setZero();
    state_eq_deriv_x(0, 0) = 1.0 - std::pow(x[1], 2);
    state_eq_deriv_x(0, 1) = -2 * x[0] * x[1] - 1.0;
    state_eq_deriv_x(1, 0) = 1;
    state_eq_deriv_x *= dt_;
    state_eq_deriv_x.diagonal().array() += 1;
    state_eq_deriv_u.setZero();
    state_eq_deriv_u(0, 0) = 1;
    state_eq_deriv_u *= dt_;
  }
  void calcRunningCostDeriv(double,
                            const StateDimVector & x,
                            const InputDimVector & u,
                            Eigen::Ref<StateDimVector> running_cost_deriv_x,
                            Eigen::Ref<InputDimVector> running_cost_deriv_u) const override
  {
    running_cost_deriv_x = x;
    running_cost_deriv_u = u;
  }
  void calcRunningCostDeriv(double t,
                            const StateDimVector & x,
                            const InputDimVector & u,
                            Eigen::Ref<StateDimVector> running_cost_deriv_x,
                            Eigen::Ref<InputDimVector> running_cost_deriv_u,
                            Eigen::Ref<StateStateDimMatrix> running_cost_deriv_xx,
                            Eigen::Ref<InputInputDimMatrix> running_cost_deriv_uu,
                            Eigen::Ref<StateInputDimMatrix> running_cost_deriv_xu) const override
  {
    calcRunningCostDeriv(t, x, u, running_cost_deriv_x, running_cost_deriv_u);
    running_cost_deriv_xx.setIdentity();
    running_cost_deriv_uu.setIdentity();
    running_cost_deriv_xu.setZero();
  }
  void calcTerminalCostDeriv(double, const StateDimVector &, Eigen::Ref<StateDimVector> terminal_cost_deriv_x)
      const override
  {
    terminal_cost_deriv_x.setZero();
  }
  void calcTerminalCostDeriv(double t,
                             const StateDimVector & x,
                             Eigen::Ref<StateDimVector> terminal_cost_deriv_x,
                             Eigen::Ref<StateStateDimMatrix> terminal_cost_deriv_xx) const override
  {
    calcTerminalCostDeriv(t, x, terminal_cost_deriv_x);
    terminal_cost_deriv_xx.setZero();
  }
  void calcIneqConstDeriv(double,
                          const StateDimVector &,
                          const InputDimVector &,
                          Eigen::Ref<IneqStateDimMatrix> ineq_const_deriv_x,
                          Eigen::Ref<IneqInputDimMatrix> ineq_const_deriv_u) const override
  {
    ineq_const_deriv_x.setZero();
    ineq_const_deriv_x(0, 1) = -1;
    ineq_const_deriv_u.setZero();
    ineq_const_deriv_u(1, 0) = -1;
    ineq_const_deriv_u(2, 0) = 1;
  }
};

/** ... as an FMPC problem: 0 <= T_a <= thrust_max. */
class FmpcProblemPlanarQuadrotor : public PlanarQuadrotorBodies<nmpc_fmpc::FmpcProblem<6, 2, 4>>
{
public:
  static constexpr bool kDynamicIneq = false;
  using PlanarQuadrotorBodies<nmpc_fmpc::FmpcProblem<6, 2, 4>>::PlanarQuadrotorBodies;
  IneqDimVector ineqConst(double, const StateDimVector &, const InputDimVector & u) const override
  {
    IneqDimVector g;
    g << -1 * u[0], u[0] - thrust_max_, -1 * u[1], u[1] - thrust_max_;
    return g;
  }
  void calcIneqConstDeriv(double,
                          const StateDimVector &,
                          const InputDimVector &,
                          Eigen::Ref<IneqStateDimMatrix> ineq_const_deriv_x,
                          Eigen::Ref<IneqInputDimMatrix> ineq_const_deriv_u) const override
  {
    ineq_const_deriv_x.setZero();
    ineq_const_deriv_u.setZero();
    ineq_const_deriv_u(0, 0) = -1;
    ineq_const_deriv_u(1, 0) = 1;
    ineq_const_deriv_u(2, 1) = -1;
    ineq_const_deriv_u(3, 1) = 1;
  }

};

/** Cart-pole whose position limits exist only for window_start <= t < window_end: the reference's DYNAMIC inequality
    dimension, nmpc_fmpc::FmpcProblem<4, 1, Eigen::Dynamic> with ineqDim(t) overridden (FmpcProblem.h:62-86).
    params: the cart-pole's 14, then [window_start, window_end]. */
class FmpcProblemCartPoleWindowed : public CartPoleBodies<nmpc_fmpc::FmpcProblem<4, 1, Eigen::Dynamic>>
{
public:
  using Base = CartPoleBodies<nmpc_fmpc::FmpcProblem<4, 1, Eigen::Dynamic>>;
  static constexpr bool kDynamicIneq = true;
  explicit FmpcProblemCartPoleWindowed(const double * p) : Base(p), window_start_(p[14]), window_end_(p[15]) {}
  int ineqDim(double t) const override
  {
    return (t >= window_start_ && t < window_end_) ? 4 : 2;
  }
  IneqDimVector ineqConst(double t, const StateDimVector & x, const InputDimVector & u) const override
  {
    constexpr double u_max = 15.0;
    constexpr double u_min = -1 * u_max;
    constexpr double x_max = 20.0;
    constexpr double x_min = -20.0;
    IneqDimVector g(ineqDim(t));
    g[0] = -1 * u[0] + u_min;
    g[1] = u[0] - u_max;
    if(ineqDim(t) == 4)
    {
      g[2] = -1 * x[0] + x_min;
      g[3] = x[0] - x_max;
    }
    return g;
  }
  void calcIneqConstDeriv(double t,
                          const StateDimVector &,
                          const InputDimVector &,
                          Eigen::Ref<IneqStateDimMatrix> ineq_const_deriv_x,
                          Eigen::Ref<IneqInputDimMatrix> ineq_const_deriv_u) const override
  {
    ineq_const_deriv_x.setZero();
    ineq_const_deriv_u.setZero();
    ineq_const_deriv_u(0, 0) = -1;
    ineq_const_deriv_u(1, 0) = 1;
    if(ineqDim(t) == 4)
    {
      ineq_const_deriv_x(2, 0) = -1;
      ineq_const_deriv_x(3, 0) = 1;
    }
  }

protected:
  double window_start_, window_end_;
};

template<class Problem, int NX, int NU, int NG>
int fmpcSolve(const double * params,
                     const ref_fmpc_config * cfg,
                     double t0,
                     const double * x0,
                     const double * x_in,
                     const double * u_in,
                     const double * lambda_in,
                     const double * s_in,
                     const double * nu_in,
                     double * x_out,
                     double * u_out,
                     double * lambda_out,
                     double * s_out,
                     double * nu_out,
                     double * K_out,
                     double * kkt_out,
                     int * n_trace_out,
                     int * status_out)
{
  // NGS: the solver's inequality template parameter (NG, or Eigen::Dynamic with NG the padded width of the I/O arrays)
  constexpr int NGS = Problem::kDynamicIneq ? Eigen::Dynamic : NG;
  using Solver = nmpc_fmpc::FmpcSolver<NX, NU, NGS>;
  auto problem = std::make_shared<Problem>(params);
  Solver solver(problem);
  auto & c = solver.config();
  c.print_level = 0;
  c.horizon_steps = cfg->horizon_steps;
  c.max_iter = cfg->max_iter;
  c.kkt_error_thre = cfg->kkt_error_thre;
  c.check_nan = cfg->check_nan != 0;
  c.init_complementary_variable = cfg->init_complementary_variable != 0;
  c.update_barrier_eps = cfg->update_barrier_eps != 0;
  c.break_if_llt_fails = cfg->break_if_llt_fails != 0;
  c.enable_line_search = cfg->enable_line_search != 0;
  c.merit_const_scale_from_lagrange_multipliers = cfg->merit_const_scale_from_lagrange_multipliers != 0;
  const int N = cfg->horizon_steps;
  typename Solver::Variable var(N);
  for(int i = 0; i <= N; i++)
    for(int d = 0; d < NX; d++)
    {
      var.x_list[i][d] = x_in[i * NX + d];
      var.lambda_list[i][d] = lambda_in[i * NX + d];
    }
  for(int i = 0; i < N; i++)
  {
    for(int d = 0; d < NU; d++) var.u_list[i][d] = u_in[i * NU + d];
    const int ng_i = problem->ineqDim(t0 + i * problem->dt());
    if constexpr(Problem::kDynamicIneq)
    {
      var.s_list[i].resize(ng_i);
      var.nu_list[i].resize(ng_i);
    }
    for(int d = 0; d < ng_i; d++)
    {
      var.s_list[i][d] = s_in[i * NG + d];
      var.nu_list[i][d] = nu_in[i * NG + d];
    }
  }
  typename Problem::StateDimVector current_x;
  for(int d = 0; d < NX; d++) current_x[d] = x0[d];
  typename Solver::Status status;
  try
  {
    status = solver.solve(t0, current_x, var);
  }
  catch(...)
  {
    return -1;
  }
  const auto & v = solver.variable();
  for(int i = 0; i <= N; i++)
    for(int d = 0; d < NX; d++)
    {
      x_out[i * NX + d] = v.x_list[i][d];
      lambda_out[i * NX + d] = v.lambda_list[i][d];
    }
  for(int i = 0; i < N; i++)
  {
    for(int d = 0; d < NU; d++) u_out[i * NU + d] = v.u_list[i][d];
    for(int d = 0; d < NG; d++)
    {
      // rows beyond a step's dimension do not exist in the reference: reported as the neutral (s, nu) = (1, 0)
      const bool real = d < (int)v.s_list[i].size();
      s_out[i * NG + d] = real ? v.s_list[i][d] : 1.0;
      nu_out[i * NG + d] = real ? v.nu_list[i][d] : 0.0;
    }
    if(solver.traceDataList().size() > 0 && status != Solver::Status::Succeeded)
      for(int a = 0; a < NU; a++)
        for(int d = 0; d < NX; d++) K_out[(i * NX + d) * NU + a] = solver.coeffList()[i].K(a, d);
  }
  const auto & tl = solver.traceDataList();
  for(size_t r = 0; r < tl.size(); r++) kkt_out[r] = tl[r].kkt_error;
  *n_trace_out = (int)tl.size();
  *status_out = static_cast<int>(status);
  return 0;
}

} // namespace

extern "C"
{

int ref_fmpc_solve(const char * model,
                   const double * params,
                   const ref_fmpc_config * cfg,
                   double t0,
                   const double * x0,
                   const double * x_in,
                   const double * u_in,
                   const double * lambda_in,
                   const double * s_in,
                   const double * nu_in,
                   double * x_out,
                   double * u_out,
                   double * lambda_out,
                   double * s_out,
                   double * nu_out,
                   double * K_out,
                   double * kkt_out,
                   int * n_trace_out,
                   int * status_out)
{
  std::string m(model);
  if(m == "fmpc_cartpole")
    return fmpcSolve<FmpcProblemCartPole, 4, 1, 4>(params, cfg, t0, x0, x_in, u_in, lambda_in, s_in, nu_in, x_out,
                                                   u_out, lambda_out, s_out, nu_out, K_out, kkt_out, n_trace_out,
                                                   status_out);
  if(m == "fmpc_oscillator")
    return fmpcSolve<FmpcProblemOscillator, 2, 1, 3>(params, cfg, t0, x0, x_in, u_in, lambda_in, s_in, nu_in, x_out,
                                                     u_out, lambda_out, s_out, nu_out, K_out, kkt_out, n_trace_out,
                                                     status_out);
  if(m == "fmpc_planar_quadrotor")
    return fmpcSolve<FmpcProblemPlanarQuadrotor, 6, 2, 4>(params, cfg, t0, x0, x_in, u_in, lambda_in, s_in, nu_in, x_out,
                                                          u_out, lambda_out, s_out, nu_out, K_out, kkt_out, n_trace_out,
                                                          status_out);
  if(m == "fmpc_cartpole_windowed")
    return fmpcSolve<FmpcProblemCartPoleWindowed, 4, 1, 4>(params, cfg, t0, x0, x_in, u_in, lambda_in, s_in, nu_in, x_out,
                                                           u_out, lambda_out, s_out, nu_out, K_out, kkt_out, n_trace_out,
                                                           status_out);
  return -2;
}

} // extern "C"
