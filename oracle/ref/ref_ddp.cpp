// TEST INFRASTRUCTURE ONLY -- builds oracle/_ref/libnmpc_ref.so: the REFERENCE's own solver headers
// (straight from /root/reference, unmodified) compiled against oracle/ref/eigen_shim, with the problem
// classes of the reference's tests restated in the tests' own Eigen idioms (ref_models.h).  Used to pin
// oracle/ (the plain restatement) against the reference's real control flow, and to generate
// tests/golden/*.npz.  DDPSolver.hpp and FmpcSolver.hpp each define calcDuration() in an anonymous
// namespace, so the two solvers live in separate translation units.
#include <array>
#include <cstring>
#include <functional>
#include <iostream>
#include <memory>
#include <string>
#include <vector>
#ifdef _OPENMP
#  include <omp.h>
#endif

#include <nmpc_ddp/BoxQP.h>
#include <nmpc_ddp/DDPSolver.h>

#include "ref_models.h"

#include <nmpc_b200/models/quadrotor.h>

namespace
{
/** A functor of include/nmpc_b200/models (the same source the CUDA kernels and oracle/ evaluate) behind the REFERENCE's
    DDPProblem interface, so that the reference's DDPSolver runs on models that none of its own tests define. */
template<class F>
class RefProblemFromFunctor : public nmpc_ddp::DDPProblem<F::NX, F::NU>
{
public:
  using Base = nmpc_ddp::DDPProblem<F::NX, F::NU>;
  using StateDimVector = typename Base::StateDimVector;
  using InputDimVector = typename Base::InputDimVector;
  using StateStateDimMatrix = typename Base::StateStateDimMatrix;
  using InputInputDimMatrix = typename Base::InputInputDimMatrix;
  using StateInputDimMatrix = typename Base::StateInputDimMatrix;
  static constexpr int NX = F::NX, NU = F::NU;

  explicit RefProblemFromFunctor(const double * p) : Base(p[0]), f_(F::fromParams(p)) {}

  static typename F::StateDimVector fx(const StateDimVector & x)
  {
    typename F::StateDimVector v;
    for(int i = 0; i < NX; i++) v[i] = x[i];
    return v;
  }
  static typename F::InputDimVector fu(const InputDimVector & u)
  {
    typename F::InputDimVector v;
    for(int i = 0; i < NU; i++) v[i] = u[i];
    return v;
  }
  StateDimVector stateEq(double t, const StateDimVector & x, const InputDimVector & u) const override
  {
    const typename F::StateDimVector n = f_.stateEq(t, fx(x), fu(u));
    StateDimVector out;
    for(int i = 0; i < NX; i++) out[i] = n[i];
    return out;
  }
  double runningCost(double t, const StateDimVector & x, const InputDimVector & u) const override
  {
    return f_.runningCost(t, fx(x), fu(u));
  }
  double terminalCost(double t, const StateDimVector & x) const override
  {
    return f_.terminalCost(t, fx(x));
  }
  void calcStateEqDeriv(double t,
                        const StateDimVector & x,
                        const InputDimVector & u,
                        Eigen::Ref<StateStateDimMatrix> state_eq_deriv_x,
                        Eigen::Ref<StateInputDimMatrix> state_eq_deriv_u) const override
  {
    typename F::StateStateDimMatrix a;
    typename F::StateInputDimMatrix b;
    f_.calcStateEqDeriv(t, fx(x), fu(u), a, b);
    for(int j = 0; j < NX; j++)
      for(int i = 0; i < NX; i++) state_eq_deriv_x(i, j) = a(i, j);
    for(int j = 0; j < NU; j++)
      for(int i = 0; i < NX; i++) state_eq_deriv_u(i, j) = b(i, j);
  }
  void calcStateEqDeriv(double,
                        const StateDimVector &,
                        const InputDimVector &,
                        Eigen::Ref<StateStateDimMatrix>,
                        Eigen::Ref<StateInputDimMatrix>,
                        std::vector<StateStateDimMatrix> &,
                        std::vector<InputInputDimMatrix> &,
                        std::vector<StateInputDimMatrix> &) const override
  {
    throw std::runtime_error("Second-order derivatives of state equation are not implemented.");
  }
  void calcRunningCostDeriv(double t,
                            const StateDimVector & x,
                            const InputDimVector & u,
                            Eigen::Ref<StateDimVector> running_cost_deriv_x,
                            Eigen::Ref<InputDimVector> running_cost_deriv_u) const override
  {
    // the functors have the second-order overload only
    typename F::StateDimVector lx;
    typename F::InputDimVector lu;
    typename F::StateStateDimMatrix lxx;
    typename F::InputInputDimMatrix luu;
    typename F::StateInputDimMatrix lxu;
    f_.calcRunningCostDeriv(t, fx(x), fu(u), lx, lu, lxx, luu, lxu);
    for(int i = 0; i < NX; i++) running_cost_deriv_x[i] = lx[i];
    for(int i = 0; i < NU; i++) running_cost_deriv_u[i] = lu[i];
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
    typename F::StateDimVector lx;
    typename F::InputDimVector lu;
    typename F::StateStateDimMatrix lxx;
    typename F::InputInputDimMatrix luu;
    typename F::StateInputDimMatrix lxu;
    f_.calcRunningCostDeriv(t, fx(x), fu(u), lx, lu, lxx, luu, lxu);
    for(int i = 0; i < NX; i++) running_cost_deriv_x[i] = lx[i];
    for(int i = 0; i < NU; i++) running_cost_deriv_u[i] = lu[i];
    for(int j = 0; j < NX; j++)
      for(int i = 0; i < NX; i++) running_cost_deriv_xx(i, j) = lxx(i, j);
    for(int j = 0; j < NU; j++)
      for(int i = 0; i < NU; i++) running_cost_deriv_uu(i, j) = luu(i, j);
    for(int j = 0; j < NU; j++)
      for(int i = 0; i < NX; i++) running_cost_deriv_xu(i, j) = lxu(i, j);
  }
  void calcTerminalCostDeriv(double t, const StateDimVector & x, Eigen::Ref<StateDimVector> terminal_cost_deriv_x)
      const override
  {
    typename F::StateDimVector vx;
    typename F::StateStateDimMatrix vxx;
    f_.calcTerminalCostDeriv(t, fx(x), vx, vxx);
    for(int i = 0; i < NX; i++) terminal_cost_deriv_x[i] = vx[i];
  }
  void calcTerminalCostDeriv(double t,
                             const StateDimVector & x,
                             Eigen::Ref<StateDimVector> terminal_cost_deriv_x,
                             Eigen::Ref<StateStateDimMatrix> terminal_cost_deriv_xx) const override
  {
    typename F::StateDimVector vx;
    typename F::StateStateDimMatrix vxx;
    f_.calcTerminalCostDeriv(t, fx(x), vx, vxx);
    for(int i = 0; i < NX; i++) terminal_cost_deriv_x[i] = vx[i];
    for(int j = 0; j < NX; j++)
      for(int i = 0; i < NX; i++) terminal_cost_deriv_xx(i, j) = vxx(i, j);
  }

protected:
  F f_;
};
// four inputs, twelve states: BoxQP<4> with several free / clamped inputs per step
using DDPProblemQuadrotor = RefProblemFromFunctor<nmpc_b200::models::Quadrotor<double>>;
} // namespace

namespace
{
using DDPProblemCartPole = CartPoleBodies<nmpc_ddp::DDPProblem<4, 1>>;
// two inputs: the control-limited backward pass with BoxQP<2> (DDPSolver.hpp:450-497)
using DDPProblemPlanarQuadrotor = PlanarQuadrotorBodies<nmpc_ddp::DDPProblem<6, 2>>;
}

extern "C"
{
// same struct layouts as oracle/capi.cpp
typedef struct
{
  int horizon_steps, max_iter, reg_type, with_input_constraint, n_alpha, reserved;
  double initial_lambda, initial_dlambda, lambda_factor, lambda_min, lambda_max, k_rel_norm_thre, lambda_thre,
      cost_update_ratio_thre, cost_update_thre;
  double alpha_list[16];
} ref_ddp_config;

/** The reference's default Configuration, as its own constructor builds it (DDPSolver.h:47-110). */
void ref_ddp_config_default(ref_ddp_config * cfg)
{
  nmpc_ddp::DDPSolver<4, 1>::Configuration c;
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->horizon_steps = c.horizon_steps;
  cfg->max_iter = c.max_iter;
  cfg->reg_type = c.reg_type;
  cfg->with_input_constraint = c.with_input_constraint;
  cfg->n_alpha = c.alpha_list.size();
  cfg->initial_lambda = c.initial_lambda;
  cfg->initial_dlambda = c.initial_dlambda;
  cfg->lambda_factor = c.lambda_factor;
  cfg->lambda_min = c.lambda_min;
  cfg->lambda_max = c.lambda_max;
  cfg->k_rel_norm_thre = c.k_rel_norm_thre;
  cfg->lambda_thre = c.lambda_thre;
  cfg->cost_update_ratio_thre = c.cost_update_ratio_thre;
  cfg->cost_update_thre = c.cost_update_thre;
  for(int i = 0; i < cfg->n_alpha; i++) cfg->alpha_list[i] = c.alpha_list[i];
}

} // extern "C"

namespace
{
/** One DDPSolver<NX, NU>::solve() with the reference's code.  Arrays as in oracle_ddp_solve_batch, B = 1;
    u_lo / u_hi [NU] constant limits. */
template<class Problem, int NX, int NU>
int refDdpSolve(const double * params,
                           const ref_ddp_config * cfg,
                           double t0,
                           const double * x0,
                           const double * u_init,
                           const double * u_lo,
                           const double * u_hi,
                           double * x_out,
                           double * u_out,
                           double * cost_out,
                           double * trace_out,
                           int * n_trace_out,
                           int * solve_ret)
{
  auto problem = std::make_shared<Problem>(params);
  nmpc_ddp::DDPSolver<NX, NU> solver(problem);
  auto & c = solver.config();
  c.print_level = 0;
  c.with_input_constraint = cfg->with_input_constraint != 0;
  c.max_iter = cfg->max_iter;
  c.horizon_steps = cfg->horizon_steps;
  c.reg_type = cfg->reg_type;
  c.initial_lambda = cfg->initial_lambda;
  c.initial_dlambda = cfg->initial_dlambda;
  c.lambda_factor = cfg->lambda_factor;
  c.lambda_min = cfg->lambda_min;
  c.lambda_max = cfg->lambda_max;
  c.k_rel_norm_thre = cfg->k_rel_norm_thre;
  c.lambda_thre = cfg->lambda_thre;
  c.alpha_list.resize(cfg->n_alpha);
  for(int i = 0; i < cfg->n_alpha; i++) c.alpha_list[i] = cfg->alpha_list[i];
  c.cost_update_ratio_thre = cfg->cost_update_ratio_thre;
  c.cost_update_thre = cfg->cost_update_thre;
  if(c.with_input_constraint)
  {
    std::array<typename Problem::InputDimVector, 2> limits;
    for(int d = 0; d < NU; d++)
    {
      limits[0][d] = u_lo[d];
      limits[1][d] = u_hi[d];
    }
    solver.setInputLimitsFunc([limits](double) { return limits; });
  }
  const int N = cfg->horizon_steps;
  typename Problem::StateDimVector current_x;
  for(int d = 0; d < NX; d++) current_x[d] = x0[d];
  std::vector<typename Problem::InputDimVector> initial_u_list(N);
  for(int i = 0; i < N; i++)
    for(int d = 0; d < NU; d++) initial_u_list[i][d] = u_init[i * NU + d];
  // the unconstrained branch prints "[DDP/Forward] Value is not expected to decrease." even at print_level 0
  std::streambuf * old = std::cout.rdbuf(nullptr);
  bool ret = false;
  try
  {
    ret = solver.solve(t0, current_x, initial_u_list);
  }
  catch(...)
  {
    std::cout.rdbuf(old);
    return -1;
  }
  std::cout.rdbuf(old);
  *solve_ret = ret ? 1 : 0;
  const auto & cd = solver.controlData();
  for(int i = 0; i <= N; i++)
  {
    for(int d = 0; d < NX; d++) x_out[i * NX + d] = cd.x_list[i][d];
    cost_out[i] = cd.cost_list[i];
  }
  for(int i = 0; i < N; i++)
    for(int d = 0; d < NU; d++) u_out[i * NU + d] = cd.u_list[i][d];
  const auto & tl = solver.traceDataList();
  std::memset(trace_out, 0, sizeof(double) * (cfg->max_iter + 1) * 9);
  for(size_t r = 0; r < tl.size(); r++)
  {
    double * tr = trace_out + r * 9;
    tr[0] = tl[r].iter, tr[1] = tl[r].cost, tr[2] = tl[r].lambda, tr[3] = tl[r].dlambda, tr[4] = tl[r].alpha;
    tr[5] = tl[r].k_rel_norm, tr[6] = tl[r].cost_update_actual, tr[7] = tl[r].cost_update_expected;
    tr[8] = tl[r].cost_update_ratio;
  }
  *n_trace_out = (int)tl.size();
  return 0;
}
} // namespace

extern "C"
{
int ref_ddp_solve_cartpole(const double * params,
                           const ref_ddp_config * cfg,
                           double t0,
                           const double * x0,
                           const double * u_init,
                           const double * u_lo,
                           const double * u_hi,
                           double * x_out,
                           double * u_out,
                           double * cost_out,
                           double * trace_out,
                           int * n_trace_out,
                           int * solve_ret)
{
  return refDdpSolve<DDPProblemCartPole, 4, 1>(params, cfg, t0, x0, u_init, u_lo, u_hi, x_out, u_out, cost_out, trace_out,
                                               n_trace_out, solve_ret);
}

/** The 3-D quadrotor functor (n_x = 12, n_u = 4) through the reference's DDPSolver<12, 4>, with BoxQP<4>. */
int ref_ddp_solve_quadrotor(const double * params,
                            const ref_ddp_config * cfg,
                            double t0,
                            const double * x0,
                            const double * u_init,
                            const double * u_lo,
                            const double * u_hi,
                            double * x_out,
                            double * u_out,
                            double * cost_out,
                            double * trace_out,
                            int * n_trace_out,
                            int * solve_ret)
{
  return refDdpSolve<DDPProblemQuadrotor, 12, 4>(params, cfg, t0, x0, u_init, u_lo, u_hi, x_out, u_out, cost_out,
                                                 trace_out, n_trace_out, solve_ret);
}

/** The planar quadrotor (n_x = 6, n_u = 2) through the reference's DDPSolver<6, 2>, with BoxQP<2> when limits are on. */
int ref_ddp_solve_planar(const double * params,
                         const ref_ddp_config * cfg,
                         double t0,
                         const double * x0,
                         const double * u_init,
                         const double * u_lo,
                         const double * u_hi,
                         double * x_out,
                         double * u_out,
                         double * cost_out,
                         double * trace_out,
                         int * n_trace_out,
                         int * solve_ret)
{
  return refDdpSolve<DDPProblemPlanarQuadrotor, 6, 2>(params, cfg, t0, x0, u_init, u_lo, u_hi, x_out, u_out, cost_out,
                                                      trace_out, n_trace_out, solve_ret);
}

/** A batch of cart-pole solves with the reference's code: one DDPSolver object per OpenMP thread, instances
    handed out dynamically ("N independent DDPSolver instances").  Outputs: first-step control, total cost,
    iteration count per instance.  The timed CPU arm of bench.py (informational: the Eigen stand-in is
    heap-backed, so this under-states what the reference achieves with real Eigen). */
int ref_ddp_solve_cartpole_batch(const double * params,
                                 const ref_ddp_config * cfg,
                                 int B,
                                 double t0,
                                 const double * x0,
                                 const double * u_init,
                                 int nthreads,
                                 double * u0_out,
                                 double * cost_out,
                                 int * iters_out)
{
  const int N = cfg->horizon_steps;
  int err = 0;
  std::streambuf * old = std::cout.rdbuf(nullptr);
#ifdef _OPENMP
  if(nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
  {
    auto problem = std::make_shared<DDPProblemCartPole>(params);
    nmpc_ddp::DDPSolver<4, 1> solver(problem);
    auto & c = solver.config();
    c.print_level = 0;
    c.with_input_constraint = false;
    c.max_iter = cfg->max_iter;
    c.horizon_steps = N;
    c.reg_type = cfg->reg_type;
    c.initial_lambda = cfg->initial_lambda;
    c.initial_dlambda = cfg->initial_dlambda;
    c.lambda_factor = cfg->lambda_factor;
    c.lambda_min = cfg->lambda_min;
    c.lambda_max = cfg->lambda_max;
    c.k_rel_norm_thre = cfg->k_rel_norm_thre;
    c.lambda_thre = cfg->lambda_thre;
    c.alpha_list.resize(cfg->n_alpha);
    for(int i = 0; i < cfg->n_alpha; i++) c.alpha_list[i] = cfg->alpha_list[i];
    c.cost_update_ratio_thre = cfg->cost_update_ratio_thre;
    c.cost_update_thre = cfg->cost_update_thre;
    std::vector<DDPProblemCartPole::InputDimVector> initial_u_list(N);
#pragma omp for schedule(dynamic, 4)
    for(int b = 0; b < B; b++)
    {
      DDPProblemCartPole::StateDimVector current_x;
      current_x << x0[4 * b + 0], x0[4 * b + 1], x0[4 * b + 2], x0[4 * b + 3];
      for(int i = 0; i < N; i++) initial_u_list[i][0] = u_init[(size_t)b * N + i];
      try
      {
        solver.solve(t0, current_x, initial_u_list);
      }
      catch(...)
      {
#pragma omp atomic write
        err = -1;
        continue;
      }
      const auto & cd = solver.controlData();
      if(u0_out) u0_out[b] = cd.u_list[0][0];
      if(cost_out) cost_out[b] = cd.cost_list.sum();
      if(iters_out) iters_out[b] = solver.traceDataList().back().iter;
    }
  }
  std::cout.rdbuf(old);
  return err;
}

/** nmpc_ddp::BoxQP<2>::solve / BoxQP<Dynamic>::solve with the reference's code (TestBoxQP.cpp cases). */
int ref_boxqp_solve2(const double * H, const double * g, const double * lower, const double * upper, int dynamic,
                     double * x_out, int * retval)
{
  std::streambuf * old = std::cout.rdbuf(nullptr);
  if(dynamic)
  {
    nmpc_ddp::BoxQP<Eigen::Dynamic> qp(2);
    Eigen::MatrixXd Hm(2, 2);
    Eigen::VectorXd gv(2), lo(2), up(2);
    for(int j = 0; j < 2; j++)
      for(int i = 0; i < 2; i++) Hm(i, j) = H[i + 2 * j];
    for(int i = 0; i < 2; i++) gv[i] = g[i], lo[i] = lower[i], up[i] = upper[i];
    Eigen::VectorXd x = qp.solve(Hm, gv, lo, up);
    x_out[0] = x[0], x_out[1] = x[1];
    *retval = qp.retval_;
  }
  else
  {
    nmpc_ddp::BoxQP<2> qp;
    Eigen::Matrix2d Hm;
    Eigen::Vector2d gv, lo, up;
    for(int j = 0; j < 2; j++)
      for(int i = 0; i < 2; i++) Hm(i, j) = H[i + 2 * j];
    for(int i = 0; i < 2; i++) gv[i] = g[i], lo[i] = lower[i], up[i] = upper[i];
    Eigen::Vector2d x = qp.solve(Hm, gv, lo, up);
    x_out[0] = x[0], x_out[1] = x[1];
    *retval = qp.retval_;
  }
  std::cout.rdbuf(old);
  return 0;
}
} // extern "C"
