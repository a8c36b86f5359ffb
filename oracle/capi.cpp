// TEST INFRASTRUCTURE ONLY -- CPU oracle for nmpc_b200 (see oracle/README.md).
//
// C entry points (ctypes-friendly) over the oracle restatements, plus the OpenMP batch driver
// that bench.py times as the CPU baseline ("N independent DDPSolver objects", one per host
// thread -- the reference has no batching of its own).  Host layouts are instance-major:
//   x0[B][NX], u_init[B][N][NU], x[B][N+1][NX], u[B][N][NU], cost[B][N+1], k[B][N][NU],
//   K[B][N][NU*NX] (column-major NU x NX per step), trace[B][max_iter+1][9].
#include <cstring>
#include <memory>
#include <cmath>
#include <string>
#include <vector>

#ifdef _OPENMP
#  include <omp.h>
#endif

#include "ddp_oracle.hpp"
#include "fmpc_oracle.hpp"
#include "models.hpp"

extern "C"
{
typedef struct
{
  int horizon_steps;
  int max_iter;
  int reg_type;
  int with_input_constraint;
  int n_alpha;
  int reserved;
  double initial_lambda;
  double initial_dlambda;
  double lambda_factor;
  double lambda_min;
  double lambda_max;
  double k_rel_norm_thre;
  double lambda_thre;
  double cost_update_ratio_thre;
  double cost_update_thre;
  double alpha_list[16];
} oracle_ddp_config;

typedef struct
{
  int horizon_steps;
  int max_iter;
  int check_nan;
  int init_complementary_variable;
  int update_barrier_eps;
  int break_if_llt_fails;
  int enable_line_search;
  int merit_const_scale_from_lagrange_multipliers;
  double kkt_error_thre;
  double initial_barrier_eps; // value of the barrier_eps_ member when solve() is entered (FmpcSolver.h:413-414: 1e-4)
} oracle_fmpc_config;
}

namespace
{
using namespace oracle;

template<class Problem, int NX, int NU>
int ddpSolveBatch(const double * params,
                  const oracle_ddp_config * cfg,
                  int B,
                  double t0,
                  const double * x0,
                  const double * u_init,
                  const double * u_lo,
                  const double * u_hi,
                  double * x_out,
                  double * u_out,
                  double * cost_out,
                  double * k_out,
                  double * K_out,
                  double * trace_out,
                  int * n_trace_out,
                  int * status_out,
                  int * iters_out,
                  int * n_fwd_out,
                  int * n_bwd_out,
                  int nthreads,
                  const double * u_lo_steps = nullptr, // [N][NU]: limits that change along the horizon
                  const double * u_hi_steps = nullptr)
{
  const int N = cfg->horizon_steps;
  const int TR = cfg->max_iter + 1;
  int err = 0;
#ifdef _OPENMP
  if(nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
  {
    auto problem = std::make_shared<Problem>(params);
    DDPSolver<NX, NU> solver(problem);
    auto & c = solver.config();
    c.print_level = 0;
    c.with_input_constraint = cfg->with_input_constraint != 0;
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
    c.alpha_list.assign(cfg->alpha_list, cfg->alpha_list + cfg->n_alpha);
    c.cost_update_ratio_thre = cfg->cost_update_ratio_thre;
    c.cost_update_thre = cfg->cost_update_thre;
    if(c.with_input_constraint)
    {
      std::array<Vec<NU>, 2> lim;
      for(int d = 0; d < NU; d++)
      {
        lim[0][d] = u_lo[d];
        lim[1][d] = u_hi[d];
      }
      if(u_lo_steps != nullptr && u_hi_steps != nullptr)
      {
        // input_limits_func_(t) of the reference is a function of time: here a table over the horizon steps
        const double dt = problem->dt();
        solver.setInputLimitsFunc([=](double t) {
          long i = std::lround((t - t0) / dt);
          i = i < 0 ? 0 : (i > N - 1 ? N - 1 : i);
          std::array<Vec<NU>, 2> l;
          for(int d = 0; d < NU; d++)
          {
            l[0][d] = u_lo_steps[(size_t)i * NU + d];
            l[1][d] = u_hi_steps[(size_t)i * NU + d];
          }
          return l;
        });
      }
      else
      {
        solver.setInputLimitsFunc([lim](double) { return lim; });
      }
    }

    std::vector<Vec<NU>> u_list(N);
#pragma omp for schedule(dynamic, 4)
    for(int b = 0; b < B; b++)
    {
      Vec<NX> cx;
      for(int d = 0; d < NX; d++) cx[d] = x0[(size_t)b * NX + d];
      for(int i = 0; i < N; i++)
        for(int d = 0; d < NU; d++) u_list[i][d] = u_init[((size_t)b * N + i) * NU + d];
      try
      {
        solver.solve(t0, cx, u_list);
      }
      catch(...)
      {
#pragma omp atomic write
        err = -1;
        continue;
      }
      const auto & cd = solver.controlData();
      if(x_out)
        for(int i = 0; i <= N; i++)
          for(int d = 0; d < NX; d++) x_out[((size_t)b * (N + 1) + i) * NX + d] = cd.x_list[i][d];
      if(u_out)
        for(int i = 0; i < N; i++)
          for(int d = 0; d < NU; d++) u_out[((size_t)b * N + i) * NU + d] = cd.u_list[i][d];
      if(cost_out)
        for(int i = 0; i <= N; i++) cost_out[(size_t)b * (N + 1) + i] = cd.cost_list[i];
      if(k_out)
        for(int i = 0; i < N; i++)
          for(int d = 0; d < NU; d++) k_out[((size_t)b * N + i) * NU + d] = solver.kList()[i][d];
      if(K_out)
        for(int i = 0; i < N; i++)
          for(int d = 0; d < NU * NX; d++) K_out[((size_t)b * N + i) * NU * NX + d] = solver.KList()[i].d[d];
      const auto & tl = solver.traceDataList();
      if(trace_out)
      {
        double * tr = trace_out + (size_t)b * TR * 9;
        std::memset(tr, 0, sizeof(double) * TR * 9);
        for(size_t r = 0; r < tl.size() && (int)r < TR; r++)
        {
          tr[r * 9 + 0] = tl[r].iter;
          tr[r * 9 + 1] = tl[r].cost;
          tr[r * 9 + 2] = tl[r].lambda;
          tr[r * 9 + 3] = tl[r].dlambda;
          tr[r * 9 + 4] = tl[r].alpha;
          tr[r * 9 + 5] = tl[r].k_rel_norm;
          tr[r * 9 + 6] = tl[r].cost_update_actual;
          tr[r * 9 + 7] = tl[r].cost_update_expected;
          tr[r * 9 + 8] = tl[r].cost_update_ratio;
        }
      }
      if(n_trace_out) n_trace_out[b] = (int)tl.size();
      if(status_out) status_out[b] = solver.retval_last;
      if(iters_out) iters_out[b] = tl.back().iter;
      if(n_fwd_out) n_fwd_out[b] = solver.n_forward_pass;
      if(n_bwd_out) n_bwd_out[b] = solver.n_backward_pass;
    }
  }
  return err;
}

template<class Problem, int NX, int NU>
int modelEval(const double * params,
              double t,
              const double * x,
              const double * u,
              double * x_next,
              double * costs,
              double * Fx,
              double * Fu,
              double * Lx,
              double * Lu,
              double * Lxx,
              double * Luu,
              double * Lxu,
              double * Vx,
              double * Vxx)
{
  Problem problem(params);
  Vec<NX> xv;
  Vec<NU> uv;
  for(int d = 0; d < NX; d++) xv[d] = x[d];
  for(int d = 0; d < NU; d++) uv[d] = u[d];
  Vec<NX> xn = problem.stateEq(t, xv, uv);
  std::memcpy(x_next, xn.d, sizeof(double) * NX);
  costs[0] = problem.runningCost(t, xv, uv);
  costs[1] = problem.terminalCost(t, xv);
  Mat<NX, NX> mFx, mLxx, mVxx;
  Mat<NX, NU> mFu, mLxu;
  Vec<NX> mLx, mVx;
  Vec<NU> mLu;
  Mat<NU, NU> mLuu;
  problem.calcStateEqDeriv(t, xv, uv, mFx, mFu);
  problem.calcRunningCostDeriv(t, xv, uv, mLx, mLu, mLxx, mLuu, mLxu);
  problem.calcTerminalCostDeriv(t, xv, mVx, mVxx);
  std::memcpy(Fx, mFx.d, sizeof(double) * NX * NX);
  std::memcpy(Fu, mFu.d, sizeof(double) * NX * NU);
  std::memcpy(Lx, mLx.d, sizeof(double) * NX);
  std::memcpy(Lu, mLu.d, sizeof(double) * NU);
  std::memcpy(Lxx, mLxx.d, sizeof(double) * NX * NX);
  std::memcpy(Luu, mLuu.d, sizeof(double) * NU * NU);
  std::memcpy(Lxu, mLxu.d, sizeof(double) * NX * NU);
  std::memcpy(Vx, mVx.d, sizeof(double) * NX);
  std::memcpy(Vxx, mVxx.d, sizeof(double) * NX * NX);
  return 0;
}

template<int N>
int boxqpSolve(const double * H,
               const double * g,
               const double * lower,
               const double * upper,
               const double * x0,
               double * x_out,
               int * retval,
               int * iters)
{
  BoxQP<N> qp;
  Mat<N, N> mH;
  Vec<N> mg, ml, mu, mx0;
  std::memcpy(mH.d, H, sizeof(double) * N * N);
  std::memcpy(mg.d, g, sizeof(double) * N);
  std::memcpy(ml.d, lower, sizeof(double) * N);
  std::memcpy(mu.d, upper, sizeof(double) * N);
  std::memcpy(mx0.d, x0, sizeof(double) * N);
  Vec<N> x = qp.solve(mH, mg, ml, mu, mx0);
  std::memcpy(x_out, x.d, sizeof(double) * N);
  *retval = qp.retval;
  *iters = qp.iter;
  return 0;
}
} // namespace

extern "C"
{
void oracle_ddp_config_default(oracle_ddp_config * cfg)
{
  oracle::DDPSolver<1, 1>::Configuration c;
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->horizon_steps = c.horizon_steps;
  cfg->max_iter = c.max_iter;
  cfg->reg_type = c.reg_type;
  cfg->with_input_constraint = c.with_input_constraint ? 1 : 0;
  cfg->n_alpha = (int)c.alpha_list.size();
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

int oracle_model_dims(const char * model, int * nx, int * nu, int * ng, int * nparams)
{
  std::string m(model);
  if(m == "cartpole")
  {
    *nx = 4, *nu = 1, *ng = 0, *nparams = DDPProblemCartPole::kNumParams;
    return 0;
  }
  if(m == "bipedal")
  {
    *nx = 2, *nu = 1, *ng = 0, *nparams = DDPProblemBipedal::kNumParams;
    return 0;
  }
  if(m == "quadrotor")
  {
    *nx = 12, *nu = 4, *ng = 0, *nparams = DDPProblemQuadrotor::kNumParams;
    return 0;
  }
  if(m == "vertical_motion")
  {
    *nx = 2, *nu = 2, *ng = 0, *nparams = DDPProblemVerticalMotion::kNumParams;
    return 0;
  }
  if(m == "centroidal_motion")
  {
    *nx = 9, *nu = 16, *ng = 0, *nparams = DDPProblemCentroidalMotion::kNumParams;
    return 0;
  }
  if(m == "planar_quadrotor")
  {
    *nx = 6, *nu = 2, *ng = 0, *nparams = DDPProblemPlanarQuadrotor::kNumParams;
    return 0;
  }
  if(m == "fmpc_cartpole")
  {
    *nx = 4, *nu = 1, *ng = 4, *nparams = FmpcProblemCartPole::kNumParams;
    return 0;
  }
  if(m == "fmpc_oscillator")
  {
    *nx = 2, *nu = 1, *ng = 3, *nparams = FmpcProblemOscillator::kNumParams;
    return 0;
  }
  if(m == "fmpc_planar_quadrotor")
  {
    *nx = 6, *nu = 2, *ng = 4, *nparams = FmpcProblemPlanarQuadrotor::kNumParams;
    return 0;
  }
  if(m == "fmpc_cartpole_windowed")
  {
    *nx = 4, *nu = 1, *ng = 4, *nparams = FmpcProblemCartPoleWindowed::kNumParams;
    return 0;
  }
  return -1;
}

int oracle_model_default_params(const char * model, double * params)
{
  std::string m(model);
  if(m == "cartpole")
    DDPProblemCartPole::defaultParams(params);
  else if(m == "bipedal")
    DDPProblemBipedal::defaultParams(params);
  else if(m == "quadrotor")
    DDPProblemQuadrotor::defaultParams(params);
  else if(m == "vertical_motion")
    DDPProblemVerticalMotion::defaultParams(params);
  else if(m == "centroidal_motion")
    DDPProblemCentroidalMotion::defaultParams(params);
  else if(m == "planar_quadrotor")
    DDPProblemPlanarQuadrotor::defaultParams(params);
  else if(m == "fmpc_cartpole")
    FmpcProblemCartPole::defaultParams(params);
  else if(m == "fmpc_oscillator")
    FmpcProblemOscillator::defaultParams(params);
  else if(m == "fmpc_planar_quadrotor")
    FmpcProblemPlanarQuadrotor::defaultParams(params);
  else if(m == "fmpc_cartpole_windowed")
    FmpcProblemCartPoleWindowed::defaultParams(params);
  else
    return -1;
  return 0;
}

int oracle_ddp_solve_batch(const char * model,
                           const double * params,
                           const oracle_ddp_config * cfg,
                           int B,
                           double t0,
                           const double * x0,
                           const double * u_init,
                           const double * u_lo,
                           const double * u_hi,
                           double * x_out,
                           double * u_out,
                           double * cost_out,
                           double * k_out,
                           double * K_out,
                           double * trace_out,
                           int * n_trace_out,
                           int * status_out,
                           int * iters_out,
                           int * n_fwd_out,
                           int * n_bwd_out,
                           int nthreads)
{
  std::string m(model);
  if(m == "cartpole")
    return ddpSolveBatch<DDPProblemCartPole, 4, 1>(params, cfg, B, t0, x0, u_init, u_lo, u_hi, x_out, u_out, cost_out,
                                                   k_out, K_out, trace_out, n_trace_out, status_out, iters_out,
                                                   n_fwd_out, n_bwd_out, nthreads);
  if(m == "bipedal")
    return ddpSolveBatch<DDPProblemBipedal, 2, 1>(params, cfg, B, t0, x0, u_init, u_lo, u_hi, x_out, u_out, cost_out,
                                                  k_out, K_out, trace_out, n_trace_out, status_out, iters_out,
                                                  n_fwd_out, n_bwd_out, nthreads);
  if(m == "quadrotor")
    return ddpSolveBatch<DDPProblemQuadrotor, 12, 4>(params, cfg, B, t0, x0, u_init, u_lo, u_hi, x_out, u_out,
                                                     cost_out, k_out, K_out, trace_out, n_trace_out, status_out,
                                                     iters_out, n_fwd_out, n_bwd_out, nthreads);
  if(m == "vertical_motion")
    return ddpSolveBatch<DDPProblemVerticalMotion, 2, 2>(params, cfg, B, t0, x0, u_init, u_lo, u_hi, x_out, u_out,
                                                         cost_out, k_out, K_out, trace_out, n_trace_out, status_out,
                                                         iters_out, n_fwd_out, n_bwd_out, nthreads);
  if(m == "centroidal_motion")
    return ddpSolveBatch<DDPProblemCentroidalMotion, 9, 16>(params, cfg, B, t0, x0, u_init, u_lo, u_hi, x_out, u_out,
                                                            cost_out, k_out, K_out, trace_out, n_trace_out, status_out,
                                                            iters_out, n_fwd_out, n_bwd_out, nthreads);
  if(m == "planar_quadrotor")
    return ddpSolveBatch<DDPProblemPlanarQuadrotor, 6, 2>(params, cfg, B, t0, x0, u_init, u_lo, u_hi, x_out, u_out,
                                                          cost_out, k_out, K_out, trace_out, n_trace_out, status_out,
                                                          iters_out, n_fwd_out, n_bwd_out, nthreads);
  return -2;
}

/** oracle_ddp_solve_batch for the cart-pole with input limits that change along the horizon (u_lo_steps /
    u_hi_steps [N][1]): the reference's input_limits_func_(t) evaluated at t_i (DDPSolver.hpp:470). */
int oracle_ddp_solve_batch_cartpole_tv(const double * params,
                                       const oracle_ddp_config * cfg,
                                       int B,
                                       double t0,
                                       const double * x0,
                                       const double * u_init,
                                       const double * u_lo_steps,
                                       const double * u_hi_steps,
                                       double * x_out,
                                       double * u_out,
                                       double * cost_out,
                                       int * status_out,
                                       int * iters_out,
                                       int nthreads)
{
  const double zero = 0.0;
  return ddpSolveBatch<DDPProblemCartPole, 4, 1>(params, cfg, B, t0, x0, u_init, &zero, &zero, x_out, u_out, cost_out,
                                                 nullptr, nullptr, nullptr, nullptr, status_out, iters_out, nullptr,
                                                 nullptr, nthreads, u_lo_steps, u_hi_steps);
}

int oracle_model_eval(const char * model,
                      const double * params,
                      double t,
                      const double * x,
                      const double * u,
                      double * x_next,
                      double * costs,
                      double * Fx,
                      double * Fu,
                      double * Lx,
                      double * Lu,
                      double * Lxx,
                      double * Luu,
                      double * Lxu,
                      double * Vx,
                      double * Vxx)
{
  std::string m(model);
  if(m == "cartpole")
    return modelEval<DDPProblemCartPole, 4, 1>(params, t, x, u, x_next, costs, Fx, Fu, Lx, Lu, Lxx, Luu, Lxu, Vx, Vxx);
  if(m == "bipedal")
    return modelEval<DDPProblemBipedal, 2, 1>(params, t, x, u, x_next, costs, Fx, Fu, Lx, Lu, Lxx, Luu, Lxu, Vx, Vxx);
  if(m == "quadrotor")
    return modelEval<DDPProblemQuadrotor, 12, 4>(params, t, x, u, x_next, costs, Fx, Fu, Lx, Lu, Lxx, Luu, Lxu, Vx, Vxx);
  if(m == "vertical_motion")
    return modelEval<DDPProblemVerticalMotion, 2, 2>(params, t, x, u, x_next, costs, Fx, Fu, Lx, Lu, Lxx, Luu, Lxu, Vx,
                                                     Vxx);
  if(m == "centroidal_motion")
    return modelEval<DDPProblemCentroidalMotion, 9, 16>(params, t, x, u, x_next, costs, Fx, Fu, Lx, Lu, Lxx, Luu, Lxu, Vx,
                                                        Vxx);
  if(m == "planar_quadrotor")
    return modelEval<DDPProblemPlanarQuadrotor, 6, 2>(params, t, x, u, x_next, costs, Fx, Fu, Lx, Lu, Lxx, Luu, Lxu, Vx,
                                                      Vxx);
  if(m == "fmpc_cartpole")
    return modelEval<FmpcProblemCartPole, 4, 1>(params, t, x, u, x_next, costs, Fx, Fu, Lx, Lu, Lxx, Luu, Lxu, Vx, Vxx);
  if(m == "fmpc_oscillator")
    return modelEval<FmpcProblemOscillator, 2, 1>(params, t, x, u, x_next, costs, Fx, Fu, Lx, Lu, Lxx, Luu, Lxu, Vx,
                                                  Vxx);
  if(m == "fmpc_planar_quadrotor")
    return modelEval<FmpcProblemPlanarQuadrotor, 6, 2>(params, t, x, u, x_next, costs, Fx, Fu, Lx, Lu, Lxx, Luu, Lxu, Vx,
                                                       Vxx);
  if(m == "fmpc_cartpole_windowed")
    return modelEval<FmpcProblemCartPoleWindowed, 4, 1>(params, t, x, u, x_next, costs, Fx, Fu, Lx, Lu, Lxx, Luu, Lxu, Vx,
                                                        Vxx);
  return -2;
}

int oracle_boxqp_solve(int n,
                       const double * H,
                       const double * g,
                       const double * lower,
                       const double * upper,
                       const double * x0,
                       double * x_out,
                       int * retval,
                       int * iters)
{
  switch(n)
  {
    case 1:
      return boxqpSolve<1>(H, g, lower, upper, x0, x_out, retval, iters);
    case 2:
      return boxqpSolve<2>(H, g, lower, upper, x0, x_out, retval, iters);
    case 3:
      return boxqpSolve<3>(H, g, lower, upper, x0, x_out, retval, iters);
    case 4:
      return boxqpSolve<4>(H, g, lower, upper, x0, x_out, retval, iters);
    default:
      return -2;
  }
}

int oracle_num_threads()
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
} // extern "C"

#include "fmpc_capi.inc"
