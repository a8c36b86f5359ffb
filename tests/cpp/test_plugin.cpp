// TEST -- a user functor that lives OUTSIDE libnmpc_b200.so (tests/plugin/): loaded with nmpc_b200_load_plugin, driven
// through the C++ facade exactly like a built-in problem, and checked against the oracle (oracle/ddp_oracle.hpp, test
// infrastructure) instantiated on the SAME functor.  g++ only: no CUDA code in this translation unit.
//   usage: test_plugin <libpendulum_plugin.so> [--no-solve]
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <random>
#include <vector>

#include <nmpc_ddp/DDPSolver.h>

#include "../../oracle/ddp_oracle.hpp"
#include "../plugin/pendulum.h"

namespace
{
/** The oracle's view of the same functor (cf. oracle::DDPProblemFromFunctor). */
class OraclePendulum : public oracle::DDPProblem<2, 1>
{
public:
  explicit OraclePendulum(const double * p) : oracle::DDPProblem<2, 1>(p[0]), f_(Pendulum<double>::fromParams(p)) {}
  template<class A, class B>
  static void copy(const A & a, B & b, int n)
  {
    for(int i = 0; i < n; i++) b.d[i] = a.d[i];
  }
  StateDimVector stateEq(double t, const StateDimVector & x, const InputDimVector & u) const override
  {
    Pendulum<double>::StateDimVector fx, fn;
    Pendulum<double>::InputDimVector fu;
    copy(x, fx, 2);
    copy(u, fu, 1);
    fn = f_.stateEq(t, fx, fu);
    StateDimVector out;
    copy(fn, out, 2);
    return out;
  }
  double runningCost(double t, const StateDimVector & x, const InputDimVector & u) const override
  {
    Pendulum<double>::StateDimVector fx;
    Pendulum<double>::InputDimVector fu;
    copy(x, fx, 2);
    copy(u, fu, 1);
    return f_.runningCost(t, fx, fu);
  }
  double terminalCost(double t, const StateDimVector & x) const override
  {
    Pendulum<double>::StateDimVector fx;
    copy(x, fx, 2);
    return f_.terminalCost(t, fx);
  }
  void calcStateEqDeriv(double t, const StateDimVector & x, const InputDimVector & u, StateStateDimMatrix & Fx,
                        StateInputDimMatrix & Fu) const override
  {
    Pendulum<double>::StateDimVector fx;
    Pendulum<double>::InputDimVector fu;
    Pendulum<double>::StateStateDimMatrix a;
    Pendulum<double>::StateInputDimMatrix b;
    copy(x, fx, 2);
    copy(u, fu, 1);
    f_.calcStateEqDeriv(t, fx, fu, a, b);
    copy(a, Fx, 4);
    copy(b, Fu, 2);
  }
  void calcRunningCostDeriv(double t, const StateDimVector & x, const InputDimVector & u, StateDimVector & Lx,
                            InputDimVector & Lu, StateStateDimMatrix & Lxx, InputInputDimMatrix & Luu,
                            StateInputDimMatrix & Lxu) const override
  {
    Pendulum<double>::StateDimVector fx, lx;
    Pendulum<double>::InputDimVector fu, lu;
    Pendulum<double>::StateStateDimMatrix lxx;
    Pendulum<double>::InputInputDimMatrix luu;
    Pendulum<double>::StateInputDimMatrix lxu;
    copy(x, fx, 2);
    copy(u, fu, 1);
    f_.calcRunningCostDeriv(t, fx, fu, lx, lu, lxx, luu, lxu);
    copy(lx, Lx, 2);
    copy(lu, Lu, 1);
    copy(lxx, Lxx, 4);
    copy(luu, Luu, 1);
    copy(lxu, Lxu, 2);
  }
  void calcTerminalCostDeriv(double t, const StateDimVector & x, StateDimVector & Vx, StateStateDimMatrix & Vxx)
      const override
  {
    Pendulum<double>::StateDimVector fx, vx;
    Pendulum<double>::StateStateDimMatrix vxx;
    copy(x, fx, 2);
    f_.calcTerminalCostDeriv(t, fx, vx, vxx);
    copy(vx, Vx, 2);
    copy(vxx, Vxx, 4);
  }

protected:
  Pendulum<double> f_;
};
} // namespace

int main(int argc, char ** argv)
{
  if(argc < 2)
  {
    std::printf("usage: test_plugin <plugin.so> [--no-solve]\n");
    return 2;
  }
  const bool no_solve = argc > 2 && std::strcmp(argv[2], "--no-solve") == 0;

  // before loading: the library does not know the problem
  int nx = 0, nu = 0, ng = 0, np = 0;
  std::printf("known_before %d\n", nmpc_b200_model_dims("pendulum", &nx, &nu, &ng, &np) == NMPC_B200_OK ? 1 : 0);
  std::printf("load_missing %d\n", nmpc_b200_load_plugin("/nonexistent/libnope.so"));
  const int rc = nmpc_b200_load_plugin(argv[1]);
  std::printf("load_rc %d\n", rc);
  if(rc != NMPC_B200_OK)
  {
    std::printf("load_error %s\n", nmpc_b200_last_error());
    return 1;
  }
  std::printf("load_again %d\n", nmpc_b200_load_plugin(argv[1]));
  const int known = nmpc_b200_model_dims("pendulum", &nx, &nu, &ng, &np) == NMPC_B200_OK ? 1 : 0;
  std::printf("known_after %d dims %d %d %d %d\n", known, nx, nu, ng, np);
  if(no_solve) return known ? 0 : 1;

  constexpr int B = 24, N = 80;
  auto problem = std::make_shared<nmpc_ddp::FunctorProblem<Pendulum<double>>>("pendulum");
  nmpc_ddp::DDPSolver<2, 1> solver(problem, B);
  solver.config().horizon_steps = N;
  solver.config().max_iter = 30;
  std::mt19937 gen(5);
  std::uniform_real_distribution<double> ang(-3.0, 3.0), vel(-1.0, 1.0);
  std::vector<double> x0(B * 2), u_init((size_t)B * N, 0.0);
  for(int b = 0; b < B; b++)
  {
    x0[2 * b] = ang(gen);
    x0[2 * b + 1] = vel(gen);
  }
  std::vector<int> status;
  solver.solveBatch(B, 0.0, x0.data(), u_init.data(), N, &status);
  std::vector<double> u((size_t)B * N), cost(B);
  std::vector<int> iters(B);
  solver.get(NMPC_B200_DDP_U, u.data(), u.size() * sizeof(double));
  solver.get(NMPC_B200_DDP_COST, cost.data(), cost.size() * sizeof(double));
  solver.get(NMPC_B200_DDP_ITERS, iters.data(), iters.size() * sizeof(int));

  std::vector<double> params(Pendulum<double>::NUM_PARAMS);
  Pendulum<double>::defaultParams(params.data());
  double max_du = 0, max_dc = 0;
  int iters_equal = 1, status_equal = 1, n_converged = 0;
  for(int b = 0; b < B; b++)
  {
    oracle::DDPSolver<2, 1> ref(std::make_shared<OraclePendulum>(params.data()));
    ref.config().print_level = 0;
    ref.config().horizon_steps = N;
    ref.config().max_iter = 30;
    oracle::Vec<2> cx;
    cx[0] = x0[2 * b];
    cx[1] = x0[2 * b + 1];
    const bool conv = ref.solve(0.0, cx, std::vector<oracle::Vec<1>>(N));
    double umax = 0, du = 0, csum = 0;
    for(int i = 0; i < N; i++)
    {
      umax = std::fmax(umax, std::fabs(ref.controlData().u_list[i][0]));
      du = std::fmax(du, std::fabs(ref.controlData().u_list[i][0] - u[(size_t)b * N + i]));
    }
    for(int i = 0; i <= N; i++) csum += ref.controlData().cost_list[i];
    max_du = std::fmax(max_du, du / (1 + umax));
    max_dc = std::fmax(max_dc, std::fabs(csum - cost[b]) / std::fabs(csum));
    if(ref.traceDataList().back().iter != iters[b]) iters_equal = 0;
    if((status[b] == 1) != conv) status_equal = 0;
    n_converged += conv ? 1 : 0;
  }
  std::printf("rel_du %.3e\nrel_dcost %.3e\niters_equal %d\nstatus_equal %d\nn_converged %d\n", max_du, max_dc, iters_equal,
              status_equal, n_converged);
  return 0;
}
