/* nmpc_b200 -- Van der Pol oscillator problem functor (device + host).
 *
 * Same problem as the reference's FmpcProblemOscillator (isri-aist/NMPC
 * nmpc_fmpc/tests/src/TestFmpcOscillator.cpp:18-135): state [x0, x1], input [u], three inequalities
 * (-x1 - 0.05 <= 0, -u - 1 <= 0, u - 0.9 <= 0).  Flat parameter layout: [dt].
 */
#pragma once

#include <nmpc_b200/matrix.h>

namespace nmpc_b200
{
namespace models
{
template<class S = double>
struct Oscillator
{
  static constexpr int NX = 2;
  static constexpr int NU = 1;
  static constexpr int NG = 3;
  static constexpr int NUM_PARAMS = 1;

  using Scalar = S;
  using StateDimVector = Matrix<S, NX, 1>;
  using InputDimVector = Matrix<S, NU, 1>;
  using IneqDimVector = Matrix<S, NG, 1>;
  using StateStateDimMatrix = Matrix<S, NX, NX>;
  using InputInputDimMatrix = Matrix<S, NU, NU>;
  using StateInputDimMatrix = Matrix<S, NX, NU>;
  using IneqStateDimMatrix = Matrix<S, NG, NX>;
  using IneqInputDimMatrix = Matrix<S, NG, NU>;

  S dt_ = S(0.01);

  static Oscillator fromParams(const double * p)
  {
    Oscillator m;
    m.dt_ = S(p[0]);
    return m;
  }
  static void defaultParams(double * p)
  {
    p[0] = 0.01;
  }
  NMPC_HD S dt() const
  {
    return dt_;
  }

  NMPC_HD StateDimVector stateEq(S t, const StateDimVector & x, const InputDimVector & u) const
  {
    return stateEq(t, x, u, dt_);
  }
  NMPC_HD StateDimVector stateEq(S, const StateDimVector & x, const InputDimVector & u, S dt) const
  {
    StateDimVector x_dot;
    x_dot[0] = (S(1.0) - x[1] * x[1]) * x[0] - x[1] + u[0];
    x_dot[1] = x[0];
    return x + dt * x_dot;
  }
  NMPC_HD S runningCost(S, const StateDimVector & x, const InputDimVector & u) const
  {
    return S(0.5) * (x.squaredNorm() + u.squaredNorm());
  }
  NMPC_HD S terminalCost(S, const StateDimVector &) const
  {
    return S(0);
  }
  NMPC_HD IneqDimVector ineqConst(S, const StateDimVector & x, const InputDimVector & u) const
  {
    IneqDimVector g;
    g[0] = S(-1) * x[1] - S(0.05);
    g[1] = S(-1) * u[0] - S(1.0);
    g[2] = u[0] - S(0.9);
    return g;
  }
  NMPC_HD void calcStateEqDeriv(S,
                                const StateDimVector & x,
                                const InputDimVector &,
                                StateStateDimMatrix & state_eq_deriv_x,
                                StateInputDimMatrix & state_eq_deriv_u) const
  {
    state_eq_deriv_x.setZero();
    state_eq_deriv_x(0, 0) = S(1.0) - x[1] * x[1];
    state_eq_deriv_x(0, 1) = S(-2) * x[0] * x[1] - S(1.0);
    state_eq_deriv_x(1, 0) = S(1);
    state_eq_deriv_x *= dt_;
    state_eq_deriv_x.addToDiagonal(S(1));

    state_eq_deriv_u.setZero();
    state_eq_deriv_u(0, 0) = S(1);
    state_eq_deriv_u *= dt_;
  }
  NMPC_HD void calcRunningCostDeriv(S,
                                    const StateDimVector & x,
                                    const InputDimVector & u,
                                    StateDimVector & running_cost_deriv_x,
                                    InputDimVector & running_cost_deriv_u,
                                    StateStateDimMatrix & running_cost_deriv_xx,
                                    InputInputDimMatrix & running_cost_deriv_uu,
                                    StateInputDimMatrix & running_cost_deriv_xu) const
  {
    running_cost_deriv_x = x;
    running_cost_deriv_u = u;
    running_cost_deriv_xx.setIdentity();
    running_cost_deriv_uu.setIdentity();
    running_cost_deriv_xu.setZero();
  }
  NMPC_HD void calcTerminalCostDeriv(S,
                                     const StateDimVector &,
                                     StateDimVector & terminal_cost_deriv_x,
                                     StateStateDimMatrix & terminal_cost_deriv_xx) const
  {
    terminal_cost_deriv_x.setZero();
    terminal_cost_deriv_xx.setZero();
  }
  NMPC_HD void calcIneqConstDeriv(S,
                                  const StateDimVector &,
                                  const InputDimVector &,
                                  IneqStateDimMatrix & ineq_const_deriv_x,
                                  IneqInputDimMatrix & ineq_const_deriv_u) const
  {
    ineq_const_deriv_x.setZero();
    ineq_const_deriv_x(0, 1) = S(-1);

    ineq_const_deriv_u.setZero();
    ineq_const_deriv_u(1, 0) = S(-1);
    ineq_const_deriv_u(2, 0) = S(1);
  }
};
} // namespace models
} // namespace nmpc_b200
