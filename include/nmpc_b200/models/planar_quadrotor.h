/* nmpc_b200 -- planar quadrotor with thrust limits: an FMPC problem with TWO inputs and four inequalities.
 *
 * The reference's FMPC tests (cart-pole, Van der Pol oscillator) have one input, so the n_u > 1 branch of
 * FmpcSolver::backwardPass -- Eigen::LDLT of G with diagonal pivoting, Eigen::FullPivLU when its info() is not
 * Success (isri-aist/NMPC nmpc_fmpc/include/nmpc_fmpc/FmpcSolver.hpp:596-617) -- is never exercised by them.  This
 * functor is the smallest natural problem that does: state [px, pz, theta, vx, vz, omega], inputs [T1, T2] (rotor
 * thrusts), explicit Euler step like every reference model,
 *
 *   px' = vx, pz' = vz, theta' = omega,
 *   vx' = -(T1 + T2) sin(theta) / m,  vz' = (T1 + T2) cos(theta) / m - g,  omega' = r (T1 - T2) / J
 *
 * quadratic running / terminal cost about (ref_px, ref_pz, 0, 0, 0, 0) and the hover thrust, with an optional
 * input CROSS weight (Luu = [[w, c], [c, w]]: w = 0, c != 0 and zero multipliers give the indefinite, zero-diagonal
 * G that makes LDLT report NumericalIssue), and 0 <= T_a <= thrust_max as g = (-T1, T1 - max, -T2, T2 - max) <= 0.
 * Method names and argument order follow nmpc_ddp::DDPProblem / nmpc_fmpc::FmpcProblem (DDPProblem.h:99-198,
 * FmpcProblem.h:94-107).  The same problem in Eigen idiom drives the reference's own FmpcSolver<6, 2, 4> for the golden
 * vectors of tests/test_fmpc_extra.py.
 *
 * Flat parameters: [dt, mass, inertia, arm, thrust_max, running_x[6], running_u, running_u_cross, terminal_x[6],
 *                   ref_px, ref_pz]
 */
#pragma once

#include <cmath>

#include <nmpc_b200/matrix.h>

namespace nmpc_b200
{
namespace models
{
template<class S = double>
struct PlanarQuadrotor
{
  static constexpr int NX = 6;
  static constexpr int NU = 2;
  static constexpr int NG = 4;
  static constexpr int NUM_PARAMS = 21;

  using Scalar = S;
  using StateDimVector = Matrix<S, NX, 1>;
  using InputDimVector = Matrix<S, NU, 1>;
  using IneqDimVector = Matrix<S, NG, 1>;
  using StateStateDimMatrix = Matrix<S, NX, NX>;
  using InputInputDimMatrix = Matrix<S, NU, NU>;
  using StateInputDimMatrix = Matrix<S, NX, NU>;
  using IneqStateDimMatrix = Matrix<S, NG, NX>;
  using IneqInputDimMatrix = Matrix<S, NG, NU>;

  S dt_ = S(0.02);
  S mass = S(1.0);
  S inertia = S(0.02);
  S arm = S(0.2);
  S thrust_max = S(10.0);
  S running_x[NX] = {S(1.0), S(1.0), S(0.1), S(0.1), S(0.1), S(0.01)};
  S running_u = S(0.01);
  S running_u_cross = S(0.0);
  S terminal_x[NX] = {S(10.0), S(10.0), S(1.0), S(1.0), S(1.0), S(0.1)};
  S ref_px = S(0.0);
  S ref_pz = S(0.0);

  static constexpr double g_ = 9.80665;

  static PlanarQuadrotor fromParams(const double * p)
  {
    PlanarQuadrotor m;
    m.dt_ = S(p[0]);
    m.mass = S(p[1]);
    m.inertia = S(p[2]);
    m.arm = S(p[3]);
    m.thrust_max = S(p[4]);
    for(int i = 0; i < NX; i++) m.running_x[i] = S(p[5 + i]);
    m.running_u = S(p[11]);
    m.running_u_cross = S(p[12]);
    for(int i = 0; i < NX; i++) m.terminal_x[i] = S(p[13 + i]);
    m.ref_px = S(p[19]);
    m.ref_pz = S(p[20]);
    return m;
  }

  static void defaultParams(double * p)
  {
    const double d[NUM_PARAMS] = {0.02, 1.0, 0.02, 0.2, 10.0, 1.0, 1.0, 0.1, 0.1, 0.1, 0.01, 0.01, 0.0,
                                  10.0, 10.0, 1.0, 1.0, 1.0, 0.1, 0.0, 0.0};
    for(int i = 0; i < NUM_PARAMS; i++) p[i] = d[i];
  }

  NMPC_HD S dt() const
  {
    return dt_;
  }

  NMPC_HD S hoverThrust() const
  {
    return S(0.5) * mass * S(g_);
  }

  NMPC_HD S stateRef(int i) const
  {
    return i == 0 ? ref_px : (i == 1 ? ref_pz : S(0));
  }

  NMPC_HD StateDimVector stateEq(S t, const StateDimVector & x, const InputDimVector & u) const
  {
    return stateEq(t, x, u, dt_);
  }

  NMPC_HD StateDimVector stateEq(S, const StateDimVector & x, const InputDimVector & u, S dt) const
  {
    const S st = sin(x[2]), ct = cos(x[2]);
    const S thrust = u[0] + u[1];
    StateDimVector x_dot;
    x_dot[0] = x[3];
    x_dot[1] = x[4];
    x_dot[2] = x[5];
    x_dot[3] = S(-1) * thrust * st / mass;
    x_dot[4] = thrust * ct / mass - S(g_);
    x_dot[5] = arm * (u[0] - u[1]) / inertia;
    return x + dt * x_dot;
  }

  NMPC_HD S runningCost(S, const StateDimVector & x, const InputDimVector & u) const
  {
    S sx = S(0);
NMPC_UNROLL
    for(int i = 0; i < NX; i++)
    {
      const S e = x[i] - stateRef(i);
      sx += running_x[i] * (e * e);
    }
    const S e0 = u[0] - hoverThrust(), e1 = u[1] - hoverThrust();
    const S su = running_u * (e0 * e0 + e1 * e1);
    return (S(0.5) * sx + S(0.5) * su) + running_u_cross * (e0 * e1);
  }

  NMPC_HD S terminalCost(S, const StateDimVector & x) const
  {
    S sx = S(0);
NMPC_UNROLL
    for(int i = 0; i < NX; i++)
    {
      const S e = x[i] - stateRef(i);
      sx += terminal_x[i] * (e * e);
    }
    return S(0.5) * sx;
  }

  NMPC_HD void calcStateEqDeriv(S,
                                const StateDimVector & x,
                                const InputDimVector & u,
                                StateStateDimMatrix & state_eq_deriv_x,
                                StateInputDimMatrix & state_eq_deriv_u) const
  {
    const S st = sin(x[2]), ct = cos(x[2]);
    const S thrust = u[0] + u[1];
    state_eq_deriv_x.setZero();
    state_eq_deriv_x(0, 3) = S(1);
    state_eq_deriv_x(1, 4) = S(1);
    state_eq_deriv_x(2, 5) = S(1);
    state_eq_deriv_x(3, 2) = S(-1) * thrust * ct / mass;
    state_eq_deriv_x(4, 2) = S(-1) * thrust * st / mass;
    state_eq_deriv_x *= dt_;
    state_eq_deriv_x.addToDiagonal(S(1));

    state_eq_deriv_u.setZero();
    state_eq_deriv_u(3, 0) = S(-1) * st / mass;
    state_eq_deriv_u(3, 1) = S(-1) * st / mass;
    state_eq_deriv_u(4, 0) = ct / mass;
    state_eq_deriv_u(4, 1) = ct / mass;
    state_eq_deriv_u(5, 0) = arm / inertia;
    state_eq_deriv_u(5, 1) = S(-1) * arm / inertia;
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
    running_cost_deriv_xx.setZero();
NMPC_UNROLL
    for(int i = 0; i < NX; i++)
    {
      running_cost_deriv_x[i] = running_x[i] * (x[i] - stateRef(i));
      running_cost_deriv_xx(i, i) = running_x[i];
    }
    const S e0 = u[0] - hoverThrust(), e1 = u[1] - hoverThrust();
    running_cost_deriv_u[0] = running_u * e0 + running_u_cross * e1;
    running_cost_deriv_u[1] = running_u * e1 + running_u_cross * e0;
    running_cost_deriv_uu(0, 0) = running_u;
    running_cost_deriv_uu(1, 1) = running_u;
    running_cost_deriv_uu(0, 1) = running_u_cross;
    running_cost_deriv_uu(1, 0) = running_u_cross;
    running_cost_deriv_xu.setZero();
  }

  NMPC_HD void calcTerminalCostDeriv(S,
                                     const StateDimVector & x,
                                     StateDimVector & terminal_cost_deriv_x,
                                     StateStateDimMatrix & terminal_cost_deriv_xx) const
  {
    terminal_cost_deriv_xx.setZero();
NMPC_UNROLL
    for(int i = 0; i < NX; i++)
    {
      terminal_cost_deriv_x[i] = terminal_x[i] * (x[i] - stateRef(i));
      terminal_cost_deriv_xx(i, i) = terminal_x[i];
    }
  }

  NMPC_HD IneqDimVector ineqConst(S, const StateDimVector &, const InputDimVector & u) const
  {
    IneqDimVector g;
    g[0] = S(-1) * u[0];
    g[1] = u[0] - thrust_max;
    g[2] = S(-1) * u[1];
    g[3] = u[1] - thrust_max;
    return g;
  }

  NMPC_HD void calcIneqConstDeriv(S,
                                  const StateDimVector &,
                                  const InputDimVector &,
                                  IneqStateDimMatrix & ineq_const_deriv_x,
                                  IneqInputDimMatrix & ineq_const_deriv_u) const
  {
    ineq_const_deriv_x.setZero();
    ineq_const_deriv_u.setZero();
    ineq_const_deriv_u(0, 0) = S(-1);
    ineq_const_deriv_u(1, 0) = S(1);
    ineq_const_deriv_u(2, 1) = S(-1);
    ineq_const_deriv_u(3, 1) = S(1);
  }
};
} // namespace models
} // namespace nmpc_b200
