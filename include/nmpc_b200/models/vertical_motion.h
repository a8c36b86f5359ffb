/* nmpc_b200 -- vertical-motion problem functor with a TIME-VARYING input dimension (device + host).
 *
 * Same problem as the reference's DDPProblemVerticalMotion, a DDPProblem<2, Eigen::Dynamic> (isri-aist/NMPC
 * nmpc_ddp/tests/src/TestDDPVerticalMotion.cpp:25-234): state [pos_z, vel_z], input = the contact forces, whose
 * NUMBER changes along the horizon (inputDim(t), :61-78: two contacts for 2 < t < 3, none for 4.5 < t < 5, one
 * otherwise), reference height ref_pos_func of the test (:245-258).
 *
 * A kernel needs compile-time sizes, so the functor declares NU = the largest dimension and inputDim(t) <= NU;
 * inputs a >= inputDim(t) are PADDING: the engine keeps them at zero and replaces their rows / columns of
 * (Fu, Lu, Luu, Lxu) by (0, 0, unit diagonal, 0) after every linearisation, which makes Quu block diagonal
 * [Quu_active, 1 (+lambda)]: the gains of the padding are exactly zero and every active quantity is the one the
 * reference computes on the reduced system (tests/golden pins this against the reference's Dynamic code path).
 *
 * Flat parameter layout: [dt, running_x0, running_x1, running_u, terminal_x0, terminal_x1, mass, ref_switch_t].
 */
#pragma once

#include <nmpc_b200/matrix.h>

namespace nmpc_b200
{
namespace models
{
template<class S = double>
struct VerticalMotion
{
  static constexpr int NX = 2;
  static constexpr int NU = 2; //!< largest inputDim(t)
  static constexpr int NUM_PARAMS = 8;

  using Scalar = S;
  using StateDimVector = Matrix<S, NX, 1>;
  using InputDimVector = Matrix<S, NU, 1>;
  using StateStateDimMatrix = Matrix<S, NX, NX>;
  using InputInputDimMatrix = Matrix<S, NU, NU>;
  using StateInputDimMatrix = Matrix<S, NX, NU>;

  S dt_ = S(0.01);
  S running_x[2] = {S(1.0), S(1e-3)}; // CostWeight (:36-46)
  S running_u = S(1e-4);
  S terminal_x[2] = {S(1.0), S(1e-3)};
  S mass_ = S(1.0);
  S ref_switch_t = S(8.0);

  static constexpr double g_ = 9.80665; // [m/s^2]

  static VerticalMotion fromParams(const double * p)
  {
    VerticalMotion m;
    m.dt_ = S(p[0]);
    m.running_x[0] = S(p[1]), m.running_x[1] = S(p[2]);
    m.running_u = S(p[3]);
    m.terminal_x[0] = S(p[4]), m.terminal_x[1] = S(p[5]);
    m.mass_ = S(p[6]);
    m.ref_switch_t = S(p[7]);
    return m;
  }
  static void defaultParams(double * p)
  {
    const double d[NUM_PARAMS] = {0.01, 1.0, 1e-3, 1e-4, 1.0, 1e-3, 1.0, 8.0};
    for(int i = 0; i < NUM_PARAMS; i++) p[i] = d[i];
  }
  NMPC_HD S dt() const
  {
    return dt_;
  }

  /** DDPProblem::inputDim(t) (DDPProblem.h:72-85; TestDDPVerticalMotion.cpp:61-78). */
  NMPC_HD int inputDim(S t) const
  {
    // Add small values to avoid numerical instability at inequality bounds
    const S epsilon_t = S(1e-6);
    t += epsilon_t;
    if(S(2.0) < t && t < S(3.0))
    {
      return 2;
    }
    else if(S(4.5) < t && t < S(5.0))
    {
      return 0;
    }
    else
    {
      return 1;
    }
  }

  /** ref_pos_func of the test (:245-258). */
  NMPC_HD S refPos(S t) const
  {
    t += S(1e-6);
    return (t < ref_switch_t) ? S(1.0) : S(0.0);
  }

  NMPC_HD StateDimVector stateEq(S, const StateDimVector & x, const InputDimVector & u) const
  {
    // x_dot << x[1], u.sum() / mass_ - g_   (padding inputs are zero)
    StateDimVector x_dot;
    x_dot[0] = x[1];
    x_dot[1] = (u[0] + u[1]) / mass_ - S(g_);
    StateDimVector out;
    out[0] = x[0] + dt_ * x_dot[0];
    out[1] = x[1] + dt_ * x_dot[1];
    return out;
  }

  NMPC_HD S runningCost(S t, const StateDimVector & x, const InputDimVector & u) const
  {
    const S e0 = x[0] - refPos(t), e1 = x[1] - S(0);
    const S cost_x = S(0.5) * (running_x[0] * (e0 * e0) + running_x[1] * (e1 * e1));
    const S cost_u = S(0.5) * running_u * (u[0] * u[0] + u[1] * u[1]);
    return cost_x + cost_u;
  }

  NMPC_HD S terminalCost(S t, const StateDimVector & x) const
  {
    const S e0 = x[0] - refPos(t), e1 = x[1] - S(0);
    return S(0.5) * (terminal_x[0] * (e0 * e0) + terminal_x[1] * (e1 * e1));
  }

  NMPC_HD void calcStateEqDeriv(S, const StateDimVector &, const InputDimVector &, StateStateDimMatrix & Fx,
                                StateInputDimMatrix & Fu) const
  {
    Fx.setZero();
    Fx(0, 1) = S(1);
    Fx *= dt_;
    Fx(0, 0) += S(1);
    Fx(1, 1) += S(1);
    Fu.setZero();
    Fu(1, 0) = S(1) / mass_;
    Fu(1, 1) = S(1) / mass_;
    Fu *= dt_;
  }

  NMPC_HD void calcRunningCostDeriv(S t, const StateDimVector & x, const InputDimVector & u, StateDimVector & Lx,
                                    InputDimVector & Lu, StateStateDimMatrix & Lxx, InputInputDimMatrix & Luu,
                                    StateInputDimMatrix & Lxu) const
  {
    Lx[0] = running_x[0] * (x[0] - refPos(t));
    Lx[1] = running_x[1] * (x[1] - S(0));
    Lxx.setZero();
    Lxx(0, 0) = running_x[0];
    Lxx(1, 1) = running_x[1];
    Lxu.setZero();
    Lu[0] = running_u * u[0];
    Lu[1] = running_u * u[1];
    Luu.setZero();
    Luu(0, 0) = S(1);
    Luu(1, 1) = S(1);
    Luu *= running_u;
  }

  NMPC_HD void calcTerminalCostDeriv(S t, const StateDimVector & x, StateDimVector & Vx, StateStateDimMatrix & Vxx) const
  {
    Vx[0] = terminal_x[0] * (x[0] - refPos(t));
    Vx[1] = terminal_x[1] * (x[1] - S(0));
    Vxx.setZero();
    Vxx(0, 0) = terminal_x[0];
    Vxx(1, 1) = terminal_x[1];
  }
};
} // namespace models
} // namespace nmpc_b200
