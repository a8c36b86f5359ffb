/* nmpc_b200 -- bipedal CoM-ZMP problem functor (device + host).
 *
 * Same problem as the reference's DDPProblemBipedal (isri-aist/NMPC
 * nmpc_ddp/tests/src/TestDDPBipedal.cpp:16-144) with the test's time-varying references
 * ref_zmp_func / omega2_func (:170-227) evaluated inside the functor (the reference passes them as
 * std::function, which cannot cross to the device).  State [CoM_pos, CoM_vel], input [ZMP].
 * Flat parameter layout: [dt, running_vel, running_zmp, terminal_pos, terminal_vel, end_t].
 */
#pragma once

#include <cmath>

#include <nmpc_b200/matrix.h>

namespace nmpc_b200
{
namespace models
{
template<class S = double>
struct Bipedal
{
  static constexpr int NX = 2;
  static constexpr int NU = 1;
  static constexpr int NUM_PARAMS = 6;

  using Scalar = S;
  using StateDimVector = Matrix<S, NX, 1>;
  using InputDimVector = Matrix<S, NU, 1>;
  using StateStateDimMatrix = Matrix<S, NX, NX>;
  using InputInputDimMatrix = Matrix<S, NU, NU>;
  using StateInputDimMatrix = Matrix<S, NX, NU>;

  S dt_ = S(0.01);
  S running_vel = S(1e-14);
  S running_zmp = S(1e-1);
  S terminal_pos = S(1e2);
  S terminal_vel = S(1.0);
  S end_t = S(20.0);

  static Bipedal fromParams(const double * p)
  {
    Bipedal m;
    m.dt_ = S(p[0]);
    m.running_vel = S(p[1]);
    m.running_zmp = S(p[2]);
    m.terminal_pos = S(p[3]);
    m.terminal_vel = S(p[4]);
    m.end_t = S(p[5]);
    return m;
  }
  static void defaultParams(double * p)
  {
    const double d[NUM_PARAMS] = {0.01, 1e-14, 1e-1, 1e2, 1.0, 20.0};
    for(int i = 0; i < NUM_PARAMS; i++) p[i] = d[i];
  }
  NMPC_HD S dt() const
  {
    return dt_;
  }

  NMPC_HD static S minJerk(S t)
  {
    const S t3 = t * t * t;
    return S(6) * (t3 * t * t) + S(-15) * (t3 * t) + S(10) * t3;
  }
  NMPC_HD static S minJerkSecondDeriv(S t)
  {
    return S(120) * (t * t * t) + S(-180) * (t * t) + S(60) * t;
  }
  NMPC_HD S refZmp(S t) const
  {
    t += S(1e-6);
    if(t <= S(1.5) || t >= end_t - S(1.5)) return S(0.0);
    return (static_cast<int>(floor((t - S(1.0)) / S(1.0))) % 2 == 0) ? S(0.15) : S(-0.15);
  }
  NMPC_HD S omega2(S t) const
  {
    t += S(1e-6);
    const S cog_pos_z_high = S(1.0);
    const S cog_pos_z_low = S(0.3);
    S cog_pos_z = S(0.0);
    S cog_acc_z = S(0.0);
    if(t < S(7.0))
    {
      cog_pos_z = cog_pos_z_high;
    }
    else if(t < S(8.0))
    {
      const S scale = cog_pos_z_low - cog_pos_z_high;
      cog_pos_z = scale * minJerk(t - S(7.0)) + cog_pos_z_high;
      cog_acc_z = scale * minJerkSecondDeriv(t - S(7.0));
    }
    else if(t < S(12.0))
    {
      cog_pos_z = cog_pos_z_low;
    }
    else if(t < S(13.0))
    {
      const S scale = cog_pos_z_high - cog_pos_z_low;
      cog_pos_z = scale * minJerk(t - S(12.0)) + cog_pos_z_low;
      cog_acc_z = scale * minJerkSecondDeriv(t - S(12.0));
    }
    else
    {
      cog_pos_z = cog_pos_z_high;
    }
    return (cog_acc_z + S(9.80665)) / cog_pos_z;
  }
  NMPC_HD StateStateDimMatrix A(S t) const
  {
    StateStateDimMatrix A;
    const S w2 = omega2(t);
    A(0, 0) = S(1) + S(0.5) * dt_ * dt_ * w2;
    A(0, 1) = dt_;
    A(1, 0) = dt_ * w2;
    A(1, 1) = S(1);
    return A;
  }
  NMPC_HD StateInputDimMatrix B(S t) const
  {
    StateInputDimMatrix B;
    const S w2 = omega2(t);
    B[0] = S(-0.5) * dt_ * dt_ * w2;
    B[1] = S(-1) * dt_ * w2;
    return B;
  }

  NMPC_HD StateDimVector stateEq(S t, const StateDimVector & x, const InputDimVector & u) const
  {
    return A(t) * x + B(t) * u;
  }
  NMPC_HD S runningCost(S t, const StateDimVector & x, const InputDimVector & u) const
  {
    const S e = u[0] - refZmp(t);
    return running_vel * S(0.5) * (x[1] * x[1]) + running_zmp * S(0.5) * (e * e);
  }
  NMPC_HD S terminalCost(S t, const StateDimVector & x) const
  {
    const S e = x[0] - refZmp(t);
    return terminal_pos * S(0.5) * (e * e) + terminal_vel * S(0.5) * (x[1] * x[1]);
  }
  NMPC_HD void calcStateEqDeriv(S t,
                                const StateDimVector &,
                                const InputDimVector &,
                                StateStateDimMatrix & state_eq_deriv_x,
                                StateInputDimMatrix & state_eq_deriv_u) const
  {
    state_eq_deriv_x = A(t);
    state_eq_deriv_u = B(t);
  }
  NMPC_HD void calcRunningCostDeriv(S t,
                                    const StateDimVector & x,
                                    const InputDimVector & u,
                                    StateDimVector & running_cost_deriv_x,
                                    InputDimVector & running_cost_deriv_u,
                                    StateStateDimMatrix & running_cost_deriv_xx,
                                    InputInputDimMatrix & running_cost_deriv_uu,
                                    StateInputDimMatrix & running_cost_deriv_xu) const
  {
    running_cost_deriv_x[0] = S(0);
    running_cost_deriv_x[1] = running_vel * x[1];
    running_cost_deriv_u[0] = running_zmp * (u[0] - refZmp(t));
    running_cost_deriv_xx.setZero();
    running_cost_deriv_xx(1, 1) = running_vel;
    running_cost_deriv_uu(0, 0) = running_zmp;
    running_cost_deriv_xu.setZero();
  }
  NMPC_HD void calcTerminalCostDeriv(S t,
                                     const StateDimVector & x,
                                     StateDimVector & terminal_cost_deriv_x,
                                     StateStateDimMatrix & terminal_cost_deriv_xx) const
  {
    terminal_cost_deriv_x[0] = terminal_pos * (x[0] - refZmp(t));
    terminal_cost_deriv_x[1] = terminal_vel * x[1];
    terminal_cost_deriv_xx.setZero();
    terminal_cost_deriv_xx(0, 0) = terminal_pos;
    terminal_cost_deriv_xx(1, 1) = terminal_vel;
  }
};
} // namespace models
} // namespace nmpc_b200
