/* TEST FIXTURE -- a problem functor that is NOT part of libnmpc_b200.so: a damped pendulum with a torque input,
   state [theta, omega], explicit Euler; swing-up cost about theta = 0.  Written the way a user of the reference
   writes a DDPProblem<2, 1> (method names / argument order of DDPProblem.h:99-198).
   params: [dt, mass, length, damping, running_x[2], running_u, terminal_x[2]] */
#pragma once

#include <cmath>

#include <nmpc_b200/matrix.h>

template<class S = double>
struct Pendulum
{
  static constexpr int NX = 2;
  static constexpr int NU = 1;
  static constexpr int NUM_PARAMS = 9;
  using Scalar = S;
  using StateDimVector = nmpc_b200::Matrix<S, NX, 1>;
  using InputDimVector = nmpc_b200::Matrix<S, NU, 1>;
  using StateStateDimMatrix = nmpc_b200::Matrix<S, NX, NX>;
  using InputInputDimMatrix = nmpc_b200::Matrix<S, NU, NU>;
  using StateInputDimMatrix = nmpc_b200::Matrix<S, NX, NU>;

  S dt_ = S(0.02), mass = S(1.0), length = S(0.5), damping = S(0.05);
  S running_x[2] = {S(1.0), S(0.1)};
  S running_u = S(0.05);
  S terminal_x[2] = {S(50.0), S(5.0)};
  static constexpr double g_ = 9.80665;

  static Pendulum fromParams(const double * p)
  {
    Pendulum m;
    m.dt_ = S(p[0]);
    m.mass = S(p[1]);
    m.length = S(p[2]);
    m.damping = S(p[3]);
    m.running_x[0] = S(p[4]);
    m.running_x[1] = S(p[5]);
    m.running_u = S(p[6]);
    m.terminal_x[0] = S(p[7]);
    m.terminal_x[1] = S(p[8]);
    return m;
  }
  static void defaultParams(double * p)
  {
    const double d[NUM_PARAMS] = {0.02, 1.0, 0.5, 0.05, 1.0, 0.1, 0.05, 50.0, 5.0};
    for(int i = 0; i < NUM_PARAMS; i++) p[i] = d[i];
  }
  NMPC_HD S dt() const
  {
    return dt_;
  }
  NMPC_HD S inertia() const
  {
    return mass * length * length;
  }
  NMPC_HD StateDimVector stateEq(S, const StateDimVector & x, const InputDimVector & u) const
  {
    StateDimVector x_dot;
    x_dot[0] = x[1];
    x_dot[1] = (u[0] - damping * x[1] + mass * S(g_) * length * sin(x[0])) / inertia();
    return x + dt_ * x_dot;
  }
  NMPC_HD S runningCost(S, const StateDimVector & x, const InputDimVector & u) const
  {
    return S(0.5) * (running_x[0] * (x[0] * x[0]) + running_x[1] * (x[1] * x[1])) + S(0.5) * running_u * (u[0] * u[0]);
  }
  NMPC_HD S terminalCost(S, const StateDimVector & x) const
  {
    return S(0.5) * (terminal_x[0] * (x[0] * x[0]) + terminal_x[1] * (x[1] * x[1]));
  }
  NMPC_HD void calcStateEqDeriv(S, const StateDimVector & x, const InputDimVector &, StateStateDimMatrix & fx,
                                StateInputDimMatrix & fu) const
  {
    fx.setZero();
    fx(0, 1) = S(1);
    fx(1, 0) = mass * S(g_) * length * cos(x[0]) / inertia();
    fx(1, 1) = S(-1) * damping / inertia();
    fx *= dt_;
    fx.addToDiagonal(S(1));
    fu.setZero();
    fu[1] = dt_ / inertia();
  }
  NMPC_HD void calcRunningCostDeriv(S, const StateDimVector & x, const InputDimVector & u, StateDimVector & lx,
                                    InputDimVector & lu, StateStateDimMatrix & lxx, InputInputDimMatrix & luu,
                                    StateInputDimMatrix & lxu) const
  {
    lx[0] = running_x[0] * x[0];
    lx[1] = running_x[1] * x[1];
    lu[0] = running_u * u[0];
    lxx.setZero();
    lxx(0, 0) = running_x[0];
    lxx(1, 1) = running_x[1];
    luu(0, 0) = running_u;
    lxu.setZero();
  }
  NMPC_HD void calcTerminalCostDeriv(S, const StateDimVector & x, StateDimVector & vx, StateStateDimMatrix & vxx) const
  {
    vx[0] = terminal_x[0] * x[0];
    vx[1] = terminal_x[1] * x[1];
    vxx.setZero();
    vxx(0, 0) = terminal_x[0];
    vxx(1, 1) = terminal_x[1];
  }
};
