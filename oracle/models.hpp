// TEST INFRASTRUCTURE ONLY -- CPU oracle for nmpc_b200 (see oracle/README.md).
//
// Problem definitions restated from the reference's test programs (the only models the
// reference ships).  Each is an independent restatement -- the product's device functors live in
// include/nmpc_b200/models/ and are compared against these by tests/.
//
// Flat parameter layouts (doubles), shared by convention with the product's C-ABI:
//   cartpole : [dt, cart_mass, pole_mass, pole_length, running_x[4], running_u, terminal_x[4], ref_pos]  (14)
//   bipedal  : [dt, running_vel, running_zmp, terminal_pos, terminal_vel, end_t]                           (6)
#pragma once

#include <cmath>

#include "ddp_oracle.hpp"

namespace oracle
{
/** Cart-pole, nmpc_ddp/tests/src/TestDDPCartPole.cpp:28-234.
    State [pos, theta, vel, omega], input [force]; ref_pos_func_ is a constant here. */
class DDPProblemCartPole : public DDPProblem<4, 1>
{
public:
  static constexpr int kNumParams = 14;
  static constexpr double g_ = 9.80665; // :230

  explicit DDPProblemCartPole(const double * p)
  : DDPProblem<4, 1>(p[0]), cart_mass(p[1]), pole_mass(p[2]), pole_length(p[3]), ref_pos(p[13])
  {
    for(int i = 0; i < 4; i++) running_x[i] = p[4 + i];
    running_u = p[8];
    for(int i = 0; i < 4; i++) terminal_x[i] = p[9 + i];
  }

  /** Defaults as the rostest launch file sets them (TestDDPCartPole.test:12-24): running_u = 0.01. */
  static void defaultParams(double * p)
  {
    const double d[kNumParams] = {0.01, 1.0, 0.5, 2.0, 0.1, 1.0, 0.01, 0.1, 0.01, 0.1, 1.0, 0.01, 0.1, 0.0};
    for(int i = 0; i < kNumParams; i++) p[i] = d[i];
  }

  StateDimVector stateEq(double t, const StateDimVector & x, const InputDimVector & u) const override
  {
    return stateEq(t, x, u, dt_); // :63-66
  }

  StateDimVector stateEq(double, const StateDimVector & x, const InputDimVector & u, double dt) const
  {
    // :68-98
    double theta = x[1];
    double vel = x[2];
    double omega = x[3];
    double f = u[0];

    double m1 = cart_mass;
    double m2 = pole_mass;
    double l = pole_length;

    double sin_theta = std::sin(theta);
    double cos_theta = std::cos(theta);
    double omega2 = std::pow(omega, 2);
    double denom = m1 + m2 * std::pow(sin_theta, 2);

    StateDimVector x_dot;
    x_dot[0] = vel;
    x_dot[1] = omega;
    x_dot[2] = (f - m2 * l * omega2 * sin_theta + m2 * g_ * sin_theta * cos_theta) / denom;
    x_dot[3] = (f * cos_theta - m2 * l * omega2 * sin_theta * cos_theta + g_ * (m1 + m2) * sin_theta) / (l * denom);

    StateDimVector x_next;
    for(int i = 0; i < 4; i++) x_next[i] = x[i] + dt * x_dot[i];
    return x_next;
  }

  double runningCost(double, const StateDimVector & x, const InputDimVector & u) const override
  {
    // :100-105   0.5 * wx . (x - ref)^2 + 0.5 * wu . u^2
    double ref_x[4] = {ref_pos, 0, 0, 0};
    double sx = 0.0;
    for(int i = 0; i < 4; i++) sx += running_x[i] * ((x[i] - ref_x[i]) * (x[i] - ref_x[i]));
    double su = running_u * (u[0] * u[0]);
    return 0.5 * sx + 0.5 * su;
  }

  double terminalCost(double, const StateDimVector & x) const override
  {
    // :107-112
    double ref_x[4] = {ref_pos, 0, 0, 0};
    double sx = 0.0;
    for(int i = 0; i < 4; i++) sx += terminal_x[i] * ((x[i] - ref_x[i]) * (x[i] - ref_x[i]));
    return 0.5 * sx;
  }

  void calcStateEqDeriv(double,
                        const StateDimVector & x,
                        const InputDimVector & u,
                        StateStateDimMatrix & Fx,
                        StateInputDimMatrix & Fu) const override
  {
    // :114-159
    double theta = x[1];
    double omega = x[3];
    double f = u[0];

    double m1 = cart_mass;
    double m2 = pole_mass;
    double l = pole_length;

    double sin_theta = std::sin(theta);
    double cos_theta = std::cos(theta);
    double omega2 = std::pow(omega, 2);
    double denom = m1 + m2 * std::pow(sin_theta, 2);

    Fx.setZero();
    Fx(0, 2) = 1;
    Fx(1, 3) = 1;
    Fx(2, 1) = ((-1 * m2 * l * omega2 * cos_theta + m2 * g_ * (1 - 2 * std::pow(sin_theta, 2))) * denom
                + -1 * (f - m2 * l * omega2 * sin_theta + m2 * g_ * sin_theta * cos_theta)
                      * (2 * m2 * sin_theta * cos_theta))
               / std::pow(denom, 2);
    Fx(2, 3) = (-2 * m2 * l * omega * sin_theta) / denom;
    Fx(3, 1) = ((-1 * f * sin_theta + -1 * m2 * l * omega2 * (1 - 2 * std::pow(sin_theta, 2))
                 + g_ * (m1 + m2) * cos_theta)
                    * denom
                + -1 * (f * cos_theta - m2 * l * omega2 * sin_theta * cos_theta + g_ * (m1 + m2) * sin_theta)
                      * (2 * m2 * sin_theta * cos_theta))
               / (l * std::pow(denom, 2));
    Fx(3, 3) = (-2 * m2 * l * omega * sin_theta * cos_theta) / (l * denom);
    for(int i = 0; i < 16; i++) Fx.d[i] *= dt_;
    for(int i = 0; i < 4; i++) Fx(i, i) += 1.0;

    Fu.setZero();
    Fu[2] = 1 / denom;
    Fu[3] = cos_theta / (l * denom);
    for(int i = 0; i < 4; i++) Fu.d[i] *= dt_;
  }

  void calcRunningCostDeriv(double,
                            const StateDimVector & x,
                            const InputDimVector & u,
                            StateDimVector & Lx,
                            InputDimVector & Lu,
                            StateStateDimMatrix & Lxx,
                            InputInputDimMatrix & Luu,
                            StateInputDimMatrix & Lxu) const override
  {
    // :187-205
    double ref_x[4] = {ref_pos, 0, 0, 0};
    for(int i = 0; i < 4; i++) Lx[i] = running_x[i] * (x[i] - ref_x[i]);
    Lu[0] = running_u * u[0];
    Lxx.setZero();
    for(int i = 0; i < 4; i++) Lxx(i, i) = running_x[i];
    Luu(0, 0) = running_u;
    Lxu.setZero();
  }

  void calcTerminalCostDeriv(double, const StateDimVector & x, StateDimVector & Vx, StateStateDimMatrix & Vxx)
      const override
  {
    // :217-227
    double ref_x[4] = {ref_pos, 0, 0, 0};
    for(int i = 0; i < 4; i++) Vx[i] = terminal_x[i] * (x[i] - ref_x[i]);
    Vxx.setZero();
    for(int i = 0; i < 4; i++) Vxx(i, i) = terminal_x[i];
  }

  double cart_mass, pole_mass, pole_length;
  double running_x[4], running_u, terminal_x[4];
  double ref_pos;
};

/** Bipedal CoM-ZMP model, nmpc_ddp/tests/src/TestDDPBipedal.cpp:16-144, with the test's
    ref_zmp_func / omega2_func (:170-222) as the time-varying references. */
class DDPProblemBipedal : public DDPProblem<2, 1>
{
public:
  static constexpr int kNumParams = 6;

  explicit DDPProblemBipedal(const double * p)
  : DDPProblem<2, 1>(p[0]), running_vel(p[1]), running_zmp(p[2]), terminal_pos(p[3]), terminal_vel(p[4]), end_t(p[5])
  {
  }

  static void defaultParams(double * p)
  {
    // dt :161, CostWeight :21-29, end_t :164
    const double d[kNumParams] = {0.01, 1e-14, 1e-1, 1e2, 1.0, 20.0};
    for(int i = 0; i < kNumParams; i++) p[i] = d[i];
  }

  static double minJerk(double t) // :151-154
  {
    return 6 * std::pow(t, 5) + -15 * std::pow(t, 4) + 10 * std::pow(t, 3);
  }
  static double minJerkSecondDeriv(double t) // :156-159
  {
    return 120 * std::pow(t, 3) + -180 * std::pow(t, 2) + 60 * t;
  }

  double refZmp(double t) const // :170-191
  {
    constexpr double epsilon_t = 1e-6;
    t += epsilon_t;
    if(t <= 1.5 || t >= end_t - 1.5)
    {
      return 0.0;
    }
    else
    {
      if(static_cast<int>(std::floor((t - 1.0) / 1.0)) % 2 == 0)
      {
        return 0.15;
      }
      else
      {
        return -0.15;
      }
    }
  }

  double omega2(double t) const // :192-227
  {
    constexpr double epsilon_t = 1e-6;
    t += epsilon_t;
    double cog_pos_z_high = 1.0;
    double cog_pos_z_low = 0.3;
    double cog_pos_z = 0.0;
    double cog_acc_z = 0.0;
    if(t < 7.0)
    {
      cog_pos_z = cog_pos_z_high;
    }
    else if(t < 8.0)
    {
      double scale = cog_pos_z_low - cog_pos_z_high;
      cog_pos_z = scale * minJerk(t - 7.0) + cog_pos_z_high;
      cog_acc_z = scale * minJerkSecondDeriv(t - 7.0);
    }
    else if(t < 12.0)
    {
      cog_pos_z = cog_pos_z_low;
    }
    else if(t < 13.0)
    {
      double scale = cog_pos_z_high - cog_pos_z_low;
      cog_pos_z = scale * minJerk(t - 12.0) + cog_pos_z_low;
      cog_acc_z = scale * minJerkSecondDeriv(t - 12.0);
    }
    else
    {
      cog_pos_z = cog_pos_z_high;
    }
    constexpr double g = 9.80665;
    return (cog_acc_z + g) / cog_pos_z;
  }

  StateStateDimMatrix A(double t) const // :126-132
  {
    StateStateDimMatrix A;
    double w2 = omega2(t);
    A(0, 0) = 1 + 0.5 * dt_ * dt_ * w2;
    A(0, 1) = dt_;
    A(1, 0) = dt_ * w2;
    A(1, 1) = 1;
    return A;
  }

  StateInputDimMatrix B(double t) const // :134-140
  {
    StateInputDimMatrix B;
    double w2 = omega2(t);
    B[0] = -0.5 * dt_ * dt_ * w2;
    B[1] = -1 * dt_ * w2;
    return B;
  }

  StateDimVector stateEq(double t, const StateDimVector & x, const InputDimVector & u) const override
  {
    return add(mul(A(t), x), mul(B(t), u)); // :38-41
  }

  double runningCost(double t, const StateDimVector & x, const InputDimVector & u) const override
  {
    // :43-47
    return running_vel * 0.5 * std::pow(x[1], 2) + running_zmp * 0.5 * std::pow(u[0] - refZmp(t), 2);
  }

  double terminalCost(double t, const StateDimVector & x) const override
  {
    // :49-53
    return terminal_pos * 0.5 * std::pow(x[0] - refZmp(t), 2) + terminal_vel * 0.5 * std::pow(x[1], 2);
  }

  void calcStateEqDeriv(double t,
                        const StateDimVector &,
                        const InputDimVector &,
                        StateStateDimMatrix & Fx,
                        StateInputDimMatrix & Fu) const override
  {
    Fx = A(t); // :55-63
    Fu = B(t);
  }

  void calcRunningCostDeriv(double t,
                            const StateDimVector & x,
                            const InputDimVector & u,
                            StateDimVector & Lx,
                            InputDimVector & Lu,
                            StateStateDimMatrix & Lxx,
                            InputInputDimMatrix & Luu,
                            StateInputDimMatrix & Lxu) const override
  {
    // :91-104
    Lx[0] = 0;
    Lx[1] = running_vel * x[1];
    Lu[0] = running_zmp * (u[0] - refZmp(t));
    Lxx(0, 0) = 0;
    Lxx(0, 1) = 0;
    Lxx(1, 0) = 0;
    Lxx(1, 1) = running_vel;
    Luu(0, 0) = running_zmp;
    Lxu[0] = 0;
    Lxu[1] = 0;
  }

  void calcTerminalCostDeriv(double t, const StateDimVector & x, StateDimVector & Vx, StateStateDimMatrix & Vxx)
      const override
  {
    // :106-122
    Vx[0] = terminal_pos * (x[0] - refZmp(t));
    Vx[1] = terminal_vel * x[1];
    Vxx(0, 0) = terminal_pos;
    Vxx(0, 1) = 0;
    Vxx(1, 0) = 0;
    Vxx(1, 1) = terminal_vel;
  }

  double running_vel, running_zmp, terminal_pos, terminal_vel, end_t;
};
} // namespace oracle

#include "fmpc_oracle.hpp"

namespace oracle
{
/** FMPC cart-pole, nmpc_fmpc/tests/src/TestFmpcCartPole.cpp:32-256.  Dynamics, costs and their
    derivatives are textually identical to the DDP cart-pole (:69-226 vs TestDDPCartPole.cpp:63-227),
    so they are delegated; the additions are the four inequalities (:118-132) and C, D (:236-249). */
class FmpcProblemCartPole : public FmpcProblem<4, 1, 4>
{
public:
  static constexpr int kNumParams = DDPProblemCartPole::kNumParams;

  explicit FmpcProblemCartPole(const double * p) : FmpcProblem<4, 1, 4>(p[0]), base_(p) {}

  static void defaultParams(double * p)
  {
    DDPProblemCartPole::defaultParams(p); // TestFmpcCartPole.test:12-24 (running_u 0.01)
  }

  StateDimVector stateEq(double t, const StateDimVector & x, const InputDimVector & u) const override
  {
    return base_.stateEq(t, x, u);
  }
  double runningCost(double t, const StateDimVector & x, const InputDimVector & u) const override
  {
    return base_.runningCost(t, x, u);
  }
  double terminalCost(double t, const StateDimVector & x) const override
  {
    return base_.terminalCost(t, x);
  }
  IneqDimVector ineqConst(double, const StateDimVector & x, const InputDimVector & u) const override
  {
    // :118-132
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
  void calcStateEqDeriv(double t,
                        const StateDimVector & x,
                        const InputDimVector & u,
                        StateStateDimMatrix & Fx,
                        StateInputDimMatrix & Fu) const override
  {
    base_.calcStateEqDeriv(t, x, u, Fx, Fu);
  }
  void calcRunningCostDeriv(double t,
                            const StateDimVector & x,
                            const InputDimVector & u,
                            StateDimVector & Lx,
                            InputDimVector & Lu,
                            StateStateDimMatrix & Lxx,
                            InputInputDimMatrix & Luu,
                            StateInputDimMatrix & Lxu) const override
  {
    base_.calcRunningCostDeriv(t, x, u, Lx, Lu, Lxx, Luu, Lxu);
  }
  void calcTerminalCostDeriv(double t, const StateDimVector & x, StateDimVector & Vx, StateStateDimMatrix & Vxx)
      const override
  {
    base_.calcTerminalCostDeriv(t, x, Vx, Vxx);
  }
  void calcIneqConstDeriv(double,
                          const StateDimVector &,
                          const InputDimVector &,
                          IneqStateDimMatrix & C,
                          IneqInputDimMatrix & D) const override
  {
    // :236-249
    C.setZero();
    C(2, 0) = -1;
    C(3, 0) = 1;
    D.setZero();
    D(0, 0) = -1;
    D(1, 0) = 1;
  }

protected:
  DDPProblemCartPole base_;
};

/** Van der Pol oscillator, nmpc_fmpc/tests/src/TestFmpcOscillator.cpp:18-135.  params: [dt]. */
class FmpcProblemOscillator : public FmpcProblem<2, 1, 3>
{
public:
  static constexpr int kNumParams = 1;

  explicit FmpcProblemOscillator(const double * p) : FmpcProblem<2, 1, 3>(p[0]) {}

  static void defaultParams(double * p)
  {
    p[0] = 0.01; // :139
  }

  StateDimVector stateEq(double t, const StateDimVector & x, const InputDimVector & u) const override
  {
    return stateEq(t, x, u, dt_);
  }
  StateDimVector stateEq(double, const StateDimVector & x, const InputDimVector & u, double dt) const
  {
    // :28-36
    StateDimVector x_dot;
    x_dot[0] = (1.0 - std::pow(x[1], 2)) * x[0] - x[1] + u[0];
    x_dot[1] = x[0];
    StateDimVector x_next;
    for(int i = 0; i < 2; i++) x_next[i] = x[i] + dt * x_dot[i];
    return x_next;
  }
  double runningCost(double, const StateDimVector & x, const InputDimVector & u) const override
  {
    return 0.5 * (squaredNorm(x) + squaredNorm(u)); // :38-43
  }
  double terminalCost(double, const StateDimVector &) const override
  {
    return 0; // :45-50
  }
  IneqDimVector ineqConst(double, const StateDimVector & x, const InputDimVector & u) const override
  {
    // :52-61
    IneqDimVector g;
    g[0] = -1 * x[1] - 0.05;
    g[1] = -1 * u[0] - 1.0;
    g[2] = u[0] - 0.9;
    return g;
  }
  void calcStateEqDeriv(double,
                        const StateDimVector & x,
                        const InputDimVector &,
                        StateStateDimMatrix & Fx,
                        StateInputDimMatrix & Fu) const override
  {
    // :63-79
    Fx.setZero();
    Fx(0, 0) = 1.0 - std::pow(x[1], 2);
    Fx(0, 1) = -2 * x[0] * x[1] - 1.0;
    Fx(1, 0) = 1;
    for(int i = 0; i < 4; i++) Fx.d[i] *= dt_;
    for(int i = 0; i < 2; i++) Fx(i, i) += 1;
    Fu.setZero();
    Fu(0, 0) = 1;
    for(int i = 0; i < 2; i++) Fu.d[i] *= dt_;
  }
  void calcRunningCostDeriv(double,
                            const StateDimVector & x,
                            const InputDimVector & u,
                            StateDimVector & Lx,
                            InputDimVector & Lu,
                            StateStateDimMatrix & Lxx,
                            InputInputDimMatrix & Luu,
                            StateInputDimMatrix & Lxu) const override
  {
    // :81-105
    Lx = x;
    Lu = u;
    Lxx.setZero();
    for(int i = 0; i < 2; i++) Lxx(i, i) = 1;
    Luu(0, 0) = 1;
    Lxu.setZero();
  }
  void calcTerminalCostDeriv(double, const StateDimVector &, StateDimVector & Vx, StateStateDimMatrix & Vxx)
      const override
  {
    Vx.setZero(); // :107-121
    Vxx.setZero();
  }
  void calcIneqConstDeriv(double,
                          const StateDimVector &,
                          const InputDimVector &,
                          IneqStateDimMatrix & C,
                          IneqInputDimMatrix & D) const override
  {
    // :123-134
    C.setZero();
    C(0, 1) = -1;
    D.setZero();
    D(1, 0) = -1;
    D(2, 0) = 1;
  }
};
} // namespace oracle

// ---- quadrotor (BASELINE.json configs[3]; no counterpart in the reference) ----
// The functor of include/nmpc_b200/models/quadrotor.h instantiated in double is the fp64 reference for
// the fp32 device run; the DDP *solver* around it is still the restatement of this directory.
#include <nmpc_b200/models/quadrotor.h>

namespace oracle
{
class DDPProblemQuadrotor : public DDPProblem<12, 4>
{
public:
  using F = nmpc_b200::models::Quadrotor<double>;
  static constexpr int kNumParams = F::NUM_PARAMS;

  explicit DDPProblemQuadrotor(const double * p) : DDPProblem<12, 4>(p[0]), f_(F::fromParams(p)) {}
  static void defaultParams(double * p)
  {
    F::defaultParams(p);
  }

  template<class A, class B>
  static void copy(const A & a, B & b, int n)
  {
    for(int i = 0; i < n; i++) b.d[i] = a.d[i];
  }
  StateDimVector stateEq(double t, const StateDimVector & x, const InputDimVector & u) const override
  {
    F::StateDimVector fx, fn;
    F::InputDimVector fu;
    copy(x, fx, 12);
    copy(u, fu, 4);
    fn = f_.stateEq(t, fx, fu);
    StateDimVector out;
    copy(fn, out, 12);
    return out;
  }
  double runningCost(double t, const StateDimVector & x, const InputDimVector & u) const override
  {
    F::StateDimVector fx;
    F::InputDimVector fu;
    copy(x, fx, 12);
    copy(u, fu, 4);
    return f_.runningCost(t, fx, fu);
  }
  double terminalCost(double t, const StateDimVector & x) const override
  {
    F::StateDimVector fx;
    copy(x, fx, 12);
    return f_.terminalCost(t, fx);
  }
  void calcStateEqDeriv(double t,
                        const StateDimVector & x,
                        const InputDimVector & u,
                        StateStateDimMatrix & Fx,
                        StateInputDimMatrix & Fu) const override
  {
    F::StateDimVector fx;
    F::InputDimVector fu;
    F::StateStateDimMatrix a;
    F::StateInputDimMatrix b;
    copy(x, fx, 12);
    copy(u, fu, 4);
    f_.calcStateEqDeriv(t, fx, fu, a, b);
    copy(a, Fx, 144);
    copy(b, Fu, 48);
  }
  void calcRunningCostDeriv(double t,
                            const StateDimVector & x,
                            const InputDimVector & u,
                            StateDimVector & Lx,
                            InputDimVector & Lu,
                            StateStateDimMatrix & Lxx,
                            InputInputDimMatrix & Luu,
                            StateInputDimMatrix & Lxu) const override
  {
    F::StateDimVector fx, lx;
    F::InputDimVector fu, lu;
    F::StateStateDimMatrix lxx;
    F::InputInputDimMatrix luu;
    F::StateInputDimMatrix lxu;
    copy(x, fx, 12);
    copy(u, fu, 4);
    f_.calcRunningCostDeriv(t, fx, fu, lx, lu, lxx, luu, lxu);
    copy(lx, Lx, 12);
    copy(lu, Lu, 4);
    copy(lxx, Lxx, 144);
    copy(luu, Luu, 16);
    copy(lxu, Lxu, 48);
  }
  void calcTerminalCostDeriv(double t, const StateDimVector & x, StateDimVector & Vx, StateStateDimMatrix & Vxx)
      const override
  {
    F::StateDimVector fx, vx;
    F::StateStateDimMatrix vxx;
    copy(x, fx, 12);
    f_.calcTerminalCostDeriv(t, fx, vx, vxx);
    copy(vx, Vx, 12);
    copy(vxx, Vxx, 144);
  }

protected:
  F f_;
};
} // namespace oracle

// ---- vertical motion with a time-varying input dimension (TestDDPVerticalMotion.cpp:25-234) ----
// The functor of include/nmpc_b200/models/vertical_motion.h restates the problem bodies once for host and device;
// what pins both is tests/golden/vertical_*: outputs of the reference's DDPSolver<2, Eigen::Dynamic> (oracle/ref).
#include <nmpc_b200/models/centroidal_motion.h>
#include <nmpc_b200/models/vertical_motion.h>
#include <nmpc_b200/models/cartpole.h>
#include <nmpc_b200/models/planar_quadrotor.h>

namespace oracle
{
/** Any host+device functor F of include/nmpc_b200/models as an oracle problem. */
template<class F>
class DDPProblemFromFunctor : public DDPProblem<F::NX, F::NU>
{
public:
  using Base = DDPProblem<F::NX, F::NU>;
  using StateDimVector = typename Base::StateDimVector;
  using InputDimVector = typename Base::InputDimVector;
  using StateStateDimMatrix = typename Base::StateStateDimMatrix;
  using InputInputDimMatrix = typename Base::InputInputDimMatrix;
  using StateInputDimMatrix = typename Base::StateInputDimMatrix;
  static constexpr int kNumParams = F::NUM_PARAMS;
  static constexpr int NX = F::NX, NU = F::NU;

  explicit DDPProblemFromFunctor(const double * p) : Base(p[0]), f_(F::fromParams(p)) {}
  static void defaultParams(double * p)
  {
    F::defaultParams(p);
  }
  template<class A, class B>
  static void copy(const A & a, B & b, int n)
  {
    for(int i = 0; i < n; i++) b.d[i] = a.d[i];
  }
  template<class G, class = void>
  struct HasInputDim : std::false_type
  {
  };
  template<class G>
  struct HasInputDim<G, std::void_t<decltype(std::declval<const G &>().inputDim(0.0))>> : std::true_type
  {
  };
  int inputDim(double t) const override
  {
    if constexpr(HasInputDim<F>::value)
      return f_.inputDim(t);
    else
      return NU;
  }
  StateDimVector stateEq(double t, const StateDimVector & x, const InputDimVector & u) const override
  {
    typename F::StateDimVector fx, fn;
    typename F::InputDimVector fu;
    copy(x, fx, NX);
    copy(u, fu, NU);
    fn = f_.stateEq(t, fx, fu);
    StateDimVector out;
    copy(fn, out, NX);
    return out;
  }
  double runningCost(double t, const StateDimVector & x, const InputDimVector & u) const override
  {
    typename F::StateDimVector fx;
    typename F::InputDimVector fu;
    copy(x, fx, NX);
    copy(u, fu, NU);
    return f_.runningCost(t, fx, fu);
  }
  double terminalCost(double t, const StateDimVector & x) const override
  {
    typename F::StateDimVector fx;
    copy(x, fx, NX);
    return f_.terminalCost(t, fx);
  }
  void calcStateEqDeriv(double t, const StateDimVector & x, const InputDimVector & u, StateStateDimMatrix & Fx,
                        StateInputDimMatrix & Fu) const override
  {
    typename F::StateDimVector fx;
    typename F::InputDimVector fu;
    typename F::StateStateDimMatrix a;
    typename F::StateInputDimMatrix b;
    copy(x, fx, NX);
    copy(u, fu, NU);
    f_.calcStateEqDeriv(t, fx, fu, a, b);
    copy(a, Fx, NX * NX);
    copy(b, Fu, NX * NU);
  }
  void calcRunningCostDeriv(double t, const StateDimVector & x, const InputDimVector & u, StateDimVector & Lx,
                            InputDimVector & Lu, StateStateDimMatrix & Lxx, InputInputDimMatrix & Luu,
                            StateInputDimMatrix & Lxu) const override
  {
    typename F::StateDimVector fx, lx;
    typename F::InputDimVector fu, lu;
    typename F::StateStateDimMatrix lxx;
    typename F::InputInputDimMatrix luu;
    typename F::StateInputDimMatrix lxu;
    copy(x, fx, NX);
    copy(u, fu, NU);
    f_.calcRunningCostDeriv(t, fx, fu, lx, lu, lxx, luu, lxu);
    copy(lx, Lx, NX);
    copy(lu, Lu, NU);
    copy(lxx, Lxx, NX * NX);
    copy(luu, Luu, NU * NU);
    copy(lxu, Lxu, NX * NU);
  }
  void calcTerminalCostDeriv(double t, const StateDimVector & x, StateDimVector & Vx, StateStateDimMatrix & Vxx)
      const override
  {
    typename F::StateDimVector fx, vx;
    typename F::StateStateDimMatrix vxx;
    copy(x, fx, NX);
    f_.calcTerminalCostDeriv(t, fx, vx, vxx);
    copy(vx, Vx, NX);
    copy(vxx, Vxx, NX * NX);
  }

protected:
  F f_;
};
using DDPProblemVerticalMotion = DDPProblemFromFunctor<nmpc_b200::models::VerticalMotion<double>>;
// centroidal motion, n_x = 9, input dimension 16 or 0 (TestDDPCentroidalMotion.cpp:18-201); pinned by
// tests/golden/reference_centroidal.npz (the reference's DDPSolver<9, Eigen::Dynamic>, oracle/ref/ref_centroidal.cpp)
using DDPProblemCentroidalMotion = DDPProblemFromFunctor<nmpc_b200::models::CentroidalMotion<double>>;
// planar quadrotor, n_x = 6, n_u = 2: the two-input control-limited backward pass (BoxQP<2>, DDPSolver.hpp:450-497);
// pinned by tests/golden/reference_ddp_planar.npz (the reference's DDPSolver<6, 2>, oracle/ref/ref_ddp.cpp)
using DDPProblemPlanarQuadrotor = DDPProblemFromFunctor<nmpc_b200::models::PlanarQuadrotor<double>>;
/** Any host+device FMPC functor F of include/nmpc_b200/models as an oracle problem (ineqDim(t) when F has it). */
template<class F>
class FmpcProblemFromFunctor : public FmpcProblem<F::NX, F::NU, F::NG>
{
public:
  using Base = FmpcProblem<F::NX, F::NU, F::NG>;
  using StateDimVector = Vec<F::NX>;
  using InputDimVector = Vec<F::NU>;
  using IneqDimVector = Vec<F::NG>;
  using StateStateDimMatrix = Mat<F::NX, F::NX>;
  using InputInputDimMatrix = Mat<F::NU, F::NU>;
  using StateInputDimMatrix = Mat<F::NX, F::NU>;
  using IneqStateDimMatrix = Mat<F::NG, F::NX>;
  using IneqInputDimMatrix = Mat<F::NG, F::NU>;
  static constexpr int kNumParams = F::NUM_PARAMS;
  static constexpr int NX = F::NX, NU = F::NU, NG = F::NG;

  explicit FmpcProblemFromFunctor(const double * p) : Base(p[0]), f_(F::fromParams(p)) {}
  static void defaultParams(double * p)
  {
    F::defaultParams(p);
  }
  template<class A, class B>
  static void copy(const A & a, B & b, int n)
  {
    for(int i = 0; i < n; i++) b.d[i] = a.d[i];
  }
  template<class G, class = void>
  struct HasIneqDim : std::false_type
  {
  };
  template<class G>
  struct HasIneqDim<G, std::void_t<decltype(std::declval<const G &>().ineqDim(0.0))>> : std::true_type
  {
  };
  int ineqDim(double t) const override
  {
    if constexpr(HasIneqDim<F>::value)
      return f_.ineqDim(t);
    else
      return NG;
  }
  StateDimVector stateEq(double t, const StateDimVector & x, const InputDimVector & u) const override
  {
    typename F::StateDimVector fx, fn;
    typename F::InputDimVector fu;
    copy(x, fx, NX);
    copy(u, fu, NU);
    fn = f_.stateEq(t, fx, fu);
    StateDimVector out;
    copy(fn, out, NX);
    return out;
  }
  double runningCost(double t, const StateDimVector & x, const InputDimVector & u) const override
  {
    typename F::StateDimVector fx;
    typename F::InputDimVector fu;
    copy(x, fx, NX);
    copy(u, fu, NU);
    return f_.runningCost(t, fx, fu);
  }
  double terminalCost(double t, const StateDimVector & x) const override
  {
    typename F::StateDimVector fx;
    copy(x, fx, NX);
    return f_.terminalCost(t, fx);
  }
  IneqDimVector ineqConst(double t, const StateDimVector & x, const InputDimVector & u) const override
  {
    typename F::StateDimVector fx;
    typename F::InputDimVector fu;
    copy(x, fx, NX);
    copy(u, fu, NU);
    const typename F::IneqDimVector g = f_.ineqConst(t, fx, fu);
    IneqDimVector out;
    copy(g, out, NG);
    return out;
  }
  void calcStateEqDeriv(double t, const StateDimVector & x, const InputDimVector & u, StateStateDimMatrix & Fx,
                        StateInputDimMatrix & Fu) const override
  {
    typename F::StateDimVector fx;
    typename F::InputDimVector fu;
    typename F::StateStateDimMatrix a;
    typename F::StateInputDimMatrix b;
    copy(x, fx, NX);
    copy(u, fu, NU);
    f_.calcStateEqDeriv(t, fx, fu, a, b);
    copy(a, Fx, NX * NX);
    copy(b, Fu, NX * NU);
  }
  void calcRunningCostDeriv(double t, const StateDimVector & x, const InputDimVector & u, StateDimVector & Lx,
                            InputDimVector & Lu, StateStateDimMatrix & Lxx, InputInputDimMatrix & Luu,
                            StateInputDimMatrix & Lxu) const override
  {
    typename F::StateDimVector fx, lx;
    typename F::InputDimVector fu, lu;
    typename F::StateStateDimMatrix lxx;
    typename F::InputInputDimMatrix luu;
    typename F::StateInputDimMatrix lxu;
    copy(x, fx, NX);
    copy(u, fu, NU);
    f_.calcRunningCostDeriv(t, fx, fu, lx, lu, lxx, luu, lxu);
    copy(lx, Lx, NX);
    copy(lu, Lu, NU);
    copy(lxx, Lxx, NX * NX);
    copy(luu, Luu, NU * NU);
    copy(lxu, Lxu, NX * NU);
  }
  void calcTerminalCostDeriv(double t, const StateDimVector & x, StateDimVector & Vx, StateStateDimMatrix & Vxx)
      const override
  {
    typename F::StateDimVector fx, vx;
    typename F::StateStateDimMatrix vxx;
    copy(x, fx, NX);
    f_.calcTerminalCostDeriv(t, fx, vx, vxx);
    copy(vx, Vx, NX);
    copy(vxx, Vxx, NX * NX);
  }
  void calcIneqConstDeriv(double t, const StateDimVector & x, const InputDimVector & u, IneqStateDimMatrix & C,
                          IneqInputDimMatrix & D) const override
  {
    typename F::StateDimVector fx;
    typename F::InputDimVector fu;
    typename F::IneqStateDimMatrix c;
    typename F::IneqInputDimMatrix d;
    copy(x, fx, NX);
    copy(u, fu, NU);
    f_.calcIneqConstDeriv(t, fx, fu, c, d);
    copy(c, C, NG * NX);
    copy(d, D, NG * NU);
  }

protected:
  F f_;
};
// two inputs (pivoted LDLT / FullPivLU gain solve) and a time-varying inequality dimension; both pinned by golden
// vectors of the reference's own FmpcSolver (oracle/ref/ref_fmpc.cpp)
using FmpcProblemPlanarQuadrotor = FmpcProblemFromFunctor<nmpc_b200::models::PlanarQuadrotor<double>>;
using FmpcProblemCartPoleWindowed = FmpcProblemFromFunctor<nmpc_b200::models::CartPoleWindowed<double>>;
} // namespace oracle
