// TEST INFRASTRUCTURE ONLY -- builds oracle/_ref/libnmpc_ref.so: the REFERENCE's own solver headers
// (straight from /root/reference, unmodified) compiled against oracle/ref/eigen_shim, with the problem
// classes of the reference's tests restated in the tests' own Eigen idioms (ref_models.h).  Used to pin
// oracle/ (the plain restatement) against the reference's real control flow, and to generate
// tests/golden/*.npz.  DDPSolver.hpp and FmpcSolver.hpp each define calcDuration() in an anonymous
// namespace, so the two solvers live in separate translation units.
#pragma once

#include <cmath>
#include <stdexcept>
#include <vector>

#include <Eigen/Dense>

/** nmpc_ddp/tests/src/TestDDPCartPole.cpp:28-234 (bodies as in the test; parameters from a flat array). */
template<class Base>
class CartPoleBodies : public Base
{
public:
  using StateDimVector = typename Base::StateDimVector;
  using InputDimVector = typename Base::InputDimVector;
  using StateStateDimMatrix = typename Base::StateStateDimMatrix;
  using InputInputDimMatrix = typename Base::InputInputDimMatrix;
  using StateInputDimMatrix = typename Base::StateInputDimMatrix;

  explicit CartPoleBodies(const double * p) : Base(p[0])
  {
    cart_mass = p[1], pole_mass = p[2], pole_length = p[3];
    running_x << p[4], p[5], p[6], p[7];
    running_u << p[8];
    terminal_x << p[9], p[10], p[11], p[12];
    ref_pos = p[13];
  }

  StateDimVector stateEq(double t, const StateDimVector & x, const InputDimVector & u) const override
  {
    return stateEq(t, x, u, this->dt_);
  }
  StateDimVector stateEq(double, const StateDimVector & x, const InputDimVector & u, double dt) const
  {
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
    return x + dt * x_dot;
  }
  double runningCost(double, const StateDimVector & x, const InputDimVector & u) const override
  {
    StateDimVector ref_x;
    ref_x << ref_pos, 0, 0, 0;
    return 0.5 * running_x.dot((x - ref_x).cwiseAbs2()) + 0.5 * running_u.dot(u.cwiseAbs2());
  }
  double terminalCost(double, const StateDimVector & x) const override
  {
    StateDimVector ref_x;
    ref_x << ref_pos, 0, 0, 0;
    return 0.5 * terminal_x.dot((x - ref_x).cwiseAbs2());
  }
  void calcStateEqDeriv(double,
                        const StateDimVector & x,
                        const InputDimVector & u,
                        Eigen::Ref<StateStateDimMatrix> state_eq_deriv_x,
                        Eigen::Ref<StateInputDimMatrix> state_eq_deriv_u) const override
  {
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
    state_eq_deriv_x.setZero();
    state_eq_deriv_x(0, 2) = 1;
    state_eq_deriv_x(1, 3) = 1;
    state_eq_deriv_x(2, 1) = ((-1 * m2 * l * omega2 * cos_theta + m2 * g_ * (1 - 2 * std::pow(sin_theta, 2))) * denom
                              + -1 * (f - m2 * l * omega2 * sin_theta + m2 * g_ * sin_theta * cos_theta)
                                    * (2 * m2 * sin_theta * cos_theta))
                             / std::pow(denom, 2);
    state_eq_deriv_x(2, 3) = (-2 * m2 * l * omega * sin_theta) / denom;
    state_eq_deriv_x(3, 1) =
        ((-1 * f * sin_theta + -1 * m2 * l * omega2 * (1 - 2 * std::pow(sin_theta, 2)) + g_ * (m1 + m2) * cos_theta)
             * denom
         + -1 * (f * cos_theta - m2 * l * omega2 * sin_theta * cos_theta + g_ * (m1 + m2) * sin_theta)
               * (2 * m2 * sin_theta * cos_theta))
        / (l * std::pow(denom, 2));
    state_eq_deriv_x(3, 3) = (-2 * m2 * l * omega * sin_theta * cos_theta) / (l * denom);
    state_eq_deriv_x *= this->dt_;
    state_eq_deriv_x.diagonal().array() += 1.0;
    state_eq_deriv_u.setZero();
    state_eq_deriv_u[2] = 1 / denom;
    state_eq_deriv_u[3] = cos_theta / (l * denom);
    state_eq_deriv_u *= this->dt_;
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
  void calcRunningCostDeriv(double,
                            const StateDimVector & x,
                            const InputDimVector & u,
                            Eigen::Ref<StateDimVector> running_cost_deriv_x,
                            Eigen::Ref<InputDimVector> running_cost_deriv_u) const override
  {
    StateDimVector ref_x;
    ref_x << ref_pos, 0, 0, 0;
    running_cost_deriv_x = running_x.cwiseProduct(x - ref_x);
    running_cost_deriv_u = running_u.cwiseProduct(u);
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
    running_cost_deriv_xx = running_x.asDiagonal();
    running_cost_deriv_uu = running_u.asDiagonal();
    running_cost_deriv_xu.setZero();
  }
  void calcTerminalCostDeriv(double,
                             const StateDimVector & x,
                             Eigen::Ref<StateDimVector> terminal_cost_deriv_x) const override
  {
    StateDimVector ref_x;
    ref_x << ref_pos, 0, 0, 0;
    terminal_cost_deriv_x = terminal_x.cwiseProduct(x - ref_x);
  }
  void calcTerminalCostDeriv(double t,
                             const StateDimVector & x,
                             Eigen::Ref<StateDimVector> terminal_cost_deriv_x,
                             Eigen::Ref<StateStateDimMatrix> terminal_cost_deriv_xx) const override
  {
    calcTerminalCostDeriv(t, x, terminal_cost_deriv_x);
    terminal_cost_deriv_xx = terminal_x.asDiagonal();
  }

  static constexpr double g_ = 9.80665;
  double cart_mass, pole_mass, pole_length, ref_pos;
  StateDimVector running_x, terminal_x;
  InputDimVector running_u;
};

/** The planar quadrotor of include/nmpc_b200/models/planar_quadrotor.h (two inputs, four thrust limits) written as a
    user of the reference would: an nmpc_fmpc::FmpcProblem<6, 2, 4> in Eigen idiom.  Not one of the reference's tests --
    it exists to drive the n_u > 1 branch of FmpcSolver::backwardPass (LDLT / FullPivLU of G, FmpcSolver.hpp:596-617).
    params: [dt, mass, inertia, arm, thrust_max, running_x[6], running_u, running_u_cross, terminal_x[6], ref_px, ref_pz]. */
template<class Base>
class PlanarQuadrotorBodies : public Base
{
public:
  using StateDimVector = typename Base::StateDimVector;
  using InputDimVector = typename Base::InputDimVector;
  using StateStateDimMatrix = typename Base::StateStateDimMatrix;
  using InputInputDimMatrix = typename Base::InputInputDimMatrix;
  using StateInputDimMatrix = typename Base::StateInputDimMatrix;
  explicit PlanarQuadrotorBodies(const double * p) : Base(p[0])
  {
    mass_ = p[1];
    inertia_ = p[2];
    arm_ = p[3];
    thrust_max_ = p[4];
    for(int i = 0; i < 6; i++) running_x_[i] = p[5 + i];
    running_u_ = p[11];
    running_u_cross_ = p[12];
    for(int i = 0; i < 6; i++) terminal_x_[i] = p[13 + i];
    ref_.setZero();
    ref_[0] = p[19];
    ref_[1] = p[20];
  }
  double hover() const
  {
    return 0.5 * mass_ * 9.80665;
  }
  StateDimVector stateEq(double, const StateDimVector & x, const InputDimVector & u) const override
  {
    const double st = std::sin(x[2]), ct = std::cos(x[2]);
    const double thrust = u[0] + u[1];
    StateDimVector x_dot;
    x_dot << x[3], x[4], x[5], -1 * thrust * st / mass_, thrust * ct / mass_ - 9.80665, arm_ * (u[0] - u[1]) / inertia_;
    return x + this->dt_ * x_dot;
  }
  double runningCost(double, const StateDimVector & x, const InputDimVector & u) const override
  {
    double sx = 0;
    for(int i = 0; i < 6; i++)
    {
      const double e = x[i] - ref_[i];
      sx += running_x_[i] * (e * e);
    }
    const double e0 = u[0] - hover(), e1 = u[1] - hover();
    const double su = running_u_ * (e0 * e0 + e1 * e1);
    return (0.5 * sx + 0.5 * su) + running_u_cross_ * (e0 * e1);
  }
  double terminalCost(double, const StateDimVector & x) const override
  {
    double sx = 0;
    for(int i = 0; i < 6; i++)
    {
      const double e = x[i] - ref_[i];
      sx += terminal_x_[i] * (e * e);
    }
    return 0.5 * sx;
  }
  void calcStateEqDeriv(double,
                        const StateDimVector & x,
                        const InputDimVector & u,
                        Eigen::Ref<StateStateDimMatrix> state_eq_deriv_x,
                        Eigen::Ref<StateInputDimMatrix> state_eq_deriv_u) const override
  {
    const double st = std::sin(x[2]), ct = std::cos(x[2]);
    const double thrust = u[0] + u[1];
    state_eq_deriv_x.setZero();
    state_eq_deriv_x(0, 3) = 1;
    state_eq_deriv_x(1, 4) = 1;
    state_eq_deriv_x(2, 5) = 1;
    state_eq_deriv_x(3, 2) = -1 * thrust * ct / mass_;
    state_eq_deriv_x(4, 2) = -1 * thrust * st / mass_;
    state_eq_deriv_x *= this->dt_;
    state_eq_deriv_x.diagonal().array() += 1;
    state_eq_deriv_u.setZero();
    state_eq_deriv_u(3, 0) = -1 * st / mass_;
    state_eq_deriv_u(3, 1) = -1 * st / mass_;
    state_eq_deriv_u(4, 0) = ct / mass_;
    state_eq_deriv_u(4, 1) = ct / mass_;
    state_eq_deriv_u(5, 0) = arm_ / inertia_;
    state_eq_deriv_u(5, 1) = -1 * arm_ / inertia_;
    state_eq_deriv_u *= this->dt_;
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
  void calcRunningCostDeriv(double,
                            const StateDimVector & x,
                            const InputDimVector & u,
                            Eigen::Ref<StateDimVector> running_cost_deriv_x,
                            Eigen::Ref<InputDimVector> running_cost_deriv_u) const override
  {
    for(int i = 0; i < 6; i++) running_cost_deriv_x[i] = running_x_[i] * (x[i] - ref_[i]);
    const double e0 = u[0] - hover(), e1 = u[1] - hover();
    running_cost_deriv_u[0] = running_u_ * e0 + running_u_cross_ * e1;
    running_cost_deriv_u[1] = running_u_ * e1 + running_u_cross_ * e0;
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
    running_cost_deriv_xx.setZero();
    for(int i = 0; i < 6; i++) running_cost_deriv_xx(i, i) = running_x_[i];
    running_cost_deriv_uu(0, 0) = running_u_;
    running_cost_deriv_uu(1, 1) = running_u_;
    running_cost_deriv_uu(0, 1) = running_u_cross_;
    running_cost_deriv_uu(1, 0) = running_u_cross_;
    running_cost_deriv_xu.setZero();
  }
  void calcTerminalCostDeriv(double, const StateDimVector & x, Eigen::Ref<StateDimVector> terminal_cost_deriv_x)
      const override
  {
    for(int i = 0; i < 6; i++) terminal_cost_deriv_x[i] = terminal_x_[i] * (x[i] - ref_[i]);
  }
  void calcTerminalCostDeriv(double t,
                             const StateDimVector & x,
                             Eigen::Ref<StateDimVector> terminal_cost_deriv_x,
                             Eigen::Ref<StateStateDimMatrix> terminal_cost_deriv_xx) const override
  {
    calcTerminalCostDeriv(t, x, terminal_cost_deriv_x);
    terminal_cost_deriv_xx.setZero();
    for(int i = 0; i < 6; i++) terminal_cost_deriv_xx(i, i) = terminal_x_[i];
  }
protected:
  double mass_, inertia_, arm_, thrust_max_, running_x_[6], running_u_, running_u_cross_, terminal_x_[6];
  StateDimVector ref_;
};
