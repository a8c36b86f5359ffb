/* nmpc_b200 -- nmpc_ddp::DDPProblem facade.
 *
 * Same class template, typedef names, constructor and method set as the reference's
 * nmpc_ddp/include/nmpc_ddp/DDPProblem.h:15-203, with two differences forced by device execution:
 *  - the fixed-size types are nmpc_b200::Matrix (Eigen is not a dependency), so the output arguments
 *    are plain references instead of Eigen::Ref;
 *  - a problem that is to be solved on the GPU must name its device functor (deviceFunctor()): host
 *    virtual methods cannot run inside a kernel.  FunctorProblem<F> below derives every virtual from a
 *    functor type F, so a problem is written once and is usable from host code exactly like a reference
 *    problem (e.g. the plant simulation problem->stateEq(...) of TestDDPCartPole.cpp:330).
 */
#pragma once

#include <stdexcept>
#include <string>
#include <vector>

#include <nmpc_b200/matrix.h>

namespace nmpc_b200
{
/** Name of a registered device functor plus its flat parameter vector (see c_api.h). */
struct DeviceFunctorBinding
{
  std::string name;
  std::vector<double> params;
};
} // namespace nmpc_b200

namespace nmpc_ddp
{
template<int StateDim, int InputDim>
class DDPProblem
{
public:
  using StateDimVector = nmpc_b200::Matrix<double, StateDim, 1>;
  using InputDimVector = nmpc_b200::Matrix<double, InputDim, 1>;
  using StateStateDimMatrix = nmpc_b200::Matrix<double, StateDim, StateDim>;
  using InputInputDimMatrix = nmpc_b200::Matrix<double, InputDim, InputDim>;
  using StateInputDimMatrix = nmpc_b200::Matrix<double, StateDim, InputDim>;
  using InputStateDimMatrix = nmpc_b200::Matrix<double, InputDim, StateDim>;

public:
  DDPProblem(double dt) : dt_(dt)
  {
    static_assert(StateDim > 0, "[DDP] Template param StateDim should be positive.");
    static_assert(InputDim >= 0, "[DDP] Template param InputDim should be non-negative: a time-varying input dimension is declared as the "
                                 "LARGEST dimension plus inputDim(t) (see nmpc_b200/models/vertical_motion.h).");
  }
  virtual ~DDPProblem() = default;

  static inline constexpr int stateDim()
  {
    return StateDim;
  }
  inline virtual int inputDim() const
  {
    return InputDim;
  }
  inline virtual int inputDim(double // t
  ) const
  {
    return inputDim();
  }
  inline double dt() const
  {
    return dt_;
  }

  virtual StateDimVector stateEq(double t, const StateDimVector & x, const InputDimVector & u) const = 0;
  virtual double runningCost(double t, const StateDimVector & x, const InputDimVector & u) const = 0;
  virtual double terminalCost(double t, const StateDimVector & x) const = 0;
  virtual void calcStateEqDeriv(double t,
                                const StateDimVector & x,
                                const InputDimVector & u,
                                StateStateDimMatrix & state_eq_deriv_x,
                                StateInputDimMatrix & state_eq_deriv_u) const = 0;
  virtual void calcRunningCostDeriv(double t,
                                    const StateDimVector & x,
                                    const InputDimVector & u,
                                    StateDimVector & running_cost_deriv_x,
                                    InputDimVector & running_cost_deriv_u,
                                    StateStateDimMatrix & running_cost_deriv_xx,
                                    InputInputDimMatrix & running_cost_deriv_uu,
                                    StateInputDimMatrix & running_cost_deriv_xu) const = 0;
  virtual void calcTerminalCostDeriv(double t,
                                     const StateDimVector & x,
                                     StateDimVector & terminal_cost_deriv_x,
                                     StateStateDimMatrix & terminal_cost_deriv_xx) const = 0;

  /** Device functor that implements this problem on the GPU.  A problem without one cannot be solved:
      there is no CPU fallback. */
  virtual nmpc_b200::DeviceFunctorBinding deviceFunctor() const
  {
    throw std::runtime_error("[nmpc_b200] this DDPProblem names no device functor; derive from "
                             "nmpc_ddp::FunctorProblem<F> or override deviceFunctor()");
  }

protected:
  const double dt_ = 0;
};

/** DDPProblem whose virtuals all forward to a (host+device) functor F registered under `name`. */
template<class F>
class FunctorProblem : public DDPProblem<F::NX, F::NU>
{
public:
  using Base = DDPProblem<F::NX, F::NU>;
  using typename Base::InputDimVector;
  using typename Base::InputInputDimMatrix;
  using typename Base::StateDimVector;
  using typename Base::StateInputDimMatrix;
  using typename Base::StateStateDimMatrix;

  FunctorProblem(const std::string & name, const std::vector<double> & params)
  : Base(params.at(0)), functor_(F::fromParams(params.data())), binding_{name, params}
  {
    if(static_cast<int>(params.size()) != F::NUM_PARAMS)
      throw std::invalid_argument("[nmpc_b200] functor '" + name + "' expects " + std::to_string(F::NUM_PARAMS)
                                  + " parameters");
  }
  explicit FunctorProblem(const std::string & name) : FunctorProblem(name, defaults()) {}

  static std::vector<double> defaults()
  {
    std::vector<double> p(F::NUM_PARAMS);
    F::defaultParams(p.data());
    return p;
  }
  const F & functor() const
  {
    return functor_;
  }

  using Base::inputDim;
  /** DDPProblem::inputDim(t) (reference DDPProblem.h:72-85): the functor's own inputDim(t) when it declares a
      time-varying input dimension (inputs a >= inputDim(t) are padding), InputDim otherwise. */
  int inputDim(double t) const override
  {
    return inputDimOf(functor_, t, 0);
  }

  StateDimVector stateEq(double t, const StateDimVector & x, const InputDimVector & u) const override
  {
    return functor_.stateEq(t, x, u);
  }
  double runningCost(double t, const StateDimVector & x, const InputDimVector & u) const override
  {
    return functor_.runningCost(t, x, u);
  }
  double terminalCost(double t, const StateDimVector & x) const override
  {
    return functor_.terminalCost(t, x);
  }
  void calcStateEqDeriv(double t,
                        const StateDimVector & x,
                        const InputDimVector & u,
                        StateStateDimMatrix & state_eq_deriv_x,
                        StateInputDimMatrix & state_eq_deriv_u) const override
  {
    functor_.calcStateEqDeriv(t, x, u, state_eq_deriv_x, state_eq_deriv_u);
  }
  void calcRunningCostDeriv(double t,
                            const StateDimVector & x,
                            const InputDimVector & u,
                            StateDimVector & running_cost_deriv_x,
                            InputDimVector & running_cost_deriv_u,
                            StateStateDimMatrix & running_cost_deriv_xx,
                            InputInputDimMatrix & running_cost_deriv_uu,
                            StateInputDimMatrix & running_cost_deriv_xu) const override
  {
    functor_.calcRunningCostDeriv(t, x, u, running_cost_deriv_x, running_cost_deriv_u, running_cost_deriv_xx,
                                  running_cost_deriv_uu, running_cost_deriv_xu);
  }
  void calcTerminalCostDeriv(double t,
                             const StateDimVector & x,
                             StateDimVector & terminal_cost_deriv_x,
                             StateStateDimMatrix & terminal_cost_deriv_xx) const override
  {
    functor_.calcTerminalCostDeriv(t, x, terminal_cost_deriv_x, terminal_cost_deriv_xx);
  }
  nmpc_b200::DeviceFunctorBinding deviceFunctor() const override
  {
    return binding_;
  }

protected:
  template<class G>
  static auto inputDimOf(const G & g, double t, int) -> decltype(g.inputDim(t))
  {
    return g.inputDim(t);
  }
  template<class G>
  static int inputDimOf(const G &, double, long)
  {
    return F::NU;
  }

  F functor_;
  nmpc_b200::DeviceFunctorBinding binding_;
};
} // namespace nmpc_ddp
