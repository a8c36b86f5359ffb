/* nmpc_b200 -- nmpc_fmpc::FmpcProblem facade (reference: nmpc_fmpc/include/nmpc_fmpc/FmpcProblem.h:14-134).
 * Adds ineqDim(), ineqConst() and calcIneqConstDeriv() to nmpc_ddp::DDPProblem; see
 * include/nmpc_ddp/DDPProblem.h for how a problem is bound to its device functor. */
#pragma once

#include <type_traits>
#include <utility>

#include <nmpc_ddp/DDPProblem.h>

namespace nmpc_b200
{
/** Does functor F declare a time-varying inequality dimension, `int ineqDim(t)`? */
template<class F, class = void>
struct HasIneqDimMethod : std::false_type
{
};
template<class F>
struct HasIneqDimMethod<F, std::void_t<decltype(std::declval<const F &>().ineqDim(0.0))>> : std::true_type
{
};
} // namespace nmpc_b200

namespace nmpc_fmpc
{
template<int StateDim, int InputDim, int IneqDim>
class FmpcProblem : public nmpc_ddp::DDPProblem<StateDim, InputDim>
{
public:
  using StateDimVector = typename nmpc_ddp::DDPProblem<StateDim, InputDim>::StateDimVector;
  using InputDimVector = typename nmpc_ddp::DDPProblem<StateDim, InputDim>::InputDimVector;
  using IneqDimVector = nmpc_b200::Matrix<double, IneqDim, 1>;
  using StateStateDimMatrix = typename nmpc_ddp::DDPProblem<StateDim, InputDim>::StateStateDimMatrix;
  using InputInputDimMatrix = typename nmpc_ddp::DDPProblem<StateDim, InputDim>::InputInputDimMatrix;
  using StateInputDimMatrix = typename nmpc_ddp::DDPProblem<StateDim, InputDim>::StateInputDimMatrix;
  using InputStateDimMatrix = typename nmpc_ddp::DDPProblem<StateDim, InputDim>::InputStateDimMatrix;
  using IneqStateDimMatrix = nmpc_b200::Matrix<double, IneqDim, StateDim>;
  using IneqInputDimMatrix = nmpc_b200::Matrix<double, IneqDim, InputDim>;

public:
  FmpcProblem(double dt) : nmpc_ddp::DDPProblem<StateDim, InputDim>(dt)
  {
    static_assert(IneqDim >= 0, "[FMPC] Template param IneqDim is the LARGEST inequality dimension of the problem.");
  }

  /** The (largest) inequality dimension. */
  inline virtual int ineqDim() const
  {
    return IneqDim;
  }
  /** The inequality dimension at time t (FmpcProblem.h:74-86).  The reference spells a time-varying dimension
      FmpcProblem<S, I, Eigen::Dynamic> with this method overridden; a kernel needs compile-time sizes, so here IneqDim
      is the largest dimension and the override returns how many LEADING rows of ineqConst / calcIneqConstDeriv are
      constraints at t.  The solver keeps the other rows neutral (s = 1, nu = 0, C = D = 0): every active quantity
      equals the reference's reduced-size result (tests/test_fmpc_extra.py, golden vectors of the reference's own
      FmpcSolver<4, 1, Eigen::Dynamic>). */
  inline virtual int ineqDim(double // t
  ) const
  {
    return ineqDim();
  }

  /** Inequality constraints; feasible iff <= 0 (FmpcProblem.h:88-94). */
  virtual IneqDimVector ineqConst(double t, const StateDimVector & x, const InputDimVector & u) const = 0;
  virtual void calcIneqConstDeriv(double t,
                                  const StateDimVector & x,
                                  const InputDimVector & u,
                                  IneqStateDimMatrix & ineq_const_deriv_x,
                                  IneqInputDimMatrix & ineq_const_deriv_u) const = 0;
};

/** FmpcProblem whose virtuals all forward to a (host+device) functor F registered under `name`. */
template<class F>
class FunctorProblem : public FmpcProblem<F::NX, F::NU, F::NG>
{
public:
  using Base = FmpcProblem<F::NX, F::NU, F::NG>;
  using typename Base::IneqDimVector;
  using typename Base::IneqInputDimMatrix;
  using typename Base::IneqStateDimMatrix;
  using typename Base::InputDimVector;
  using typename Base::InputInputDimMatrix;
  using typename Base::StateDimVector;
  using typename Base::StateInputDimMatrix;
  using typename Base::StateStateDimMatrix;

  FunctorProblem(const std::string & name, const std::vector<double> & params)
  : Base(params.at(0)), functor_(F::fromParams(params.data())), binding_{name, params}
  {
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
  int ineqDim(double t) const override
  {
    if constexpr(nmpc_b200::HasIneqDimMethod<F>::value)
      return functor_.ineqDim(t);
    else
      return F::NG;
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
  IneqDimVector ineqConst(double t, const StateDimVector & x, const InputDimVector & u) const override
  {
    return functor_.ineqConst(t, x, u);
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
  void calcIneqConstDeriv(double t,
                          const StateDimVector & x,
                          const InputDimVector & u,
                          IneqStateDimMatrix & ineq_const_deriv_x,
                          IneqInputDimMatrix & ineq_const_deriv_u) const override
  {
    functor_.calcIneqConstDeriv(t, x, u, ineq_const_deriv_x, ineq_const_deriv_u);
  }
  nmpc_b200::DeviceFunctorBinding deviceFunctor() const override
  {
    return binding_;
  }

protected:
  F functor_;
  nmpc_b200::DeviceFunctorBinding binding_;
};
} // namespace nmpc_fmpc
