// TEST INFRASTRUCTURE ONLY -- CPU oracle for nmpc_b200 (see oracle/README.md).
//
// Plain C++ restatement of nmpc_ddp::DDPSolver<StateDim, InputDim> (fixed dimensions), one
// function per reference function, same statement order:
//   solve         /root/reference/nmpc_ddp/include/nmpc_ddp/DDPSolver.hpp:27-141
//   procOnce      /root/reference/nmpc_ddp/include/nmpc_ddp/DDPSolver.hpp:144-340
//   backwardPass  /root/reference/nmpc_ddp/include/nmpc_ddp/DDPSolver.hpp:343-534
//   forwardPass   /root/reference/nmpc_ddp/include/nmpc_ddp/DDPSolver.hpp:537-560
//   Configuration /root/reference/nmpc_ddp/include/nmpc_ddp/DDPSolver.h:47-110 (defaults)
//   problem API   /root/reference/nmpc_ddp/include/nmpc_ddp/DDPProblem.h:99-198
// The dense arithmetic is Eigen's in the reference; see linalg.hpp for what is restated.
// Matrix products are associated left to right, as Eigen evaluates `A.transpose() * V * B`.
#pragma once

#include <array>
#include <cmath>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "boxqp_oracle.hpp"
#include "linalg.hpp"

namespace oracle
{
/** Mirror of nmpc_ddp::DDPProblem (DDPProblem.h:15-203); only the overloads the solvers call. */
template<int NX, int NU>
class DDPProblem
{
public:
  using StateDimVector = Vec<NX>;
  using InputDimVector = Vec<NU>;
  using StateStateDimMatrix = Mat<NX, NX>;
  using InputInputDimMatrix = Mat<NU, NU>;
  using StateInputDimMatrix = Mat<NX, NU>;
  using InputStateDimMatrix = Mat<NU, NX>;

  explicit DDPProblem(double dt) : dt_(dt) {}
  virtual ~DDPProblem() = default;

  static constexpr int stateDim()
  {
    return NX;
  }
  static constexpr int inputDim()
  {
    return NU;
  }
  double dt() const
  {
    return dt_;
  }
  /** DDPProblem::inputDim(t) (DDPProblem.h:72-85).  The reference sizes every per-step vector and matrix by it
      (DDPProblem<StateDim, Eigen::Dynamic>); this restatement keeps fixed sizes NU = the largest dimension and treats
      inputs a >= inputDim(t) as padding: zero in the input sequence, decoupled in the linearisation (Fu(:,a) = 0,
      Lu(a) = 0, Lxu(:,a) = 0, Luu(a,:) = Luu(:,a) = e_a), which yields the reduced system's gains and value function
      for the active inputs and exactly zero gains for the padding.  Pinned against the reference's own Dynamic code
      path by tests/golden (vertical_*). */
  virtual int inputDim(double /* t */) const
  {
    return NU;
  }

  virtual StateDimVector stateEq(double t, const StateDimVector & x, const InputDimVector & u) const = 0; // :99
  virtual double runningCost(double t, const StateDimVector & x, const InputDimVector & u) const = 0; // :107
  virtual double terminalCost(double t, const StateDimVector & x) const = 0; // :114
  virtual void calcStateEqDeriv(double t,
                                const StateDimVector & x,
                                const InputDimVector & u,
                                StateStateDimMatrix & Fx,
                                StateInputDimMatrix & Fu) const = 0; // :123-127
  virtual void calcRunningCostDeriv(double t,
                                    const StateDimVector & x,
                                    const InputDimVector & u,
                                    StateDimVector & Lx,
                                    InputDimVector & Lu,
                                    StateStateDimMatrix & Lxx,
                                    InputInputDimMatrix & Luu,
                                    StateInputDimMatrix & Lxu) const = 0; // :171-178
  virtual void calcTerminalCostDeriv(double t,
                                     const StateDimVector & x,
                                     StateDimVector & Vx,
                                     StateStateDimMatrix & Vxx) const = 0; // :195-198

protected:
  const double dt_;
};

template<int NX, int NU>
class DDPSolver
{
public:
  using StateDimVector = Vec<NX>;
  using InputDimVector = Vec<NU>;
  using StateStateDimMatrix = Mat<NX, NX>;
  using InputInputDimMatrix = Mat<NU, NU>;
  using StateInputDimMatrix = Mat<NX, NU>;
  using InputStateDimMatrix = Mat<NU, NX>;

  struct Configuration // DDPSolver.h:47-110
  {
    Configuration()
    {
      // alpha_list = 10^linspace(0, -3, 11)  (DDPSolver.h:53-59)
      int list_size = 11;
      alpha_list.resize(list_size);
      for(int i = 0; i < list_size; i++)
      {
        // Eigen::VectorXd::LinSpaced(size, low, high)[i] = low + i * (high - low) / (size - 1)
        double e = 0.0 + i * ((-3.0 - 0.0) / (list_size - 1));
        if(i == list_size - 1) e = -3.0;
        alpha_list[i] = std::pow(10, e);
      }
    }
    int print_level = 1;
    bool use_state_eq_second_derivative = false;
    bool with_input_constraint = false;
    int max_iter = 500;
    int horizon_steps = 100;
    int reg_type = 1;
    double initial_lambda = 1e-4;
    double initial_dlambda = 1.0;
    double lambda_factor = 1.6;
    double lambda_min = 1e-6;
    double lambda_max = 1e10;
    double k_rel_norm_thre = 1e-4;
    double lambda_thre = 1e-5;
    std::vector<double> alpha_list;
    double cost_update_ratio_thre = 0;
    double cost_update_thre = 1e-7;
  };

  struct ControlData // DDPSolver.h:113-123
  {
    std::vector<StateDimVector> x_list;
    std::vector<InputDimVector> u_list;
    std::vector<double> cost_list;
    double costSum() const
    {
      double s = 0.0;
      for(double c : cost_list) s += c;
      return s;
    }
  };

  struct Derivative // DDPSolver.h:126-176 (first-order dynamics only; second order throws in the reference)
  {
    StateStateDimMatrix Fx;
    StateInputDimMatrix Fu;
    StateDimVector Lx;
    InputDimVector Lu;
    StateStateDimMatrix Lxx;
    InputInputDimMatrix Luu;
    StateInputDimMatrix Lxu;
  };

  struct TraceData // DDPSolver.h:179-216 (durations omitted: wall-clock, not part of parity)
  {
    int iter = 0;
    double cost = 0;
    double lambda = 0;
    double dlambda = 0;
    double alpha = 0;
    double k_rel_norm = 0;
    double cost_update_actual = 0;
    double cost_update_expected = 0;
    double cost_update_ratio = 0;
  };

  explicit DDPSolver(const std::shared_ptr<DDPProblem<NX, NU>> & problem) : problem_(problem) {}

  Configuration & config()
  {
    return config_;
  }
  const ControlData & controlData() const
  {
    return control_data_;
  }
  const std::vector<TraceData> & traceDataList() const
  {
    return trace_data_list_;
  }
  void setInputLimitsFunc(const std::function<std::array<InputDimVector, 2>(double)> & f)
  {
    input_limits_func_ = f;
  }

  // ---- counters the reference does not keep (used for roofline accounting only) ----
  int n_forward_pass = 0; //!< forwardPass() calls over the whole solve
  int n_backward_pass = 0; //!< backwardPass() calls over the whole solve (incl. failed sweeps)
  int retval_last = 0; //!< last procOnce return value (-1/0/1)

  bool solve(double current_t, const StateDimVector & current_x, const std::vector<InputDimVector> & initial_u_list)
  {
    // DDPSolver.hpp:36-38
    current_t_ = current_t;
    lambda_ = config_.initial_lambda;
    dlambda_ = config_.initial_dlambda;
    n_forward_pass = 0;
    n_backward_pass = 0;

    // :41-45
    if(static_cast<int>(initial_u_list.size()) != config_.horizon_steps)
    {
      throw std::invalid_argument("initial_u_list length should be " + std::to_string(config_.horizon_steps) + " but "
                                  + std::to_string(initial_u_list.size()) + ".");
    }

    // :61-80
    const int N = config_.horizon_steps;
    candidate_control_data_.x_list.resize(N + 1);
    candidate_control_data_.u_list.resize(N);
    candidate_control_data_.cost_list.resize(N + 1);
    derivative_list_.resize(N);
    k_list_.resize(N);
    K_list_.resize(N);

    // :83-95 initial rollout
    control_data_.u_list = initial_u_list;
    control_data_.x_list.resize(N + 1);
    control_data_.cost_list.resize(N + 1);
    control_data_.x_list[0] = current_x;
    for(int i = 0; i < N; i++)
    {
      double t = current_t_ + i * problem_->dt();
      for(int a = problem_->inputDim(t); a < NU; a++) control_data_.u_list[i][a] = 0.0; // padding inputs
      control_data_.x_list[i + 1] = problem_->stateEq(t, control_data_.x_list[i], control_data_.u_list[i]);
      control_data_.cost_list[i] = problem_->runningCost(t, control_data_.x_list[i], control_data_.u_list[i]);
    }
    double terminal_t = current_t_ + N * problem_->dt();
    control_data_.cost_list[N] = problem_->terminalCost(terminal_t, control_data_.x_list[N]);

    // :98-104
    trace_data_list_.clear();
    TraceData initial_trace_data;
    initial_trace_data.iter = 0;
    initial_trace_data.cost = control_data_.costSum();
    initial_trace_data.lambda = lambda_;
    initial_trace_data.dlambda = dlambda_;
    trace_data_list_.push_back(initial_trace_data);

    // :115-123
    int retval = 0;
    for(int iter = 1; iter <= config_.max_iter; iter++)
    {
      retval = procOnce(iter);
      if(retval != 0)
      {
        break;
      }
    }
    retval_last = retval;
    return retval == 1; // :140
  }

  // exposed for kernel-level parity tests
  const std::vector<InputDimVector> & kList() const
  {
    return k_list_;
  }
  const std::vector<InputStateDimMatrix> & KList() const
  {
    return K_list_;
  }
  const std::vector<Derivative> & derivativeList() const
  {
    return derivative_list_;
  }

protected:
  int procOnce(int iter)
  {
    // :152-154
    trace_data_list_.push_back(TraceData());
    const size_t trace_idx = trace_data_list_.size() - 1;
    trace_data_list_[trace_idx].iter = iter;
    const int N = config_.horizon_steps;

    // Step 1 (:157-185)
    for(int i = 0; i < N; i++)
    {
      auto & derivative = derivative_list_[i];
      double t = current_t_ + i * problem_->dt();
      const StateDimVector & x = control_data_.x_list[i];
      const InputDimVector & u = control_data_.u_list[i];
      if(config_.use_state_eq_second_derivative)
      {
        // the reference reaches the throw in backwardPass (:393); nothing second-order is ever used
        throw std::runtime_error("Vector-tensor product is not implemented yet.");
      }
      problem_->calcStateEqDeriv(t, x, u, derivative.Fx, derivative.Fu);
      problem_->calcRunningCostDeriv(t, x, u, derivative.Lx, derivative.Lu, derivative.Lxx, derivative.Luu,
                                     derivative.Lxu);
      for(int a = problem_->inputDim(t); a < NU; a++) // decouple the padding inputs (see DDPProblem::inputDim)
      {
        for(int r = 0; r < NX; r++)
        {
          derivative.Fu(r, a) = 0.0;
          derivative.Lxu(r, a) = 0.0;
        }
        derivative.Lu[a] = 0.0;
        for(int c = 0; c < NU; c++)
        {
          derivative.Luu(a, c) = 0.0;
          derivative.Luu(c, a) = 0.0;
        }
        derivative.Luu(a, a) = 1.0;
      }
    }
    double terminal_t = current_t_ + N * problem_->dt();
    problem_->calcTerminalCostDeriv(terminal_t, control_data_.x_list[N], last_Vx_, last_Vxx_);

    // Step 2 (:188-214)
    while(!backwardPass())
    {
      dlambda_ = std::max(dlambda_ * config_.lambda_factor, config_.lambda_factor);
      lambda_ = std::max(lambda_ * dlambda_, config_.lambda_min);
      if(lambda_ > config_.lambda_max)
      {
        return -1;
      }
    }

    // small-gradient termination (:217-231)
    double k_rel_norm = 0;
    for(int i = 0; i < N; i++)
    {
      k_rel_norm = std::max(k_rel_norm, norm(k_list_[i]) / (norm(control_data_.u_list[i]) + 1.0));
    }
    trace_data_list_[trace_idx].k_rel_norm = k_rel_norm;
    if(k_rel_norm < config_.k_rel_norm_thre && lambda_ < config_.lambda_thre)
    {
      return 1;
    }

    // Step 3 (:234-274)
    bool forward_pass_success = false;
    double cost_update_actual = 0;
    {
      double alpha = 0;
      double cost_update_expected = 0;
      double cost_update_ratio = 0;
      for(size_t i = 0; i < config_.alpha_list.size(); i++)
      {
        alpha = config_.alpha_list[i];

        forwardPass(alpha);

        cost_update_actual = control_data_.costSum() - candidate_control_data_.costSum();
        cost_update_expected = -1 * alpha * (dV_[0] + alpha * dV_[1]);
        cost_update_ratio = cost_update_actual / cost_update_expected;
        if(cost_update_expected < 0)
        {
          cost_update_ratio = (cost_update_actual >= 0 ? 1 : -1);
        }
        if(cost_update_ratio > config_.cost_update_ratio_thre)
        {
          forward_pass_success = true;
          break;
        }
      }
      trace_data_list_[trace_idx].alpha = alpha;
      trace_data_list_[trace_idx].cost_update_actual = cost_update_actual;
      trace_data_list_[trace_idx].cost_update_expected = cost_update_expected;
      trace_data_list_[trace_idx].cost_update_ratio = cost_update_ratio;
    }

    // Step 4 (:280-333)
    int retval = 0;
    if(forward_pass_success)
    {
      control_data_.x_list = candidate_control_data_.x_list;
      control_data_.u_list = candidate_control_data_.u_list;
      control_data_.cost_list = candidate_control_data_.cost_list;

      if(cost_update_actual < config_.cost_update_thre)
      {
        retval = 1;
      }

      dlambda_ = std::min(dlambda_ / config_.lambda_factor, 1 / config_.lambda_factor);
      if(lambda_ >= config_.lambda_min)
      {
        lambda_ *= dlambda_;
      }
      else
      {
        lambda_ = 0;
      }
    }
    else
    {
      dlambda_ = std::max(dlambda_ * config_.lambda_factor, config_.lambda_factor);
      lambda_ = std::max(lambda_ * dlambda_, config_.lambda_min);
      if(lambda_ > config_.lambda_max)
      {
        retval = -1;
      }
    }

    // :335-337
    trace_data_list_[trace_idx].cost = control_data_.costSum();
    trace_data_list_[trace_idx].lambda = lambda_;
    trace_data_list_[trace_idx].dlambda = dlambda_;

    return retval;
  }

  bool backwardPass()
  {
    n_backward_pass++;
    const int N = config_.horizon_steps;

    // :346-363
    StateDimVector Vx = last_Vx_;
    StateStateDimMatrix Vxx = last_Vxx_;
    StateStateDimMatrix Vxx_reg;

    InputDimVector Qu;
    StateDimVector Qx;
    InputStateDimMatrix Qux;
    InputInputDimMatrix Quu;
    StateStateDimMatrix Qxx;
    InputStateDimMatrix Qux_reg;
    InputInputDimMatrix Quu_F;

    InputDimVector k;
    InputStateDimMatrix K;

    dV_[0] = 0.0; // :365
    dV_[1] = 0.0;

    for(int i = N - 1; i >= 0; i--)
    {
      // :370-381
      double t = current_t_ + i * problem_->dt();
      const StateStateDimMatrix & Fx = derivative_list_[i].Fx;
      const StateInputDimMatrix & Fu = derivative_list_[i].Fu;
      const StateDimVector & Lx = derivative_list_[i].Lx;
      const InputDimVector & Lu = derivative_list_[i].Lu;
      const StateStateDimMatrix & Lxx = derivative_list_[i].Lxx;
      const InputInputDimMatrix & Luu = derivative_list_[i].Luu;
      const StateInputDimMatrix & Lxu = derivative_list_[i].Lxu;

      // Q (:386-408)
      Qu = add(Lu, mulT(Fu, Vx));
      Qx = add(Lx, mulT(Fx, Vx));
      Qux = add(transpose(Lxu), mul(mulT(Fu, Vxx), Fx));
      Quu = add(Luu, mul(mulT(Fu, Vxx), Fu));
      Qxx = add(Lxx, mul(mulT(Fx, Vxx), Fx));

      // regularisation (:421-441)
      Vxx_reg = Vxx;
      if(config_.reg_type == 2)
      {
        for(int d = 0; d < NX; d++) Vxx_reg(d, d) += lambda_;
      }
      Qux_reg = add(transpose(Lxu), mul(mulT(Fu, Vxx_reg), Fx));
      Quu_F = add(Luu, mul(mulT(Fu, Vxx_reg), Fu));
      if(config_.reg_type == 1)
      {
        for(int d = 0; d < NU; d++) Quu_F(d, d) += lambda_;
      }

      // gains (:448-517)
      if(NU > 0)
      {
        if(config_.with_input_constraint)
        {
          // :452-467 warm start from the next step's feedforward term
          InputDimVector initial_k;
          if(i == N - 1)
          {
            initial_k.setZero();
          }
          else if(problem_->inputDim(t) == problem_->inputDim(t + problem_->dt()))
          {
            initial_k = k_list_[i + 1]; // k_list_[i + 1].size() == input_dim (:459)
          }
          else
          {
            initial_k.setZero(); // the next step has another input dimension (:463-466)
          }

          // :469-480 a fresh default-configured BoxQP per step
          BoxQP<NU> qp;
          const auto u_limits = input_limits_func_(t);
          k = qp.solve(Quu_F, Qu, sub(u_limits[0], control_data_.u_list[i]), sub(u_limits[1], control_data_.u_list[i]),
                       initial_k);
          if(qp.retval < 0)
          {
            return false;
          }

          // :482-496 feedback gain only on the free dimensions
          K.setZero();
          if(qp.n_free > 0)
          {
            for(int c = 0; c < NX; c++)
            {
              double rhs[NU > 0 ? NU : 1];
              for(int j = 0; j < qp.n_free; j++) rhs[j] = Qux_reg(qp.free_idxs[j], c);
              lltSolveInPlace(qp.llt_free, qp.n_free, qp.n_free, rhs);
              for(int j = 0; j < qp.n_free; j++) K(qp.free_idxs[j], c) = -1 * rhs[j];
            }
          }
        }
        else
        {
          // :500-510
          InputInputDimMatrix llt = Quu_F;
          if(!lltInPlace(llt.d, NU, NU))
          {
            return false;
          }
          k = Qu;
          lltSolveInPlace(llt.d, NU, NU, k.d);
          for(int d = 0; d < NU; d++) k[d] = -1 * k[d];
          K = Qux_reg;
          for(int c = 0; c < NX; c++)
          {
            lltSolveInPlace(llt.d, NU, NU, &K.d[c * NU]);
            for(int d = 0; d < NU; d++) K(d, c) = -1 * K(d, c);
          }
        }
      }

      // cost-to-go (:522-526)
      dV_[0] += dot(k, Qu);
      dV_[1] += 0.5 * dot(k, mul(Quu, k));
      Vec<NX> Vx_new = add(add(add(Qx, mul(mulT(K, Quu), k)), mulT(K, Qu)), mulT(Qux, k));
      StateStateDimMatrix Vxx_new = add(add(add(Qxx, mul(mulT(K, Quu), K)), mulT(K, Qux)), mulT(Qux, K));
      Vx = Vx_new;
      for(int c = 0; c < NX; c++)
        for(int r = 0; r < NX; r++) Vxx(r, c) = 0.5 * (Vxx_new(r, c) + Vxx_new(c, r));

      // :529-530
      k_list_[i] = k;
      K_list_[i] = K;
    }
    return true;
  }

  void forwardPass(double alpha)
  {
    n_forward_pass++;
    const int N = config_.horizon_steps;
    candidate_control_data_.x_list[0] = control_data_.x_list[0]; // :540
    for(int i = 0; i < N; i++)
    {
      // :545-546
      StateDimVector dx = sub(candidate_control_data_.x_list[i], control_data_.x_list[i]);
      InputDimVector Kdx = mul(K_list_[i], dx);
      for(int d = 0; d < NU; d++)
      {
        candidate_control_data_.u_list[i][d] = control_data_.u_list[i][d] + alpha * k_list_[i][d] + Kdx[d];
      }
      // :551-555
      double t = current_t_ + i * problem_->dt();
      candidate_control_data_.x_list[i + 1] =
          problem_->stateEq(t, candidate_control_data_.x_list[i], candidate_control_data_.u_list[i]);
      candidate_control_data_.cost_list[i] =
          problem_->runningCost(t, candidate_control_data_.x_list[i], candidate_control_data_.u_list[i]);
    }
    double terminal_t = current_t_ + N * problem_->dt(); // :557-559
    candidate_control_data_.cost_list[N] = problem_->terminalCost(terminal_t, candidate_control_data_.x_list[N]);
  }

protected:
  Configuration config_;
  std::shared_ptr<DDPProblem<NX, NU>> problem_;
  std::vector<TraceData> trace_data_list_;
  std::function<std::array<InputDimVector, 2>(double)> input_limits_func_;
  double current_t_ = 0;
  double lambda_ = 0;
  double dlambda_ = 0;
  ControlData control_data_;
  ControlData candidate_control_data_;
  std::vector<InputDimVector> k_list_;
  std::vector<InputStateDimMatrix> K_list_;
  std::vector<Derivative> derivative_list_;
  StateDimVector last_Vx_;
  StateStateDimMatrix last_Vxx_;
  double dV_[2] = {0, 0};
};
} // namespace oracle
