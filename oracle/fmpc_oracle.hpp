// TEST INFRASTRUCTURE ONLY -- CPU oracle for nmpc_b200 (see oracle/README.md).
//
// Plain C++ restatement of nmpc_fmpc::FmpcSolver<StateDim, InputDim, IneqDim> (fixed dimensions):
//   Variable / Coefficient  /root/reference/nmpc_fmpc/include/nmpc_fmpc/FmpcSolver.h:117-229
//   solve            /root/reference/nmpc_fmpc/include/nmpc_fmpc/FmpcSolver.hpp:158-257
//   checkVariable    .hpp:285-362
//   procOnce         .hpp:365-493
//   calcKktError     .hpp:496-521
//   backwardPass     .hpp:524-665
//   forwardPass      .hpp:668-708
//   updateVariables  .hpp:711-834
//   setupMeritFunc   .hpp:837-933, calcMeritFunc .hpp:936-982
//   l1NormDirectionalDeriv  /root/reference/nmpc_fmpc/include/nmpc_fmpc/MathUtils.h:17-38
//   problem API      /root/reference/nmpc_fmpc/include/nmpc_fmpc/FmpcProblem.h:88-107
// Models: nmpc_fmpc/tests/src/TestFmpcCartPole.cpp:32-256, TestFmpcOscillator.cpp:18-135.
#pragma once

#include <algorithm>
#include <cmath>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "ddp_oracle.hpp"
#include "linalg.hpp"

namespace oracle
{
template<int NX, int NU, int NG>
class FmpcProblem : public DDPProblem<NX, NU>
{
public:
  using IneqDimVector = Vec<NG>;
  using IneqStateDimMatrix = Mat<NG, NX>;
  using IneqInputDimMatrix = Mat<NG, NU>;

  explicit FmpcProblem(double dt) : DDPProblem<NX, NU>(dt) {}

  static constexpr int ineqDim()
  {
    return NG;
  }
  /** FmpcProblem<.., Eigen::Dynamic> (FmpcProblem.h:62-86): the inequality dimension at time t, at most NG.  Rows
      j >= ineqDim(t) of a step are PADDING: the solver keeps s = 1, nu = 0 there and treats g + s, C, D as zero,
      which leaves every active quantity at the value the reference computes with vectors of size ineqDim(t)
      (pinned by tests/golden/reference_fmpc_dynamic.npz, produced by the reference's FmpcSolver<4, 1, Dynamic>). */
  virtual int ineqDim(double) const
  {
    return NG;
  }
  virtual IneqDimVector ineqConst(double t, const Vec<NX> & x, const Vec<NU> & u) const = 0; // FmpcProblem.h:94
  virtual void calcIneqConstDeriv(double t,
                                  const Vec<NX> & x,
                                  const Vec<NU> & u,
                                  IneqStateDimMatrix & C,
                                  IneqInputDimMatrix & D) const = 0; // FmpcProblem.h:103-107
};

/** Eigen::LDLT<Matrix<n,n>>::compute + info(), restated (diagonal pivoting, Eigen 3.4
    internal::ldlt_inplace<Lower>::unblocked): info is NumericalIssue only when a valid pivot
    follows a zero pivot; NaN pivots count as "zero" and do not fail by themselves. */
struct LdltFactor
{
  static constexpr int MAXN = 16;
  double a[MAXN * MAXN];
  int perm[MAXN];
  int n = 0;
  bool success = true;

  void compute(const double * a_in, int n_)
  {
    n = n_;
    for(int j = 0; j < n; j++)
      for(int i = 0; i < n; i++) a[i + j * n] = a_in[i + j * n];
    for(int i = 0; i < n; i++) perm[i] = i;
    bool found_zero_pivot = false;
    success = true;
    if(n <= 1)
    {
      return; // Eigen: size <= 1 => Success unconditionally
    }
    for(int k = 0; k < n; k++)
    {
      int piv = k;
      double best = std::fabs(a[k + k * n]);
      for(int i = k + 1; i < n; i++)
      {
        double v = std::fabs(a[i + i * n]);
        if(v > best)
        {
          best = v;
          piv = i;
        }
      }
      if(piv != k)
      {
        // Eigen swaps within the lower triangle only; a full symmetric swap of a symmetric
        // working copy yields the same lower triangle
        for(int j = 0; j < n; j++) std::swap(a[k + j * n], a[piv + j * n]);
        for(int i = 0; i < n; i++) std::swap(a[i + k * n], a[i + piv * n]);
        std::swap(perm[k], perm[piv]);
      }
      // A(k,k) -= A10 * (D0 * A10^T);  A21 -= A20 * (D0 * A10^T);  A21 /= A(k,k)
      double dk = a[k + k * n];
      for(int j = 0; j < k; j++) dk -= a[k + j * n] * (a[j + j * n] * a[k + j * n]);
      a[k + k * n] = dk;
      for(int i = k + 1; i < n; i++)
      {
        double s = a[i + k * n];
        for(int j = 0; j < k; j++) s -= a[i + j * n] * (a[j + j * n] * a[k + j * n]);
        a[i + k * n] = s;
      }
      bool pivot_is_valid = (std::fabs(dk) > 0.0);
      if(pivot_is_valid)
      {
        for(int i = k + 1; i < n; i++) a[i + k * n] /= dk;
      }
      else
      {
        for(int i = k + 1; i < n; i++) success = success && (a[i + k * n] == 0.0);
      }
      if(found_zero_pivot && pivot_is_valid)
        success = false;
      else if(!pivot_is_valid)
        found_zero_pivot = true;
    }
  }

  /** x = A^-1 b with Eigen::LDLT::solve's pseudo-inverse of D (|d| <= DBL_MIN => 0). */
  void solveInPlace(double * b) const
  {
    double y[MAXN];
    for(int i = 0; i < n; i++) y[i] = b[perm[i]];
    for(int i = 0; i < n; i++)
      for(int j = 0; j < i; j++) y[i] -= a[i + j * n] * y[j];
    const double tol = std::numeric_limits<double>::min();
    for(int i = 0; i < n; i++)
    {
      double dk = a[i + i * n];
      if(std::fabs(dk) > tol)
        y[i] /= dk;
      else
        y[i] = 0.0;
    }
    for(int i = n - 1; i >= 0; i--)
      for(int j = i + 1; j < n; j++) y[i] -= a[j + i * n] * y[j];
    for(int i = 0; i < n; i++) b[perm[i]] = y[i];
  }
};

/** Eigen::FullPivLU<Matrix<n,n>> (Eigen 3.4: complete pivoting, first maximum of |a| in column-major order of the
    trailing block; solve() keeps the leading `rank` pivots, rank = #{|pivot| > epsilon * n * max|pivot|}, the other
    solution components are zero).  The fallback of FmpcSolver.hpp:614-616. */
struct FullPivLuFactor
{
  static constexpr int MAXN = 16;
  double a[MAXN * MAXN];
  int row_tr[MAXN];
  int col_perm[MAXN];
  int n = 0;
  int rank = 0;

  void compute(const double * a_in, int n_)
  {
    n = n_;
    for(int j = 0; j < n; j++)
      for(int i = 0; i < n; i++) a[i + j * n] = a_in[i + j * n];
    for(int i = 0; i < n; i++)
    {
      row_tr[i] = i;
      col_perm[i] = i;
    }
    double maxpivot = 0.0;
    int nonzero = n;
    for(int k = 0; k < n; k++)
    {
      int pr = k, pc = k;
      double best = -1.0;
      for(int j = k; j < n; j++)
        for(int i = k; i < n; i++)
        {
          const double v = std::fabs(a[i + j * n]);
          if(v > best)
          {
            best = v;
            pr = i;
            pc = j;
          }
        }
      if(best == 0.0)
      {
        nonzero = k;
        break;
      }
      if(best > maxpivot) maxpivot = best;
      row_tr[k] = pr;
      if(pr != k)
        for(int j = 0; j < n; j++) std::swap(a[k + j * n], a[pr + j * n]);
      if(pc != k)
      {
        for(int i = 0; i < n; i++) std::swap(a[i + k * n], a[i + pc * n]);
        std::swap(col_perm[k], col_perm[pc]);
      }
      const double p = a[k + k * n];
      for(int i = k + 1; i < n; i++) a[i + k * n] /= p;
      for(int j = k + 1; j < n; j++)
        for(int i = k + 1; i < n; i++) a[i + j * n] -= a[i + k * n] * a[k + j * n];
    }
    const double thr = std::numeric_limits<double>::epsilon() * n * maxpivot;
    rank = 0;
    for(int i = 0; i < nonzero; i++)
      if(std::fabs(a[i + i * n]) > thr) rank++;
  }

  void solveInPlace(double * b) const
  {
    double c[MAXN];
    for(int i = 0; i < n; i++) c[i] = b[i];
    for(int k = 0; k < n; k++)
      if(row_tr[k] != k) std::swap(c[k], c[row_tr[k]]);
    for(int k = 0; k < n; k++)
      for(int i = k + 1; i < n; i++) c[i] -= a[i + k * n] * c[k];
    for(int i = rank - 1; i >= 0; i--)
    {
      double s = c[i];
      for(int j = i + 1; j < rank; j++) s -= a[i + j * n] * c[j];
      c[i] = s / a[i + i * n];
    }
    for(int i = 0; i < n; i++) b[col_perm[i]] = (i < rank) ? c[i] : 0.0;
  }
};

/** MathUtils.h:17-38 */
template<int IN, int OUT>
inline double l1NormDirectionalDeriv(const Vec<OUT> & func, const Mat<OUT, IN> & jac, const Vec<IN> & dir)
{
  double deriv = 0.0;
  for(int i = 0; i < OUT; i++)
  {
    double jd = 0.0;
    for(int j = 0; j < IN; j++) jd += jac(i, j) * dir[j];
    if(func[i] > 0)
      deriv += jd;
    else if(func[i] < 0)
      deriv += -1 * jd;
    else
      deriv += std::fabs(jd);
  }
  return deriv;
}

template<int NX, int NU, int NG>
class FmpcSolver
{
public:
  using StateDimVector = Vec<NX>;
  using InputDimVector = Vec<NU>;
  using IneqDimVector = Vec<NG>;
  using StateStateDimMatrix = Mat<NX, NX>;
  using InputInputDimMatrix = Mat<NU, NU>;
  using StateInputDimMatrix = Mat<NX, NU>;
  using InputStateDimMatrix = Mat<NU, NX>;
  using IneqStateDimMatrix = Mat<NG, NX>;
  using IneqInputDimMatrix = Mat<NG, NU>;

  struct Configuration // FmpcSolver.h:58-89
  {
    int print_level = 1;
    int horizon_steps = 100;
    int max_iter = 10;
    double kkt_error_thre = 1e-4;
    bool check_nan = true;
    bool init_complementary_variable = false;
    bool update_barrier_eps = true;
    bool break_if_llt_fails = false;
    bool enable_line_search = false;
    bool merit_const_scale_from_lagrange_multipliers = false;
  };

  enum class Status // FmpcSolver.h:92-114
  {
    Uninitialized = 0,
    Succeeded = 1,
    ErrorInForward = 2,
    ErrorInBackward = 3,
    ErrorInUpdate = 4,
    MaxIterationReached = 5,
    IterationContinued = 6
  };

  struct Variable // FmpcSolver.h:117-158
  {
    explicit Variable(int _horizon_steps = 0) : horizon_steps(_horizon_steps)
    {
      x_list.resize(horizon_steps + 1);
      u_list.resize(horizon_steps);
      lambda_list.resize(horizon_steps + 1);
      s_list.resize(horizon_steps);
      nu_list.resize(horizon_steps);
    }
    void reset(double _x, double _u, double _lambda, double _s, double _nu)
    {
      for(auto & x : x_list) x.setConstant(_x);
      for(auto & u : u_list) u.setConstant(_u);
      for(auto & l : lambda_list) l.setConstant(_lambda);
      for(auto & s : s_list) s.setConstant(_s);
      for(auto & nu : nu_list) nu.setConstant(_nu);
    }
    bool containsNaN() const
    {
      for(auto & x : x_list)
        if(hasNaNOrInf(x)) return true;
      for(auto & u : u_list)
        if(hasNaNOrInf(u)) return true;
      for(auto & l : lambda_list)
        if(hasNaNOrInf(l)) return true;
      for(auto & s : s_list)
        if(hasNaNOrInf(s)) return true;
      for(auto & nu : nu_list)
        if(hasNaNOrInf(nu)) return true;
      return false;
    }
    int horizon_steps;
    std::vector<StateDimVector> x_list;
    std::vector<InputDimVector> u_list;
    std::vector<StateDimVector> lambda_list;
    std::vector<IneqDimVector> s_list;
    std::vector<IneqDimVector> nu_list;
  };

  struct Coefficient // FmpcSolver.h:161-229
  {
    bool terminal = false;
    StateStateDimMatrix A;
    StateInputDimMatrix B;
    IneqStateDimMatrix C;
    IneqInputDimMatrix D;
    StateDimVector Lx;
    InputDimVector Lu;
    StateStateDimMatrix Lxx;
    InputInputDimMatrix Luu;
    StateInputDimMatrix Lxu;
    StateDimVector x_bar;
    IneqDimVector g_bar;
    StateDimVector Lx_bar;
    InputDimVector Lu_bar;
    InputDimVector k;
    InputStateDimMatrix K;
    StateDimVector s;
    StateStateDimMatrix P;

    bool containsNaN() const // .hpp:133-155 (the terminal entry only holds Lx, Lxx, Lx_bar, s, P)
    {
      if(!terminal)
      {
        if(hasNaNOrInf(A) || hasNaNOrInf(B) || hasNaNOrInf(C) || hasNaNOrInf(D)) return true;
        if(hasNaNOrInf(Lu) || hasNaNOrInf(Luu) || hasNaNOrInf(Lxu)) return true;
        if(hasNaNOrInf(x_bar) || hasNaNOrInf(g_bar) || hasNaNOrInf(Lu_bar)) return true;
        if(hasNaNOrInf(k) || hasNaNOrInf(K)) return true;
      }
      if(hasNaNOrInf(Lx) || hasNaNOrInf(Lxx) || hasNaNOrInf(Lx_bar)) return true;
      if(hasNaNOrInf(s) || hasNaNOrInf(P)) return true;
      return false;
    }
  };

  struct TraceData // FmpcSolver.h:232-251 (durations omitted)
  {
    int iter = 0;
    double kkt_error = 0;
    // extras the reference does not record (for per-iteration parity diagnostics)
    double barrier_eps = 0;
    double alpha_s = 0;
    double alpha_nu = 0;
  };

  explicit FmpcSolver(const std::shared_ptr<FmpcProblem<NX, NU, NG>> & problem) : problem_(problem) {}

  Configuration & config()
  {
    return config_;
  }
  const Variable & variable() const
  {
    return variable_;
  }
  const std::vector<Coefficient> & coeffList() const
  {
    return coeff_list_;
  }
  const std::vector<TraceData> & traceDataList() const
  {
    return trace_data_list_;
  }
  double & barrierEps()
  {
    return barrier_eps_;
  }

  Status solve(double current_t, const StateDimVector & current_x, const Variable & initial_variable)
  {
    // .hpp:166-169
    current_t_ = current_t;
    current_x_ = current_x;
    variable_ = initial_variable;
    const int N = config_.horizon_steps;

    // .hpp:172-188
    if(config_.init_complementary_variable)
    {
      constexpr double initial_barrier_eps = 1e-4;
      constexpr double complementary_variable_margin_rate = 1e-2;
      constexpr double complementary_variable_min = 1e-2;

      barrier_eps_ = initial_barrier_eps;
      for(int i = 0; i < N; i++)
      {
        double t = current_t_ + i * problem_->dt();
        IneqDimVector g = problem_->ineqConst(t, variable_.x_list[i], variable_.u_list[i]);
        for(int j = 0; j < NG; j++)
        {
          variable_.s_list[i][j] =
              (1.0 + complementary_variable_margin_rate) * std::fmax(-1 * g[j], complementary_variable_min);
          variable_.nu_list[i][j] = (1.0 + complementary_variable_margin_rate)
                                    * std::fmax(barrier_eps_ * (1.0 / variable_.s_list[i][j]), complementary_variable_min);
        }
      }
    }

    for(int i = 0; i < N; i++) // padding rows of a time-varying inequality dimension
    {
      const int ng_i = problem_->ineqDim(current_t_ + i * problem_->dt());
      for(int j = ng_i; j < NG; j++)
      {
        variable_.s_list[i][j] = 1.0;
        variable_.nu_list[i][j] = 0.0;
      }
    }

    checkVariable(); // .hpp:191

    if(delta_variable_.horizon_steps != N) // .hpp:194-198
    {
      delta_variable_ = Variable(N);
    }

    // .hpp:201-220: N stage entries (resize truncates the previous terminal entry) + 1 terminal entry
    coeff_list_.resize(N);
    coeff_list_.emplace_back();
    coeff_list_[N].terminal = true;

    trace_data_list_.clear(); // .hpp:223

    // .hpp:232-244
    Status status = Status::Uninitialized;
    for(int iter = 1; iter <= config_.max_iter; iter++)
    {
      status = procOnce(iter);
      if(status != Status::IterationContinued)
      {
        break;
      }
    }
    if(status == Status::IterationContinued)
    {
      status = Status::MaxIterationReached;
    }
    return status;
  }

protected:
  void checkVariable() const
  {
    const int N = config_.horizon_steps;
    // .hpp:288-312
    if(static_cast<int>(variable_.x_list.size()) != N + 1)
      throw std::invalid_argument("[FMPC] x_list length should be " + std::to_string(N + 1) + " but "
                                  + std::to_string(variable_.x_list.size()) + ".");
    if(static_cast<int>(variable_.u_list.size()) != N)
      throw std::invalid_argument("[FMPC] u_list length should be " + std::to_string(N) + " but "
                                  + std::to_string(variable_.u_list.size()) + ".");
    if(static_cast<int>(variable_.lambda_list.size()) != N + 1)
      throw std::invalid_argument("[FMPC] lambda_list length should be " + std::to_string(N + 1) + " but "
                                  + std::to_string(variable_.lambda_list.size()) + ".");
    if(static_cast<int>(variable_.s_list.size()) != N)
      throw std::invalid_argument("[FMPC] s_list length should be " + std::to_string(N) + " but "
                                  + std::to_string(variable_.s_list.size()) + ".");
    if(static_cast<int>(variable_.nu_list.size()) != N)
      throw std::invalid_argument("[FMPC] nu_list length should be " + std::to_string(N) + " but "
                                  + std::to_string(variable_.nu_list.size()) + ".");
    // .hpp:348-361
    for(int i = 0; i < N; i++)
    {
      double t = current_t_ + i * problem_->dt();
      for(int j = 0; j < NG; j++)
      {
        if(variable_.s_list[i][j] < 0)
          throw std::runtime_error("[FMPC] s_list[i] must be non-negative. i: " + std::to_string(i)
                                   + ", time: " + std::to_string(t));
      }
      for(int j = 0; j < NG; j++)
      {
        if(variable_.nu_list[i][j] < 0)
          throw std::runtime_error("[FMPC] nu_list[i] must be non-negative. i: " + std::to_string(i)
                                   + ", time: " + std::to_string(t));
      }
    }
  }

  Status procOnce(int iter)
  {
    const int N = config_.horizon_steps;
    trace_data_list_.emplace_back(); // .hpp:373-375
    const size_t trace_idx = trace_data_list_.size() - 1;
    trace_data_list_[trace_idx].iter = iter;

    // barrier parameter (.hpp:378-399)
    if(config_.update_barrier_eps)
    {
      double s_nu_ave = 0.0;
      int total_ineq_dim = 0;
      for(int i = 0; i < N; i++)
      {
        s_nu_ave += dot(variable_.s_list[i], variable_.nu_list[i]);
        total_ineq_dim += problem_->ineqDim(current_t_ + i * problem_->dt()); // s_list[i].size()
      }
      s_nu_ave /= total_ineq_dim;

      double sigma = 0.5;
      constexpr double barrier_eps_min = 1e-8;
      constexpr double barrier_eps_max = 1e6;
      barrier_eps_ = std::clamp(sigma * s_nu_ave, barrier_eps_min, barrier_eps_max);
    }
    trace_data_list_[trace_idx].barrier_eps = barrier_eps_;

    // Step 1 (.hpp:402-440)
    {
      double dt = problem_->dt();
      for(int i = 0; i < N; i++)
      {
        auto & coeff = coeff_list_[i];
        double t = current_t_ + i * dt;
        const StateDimVector & x = variable_.x_list[i];
        const StateDimVector & next_x = variable_.x_list[i + 1];
        const InputDimVector & u = variable_.u_list[i];
        const StateDimVector & lambda = variable_.lambda_list[i];
        const StateDimVector & next_lambda = variable_.lambda_list[i + 1];
        const IneqDimVector & s = variable_.s_list[i];
        const IneqDimVector & nu = variable_.nu_list[i];

        problem_->calcStateEqDeriv(t, x, u, coeff.A, coeff.B);
        problem_->calcIneqConstDeriv(t, x, u, coeff.C, coeff.D);
        problem_->calcRunningCostDeriv(t, x, u, coeff.Lx, coeff.Lu, coeff.Lxx, coeff.Luu, coeff.Lxu);

        coeff.x_bar = sub(problem_->stateEq(t, x, u), next_x); // (2.23c)
        coeff.g_bar = add(problem_->ineqConst(t, x, u), s); // (2.23d)
        for(int j = problem_->ineqDim(t); j < NG; j++) // padding rows: no constraint
        {
          coeff.g_bar[j] = 0.0;
          for(int c = 0; c < NX; c++) coeff.C(j, c) = 0.0;
          for(int c = 0; c < NU; c++) coeff.D(j, c) = 0.0;
        }
        // (2.25b)  -1 * lambda + dt * Lx + A^T next_lambda + C^T nu
        coeff.Lx_bar = add(add(add(scale(-1, lambda), scale(dt, coeff.Lx)), mulT(coeff.A, next_lambda)),
                           mulT(coeff.C, nu));
        // (2.25c)  dt * Lu + B^T next_lambda + D^T nu
        coeff.Lu_bar = add(add(scale(dt, coeff.Lu), mulT(coeff.B, next_lambda)), mulT(coeff.D, nu));
      }
      {
        auto & terminal_coeff = coeff_list_[N];
        double terminal_t = current_t_ + N * dt;
        const StateDimVector & terminal_x = variable_.x_list[N];
        const StateDimVector & terminal_lambda = variable_.lambda_list[N];
        problem_->calcTerminalCostDeriv(terminal_t, terminal_x, terminal_coeff.Lx, terminal_coeff.Lxx);
        terminal_coeff.Lx_bar = sub(terminal_coeff.Lx, terminal_lambda); // (2.25a)
      }
    }

    // KKT error with barrier_eps = 0 (.hpp:443-448)
    double kkt_error = calcKktError(0.0);
    trace_data_list_[trace_idx].kkt_error = kkt_error;
    if(kkt_error <= config_.kkt_error_thre)
    {
      return Status::Succeeded;
    }

    if(!backwardPass()) return Status::ErrorInBackward; // .hpp:451-462
    if(!forwardPass()) return Status::ErrorInForward; // .hpp:465-476
    if(!updateVariables(trace_idx)) return Status::ErrorInUpdate; // .hpp:479-490
    return Status::IterationContinued;
  }

  double calcKktError(double barrier_eps) const
  {
    const int N = config_.horizon_steps;
    double kkt_error = 0;
    kkt_error += squaredNorm(sub(current_x_, variable_.x_list[0])); // .hpp:501
    for(int i = 0; i < N; i++)
    {
      const auto & coeff = coeff_list_[i];
      kkt_error += squaredNorm(coeff.x_bar);
      kkt_error += squaredNorm(coeff.g_bar);
      kkt_error += squaredNorm(coeff.Lx_bar);
      kkt_error += squaredNorm(coeff.Lu_bar);
      double comp = 0.0; // .hpp:510-511
      for(int j = 0; j < NG; j++)
      {
        double v = std::fmax(variable_.s_list[i][j] * variable_.nu_list[i][j] - barrier_eps, 0.0);
        comp += v * v;
      }
      kkt_error += comp;
    }
    kkt_error += squaredNorm(coeff_list_[N].Lx_bar); // .hpp:514-515
    return std::sqrt(kkt_error);
  }

  bool backwardPass()
  {
    const int N = config_.horizon_steps;
    StateStateDimMatrix Qxx_tilde;
    InputInputDimMatrix Quu_tilde;
    StateInputDimMatrix Qxu_tilde;
    StateDimVector Lx_tilde;
    InputDimVector Lu_tilde;
    StateStateDimMatrix F;
    StateInputDimMatrix H;
    InputInputDimMatrix G;
    InputDimVector k;
    InputStateDimMatrix K;
    StateDimVector s;
    StateStateDimMatrix P;

    {
      auto & terminal_coeff = coeff_list_[N]; // .hpp:544-548
      s = scale(-1, terminal_coeff.Lx_bar); // (2.34)
      P = terminal_coeff.Lxx;
      terminal_coeff.s = s;
      terminal_coeff.P = P;
    }

    for(int i = N - 1; i >= 0; i--)
    {
      double dt = problem_->dt();
      auto & coeff = coeff_list_[i];
      const auto & A = coeff.A;
      const auto & B = coeff.B;
      const auto & C = coeff.C;
      const auto & D = coeff.D;
      const auto & Lxx = coeff.Lxx;
      const auto & Luu = coeff.Luu;
      const auto & Lxu = coeff.Lxu;
      const auto & x_bar = coeff.x_bar;
      const auto & g_bar = coeff.g_bar;
      const auto & Lx_bar = coeff.Lx_bar;
      const auto & Lu_bar = coeff.Lu_bar;

      // pre-process (.hpp:572-583)
      IneqDimVector nu_s, tilde_sub;
      for(int j = 0; j < NG; j++)
      {
        nu_s[j] = variable_.nu_list[i][j] / variable_.s_list[i][j];
        tilde_sub[j] = nu_s[j] * g_bar[j] - variable_.nu_list[i][j] + barrier_eps_ * (1.0 / variable_.s_list[i][j]);
      }
      // C^T diag(nu_s) evaluated first (left to right), then times C / D
      Mat<NX, NG> Ct_ns;
      Mat<NU, NG> Dt_ns;
      for(int j = 0; j < NG; j++)
      {
        for(int r = 0; r < NX; r++) Ct_ns(r, j) = C(j, r) * nu_s[j];
        for(int r = 0; r < NU; r++) Dt_ns(r, j) = D(j, r) * nu_s[j];
      }
      Qxx_tilde = add(scale(dt, Lxx), mul(Ct_ns, C)); // (2.28c)
      Quu_tilde = add(scale(dt, Luu), mul(Dt_ns, D)); // (2.28e)
      Qxu_tilde = add(scale(dt, Lxu), mul(Ct_ns, D)); // (2.28d)
      Lx_tilde = add(Lx_bar, mulT(C, tilde_sub)); // (2.28f)
      Lu_tilde = add(Lu_bar, mulT(D, tilde_sub)); // (2.28g)

      F = add(Qxx_tilde, mul(mulT(A, P), A)); // (2.35b)
      H = add(Qxu_tilde, mul(mulT(A, P), B)); // (2.35c)
      G = add(Quu_tilde, mul(mulT(B, P), B)); // (2.35d)

      // gain solve (.hpp:592-624)
      if(NU > 0)
      {
        LdltFactor ldlt;
        ldlt.compute(G.d, NU);
        if(ldlt.success)
        {
          // k = -G^-1 (B^T (P x_bar - s) + Lu_tilde); K = -G^-1 H^T   (2.35e)
          InputDimVector rhs = add(mulT(B, sub(mul(P, x_bar), s)), Lu_tilde);
          ldlt.solveInPlace(rhs.d);
          k = scale(-1, rhs);
          InputStateDimMatrix Ht = transpose(H);
          for(int c = 0; c < NX; c++) ldlt.solveInPlace(&Ht.d[c * NU]);
          K = scale(-1, Ht);
        }
        else
        {
          if(config_.break_if_llt_fails)
          {
            return false;
          }
          // Eigen::FullPivLU fallback (.hpp:614-616)
          FullPivLuFactor lu;
          lu.compute(G.d, NU);
          InputDimVector rhs = add(mulT(B, sub(mul(P, x_bar), s)), Lu_tilde);
          lu.solveInPlace(rhs.d);
          k = scale(-1, rhs);
          InputStateDimMatrix Ht = transpose(H);
          for(int c = 0; c < NX; c++) lu.solveInPlace(&Ht.d[c * NU]);
          K = scale(-1, Ht);
        }
      }

      // post-process (.hpp:633-637)
      s = sub(sub(mulT(A, sub(s, mul(P, x_bar))), Lx_tilde), mul(H, k)); // (2.35a)
      StateStateDimMatrix P_new = sub(F, mul(mulT(K, G), K));
      for(int c = 0; c < NX; c++)
        for(int r = 0; r < NX; r++) P(r, c) = 0.5 * (P_new(r, c) + P_new(c, r));

      coeff.k = k; // .hpp:643-646
      coeff.K = K;
      coeff.s = s;
      coeff.P = P;
    }

    if(config_.check_nan) // .hpp:649-662
    {
      for(const auto & coeff : coeff_list_)
      {
        if(coeff.containsNaN()) return false;
      }
    }
    return true;
  }

  bool forwardPass()
  {
    const int N = config_.horizon_steps;
    delta_variable_.x_list[0] = sub(current_x_, variable_.x_list[0]); // .hpp:670

    for(int i = 0; i < N + 1; i++) // .hpp:672-684
    {
      const auto & coeff = coeff_list_[i];
      delta_variable_.lambda_list[i] = sub(mul(coeff.P, delta_variable_.x_list[i]), coeff.s); // (2.33)
      if(i < N)
      {
        delta_variable_.u_list[i] = add(mul(coeff.K, delta_variable_.x_list[i]), coeff.k); // (2.36)
        delta_variable_.x_list[i + 1] =
            add(add(mul(coeff.A, delta_variable_.x_list[i]), mul(coeff.B, delta_variable_.u_list[i])),
                coeff.x_bar); // (2.26b)
      }
    }

    for(int i = 0; i < N; i++) // .hpp:686-696
    {
      const auto & coeff = coeff_list_[i];
      delta_variable_.s_list[i] =
          scale(-1, add(add(mul(coeff.C, delta_variable_.x_list[i]), mul(coeff.D, delta_variable_.u_list[i])),
                        coeff.g_bar)); // (2.27a)
      for(int j = 0; j < NG; j++)
      {
        delta_variable_.nu_list[i][j] =
            -1 * (variable_.nu_list[i][j] * (delta_variable_.s_list[i][j] + variable_.s_list[i][j]) - barrier_eps_)
            / variable_.s_list[i][j]; // (2.27b)
      }
      for(int j = problem_->ineqDim(current_t_ + i * problem_->dt()); j < NG; j++) // padding rows stay put
      {
        delta_variable_.s_list[i][j] = 0.0;
        delta_variable_.nu_list[i][j] = 0.0;
      }
    }

    if(config_.check_nan && delta_variable_.containsNaN()) // .hpp:698-705
    {
      return false;
    }
    return true;
  }

  bool updateVariables(size_t trace_idx)
  {
    const int N = config_.horizon_steps;
    // fraction-to-boundary (.hpp:714-750)
    double alpha_s_max = 1.0;
    double alpha_nu_max = 1.0;
    {
      constexpr double margin_ratio = 0.995;
      for(int i = 0; i < N; i++)
      {
        const IneqDimVector & s = variable_.s_list[i];
        const IneqDimVector & nu = variable_.nu_list[i];
        const IneqDimVector & delta_s = delta_variable_.s_list[i];
        const IneqDimVector & delta_nu = delta_variable_.nu_list[i];
        for(int ineq_idx = 0; ineq_idx < NG; ineq_idx++)
        {
          if(delta_s[ineq_idx] < 0)
          {
            alpha_s_max = std::min(alpha_s_max, -1 * margin_ratio * s[ineq_idx] / delta_s[ineq_idx]);
          }
          if(delta_nu[ineq_idx] < 0)
          {
            alpha_nu_max = std::min(alpha_nu_max, -1 * margin_ratio * nu[ineq_idx] / delta_nu[ineq_idx]);
          }
        }
      }
      if(!(alpha_s_max > 0.0 && alpha_s_max <= 1.0 && alpha_nu_max > 0.0 && alpha_nu_max <= 1.0))
      {
        return false;
      }
    }

    // line search (.hpp:753-793)
    double alpha_s = alpha_s_max;
    double alpha_nu = alpha_nu_max;
    if(config_.enable_line_search)
    {
      setupMeritFunc();

      constexpr double armijo_scale = 1e-3;
      constexpr double alpha_s_update_ratio = 0.5;
      constexpr double alpha_s_min = 1e-10;
      Variable ls_variable = variable_;
      while(true)
      {
        if(alpha_s < alpha_s_min)
        {
          break;
        }
        for(int i = 0; i < N + 1; i++)
        {
          for(int d = 0; d < NX; d++)
            ls_variable.x_list[i][d] = variable_.x_list[i][d] + alpha_s * delta_variable_.x_list[i][d];
          if(i < N)
          {
            for(int d = 0; d < NU; d++)
              ls_variable.u_list[i][d] = variable_.u_list[i][d] + alpha_s * delta_variable_.u_list[i][d];
            for(int d = 0; d < NG; d++)
              ls_variable.s_list[i][d] = variable_.s_list[i][d] + alpha_s * delta_variable_.s_list[i][d];
          }
        }
        double merit_func_new = calcMeritFunc(ls_variable);
        if(merit_func_new < merit_func_ + armijo_scale * alpha_s * merit_deriv_)
        {
          break;
        }
        alpha_s *= alpha_s_update_ratio;
      }
    }
    trace_data_list_[trace_idx].alpha_s = alpha_s;
    trace_data_list_[trace_idx].alpha_nu = alpha_nu;

    // .hpp:802-831; min_positive_value = numeric_limits<double>::lowest() makes the clamp a no-op
    for(int i = 0; i < N + 1; i++)
    {
      for(int d = 0; d < NX; d++)
      {
        variable_.x_list[i][d] += alpha_s * delta_variable_.x_list[i][d];
        variable_.lambda_list[i][d] += alpha_nu * delta_variable_.lambda_list[i][d];
      }
      if(i < N)
      {
        for(int d = 0; d < NU; d++) variable_.u_list[i][d] += alpha_s * delta_variable_.u_list[i][d];
        for(int d = 0; d < NG; d++)
        {
          variable_.s_list[i][d] += alpha_s * delta_variable_.s_list[i][d];
          variable_.nu_list[i][d] += alpha_nu * delta_variable_.nu_list[i][d];
        }
        constexpr double min_positive_value = std::numeric_limits<double>::lowest();
        for(int d = 0; d < NG; d++)
        {
          variable_.s_list[i][d] = std::fmax(variable_.s_list[i][d], min_positive_value);
          variable_.nu_list[i][d] = std::fmax(variable_.nu_list[i][d], min_positive_value);
        }
      }
    }
    return true;
  }

  void setupMeritFunc()
  {
    const int N = config_.horizon_steps;
    double merit_func_obj = 0.0;
    double merit_func_const = 0.0;
    double merit_deriv_obj = 0.0;
    double merit_deriv_const = 0.0;
    double dt = problem_->dt();
    StateStateDimMatrix neg_identity;
    neg_identity.setZero();
    for(int d = 0; d < NX; d++) neg_identity(d, d) = -1.0;
    Mat<NG, NG> ineq_identity;
    ineq_identity.setZero();
    for(int d = 0; d < NG; d++) ineq_identity(d, d) = 1.0;

    {
      StateDimVector const_func = sub(current_x_, variable_.x_list[0]); // .hpp:846-849
      for(int d = 0; d < NX; d++) merit_func_const += std::fabs(const_func[d]);
      merit_deriv_const += l1NormDirectionalDeriv(const_func, neg_identity, delta_variable_.x_list[0]);
    }

    for(int i = 0; i < N; i++) // .hpp:852-891
    {
      double t = current_t_ + i * dt;
      const auto & x = variable_.x_list[i];
      const auto & u = variable_.u_list[i];
      const auto & s = variable_.s_list[i];
      const auto & next_x = variable_.x_list[i + 1];
      const auto & delta_x = delta_variable_.x_list[i];
      const auto & delta_u = delta_variable_.u_list[i];
      const auto & delta_s = delta_variable_.s_list[i];
      const auto & delta_next_x = delta_variable_.x_list[i + 1];
      const auto & coeff = coeff_list_[i];

      merit_func_obj += problem_->runningCost(t, x, u) * dt;
      merit_deriv_obj += (dot(coeff.Lx, delta_x) + dot(coeff.Lu, delta_u)) * dt;

      double log_sum = 0.0, inv_dot = 0.0;
      for(int d = 0; d < NG; d++)
      {
        log_sum += std::log(s[d]);
        inv_dot += (1.0 / s[d]) * delta_s[d];
      }
      merit_func_obj += -1 * barrier_eps_ * log_sum;
      merit_deriv_obj += -1 * barrier_eps_ * inv_dot;

      {
        StateDimVector const_func = sub(problem_->stateEq(t, x, u), next_x);
        for(int d = 0; d < NX; d++) merit_func_const += std::fabs(const_func[d]);
        merit_deriv_const += l1NormDirectionalDeriv(const_func, coeff.A, delta_x);
        merit_deriv_const += l1NormDirectionalDeriv(const_func, coeff.B, delta_u);
        merit_deriv_const += l1NormDirectionalDeriv(const_func, neg_identity, delta_next_x);
      }
      {
        IneqDimVector const_func = add(problem_->ineqConst(t, x, u), s);
        for(int d = problem_->ineqDim(t); d < NG; d++) const_func[d] = 0.0; // padding rows
        for(int d = 0; d < NG; d++) merit_func_const += std::fabs(const_func[d]);
        merit_deriv_const += l1NormDirectionalDeriv(const_func, coeff.C, delta_x);
        merit_deriv_const += l1NormDirectionalDeriv(const_func, coeff.D, delta_u);
        merit_deriv_const += l1NormDirectionalDeriv(const_func, ineq_identity, delta_s);
      }
    }

    {
      double terminal_t = current_t_ + N * dt; // .hpp:893-901
      merit_func_obj += problem_->terminalCost(terminal_t, variable_.x_list[N]);
      merit_deriv_obj += dot(coeff_list_[N].Lx, delta_variable_.x_list[N]);
    }

    constexpr double merit_const_scale_min = 1e-3; // .hpp:903-923
    if(config_.merit_const_scale_from_lagrange_multipliers)
    {
      merit_const_scale_ = merit_const_scale_min;
      for(int i = 0; i < N + 1; i++)
      {
        for(int d = 0; d < NX; d++)
          merit_const_scale_ = std::max(merit_const_scale_, std::fabs(variable_.lambda_list[i][d]));
        if(i < N)
        {
          for(int d = 0; d < NG; d++)
            merit_const_scale_ = std::max(merit_const_scale_, std::fabs(variable_.nu_list[i][d]));
        }
      }
    }
    else
    {
      constexpr double rho = 0.5;
      merit_const_scale_ = std::max(merit_deriv_obj / ((1.0 - rho) * merit_func_const), merit_const_scale_min);
    }

    merit_func_ = merit_func_obj + merit_const_scale_ * merit_func_const; // .hpp:925-926
    merit_deriv_ = merit_deriv_obj + merit_const_scale_ * merit_deriv_const;
  }

  double calcMeritFunc(const Variable & variable) const
  {
    const int N = config_.horizon_steps;
    double merit_func_obj = 0.0;
    double merit_func_const = 0.0;
    double dt = problem_->dt();
    {
      StateDimVector const_func = sub(current_x_, variable.x_list[0]);
      for(int d = 0; d < NX; d++) merit_func_const += std::fabs(const_func[d]);
    }
    for(int i = 0; i < N; i++)
    {
      double t = current_t_ + i * dt;
      const auto & x = variable.x_list[i];
      const auto & u = variable.u_list[i];
      const auto & s = variable.s_list[i];
      const auto & next_x = variable.x_list[i + 1];
      merit_func_obj += problem_->runningCost(t, x, u) * dt;
      double log_sum = 0.0;
      for(int d = 0; d < NG; d++) log_sum += std::log(s[d]);
      merit_func_obj += -1 * barrier_eps_ * log_sum;
      {
        StateDimVector const_func = sub(problem_->stateEq(t, x, u), next_x);
        for(int d = 0; d < NX; d++) merit_func_const += std::fabs(const_func[d]);
      }
      {
        IneqDimVector const_func = add(problem_->ineqConst(t, x, u), s);
        for(int d = problem_->ineqDim(t); d < NG; d++) const_func[d] = 0.0; // padding rows
        for(int d = 0; d < NG; d++) merit_func_const += std::fabs(const_func[d]);
      }
    }
    {
      double terminal_t = current_t_ + N * dt;
      merit_func_obj += problem_->terminalCost(terminal_t, variable.x_list[N]);
    }
    return merit_func_obj + merit_const_scale_ * merit_func_const;
  }

protected:
  Configuration config_;
  std::shared_ptr<FmpcProblem<NX, NU, NG>> problem_;
  Variable variable_;
  Variable delta_variable_;
  std::vector<Coefficient> coeff_list_;
  std::vector<TraceData> trace_data_list_;
  double current_t_ = 0;
  StateDimVector current_x_ = StateDimVector::Zero();
  double barrier_eps_ = 1e-4; // FmpcSolver.h:413-414; persists across solve() calls
  double merit_const_scale_ = 0.0;
  double merit_func_ = 0.0;
  double merit_deriv_ = 0.0;
};
} // namespace oracle
