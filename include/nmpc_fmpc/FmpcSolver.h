/* nmpc_b200 -- nmpc_fmpc::FmpcSolver facade over the C ABI (libnmpc_b200.so).
 *
 * Same public surface as the reference's nmpc_fmpc/include/nmpc_fmpc/FmpcSolver.h:22-424: Configuration
 * (:58-89), Status (:92-114), Variable (:117-158, reset / containsNaN), TraceData (:232-251),
 * ComputationDuration (:254-288), solve(current_t, current_x, initial_variable), variable(), coeffList()
 * (k and K gains, used for first-step feedback in TestFmpcCartPole.cpp:351-356), traceDataList(),
 * dumpTraceDataList().  checkVariable()'s exceptions are reproduced (FmpcSolver.hpp:285-362).
 */
#pragma once

#include <cmath>
#include <fstream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include <nmpc_b200/c_api.h>
#include <nmpc_ddp/DDPSolver.h>
#include <nmpc_fmpc/FmpcProblem.h>

namespace nmpc_fmpc
{
template<int StateDim, int InputDim, int IneqDim>
class FmpcSolver
{
public:
  using StateDimVector = typename FmpcProblem<StateDim, InputDim, IneqDim>::StateDimVector;
  using InputDimVector = typename FmpcProblem<StateDim, InputDim, IneqDim>::InputDimVector;
  using IneqDimVector = typename FmpcProblem<StateDim, InputDim, IneqDim>::IneqDimVector;
  using StateStateDimMatrix = typename FmpcProblem<StateDim, InputDim, IneqDim>::StateStateDimMatrix;
  using InputStateDimMatrix = typename FmpcProblem<StateDim, InputDim, IneqDim>::InputStateDimMatrix;

public:
  /*! \brief Configuration (FmpcSolver.h:58-89). */
  struct Configuration
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

  /*! \brief Result status (FmpcSolver.h:92-114). */
  enum class Status
  {
    Uninitialized = 0,
    Succeeded = 1,
    ErrorInForward = 2,
    ErrorInBackward = 3,
    ErrorInUpdate = 4,
    MaxIterationReached = 5,
    IterationContinued = 6
  };

  /*! \brief Optimization variables (FmpcSolver.h:117-158). */
  struct Variable
  {
    Variable(int _horizon_steps = 0) : horizon_steps(_horizon_steps)
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
      for(auto & lambda : lambda_list) lambda.setConstant(_lambda);
      for(auto & s : s_list) s.setConstant(_s);
      for(auto & nu : nu_list) nu.setConstant(_nu);
    }
    bool containsNaN() const
    {
      auto bad = [](const auto & list) {
        for(const auto & v : list)
          for(int i = 0; i < v.size(); i++)
            if(std::isnan(v.d[i]) || std::isinf(v.d[i])) return true;
        return false;
      };
      return bad(x_list) || bad(u_list) || bad(lambda_list) || bad(s_list) || bad(nu_list);
    }
    int horizon_steps;
    std::vector<StateDimVector> x_list;
    std::vector<InputDimVector> u_list;
    std::vector<StateDimVector> lambda_list;
    std::vector<IneqDimVector> s_list;
    std::vector<IneqDimVector> nu_list;
    int print_level = 1;
  };

  /*! \brief Gains of the linearised KKT system (the k, K members of FmpcSolver.h:161-229 Coefficient). */
  struct Coefficient
  {
    InputDimVector k;
    InputStateDimMatrix K;
  };

  /*! \brief Data to trace optimization loop (FmpcSolver.h:232-251). */
  struct TraceData
  {
    int iter = 0;
    double kkt_error = 0;
    double duration_coeff = 0;
    double duration_backward = 0;
    double duration_forward = 0;
    double duration_update = 0;
  };

  /*! \brief Data of computation duration [msec] (FmpcSolver.h:254-288). */
  struct ComputationDuration
  {
    double solve = 0;
    double setup = 0;
    double opt = 0;
    double coeff = 0;
    double backward = 0;
    double forward = 0;
    double update = 0;
    double gain_pre = 0;
    double gain_solve = 0;
    double gain_post = 0;
    double fraction = 0;
  };

public:
  FmpcSolver(const std::shared_ptr<FmpcProblem<StateDim, InputDim, IneqDim>> & problem, int device = 0)
  : problem_(problem), device_(device)
  {
  }
  ~FmpcSolver()
  {
    if(handle_) nmpc_b200_fmpc_destroy(handle_);
  }
  FmpcSolver(const FmpcSolver &) = delete;
  FmpcSolver & operator=(const FmpcSolver &) = delete;

  inline Configuration & config()
  {
    return config_;
  }

  /** \brief Solve optimization (FmpcSolver.hpp:158-257). */
  Status solve(double current_t, const StateDimVector & current_x, const Variable & initial_variable)
  {
    const int N = config_.horizon_steps;
    // checkVariable(): sequence lengths (FmpcSolver.hpp:288-312)
    auto check_len = [&](const char * name, size_t have, int want) {
      if(static_cast<int>(have) != want)
        throw std::invalid_argument(std::string("[FMPC] ") + name + " length should be " + std::to_string(want)
                                    + " but " + std::to_string(have) + ".");
    };
    check_len("x_list", initial_variable.x_list.size(), N + 1);
    check_len("u_list", initial_variable.u_list.size(), N);
    check_len("lambda_list", initial_variable.lambda_list.size(), N + 1);
    check_len("s_list", initial_variable.s_list.size(), N);
    check_len("nu_list", initial_variable.nu_list.size(), N);

    ensureHandle();
    auto flat = [](const auto & list, int dim) {
      std::vector<double> out(list.size() * (dim > 0 ? dim : 1));
      for(size_t i = 0; i < list.size(); i++)
        for(int d = 0; d < dim; d++) out[i * dim + d] = list[i].d[d];
      return out;
    };
    const auto x = flat(initial_variable.x_list, StateDim), u = flat(initial_variable.u_list, InputDim),
               l = flat(initial_variable.lambda_list, StateDim), s = flat(initial_variable.s_list, IneqDim),
               nu = flat(initial_variable.nu_list, IneqDim);
    nmpc_b200::throwOnError(nmpc_b200_fmpc_enable_timing(handle_, 1));
    nmpc_b200::throwOnError(nmpc_b200_fmpc_solve(handle_, 1, current_t, current_x.d, x.data(), u.data(), l.data(),
                                                 s.data(), nu.data(), N, 0, nullptr));
    // results
    variable_ = Variable(N);
    auto unflat = [&](int field, auto & list, int dim) {
      std::vector<double> buf(list.size() * (dim > 0 ? dim : 1));
      nmpc_b200::throwOnError(nmpc_b200_fmpc_get(handle_, field, buf.data(), list.size() * dim * sizeof(double), 0,
                                                 nullptr));
      for(size_t i = 0; i < list.size(); i++)
        for(int d = 0; d < dim; d++) list[i].d[d] = buf[i * dim + d];
    };
    unflat(NMPC_B200_FMPC_X, variable_.x_list, StateDim);
    unflat(NMPC_B200_FMPC_U, variable_.u_list, InputDim);
    unflat(NMPC_B200_FMPC_LAMBDA, variable_.lambda_list, StateDim);
    unflat(NMPC_B200_FMPC_S, variable_.s_list, IneqDim);
    unflat(NMPC_B200_FMPC_NU, variable_.nu_list, IneqDim);
    coeff_list_.assign(N, Coefficient());
    {
      std::vector<double> kb((size_t)N * InputDim), Kb((size_t)N * InputDim * StateDim);
      nmpc_b200::throwOnError(nmpc_b200_fmpc_get(handle_, NMPC_B200_FMPC_K_FF, kb.data(), kb.size() * 8, 0, nullptr));
      nmpc_b200::throwOnError(nmpc_b200_fmpc_get(handle_, NMPC_B200_FMPC_K_FB, Kb.data(), Kb.size() * 8, 0, nullptr));
      for(int i = 0; i < N; i++)
      {
        for(int d = 0; d < InputDim; d++) coeff_list_[i].k.d[d] = kb[(size_t)i * InputDim + d];
        for(int d = 0; d < InputDim * StateDim; d++) coeff_list_[i].K.d[d] = Kb[(size_t)i * InputDim * StateDim + d];
      }
    }
    int status = 0, n_trace = 0;
    nmpc_b200::throwOnError(nmpc_b200_fmpc_get(handle_, NMPC_B200_FMPC_STATUS, &status, sizeof(int), 0, nullptr));
    nmpc_b200::throwOnError(nmpc_b200_fmpc_get(handle_, NMPC_B200_FMPC_N_TRACE, &n_trace, sizeof(int), 0, nullptr));
    std::vector<double> tr((size_t)(config_.max_iter > 0 ? config_.max_iter : 1) * 5);
    nmpc_b200::throwOnError(nmpc_b200_fmpc_get(handle_, NMPC_B200_FMPC_TRACE, tr.data(),
                                               (size_t)config_.max_iter * 5 * sizeof(double), 0, nullptr));
    trace_data_list_.clear();
    for(int r = 0; r < n_trace; r++)
    {
      TraceData t;
      t.iter = static_cast<int>(tr[(size_t)r * 5]);
      t.kkt_error = tr[(size_t)r * 5 + 1];
      trace_data_list_.push_back(t);
    }
    double ms[8];
    int launches[4];
    nmpc_b200::throwOnError(nmpc_b200_fmpc_get_durations(handle_, ms, launches));
    computation_duration_ = ComputationDuration();
    computation_duration_.solve = ms[0];
    computation_duration_.setup = ms[1] + ms[7];
    computation_duration_.opt = ms[2];
    computation_duration_.coeff = ms[3];
    computation_duration_.backward = ms[4];
    computation_duration_.forward = ms[5];
    computation_duration_.update = ms[6];
    return static_cast<Status>(status);
  }

  inline const Variable & variable() const
  {
    return variable_;
  }
  inline const std::vector<Coefficient> & coeffList() const
  {
    return coeff_list_;
  }
  inline const std::vector<TraceData> & traceDataList() const
  {
    return trace_data_list_;
  }
  inline const ComputationDuration & computationDuration() const
  {
    return computation_duration_;
  }

  /** \brief Dump trace data list: same columns as FmpcSolver.hpp:259-283. */
  void dumpTraceDataList(const std::string & file_path) const
  {
    std::ofstream ofs(file_path);
    ofs << "iter kkt_error duration_coeff duration_backward duration_forward duration_update" << std::endl;
    for(const auto & t : trace_data_list_)
    {
      ofs << t.iter << " " << t.kkt_error << " " << t.duration_coeff << " " << t.duration_backward << " "
          << t.duration_forward << " " << t.duration_update << std::endl;
    }
  }

protected:
  void ensureHandle()
  {
    nmpc_b200_fmpc_config c;
    nmpc_b200_fmpc_config_default(&c);
    c.horizon_steps = config_.horizon_steps;
    c.max_iter = config_.max_iter;
    c.kkt_error_thre = config_.kkt_error_thre;
    c.check_nan = config_.check_nan;
    c.init_complementary_variable = config_.init_complementary_variable;
    c.update_barrier_eps = config_.update_barrier_eps;
    c.break_if_llt_fails = config_.break_if_llt_fails;
    c.enable_line_search = config_.enable_line_search;
    c.merit_const_scale_from_lagrange_multipliers = config_.merit_const_scale_from_lagrange_multipliers;
    if(!handle_)
    {
      const nmpc_b200::DeviceFunctorBinding binding = problem_->deviceFunctor();
      nmpc_b200::throwOnError(nmpc_b200_fmpc_create(binding.name.c_str(), binding.params.data(),
                                                    static_cast<int>(binding.params.size()), &c, 1, device_,
                                                    &handle_));
    }
    else
    {
      nmpc_b200::throwOnError(nmpc_b200_fmpc_set_config(handle_, &c));
    }
  }

protected:
  Configuration config_;
  std::shared_ptr<FmpcProblem<StateDim, InputDim, IneqDim>> problem_;
  Variable variable_;
  std::vector<Coefficient> coeff_list_;
  std::vector<TraceData> trace_data_list_;
  ComputationDuration computation_duration_;
  nmpc_b200_fmpc * handle_ = nullptr;
  int device_ = 0;
};
} // namespace nmpc_fmpc
