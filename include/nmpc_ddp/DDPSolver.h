/* nmpc_b200 -- nmpc_ddp::DDPSolver facade over the C ABI (libnmpc_b200.so).
 *
 * Same public surface as the reference's nmpc_ddp/include/nmpc_ddp/DDPSolver.h:23-375:
 * Configuration (:47-110), ControlData (:113-123), TraceData (:179-216), ComputationDuration (:219-247),
 * DDPSolver(problem), config(), solve(current_t, current_x, initial_u_list), setInputLimitsFunc,
 * controlData(), traceDataList(), computationDuration(), dumpTraceDataList(); plus solveBatch() for B
 * independent instances on one GPU.  Exceptions are the reference's: std::invalid_argument for a wrong
 * initial_u_list length (DDPSolver.hpp:41-45), std::runtime_error otherwise.
 */
#pragma once

#include <array>
#include <cmath>
#include <fstream>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include <nmpc_b200/c_api.h>
#include <nmpc_ddp/DDPProblem.h>

namespace nmpc_b200
{
/** Dynamic vector of doubles with the few Eigen::VectorXd members the solver API exposes. */
class VectorX : public std::vector<double>
{
public:
  using std::vector<double>::vector;
  double sum() const
  {
    double s = 0;
    for(double v : *this) s += v;
    return s;
  }
};

inline void throwOnError(int status)
{
  if(status == NMPC_B200_OK) return;
  const std::string msg = nmpc_b200_last_error();
  if(status == NMPC_B200_ERR_INVALID_ARGUMENT) throw std::invalid_argument(msg);
  throw std::runtime_error(msg);
}
} // namespace nmpc_b200

namespace nmpc_ddp
{
template<int StateDim, int InputDim>
class DDPSolver
{
public:
  using StateDimVector = typename DDPProblem<StateDim, InputDim>::StateDimVector;
  using InputDimVector = typename DDPProblem<StateDim, InputDim>::InputDimVector;
  using StateStateDimMatrix = typename DDPProblem<StateDim, InputDim>::StateStateDimMatrix;
  using InputInputDimMatrix = typename DDPProblem<StateDim, InputDim>::InputInputDimMatrix;
  using StateInputDimMatrix = typename DDPProblem<StateDim, InputDim>::StateInputDimMatrix;
  using InputStateDimMatrix = typename DDPProblem<StateDim, InputDim>::InputStateDimMatrix;

public:
  /*! \brief Configuration (DDPSolver.h:47-110). */
  struct Configuration
  {
    Configuration()
    {
      int list_size = 11;
      alpha_list.resize(list_size);
      for(int i = 0; i < list_size; i++)
      {
        double e = (i == list_size - 1) ? -3.0 : 0.0 + i * ((-3.0 - 0.0) / (list_size - 1));
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
    nmpc_b200::VectorX alpha_list;
    double cost_update_ratio_thre = 0;
    double cost_update_thre = 1e-7;
  };

  /*! \brief Control data (DDPSolver.h:113-123). */
  struct ControlData
  {
    std::vector<StateDimVector> x_list;
    std::vector<InputDimVector> u_list;
    nmpc_b200::VectorX cost_list;
  };

  /*! \brief Data to trace optimization loop (DDPSolver.h:179-216). */
  struct TraceData
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
    double duration_derivative = 0;
    double duration_backward = 0;
    double duration_forward = 0;
  };

  /*! \brief Data of computation duration [msec] (DDPSolver.h:219-247); Q/reg/gain are fused in one kernel. */
  struct ComputationDuration
  {
    double solve = 0;
    double setup = 0;
    double opt = 0;
    double derivative = 0;
    double backward = 0;
    double forward = 0;
    double Q = 0;
    double reg = 0;
    double gain = 0;
  };

public:
  /** \param batch_capacity number of instances solveBatch() may be given (solve() uses one)
      \param device CUDA device ordinal */
  DDPSolver(const std::shared_ptr<DDPProblem<StateDim, InputDim>> & problem, int batch_capacity = 1, int device = 0)
  : problem_(problem), batch_capacity_(batch_capacity), device_(device)
  {
  }
  ~DDPSolver()
  {
    if(handle_) nmpc_b200_ddp_destroy(handle_);
  }
  DDPSolver(const DDPSolver &) = delete;
  DDPSolver & operator=(const DDPSolver &) = delete;

  inline Configuration & config()
  {
    return config_;
  }
  inline const Configuration & config() const
  {
    return config_;
  }

  /** \brief Solve optimization (DDPSolver.hpp:27-141); \return whether converged (retval == 1). */
  bool solve(double current_t, const StateDimVector & current_x, const std::vector<InputDimVector> & initial_u_list)
  {
    std::vector<double> u(initial_u_list.size() * (InputDim > 0 ? InputDim : 1));
    for(size_t i = 0; i < initial_u_list.size(); i++)
      for(int d = 0; d < InputDim; d++) u[i * InputDim + d] = initial_u_list[i][d];
    std::vector<int> status;
    solveBatch(1, current_t, current_x.d, u.data(), static_cast<int>(initial_u_list.size()), &status);
    fetchSingle();
    return status[0] == 1;
  }

  /** \brief Solve B independent instances: x0[B][StateDim], u_init[B][n_steps][InputDim] (host arrays).
      \param status_out per-instance last procOnce() value: 1 converged, 0 max_iter reached, -1 failure */
  void solveBatch(int B,
                  double current_t,
                  const double * x0,
                  const double * u_init,
                  int n_steps,
                  std::vector<int> * status_out = nullptr)
  {
    ensureHandle();
    if(config_.with_input_constraint) applyInputLimits(current_t);
    nmpc_b200::throwOnError(nmpc_b200_ddp_enable_timing(handle_, 1));
    nmpc_b200::throwOnError(nmpc_b200_ddp_solve(handle_, B, current_t, x0, u_init, n_steps, 0, nullptr));
    last_B_ = B;
    if(status_out)
    {
      status_out->resize(B);
      get(NMPC_B200_DDP_STATUS, status_out->data(), sizeof(int) * B);
    }
    double ms[8];
    int launches[4];
    nmpc_b200::throwOnError(nmpc_b200_ddp_get_durations(handle_, ms, launches));
    computation_duration_ = ComputationDuration();
    computation_duration_.solve = ms[0];
    computation_duration_.setup = ms[1] + ms[6];
    computation_duration_.opt = ms[2];
    computation_duration_.derivative = ms[3];
    computation_duration_.backward = ms[4];
    computation_duration_.forward = ms[5];
  }

  /** \brief The receding-horizon loops that call solve() every tick in the reference (TestDDPBipedal.cpp:243-268,
      TestDDPCartPole.cpp:313-343 + :388-396), for B instances, every tick on the device (nmpc_b200_ddp_run_mpc).
      x0[B][StateDim], u_init[B][n_steps][InputDim]; logs (may be null): x_log[B][n_ticks+1][StateDim],
      u_log[B][n_ticks][InputDim], iters_log[B][n_ticks], status_log[B][n_ticks].  Afterwards get() / controlData()
      describe the last tick's solve. */
  void runMpc(int B,
              double current_t,
              const double * x0,
              const double * u_init,
              int n_steps,
              const nmpc_b200_mpc_config & mpc,
              double * x_log,
              double * u_log,
              int * iters_log = nullptr,
              int * status_log = nullptr)
  {
    ensureHandle();
    if(config_.with_input_constraint || mpc.clamp_u0)
    {
      applyInputLimits(current_t);
      applyInputLimitsMpc(current_t, mpc);
    }
    nmpc_b200::throwOnError(nmpc_b200_ddp_run_mpc(handle_, B, current_t, x0, u_init, n_steps, &mpc, x_log, u_log,
                                                  iters_log, status_log, 0, nullptr));
    last_B_ = B;
  }

  /** input_limits_func_ at every (tick, horizon step) time of the loop, as the reference evaluates it at every
      solve (DDPSolver.hpp:470); uploaded only when it actually depends on time. */
  void applyInputLimitsMpc(double current_t, const nmpc_b200_mpc_config & mpc)
  {
    const int N = config_.horizon_steps;
    const size_t per_tick = static_cast<size_t>(N) * (InputDim > 0 ? InputDim : 1);
    std::vector<double> lo(per_tick * mpc.n_ticks), hi(per_tick * mpc.n_ticks);
    bool varies = false;
    for(int k = 0; k < mpc.n_ticks; k++)
      for(int i = 0; i < N; i++)
      {
        const auto limits = input_limits_func_(current_t + k * mpc.tick_dt + i * problem_->dt());
        for(int d = 0; d < InputDim; d++)
        {
          const size_t e = k * per_tick + static_cast<size_t>(i) * InputDim + d;
          lo[e] = limits[0][d];
          hi[e] = limits[1][d];
          if(lo[e] != lo[d] || hi[e] != hi[d]) varies = true;
        }
      }
    if(varies)
      nmpc_b200::throwOnError(nmpc_b200_ddp_set_input_limits_mpc(handle_, mpc.n_ticks, N, lo.data(), hi.data()));
  }

  /** \brief Copy a result field of the last solveBatch() (see nmpc_b200_ddp_field). */
  void get(int field, void * dst, size_t bytes) const
  {
    nmpc_b200::throwOnError(nmpc_b200_ddp_get(handle_, field, dst, bytes, 0, nullptr));
  }

  inline void setInputLimitsFunc(const std::function<std::array<InputDimVector, 2>(double)> & input_limits_func)
  {
    input_limits_func_ = input_limits_func;
  }
  inline const ControlData & controlData() const
  {
    return control_data_;
  }
  inline const std::vector<TraceData> & traceDataList() const
  {
    return trace_data_list_;
  }
  inline const ComputationDuration & computationDuration() const
  {
    return computation_duration_;
  }

  /** \brief Dump trace data list: same columns as DDPSolver.hpp:563-598. */
  void dumpTraceDataList(const std::string & file_path) const
  {
    std::ofstream ofs(file_path);
    ofs << "iter cost lambda dlambda alpha k_rel_norm cost_update_actual cost_update_expected cost_update_ratio "
           "duration_derivative duration_backward duration_forward"
        << std::endl;
    for(const auto & t : trace_data_list_)
    {
      ofs << t.iter << " " << t.cost << " " << t.lambda << " " << t.dlambda << " " << t.alpha << " " << t.k_rel_norm
          << " " << t.cost_update_actual << " " << t.cost_update_expected << " " << t.cost_update_ratio << " "
          << t.duration_derivative << " " << t.duration_backward << " " << t.duration_forward << std::endl;
    }
  }

protected:
  /** input_limits_func_(current_t + i dt) for every horizon step, as backwardPass() evaluates it (DDPSolver.hpp:470). */
  void applyInputLimits(double current_t)
  {
    if(!input_limits_func_) throw std::runtime_error("[DDP] input limits function is not set.");
    const int N = config_.horizon_steps;
    const int nu = InputDim > 0 ? InputDim : 1;
    std::vector<double> lo(static_cast<size_t>(N) * nu), hi(static_cast<size_t>(N) * nu);
    for(int i = 0; i < N; i++)
    {
      const auto limits = input_limits_func_(current_t + i * problem_->dt());
      for(int d = 0; d < InputDim; d++)
      {
        lo[static_cast<size_t>(i) * InputDim + d] = limits[0][d];
        hi[static_cast<size_t>(i) * InputDim + d] = limits[1][d];
      }
    }
    nmpc_b200::throwOnError(nmpc_b200_ddp_set_input_limits_horizon(handle_, N, lo.data(), hi.data()));
  }

public:
  /** The C-ABI form of a Configuration (also used by ShardedDDPSolver). */
  static nmpc_b200_ddp_config toC(const Configuration & config_)
  {
    nmpc_b200_ddp_config c;
    nmpc_b200_ddp_config_default(&c);
    c.horizon_steps = config_.horizon_steps;
    c.max_iter = config_.max_iter;
    c.reg_type = config_.reg_type;
    c.with_input_constraint = config_.with_input_constraint ? 1 : 0;
    c.use_state_eq_second_derivative = config_.use_state_eq_second_derivative ? 1 : 0;
    if(config_.alpha_list.size() > 16) throw std::invalid_argument("alpha_list longer than 16");
    c.n_alpha = static_cast<int>(config_.alpha_list.size());
    for(size_t i = 0; i < config_.alpha_list.size(); i++) c.alpha_list[i] = config_.alpha_list[i];
    c.initial_lambda = config_.initial_lambda;
    c.initial_dlambda = config_.initial_dlambda;
    c.lambda_factor = config_.lambda_factor;
    c.lambda_min = config_.lambda_min;
    c.lambda_max = config_.lambda_max;
    c.k_rel_norm_thre = config_.k_rel_norm_thre;
    c.lambda_thre = config_.lambda_thre;
    c.cost_update_ratio_thre = config_.cost_update_ratio_thre;
    c.cost_update_thre = config_.cost_update_thre;
    return c;
  }

protected:
  nmpc_b200_ddp_config cConfig() const
  {
    return toC(config_);
  }

  void ensureHandle()
  {
    const nmpc_b200_ddp_config c = cConfig();
    if(!handle_)
    {
      const nmpc_b200::DeviceFunctorBinding binding = problem_->deviceFunctor();
      int nx = 0, nu = 0;
      nmpc_b200::throwOnError(nmpc_b200_model_dims(binding.name.c_str(), &nx, &nu, nullptr, nullptr));
      if(nx != StateDim || nu != InputDim)
        throw std::runtime_error("[nmpc_b200] device functor '" + binding.name + "' has dimensions "
                                 + std::to_string(nx) + "x" + std::to_string(nu));
      nmpc_b200::throwOnError(nmpc_b200_ddp_create(binding.name.c_str(), binding.params.data(),
                                                   static_cast<int>(binding.params.size()), &c, batch_capacity_,
                                                   device_, &handle_));
    }
    else
    {
      nmpc_b200::throwOnError(nmpc_b200_ddp_set_config(handle_, &c));
    }
  }

  /** Fill control_data_ / trace_data_list_ from instance 0 of the last solve. */
  void fetchSingle()
  {
    const int N = config_.horizon_steps;
    std::vector<double> x((size_t)last_B_ * (N + 1) * StateDim), u((size_t)last_B_ * N * (InputDim > 0 ? InputDim : 1)),
        c((size_t)last_B_ * (N + 1)), tr((size_t)last_B_ * (config_.max_iter + 1) * 9);
    std::vector<int> n_trace(last_B_);
    get(NMPC_B200_DDP_X, x.data(), x.size() * sizeof(double));
    get(NMPC_B200_DDP_U, u.data(), (size_t)last_B_ * N * InputDim * sizeof(double));
    get(NMPC_B200_DDP_COST_LIST, c.data(), c.size() * sizeof(double));
    get(NMPC_B200_DDP_TRACE, tr.data(), tr.size() * sizeof(double));
    get(NMPC_B200_DDP_N_TRACE, n_trace.data(), n_trace.size() * sizeof(int));
    control_data_.x_list.resize(N + 1);
    control_data_.u_list.resize(N);
    control_data_.cost_list.resize(N + 1);
    for(int i = 0; i <= N; i++)
    {
      for(int d = 0; d < StateDim; d++) control_data_.x_list[i][d] = x[(size_t)i * StateDim + d];
      control_data_.cost_list[i] = c[i];
    }
    for(int i = 0; i < N; i++)
      for(int d = 0; d < InputDim; d++) control_data_.u_list[i][d] = u[(size_t)i * InputDim + d];
    // duration_derivative / _backward / _forward per trace entry (DDPSolver.h:208-215): the stage events of the batch
    std::vector<double> dur((size_t)(config_.max_iter + 1) * 4, 0.0);
    int n_dur = 0;
    nmpc_b200::throwOnError(nmpc_b200_ddp_get_iteration_durations(handle_, dur.data(), config_.max_iter + 1, &n_dur));
    trace_data_list_.clear();
    for(int r = 0; r < n_trace[0]; r++)
    {
      const double * row = &tr[(size_t)r * 9];
      TraceData t;
      if(r < n_dur)
      {
        t.duration_derivative = dur[(size_t)r * 4 + 0];
        t.duration_backward = dur[(size_t)r * 4 + 1];
        t.duration_forward = dur[(size_t)r * 4 + 2] + dur[(size_t)r * 4 + 3];
      }
      t.iter = static_cast<int>(row[0]);
      t.cost = row[1];
      t.lambda = row[2];
      t.dlambda = row[3];
      t.alpha = row[4];
      t.k_rel_norm = row[5];
      t.cost_update_actual = row[6];
      t.cost_update_expected = row[7];
      t.cost_update_ratio = row[8];
      trace_data_list_.push_back(t);
    }
  }

protected:
  Configuration config_;
  std::shared_ptr<DDPProblem<StateDim, InputDim>> problem_;
  std::vector<TraceData> trace_data_list_;
  std::function<std::array<InputDimVector, 2>(double)> input_limits_func_;
  ComputationDuration computation_duration_;
  ControlData control_data_;
  nmpc_b200_ddp * handle_ = nullptr;
  int batch_capacity_ = 1;
  int device_ = 0;
  int last_B_ = 0;
};
} // namespace nmpc_ddp
