/* nmpc_b200 -- a batch of nmpc_ddp::DDPSolver objects sharded over several GPUs of one box, from ONE process.
 *
 * The reference runs one DDPSolver object per problem on one host thread (nmpc_ddp/include/nmpc_ddp/DDPSolver.h:329-374);
 * instances never interact, so a batch splits into contiguous chunks, one per device, with no collective on the data
 * path.  This facade sits on nmpc_b200_ddp_create_sharded (include/nmpc_b200/c_api.h, "several GPUs, one box"): same
 * Configuration, same problem binding and the same per-instance results as nmpc_ddp::DDPSolver::solveBatch on one GPU.
 */
#pragma once

#include <nmpc_ddp/DDPSolver.h>

namespace nmpc_ddp
{
template<int StateDim, int InputDim>
class ShardedDDPSolver
{
public:
  using Single = DDPSolver<StateDim, InputDim>;
  using Configuration = typename Single::Configuration;
  using StateDimVector = typename Single::StateDimVector;
  using InputDimVector = typename Single::InputDimVector;

  /** \param total_capacity largest batch of any later solveBatch()
      \param devices CUDA ordinals, one shard each (empty: every visible device) */
  ShardedDDPSolver(const std::shared_ptr<DDPProblem<StateDim, InputDim>> & problem,
                   int total_capacity,
                   const std::vector<int> & devices = {})
  : problem_(problem), total_capacity_(total_capacity), devices_(devices)
  {
  }
  ~ShardedDDPSolver()
  {
    if(handle_) nmpc_b200_ddp_sharded_destroy(handle_);
  }
  ShardedDDPSolver(const ShardedDDPSolver &) = delete;
  ShardedDDPSolver & operator=(const ShardedDDPSolver &) = delete;

  inline Configuration & config()
  {
    return config_;
  }

  /** \brief Input limits constant over the horizon, for every shard (DDPSolver::setInputLimitsFunc, DDPSolver.h:282-285). */
  void setInputLimits(const InputDimVector & lower, const InputDimVector & upper)
  {
    ensureHandle();
    nmpc_b200::throwOnError(nmpc_b200_ddp_sharded_set_input_limits(handle_, lower.d, upper.d));
  }

  /** \brief DDPSolver::solve for B instances: x0[B][StateDim], u_init[B][n_steps][InputDim] (host arrays).
      \return per-instance value of solve() */
  std::vector<bool> solveBatch(int B, double current_t, const double * x0, const double * u_init, int n_steps)
  {
    ensureHandle();
    nmpc_b200::throwOnError(nmpc_b200_ddp_sharded_solve(handle_, B, current_t, x0, u_init, n_steps));
    last_B_ = B;
    std::vector<int> status(B);
    get(NMPC_B200_DDP_STATUS, status.data(), sizeof(int) * status.size());
    std::vector<bool> converged(B);
    for(int b = 0; b < B; b++) converged[b] = status[b] == 1;
    return converged;
  }

  /** \brief A result field of the last solveBatch() over all shards, in instance order (see nmpc_b200_ddp_field). */
  void get(int field, void * dst, size_t bytes, int dst_device = -1) const
  {
    nmpc_b200::throwOnError(nmpc_b200_ddp_sharded_get(handle_, field, dst, bytes, dst_device));
  }

  /** \brief controlData().u_list[0] of every instance: what an MPC tick applies. */
  std::vector<InputDimVector> firstInputs() const
  {
    std::vector<double> raw(static_cast<size_t>(last_B_) * InputDim);
    get(NMPC_B200_DDP_U0, raw.data(), sizeof(double) * raw.size());
    std::vector<InputDimVector> out(last_B_);
    for(int b = 0; b < last_B_; b++)
      for(int d = 0; d < InputDim; d++) out[b][d] = raw[static_cast<size_t>(b) * InputDim + d];
    return out;
  }

  int numShards() const
  {
    return nmpc_b200_ddp_sharded_num_shards(handle_);
  }

protected:
  void ensureHandle()
  {
    const nmpc_b200_ddp_config c = Single::toC(config_);
    if(!handle_)
    {
      const nmpc_b200::DeviceFunctorBinding binding = problem_->deviceFunctor();
      nmpc_b200::throwOnError(nmpc_b200_ddp_create_sharded(binding.name.c_str(), binding.params.data(),
                                                           static_cast<int>(binding.params.size()), &c, total_capacity_,
                                                           devices_.empty() ? nullptr : devices_.data(),
                                                           static_cast<int>(devices_.size()), &handle_));
    }
    else
    {
      nmpc_b200::throwOnError(nmpc_b200_ddp_sharded_set_config(handle_, &c));
    }
  }

  std::shared_ptr<DDPProblem<StateDim, InputDim>> problem_;
  int total_capacity_;
  std::vector<int> devices_;
  Configuration config_;
  nmpc_b200_ddp_sharded * handle_ = nullptr;
  int last_B_ = 0;
};
} // namespace nmpc_ddp
