/* nmpc_b200 -- host side of the batched DDP engine for one functor type M.
 *
 * Owns the device workspace for `capacity` instances, converts the instance-major arrays of the
 * C ABI to the batch-innermost device layout, and enqueues K0, then max_iter x {K1, K2, K3} -- for n_x < 8
 * max_iter x {K1+K2 fused, K3} -- on one stream with no host round trip (per-instance lambda/status/iteration
 * counters live on the device; finished instances' threads return immediately).  Kernel variants are chosen here by
 * n_x and batch size (thresholds measured on B200); runMpc() chains whole solves tick after tick.  For max_iter > kCheckStride the host polls an
 * active-instance counter every kCheckStride iterations so that a batch that has converged does not
 * pay for hundreds of empty launches (DDPSolver::Configuration::max_iter defaults to 500).
 */
#pragma once

#include <algorithm>
#include <cctype>
#include <string>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "common.cuh"
#include "ddp_kernels.cuh"
#include "ddp_backward_coop.cuh"
#include "ddp_backward_quad.cuh"
#include "ddp_backward_wide.cuh"
#include "ddp_backward_fused.cuh"
#include "ddp_backward_lanes.cuh"
#include "ddp_forward_phased.cuh"
#include "ddp_forward_split.cuh"
#include "ddp_solve_tile.cuh"
#include "ddp_mpc.cuh"
#include "registry.h"

namespace nmpc_b200
{
namespace ddp
{
constexpr int kCheckStride = 16;

static __global__ void count_active_kernel(const int * status, int B, int * counter)
{
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = (b < B) && (status[b] == 0);
  const unsigned m = __ballot_sync(0xffffffffu, active);
  if((threadIdx.x & 31) == 0 && m != 0) atomicAdd(counter, __popc(m));
}

template<class S>
__global__ void extract_u0_kernel(const S * u0buf, const S * u1buf, const int * sel, double * dst, int B, int NU, int Bp)
{
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if(b >= B) return;
  const S * u = sel[b] ? u1buf : u0buf;
  for(int d = 0; d < NU; d++) dst[(size_t)b * NU + d] = double(u[(size_t)d * Bp + b]);
}

/** Kernel launch, optionally with the programmatic-stream-serialization attribute (see pdlPrologue()).
    Measured on B200 (cart-pole, B=4096, M-fixed): letting the ~45 dependent kernels of a solve pre-launch made
    the step SLOWER (2.87 ms vs 2.18 ms) -- the early-resident grids of the wide linearisation kernel get in the
    way of the latency-bound sweeps -- so the attribute is off unless NMPC_B200_PDL=1. */
template<class... KArgs, class... Args>
inline void launchPdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&... args)
{
  static const bool use_pdl = [] {
    const char * env = std::getenv("NMPC_B200_PDL");
    return env != nullptr && env[0] == '1';
  }();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = use_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  NMPC_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...));
}

/** The engine's kernel-selection knobs (nmpc_b200_ddp_set_tuning / _get_tuning, c_api.h).  Defaults come from the
    measurements on a 148-SM B200 recorded in DESIGN.md, scaled by the SM count of the device the engine runs on;
    -1 = "decide from the problem size".  An environment variable NMPC_B200_<KEY IN CAPITALS> overrides a default at
    construction (developer switch for tools/exp_variants.py); a call to setTuning overrides both. */
struct DdpTuning
{
  int backward_lanes = 1; //!< n_x < 8, small batches: 0 thread per instance, 1 G lanes + smem exchange, 2 shuffles
  int backward_lanes_tiles_per_cta = 1; //!< 32-instance tiles per CTA of the lanes kernel
  int backward_lanes_max_batch = 0; //!< one tile per SM; beyond, the fused thread-per-instance sweep wins
  int backward_fused = 1; //!< linearisation fused into the backward kernel (n_x < 8)
  int backward_quad = -1; //!< four warps per tile (n_x >= 8): -1 by size, 0 never, 1 always
  int backward_quad_max_batch = 0;
  int backward_group_size = -1; //!< lanes per instance of the cooperative sweep: -1 by size, 1, or kCoopGS
  int backward_coop_max_batch = 0;
  int backward_wide = 1; //!< n_u >= 8 without limits: the n_u side spread over a lane group
  int forward_lanes = -1; //!< -1 by size; 1 thread per instance; 4 / 16 lanes speculate; 3 = three-phase line search
  int forward_phased_max_batch = 0;
  int forward_split = 1; //!< phased line search with rollout / cost / loader warps
  int forward_split_max_batch = 0;
  int threads_per_block = -1;
  int solve_tile = 0; //!< persistent CTA per tile for the whole solve (measured slower)
  int solve_tile_max_batch = 0;
  int sm_count = 148;

  void scaleToDevice(int sms)
  {
    sm_count = sms;
    backward_lanes_max_batch = sms * 32; // one 32-instance tile per SM
    backward_quad_max_batch = sms * 4 * 32;
    backward_coop_max_batch = sms * 128;
    forward_phased_max_batch = sms * 332; // measured cross-over 49152 on 148 SMs
    forward_split_max_batch = sms * 83; // measured cross-over 12288 on 148 SMs
    solve_tile_max_batch = sms * 32;
  }

  /** Pointer to the knob called `key`, or nullptr. */
  int * find(const std::string & key)
  {
#define NMPC_B200_TUNING_KEY(k) \
  if(key == #k) return &k;
    NMPC_B200_TUNING_KEY(backward_lanes)
    NMPC_B200_TUNING_KEY(backward_lanes_tiles_per_cta)
    NMPC_B200_TUNING_KEY(backward_lanes_max_batch)
    NMPC_B200_TUNING_KEY(backward_fused)
    NMPC_B200_TUNING_KEY(backward_quad)
    NMPC_B200_TUNING_KEY(backward_quad_max_batch)
    NMPC_B200_TUNING_KEY(backward_group_size)
    NMPC_B200_TUNING_KEY(backward_coop_max_batch)
    NMPC_B200_TUNING_KEY(backward_wide)
    NMPC_B200_TUNING_KEY(forward_lanes)
    NMPC_B200_TUNING_KEY(forward_phased_max_batch)
    NMPC_B200_TUNING_KEY(forward_split)
    NMPC_B200_TUNING_KEY(forward_split_max_batch)
    NMPC_B200_TUNING_KEY(threads_per_block)
    NMPC_B200_TUNING_KEY(solve_tile)
    NMPC_B200_TUNING_KEY(solve_tile_max_batch)
#undef NMPC_B200_TUNING_KEY
    return nullptr;
  }

  static const char * const * keys(int & n)
  {
    static const char * const k[] = {"backward_lanes", "backward_lanes_tiles_per_cta", "backward_lanes_max_batch",
                                     "backward_fused", "backward_quad", "backward_quad_max_batch", "backward_group_size",
                                     "backward_coop_max_batch", "backward_wide", "forward_lanes", "forward_phased_max_batch",
                                     "forward_split", "forward_split_max_batch", "threads_per_block", "solve_tile",
                                     "solve_tile_max_batch"};
    n = (int)(sizeof(k) / sizeof(k[0]));
    return k;
  }

  /** NMPC_B200_<KEY> environment overrides (plus the older short names the experiment scripts use). */
  void overlayEnvironment()
  {
    int n = 0;
    const char * const * k = keys(n);
    for(int i = 0; i < n; i++)
    {
      std::string name = "NMPC_B200_";
      for(const char * c = k[i]; *c; c++) name += (char)std::toupper((unsigned char)*c);
      if(const char * env = std::getenv(name.c_str())) *find(k[i]) = std::atoi(env);
    }
    static const char * const alias[][2] = {{"NMPC_B200_BWD_LANES", "backward_lanes"},
                                            {"NMPC_B200_BWD_LANES_TPC", "backward_lanes_tiles_per_cta"},
                                            {"NMPC_B200_BWD_LANES_MAXB", "backward_lanes_max_batch"},
                                            {"NMPC_B200_BWD_FUSED", "backward_fused"},
                                            {"NMPC_B200_BWD_QUAD", "backward_quad"},
                                            {"NMPC_B200_BWD_GS", "backward_group_size"},
                                            {"NMPC_B200_BWD_WIDE", "backward_wide"},
                                            {"NMPC_B200_FWD_GA", "forward_lanes"},
                                            {"NMPC_B200_FWD_SPLIT", "forward_split"},
                                            {"NMPC_B200_FWD_SPLIT_MAXB", "forward_split_max_batch"},
                                            {"NMPC_B200_TPB", "threads_per_block"},
                                            {"NMPC_B200_TILE", "solve_tile"},
                                            {"NMPC_B200_TILE_MAXB", "solve_tile_max_batch"}};
    for(const auto & a : alias)
      if(const char * env = std::getenv(a[0])) *find(a[1]) = std::atoi(env);
  }
};

template<class M>
class DdpEngine : public DdpEngineBase
{
public:
  using S = typename M::Scalar;
  static constexpr int NX = M::NX;
  static constexpr int NU = M::NU;
  using L = BlockLayout<NX, NU>;

  DdpEngine(const double * params, const nmpc_b200_ddp_config & cfg, int batch_capacity, int device)
  : model_(M::fromParams(params)), device_(device), capacity_(batch_capacity)
  {
    if(batch_capacity <= 0) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "batch_capacity must be positive");
    DeviceGuard guard(device_);
    NMPC_CUDA_CHECK(cudaStreamCreateWithFlags(&own_stream_, cudaStreamNonBlocking));
    NMPC_CUDA_CHECK(cudaMallocHost(reinterpret_cast<void **>(&h_counter_), sizeof(int)));
    Bp_ = ((capacity_ + 127) / 128) * 128;
    int sms = 0;
    NMPC_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device_));
    tune_.scaleToDevice(sms);
    tune_.overlayEnvironment();
    if(kThreadSweepFits)
    {
      NMPC_CUDA_CHECK(cudaFuncSetAttribute(backward_kernel<M, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)backwardSmemBytes(maxThreadsPerBlock())));
      NMPC_CUDA_CHECK(cudaFuncSetAttribute(backward_kernel<M, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)backwardSmemBytes(maxThreadsPerBlock())));
    }
    applyConfig(cfg, true);
  }

  ~DdpEngine() override
  {
    DeviceGuard guard(device_);
    cudaStreamSynchronize(own_stream_);
    for(auto e : events_) cudaEventDestroy(e);
    cudaStreamDestroy(own_stream_);
    cudaFreeHost(h_counter_);
  }

  void setConfig(const nmpc_b200_ddp_config & cfg) override
  {
    DeviceGuard guard(device_);
    applyConfig(cfg, false);
  }

  const nmpc_b200_ddp_config & config() const override
  {
    return cfg_;
  }

  void setTuning(const char * key, int value) override
  {
    int * knob = key ? tune_.find(key) : nullptr;
    if(knob == nullptr) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, std::string("unknown tuning key '") + (key ? key : "") + "'");
    if(*knob == value) return;
    DeviceGuard guard(device_);
    *knob = value;
    // the derivative buffers exist only for the unfused pipeline: the choice is made at allocation
    if(use_fused_ != backwardUsesFused(capacity_)) allocate();
  }

  int getTuning(const char * key) override
  {
    int * knob = key ? tune_.find(key) : nullptr;
    if(knob == nullptr) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, std::string("unknown tuning key '") + (key ? key : "") + "'");
    return *knob;
  }

  void setInputLimits(const double * lower, const double * upper) override
  {
    // limits constant over the horizon: the same pair at every step
    const int N = cfg_.horizon_steps;
    std::vector<double> lo((size_t)N * NU), hi((size_t)N * NU);
    for(int i = 0; i < N; i++)
      for(int d = 0; d < NU; d++)
      {
        lo[(size_t)i * NU + d] = lower[d];
        hi[(size_t)i * NU + d] = upper[d];
      }
    setInputLimitsHorizon(N, lo.data(), hi.data());
    limits_vary_ = false;
    mpc_limit_ticks_ = 0;
  }

  /** input_limits_func_(current_t + i dt) for i = 0 .. N-1 (DDPSolver.h:282-285, DDPSolver.hpp:470): lower / upper
      [n_steps][NU]. */
  void setInputLimitsHorizon(int n_steps, const double * lower, const double * upper) override
  {
    DeviceGuard guard(device_);
    const int N = cfg_.horizon_steps;
    if(n_steps != N)
      throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "input limits are needed for " + std::to_string(N) + " steps but "
                                                      + std::to_string(n_steps) + " were given");
    if(lower == nullptr || upper == nullptr) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "null input limits");
    u_lo_.assign(lower, lower + (size_t)N * NU);
    u_hi_.assign(upper, upper + (size_t)N * NU);
    std::vector<S> lo((size_t)N * NU), hi((size_t)N * NU);
    limits_vary_ = false;
    for(size_t e = 0; e < lo.size(); e++)
    {
      lo[e] = S(lower[e]);
      hi[e] = S(upper[e]);
      if(lower[e] != lower[e % NU] || upper[e] != upper[e % NU]) limits_vary_ = true;
    }
    if(NU > 0)
    {
      NMPC_CUDA_CHECK(cudaMemcpy(d_u_lo_.ptr, lo.data(), sizeof(S) * lo.size(), cudaMemcpyHostToDevice));
      NMPC_CUDA_CHECK(cudaMemcpy(d_u_hi_.ptr, hi.data(), sizeof(S) * hi.size(), cudaMemcpyHostToDevice));
    }
    have_limits_ = true;
  }

  /** input_limits_func_(current_t + tick * tick_dt + i dt) for every tick of the device-resident MPC loop and every
      horizon step (DDPSolver.hpp:470 evaluates the function anew at every solve): lower / upper [n_ticks][n_steps][NU]. */
  void setInputLimitsMpc(int n_ticks, int n_steps, const double * lower, const double * upper) override
  {
    DeviceGuard guard(device_);
    const int N = cfg_.horizon_steps;
    if(n_steps != N)
      throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "input limits are needed for " + std::to_string(N) + " steps but "
                                                      + std::to_string(n_steps) + " were given");
    if(n_ticks <= 0 || lower == nullptr || upper == nullptr)
      throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "null or empty per-tick input limits");
    const size_t n = (size_t)n_ticks * N * NU;
    std::vector<S> lo(n), hi(n);
    for(size_t e = 0; e < n; e++)
    {
      lo[e] = S(lower[e]);
      hi[e] = S(upper[e]);
    }
    mpc_lo_.allocate(n > 0 ? n : 1);
    mpc_hi_.allocate(n > 0 ? n : 1);
    if(n > 0)
    {
      NMPC_CUDA_CHECK(cudaMemcpy(mpc_lo_.ptr, lo.data(), sizeof(S) * n, cudaMemcpyHostToDevice));
      NMPC_CUDA_CHECK(cudaMemcpy(mpc_hi_.ptr, hi.data(), sizeof(S) * n, cudaMemcpyHostToDevice));
    }
    mpc_limit_ticks_ = n_ticks;
  }

  void solve(int B, double current_t, const double * x0, const double * u_init, int n_u_steps, bool on_device, void * stream)
      override
  {
    DeviceGuard guard(device_);
    cudaStream_t st = beginSolve(B, n_u_steps, x0, u_init, stream);
    n_events_used_ = 0;
    record(st); // 0: start
    stageInputs(B, x0, u_init, on_device, st);
    runIterations(B, current_t, st);
  }

  /** The reference's MPC loops for the whole batch, tick after tick on the device (c_api.h, ddp_mpc.cuh). */
  void runMpc(int B,
              double current_t,
              const double * x0,
              const double * u_init,
              int n_u_steps,
              const nmpc_b200_mpc_config & mpc,
              double * x_log,
              double * u_log,
              int * iters_log,
              int * status_log,
              bool on_device,
              void * stream) override
  {
    DeviceGuard guard(device_);
    if(mpc.n_ticks <= 0) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "n_ticks must be positive");
    if(mpc.plant != 0 && mpc.plant != 1) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "plant must be 0 or 1");
    if(mpc.plant == 1 && !HasStateEqDt<M>::value)
      throw Error(NMPC_B200_ERR_UNSUPPORTED, "plant = 1 needs a functor with stateEq(t, x, u, dt)");
    if(mpc.plant == 1 && mpc.n_substeps <= 0) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "n_substeps must be positive");
    if(mpc.clamp_u0 && !have_limits_) throw Error(NMPC_B200_ERR_RUNTIME, "clamp_u0 is set but no input limits were given");
    const bool needs_limits = mpc.clamp_u0 || cfg_.with_input_constraint;
    const bool per_tick_limits = needs_limits && mpc_limit_ticks_ >= mpc.n_ticks;
    if(limits_vary_ && needs_limits && !per_tick_limits)
      throw Error(NMPC_B200_ERR_UNSUPPORTED,
                  "the input limits change along the horizon, so every tick of the device-resident MPC loop needs its own "
                  "table: give them with nmpc_b200_ddp_set_input_limits_mpc for at least n_ticks ticks");
    cudaStream_t st = beginSolve(B, n_u_steps, x0, u_init, stream);
    const size_t T = mpc.n_ticks;
    const size_t Bp = Bp_;
    MpcLogs<S> logs{};
    if(x_log) logs.x = (mpc_x_.reserve((T + 1) * NX * Bp), mpc_x_.ptr);
    if(u_log) logs.u = (mpc_u_.reserve(T * NU * Bp), mpc_u_.ptr);
    if(iters_log || status_log) mpc_i_.reserve(2 * T * Bp);
    if(iters_log) logs.iters = mpc_i_.ptr;
    if(status_log) logs.status = mpc_i_.ptr + T * Bp;
    MpcParams<S> mp{};
    mp.n_ticks = mpc.n_ticks;
    mp.plant = mpc.plant;
    mp.shift_inputs = mpc.shift_inputs;
    mp.clamp_u0 = mpc.clamp_u0;
    mp.n_substeps = mpc.n_substeps;
    mp.tick_dt = S(mpc.tick_dt);
    mp.sim_dt = S(mpc.sim_dt);

    n_events_used_ = 0;
    record(st);
    stageInputs(B, x0, u_init, on_device, st);
    for(int tick = 0; tick < mpc.n_ticks; tick++)
    {
      const double t = current_t + tick * mpc.tick_dt;
      if(tick > 0)
      {
        // per-tick event bookkeeping restarts so that computationDuration() describes the last solve
        n_events_used_ = 0;
        record(st);
        record(st);
      }
      if(per_tick_limits)
      {
        // this tick's horizon: input_limits_func_(t + i dt), i = 0 .. N-1 (the clamp of the applied input reads step 0)
        ws_.u_lo = mpc_lo_.ptr + (size_t)tick * cfg_.horizon_steps * NU;
        ws_.u_hi = mpc_hi_.ptr + (size_t)tick * cfg_.horizon_steps * NU;
      }
      runIterations(B, t, st);
      mpc_advance_kernel<M><<<(B + 127) / 128, 128, 0, st>>>(model_, ws_, prm_, mp, logs, tick, S(t));
      NMPC_CUDA_CHECK(cudaGetLastError());
    }
    ws_.u_lo = d_u_lo_.ptr;
    ws_.u_hi = d_u_hi_.ptr;
    // logs back to instance-major
    auto out_f64 = [&](const S * src, double * dst, int R) {
      if(dst == nullptr) return;
      const size_t need = sizeof(double) * (size_t)B * R;
      double * d_out = on_device ? dst : stageOut(need);
      launchGatherRows<S, double>(src, src, nullptr, nullptr, 0, 1, d_out, B, R, Bp_, st);
      if(!on_device)
      {
        NMPC_CUDA_CHECK(cudaMemcpyAsync(dst, d_out, need, cudaMemcpyDeviceToHost, st));
        NMPC_CUDA_CHECK(cudaStreamSynchronize(st));
      }
    };
    auto out_i32 = [&](const int * src, int * dst, int R) {
      if(dst == nullptr) return;
      const size_t need = sizeof(int) * (size_t)B * R;
      int * d_out = on_device ? dst : reinterpret_cast<int *>(stageOut(need));
      launchGatherRows<int, int>(src, src, nullptr, nullptr, 0, 1, d_out, B, R, Bp_, st);
      if(!on_device)
      {
        NMPC_CUDA_CHECK(cudaMemcpyAsync(dst, d_out, need, cudaMemcpyDeviceToHost, st));
        NMPC_CUDA_CHECK(cudaStreamSynchronize(st));
      }
    };
    out_f64(logs.x, x_log, (mpc.n_ticks + 1) * NX);
    out_f64(logs.u, u_log, mpc.n_ticks * NU);
    out_i32(logs.iters, iters_log, mpc.n_ticks);
    out_i32(logs.status, status_log, mpc.n_ticks);
    NMPC_CUDA_CHECK(cudaGetLastError());
  }

  void get(int what, void * dst, size_t dst_bytes, bool dst_on_device, void * stream) override
  {
    DeviceGuard guard(device_);
    if(B_ <= 0) throw Error(NMPC_B200_ERR_RUNTIME, "get() before solve()");
    if(dst == nullptr) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "null destination");
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : last_stream_;
    const int N = cfg_.horizon_steps;
    const int B = B_;
    size_t need = 0;
    bool is_int = false;
    const int * int_src = nullptr;
    int R = 0;
    switch(what)
    {
      case NMPC_B200_DDP_X:
        R = (N + 1) * NX;
        break;
      case NMPC_B200_DDP_U:
        R = N * NU;
        break;
      case NMPC_B200_DDP_COST_LIST:
        R = N + 1;
        break;
      case NMPC_B200_DDP_K_FF:
        R = N * NU;
        break;
      case NMPC_B200_DDP_K_FB:
        R = N * NU * NX;
        break;
      case NMPC_B200_DDP_TRACE:
        R = (cfg_.max_iter + 1) * kTraceFields;
        break;
      case NMPC_B200_DDP_COST:
        R = 1;
        break;
      case NMPC_B200_DDP_U0:
        R = NU;
        break;
      case NMPC_B200_DDP_STATUS:
        is_int = true;
        int_src = ws_.status;
        break;
      case NMPC_B200_DDP_ITERS:
        is_int = true;
        int_src = ws_.iters;
        break;
      case NMPC_B200_DDP_N_FORWARD:
        is_int = true;
        int_src = ws_.n_fwd;
        break;
      case NMPC_B200_DDP_N_BACKWARD:
        is_int = true;
        int_src = ws_.n_bwd;
        break;
      case NMPC_B200_DDP_N_TRACE:
        is_int = true;
        break;
      default:
        throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "unknown DDP field " + std::to_string(what));
    }
    need = is_int ? sizeof(int) * (size_t)B : sizeof(double) * (size_t)B * R;
    if(dst_bytes < need)
    {
      throw Error(NMPC_B200_ERR_INVALID_ARGUMENT,
                  "destination too small: " + std::to_string(dst_bytes) + " < " + std::to_string(need));
    }

    record(st);
    if(is_int)
    {
      if(what == NMPC_B200_DDP_N_TRACE)
      {
        // traceDataList().size() == last iteration + 1
        std::vector<int> it(B);
        NMPC_CUDA_CHECK(cudaMemcpyAsync(it.data(), ws_.iters, need, cudaMemcpyDeviceToHost, st));
        NMPC_CUDA_CHECK(cudaStreamSynchronize(st));
        for(auto & v : it) v += 1;
        NMPC_CUDA_CHECK(cudaMemcpy(dst, it.data(), need, dst_on_device ? cudaMemcpyHostToDevice : cudaMemcpyHostToHost));
      }
      else
      {
        NMPC_CUDA_CHECK(
            cudaMemcpyAsync(dst, int_src, need, dst_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
        if(!dst_on_device) NMPC_CUDA_CHECK(cudaStreamSynchronize(st));
      }
      record(st);
      return;
    }

    double * d_out = dst_on_device ? static_cast<double *>(dst) : stageOut(need);
    switch(what)
    {
      case NMPC_B200_DDP_X:
        launchGatherRows<S, double>(ws_.x[0], ws_.x[1], ws_.sel, nullptr, 0, 1, d_out, B, R, Bp_, st);
        break;
      case NMPC_B200_DDP_U:
        launchGatherRows<S, double>(ws_.u[0], ws_.u[1], ws_.sel, nullptr, 0, 1, d_out, B, R, Bp_, st);
        break;
      case NMPC_B200_DDP_COST_LIST:
        launchGatherRows<S, double>(ws_.cost[0], ws_.cost[1], ws_.sel, nullptr, 0, 1, d_out, B, R, Bp_, st);
        break;
      case NMPC_B200_DDP_K_FF:
        launchGatherRows<S, double>(ws_.kff, ws_.kff, nullptr, nullptr, 0, 1, d_out, B, R, Bp_, st);
        break;
      case NMPC_B200_DDP_K_FB:
        launchGatherRows<S, double>(ws_.kfb, ws_.kfb, nullptr, nullptr, 0, 1, d_out, B, R, Bp_, st);
        break;
      case NMPC_B200_DDP_TRACE:
        // rows past the last executed iteration were never written: read them as zero
        launchGatherRows<S, double>(ws_.trace, ws_.trace, nullptr, ws_.iters, 1, kTraceFields, d_out, B, R, Bp_, st);
        break;
      case NMPC_B200_DDP_COST:
        launchGatherRows<S, double>(ws_.cost_sum, ws_.cost_sum, nullptr, nullptr, 0, 1, d_out, B, R, Bp_, st);
        break;
      case NMPC_B200_DDP_U0:
        extract_u0_kernel<S><<<(B + 127) / 128, 128, 0, st>>>(ws_.u[0], ws_.u[1], ws_.sel, d_out, B, NU, Bp_);
        break;
    }
    NMPC_CUDA_CHECK(cudaGetLastError());
    if(!dst_on_device)
    {
      NMPC_CUDA_CHECK(cudaMemcpyAsync(dst, d_out, need, cudaMemcpyDeviceToHost, st));
      NMPC_CUDA_CHECK(cudaStreamSynchronize(st));
    }
    record(st);
  }

  void sync() override
  {
    DeviceGuard guard(device_);
    NMPC_CUDA_CHECK(cudaStreamSynchronize(last_stream_ ? last_stream_ : own_stream_));
  }

  void enableTiming(bool enable) override
  {
    timing_ = enable;
  }

  /** TraceData::duration_derivative / _backward / _forward of every iteration (DDPSolver.h:208-215) from the stage
      events of the last solve: ms[rows][4] = {derivative, backward, first line-search candidate, other candidates};
      row 0 is the initial rollout (column 2).  Returns the number of rows filled. */
  int getIterationDurations(double * ms, int rows) override
  {
    DeviceGuard guard(device_);
    if(!timing_ || n_events_used_ < 4 || rows <= 0) return 0;
    NMPC_CUDA_CHECK(cudaStreamSynchronize(last_stream_));
    auto el = [&](int a, int b) {
      float t = 0.f;
      cudaEventElapsedTime(&t, events_[a], events_[b]);
      return double(t);
    };
    for(int i = 0; i < 4 * rows; i++) ms[i] = 0.0;
    if(persistent_last_)
    {
      if(!persistent_timed_) return 0;
      const int n = std::min(rows, cfg_.max_iter + 1);
      std::vector<unsigned long long> ns(4 * (size_t)(cfg_.max_iter + 1));
      NMPC_CUDA_CHECK(cudaMemcpy(ns.data(), stage_ns_.ptr, sizeof(unsigned long long) * ns.size(), cudaMemcpyDeviceToHost));
      ms[2] = 1e-6 * double(ns[1]);
      for(int it = 1; it < n; it++)
      {
        ms[4 * it + 1] = 1e-6 * double(ns[4 * (size_t)it]);
        ms[4 * it + 2] = 1e-6 * double(ns[4 * (size_t)it + 1]);
      }
      return n;
    }
    ms[2] = el(1, 2);
    const int n = std::min(rows, iters_launched_ + 1);
    for(int it = 1; it < n; it++)
    {
      const int e = iter_event_base_ + kEventsPerIter * (it - 1);
      ms[4 * it + 0] = el(e - 1, e);
      ms[4 * it + 1] = el(e, e + 1);
      ms[4 * it + 2] = el(e + 1, e + 2);
      ms[4 * it + 3] = el(e + 2, e + 3);
    }
    return n;
  }

  void getDurations(double * ms, int * launches) override
  {
    DeviceGuard guard(device_);
    for(int i = 0; i < 8; i++) ms[i] = 0.0;
    if(launches)
      for(int i = 0; i < 4; i++) launches[i] = launches_[i];
    if(!timing_ || n_events_used_ < 4) return;
    NMPC_CUDA_CHECK(cudaStreamSynchronize(last_stream_));
    auto el = [&](int a, int b) {
      float t = 0.f;
      cudaEventElapsedTime(&t, events_[a], events_[b]);
      return double(t);
    };
    const int end_opt = iter_event_base_ + kEventsPerIter * iters_launched_;
    ms[6] = el(0, 1); // copy_in
    ms[1] = el(1, 2); // setup: layout + initial rollout
    double der = 0, bwd = 0, fwd = 0;
    for(int it = 0; it < iters_launched_; it++)
    {
      const int e = iter_event_base_ + kEventsPerIter * it;
      der += el(e - 1, e);
      bwd += el(e, e + 1);
      fwd += el(e + 1, e + 3);
    }
    ms[3] = der;
    ms[4] = bwd;
    ms[5] = fwd;
    ms[2] = el(2, end_opt); // opt (includes any active-count polls)
    ms[0] = el(0, end_opt); // solve
    if(persistent_last_ && persistent_timed_)
    {
      // one kernel: the stage durations come from its own %globaltimer stamps (maximum over the tiles, per iteration)
      std::vector<unsigned long long> ns(4 * (size_t)(cfg_.max_iter + 1));
      NMPC_CUDA_CHECK(cudaMemcpy(ns.data(), stage_ns_.ptr, sizeof(unsigned long long) * ns.size(), cudaMemcpyDeviceToHost));
      const double k0 = 1e-6 * double(ns[1]);
      double b = 0, f = 0;
      const int tiles = (B_ + kTile - 1) / kTile;
      const bool dump = std::getenv("NMPC_B200_TILE_DUMP") != nullptr;
      for(int it = 1; it <= cfg_.max_iter; it++)
      {
        b += 1e-6 * double(ns[4 * (size_t)it]);
        f += 1e-6 * double(ns[4 * (size_t)it + 1]);
        if(dump)
          std::fprintf(stderr, "iter %d: bwd max %.1f mean %.1f us, fwd max %.1f mean %.1f us\n", it, 1e-3 * double(ns[4 * (size_t)it]),
                       1e-3 * double(ns[4 * (size_t)it + 2]) / tiles, 1e-3 * double(ns[4 * (size_t)it + 1]),
                       1e-3 * double(ns[4 * (size_t)it + 3]) / tiles);
      }
      ms[1] += k0;
      ms[2] = std::max(0.0, ms[2] - k0);
      ms[4] = b;
      ms[5] = f;
    }
    // copy_out: every get() since the last solve() recorded a pair after end_opt
    double out = 0;
    for(int e = end_opt + 1; e + 1 < n_events_used_; e += 2) out += el(e, e + 1);
    ms[7] = out;
  }

protected:
  /** Argument checks of DDPSolver::solve (DDPSolver.hpp:41-56, :391-414) and per-solve bookkeeping. */
  cudaStream_t beginSolve(int B, int n_u_steps, const double * x0, const double * u_init, void * stream)
  {
    const int N = cfg_.horizon_steps;
    // DDPSolver.hpp:41-45
    if(n_u_steps != N)
    {
      throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "initial_u_list length should be " + std::to_string(N) + " but "
                                                      + std::to_string(n_u_steps) + ".");
    }
    if(B <= 0 || B > capacity_)
    {
      throw Error(NMPC_B200_ERR_CAPACITY,
                  "batch " + std::to_string(B) + " outside (0, capacity " + std::to_string(capacity_) + "]");
    }
    if(x0 == nullptr || u_init == nullptr) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "null input array");
    // DDPSolver.hpp:391-414: the reference throws as soon as the backward pass runs
    if(cfg_.use_state_eq_second_derivative)
    {
      throw Error(NMPC_B200_ERR_RUNTIME, "Vector-tensor product is not implemented yet.");
    }
    if(cfg_.with_input_constraint && !have_limits_)
    {
      throw Error(NMPC_B200_ERR_RUNTIME, "with_input_constraint is set but no input limits were given");
    }
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : own_stream_;
    last_stream_ = st;
    B_ = B;
    ws_.B = B;
    return st;
  }

  /** Instance-major x0 / initial_u_list (host or device) -> x_list[0] and u_list of trajectory buffer 0. */
  void stageInputs(int B, const double * x0, const double * u_init, bool on_device, cudaStream_t st)
  {
    const int N = cfg_.horizon_steps;
    const double * d_x0 = x0;
    const double * d_u = u_init;
    if(!on_device)
    {
      NMPC_CUDA_CHECK(cudaMemcpyAsync(stage_in_x_.ptr, x0, sizeof(double) * B * NX, cudaMemcpyHostToDevice, st));
      NMPC_CUDA_CHECK(
          cudaMemcpyAsync(stage_in_u_.ptr, u_init, sizeof(double) * (size_t)B * N * NU, cudaMemcpyHostToDevice, st));
      d_x0 = stage_in_x_.ptr;
      d_u = stage_in_u_.ptr;
    }
    record(st); // 1: inputs on device
    launchScatterRows<double, S>(d_x0, ws_.x[0], B, NX, Bp_, st);
    launchScatterRows<double, S>(d_u, ws_.u[0], B, N * NU, Bp_, st);
  }

  /** K0, then max_iter x {K1, K2, K3} from the inputs staged in trajectory buffer 0. */
  void runIterations(int B, double current_t, cudaStream_t st)
  {
    const int N = cfg_.horizon_steps;
    prm_.t0 = S(current_t);
    launches_[0] = launches_[1] = launches_[2] = launches_[3] = 0;
    persistent_last_ = false;
    if constexpr(kLanesOk)
    {
      if(usePersistent(B))
      {
        launchSolveTile(B, st);
        return;
      }
    }
    const int tpb = threadsPerBlock(B);
    const int grid = (B + tpb - 1) / tpb;
    if(initialRolloutUsesSplit(B))
    {
      // the line search's rollout / cost / loader roles (ddp_forward_split.cuh): 39 -> 27 us at B = 4096
      using SL = SplitLayout<M>;
      const size_t smem = sizeof(S) * (SL::inElems(kTile) + SL::outElems(kTile))
                          + sizeof(unsigned long long) * 2 * (kSplitIn + kSplitOut) + 16;
      bool & attr_set = init_split_attr_set_;
      if(!attr_set)
      {
        NMPC_CUDA_CHECK(cudaFuncSetAttribute(rollout_init_split_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
      }
      launchPdl(rollout_init_split_kernel<M>, dim3((B + kTile - 1) / kTile), dim3(96), smem, st, model_, ws_, prm_);
    }
    else
      launchPdl(rollout_init_kernel<M>, dim3(grid), dim3(tpb), 0, st, model_, ws_, prm_);
    launches_[0]++;
    record(st); // 2: setup done

    const int tpb1 = 128;
    const dim3 grid1((B + tpb1 - 1) / tpb1, N + 1);
    iter_event_base_ = n_events_used_;
    iters_launched_ = 0;
    const bool fused = use_fused_; // Step 1 then happens inside the backward kernel
    for(int iter = 1; iter <= cfg_.max_iter; iter++)
    {
      if(!fused) launchPdl(linearize_kernel<M>, grid1, dim3(tpb1), 0, st, model_, ws_, prm_);
      record(st);
      launchBackward(B, tpb, grid, iter, st);
      record(st);
      launchForward(B, tpb, grid, iter, st); // records the event between the two phases of the line search itself
      record(st);
      if(!fused) launches_[1]++;
      launches_[2]++;
      launches_[3]++;
      iters_launched_ = iter;
      if(iter % kCheckStride == 0 && iter < cfg_.max_iter)
      {
        NMPC_CUDA_CHECK(cudaMemsetAsync(d_counter_.ptr, 0, sizeof(int), st));
        count_active_kernel<<<(B + 255) / 256, 256, 0, st>>>(ws_.status, B, d_counter_.ptr);
        NMPC_CUDA_CHECK(cudaMemcpyAsync(h_counter_, d_counter_.ptr, sizeof(int), cudaMemcpyDeviceToHost, st));
        NMPC_CUDA_CHECK(cudaStreamSynchronize(st));
        if(*h_counter_ == 0) break;
      }
    }
    record(st); // end of optimisation loop
    NMPC_CUDA_CHECK(cudaGetLastError());
  }

  /** One persistent CTA per 32-instance tile for the whole solve (ddp_solve_tile.cuh): while every tile has an SM of
      its own.  Beyond that the tiles would queue up behind whole solves and the stage kernels take over. */
  bool usePersistent(int B) const
  {
    return kLanesOk && use_fused_ && tune_.solve_tile != 0 && B <= tune_.solve_tile_max_batch;
  }

  void launchSolveTile(int B, cudaStream_t st)
  {
    if constexpr(kLanesOk)
    {
      using XS = XchSmem<S, LaneLayout<M>::G>;
      ensureFanout();
      const size_t smem = TileSmem<M>::bytes();
      const int tiles = (B + kTile - 1) / kTile;
      unsigned long long * stage_ns = nullptr;
      if(timing_)
      {
        stage_ns_.reserve(4 * (size_t)(cfg_.max_iter + 1));
        NMPC_CUDA_CHECK(cudaMemsetAsync(stage_ns_.ptr, 0, sizeof(unsigned long long) * 4 * (size_t)(cfg_.max_iter + 1), st));
        stage_ns = stage_ns_.ptr;
      }
      record(st); // 2: (the initial rollout is inside the kernel)
      iter_event_base_ = n_events_used_;
      iters_launched_ = 0;
      auto launch = [&](auto kernel, bool & attr_set) {
        if(!attr_set)
        {
          NMPC_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          attr_set = true;
        }
        launchPdl(kernel, dim3(tiles), dim3(kTileWarps * 32), smem, st, model_, ws_, prm_, fan_, stage_ns);
      };
      if(cfg_.with_input_constraint)
        launch(ddp_solve_tile_kernel<M, true, XS>, tile_attr_set_[1]);
      else
        launch(ddp_solve_tile_kernel<M, false, XS>, tile_attr_set_[0]);
      launches_[0] = 1;
      persistent_last_ = true;
      persistent_timed_ = timing_;
      record(st); // end of optimisation loop
      NMPC_CUDA_CHECK(cudaGetLastError());
    }
  }

  static constexpr int kEventsPerIter = 4; //!< after Step 1, Step 2, the first line-search candidate, the other candidates
  static constexpr int kMaxThreadsPerBlock = 128;
  static constexpr bool kHasBoxQP = true;
  static constexpr size_t kQuadSmemLimit = 200 * 1024; //!< shared memory the column-split K2 may use per CTA
  static constexpr int kPhased = 3; //!< K3 variant id: three-phase line search
  // group size of the cooperative K2: the power of two >= NX, capped at a warp
  static constexpr int kCoopGS = (NX <= 1) ? 1 : (NX <= 2) ? 2 : (NX <= 4) ? 4 : (NX <= 8) ? 8 : (NX <= 16) ? 16 : 32;

  /** K2 stages two derivative blocks per thread in shared memory (cp.async ring). */
  static constexpr size_t backwardSmemBytes(int tpb)
  {
    return 2 * sizeof(S) * (size_t)L::SIZE * tpb + 16 * (size_t)(tpb / 32) + 128;
  }
  // the thread-per-instance sweep needs that ring for at least one warp; for large n_u (centroidal motion, 9 x 16:
  // 374 KB per warp) it does not fit an SM's shared memory and the cooperative sweep is the only K2 variant
  static constexpr bool kThreadSweepFits = backwardSmemBytes(32) <= 227 * 1024;

  /** K3 variant: lanes per instance that evaluate line-search candidates concurrently.  Small
      batches cannot fill the 148 SMs with one thread per instance, so they spend lanes on speculation. */
  int forwardLanesPerInstance(int B) const
  {
    const int v = tune_.forward_lanes;
    if(v == 1 || v == 4 || v == 16 || v == kPhased) return v;
    // latency-bound regime: minimise the number of sequential rollouts.  Measured on B200 (cart-pole, M-fixed):
    // phased beats the in-warp fan-out up to B = 32768 (0.31 vs 0.37 ms) and loses at 131072 (1.18 vs 0.83 ms),
    // where evaluating all remaining candidates of every failed instance costs throughput.
    if(B <= tune_.forward_phased_max_batch) return kPhased;
    return 1;
  }

  /** K2 variant: lanes per instance (columns of the n_x x n_x matrices are spread over the group). */
  int backwardLanesPerInstance(int B) const
  {
    if(!kThreadSweepFits) return kCoopGS;
    const int v = tune_.backward_group_size;
    if(v == 1 || v == kCoopGS) return v;
    // One thread per instance keeps every matrix in registers and is the faster variant while they fit
    // (measured on B200, cart-pole 4x1, B=4096: 84 us vs 113 us per sweep).  From n_x = 8 on the register
    // file overflows (n_x = 12: 19 KB of spills per thread) and the cooperative variant takes over.
    if(NX < 8) return 1;
    return (B <= tune_.backward_coop_max_batch) ? kCoopGS : 1;
  }

  template<bool CONSTRAINED>
  void launchBackwardCoop(int B, int iter, cudaStream_t st)
  {
    using C = CoopLayout<M, kCoopGS>;
    constexpr int kWarps = 2;
    const int grid = (B + kWarps * C::IPW - 1) / (kWarps * C::IPW);
    const size_t smem = sizeof(S) * (size_t)kWarps * C::WARP_ELEMS;
    bool & attr_set = attr_set_[0 + (CONSTRAINED ? 1 : 0)]; // per engine: function attributes are per device
    if(!attr_set)
    {
      cudaFuncSetAttribute(backward_coop_kernel<M, kCoopGS, CONSTRAINED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)smem);
      attr_set = true;
    }
    launchPdl(backward_coop_kernel<M, kCoopGS, CONSTRAINED>, dim3(grid), dim3(kWarps * 32), smem, st, model_, ws_, prm_, iter);
  }

  /** K1 + K2 fused (producer warp + consumer warp per 32-instance tile, ddp_backward_fused.cuh): the default while the
      thread-per-instance sweep is the K2 variant in use, i.e. for n_x < 8.  The knob backward_fused = 0 restores the
      three-kernel pipeline. */
  bool backwardUsesFused(int B) const
  {
    if(NX >= 8) return false;
    if(tune_.backward_fused == 0) return false;
    if(tune_.backward_quad == 1) return false;
    if(tune_.backward_group_size > 1) return false;
    (void)B;
    return true;
  }

  template<bool CONSTRAINED>
  void launchBackwardFused(int B, int iter, cudaStream_t st)
  {
    const size_t smem = FusedLayout<M>::bytes();
    bool & attr_set = attr_set_[2 + (CONSTRAINED ? 1 : 0)]; // per engine: function attributes are per device
    if(!attr_set)
    {
      NMPC_CUDA_CHECK(cudaFuncSetAttribute(backward_fused_kernel<M, CONSTRAINED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
      attr_set = true;
    }
    launchPdl(backward_fused_kernel<M, CONSTRAINED>, dim3((B + kTile - 1) / kTile), dim3(64), smem, st, model_, ws_, prm_,
              iter);
  }

  /** K1 + K2 with G lanes per instance (ddp_backward_lanes.cuh): n_x <= 4, latency-bound batches. */
  static constexpr bool kLanesOk = NX <= 4;
  template<bool CONSTRAINED, int P, class XCH, int TPC>
  void launchBackwardLanesT(int B, int iter, cudaStream_t st, int slot)
  {
    if constexpr(kLanesOk)
    {
      using LL = LaneLayout<M>;
      bool & attr_set = lanes_attr_set_[slot];
      if(!attr_set)
      {
        NMPC_CUDA_CHECK(cudaFuncSetAttribute(backward_lanes_kernel<M, CONSTRAINED, P, XCH, TPC>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(TPC * LL::bytes())));
        attr_set = true;
      }
      const int tiles = (B + kTile - 1) / kTile;
      launchPdl(backward_lanes_kernel<M, CONSTRAINED, P, XCH, TPC>, dim3((tiles + TPC - 1) / TPC),
                dim3((LL::CW + P) * 32 * TPC), TPC * LL::bytes(), st, model_, ws_, prm_, iter);
    }
  }
  template<bool CONSTRAINED>
  void launchBackwardLanes(int B, int iter, cudaStream_t st)
  {
    if constexpr(kLanesOk)
    {
      using LL = LaneLayout<M>;
      using XS = XchSmem<S, LL::G>;
      using XH = XchShfl<S, LL::G>;
      const int c = CONSTRAINED ? 4 : 0;
      if(tune_.backward_lanes == 2)
      {
        if(tune_.backward_lanes_tiles_per_cta == 2) launchBackwardLanesT<CONSTRAINED, 2, XH, 2>(B, iter, st, c + 3);
        else launchBackwardLanesT<CONSTRAINED, 2, XH, 1>(B, iter, st, c + 2);
      }
      else
      {
        if(tune_.backward_lanes_tiles_per_cta == 2) launchBackwardLanesT<CONSTRAINED, 2, XS, 2>(B, iter, st, c + 1);
        else launchBackwardLanesT<CONSTRAINED, 2, XS, 1>(B, iter, st, c + 0);
      }
    }
  }

  /** K2 variant for latency-bound batches: four warps per 32-instance tile (ddp_backward_quad.cuh). */
  bool backwardUsesQuad(int B) const
  {
    if(QuadLayout<M>::bytes() > kQuadSmemLimit) return false;
    if(tune_.backward_quad == 0) return false;
    if(tune_.backward_quad == 1) return true;
    // Measured on B200: for n_x = 4 (cart-pole, B=4096) the three barriers per step cost what the split saves
    // (107 us vs 84 us per sweep; both variants are bound by one warp's dependent chain); for n_x = 12 (quadrotor
    // fp32, B=8192) the split beats both alternatives (14.5 ms vs 20.8 ms per 10 sweeps for the in-warp variant).
    if(NX < 8) return false;
    return B <= tune_.backward_quad_max_batch;
  }

  template<bool CONSTRAINED>
  void launchBackwardQuad(int B, int iter, cudaStream_t st)
  {
    const size_t smem = QuadLayout<M>::bytes();
    bool & attr_set = attr_set_[4 + (CONSTRAINED ? 1 : 0)]; // per engine: function attributes are per device
    if(!attr_set)
    {
      NMPC_CUDA_CHECK(cudaFuncSetAttribute(backward_quad_kernel<M, CONSTRAINED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
      attr_set = true;
    }
    launchPdl(backward_quad_kernel<M, CONSTRAINED>, dim3((B + kTile - 1) / kTile), dim3(QuadLayout<M>::W * 32), smem, st, model_,
              ws_, prm_, iter);
  }

  /** K2 variant for many inputs (n_u >= 8, no input limits): the n_u side spread over the lanes of a group
      (ddp_backward_wide.cuh).  Measured on B200, centroidal motion 9 x 16, B = 1024, one sweep: see DESIGN.md;
      the knob backward_wide = 0 falls back to the cooperative variant that recomputes the n_u x n_u part in every lane. */
  static constexpr int kWideGS = (NU <= 16 && NX < 16) ? 16 : 32;
  static constexpr bool kWideOk = NU >= 8 && NU <= kWideGS && NX < kWideGS;
  bool backwardUsesWide() const
  {
    return kWideOk && tune_.backward_wide != 0;
  }

  void launchBackwardWide(int B, int iter, cudaStream_t st)
  {
    if constexpr(kWideOk)
    {
      using C = WideLayout<M, kWideGS>;
      bool & attr_set = attr_set_[9]; // per engine: function attributes are per device
      if(!attr_set)
      {
        NMPC_CUDA_CHECK(cudaFuncSetAttribute(backward_wide_kernel<M, kWideGS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)C::bytes()));
        attr_set = true;
      }
      launchPdl(backward_wide_kernel<M, kWideGS>, dim3((B + C::IPW - 1) / C::IPW), dim3(32), C::bytes(), st, ws_, prm_, iter);
    }
  }

  void launchBackward(int B, int tpb, int grid, int iter, cudaStream_t st)
  {
    if constexpr(kWideOk)
    {
      if(!cfg_.with_input_constraint && backwardUsesWide())
      {
        launchBackwardWide(B, iter, st);
        return;
      }
    }
    if constexpr(NX < 8)
    {
      if(use_fused_ && kLanesOk && tune_.backward_lanes != 0 && B <= tune_.backward_lanes_max_batch)
      {
        if(cfg_.with_input_constraint)
          launchBackwardLanes<true>(B, iter, st);
        else
          launchBackwardLanes<false>(B, iter, st);
        return;
      }
      if(use_fused_)
      {
        if(cfg_.with_input_constraint)
          launchBackwardFused<true>(B, iter, st);
        else
          launchBackwardFused<false>(B, iter, st);
        return;
      }
    }
    if constexpr(QuadLayout<M>::bytes() <= kQuadSmemLimit)
    {
      if(backwardUsesQuad(B))
      {
        if(cfg_.with_input_constraint)
          launchBackwardQuad<true>(B, iter, st);
        else
          launchBackwardQuad<false>(B, iter, st);
        return;
      }
    }
    if(backwardLanesPerInstance(B) > 1)
    {
      if(cfg_.with_input_constraint)
        launchBackwardCoop<true>(B, iter, st);
      else
        launchBackwardCoop<false>(B, iter, st);
      return;
    }
    if(cfg_.with_input_constraint)
      launchPdl(backward_kernel<M, true>, dim3(grid), dim3(tpb), backwardSmemBytes(tpb), st, model_, ws_, prm_, iter);
    else
      launchPdl(backward_kernel<M, false>, dim3(grid), dim3(tpb), backwardSmemBytes(tpb), st, model_, ws_, prm_, iter);
  }

  template<int GA>
  void launchForwardSpec(int B, int iter, cudaStream_t st)
  {
    constexpr int kWarps = 4;
    constexpr int ipw = 32 / GA;
    const int grid = (B + kWarps * ipw - 1) / (kWarps * ipw);
    const size_t smem = sizeof(S) * (size_t)kWarps * 4 * FwdOperands<NX, NU>::SIZE * ipw;
    bool & attr_set = attr_set_[6 + (GA == 16 ? 1 : 0)]; // per engine: function attributes are per device
    if(!attr_set)
    {
      cudaFuncSetAttribute(forward_spec_kernel<M, GA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      attr_set = true;
    }
    launchPdl(forward_spec_kernel<M, GA>, dim3(grid), dim3(kWarps * 32), smem, st, model_, ws_, prm_, iter);
  }

  // the split rings (two steps per stage) must fit an SM's shared memory: not for many inputs (centroidal motion 9 x 16)
  static constexpr size_t kSplitFirstBytes = sizeof(S) * (SplitLayout<M>::inElems(kTile) + SplitLayout<M>::outElems(kTile)) + 256;
  static constexpr bool kSplitFits = kSplitFirstBytes <= 200 * 1024 && FanSmem<M, kFanSplitWarps>::bytes() <= 200 * 1024;

  /** K0 in the split-role form: same regime as the split line search (a functor with a time-varying input dimension
      keeps the thread-per-instance K0, whose padding rule it shares with the MPC-loop kernel). */
  bool initialRolloutUsesSplit(int B) const
  {
    return kSplitFits && !HasInputDim<M>::value && tune_.forward_split != 0 && B <= tune_.forward_split_max_batch
           && forwardLanesPerInstance(B) == kPhased;
  }

  /** Three-phase line search (ddp_forward_phased.cuh). */
  void launchForwardPhased(int B, int iter, cudaStream_t st)
  {
    using O = FwdOperands<NX, NU>;
    ensureFanout();
    const bool split = kSplitFits && tune_.forward_split != 0 && B <= tune_.forward_split_max_batch;
    if(split)
    {
      // one 32-instance tile per CTA: rollout warp + cost warp + loader warp (ddp_forward_split.cuh)
      using SL = SplitLayout<M>;
      const size_t smem = sizeof(S) * (SL::inElems(kTile) + SL::outElems(kTile))
                          + sizeof(unsigned long long) * 2 * (kSplitIn + kSplitOut) + 16;
      bool & attr_set = lanes_attr_set_[10];
      if(!attr_set)
      {
        NMPC_CUDA_CHECK(cudaFuncSetAttribute(forward_first_split_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
      }
      launchPdl(forward_first_split_kernel<M>, dim3((B + kTile - 1) / kTile), dim3(96), smem, st, model_, ws_, prm_, fan_, iter);
    }
    else
    {
      // one 32-instance tile per CTA (compute warp + loader warp): 4096 instances cover 128 SMs
      const size_t smem = sizeof(S) * (size_t)kFirstDepth * O::SIZE * kTile + sizeof(unsigned long long) * 2 * kFirstDepth + 16;
      bool & attr_set = attr_set_[8]; // per engine: function attributes are per device
      if(!attr_set)
      {
        cudaFuncSetAttribute(forward_first_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
      }
      launchPdl(forward_first_kernel<M>, dim3((B + kTile - 1) / kTile), dim3(64), smem, st, model_, ws_, prm_, fan_, iter);
    }
    record(st); // first candidate done
    if(split)
    {
      using FS = FanSmem<M, kFanSplitWarps>;
      constexpr int ipc = FS::IPC; // listed instances per CTA
      const size_t smem = FS::bytes();
      bool & attr_set = lanes_attr_set_[11];
      if(!attr_set)
      {
        NMPC_CUDA_CHECK(cudaFuncSetAttribute(forward_fanout_split_kernel<M, kFanSplitWarps>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
      }
      const int grid = (B + ipc - 1) / ipc; // worst case: every instance listed
      launchPdl(forward_fanout_split_kernel<M, kFanSplitWarps>, dim3(grid), dim3((2 * kFanSplitWarps + 1) * 32), smem, st, model_,
                ws_, prm_, fan_, iter);
    }
    else
    {
      constexpr int ipc = kFanWarps * (32 / kFanLanes); // listed instances per CTA
      const size_t smem = sizeof(S) * (size_t)kFanDepth * ipc * O::SIZE + sizeof(unsigned long long) * 2 * kFanDepth + 16;
      const int grid = (B + ipc - 1) / ipc; // worst case: every instance listed
      launchPdl(forward_fanout_kernel<M>, dim3(grid), dim3((kFanWarps + 1) * 32), smem, st, model_, ws_, prm_, fan_, iter);
    }
  }

  void ensureFanout()
  {
    const size_t N = cfg_.horizon_steps;
    const size_t items = (size_t)Bp_ * kFanLanes;
    if(fan_.items == items && fan_scratch_.count == items * ((N + 1) * NX + N * NU + (N + 1))) return;
    fan_scratch_.allocate(items * ((N + 1) * NX + N * NU + (N + 1)));
    fan_ints_.allocate(2 * (size_t)Bp_);
    fan_.count = d_fan_count_.ptr;
    fan_.list = fan_ints_.ptr;
    fan_.commit_item = fan_ints_.ptr + Bp_;
    fan_.sx = fan_scratch_.ptr;
    fan_.su = fan_.sx + items * (N + 1) * NX;
    fan_.sc = fan_.su + items * N * NU;
    fan_.items = items;
  }

  void launchForward(int B, int tpb, int grid, int iter, cudaStream_t st)
  {
    switch(forwardLanesPerInstance(B))
    {
      case 16:
        launchForwardSpec<16>(B, iter, st);
        break;
      case 4:
        launchForwardSpec<4>(B, iter, st);
        break;
      case kPhased:
        launchForwardPhased(B, iter, st);
        return;
      default:
        launchPdl(forward_kernel<M>, dim3(grid), dim3(tpb), 0, st, model_, ws_, prm_, iter);
    }
    record(st); // single-kernel line search: the whole of it counts as "first candidate"
  }

  /** Largest CTA whose two-stage block ring fits the 227 KB of shared memory an SM offers. */
  static int maxThreadsPerBlock()
  {
    int tpb = kMaxThreadsPerBlock;
    while(tpb > 32 && backwardSmemBytes(tpb) > 200 * 1024) tpb -= 32;
    return tpb;
  }

  int threadsPerBlock(int B) const
  {
    const int cap = maxThreadsPerBlock();
    const int v = tune_.threads_per_block;
    if(v >= 32 && v <= cap && v % 32 == 0) return v;
    // spread small batches over all SMs: one warp per CTA until every SM has a few warps
    if(B <= tune_.sm_count * 32 * 2) return 32;
    if(B <= tune_.sm_count * 64 * 4) return std::min(64, cap);
    return std::min(128, cap);
  }

  void record(cudaStream_t st)
  {
    if(!timing_) return;
    if(n_events_used_ >= (int)events_.size())
    {
      cudaEvent_t e;
      NMPC_CUDA_CHECK(cudaEventCreate(&e));
      events_.push_back(e);
    }
    NMPC_CUDA_CHECK(cudaEventRecord(events_[n_events_used_], st));
    n_events_used_++;
  }

  double * stageOut(size_t bytes)
  {
    if(stage_out_.bytes() < bytes) stage_out_.allocate((bytes + sizeof(double) - 1) / sizeof(double));
    return stage_out_.ptr;
  }

  void applyConfig(const nmpc_b200_ddp_config & cfg, bool first)
  {
    if(cfg.horizon_steps <= 0) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "horizon_steps must be positive");
    if(cfg.max_iter < 0) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "max_iter must be non-negative");
    if(cfg.n_alpha < 0 || cfg.n_alpha > kMaxAlpha)
      throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "alpha_list longer than " + std::to_string(kMaxAlpha));
    if(cfg.reg_type != 1 && cfg.reg_type != 2 && cfg.reg_type != 0)
      throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "reg_type must be 1 or 2");
    const bool realloc_needed = first || cfg.horizon_steps != cfg_.horizon_steps || cfg.max_iter != cfg_.max_iter;
    cfg_ = cfg;
    prm_.N = cfg.horizon_steps;
    prm_.max_iter = cfg.max_iter;
    prm_.reg_type = cfg.reg_type;
    prm_.with_input_constraint = cfg.with_input_constraint;
    prm_.n_alpha = cfg.n_alpha;
    prm_.initial_lambda = S(cfg.initial_lambda);
    prm_.initial_dlambda = S(cfg.initial_dlambda);
    prm_.lambda_factor = S(cfg.lambda_factor);
    prm_.lambda_min = S(cfg.lambda_min);
    prm_.lambda_max = S(cfg.lambda_max);
    prm_.k_rel_norm_thre = S(cfg.k_rel_norm_thre);
    prm_.lambda_thre = S(cfg.lambda_thre);
    prm_.cost_update_ratio_thre = S(cfg.cost_update_ratio_thre);
    prm_.cost_update_thre = S(cfg.cost_update_thre);
    for(int i = 0; i < kMaxAlpha; i++) prm_.alpha_list[i] = (i < cfg.n_alpha) ? S(cfg.alpha_list[i]) : S(0);
    if(realloc_needed) allocate();
  }

  void allocate()
  {
    const size_t N = cfg_.horizon_steps;
    const size_t Bp = Bp_;
    NMPC_CUDA_CHECK(cudaStreamSynchronize(own_stream_));
    for(int s = 0; s < 2; s++)
    {
      x_[s].allocate((N + 1) * NX * Bp);
      u_[s].allocate(N * NU * Bp);
      cost_[s].allocate((N + 1) * Bp);
      ws_.x[s] = x_[s].ptr;
      ws_.u[s] = u_[s].ptr;
      ws_.cost[s] = cost_[s].ptr;
    }
    // the K1 -> K2 derivative tiles and terminal derivatives exist only in the three-kernel pipeline (4.8 GB at
    // B = 131072 for cart-pole); the fused K1+K2 keeps them in shared memory.  The variant is fixed per engine.
    use_fused_ = backwardUsesFused(capacity_);
    if(use_fused_)
    {
      deriv_.release();
      vterm_.release();
    }
    else
    {
      deriv_.allocate(N * L::SIZE * Bp);
      vterm_.allocate((size_t)(NX + NX * NX) * Bp);
    }
    kff_.allocate(N * NU * Bp);
    kfb_.allocate(N * NU * NX * Bp);
    trace_.allocate((size_t)(cfg_.max_iter + 1) * kTraceFields * Bp);
    scal_.allocate(5 * Bp);
    ints_.allocate(5 * Bp);
    d_u_lo_.allocate(NU > 0 ? N * NU : 1);
    d_u_hi_.allocate(NU > 0 ? N * NU : 1);
    d_counter_.allocate(1);
    d_fan_count_.allocate(1);
    NMPC_CUDA_CHECK(cudaMemset(d_fan_count_.ptr, 0, sizeof(int)));
    ws_.fan_count = d_fan_count_.ptr;
    fan_ = FwdFanout<S>{};
    fan_scratch_.release();
    stage_in_x_.allocate((size_t)capacity_ * NX);
    stage_in_u_.allocate((size_t)capacity_ * N * NU);
    NMPC_CUDA_CHECK(cudaMemset(scal_.ptr, 0, scal_.bytes()));
    NMPC_CUDA_CHECK(cudaMemset(ints_.ptr, 0, ints_.bytes()));
    ws_.Bp = Bp_;
    ws_.B = 0;
    ws_.deriv = deriv_.ptr;
    ws_.vterm = vterm_.ptr;
    ws_.kff = kff_.ptr;
    ws_.kfb = kfb_.ptr;
    ws_.trace = trace_.ptr;
    ws_.lambda = scal_.ptr;
    ws_.dlambda = scal_.ptr + Bp;
    ws_.cost_sum = scal_.ptr + 2 * Bp;
    ws_.dV = scal_.ptr + 3 * Bp;
    ws_.u_lo = d_u_lo_.ptr;
    ws_.u_hi = d_u_hi_.ptr;
    ws_.status = ints_.ptr;
    ws_.sel = ints_.ptr + Bp;
    ws_.iters = ints_.ptr + 2 * Bp;
    ws_.n_fwd = ints_.ptr + 3 * Bp;
    ws_.n_bwd = ints_.ptr + 4 * Bp;
    if(have_limits_)
    {
      // a changed horizon keeps constant limits; step-wise limits have to be given again for the new horizon
      if(limits_vary_ || u_lo_.size() < (size_t)NU)
        have_limits_ = false;
      else
      {
        const std::vector<double> lo(u_lo_.begin(), u_lo_.begin() + NU), hi(u_hi_.begin(), u_hi_.begin() + NU);
        setInputLimits(lo.data(), hi.data());
      }
    }
    B_ = 0;
  }

  M model_;
  int device_;
  int capacity_;
  int Bp_ = 0;
  int B_ = 0;
  nmpc_b200_ddp_config cfg_{};
  SolverParams<S> prm_{};
  Workspace<S> ws_{};
  cudaStream_t own_stream_ = nullptr;
  cudaStream_t last_stream_ = nullptr;
  DeviceBuffer<S> x_[2], u_[2], cost_[2], deriv_, vterm_, kff_, kfb_, trace_, scal_, d_u_lo_, d_u_hi_;
  DeviceBuffer<int> ints_, d_counter_, d_fan_count_, fan_ints_;
  DeviceBuffer<S> fan_scratch_, mpc_x_, mpc_u_, mpc_lo_, mpc_hi_;
  int mpc_limit_ticks_ = 0; //!< ticks covered by the per-tick limit tables (setInputLimitsMpc)
  DeviceBuffer<int> mpc_i_;
  FwdFanout<S> fan_{};
  DeviceBuffer<double> stage_in_x_, stage_in_u_, stage_out_;
  int * h_counter_ = nullptr;
  std::vector<double> u_lo_, u_hi_;
  bool have_limits_ = false;
  bool attr_set_[10] = {false, false, false, false, false, false, false, false, false, false};
  bool use_fused_ = false; //!< K1 fused into K2 (decided once, at allocation)
  bool lanes_attr_set_[12] = {};
  bool init_split_attr_set_ = false;
  DdpTuning tune_;
  bool tile_attr_set_[2] = {false, false};
  bool persistent_last_ = false; //!< the last solve ran in the persistent kernel
  bool persistent_timed_ = false;
  DeviceBuffer<unsigned long long> stage_ns_;
  bool limits_vary_ = false; //!< the limits differ between horizon steps
  bool timing_ = false;
  std::vector<cudaEvent_t> events_;
  int n_events_used_ = 0;
  int iter_event_base_ = 0;
  int iters_launched_ = 0;
  int launches_[4] = {0, 0, 0, 0};
};
} // namespace ddp
} // namespace nmpc_b200
