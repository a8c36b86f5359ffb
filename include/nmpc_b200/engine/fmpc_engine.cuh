/* nmpc_b200 -- host side of the batched FMPC engine for one functor type M (see fmpc_kernels.cuh). */
#pragma once

#include <algorithm>
#include <cstdlib>
#include <memory>
#include <vector>

#include "common.cuh"
#include "fmpc_kernels.cuh"
#include "fmpc_mpc.cuh"
#include "registry.h"

namespace nmpc_b200
{
namespace fmpc
{
template<class S>
__global__ void extract_first_rows_kernel(const S * src, double * dst, int B, int R, int Bp)
{
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if(b >= B) return;
  for(int d = 0; d < R; d++) dst[(size_t)b * R + d] = double(src[(size_t)d * Bp + b]);
}

template<class M>
class FmpcEngine : public FmpcEngineBase
{
public:
  using S = typename M::Scalar;
  static constexpr int NX = M::NX;
  static constexpr int NU = M::NU;
  static constexpr int NG = M::NG;
  using L = CoeffLayout<NX, NU, NG>;

  FmpcEngine(const double * params, const nmpc_b200_fmpc_config & cfg, int batch_capacity, int device)
  : model_(M::fromParams(params)), device_(device), capacity_(batch_capacity)
  {
    if(batch_capacity <= 0) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "batch_capacity must be positive");
    DeviceGuard guard(device_);
    NMPC_CUDA_CHECK(cudaStreamCreateWithFlags(&own_stream_, cudaStreamNonBlocking));
    NMPC_CUDA_CHECK(cudaMallocHost(reinterpret_cast<void **>(&h_flag_), sizeof(int)));
    Bp_ = ((capacity_ + 127) / 128) * 128;
    NMPC_CUDA_CHECK(cudaDeviceGetAttribute(&sm_count_, cudaDevAttrMultiProcessorCount, device_));
    applyConfig(cfg, true);
  }

  ~FmpcEngine() override
  {
    DeviceGuard guard(device_);
    cudaStreamSynchronize(own_stream_);
    for(auto e : events_) cudaEventDestroy(e);
    cudaStreamDestroy(own_stream_);
    cudaFreeHost(h_flag_);
  }

  void setConfig(const nmpc_b200_fmpc_config & cfg) override
  {
    DeviceGuard guard(device_);
    applyConfig(cfg, false);
  }

  void solve(int B,
             double current_t,
             const double * x0,
             const double * x,
             const double * u,
             const double * lambda,
             const double * s,
             const double * nu,
             int n_steps,
             bool on_device,
             void * stream) override
  {
    DeviceGuard guard(device_);
    cudaStream_t st = beginSolve(B, x0, x, u, lambda, s, nu, n_steps, stream);
    n_events_used_ = 0;
    record(st); // 0
    stageInputs(B, x0, x, u, lambda, s, nu, on_device, st);
    prm_.keep_barrier_eps = 0;
    runIterations(B, current_t, st, true);
  }

  /** The reference's FMPC loops for the whole batch, tick after tick on the device (c_api.h, fmpc_mpc.cuh). */
  void runMpc(int B,
              double current_t,
              const double * x0,
              const double * x,
              const double * u,
              const double * lambda,
              const double * s,
              const double * nu,
              int n_steps,
              const nmpc_b200_mpc_config & mpc,
              double * x_log,
              double * u_log,
              double * kkt_log,
              int * status_log,
              bool on_device,
              void * stream) override
  {
    DeviceGuard guard(device_);
    if(mpc.n_ticks <= 0) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "n_ticks must be positive");
    if(mpc.plant != 0 && mpc.plant != 1) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "plant must be 0 or 1");
    if(mpc.plant == 1 && !ddp::HasStateEqDt<M>::value)
      throw Error(NMPC_B200_ERR_UNSUPPORTED, "plant = 1 needs a functor with stateEq(t, x, u, dt)");
    if(mpc.plant == 1 && mpc.n_substeps <= 0) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "n_substeps must be positive");
    if(mpc.shift_inputs || mpc.clamp_u0)
      throw Error(NMPC_B200_ERR_INVALID_ARGUMENT,
                  "shift_inputs / clamp_u0 have no FMPC counterpart: the Variable is passed on as it is "
                  "(TestFmpcOscillator.cpp:189)");
    cudaStream_t st = beginSolve(B, x0, x, u, lambda, s, nu, n_steps, stream);
    const size_t T = mpc.n_ticks, Bp = Bp_;
    MpcLogs<S> logs{};
    if(x_log) logs.x = (mpc_x_.reserve((T + 1) * NX * Bp), mpc_x_.ptr);
    if(u_log) logs.u = (mpc_u_.reserve(T * NU * Bp), mpc_u_.ptr);
    if(kkt_log) logs.kkt = (mpc_k_.reserve(T * Bp), mpc_k_.ptr);
    if(status_log) logs.status = (mpc_i_.reserve(T * Bp), mpc_i_.ptr);
    ddp::MpcParams<S> mp{};
    mp.n_ticks = mpc.n_ticks;
    mp.plant = mpc.plant;
    mp.n_substeps = mpc.n_substeps;
    mp.tick_dt = S(mpc.tick_dt);
    mp.sim_dt = S(mpc.sim_dt);

    n_events_used_ = 0;
    record(st);
    stageInputs(B, x0, x, u, lambda, s, nu, on_device, st);
    for(int tick = 0; tick < mpc.n_ticks; tick++)
    {
      const double t = current_t + tick * mpc.tick_dt;
      if(tick > 0)
      {
        n_events_used_ = 0;
        record(st);
        record(st);
      }
      // barrier_eps_ is a member that persists across solve() calls (FmpcSolver.h:413-414)
      prm_.keep_barrier_eps = (tick > 0) ? 1 : 0;
      // the non-negativity check of checkVariable() waits for the device only at the first tick; the later
      // warm starts are the solver's own output and are checked once, after the loop
      runIterations(B, t, st, tick == 0);
      fmpc_mpc_advance_kernel<M><<<(B + 127) / 128, 128, 0, st>>>(model_, ws_, mp, logs, mpc.feedback, tick, S(t));
      NMPC_CUDA_CHECK(cudaGetLastError());
    }
    prm_.keep_barrier_eps = 0;
    NMPC_CUDA_CHECK(cudaMemcpyAsync(h_flag_, d_flag_.ptr, sizeof(int), cudaMemcpyDeviceToHost, st));
    auto out_f64 = [&](const S * src, double * dst, int R) {
      if(dst == nullptr) return;
      const size_t need = sizeof(double) * (size_t)B * R;
      double * d_out = dst;
      if(!on_device)
      {
        if(stage_out_.bytes() < need) stage_out_.allocate(need / sizeof(double) + 1);
        d_out = stage_out_.ptr;
      }
      launchGatherRows<S, double>(src, src, nullptr, nullptr, 0, 1, d_out, B, R, Bp_, st);
      if(!on_device)
      {
        NMPC_CUDA_CHECK(cudaMemcpyAsync(dst, d_out, need, cudaMemcpyDeviceToHost, st));
        NMPC_CUDA_CHECK(cudaStreamSynchronize(st));
      }
    };
    out_f64(logs.x, x_log, (mpc.n_ticks + 1) * NX);
    out_f64(logs.u, u_log, mpc.n_ticks * NU);
    out_f64(logs.kkt, kkt_log, mpc.n_ticks);
    if(status_log)
    {
      const size_t need = sizeof(int) * (size_t)B * mpc.n_ticks;
      int * d_out = status_log;
      if(!on_device)
      {
        if(stage_out_.bytes() < need) stage_out_.allocate(need / sizeof(double) + 1);
        d_out = reinterpret_cast<int *>(stage_out_.ptr);
      }
      launchGatherRows<int, int>(logs.status, logs.status, nullptr, nullptr, 0, 1, d_out, B, mpc.n_ticks, Bp_, st);
      if(!on_device) NMPC_CUDA_CHECK(cudaMemcpyAsync(status_log, d_out, need, cudaMemcpyDeviceToHost, st));
    }
    NMPC_CUDA_CHECK(cudaStreamSynchronize(st));
    NMPC_CUDA_CHECK(cudaGetLastError());
    if(*h_flag_ != 0)
    {
      throw Error(NMPC_B200_ERR_RUNTIME, "[FMPC] s_list[i] / nu_list[i] must be non-negative.");
    }
  }

  void get(int what, void * dst, size_t dst_bytes, bool dst_on_device, void * stream) override
  {
    DeviceGuard guard(device_);
    if(B_ <= 0) throw Error(NMPC_B200_ERR_RUNTIME, "get() before solve()");
    if(dst == nullptr) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "null destination");
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : last_stream_;
    const int N = cfg_.horizon_steps;
    const int B = B_;
    const S * src = nullptr;
    const int * isrc = nullptr;
    const int * row_limit = nullptr;
    int R = 0;
    switch(what)
    {
      case NMPC_B200_FMPC_X:
        src = ws_.x, R = (N + 1) * NX;
        break;
      case NMPC_B200_FMPC_U:
        src = ws_.u, R = N * NU;
        break;
      case NMPC_B200_FMPC_LAMBDA:
        src = ws_.lam, R = (N + 1) * NX;
        break;
      case NMPC_B200_FMPC_S:
        src = ws_.s, R = N * NG;
        break;
      case NMPC_B200_FMPC_NU:
        src = ws_.nu, R = N * NG;
        break;
      case NMPC_B200_FMPC_K_FF:
        src = ws_.kff, R = N * NU;
        break;
      case NMPC_B200_FMPC_K_FB:
        src = ws_.kfb, R = N * NU * NX;
        break;
      case NMPC_B200_FMPC_TRACE:
        src = ws_.trace, R = cfg_.max_iter * kTraceFields, row_limit = ws_.n_trace;
        break;
      case NMPC_B200_FMPC_U0:
        src = ws_.u, R = NU;
        break;
      case NMPC_B200_FMPC_STATUS:
        isrc = ws_.status;
        break;
      case NMPC_B200_FMPC_N_TRACE:
        isrc = ws_.n_trace;
        break;
      default:
        throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "unknown FMPC field " + std::to_string(what));
    }
    const size_t need = isrc ? sizeof(int) * (size_t)B : sizeof(double) * (size_t)B * R;
    if(dst_bytes < need)
      throw Error(NMPC_B200_ERR_INVALID_ARGUMENT,
                  "destination too small: " + std::to_string(dst_bytes) + " < " + std::to_string(need));
    if(isrc)
    {
      NMPC_CUDA_CHECK(
          cudaMemcpyAsync(dst, isrc, need, dst_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
      if(!dst_on_device) NMPC_CUDA_CHECK(cudaStreamSynchronize(st));
      return;
    }
    double * d_out = static_cast<double *>(dst);
    if(!dst_on_device)
    {
      if(stage_out_.bytes() < need) stage_out_.allocate(need / sizeof(double) + 1);
      d_out = stage_out_.ptr;
    }
    if(what == NMPC_B200_FMPC_U0)
      extract_first_rows_kernel<S><<<(B + 127) / 128, 128, 0, st>>>(src, d_out, B, R, Bp_);
    else
      launchGatherRows<S, double>(src, src, nullptr, row_limit, 0, kTraceFields, d_out, B, R, Bp_, st);
    NMPC_CUDA_CHECK(cudaGetLastError());
    if(!dst_on_device)
    {
      NMPC_CUDA_CHECK(cudaMemcpyAsync(dst, d_out, need, cudaMemcpyDeviceToHost, st));
      NMPC_CUDA_CHECK(cudaStreamSynchronize(st));
    }
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
    const int end_opt = iter_event_base_ + 4 * iters_launched_;
    ms[7] = el(0, 1);
    ms[1] = el(1, 2);
    for(int it = 0; it < iters_launched_; it++)
    {
      const int e = iter_event_base_ + 4 * it;
      ms[3] += el(e - 1, e);
      ms[4] += el(e, e + 1);
      ms[5] += el(e + 1, e + 2);
      ms[6] += el(e + 2, e + 3);
    }
    ms[2] = el(2, end_opt);
    ms[0] = el(0, end_opt);
  }

protected:
  /** checkVariable() sequence lengths (FmpcSolver.hpp:288-312) and per-solve bookkeeping. */
  cudaStream_t beginSolve(int B,
                          const double * x0,
                          const double * x,
                          const double * u,
                          const double * lambda,
                          const double * s,
                          const double * nu,
                          int n_steps,
                          void * stream)
  {
    const int N = cfg_.horizon_steps;
    if(n_steps != N)
    {
      throw Error(NMPC_B200_ERR_INVALID_ARGUMENT,
                  "[FMPC] u_list length should be " + std::to_string(N) + " but " + std::to_string(n_steps) + ".");
    }
    if(B <= 0 || B > capacity_)
    {
      throw Error(NMPC_B200_ERR_CAPACITY,
                  "batch " + std::to_string(B) + " outside (0, capacity " + std::to_string(capacity_) + "]");
    }
    if(!x0 || !x || !u || !lambda || !s || !nu) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "null input array");
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : own_stream_;
    last_stream_ = st;
    B_ = B;
    ws_.B = B;
    return st;
  }

  void stageInputs(int B,
                   const double * x0,
                   const double * x,
                   const double * u,
                   const double * lambda,
                   const double * s,
                   const double * nu,
                   bool on_device,
                   cudaStream_t st)
  {
    const int N = cfg_.horizon_steps;
    const size_t nx1 = (size_t)(N + 1) * NX, nun = (size_t)N * NU, ngn = (size_t)N * NG;
    const double * srcs[6] = {x0, x, u, lambda, s, nu};
    const size_t rows[6] = {(size_t)NX, nx1, nun, nx1, ngn, ngn};
    S * dsts[6] = {ws_.x0, ws_.x, ws_.u, ws_.lam, ws_.s, ws_.nu};
    size_t off = 0;
    for(int a = 0; a < 6; a++)
    {
      const double * d_src = srcs[a];
      if(!on_device)
      {
        NMPC_CUDA_CHECK(
            cudaMemcpyAsync(stage_in_.ptr + off, srcs[a], sizeof(double) * B * rows[a], cudaMemcpyHostToDevice, st));
        d_src = stage_in_.ptr + off;
        off += (size_t)capacity_ * rows[a];
      }
      launchScatterRows<double, S>(d_src, dsts[a], B, (int)rows[a], Bp_, st);
    }
    record(st); // 1: inputs in device layout
    NMPC_CUDA_CHECK(cudaMemsetAsync(d_flag_.ptr, 0, sizeof(int), st));
  }

  /** F0, then max_iter x {F1, F2, F3, F4} on the Variable resident in the workspace. */
  void runIterations(int B, double current_t, cudaStream_t st, bool wait_for_check)
  {
    const int N = cfg_.horizon_steps;
    prm_.t0 = S(current_t);
    for(int & l : launches_) l = 0;
    const int tpb = threadsPerBlock(B);
    const int grid = (B + tpb - 1) / tpb;
    const int tpb1 = 128;
    const dim3 gridN((B + tpb1 - 1) / tpb1, N), gridN1((B + tpb1 - 1) / tpb1, N + 1);

    // the sweeps run one compute warp + one loader warp per 32-instance tile (fmpc_kernels.cuh)
    using R2 = BackwardRows<NX, NU, NG>;
    using R3 = ForwardRows<NX, NU, NG>;
    const int tpb2 = 64, tpb3 = 64, grid2 = (B + kTile - 1) / kTile, grid3 = grid2;
    const size_t smem2 = R2::bytes(sizeof(S)), smem3 = R3::bytes(sizeof(S));
    bool & attr_set = attr_set_[0]; // per engine: function attributes are per device
    if(!attr_set)
    {
      NMPC_CUDA_CHECK(cudaFuncSetAttribute(fmpc_backward_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      NMPC_CUDA_CHECK(cudaFuncSetAttribute(fmpc_forward_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
      attr_set = true;
    }

    fmpc_init_kernel<M><<<gridN, tpb1, 0, st>>>(model_, ws_, prm_);
    if(wait_for_check)
    {
      // checkVariable(): s, nu must be non-negative (FmpcSolver.hpp:348-361) -- the reference throws, so
      // this is the one place where solve() waits for the device
      NMPC_CUDA_CHECK(cudaMemcpyAsync(h_flag_, d_flag_.ptr, sizeof(int), cudaMemcpyDeviceToHost, st));
      NMPC_CUDA_CHECK(cudaStreamSynchronize(st));
      if(*h_flag_ != 0)
      {
        B_ = 0;
        throw Error(NMPC_B200_ERR_RUNTIME, "[FMPC] s_list[i] / nu_list[i] must be non-negative.");
      }
    }
    record(st); // 2: setup done

    iter_event_base_ = n_events_used_;
    iters_launched_ = 0;
    for(int iter = 1; iter <= cfg_.max_iter; iter++)
    {
      fmpc_coeff_kernel<M><<<gridN1, tpb1, 0, st>>>(model_, ws_, prm_);
      record(st);
      fmpc_backward_kernel<M><<<grid2, tpb2, smem2, st>>>(model_, ws_, prm_, iter);
      record(st);
      fmpc_forward_kernel<M><<<grid3, tpb3, smem3, st>>>(model_, ws_, prm_, iter);
      if(cfg_.enable_line_search) fmpc_linesearch_kernel<M><<<grid, tpb, 0, st>>>(model_, ws_, prm_, iter);
      record(st);
      fmpc_update_kernel<M><<<gridN1, tpb1, 0, st>>>(ws_, prm_);
      record(st);
      for(int & l : launches_) l++;
      iters_launched_ = iter;
    }
    fmpc_finalize_kernel<<<(B + 255) / 256, 256, 0, st>>>(ws_.status, B);
    record(st);
    NMPC_CUDA_CHECK(cudaGetLastError());
  }

  /** One warp per CTA until every SM has a few warps, then larger CTAs (NMPC_B200_THREADS_PER_BLOCK pins it). */
  int threadsPerBlock(int B) const
  {
    static const int pinned = [] {
      const char * env = std::getenv("NMPC_B200_THREADS_PER_BLOCK");
      if(env == nullptr) env = std::getenv("NMPC_B200_TPB");
      return env != nullptr ? std::atoi(env) : -1;
    }();
    if(pinned >= 32 && pinned <= 128 && pinned % 32 == 0) return pinned;
    if(B <= sm_count_ * 32 * 2) return 32;
    if(B <= sm_count_ * 64 * 4) return 64;
    return 128;
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

  void applyConfig(const nmpc_b200_fmpc_config & cfg, bool first)
  {
    if(cfg.horizon_steps <= 0) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "horizon_steps must be positive");
    if(cfg.max_iter < 0) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "max_iter must be non-negative");
    const bool realloc_needed = first || cfg.horizon_steps != cfg_.horizon_steps || cfg.max_iter != cfg_.max_iter;
    cfg_ = cfg;
    prm_.N = cfg.horizon_steps;
    prm_.max_iter = cfg.max_iter;
    prm_.check_nan = cfg.check_nan;
    prm_.init_complementary_variable = cfg.init_complementary_variable;
    prm_.update_barrier_eps = cfg.update_barrier_eps;
    prm_.break_if_llt_fails = cfg.break_if_llt_fails;
    prm_.merit_const_scale_from_lagrange_multipliers = cfg.merit_const_scale_from_lagrange_multipliers;
    prm_.keep_barrier_eps = 0;
    prm_.kkt_error_thre = S(cfg.kkt_error_thre);
    prm_.initial_barrier_eps = S(cfg.initial_barrier_eps);
    if(realloc_needed) allocate();
  }

  void allocate()
  {
    const size_t N = cfg_.horizon_steps;
    const size_t Bp = Bp_;
    NMPC_CUDA_CHECK(cudaStreamSynchronize(own_stream_));
    const size_t nx1 = (N + 1) * NX, nun = N * NU, ngn = N * NG;
    // one slab: x0, x, u, lam, s, nu, dx, du, dlam, ds, dnu
    const size_t var_elems = NX + 2 * (nx1 + nun + nx1 + ngn + ngn);
    vars_.allocate(var_elems * Bp);
    S * p = vars_.ptr;
    auto take = [&](size_t rows) {
      S * r = p;
      p += rows * Bp;
      return r;
    };
    ws_.x0 = take(NX);
    ws_.x = take(nx1);
    ws_.u = take(nun);
    ws_.lam = take(nx1);
    ws_.s = take(ngn);
    ws_.nu = take(ngn);
    ws_.dx = take(nx1);
    ws_.du = take(nun);
    ws_.dlam = take(nx1);
    ws_.ds = take(ngn);
    ws_.dnu = take(ngn);
    coeff_.allocate(N * L::SIZE * Bp);
    term_.allocate((size_t)L::T_SIZE * Bp);
    gains_.allocate((nun + nun * NX + nx1 + nx1 * NX) * Bp);
    kkt_.allocate((N + 2) * Bp);
    trace_.allocate((size_t)(cfg_.max_iter > 0 ? cfg_.max_iter : 1) * kTraceFields * Bp);
    scal_.allocate(3 * Bp);
    ints_.allocate(2 * Bp);
    d_flag_.allocate(1);
    stage_in_.allocate((size_t)capacity_ * (NX + nx1 + nun + nx1 + ngn + ngn));
    NMPC_CUDA_CHECK(cudaMemset(scal_.ptr, 0, scal_.bytes()));
    NMPC_CUDA_CHECK(cudaMemset(ints_.ptr, 0, ints_.bytes()));
    ws_.Bp = Bp_;
    ws_.B = 0;
    ws_.coeff = coeff_.ptr;
    ws_.term = term_.ptr;
    ws_.kff = gains_.ptr;
    ws_.kfb = ws_.kff + nun * Bp;
    ws_.sv = ws_.kfb + nun * NX * Bp;
    ws_.P = ws_.sv + nx1 * Bp;
    ws_.kkt = kkt_.ptr;
    ws_.trace = trace_.ptr;
    ws_.barrier_eps = scal_.ptr;
    ws_.alpha = scal_.ptr + Bp;
    ws_.status = ints_.ptr;
    ws_.n_trace = ints_.ptr + Bp;
    ws_.bad_input = d_flag_.ptr;
    B_ = 0;
  }

  M model_;
  int device_;
  int capacity_;
  int sm_count_ = 148;
  int Bp_ = 0;
  int B_ = 0;
  nmpc_b200_fmpc_config cfg_{};
  SolverParams<S> prm_{};
  Workspace<S> ws_{};
  cudaStream_t own_stream_ = nullptr;
  cudaStream_t last_stream_ = nullptr;
  DeviceBuffer<S> vars_, coeff_, term_, gains_, kkt_, trace_, scal_, mpc_x_, mpc_u_, mpc_k_;
  DeviceBuffer<int> ints_, d_flag_, mpc_i_;
  DeviceBuffer<double> stage_in_, stage_out_;
  int * h_flag_ = nullptr;
  bool timing_ = false;
  bool attr_set_[2] = {false, false};
  std::vector<cudaEvent_t> events_;
  int n_events_used_ = 0;
  int iter_event_base_ = 0;
  int iters_launched_ = 0;
  int launches_[4] = {0, 0, 0, 0};
};
} // namespace fmpc
} // namespace nmpc_b200
