/* nmpc_b200 -- host side of the batched FMPC engine for one functor type M (see fmpc_kernels.cuh). */
#pragma once

#include <cstdlib>
#include <memory>
#include <vector>

#include "common.cuh"
#include "fmpc_kernels.cuh"
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
    const int N = cfg_.horizon_steps;
    // checkVariable(): sequence lengths (FmpcSolver.hpp:288-312)
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
    if(cfg_.enable_line_search)
    {
      throw Error(NMPC_B200_ERR_UNSUPPORTED,
                  "enable_line_search (merit-function line search, FmpcSolver.hpp:755-793) is not implemented on the "
                  "device yet");
    }

    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : own_stream_;
    last_stream_ = st;
    B_ = B;
    ws_.B = B;
    prm_.t0 = S(current_t);
    n_events_used_ = 0;
    for(int & l : launches_) l = 0;

    record(st); // 0
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

    const int tpb = threadsPerBlock(B);
    const int grid = (B + tpb - 1) / tpb;
    const int tpb1 = 128;
    const dim3 gridN((B + tpb1 - 1) / tpb1, N), gridN1((B + tpb1 - 1) / tpb1, N + 1);

    NMPC_CUDA_CHECK(cudaMemsetAsync(d_flag_.ptr, 0, sizeof(int), st));
    fmpc_init_kernel<M><<<gridN, tpb1, 0, st>>>(model_, ws_, prm_);
    // checkVariable(): s, nu must be non-negative (FmpcSolver.hpp:348-361) -- the reference throws, so
    // this is the one place where solve() waits for the device
    NMPC_CUDA_CHECK(cudaMemcpyAsync(h_flag_, d_flag_.ptr, sizeof(int), cudaMemcpyDeviceToHost, st));
    NMPC_CUDA_CHECK(cudaStreamSynchronize(st));
    if(*h_flag_ != 0)
    {
      B_ = 0;
      throw Error(NMPC_B200_ERR_RUNTIME, "[FMPC] s_list[i] / nu_list[i] must be non-negative.");
    }
    record(st); // 2: setup done

    iter_event_base_ = n_events_used_;
    iters_launched_ = 0;
    for(int iter = 1; iter <= cfg_.max_iter; iter++)
    {
      fmpc_coeff_kernel<M><<<gridN1, tpb1, 0, st>>>(model_, ws_, prm_);
      record(st);
      fmpc_backward_kernel<M><<<grid, tpb, 0, st>>>(model_, ws_, prm_, iter);
      record(st);
      fmpc_forward_kernel<M><<<grid, tpb, 0, st>>>(model_, ws_, prm_, iter);
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
  static int threadsPerBlock(int B)
  {
    if(const char * env = std::getenv("NMPC_B200_TPB"))
    {
      int v = std::atoi(env);
      if(v >= 32 && v <= 128 && v % 32 == 0) return v;
    }
    if(B <= 148 * 32 * 2) return 32;
    if(B <= 148 * 64 * 4) return 64;
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
  int Bp_ = 0;
  int B_ = 0;
  nmpc_b200_fmpc_config cfg_{};
  SolverParams<S> prm_{};
  Workspace<S> ws_{};
  cudaStream_t own_stream_ = nullptr;
  cudaStream_t last_stream_ = nullptr;
  DeviceBuffer<S> vars_, coeff_, term_, gains_, kkt_, trace_, scal_;
  DeviceBuffer<int> ints_, d_flag_;
  DeviceBuffer<double> stage_in_, stage_out_;
  int * h_flag_ = nullptr;
  bool timing_ = false;
  std::vector<cudaEvent_t> events_;
  int n_events_used_ = 0;
  int iter_event_base_ = 0;
  int iters_launched_ = 0;
  int launches_[4] = {0, 0, 0, 0};
};
} // namespace fmpc
} // namespace nmpc_b200
