/* nmpc_b200 -- several GPUs of one box (include/nmpc_b200/c_api.h, "several GPUs, one box").
 *
 * One process, many devices: nmpc_b200_ddp_sharded owns one single-GPU solver handle and one host worker thread per
 * device; a solve hands every worker its contiguous chunk of the host arrays, so the chunks are staged and solved
 * concurrently, and the host thread that called returns when the last worker has.  Results come back either to host
 * memory (each worker copies its rows into its slice) or into ONE device buffer, every shard's gather kernel storing its
 * rows directly into that device's memory over NVLink.
 *
 * One process per device: nmpc_b200_peer_* give the same direct stores between processes through a CUDA IPC mapping,
 * with a release/acquire flag word per rank instead of a collective.
 *
 * The reference has no counterpart: it runs one DDPSolver object per problem on one host thread
 * (DDPSolver.h:329-374); what is kept is that instances never interact. */
#include <nmpc_b200/c_api.h>

#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <nmpc_b200/engine/common.cuh>
#include <nmpc_b200/engine/registry.h>

namespace nmpc_b200
{
void setLastError(const std::string & msg);

namespace
{
/** A host thread that runs one job at a time for one shard; the job's failure is kept for wait(). */
class Worker
{
public:
  Worker() : thread_([this] { loop(); }) {}
  ~Worker()
  {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
    }
    cv_.notify_all();
    thread_.join();
  }
  void post(std::function<void()> job)
  {
    {
      std::lock_guard<std::mutex> lk(m_);
      job_ = std::move(job);
      busy_ = true;
      code_ = NMPC_B200_OK;
    }
    cv_.notify_all();
  }
  /** Blocks until the posted job is done; returns its status and message. */
  int wait(std::string & msg)
  {
    std::unique_lock<std::mutex> lk(m_);
    cv_.wait(lk, [this] { return !busy_; });
    msg = msg_;
    return code_;
  }

private:
  void loop()
  {
    for(;;)
    {
      std::function<void()> job;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [this] { return stop_ || (busy_ && job_); });
        if(stop_) return;
        job = std::move(job_);
        job_ = nullptr;
      }
      int code = NMPC_B200_OK;
      std::string msg;
      try
      {
        job();
      }
      catch(const Error & e)
      {
        code = e.code;
        msg = e.what();
      }
      catch(const std::exception & e)
      {
        code = NMPC_B200_ERR_RUNTIME;
        msg = e.what();
      }
      {
        std::lock_guard<std::mutex> lk(m_);
        code_ = code;
        msg_ = msg;
        busy_ = false;
      }
      cv_.notify_all();
    }
  }
  std::mutex m_;
  std::condition_variable cv_;
  std::function<void()> job_;
  bool busy_ = false, stop_ = false;
  int code_ = NMPC_B200_OK;
  std::string msg_;
  std::thread thread_; // last: starts when the members above exist
};

struct Shard
{
  int device = 0;
  nmpc_b200_ddp * handle = nullptr;
  std::unique_ptr<Worker> worker;
};

struct FmpcShard
{
  int device = 0;
  nmpc_b200_fmpc * handle = nullptr;
  std::unique_ptr<Worker> worker;
};

/** Runs job(shard index) on every shard's worker and waits for all; the first failure is rethrown. */
template<class ShardVector>
void forAllShards(ShardVector & shards, const std::function<void(int)> & job)
{
  for(size_t i = 0; i < shards.size(); i++) shards[i].worker->post([&job, i] { job((int)i); });
  int code = NMPC_B200_OK;
  std::string msg;
  for(size_t i = 0; i < shards.size(); i++)
  {
    std::string m;
    const int c = shards[i].worker->wait(m);
    if(c != NMPC_B200_OK && code == NMPC_B200_OK)
    {
      code = c;
      msg = "shard " + std::to_string(i) + " (device " + std::to_string(shards[i].device) + "): " + m;
    }
  }
  if(code != NMPC_B200_OK) throw Error(code, msg);
}

/** Peer access from `from` to `to` for direct stores (idempotent). */
void enablePeerStores(int from, int to)
{
  if(from == to) return;
  DeviceGuard guard(from);
  int can = 0;
  NMPC_CUDA_CHECK(cudaDeviceCanAccessPeer(&can, from, to));
  if(!can)
    throw Error(NMPC_B200_ERR_UNSUPPORTED,
                "device " + std::to_string(from) + " cannot store into device " + std::to_string(to) + " (no peer access)");
  const cudaError_t err = cudaDeviceEnablePeerAccess(to, 0);
  if(err != cudaSuccess && err != cudaErrorPeerAccessAlreadyEnabled) NMPC_CUDA_CHECK(err);
  cudaGetLastError();
}

/** Bytes of one instance's row of field `what` (nmpc_b200_fmpc_field). */
size_t fmpcFieldRowBytes(int what, int nx, int nu, int ng, const nmpc_b200_fmpc_config & c)
{
  const size_t N = c.horizon_steps;
  switch(what)
  {
    case NMPC_B200_FMPC_X:
    case NMPC_B200_FMPC_LAMBDA:
      return sizeof(double) * (N + 1) * nx;
    case NMPC_B200_FMPC_U:
    case NMPC_B200_FMPC_K_FF:
      return sizeof(double) * N * nu;
    case NMPC_B200_FMPC_S:
    case NMPC_B200_FMPC_NU:
      return sizeof(double) * N * ng;
    case NMPC_B200_FMPC_K_FB:
      return sizeof(double) * N * nu * nx;
    case NMPC_B200_FMPC_TRACE:
      return sizeof(double) * (size_t)c.max_iter * 5;
    case NMPC_B200_FMPC_U0:
      return sizeof(double) * nu;
    case NMPC_B200_FMPC_STATUS:
    case NMPC_B200_FMPC_N_TRACE:
      return sizeof(int);
    default:
      throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "unknown field " + std::to_string(what));
  }
}

void shardRange(int B, int n, int s, int & begin, int & end)
{
  const int base = B / n, rem = B % n;
  begin = s * base + (s < rem ? s : rem);
  end = begin + base + (s < rem ? 1 : 0);
}

void check(int status)
{
  if(status != NMPC_B200_OK) throw Error(status, nmpc_b200_last_error());
}

/** Bytes of one instance's row of field `what` (the [B][...] layouts of nmpc_b200_ddp_field). */
size_t fieldRowBytes(int what, int nx, int nu, const nmpc_b200_ddp_config & c)
{
  const size_t N = c.horizon_steps;
  switch(what)
  {
    case NMPC_B200_DDP_X:
      return sizeof(double) * (N + 1) * nx;
    case NMPC_B200_DDP_U:
    case NMPC_B200_DDP_K_FF:
      return sizeof(double) * N * nu;
    case NMPC_B200_DDP_COST_LIST:
      return sizeof(double) * (N + 1);
    case NMPC_B200_DDP_K_FB:
      return sizeof(double) * N * nu * nx;
    case NMPC_B200_DDP_TRACE:
      return sizeof(double) * (size_t)(c.max_iter + 1) * 9;
    case NMPC_B200_DDP_COST:
      return sizeof(double);
    case NMPC_B200_DDP_U0:
      return sizeof(double) * nu;
    case NMPC_B200_DDP_STATUS:
    case NMPC_B200_DDP_ITERS:
    case NMPC_B200_DDP_N_FORWARD:
    case NMPC_B200_DDP_N_BACKWARD:
    case NMPC_B200_DDP_N_TRACE:
      return sizeof(int);
    default:
      throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "unknown field " + std::to_string(what));
  }
}

template<class F>
int guarded(F && f)
{
  try
  {
    f();
    return NMPC_B200_OK;
  }
  catch(const Error & e)
  {
    setLastError(e.what());
    return e.code;
  }
  catch(const std::exception & e)
  {
    setLastError(e.what());
    return NMPC_B200_ERR_RUNTIME;
  }
}

/* ------------------------------------------------------------------ flags between processes ---- */
__global__ void peer_signal_kernel(unsigned long long * flag, unsigned long long value)
{
  // everything this stream stored before (the gather kernel's rows) becomes visible before the flag does
  __threadfence_system();
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(value) : "memory");
}

__global__ void peer_wait_kernel(unsigned long long * flags, int n_flags, unsigned long long value, long long timeout_ns)
{
  const int i = threadIdx.x;
  if(i >= n_flags) return;
  long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for(;;)
  {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + i) : "memory");
    if(v >= value) break;
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if(t - t0 > timeout_ns)
    {
      atomicAdd(flags + n_flags, 1ull); // the time-out word
      break;
    }
    __nanosleep(200);
  }
}
} // namespace
} // namespace nmpc_b200

using namespace nmpc_b200;

struct nmpc_b200_ddp_sharded
{
  std::vector<Shard> shards;
  int nx = 0, nu = 0, capacity = 0, last_B = 0;
  nmpc_b200_ddp_config cfg;
  std::vector<char> peer_enabled; // [shard]: peer access to the last dst_device enabled
  int peer_device = -1;

  ~nmpc_b200_ddp_sharded()
  {
    for(auto & s : shards)
    {
      s.worker.reset(); // joins
      if(s.handle) nmpc_b200_ddp_destroy(s.handle);
    }
  }

  /** Runs job(shard index) on every shard's worker and waits for all; the first failure is rethrown. */
  void forAll(const std::function<void(int)> & job)
  {
    for(size_t i = 0; i < shards.size(); i++) shards[i].worker->post([&job, i] { job((int)i); });
    int code = NMPC_B200_OK;
    std::string msg;
    for(size_t i = 0; i < shards.size(); i++)
    {
      std::string m;
      const int c = shards[i].worker->wait(m);
      if(c != NMPC_B200_OK && code == NMPC_B200_OK)
      {
        code = c;
        msg = "shard " + std::to_string(i) + " (device " + std::to_string(shards[i].device) + "): " + m;
      }
    }
    if(code != NMPC_B200_OK) throw Error(code, msg);
  }
};

#define NMPC_REQUIRE_SHARDED(h) \
  if((h) == nullptr || (h)->shards.empty()) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "null handle")

extern "C"
{
int nmpc_b200_ddp_create_sharded(const char * model,
                                 const double * params,
                                 int n_params,
                                 const nmpc_b200_ddp_config * cfg,
                                 int total_capacity,
                                 const int * devices,
                                 int n_devices,
                                 nmpc_b200_ddp_sharded ** out)
{
  return guarded([&] {
    if(out == nullptr) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "null output handle");
    *out = nullptr;
    if(total_capacity <= 0) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "total_capacity must be positive");
    const int visible = nmpc_b200_device_count();
    if(visible <= 0) throw Error(NMPC_B200_ERR_NO_DEVICE, "no usable CUDA device; nmpc_b200 has no CPU fallback");
    if(n_devices <= 0)
    {
      n_devices = visible;
      devices = nullptr;
    }
    auto h = std::make_unique<nmpc_b200_ddp_sharded>();
    int ng = 0, np = 0;
    check(nmpc_b200_model_dims(model, &h->nx, &h->nu, &ng, &np));
    if(cfg)
      h->cfg = *cfg;
    else
      nmpc_b200_ddp_config_default(&h->cfg);
    h->capacity = total_capacity;
    const int per_shard = (total_capacity + n_devices - 1) / n_devices;
    h->shards.resize(n_devices);
    for(int s = 0; s < n_devices; s++)
    {
      Shard & sh = h->shards[s];
      sh.device = devices ? devices[s] : s;
      check(nmpc_b200_ddp_create(model, params, n_params, &h->cfg, per_shard, sh.device, &sh.handle));
      sh.worker = std::make_unique<Worker>();
    }
    h->peer_enabled.assign(n_devices, 0);
    *out = h.release();
  });
}

int nmpc_b200_ddp_sharded_destroy(nmpc_b200_ddp_sharded * h)
{
  return guarded([&] { delete h; });
}

int nmpc_b200_ddp_sharded_num_shards(const nmpc_b200_ddp_sharded * h)
{
  return h ? (int)h->shards.size() : 0;
}

nmpc_b200_ddp * nmpc_b200_ddp_sharded_shard(nmpc_b200_ddp_sharded * h, int shard)
{
  if(h == nullptr || shard < 0 || shard >= (int)h->shards.size()) return nullptr;
  return h->shards[shard].handle;
}

int nmpc_b200_ddp_sharded_range(const nmpc_b200_ddp_sharded * h, int B, int shard, int * begin, int * end, int * device)
{
  return guarded([&] {
    NMPC_REQUIRE_SHARDED(h);
    if(shard < 0 || shard >= (int)h->shards.size() || B < 0)
      throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "shard index or batch size out of range");
    int b = 0, e = 0;
    shardRange(B, (int)h->shards.size(), shard, b, e);
    if(begin) *begin = b;
    if(end) *end = e;
    if(device) *device = h->shards[shard].device;
  });
}

int nmpc_b200_ddp_sharded_set_config(nmpc_b200_ddp_sharded * h, const nmpc_b200_ddp_config * cfg)
{
  return guarded([&] {
    NMPC_REQUIRE_SHARDED(h);
    if(cfg == nullptr) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "null config");
    for(auto & s : h->shards) check(nmpc_b200_ddp_set_config(s.handle, cfg));
    h->cfg = *cfg;
  });
}

int nmpc_b200_ddp_sharded_set_input_limits(nmpc_b200_ddp_sharded * h, const double * lower, const double * upper)
{
  return guarded([&] {
    NMPC_REQUIRE_SHARDED(h);
    for(auto & s : h->shards) check(nmpc_b200_ddp_set_input_limits(s.handle, lower, upper));
  });
}

int nmpc_b200_ddp_sharded_solve(nmpc_b200_ddp_sharded * h,
                                int B,
                                double current_t,
                                const double * x0,
                                const double * u_init,
                                int n_u_steps)
{
  return guarded([&] {
    NMPC_REQUIRE_SHARDED(h);
    if(B <= 0 || B > h->capacity)
      throw Error(NMPC_B200_ERR_INVALID_ARGUMENT,
                  "batch size " + std::to_string(B) + " outside (0, " + std::to_string(h->capacity) + "]");
    if(x0 == nullptr || u_init == nullptr) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "null input array");
    const int n = (int)h->shards.size();
    const size_t x_row = h->nx, u_row = (size_t)(n_u_steps > 0 ? n_u_steps : 0) * h->nu;
    h->last_B = 0;
    h->forAll([&](int s) {
      int b = 0, e = 0;
      shardRange(B, n, s, b, e);
      if(e == b) return; // fewer instances than shards
      Shard & sh = h->shards[s];
      check(nmpc_b200_ddp_solve(sh.handle, e - b, current_t, x0 + b * x_row, u_init + b * u_row, n_u_steps, 0, nullptr));
      check(nmpc_b200_ddp_sync(sh.handle));
    });
    h->last_B = B;
  });
}

int nmpc_b200_ddp_sharded_get(nmpc_b200_ddp_sharded * h, int what, void * dst, size_t dst_bytes, int dst_device)
{
  return guarded([&] {
    NMPC_REQUIRE_SHARDED(h);
    if(h->last_B <= 0) throw Error(NMPC_B200_ERR_RUNTIME, "get() before solve()");
    if(dst == nullptr) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "null destination");
    const size_t row = fieldRowBytes(what, h->nx, h->nu, h->cfg);
    const int B = h->last_B, n = (int)h->shards.size();
    if(dst_bytes < row * B)
      throw Error(NMPC_B200_ERR_INVALID_ARGUMENT,
                  "destination holds " + std::to_string(dst_bytes) + " bytes, the field needs " + std::to_string(row * B));
    if(dst_device >= 0 && dst_device != h->peer_device)
    {
      h->peer_enabled.assign(n, 0);
      h->peer_device = dst_device;
    }
    h->forAll([&](int s) {
      int b = 0, e = 0;
      shardRange(B, n, s, b, e);
      if(e == b) return;
      Shard & sh = h->shards[s];
      if(dst_device >= 0 && dst_device != sh.device && !h->peer_enabled[s])
      {
        DeviceGuard guard(sh.device);
        int can = 0;
        NMPC_CUDA_CHECK(cudaDeviceCanAccessPeer(&can, sh.device, dst_device));
        if(!can)
          throw Error(NMPC_B200_ERR_UNSUPPORTED, "device " + std::to_string(sh.device) + " cannot store into device "
                                                     + std::to_string(dst_device) + " (no peer access)");
        const cudaError_t err = cudaDeviceEnablePeerAccess(dst_device, 0);
        if(err != cudaSuccess && err != cudaErrorPeerAccessAlreadyEnabled) NMPC_CUDA_CHECK(err);
        cudaGetLastError();
        h->peer_enabled[s] = 1;
      }
      check(nmpc_b200_ddp_get(sh.handle, what, static_cast<char *>(dst) + row * b, row * (e - b), dst_device >= 0 ? 1 : 0,
                              nullptr));
      check(nmpc_b200_ddp_sync(sh.handle));
    });
  });
}

/* --------------------------------------------------------------------------------- FMPC ---- */
} // extern "C"

struct nmpc_b200_fmpc_sharded
{
  std::vector<FmpcShard> shards;
  int nx = 0, nu = 0, ng = 0, capacity = 0, last_B = 0;
  nmpc_b200_fmpc_config cfg;

  ~nmpc_b200_fmpc_sharded()
  {
    for(auto & s : shards)
    {
      s.worker.reset();
      if(s.handle) nmpc_b200_fmpc_destroy(s.handle);
    }
  }
};

extern "C"
{
int nmpc_b200_fmpc_create_sharded(const char * model,
                                  const double * params,
                                  int n_params,
                                  const nmpc_b200_fmpc_config * cfg,
                                  int total_capacity,
                                  const int * devices,
                                  int n_devices,
                                  nmpc_b200_fmpc_sharded ** out)
{
  return guarded([&] {
    if(out == nullptr) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "null output handle");
    *out = nullptr;
    if(total_capacity <= 0) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "total_capacity must be positive");
    const int visible = nmpc_b200_device_count();
    if(visible <= 0) throw Error(NMPC_B200_ERR_NO_DEVICE, "no usable CUDA device; nmpc_b200 has no CPU fallback");
    if(n_devices <= 0)
    {
      n_devices = visible;
      devices = nullptr;
    }
    auto h = std::make_unique<nmpc_b200_fmpc_sharded>();
    int np = 0;
    check(nmpc_b200_model_dims(model, &h->nx, &h->nu, &h->ng, &np));
    if(cfg)
      h->cfg = *cfg;
    else
      nmpc_b200_fmpc_config_default(&h->cfg);
    h->capacity = total_capacity;
    const int per_shard = (total_capacity + n_devices - 1) / n_devices;
    h->shards.resize(n_devices);
    for(int s = 0; s < n_devices; s++)
    {
      FmpcShard & sh = h->shards[s];
      sh.device = devices ? devices[s] : s;
      check(nmpc_b200_fmpc_create(model, params, n_params, &h->cfg, per_shard, sh.device, &sh.handle));
      sh.worker = std::make_unique<Worker>();
    }
    *out = h.release();
  });
}

int nmpc_b200_fmpc_sharded_destroy(nmpc_b200_fmpc_sharded * h)
{
  return guarded([&] { delete h; });
}

int nmpc_b200_fmpc_sharded_num_shards(const nmpc_b200_fmpc_sharded * h)
{
  return h ? (int)h->shards.size() : 0;
}

int nmpc_b200_fmpc_sharded_set_config(nmpc_b200_fmpc_sharded * h, const nmpc_b200_fmpc_config * cfg)
{
  return guarded([&] {
    NMPC_REQUIRE_SHARDED(h);
    if(cfg == nullptr) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "null config");
    for(auto & s : h->shards) check(nmpc_b200_fmpc_set_config(s.handle, cfg));
    h->cfg = *cfg;
  });
}

int nmpc_b200_fmpc_sharded_solve(nmpc_b200_fmpc_sharded * h,
                                 int B,
                                 double current_t,
                                 const double * x0,
                                 const double * x,
                                 const double * u,
                                 const double * lambda,
                                 const double * s,
                                 const double * nu,
                                 int n_steps)
{
  return guarded([&] {
    NMPC_REQUIRE_SHARDED(h);
    if(B <= 0 || B > h->capacity)
      throw Error(NMPC_B200_ERR_INVALID_ARGUMENT,
                  "batch size " + std::to_string(B) + " outside (0, " + std::to_string(h->capacity) + "]");
    if(!x0 || !x || !u || !lambda || !s || !nu) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "null input array");
    const int n = (int)h->shards.size();
    const size_t N = n_steps > 0 ? n_steps : 0;
    const size_t r_x0 = h->nx, r_x = (N + 1) * h->nx, r_u = N * h->nu, r_g = N * h->ng;
    h->last_B = 0;
    forAllShards(h->shards, [&](int i) {
      int b = 0, e = 0;
      shardRange(B, n, i, b, e);
      if(e == b) return;
      FmpcShard & sh = h->shards[i];
      check(nmpc_b200_fmpc_solve(sh.handle, e - b, current_t, x0 + b * r_x0, x + b * r_x, u + b * r_u, lambda + b * r_x,
                                 s + b * r_g, nu + b * r_g, n_steps, 0, nullptr));
      check(nmpc_b200_fmpc_sync(sh.handle));
    });
    h->last_B = B;
  });
}

int nmpc_b200_fmpc_sharded_get(nmpc_b200_fmpc_sharded * h, int what, void * dst, size_t dst_bytes, int dst_device)
{
  return guarded([&] {
    NMPC_REQUIRE_SHARDED(h);
    if(h->last_B <= 0) throw Error(NMPC_B200_ERR_RUNTIME, "get() before solve()");
    if(dst == nullptr) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "null destination");
    const size_t row = fmpcFieldRowBytes(what, h->nx, h->nu, h->ng, h->cfg);
    const int B = h->last_B, n = (int)h->shards.size();
    if(dst_bytes < row * B)
      throw Error(NMPC_B200_ERR_INVALID_ARGUMENT,
                  "destination holds " + std::to_string(dst_bytes) + " bytes, the field needs " + std::to_string(row * B));
    forAllShards(h->shards, [&](int i) {
      int b = 0, e = 0;
      shardRange(B, n, i, b, e);
      if(e == b) return;
      FmpcShard & sh = h->shards[i];
      if(dst_device >= 0) enablePeerStores(sh.device, dst_device);
      check(nmpc_b200_fmpc_get(sh.handle, what, static_cast<char *>(dst) + row * b, row * (e - b), dst_device >= 0 ? 1 : 0,
                               nullptr));
      check(nmpc_b200_fmpc_sync(sh.handle));
    });
  });
}

/* ------------------------------------------------------------------------ between processes ---- */
int nmpc_b200_peer_buffer_create(size_t bytes, int device, void ** ptr, unsigned char handle[64])
{
  return guarded([&] {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the C ABI carries the IPC handle as 64 bytes");
    if(ptr == nullptr || handle == nullptr || bytes == 0) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "bad argument");
    DeviceGuard guard(device);
    void * p = nullptr;
    NMPC_CUDA_CHECK(cudaMalloc(&p, bytes));
    NMPC_CUDA_CHECK(cudaMemset(p, 0, bytes));
    NMPC_CUDA_CHECK(cudaDeviceSynchronize());
    cudaIpcMemHandle_t hd;
    NMPC_CUDA_CHECK(cudaIpcGetMemHandle(&hd, p));
    memcpy(handle, &hd, 64);
    *ptr = p;
  });
}

int nmpc_b200_peer_buffer_open(const unsigned char handle[64], int device, void ** ptr)
{
  return guarded([&] {
    if(ptr == nullptr || handle == nullptr) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "bad argument");
    DeviceGuard guard(device);
    cudaIpcMemHandle_t hd;
    memcpy(&hd, handle, 64);
    NMPC_CUDA_CHECK(cudaIpcOpenMemHandle(ptr, hd, cudaIpcMemLazyEnablePeerAccess));
  });
}

int nmpc_b200_peer_buffer_close(void * ptr, int device)
{
  return guarded([&] {
    DeviceGuard guard(device);
    NMPC_CUDA_CHECK(cudaIpcCloseMemHandle(ptr));
  });
}

int nmpc_b200_peer_buffer_destroy(void * ptr, int device)
{
  return guarded([&] {
    DeviceGuard guard(device);
    NMPC_CUDA_CHECK(cudaFree(ptr));
  });
}

int nmpc_b200_peer_signal(void * flag, unsigned long long value, int device, void * stream)
{
  return guarded([&] {
    if(flag == nullptr) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "null flag");
    DeviceGuard guard(device);
    peer_signal_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<unsigned long long *>(flag), value);
    NMPC_CUDA_CHECK(cudaGetLastError());
  });
}

int nmpc_b200_peer_wait(void * flags, int n_flags, unsigned long long value, int timeout_ms, int device, void * stream)
{
  return guarded([&] {
    if(flags == nullptr || n_flags <= 0 || n_flags > 1024) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "bad flags");
    DeviceGuard guard(device);
    peer_wait_kernel<<<1, ((n_flags + 31) / 32) * 32, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<unsigned long long *>(flags), n_flags, value, (long long)(timeout_ms > 0 ? timeout_ms : 2000) * 1000000ll);
    NMPC_CUDA_CHECK(cudaGetLastError());
  });
}

int nmpc_b200_peer_check(void * flags, int n_flags, int device, void * stream)
{
  return guarded([&] {
    if(flags == nullptr || n_flags <= 0) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "bad flags");
    DeviceGuard guard(device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned long long timed_out = 0;
    NMPC_CUDA_CHECK(cudaMemcpyAsync(&timed_out, static_cast<unsigned long long *>(flags) + n_flags, sizeof(timed_out),
                                    cudaMemcpyDeviceToHost, st));
    NMPC_CUDA_CHECK(cudaStreamSynchronize(st));
    if(timed_out != 0)
    {
      NMPC_CUDA_CHECK(cudaMemsetAsync(static_cast<unsigned long long *>(flags) + n_flags, 0, sizeof(timed_out), st));
      NMPC_CUDA_CHECK(cudaStreamSynchronize(st));
      throw Error(NMPC_B200_ERR_RUNTIME, "peer_wait timed out " + std::to_string(timed_out) + " time(s): a rank did not signal");
    }
  });
}
}
