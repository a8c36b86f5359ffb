/* nmpc_b200 -- extern "C" boundary (include/nmpc_b200/c_api.h).  Catches every C++ exception and
 * turns it into a status code + thread-local message. */
#include <nmpc_b200/c_api.h>

#include <cstring>
#include <exception>
#include <memory>
#include <string>

#include <dlfcn.h>

#include <nmpc_b200/engine/common.cuh>
#include <nmpc_b200/engine/registry.h>

struct nmpc_b200_ddp
{
  std::unique_ptr<nmpc_b200::DdpEngineBase> engine;
};

struct nmpc_b200_fmpc
{
  std::unique_ptr<nmpc_b200::FmpcEngineBase> engine;
};

namespace nmpc_b200
{
namespace
{
thread_local std::string g_last_error;
}

void setLastError(const std::string & msg)
{
  g_last_error = msg;
}

namespace
{
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
  catch(...)
  {
    setLastError("unknown error");
    return NMPC_B200_ERR_RUNTIME;
  }
}

const ModelEntry & findModel(const char * model)
{
  if(model == nullptr) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "null model name");
  const ModelEntry * e = registryFind(model);
  if(e == nullptr) throw Error(NMPC_B200_ERR_UNKNOWN_MODEL, std::string("unknown problem functor '") + model + "'");
  return *e;
}

void requireDevice(int device)
{
  int n = 0;
  cudaError_t err = cudaGetDeviceCount(&n);
  if(err != cudaSuccess || n <= 0)
  {
    cudaGetLastError();
    throw Error(NMPC_B200_ERR_NO_DEVICE,
                std::string("no usable CUDA device (") + (err != cudaSuccess ? cudaGetErrorString(err) : "count 0")
                    + "); nmpc_b200 has no CPU fallback");
  }
  if(device < 0 || device >= n)
    throw Error(NMPC_B200_ERR_NO_DEVICE,
                "device ordinal " + std::to_string(device) + " out of range [0, " + std::to_string(n) + ")");
}
} // namespace
} // namespace nmpc_b200

using namespace nmpc_b200;

extern "C"
{
const char * nmpc_b200_last_error(void)
{
  return g_last_error.c_str();
}

int nmpc_b200_version(void)
{
  return NMPC_B200_VERSION;
}

int nmpc_b200_device_count(void)
{
  int n = 0;
  if(cudaGetDeviceCount(&n) != cudaSuccess)
  {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int nmpc_b200_model_dims(const char * model, int * nx, int * nu, int * ng, int * n_params)
{
  return guarded([&] {
    const ModelEntry & e = findModel(model);
    if(nx) *nx = e.nx;
    if(nu) *nu = e.nu;
    if(ng) *ng = e.ng;
    if(n_params) *n_params = e.n_params;
  });
}

int nmpc_b200_model_default_params(const char * model, double * params)
{
  return guarded([&] {
    const ModelEntry & e = findModel(model);
    if(params == nullptr) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "null params");
    e.default_params(params);
  });
}

int nmpc_b200_model_count(void)
{
  return (int)registryNames().size();
}

const char * nmpc_b200_model_name(int index)
{
  const auto & names = registryNames();
  if(index < 0 || index >= (int)names.size()) return nullptr;
  return names[index].c_str();
}

int nmpc_b200_model_eval(const char * model,
                         const double * params,
                         int n_params,
                         int device,
                         int n,
                         const double * t,
                         const double * x,
                         const double * u,
                         double * x_next,
                         double * running_cost,
                         double * terminal_cost,
                         double * Fx,
                         double * Fu,
                         double * Lx,
                         double * Lu,
                         double * Lxx,
                         double * Luu,
                         double * Lxu,
                         double * Vx,
                         double * Vxx,
                         double * g,
                         double * C,
                         double * D)
{
  return guarded([&] {
    const ModelEntry & e = findModel(model);
    if(n_params != e.n_params)
      throw Error(NMPC_B200_ERR_INVALID_ARGUMENT,
                  "expected " + std::to_string(e.n_params) + " parameters, got " + std::to_string(n_params));
    if(!e.eval) throw Error(NMPC_B200_ERR_UNSUPPORTED, "functor has no evaluator");
    requireDevice(device);
    ModelEvalOutputs out;
    out.x_next = x_next;
    out.running_cost = running_cost;
    out.terminal_cost = terminal_cost;
    out.Fx = Fx;
    out.Fu = Fu;
    out.Lx = Lx;
    out.Lu = Lu;
    out.Lxx = Lxx;
    out.Luu = Luu;
    out.Lxu = Lxu;
    out.Vx = Vx;
    out.Vxx = Vxx;
    out.g = g;
    out.C = C;
    out.D = D;
    e.eval(params, device, n, t, x, u, out);
  });
}

/* ------------------------------------------------------------------------------- DDP ---- */

void nmpc_b200_ddp_config_default(nmpc_b200_ddp_config * cfg)
{
  // DDPSolver.h:47-110
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->horizon_steps = 100;
  cfg->max_iter = 500;
  cfg->reg_type = 1;
  cfg->with_input_constraint = 0;
  cfg->n_alpha = 11;
  cfg->use_state_eq_second_derivative = 0;
  cfg->initial_lambda = 1e-4;
  cfg->initial_dlambda = 1.0;
  cfg->lambda_factor = 1.6;
  cfg->lambda_min = 1e-6;
  cfg->lambda_max = 1e10;
  cfg->k_rel_norm_thre = 1e-4;
  cfg->lambda_thre = 1e-5;
  cfg->cost_update_ratio_thre = 0;
  cfg->cost_update_thre = 1e-7;
  // alpha_list = 10^LinSpaced(11, 0, -3): exponents 0, -0.3, ..., -3 (DDPSolver.h:53-59)
  for(int i = 0; i < 11; i++)
  {
    double e = (i == 10) ? -3.0 : 0.0 + i * ((-3.0 - 0.0) / 10);
    cfg->alpha_list[i] = std::pow(10.0, e);
  }
}

int nmpc_b200_ddp_create(const char * model,
                         const double * params,
                         int n_params,
                         const nmpc_b200_ddp_config * cfg,
                         int batch_capacity,
                         int device,
                         nmpc_b200_ddp ** out)
{
  return guarded([&] {
    if(out == nullptr) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "null output handle");
    *out = nullptr;
    const ModelEntry & e = findModel(model);
    if(!e.make_ddp) throw Error(NMPC_B200_ERR_UNSUPPORTED, std::string("functor '") + model + "' has no DDP kernels");
    if(params == nullptr || n_params != e.n_params)
      throw Error(NMPC_B200_ERR_INVALID_ARGUMENT,
                  "expected " + std::to_string(e.n_params) + " parameters, got " + std::to_string(n_params));
    nmpc_b200_ddp_config c;
    if(cfg)
      c = *cfg;
    else
      nmpc_b200_ddp_config_default(&c);
    requireDevice(device);
    auto h = std::make_unique<nmpc_b200_ddp>();
    h->engine = e.make_ddp(params, c, batch_capacity, device);
    *out = h.release();
  });
}

int nmpc_b200_ddp_destroy(nmpc_b200_ddp * h)
{
  return guarded([&] { delete h; });
}

#define NMPC_REQUIRE_HANDLE(h) \
  if((h) == nullptr || !(h)->engine) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "null handle")

int nmpc_b200_ddp_set_config(nmpc_b200_ddp * h, const nmpc_b200_ddp_config * cfg)
{
  return guarded([&] {
    NMPC_REQUIRE_HANDLE(h);
    if(cfg == nullptr) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "null config");
    h->engine->setConfig(*cfg);
  });
}

int nmpc_b200_ddp_get_config(const nmpc_b200_ddp * h, nmpc_b200_ddp_config * cfg)
{
  return guarded([&] {
    NMPC_REQUIRE_HANDLE(h);
    if(cfg == nullptr) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "null config");
    *cfg = h->engine->config();
  });
}

int nmpc_b200_ddp_set_input_limits(nmpc_b200_ddp * h, const double * lower, const double * upper)
{
  return guarded([&] {
    NMPC_REQUIRE_HANDLE(h);
    if(lower == nullptr || upper == nullptr) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "null limits");
    h->engine->setInputLimits(lower, upper);
  });
}

int nmpc_b200_ddp_set_input_limits_horizon(nmpc_b200_ddp * h, int n_steps, const double * lower, const double * upper)
{
  return guarded([&] {
    NMPC_REQUIRE_HANDLE(h);
    h->engine->setInputLimitsHorizon(n_steps, lower, upper);
  });
}

int nmpc_b200_ddp_set_input_limits_mpc(nmpc_b200_ddp * h, int n_ticks, int n_steps, const double * lower, const double * upper)
{
  return guarded([&] {
    NMPC_REQUIRE_HANDLE(h);
    h->engine->setInputLimitsMpc(n_ticks, n_steps, lower, upper);
  });
}

int nmpc_b200_ddp_solve(nmpc_b200_ddp * h,
                        int B,
                        double current_t,
                        const double * x0,
                        const double * u_init,
                        int n_u_steps,
                        int on_device,
                        void * stream)
{
  return guarded([&] {
    NMPC_REQUIRE_HANDLE(h);
    h->engine->solve(B, current_t, x0, u_init, n_u_steps, on_device != 0, stream);
  });
}

int nmpc_b200_ddp_run_mpc(nmpc_b200_ddp * h,
                          int B,
                          double current_t,
                          const double * x0,
                          const double * u_init,
                          int n_u_steps,
                          const nmpc_b200_mpc_config * mpc,
                          double * x_log,
                          double * u_log,
                          int * iters_log,
                          int * status_log,
                          int on_device,
                          void * stream)
{
  return guarded([&] {
    NMPC_REQUIRE_HANDLE(h);
    if(mpc == nullptr) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "null mpc configuration");
    h->engine->runMpc(B, current_t, x0, u_init, n_u_steps, *mpc, x_log, u_log, iters_log, status_log, on_device != 0,
                      stream);
  });
}

int nmpc_b200_ddp_get(nmpc_b200_ddp * h, int what, void * dst, size_t dst_bytes, int dst_on_device, void * stream)
{
  return guarded([&] {
    NMPC_REQUIRE_HANDLE(h);
    h->engine->get(what, dst, dst_bytes, dst_on_device != 0, stream);
  });
}

int nmpc_b200_ddp_set_tuning(nmpc_b200_ddp * h, const char * key, int value)
{
  return guarded([&] {
    NMPC_REQUIRE_HANDLE(h);
    h->engine->setTuning(key, value);
  });
}

int nmpc_b200_ddp_get_tuning(nmpc_b200_ddp * h, const char * key, int * value)
{
  return guarded([&] {
    NMPC_REQUIRE_HANDLE(h);
    if(value == nullptr) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "null value");
    *value = h->engine->getTuning(key);
  });
}

int nmpc_b200_ddp_sync(nmpc_b200_ddp * h)
{
  return guarded([&] {
    NMPC_REQUIRE_HANDLE(h);
    h->engine->sync();
  });
}

int nmpc_b200_ddp_enable_timing(nmpc_b200_ddp * h, int enable)
{
  return guarded([&] {
    NMPC_REQUIRE_HANDLE(h);
    h->engine->enableTiming(enable != 0);
  });
}

int nmpc_b200_ddp_get_durations(nmpc_b200_ddp * h, double * ms, int * launches)
{
  return guarded([&] {
    NMPC_REQUIRE_HANDLE(h);
    if(ms == nullptr) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "null output");
    h->engine->getDurations(ms, launches);
  });
}

int nmpc_b200_load_plugin(const char * path)
{
  return guarded([&] {
    if(path == nullptr) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "null plugin path");
    // RTLD_GLOBAL is not needed: the plugin's registrars call registryEntry() of THIS library when its statics run
    void * handle = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if(handle == nullptr)
    {
      const char * msg = dlerror();
      throw Error(NMPC_B200_ERR_RUNTIME, std::string("cannot load plugin '") + path + "': " + (msg ? msg : "unknown error"));
    }
  });
}

int nmpc_b200_ddp_get_iteration_durations(nmpc_b200_ddp * h, double * ms, int rows, int * rows_filled)
{
  return guarded([&] {
    NMPC_REQUIRE_HANDLE(h);
    if(ms == nullptr || rows <= 0) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "null or empty output");
    const int n = h->engine->getIterationDurations(ms, rows);
    if(rows_filled != nullptr) *rows_filled = n;
  });
}

/* ------------------------------------------------------------------------------ FMPC ---- */

void nmpc_b200_fmpc_config_default(nmpc_b200_fmpc_config * cfg)
{
  // FmpcSolver.h:58-89
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->horizon_steps = 100;
  cfg->max_iter = 10;
  cfg->check_nan = 1;
  cfg->init_complementary_variable = 0;
  cfg->update_barrier_eps = 1;
  cfg->break_if_llt_fails = 0;
  cfg->enable_line_search = 0;
  cfg->merit_const_scale_from_lagrange_multipliers = 0;
  cfg->kkt_error_thre = 1e-4;
  cfg->initial_barrier_eps = 1e-4;
}

int nmpc_b200_fmpc_create(const char * model,
                          const double * params,
                          int n_params,
                          const nmpc_b200_fmpc_config * cfg,
                          int batch_capacity,
                          int device,
                          nmpc_b200_fmpc ** out)
{
  return guarded([&] {
    if(out == nullptr) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "null output handle");
    *out = nullptr;
    const ModelEntry & e = findModel(model);
    if(!e.make_fmpc)
      throw Error(NMPC_B200_ERR_UNSUPPORTED, std::string("functor '") + model + "' has no FMPC kernels");
    if(params == nullptr || n_params != e.n_params)
      throw Error(NMPC_B200_ERR_INVALID_ARGUMENT,
                  "expected " + std::to_string(e.n_params) + " parameters, got " + std::to_string(n_params));
    nmpc_b200_fmpc_config c;
    if(cfg)
      c = *cfg;
    else
      nmpc_b200_fmpc_config_default(&c);
    requireDevice(device);
    auto h = std::make_unique<nmpc_b200_fmpc>();
    h->engine = e.make_fmpc(params, c, batch_capacity, device);
    *out = h.release();
  });
}

int nmpc_b200_fmpc_destroy(nmpc_b200_fmpc * h)
{
  return guarded([&] { delete h; });
}

int nmpc_b200_fmpc_set_config(nmpc_b200_fmpc * h, const nmpc_b200_fmpc_config * cfg)
{
  return guarded([&] {
    NMPC_REQUIRE_HANDLE(h);
    if(cfg == nullptr) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "null config");
    h->engine->setConfig(*cfg);
  });
}

int nmpc_b200_fmpc_solve(nmpc_b200_fmpc * h,
                         int B,
                         double current_t,
                         const double * x0,
                         const double * x,
                         const double * u,
                         const double * lambda,
                         const double * s,
                         const double * nu,
                         int n_steps,
                         int on_device,
                         void * stream)
{
  return guarded([&] {
    NMPC_REQUIRE_HANDLE(h);
    h->engine->solve(B, current_t, x0, x, u, lambda, s, nu, n_steps, on_device != 0, stream);
  });
}

int nmpc_b200_fmpc_run_mpc(nmpc_b200_fmpc * h,
                           int B,
                           double current_t,
                           const double * x0,
                           const double * x,
                           const double * u,
                           const double * lambda,
                           const double * s,
                           const double * nu,
                           int n_steps,
                           const nmpc_b200_mpc_config * mpc,
                           double * x_log,
                           double * u_log,
                           double * kkt_log,
                           int * status_log,
                           int on_device,
                           void * stream)
{
  return guarded([&] {
    NMPC_REQUIRE_HANDLE(h);
    if(mpc == nullptr) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "null mpc configuration");
    h->engine->runMpc(B, current_t, x0, x, u, lambda, s, nu, n_steps, *mpc, x_log, u_log, kkt_log, status_log,
                      on_device != 0, stream);
  });
}

int nmpc_b200_fmpc_get(nmpc_b200_fmpc * h, int what, void * dst, size_t dst_bytes, int dst_on_device, void * stream)
{
  return guarded([&] {
    NMPC_REQUIRE_HANDLE(h);
    h->engine->get(what, dst, dst_bytes, dst_on_device != 0, stream);
  });
}

int nmpc_b200_fmpc_sync(nmpc_b200_fmpc * h)
{
  return guarded([&] {
    NMPC_REQUIRE_HANDLE(h);
    h->engine->sync();
  });
}

int nmpc_b200_fmpc_enable_timing(nmpc_b200_fmpc * h, int enable)
{
  return guarded([&] {
    NMPC_REQUIRE_HANDLE(h);
    h->engine->enableTiming(enable != 0);
  });
}

int nmpc_b200_fmpc_get_durations(nmpc_b200_fmpc * h, double * ms, int * launches)
{
  return guarded([&] {
    NMPC_REQUIRE_HANDLE(h);
    if(ms == nullptr) throw Error(NMPC_B200_ERR_INVALID_ARGUMENT, "null output");
    h->engine->getDurations(ms, launches);
  });
}
} // extern "C"
