/* nmpc_b200 -- registry of problem functors compiled into (or linked against) libnmpc_b200.so.
 *
 * The reference binds a problem to a solver through a shared_ptr to a class with virtual methods
 * (DDPSolver.h:255, :332).  Virtual host methods cannot run inside a kernel, so here a problem is a
 * trivially copyable functor type; registering it instantiates the stage kernels for that type and
 * makes it reachable by name through the C ABI (nmpc_b200_ddp_create(model, ...)).
 * User code adds a functor with NMPC_B200_REGISTER_DDP_MODEL / NMPC_B200_REGISTER_FMPC_MODEL in a
 * .cu file compiled with nvcc for sm_100a and linked with the library.
 */
#pragma once

#include <functional>
#include <memory>
#include <string>
#include <vector>

#include <nmpc_b200/c_api.h>

namespace nmpc_b200
{
/** Batched DDP solver bound to one functor type (type-erased view used by the C ABI). */
class DdpEngineBase
{
public:
  virtual ~DdpEngineBase() = default;
  virtual void setConfig(const nmpc_b200_ddp_config & cfg) = 0;
  virtual const nmpc_b200_ddp_config & config() const = 0;
  virtual void setInputLimits(const double * lower, const double * upper) = 0;
  virtual void setInputLimitsHorizon(int n_steps, const double * lower, const double * upper) = 0;
  virtual void setInputLimitsMpc(int n_ticks, int n_steps, const double * lower, const double * upper) = 0;
  virtual void solve(int B,
                     double current_t,
                     const double * x0,
                     const double * u_init,
                     int n_u_steps,
                     bool on_device,
                     void * stream) = 0;
  virtual void runMpc(int B,
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
                      void * stream) = 0;
  virtual void get(int what, void * dst, size_t dst_bytes, bool dst_on_device, void * stream) = 0;
  virtual void sync() = 0;
  virtual void enableTiming(bool enable) = 0;
  virtual void getDurations(double * ms, int * launches) = 0;
  virtual int getIterationDurations(double * ms, int rows) = 0;
  virtual void setTuning(const char * key, int value) = 0;
  virtual int getTuning(const char * key) = 0;
};

/** Batched FMPC solver bound to one functor type. */
class FmpcEngineBase
{
public:
  virtual ~FmpcEngineBase() = default;
  virtual void setConfig(const nmpc_b200_fmpc_config & cfg) = 0;
  virtual void solve(int B,
                     double current_t,
                     const double * x0,
                     const double * x,
                     const double * u,
                     const double * lambda,
                     const double * s,
                     const double * nu,
                     int n_steps,
                     bool on_device,
                     void * stream) = 0;
  virtual void runMpc(int B,
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
                      void * stream) = 0;
  virtual void get(int what, void * dst, size_t dst_bytes, bool dst_on_device, void * stream) = 0;
  virtual void sync() = 0;
  virtual void enableTiming(bool enable) = 0;
  virtual void getDurations(double * ms, int * launches) = 0;
};

/** Host outputs of a device-side functor evaluation; any pointer may be null. */
struct ModelEvalOutputs
{
  double * x_next = nullptr;
  double * running_cost = nullptr;
  double * terminal_cost = nullptr;
  double * Fx = nullptr;
  double * Fu = nullptr;
  double * Lx = nullptr;
  double * Lu = nullptr;
  double * Lxx = nullptr;
  double * Luu = nullptr;
  double * Lxu = nullptr;
  double * Vx = nullptr;
  double * Vxx = nullptr;
  double * g = nullptr;
  double * C = nullptr;
  double * D = nullptr;
};

struct ModelEntry
{
  std::string name;
  int nx = 0, nu = 0, ng = 0, n_params = 0;
  std::function<void(double *)> default_params;
  std::function<std::unique_ptr<DdpEngineBase>(const double * params,
                                               const nmpc_b200_ddp_config & cfg,
                                               int batch_capacity,
                                               int device)>
      make_ddp;
  std::function<std::unique_ptr<FmpcEngineBase>(const double * params,
                                                const nmpc_b200_fmpc_config & cfg,
                                                int batch_capacity,
                                                int device)>
      make_fmpc;
  std::function<void(const double * params,
                     int device,
                     int n,
                     const double * t,
                     const double * x,
                     const double * u,
                     const ModelEvalOutputs & out)>
      eval;
};

/** Name -> entry; entries are created on first use and later registrations merge into them. */
ModelEntry & registryEntry(const std::string & name);
const ModelEntry * registryFind(const std::string & name);
const std::vector<std::string> & registryNames();
} // namespace nmpc_b200
