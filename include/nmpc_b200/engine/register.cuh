/* nmpc_b200 -- registration of problem functors (instantiates the kernels for the functor type). */
#pragma once

#include "ddp_engine.cuh"
#include "fmpc_engine.cuh"
#include "model_eval.cuh"
#include "registry.h"

namespace nmpc_b200
{
template<class M>
struct DdpRegistrar
{
  explicit DdpRegistrar(const char * name)
  {
    ModelEntry & e = registryEntry(name);
    e.nx = M::NX;
    e.nu = M::NU;
    e.ng = ineqDimOf<M>();
    e.n_params = M::NUM_PARAMS;
    e.default_params = [](double * p) { M::defaultParams(p); };
    e.make_ddp = [](const double * params, const nmpc_b200_ddp_config & cfg, int cap, int dev) {
      return std::unique_ptr<DdpEngineBase>(new ddp::DdpEngine<M>(params, cfg, cap, dev));
    };
    e.eval = [](const double * params, int dev, int n, const double * t, const double * x, const double * u,
                const ModelEvalOutputs & out) { modelEval<M>(params, dev, n, t, x, u, out); };
  }
};

/** The functor's pointwise evaluation only (nmpc_b200_model_eval): no solver kernels are instantiated. */
template<class M>
struct EvalRegistrar
{
  explicit EvalRegistrar(const char * name)
  {
    ModelEntry & e = registryEntry(name);
    e.nx = M::NX;
    e.nu = M::NU;
    e.ng = ineqDimOf<M>();
    e.n_params = M::NUM_PARAMS;
    e.default_params = [](double * p) { M::defaultParams(p); };
    e.eval = [](const double * params, int dev, int n, const double * t, const double * x, const double * u,
                const ModelEvalOutputs & out) { modelEval<M>(params, dev, n, t, x, u, out); };
  }
};

template<class M>
struct FmpcRegistrar
{
  explicit FmpcRegistrar(const char * name)
  {
    ModelEntry & e = registryEntry(name);
    e.nx = M::NX;
    e.nu = M::NU;
    e.ng = M::NG;
    e.n_params = M::NUM_PARAMS;
    e.default_params = [](double * p) { M::defaultParams(p); };
    e.make_fmpc = [](const double * params, const nmpc_b200_fmpc_config & cfg, int cap, int dev) {
      return std::unique_ptr<FmpcEngineBase>(new fmpc::FmpcEngine<M>(params, cfg, cap, dev));
    };
    e.eval = [](const double * params, int dev, int n, const double * t, const double * x, const double * u,
                const ModelEvalOutputs & out) { modelEval<M>(params, dev, n, t, x, u, out); };
  }
};
} // namespace nmpc_b200

#define NMPC_B200_CONCAT_(a, b) a##b
#define NMPC_B200_CONCAT(a, b) NMPC_B200_CONCAT_(a, b)
/** Make functor type MODEL available to nmpc_b200_ddp_create() under NAME. */
#define NMPC_B200_REGISTER_DDP_MODEL(NAME, ...) \
  static ::nmpc_b200::DdpRegistrar<__VA_ARGS__> NMPC_B200_CONCAT(nmpc_b200_ddp_registrar_, __COUNTER__)(NAME)
/** Make functor type MODEL (with ineqConst / calcIneqConstDeriv) available to nmpc_b200_fmpc_create() under NAME. */
#define NMPC_B200_REGISTER_FMPC_MODEL(NAME, ...) \
  static ::nmpc_b200::FmpcRegistrar<__VA_ARGS__> NMPC_B200_CONCAT(nmpc_b200_fmpc_registrar_, __COUNTER__)(NAME)
/** Make functor type MODEL available to nmpc_b200_model_eval() only (no solver can be created for NAME). */
#define NMPC_B200_REGISTER_EVAL_ONLY(NAME, ...) \
  static ::nmpc_b200::EvalRegistrar<__VA_ARGS__> NMPC_B200_CONCAT(nmpc_b200_eval_registrar_, __COUNTER__)(NAME)
