/* nmpc_b200 -- evaluate a problem functor on the device at n sample points (derivative checks,
 * reference: nmpc_ddp/tests/src/TestDDPCartPole.cpp:609-649). */
#pragma once

#include <type_traits>
#include <vector>

#include "common.cuh"
#include "registry.h"

namespace nmpc_b200
{
template<class M, class = void>
struct HasIneq : std::false_type
{
};
template<class M>
struct HasIneq<M, std::void_t<decltype(&M::ineqConst)>> : std::true_type
{
};

template<class M>
constexpr int ineqDimOf()
{
  if constexpr(HasIneq<M>::value)
    return M::NG;
  else
    return 0;
}

/** out layout per point: [x_next NX | rc 1 | tc 1 | Fx | Fu | Lx | Lu | Lxx | Luu | Lxu | Vx | Vxx | g | C | D] */
template<class M>
struct EvalLayout
{
  static constexpr int NX = M::NX, NU = M::NU, NG = ineqDimOf<M>();
  static constexpr int XN = 0;
  static constexpr int RC = XN + NX;
  static constexpr int TC = RC + 1;
  static constexpr int FX = TC + 1;
  static constexpr int FU = FX + NX * NX;
  static constexpr int LX = FU + NX * NU;
  static constexpr int LU = LX + NX;
  static constexpr int LXX = LU + NU;
  static constexpr int LUU = LXX + NX * NX;
  static constexpr int LXU = LUU + NU * NU;
  static constexpr int VX = LXU + NX * NU;
  static constexpr int VXX = VX + NX;
  static constexpr int G = VXX + NX * NX;
  static constexpr int CC = G + NG;
  static constexpr int DD = CC + NG * NX;
  static constexpr int SIZE = DD + NG * NU;
};

template<class M>
__global__ void model_eval_kernel(const __grid_constant__ M model,
                                  int n,
                                  const double * __restrict__ t,
                                  const double * __restrict__ x,
                                  const double * __restrict__ u,
                                  double * __restrict__ out)
{
  using S = typename M::Scalar;
  using E = EvalLayout<M>;
  constexpr int NX = M::NX, NU = M::NU;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if(p >= n) return;
  Matrix<S, NX, 1> xv;
  Matrix<S, NU, 1> uv;
  for(int d = 0; d < NX; d++) xv[d] = S(x[(size_t)p * NX + d]);
  for(int d = 0; d < NU; d++) uv[d] = S(u[(size_t)p * NU + d]);
  const S tv = S(t[p]);
  double * o = out + (size_t)p * E::SIZE;

  Matrix<S, NX, 1> xn = model.stateEq(tv, xv, uv);
  for(int d = 0; d < NX; d++) o[E::XN + d] = double(xn[d]);
  o[E::RC] = double(model.runningCost(tv, xv, uv));
  o[E::TC] = double(model.terminalCost(tv, xv));
  Matrix<S, NX, NX> Fx, Lxx, Vxx;
  Matrix<S, NX, NU> Fu, Lxu;
  Matrix<S, NX, 1> Lx, Vx;
  Matrix<S, NU, 1> Lu;
  Matrix<S, NU, NU> Luu;
  model.calcStateEqDeriv(tv, xv, uv, Fx, Fu);
  model.calcRunningCostDeriv(tv, xv, uv, Lx, Lu, Lxx, Luu, Lxu);
  model.calcTerminalCostDeriv(tv, xv, Vx, Vxx);
  for(int d = 0; d < NX * NX; d++) o[E::FX + d] = double(Fx.d[d]);
  for(int d = 0; d < NX * NU; d++) o[E::FU + d] = double(Fu.d[d]);
  for(int d = 0; d < NX; d++) o[E::LX + d] = double(Lx.d[d]);
  for(int d = 0; d < NU; d++) o[E::LU + d] = double(Lu.d[d]);
  for(int d = 0; d < NX * NX; d++) o[E::LXX + d] = double(Lxx.d[d]);
  for(int d = 0; d < NU * NU; d++) o[E::LUU + d] = double(Luu.d[d]);
  for(int d = 0; d < NX * NU; d++) o[E::LXU + d] = double(Lxu.d[d]);
  for(int d = 0; d < NX; d++) o[E::VX + d] = double(Vx.d[d]);
  for(int d = 0; d < NX * NX; d++) o[E::VXX + d] = double(Vxx.d[d]);
  if constexpr(HasIneq<M>::value)
  {
    constexpr int NG = M::NG;
    Matrix<S, NG, 1> g = model.ineqConst(tv, xv, uv);
    Matrix<S, NG, NX> C;
    Matrix<S, NG, NU> D;
    model.calcIneqConstDeriv(tv, xv, uv, C, D);
    for(int d = 0; d < NG; d++) o[E::G + d] = double(g.d[d]);
    for(int d = 0; d < NG * NX; d++) o[E::CC + d] = double(C.d[d]);
    for(int d = 0; d < NG * NU; d++) o[E::DD + d] = double(D.d[d]);
  }
}

template<class M>
void modelEval(const double * params,
               int device,
               int n,
               const double * t,
               const double * x,
               const double * u,
               const ModelEvalOutputs & out)
{
  using E = EvalLayout<M>;
  constexpr int NX = M::NX, NU = M::NU, NG = E::NG;
  if(n <= 0) return;
  DeviceGuard guard(device);
  M model = M::fromParams(params);
  DeviceBuffer<double> dt, dx, du, dout;
  dt.allocate(n);
  dx.allocate((size_t)n * NX);
  du.allocate((size_t)n * (NU > 0 ? NU : 1));
  dout.allocate((size_t)n * E::SIZE);
  NMPC_CUDA_CHECK(cudaMemcpy(dt.ptr, t, sizeof(double) * n, cudaMemcpyHostToDevice));
  NMPC_CUDA_CHECK(cudaMemcpy(dx.ptr, x, sizeof(double) * n * NX, cudaMemcpyHostToDevice));
  if(NU > 0) NMPC_CUDA_CHECK(cudaMemcpy(du.ptr, u, sizeof(double) * n * NU, cudaMemcpyHostToDevice));
  model_eval_kernel<M><<<(n + 63) / 64, 64>>>(model, n, dt.ptr, dx.ptr, du.ptr, dout.ptr);
  NMPC_CUDA_CHECK(cudaGetLastError());
  std::vector<double> h((size_t)n * E::SIZE);
  NMPC_CUDA_CHECK(cudaMemcpy(h.data(), dout.ptr, sizeof(double) * h.size(), cudaMemcpyDeviceToHost));
  auto unpack = [&](double * dst, int off, int len) {
    if(!dst) return;
    for(int p = 0; p < n; p++)
      for(int d = 0; d < len; d++) dst[(size_t)p * len + d] = h[(size_t)p * E::SIZE + off + d];
  };
  unpack(out.x_next, E::XN, NX);
  unpack(out.running_cost, E::RC, 1);
  unpack(out.terminal_cost, E::TC, 1);
  unpack(out.Fx, E::FX, NX * NX);
  unpack(out.Fu, E::FU, NX * NU);
  unpack(out.Lx, E::LX, NX);
  unpack(out.Lu, E::LU, NU);
  unpack(out.Lxx, E::LXX, NX * NX);
  unpack(out.Luu, E::LUU, NU * NU);
  unpack(out.Lxu, E::LXU, NX * NU);
  unpack(out.Vx, E::VX, NX);
  unpack(out.Vxx, E::VXX, NX * NX);
  unpack(out.g, E::G, NG);
  unpack(out.C, E::CC, NG * NX);
  unpack(out.D, E::DD, NG * NU);
}
} // namespace nmpc_b200
