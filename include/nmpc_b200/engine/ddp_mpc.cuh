/* nmpc_b200 -- the tick-to-tick part of a receding-horizon (MPC) loop, on the device.
 *
 * The reference's solvers are driven by host loops (TestDDPBipedal.cpp:243-268, TestDDPCartPole.cpp:313-343 and
 * :388-396, TestFmpcOscillator.cpp:166-190): solve, take u_list[0], advance the plant, build the next warm start.
 * For a batch that loop body is one kernel per tick, so consecutive solves never leave the GPU.
 */
#pragma once

#include <type_traits>
#include <utility>

#include "ddp_kernels.cuh"

namespace nmpc_b200
{
namespace ddp
{
template<class S>
struct MpcParams
{
  int n_ticks;
  int plant; //!< 0: x <- x_list[1]; 1: x <- stateEq(t, x, u, sim_dt) n_substeps times
  int shift_inputs;
  int clamp_u0;
  int n_substeps;
  S tick_dt;
  S sim_dt;
};

/** Device logs of a loop, batch innermost like every other engine array. */
template<class S>
struct MpcLogs
{
  S * x; //!< [n_ticks+1][NX][Bp]
  S * u; //!< [n_ticks][NU][Bp]
  int * iters; //!< [n_ticks][Bp]
  int * status; //!< [n_ticks][Bp]
};

/** Does the functor offer stateEq(t, x, u, dt) (the plant-simulation overload of the reference's test problems,
    e.g. TestDDPCartPole.cpp:69-98)? */
template<class M, class = void>
struct HasStateEqDt : std::false_type
{
};
template<class M>
struct HasStateEqDt<M,
                    std::void_t<decltype(std::declval<const M &>().stateEq(std::declval<typename M::Scalar>(),
                                                                           std::declval<const typename M::StateDimVector &>(),
                                                                           std::declval<const typename M::InputDimVector &>(),
                                                                           std::declval<typename M::Scalar>()))>> : std::true_type
{
};

/** After the solve of tick `tick` (at time t): log, apply u_list[0], advance current_x, write the next solve's
    inputs (x_list[0] and initial_u_list) into trajectory buffer 0, where K0 expects them. */
template<class M>
__global__ void mpc_advance_kernel(const __grid_constant__ M model,
                                   const __grid_constant__ Workspace<typename M::Scalar> ws,
                                   const __grid_constant__ SolverParams<typename M::Scalar> prm,
                                   const __grid_constant__ MpcParams<typename M::Scalar> mp,
                                   const __grid_constant__ MpcLogs<typename M::Scalar> logs,
                                   int tick,
                                   typename M::Scalar t)
{
  using S = typename M::Scalar;
  constexpr int NX = M::NX, NU = M::NU;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if(b >= ws.B) return;
  const size_t Bp = ws.Bp;
  const int N = prm.N;
  const int sel = ws.sel[b];
  // no __restrict__: with sel == 0 the warm-start shift below reads and writes the same buffer
  const S * xs = ws.x[sel];
  const S * us = ws.u[sel];

  Matrix<S, NX, 1> x;
  Matrix<S, NU, 1> u;
#pragma unroll
  for(int d = 0; d < NX; d++) x[d] = xs[(size_t)d * Bp + b];
#pragma unroll
  for(int d = 0; d < NU; d++)
  {
    S v = us[(size_t)d * Bp + b];
    if(mp.clamp_u0) v = fmin(fmax(v, ws.u_lo[d]), ws.u_hi[d]); // cwiseMax(lower).cwiseMin(upper)
    u[d] = v;
  }
  if(logs.x)
  {
#pragma unroll
    for(int d = 0; d < NX; d++) logs.x[((size_t)tick * NX + d) * Bp + b] = x[d];
  }
  if(logs.u)
  {
#pragma unroll
    for(int d = 0; d < NU; d++) logs.u[((size_t)tick * NU + d) * Bp + b] = u[d];
  }
  if(logs.iters) logs.iters[(size_t)tick * Bp + b] = ws.iters[b];
  if(logs.status) logs.status[(size_t)tick * Bp + b] = ws.status[b];

  // plant
  if(mp.plant == 0)
  {
#pragma unroll
    for(int d = 0; d < NX; d++) x[d] = xs[((size_t)NX + d) * Bp + b];
  }
  else
  {
    if constexpr(HasStateEqDt<M>::value)
    {
      for(int s = 0; s < mp.n_substeps; s++) x = model.stateEq(t + s * mp.sim_dt, x, u, mp.sim_dt);
    }
  }
  if(tick == mp.n_ticks - 1)
  {
    if(logs.x)
    {
#pragma unroll
      for(int d = 0; d < NX; d++) logs.x[((size_t)(tick + 1) * NX + d) * Bp + b] = x[d];
    }
    return; // the handle keeps the last solve intact (controlData() after the loop)
  }

  // warm start of the next solve; with sel == 0 the shift is in place (entry i+1 is read before entry i is written)
  S * ud = ws.u[0];
  if(mp.shift_inputs)
  {
    for(int i = 0; i + 1 < N; i++)
    {
#pragma unroll
      for(int d = 0; d < NU; d++) ud[((size_t)i * NU + d) * Bp + b] = us[((size_t)(i + 1) * NU + d) * Bp + b];
    }
    // the new last entry repeats the old one (TestDDPBipedal.cpp:267) ...
    bool repeat_last = true;
    if constexpr(HasInputDim<M>::value)
    {
      // ... unless the input dimension at the new terminal time differs: then it is Zero(terminal_input_dim)
      // (TestDDPVerticalMotion.cpp:306-315)
      const S dt = model.dt();
      repeat_last = model.inputDim(t + (N - 1) * dt) == model.inputDim(t + N * dt);
    }
    if(!repeat_last)
    {
#pragma unroll
      for(int d = 0; d < NU; d++) ud[((size_t)(N - 1) * NU + d) * Bp + b] = S(0);
    }
    else if(sel != 0)
    {
#pragma unroll
      for(int d = 0; d < NU; d++) ud[((size_t)(N - 1) * NU + d) * Bp + b] = us[((size_t)(N - 1) * NU + d) * Bp + b];
    }
  }
  else if(sel != 0)
  {
    for(int i = 0; i < N; i++)
    {
#pragma unroll
      for(int d = 0; d < NU; d++) ud[((size_t)i * NU + d) * Bp + b] = us[((size_t)i * NU + d) * Bp + b];
    }
  }
#pragma unroll
  for(int d = 0; d < NX; d++) ws.x[0][(size_t)d * Bp + b] = x[d];
}
} // namespace ddp
} // namespace nmpc_b200
