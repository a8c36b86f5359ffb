/* nmpc_b200 -- tick-to-tick part of the FMPC receding-horizon loops, on the device
 * (TestFmpcOscillator.cpp:166-190, TestFmpcCartPole.cpp:344-357 + :405-412): take u_list[0], optionally add the
 * feedback term K_0 (x_list[0] - current_x), integrate the plant with sim_dt, keep the Variable as the next warm start.
 */
#pragma once

#include "ddp_mpc.cuh"
#include "fmpc_kernels.cuh"

namespace nmpc_b200
{
namespace fmpc
{
template<class S>
struct MpcLogs
{
  S * x; //!< [n_ticks+1][NX][Bp]
  S * u; //!< [n_ticks][NU][Bp]  u_list[0] of every tick
  S * kkt; //!< [n_ticks][Bp]     traceDataList().back().kkt_error
  int * status; //!< [n_ticks][Bp]
};

template<class M>
__global__ void fmpc_mpc_advance_kernel(const __grid_constant__ M model,
                                        const __grid_constant__ Workspace<typename M::Scalar> ws,
                                        const __grid_constant__ ddp::MpcParams<typename M::Scalar> mp,
                                        const __grid_constant__ MpcLogs<typename M::Scalar> logs,
                                        int feedback,
                                        int tick,
                                        typename M::Scalar t)
{
  using S = typename M::Scalar;
  constexpr int NX = M::NX, NU = M::NU;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if(b >= ws.B) return;
  const size_t Bp = ws.Bp;

  Matrix<S, NX, 1> x, xl0;
  Matrix<S, NU, 1> u0;
#pragma unroll
  for(int d = 0; d < NX; d++) x[d] = ws.x0[(size_t)d * Bp + b];
#pragma unroll
  for(int d = 0; d < NX; d++) xl0[d] = ws.x[(size_t)d * Bp + b];
#pragma unroll
  for(int d = 0; d < NU; d++) u0[d] = ws.u[(size_t)d * Bp + b];
  if(logs.x)
  {
#pragma unroll
    for(int d = 0; d < NX; d++) logs.x[((size_t)tick * NX + d) * Bp + b] = x[d];
  }
  if(logs.u)
  {
#pragma unroll
    for(int d = 0; d < NU; d++) logs.u[((size_t)tick * NU + d) * Bp + b] = u0[d];
  }
  if(logs.status) logs.status[(size_t)tick * Bp + b] = ws.status[b];
  if(logs.kkt)
  {
    const int n = ws.n_trace[b];
    logs.kkt[(size_t)tick * Bp + b] = (n > 0) ? ws.trace[((size_t)(n - 1) * kTraceFields + 1) * Bp + b] : S(0);
  }

  if(mp.plant == 0)
  {
#pragma unroll
    for(int d = 0; d < NX; d++) x[d] = ws.x[((size_t)NX + d) * Bp + b];
  }
  else
  {
    if constexpr(ddp::HasStateEqDt<M>::value)
    {
      S K0[NU * NX];
#pragma unroll
      for(int d = 0; d < NU * NX; d++) K0[d] = feedback ? ws.kfb[(size_t)d * Bp + b] : S(0);
      for(int s = 0; s < mp.n_substeps; s++)
      {
        Matrix<S, NU, 1> u = u0;
        if(feedback)
        {
          // u += coeffList().front().K * (variable().x_list[0] - current_x)   (TestFmpcCartPole.cpp:352-355)
#pragma unroll
          for(int a = 0; a < NU; a++)
          {
            S acc = S(0);
#pragma unroll
            for(int j = 0; j < NX; j++) acc += K0[a + j * NU] * (xl0[j] - x[j]);
            u[a] = u[a] + acc;
          }
        }
        x = model.stateEq(t + s * mp.sim_dt, x, u, mp.sim_dt);
      }
    }
  }
  if(logs.x && tick == mp.n_ticks - 1)
  {
#pragma unroll
    for(int d = 0; d < NX; d++) logs.x[((size_t)(tick + 1) * NX + d) * Bp + b] = x[d];
  }
  if(tick == mp.n_ticks - 1) return;
#pragma unroll
  for(int d = 0; d < NX; d++) ws.x0[(size_t)d * Bp + b] = x[d];
}
} // namespace fmpc
} // namespace nmpc_b200
