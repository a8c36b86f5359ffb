/* nmpc_b200 -- K1 + K2 in one kernel: a producer warp linearises, the consumer warp runs the Riccati sweep.
 *
 * In the three-kernel pipeline K1 writes N derivative tiles per instance to HBM (46 scalars per step for cart-pole:
 * 150 MB per iteration at B=4096, 4.8 GB at B=131072) only for K2 to read them back once.  K2's sweep is one warp's
 * dependent chain (1500 cycles per step, ncu), so three of the SM's four schedulers idle while it runs.  Here a second
 * warp of the CTA evaluates the functor's derivatives for the SAME 32 instances, step N-1 first, straight into a
 * DEPTH-stage shared-memory ring; the consumer warp is ddp::backwardSweep unchanged except for where its tile comes
 * from (ProducerFeed).  Producer and consumer meet at full/empty mbarriers (32 arrivals each: every lane releases its
 * own column of the tile, every lane acquires what it reads).  One step of the producer (sincos + ~150 flops + 46
 * shared stores) is several times shorter than one step of the consumer, so the consumer never waits after the
 * first tile.  What disappears: the K1 launch (35 us of a 212 us iteration at B=4096), the derivative buffer's HBM
 * round trip, and the terminal-derivative buffer (the consumer evaluates calcTerminalCostDeriv itself).
 *
 * Every scalar is computed by the same expression as in linearize_kernel / backward_kernel, so results are
 * bit-identical to the three-kernel pipeline.
 *
 * Reference: DDPSolver.hpp:157-185 (Step 1), :188-231 (Step 2), :343-534 (backwardPass).
 */
#pragma once

#include "ddp_kernels.cuh"

namespace nmpc_b200
{
namespace ddp
{
#ifndef NMPC_B200_FUSED_DEPTH
#  define NMPC_B200_FUSED_DEPTH 2
#endif
/** Ring stages between producer and consumer.  The ring (11.8 KB per stage for cart-pole) caps the CTAs per SM at
    large batch: 4 stages -> 4 CTAs per SM by shared memory, 2 stages -> 5 (then registers bound). */
constexpr int kFusedDepth = NMPC_B200_FUSED_DEPTH;

template<class M>
struct FusedLayout
{
  using L = BlockLayout<M::NX, M::NU>;
  static constexpr size_t ringBytes()
  {
    return sizeof(typename M::Scalar) * (size_t)kFusedDepth * L::SIZE * kTile;
  }
  static constexpr size_t bytes()
  {
    return ringBytes() + sizeof(unsigned long long) * 2 * kFusedDepth + 16;
  }
};

/** Producer side of one sweep: tiles of steps N-1 ... 0 for this lane's instance. */
template<class M>
__device__ __forceinline__ void produceSweep(const M & model_in_constant_bank,
                                             const Workspace<typename M::Scalar> & ws,
                                             const SolverParams<typename M::Scalar> & prm,
                                             int b,
                                             int lane,
                                             const typename M::Scalar * __restrict__ xs,
                                             const typename M::Scalar * __restrict__ us,
                                             typename M::Scalar * __restrict__ ring,
                                             unsigned long long * full,
                                             unsigned long long * empty,
                                             unsigned & fill)
{
  using S = typename M::Scalar;
  constexpr int NX = M::NX, NU = M::NU;
  using L = BlockLayout<NX, NU>;
  const M model = model_in_constant_bank; // out of the kernel-parameter constant bank, once (see forwardRolloutRing)
  const S t0 = prm.t0;
  const size_t Bp = ws.Bp;
  const int N = prm.N;

  // (x_i, u_i) are fetched two steps ahead of their use
  S xb[3][NX], ub[3][NU];
  auto load = [&](int slot, int step) {
    const int i = step > 0 ? step : 0;
#pragma unroll
    for(int d = 0; d < NX; d++) xb[slot][d] = xs[((size_t)i * NX + d) * Bp + b];
#pragma unroll
    for(int d = 0; d < NU; d++) ub[slot][d] = us[((size_t)i * NU + d) * Bp + b];
  };
  load(0, N - 1);
  load(1, N - 2);

  for(int i = N - 1; i >= 0; i--)
  {
    load(2, i - 2);
    Matrix<S, NX, 1> x;
    Matrix<S, NU, 1> u;
#pragma unroll
    for(int d = 0; d < NX; d++) x[d] = xb[0][d];
#pragma unroll
    for(int d = 0; d < NU; d++) u[d] = ub[0][d];

    Matrix<S, NX, NX> Fx, Lxx;
    Matrix<S, NX, NU> Fu, Lxu;
    Matrix<S, NX, 1> Lx;
    Matrix<S, NU, 1> Lu;
    Matrix<S, NU, NU> Luu;
    linearizeStep<M>(model, t0 + i * model.dt(), x, u, Fx, Fu, Lx, Lu, Lxx, Luu, Lxu);

    const unsigned st = fill % kFusedDepth;
    if(fill >= (unsigned)kFusedDepth) mbarWait(&empty[st], ((fill / kFusedDepth) - 1u) & 1u); // the consumer is done with it
    S * blk = ring + (size_t)st * L::SIZE * kTile + lane;
#pragma unroll
    for(int d = 0; d < NX * NX; d++) blk[(L::FX + d) * kTile] = Fx.d[d];
#pragma unroll
    for(int d = 0; d < NX * NU; d++) blk[(L::FU + d) * kTile] = Fu.d[d];
#pragma unroll
    for(int d = 0; d < NX; d++) blk[(L::LX + d) * kTile] = Lx.d[d];
#pragma unroll
    for(int d = 0; d < NU; d++) blk[(L::LU + d) * kTile] = Lu.d[d];
#pragma unroll
    for(int d = 0; d < NX * NX; d++) blk[(L::LXX + d) * kTile] = Lxx.d[d];
#pragma unroll
    for(int d = 0; d < NU * NU; d++) blk[(L::LUU + d) * kTile] = Luu.d[d];
#pragma unroll
    for(int d = 0; d < NX * NU; d++) blk[(L::LXU + d) * kTile] = Lxu.d[d];
    mbarArrive(&full[st]); // release: this lane's column of the tile
    fill++;

#pragma unroll
    for(int d = 0; d < NX; d++)
    {
      xb[0][d] = xb[1][d];
      xb[1][d] = xb[2][d];
    }
#pragma unroll
    for(int d = 0; d < NU; d++)
    {
      ub[0][d] = ub[1][d];
      ub[1][d] = ub[2][d];
    }
  }
}

/** procOnce() Steps 1-2 (DDPSolver.hpp:157-231) for one 32-instance tile: warp 0 consumes, warp 1 produces. */
template<class M, bool CONSTRAINED>
__global__ void __launch_bounds__(64) backward_fused_kernel(const __grid_constant__ M model,
                                                            const __grid_constant__ Workspace<typename M::Scalar> ws,
                                                            const __grid_constant__ SolverParams<typename M::Scalar> prm,
                                                            int iter)
{
  pdlPrologue();
  using S = typename M::Scalar;
  using L = BlockLayout<M::NX, M::NU>;
  constexpr unsigned kFull = 0xffffffffu;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  S * ring = reinterpret_cast<S *>(smem_raw);
  unsigned long long * full = reinterpret_cast<unsigned long long *>(smem_raw + FusedLayout<M>::ringBytes());
  unsigned long long * empty = full + kFusedDepth;
  if(threadIdx.x == 0)
  {
    for(int st = 0; st < kFusedDepth; st++)
    {
      mbarInit(&full[st], 32);
      mbarInit(&empty[st], 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    if(blockIdx.x == 0) *ws.fan_count = 0; // the previous iteration's line-search work list is consumed (was K1's job)
  }

  const int bg = blockIdx.x * kTile + lane; // ws.Bp is a multiple of 128: padded lanes read valid memory, never write
  const int b = bg;
  const bool live = (bg < ws.B) && (ws.status[bg < ws.B ? bg : 0] == 0);
  // both warps see the same 32 verdicts, so this exit is uniform over the CTA; it is also the barrier that
  // publishes the mbarrier initialisation
  if(!__syncthreads_or(live)) return;

  const int sel = live ? ws.sel[b] : 0;
  const S * us = ws.u[sel];
  const S * xs = ws.x[sel];
  unsigned fill = 0;

  // warp 1 produces, warp 0 consumes; both meet at ONE barrier after every sweep, where the consumer's lanes vote on
  // whether lambda must grow and the sweep be repeated
  S lambda = (warp == 0 && live) ? ws.lambda[b] : S(0);
  S dlambda = (warp == 0 && live) ? ws.dlambda[b] : S(0);
  int n_bwd = (warp == 0 && live) ? ws.n_bwd[b] : 0;
  S dV0 = S(0), dV1 = S(0), k_rel_norm = S(0);
  bool need = (warp == 0) && live;
  bool failed = false;
  ProducerFeed<S, L::SIZE, kFusedDepth> feed{ring, full, empty, fill, lane};
  while(true)
  {
    if(warp == 1)
    {
      produceSweep<M>(model, ws, prm, b, lane, xs, us, ring, full, empty, fill);
    }
    else
    {
      if(need) n_bwd++;
      const bool ok = backwardSweep<M, CONSTRAINED>(model, ws, prm, b, lane, us, xs, feed, need, lambda, dV0, dV1, k_rel_norm);
      if(need)
      {
        if(ok)
        {
          need = false;
        }
        else
        {
          // increase lambda (:194-204)
          dlambda = fmax(dlambda * prm.lambda_factor, prm.lambda_factor);
          lambda = fmax(lambda * dlambda, prm.lambda_min);
          if(lambda > prm.lambda_max)
          {
            failed = true;
            need = false;
          }
        }
      }
    }
    if(!__syncthreads_or(need ? 1 : 0)) break;
  }
  if(warp == 1) return;
  if(!live) return;
  ws.n_bwd[b] = n_bwd;
  ws.lambda[b] = lambda;
  ws.dlambda[b] = dlambda;
  if(failed)
  {
    // return -1 before k_rel_norm / cost / lambda of the trace entry are written (:203)
    ws.status[b] = -1;
    ws.iters[b] = iter;
    writeTrace<S>(ws, b, iter, S(iter), S(0), S(0), S(0), S(0), S(0), S(0), S(0), S(0));
    return;
  }
  ws.dV[b] = dV0;
  ws.dV[(size_t)ws.Bp + b] = dV1;
  if(k_rel_norm < prm.k_rel_norm_thre && lambda < prm.lambda_thre)
  {
    // return 1 with only iter and k_rel_norm set in the trace entry (:222-230)
    ws.status[b] = 1;
    ws.iters[b] = iter;
    writeTrace<S>(ws, b, iter, S(iter), S(0), S(0), S(0), S(0), k_rel_norm, S(0), S(0), S(0));
    return;
  }
  // hand k_rel_norm to the forward kernel through the trace row
  ws.trace[((size_t)iter * kTraceFields + 5) * ws.Bp + b] = k_rel_norm;
}
} // namespace ddp
} // namespace nmpc_b200
