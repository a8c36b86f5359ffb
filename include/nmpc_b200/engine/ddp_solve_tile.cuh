/* nmpc_b200 -- the whole DDPSolver::solve() of one 32-instance tile in ONE persistent CTA (DDPSolver.hpp:27-340).
 *
 * With a kernel per stage, a latency-bound batch (BASELINE.json configs[1]: 4096 instances = 128 tiles) pays ~9 us of
 * launch gap, prologue (status -> sel -> trajectory pointer chains, mbarrier set-up, first tile of the producers) and
 * drain for each of the three launches of an iteration -- 27 us of a 130 us iteration (tools/exp_variants.py
 * --horizon 2).  Instances never interact, so nothing forces the tiles of a batch to move in lock step: here a CTA of
 * nine warps keeps ITS tile from the initial rollout to the last iteration and changes roles at CTA barriers:
 *
 *   initial rollout (:83-104)       warp 0 rolls out, warp 4 evaluates costs / stores, warp 8 loads   (ddp_forward_split.cuh, INIT)
 *   per iteration
 *     Steps 1-2 (:157-231)          warps 0-3 sweep with four lanes per instance, warps 4-5 linearise (ddp_backward_lanes.cuh)
 *     Step 3, alpha_list[0]         warp 0 / warp 4 / warp 8 as above; failed instances go to a list in shared memory
 *     Step 3, other candidates      rounds of eight listed instances: warps 0-3 roll out (16 candidates lanes per
 *                                   instance), warps 4-7 evaluate, warp 8 loads; winners are copied from scratch
 *     Step 4 (:280-339)             by the cost lanes (lineSearchFinish)
 *
 * A tile leaves as soon as all of its instances have terminated (no host polling for the reference's default
 * max_iter = 500).  All stage code is the code of the stand-alone kernels; the rings' mbarriers keep their phase
 * from stage to stage (every thread counts the ring stages of every role).  Per-iteration stage durations
 * (TraceData::duration_*, DDPSolver.h:208-215) are taken with %globaltimer by thread 0 and reduced over the tiles
 * with atomicMax.
 */
#pragma once

#include "ddp_backward_lanes.cuh"
#include "ddp_forward_split.cuh"

namespace nmpc_b200
{
namespace ddp
{
constexpr int kTileWarps = 2 * kFanWarps + 1; //!< warps of a persistent CTA
constexpr int kTileProducers = 2; //!< of which linearise during the backward pass

template<class M>
struct TileSmem
{
  using S = typename M::Scalar;
  using SL = SplitLayout<M>;
  static constexpr size_t firstInBytes()
  {
    return ((sizeof(S) * SL::inElems(kTile) + sizeof(unsigned long long) * 2 * kSplitIn + 127) / 128) * 128;
  }
  static constexpr size_t listBytes()
  {
    return 256; // 32 listed instances + their count
  }
  static constexpr size_t bytes()
  {
    return LaneLayout<M>::bytes() + FanSmem<M>::bytes() + firstInBytes() + listBytes();
  }
};

__device__ __forceinline__ unsigned long long globalTimerNs()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;\n" : "=l"(t));
  return t;
}

/* The stages are separate (non-inlined) functions: each gets the registers IT needs.  Inlined into one kernel body,
   the nine-warp CTA's limit of 168 registers per thread (three warps share one of the SM's four 16 K register files)
   made ptxas spill inside the sweep's step loop (+9 local-memory accesses per step, profiles/r2_tile_*). */

/** Steps 1-2 of one iteration for the tile starting at instance tb; returns the ring's tile count afterwards. */
template<class M, bool CONSTRAINED, class XCH>
__device__ __noinline__ unsigned tileBackwardPhase(const M * model_p,
                                                   const Workspace<typename M::Scalar> * ws_p,
                                                   const SolverParams<typename M::Scalar> * prm_p,
                                                   unsigned char * smem_raw,
                                                   int tb,
                                                   int iter,
                                                   unsigned bwd_fill)
{
  using LL = LaneLayout<M>;
  constexpr int G = LL::G, IPW = LL::IPW, CW = LL::CW;
  constexpr int P = kTileProducers;
  const Workspace<typename M::Scalar> & ws = *ws_p;
  const SolverParams<typename M::Scalar> & prm = *prm_p;
  const LaneSmem<M> sm_bwd(smem_raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int role = (warp < CW) ? kLaneConsumer : (warp < CW + P ? kLaneProducer : kLaneIdle);
  const int t = (role == kLaneConsumer) ? warp * IPW + lane / G : lane;
  const int b = tb + t;
  const bool live = (b < ws.B) && (ws.status[b < ws.B ? b : 0] == 0);
  const int sel = live ? ws.sel[b] : 0;
  laneBackward<M, CONSTRAINED, P, XCH>(*model_p, ws, prm, sm_bwd, role, b, t, lane, warp - CW, live, ws.x[sel], ws.u[sel], iter,
                                       bwd_fill);
  return bwd_fill;
}

/** The initial rollout (INIT) or the first line-search candidate of the tile: warp 0 rolls out, warp kFanWarps
    evaluates / stores / decides, warp 2 kFanWarps loads.  Instances whose first candidate fails are listed. */
template<class M, bool INIT>
__device__ __noinline__ void tileFirstPhase(const M * model_p,
                                            const Workspace<typename M::Scalar> * ws_p,
                                            const SolverParams<typename M::Scalar> * prm_p,
                                            unsigned char * smem_raw,
                                            int tb,
                                            int iter,
                                            unsigned first_in_n,
                                            unsigned out0_n)
{
  using S = typename M::Scalar;
  constexpr int NX = M::NX;
  using LL = LaneLayout<M>;
  using SL = SplitLayout<M>;
  using O = typename SL::O;
  const Workspace<S> & ws = *ws_p;
  const SolverParams<S> & prm = *prm_p;
  const FanSmem<M> sm_fan(smem_raw + LL::bytes());
  S * first_in = reinterpret_cast<S *>(smem_raw + LL::bytes() + FanSmem<M>::bytes());
  unsigned long long * first_full = reinterpret_cast<unsigned long long *>(first_in + SL::inElems(kTile));
  unsigned long long * first_empty = first_full + kSplitIn;
  int * list = reinterpret_cast<int *>(smem_raw + LL::bytes() + FanSmem<M>::bytes() + TileSmem<M>::firstInBytes());
  int * n_listed = list + kTile;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int N = prm.N;
  const int bl_raw = tb + lane;
  const int bl = (bl_raw < ws.B) ? bl_raw : (ws.B - 1);
  const int sel = INIT ? 0 : ws.sel[bl];

  if(warp == 2 * kFanWarps)
    splitLoadTile<M>(ws, N, lane, bl, sel, first_in, first_full, first_empty, first_in_n);
  else if(warp == 0)
  {
    const M model = *model_p;
    Matrix<S, NX, 1> x;
#pragma unroll
    for(int d = 0; d < NX; d++) x[d] = ws.x[sel][(size_t)d * ws.Bp + bl];
    splitRollout<M, O::SIZE * kTile, kTile, INIT>(model, prm.t0, N, INIT ? S(0) : prm.alpha_list[0], x, first_in, lane, first_full,
                                                  first_empty, sm_fan.outCol(0, lane), sm_fan.outFull(0), sm_fan.outEmpty(0),
                                                  first_in_n, out0_n);
  }
  else if(warp == kFanWarps)
  {
    const M model = *model_p;
    if constexpr(INIT)
    {
      const bool mine = bl_raw < ws.B;
      const FwdDest<S> dst{ws.x[0], ws.u[0], ws.cost[0], (size_t)ws.Bp, (size_t)bl};
      const S csum = splitCost<M>(model, prm.t0, N, sm_fan.outCol(0, lane), sm_fan.outFull(0), sm_fan.outEmpty(0), mine, dst,
                                  out0_n);
      if(mine)
      {
        // lambda / dlambda are reset by every solve() (:36-38); iter-0 trace entry (:98-104)
        ws.lambda[bl] = prm.initial_lambda;
        ws.dlambda[bl] = prm.initial_dlambda;
        ws.cost_sum[bl] = csum;
        ws.status[bl] = 0;
        ws.sel[bl] = 0;
        ws.iters[bl] = 0;
        ws.n_fwd[bl] = 0;
        ws.n_bwd[bl] = 0;
        writeTrace<S>(ws, bl, 0, S(0), csum, prm.initial_lambda, prm.initial_dlambda, S(0), S(0), S(0), S(0), S(0));
      }
    }
    else
    {
      const bool active = (bl_raw < ws.B) && (ws.status[bl] == 0);
      const bool work = active && prm.n_alpha > 0;
      const S alpha = prm.alpha_list[0];
      const S cost_new = splitCost<M>(model, prm.t0, N, sm_fan.outCol(0, lane), sm_fan.outFull(0), sm_fan.outEmpty(0), work,
                                      candidateBuffer<S>(ws, sel, bl), out0_n);
      const S cost_cur = ws.cost_sum[bl];
      S actual = S(0), expected = S(0), ratio = S(0);
      bool success = false;
      if(work)
        success = lineSearchTest<S>(prm, cost_cur, cost_new, alpha, ws.dV[bl], ws.dV[(size_t)ws.Bp + bl], actual, expected,
                                    ratio);
      const bool decided = success || prm.n_alpha <= 1;
      if(active && decided)
        lineSearchFinish<S>(ws, prm, bl, iter, sel, success, work ? alpha : S(0), actual, expected, ratio, cost_cur, cost_new,
                            work ? 1 : 0);
      // the others queue up for the remaining candidates
      const unsigned listed = __ballot_sync(0xffffffffu, active && !decided);
      if(active && !decided) list[__popc(listed & ((1u << lane) - 1u))] = bl;
      if(lane == 0) *n_listed = __popc(listed);
    }
  }
}

/** One round of the remaining candidates for the listed instances list[slot0 ..). */
template<class M>
__device__ __noinline__ void tileFanPhase(const M * model_p,
                                          const Workspace<typename M::Scalar> * ws_p,
                                          const SolverParams<typename M::Scalar> * prm_p,
                                          const FwdFanout<typename M::Scalar> * fan_p,
                                          unsigned char * smem_raw,
                                          int tb,
                                          int iter,
                                          int n_list,
                                          int slot0,
                                          unsigned fan_in_n,
                                          unsigned out_n)
{
  using LL = LaneLayout<M>;
  const FanSmem<M> sm_fan(smem_raw + LL::bytes());
  const int * list = reinterpret_cast<const int *>(smem_raw + LL::bytes() + FanSmem<M>::bytes() + TileSmem<M>::firstInBytes());
  fanoutRound<M>(*model_p, *ws_p, *prm_p, *fan_p, sm_fan, iter, list, n_list, slot0, (size_t)(tb + slot0), threadIdx.x >> 5,
                 threadIdx.x & 31, fan_in_n, out_n);
}

/** stage_ns: [max_iter + 1][4] or nullptr; row 0 = {0, initial rollout, ...}, row i = {Steps 1-2, Steps 3-4, ...} of
    iteration i: columns 0-1 the maximum over the tiles of the batch, columns 2-3 the sum over the tiles. */
template<class M, bool CONSTRAINED, class XCH>
__global__ void __launch_bounds__(kTileWarps * 32, 1)
    ddp_solve_tile_kernel(const __grid_constant__ M model_in_constant_bank,
                          const __grid_constant__ Workspace<typename M::Scalar> ws,
                          const __grid_constant__ SolverParams<typename M::Scalar> prm,
                          const __grid_constant__ FwdFanout<typename M::Scalar> fan,
                          unsigned long long * stage_ns)
{
  using S = typename M::Scalar;
  constexpr int NX = M::NX;
  using LL = LaneLayout<M>;
  using SL = SplitLayout<M>;
  using O = typename SL::O;
  constexpr int G = LL::G, IPW = LL::IPW, CW = LL::CW;
  constexpr int P = kTileProducers;
  static_assert(CW + P <= kTileWarps, "backward roles exceed the CTA");
  static_assert(NX <= G, "one column per lane");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const LaneSmem<M> sm_bwd(smem_raw);
  const FanSmem<M> sm_fan(smem_raw + LL::bytes());
  S * first_in = reinterpret_cast<S *>(smem_raw + LL::bytes() + FanSmem<M>::bytes());
  unsigned long long * first_full = reinterpret_cast<unsigned long long *>(first_in + SL::inElems(kTile));
  unsigned long long * first_empty = first_full + kSplitIn;
  int * list = reinterpret_cast<int *>(smem_raw + LL::bytes() + FanSmem<M>::bytes() + TileSmem<M>::firstInBytes());
  int * n_listed = list + kTile;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tb = blockIdx.x * kTile; // first instance of the tile; ws.Bp is a multiple of 128, so tb + 31 < Bp
  if(threadIdx.x == 0)
  {
    sm_bwd.initBarriers();
    sm_fan.initBarriers();
    for(int st = 0; st < kSplitIn; st++)
    {
      mbarInit(&first_full[st], 32);
      mbarInit(&first_empty[st], 32);
    }
    *n_listed = 0;
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  // ring stages passed so far (every thread counts for every ring: roles change from stage to stage)
  unsigned bwd_fill = 0, first_in_n = 0, fan_in_n = 0, out0_n = 0, outx_n = 0;
  const unsigned in_stages = splitInStages(prm.N), out_stages = splitOutStages(prm.N);
  // the instance this lane serves in the thread-per-instance roles (rollout / cost / loader of the first candidate)
  const int bl_raw = tb + lane;
  const int bl = (bl_raw < ws.B) ? bl_raw : (ws.B - 1);
  unsigned long long t_mark = 0;
  if(stage_ns != nullptr && threadIdx.x == 0) t_mark = globalTimerNs();
  auto stamp = [&](int slot) {
    if(stage_ns != nullptr && threadIdx.x == 0)
    {
      const unsigned long long now = globalTimerNs();
      atomicMax(&stage_ns[slot], now - t_mark);
      atomicAdd(&stage_ns[slot + 2], now - t_mark);
      t_mark = now;
    }
  };

  // ------------------------------------------------------------------ solve(): initial rollout (:83-104)
  tileFirstPhase<M, true>(&model_in_constant_bank, &ws, &prm, smem_raw, tb, 0, first_in_n, out0_n);
  first_in_n += in_stages;
  out0_n += out_stages;
  __syncthreads();
  stamp(1);

  for(int iter = 1; iter <= prm.max_iter; iter++)
  {
    // a tile whose instances have all terminated is done
    const bool running = (bl_raw < ws.B) && (ws.status[bl] == 0);
    if(!__syncthreads_or(running)) break;
    if(threadIdx.x == 0) *n_listed = 0; // rewritten by the cost warp of the first candidate, read two barriers later

    // ---------------------------------------------------------------- Steps 1-2 (:157-231)
    bwd_fill = tileBackwardPhase<M, CONSTRAINED, XCH>(&model_in_constant_bank, &ws, &prm, smem_raw, tb, iter, bwd_fill);
    __syncthreads(); // gains, dV, lambda and the termination verdicts of this tile are visible to all its warps
    stamp(4 * iter);

    // ---------------------------------------------------------------- Step 3, alpha_list[0] (:234-279)
    const bool active = (bl_raw < ws.B) && (ws.status[bl] == 0);
    if(__syncthreads_or(active))
    {
      tileFirstPhase<M, false>(&model_in_constant_bank, &ws, &prm, smem_raw, tb, iter, first_in_n, out0_n);
      first_in_n += in_stages;
      out0_n += out_stages;
    }
    __syncthreads();

    // ---------------------------------------------------------------- Step 3, the other candidates
    const int n_list = *n_listed;
    for(int slot0 = 0; slot0 < n_list; slot0 += FanSmem<M>::IPC)
    {
      tileFanPhase<M>(&model_in_constant_bank, &ws, &prm, &fan, smem_raw, tb, iter, n_list, slot0, fan_in_n,
                      (warp % kFanWarps == 0) ? out0_n : outx_n);
      fan_in_n += in_stages;
      out0_n += out_stages;
      outx_n += out_stages;
    }
    __syncthreads(); // sel / status / lambda written by the cost lanes are visible; the list may be rewritten
    stamp(4 * iter + 1);
  }
}
} // namespace ddp
} // namespace nmpc_b200
