/* nmpc_b200 -- forwardPass() split over three warp roles (procOnce() Steps 3-4, DDPSolver.hpp:234-339, :537-560).
 *
 * One rollout is a dependent chain of N steps, and at a few thousand instances nothing else is in flight on the SM, so
 * the line search costs (instructions on the chain) x (a lone warp's ~3.5 cycles per instruction).  In
 * ddp_forward_phased.cuh the compute warp carries the whole step: u' = u + alpha k + K (x' - x), runningCost, stateEq,
 * the stores of the candidate and their address arithmetic (~190 instructions, 660 cycles per step).  Here a step is
 * split by what is ON the chain:
 *
 *   loader warp    streams {x_i, u_i, k_i, K_i} of the current trajectory into a shared-memory ring (cp.async)
 *   ROLLOUT warp   u'_i, then x'_{i+1} = stateEq(t_i, x'_i, u'_i); publishes (x'_i, u'_i) to a second ring     [chain]
 *   COST warp      runningCost(t_i, x'_i, u'_i), the sum, and every global store of the candidate       [off the chain]
 *
 * A ring stage holds kSPS = 2 steps, so the rollout warp passes its mbarriers once per two steps and the two steps
 * form one basic block.  A functor may split stateEq (RolloutCarry in ddp_kernels.cuh; models/cartpole.h statePrePair):
 * theta_{i+1} = theta_i + dt omega_i needs neither the input nor the trigonometry of step i, so the sincos and
 * reciprocal chains of BOTH steps of a stage are evaluated side by side before the stage's two short state updates
 * (28.9 -> 26.6 us per first-candidate pass; leaving the overlap to the compiler -- one step ahead, chains in separate
 * statements -- was slower than no split at all, 34 us: ptxas emitted the two chains one after the other).
 * Arithmetic and summation order are those of forwardRollout (ddp_kernels.cuh): the same costs and trajectories.
 */
#pragma once

#include "ddp_forward_phased.cuh"

namespace nmpc_b200
{
namespace ddp
{
constexpr int kSPS = 2; //!< steps per ring stage
constexpr int kSplitIn = 4; //!< stages of the loader -> rollout ring
constexpr int kSplitOut = 4; //!< stages of the rollout -> cost ring
/** Rollout warps (= pairs of listed instances) per CTA of the phase-2 KERNEL.  Measured at B = 4096, M-fixed: 4 warps
    (rollout and cost warps sharing the four schedulers) 39.0 us, 2 warps 35.8 us, 1 warp 33.8 us per launch: a rollout
    warp wants a scheduler to itself.  (The persistent tile kernel keeps four: its CTA is one tile's 9 warps.) */
constexpr int kFanSplitWarps = 1;

template<class M>
struct SplitLayout
{
  using S = typename M::Scalar;
  using O = FwdOperands<M::NX, M::NU>;
  static constexpr int OUT = M::NX + M::NU; //!< (x'_i, u'_i)
  /** Per group of `cols` ring columns (one column per rollout lane). */
  static constexpr size_t inElems(int cols)
  {
    return (size_t)kSplitIn * kSPS * O::SIZE * cols;
  }
  static constexpr size_t outElems(int cols)
  {
    return (size_t)kSplitOut * kSPS * OUT * cols;
  }
};

/** Rollout role.  Operand e of step slot q of in-stage st is in_ring[(st * kSPS + q) * IN_STAGE + in_off + e * IN_ES];
    this lane's out column is out_col (element r of slot q of out-stage so at out_col[((so * kSPS + q) * OUT + r) * 32]).
    Every lane of the warp runs the loop. */
template<class M, int IN_STAGE, int IN_ES, bool INIT = false>
__device__ __forceinline__ void splitRollout(const M & model,
                                             typename M::Scalar t0,
                                             int N,
                                             typename M::Scalar alpha,
                                             Matrix<typename M::Scalar, M::NX, 1> x,
                                             const typename M::Scalar * in_ring,
                                             int in_off,
                                             unsigned long long * in_full,
                                             unsigned long long * in_empty,
                                             typename M::Scalar * out_col,
                                             unsigned long long * out_full,
                                             unsigned long long * out_empty,
                                             unsigned in_base = 0,
                                             unsigned out_base = 0)
{
  using S = typename M::Scalar;
  constexpr int NX = M::NX, NU = M::NU;
  using O = FwdOperands<NX, NU>;
  constexpr int OUT = NX + NU;
  const S dt = model.dt();
  static_assert(kSPS == 2, "RolloutCarry prepares two steps at a time");
  RolloutCarry<M> carry;

  auto step = [&](const S * op, S * oc, S t, int q) {
    S xr[NX], ur[NU], kr[NU], Kr[NU * NX];
#pragma unroll
    for(int d = 0; d < NX; d++) xr[d] = op[(size_t)(O::X + d) * IN_ES];
#pragma unroll
    for(int d = 0; d < NU; d++) ur[d] = op[(size_t)(O::U + d) * IN_ES];
#pragma unroll
    for(int d = 0; d < NU; d++) kr[d] = op[(size_t)(O::KFF + d) * IN_ES];
#pragma unroll
    for(int d = 0; d < NU * NX; d++) Kr[d] = op[(size_t)(O::KFB + d) * IN_ES];
    Matrix<S, NU, 1> u;
    if constexpr(INIT)
    {
      // solve()'s initial rollout (:83-96): the given inputs as they are; padding entries of a time-varying input
      // dimension are forced to zero (and stored so by the cost role)
#pragma unroll
      for(int c = 0; c < NU; c++) u[c] = ur[c];
      if constexpr(HasInputDim<M>::value)
      {
        const int nu_act = model.inputDim(t);
#pragma unroll
        for(int c = 0; c < NU; c++)
          if(c >= nu_act) u[c] = S(0);
      }
    }
    else
    {
#pragma unroll
      for(int c = 0; c < NU; c++)
      {
        S acc = S(0);
#pragma unroll
        for(int j = 0; j < NX; j++) acc += Kr[c + j * NU] * (x[j] - xr[j]);
        u[c] = (ur[c] + alpha * kr[c]) + acc; // u' = u + alpha k + K (x' - x)   (:545-546)
      }
    }
#pragma unroll
    for(int d = 0; d < NX; d++) oc[(size_t)d * kTile] = x[d];
#pragma unroll
    for(int d = 0; d < NU; d++) oc[(size_t)(NX + d) * kTile] = u[d];
    x = carry.advance(model, t, x, u, q);
  };

  const int n_pairs = N / kSPS;
  S fi = S(0); // == S(i) exactly
  int g = 0;
  for(; g < n_pairs; g++, fi += S(kSPS))
  {
    const unsigned gi = in_base + (unsigned)g, go = out_base + (unsigned)g;
    const unsigned st = gi % kSplitIn, so = go % kSplitOut;
    carry.prepare(model, t0 + fi * dt, x); // both steps' trigonometry, side by side (placing it after the waits: same time)
    mbarWait(&in_full[st], (gi / kSplitIn) & 1u);
    if(go >= (unsigned)kSplitOut) mbarWait(&out_empty[so], ((go / kSplitOut) - 1u) & 1u);
    const S * op = in_ring + (size_t)st * kSPS * IN_STAGE + in_off;
    S * oc = out_col + (size_t)so * kSPS * OUT * kTile;
#pragma unroll
    for(int q = 0; q < kSPS; q++) step(op + (size_t)q * IN_STAGE, oc + (size_t)q * OUT * kTile, t0 + (fi + S(q)) * dt, q);
    mbarArrive(&in_empty[st]);
    mbarArrive(&out_full[so]);
  }
  // tail: the odd last step (if any) and the terminal state share one out-stage; otherwise the terminal state alone
  {
    const unsigned gi = in_base + (unsigned)g, go = out_base + (unsigned)g;
    const unsigned so = go % kSplitOut;
    if(go >= (unsigned)kSplitOut) mbarWait(&out_empty[so], ((go / kSplitOut) - 1u) & 1u);
    S * oc = out_col + (size_t)so * kSPS * OUT * kTile;
    int q = 0;
    if(N % kSPS != 0)
    {
      const unsigned st = gi % kSplitIn;
      mbarWait(&in_full[st], (gi / kSplitIn) & 1u);
      carry.prepare(model, t0 + fi * dt, x);
      step(in_ring + (size_t)st * kSPS * IN_STAGE + in_off, oc, t0 + fi * dt, 0);
      mbarArrive(&in_empty[st]);
      q = 1;
    }
#pragma unroll
    for(int d = 0; d < NX; d++) oc[((size_t)q * OUT + d) * kTile] = x[d];
    mbarArrive(&out_full[so]);
  }
}

/** Ring stages one rollout of N steps passes through. */
__host__ __device__ constexpr unsigned splitInStages(int N)
{
  return (unsigned)((N + kSPS - 1) / kSPS);
}
__host__ __device__ constexpr unsigned splitOutStages(int N)
{
  return (unsigned)((N + 1 + kSPS - 1) / kSPS);
}

/** Cost role: running / terminal costs of the candidate published by the rollout role, their sum (in the order of
    forwardRollout), and -- for `store` lanes -- the candidate trajectory written to dst. */
template<class M>
__device__ __forceinline__ typename M::Scalar splitCost(const M & model,
                                                        typename M::Scalar t0,
                                                        int N,
                                                        const typename M::Scalar * out_col,
                                                        unsigned long long * out_full,
                                                        unsigned long long * out_empty,
                                                        bool store,
                                                        const FwdDest<typename M::Scalar> & dst,
                                                        unsigned out_base = 0)
{
  using S = typename M::Scalar;
  constexpr int NX = M::NX, NU = M::NU;
  constexpr int OUT = NX + NU;
  const S dt = model.dt();
  S * xs_ptr = dst.x + dst.col;
  S * us_ptr = dst.u + dst.col;
  S * cs_ptr = dst.c + dst.col;
  const size_t Bd = dst.stride;
  S csum = S(0);
  S fi = S(0);
  const int n_stages = (N + 1 + kSPS - 1) / kSPS;
  for(int g = 0; g < n_stages; g++)
  {
    const unsigned go = out_base + (unsigned)g;
    const unsigned so = go % kSplitOut;
    mbarWait(&out_full[so], (go / kSplitOut) & 1u);
    const S * oc = out_col + (size_t)so * kSPS * OUT * kTile;
    S xv[kSPS][NX], uv[kSPS][NU];
#pragma unroll
    for(int q = 0; q < kSPS; q++)
    {
#pragma unroll
      for(int d = 0; d < NX; d++) xv[q][d] = oc[((size_t)q * OUT + d) * kTile];
#pragma unroll
      for(int d = 0; d < NU; d++) uv[q][d] = oc[((size_t)q * OUT + NX + d) * kTile];
    }
    mbarArrive(&out_empty[so]);
#pragma unroll
    for(int q = 0; q < kSPS; q++)
    {
      const int i = g * kSPS + q;
      if(i > N) break;
      Matrix<S, NX, 1> x;
#pragma unroll
      for(int d = 0; d < NX; d++) x[d] = xv[q][d];
      S c;
      if(i < N)
      {
        Matrix<S, NU, 1> u;
#pragma unroll
        for(int d = 0; d < NU; d++) u[d] = uv[q][d];
        c = model.runningCost(t0 + fi * dt, x, u);
        if(store)
        {
#pragma unroll
          for(int d = 0; d < NU; d++) us_ptr[(size_t)d * Bd] = u[d];
        }
      }
      else
        c = model.terminalCost(t0 + fi * dt, x);
      if(store)
      {
#pragma unroll
        for(int d = 0; d < NX; d++) xs_ptr[(size_t)d * Bd] = x[d];
        *cs_ptr = c;
      }
      csum += c;
      xs_ptr += (size_t)NX * Bd;
      us_ptr += (size_t)NU * Bd;
      cs_ptr += Bd;
      fi += S(1);
    }
  }
  return csum;
}

/** Loader role for one 32-instance tile: lane l streams the operands of its own instance, kSPS steps per stage. */
template<class M>
__device__ __forceinline__ void splitLoadTile(const Workspace<typename M::Scalar> & ws,
                                              int N,
                                              int lane,
                                              int b,
                                              int sel,
                                              typename M::Scalar * in_ring,
                                              unsigned long long * in_full,
                                              unsigned long long * in_empty,
                                              unsigned in_base = 0)
{
  using S = typename M::Scalar;
  constexpr int NX = M::NX, NU = M::NU;
  using O = FwdOperands<NX, NU>;
  const size_t Bp = ws.Bp;
  const S * row_ptr[O::SIZE];
  long long row_stride[O::SIZE];
#pragma unroll
  for(int e = 0; e < O::SIZE; e++)
  {
    if(e < O::U)
    {
      row_ptr[e] = ws.x[sel] + (size_t)(e - O::X) * Bp + b;
      row_stride[e] = (long long)NX * (long long)Bp;
    }
    else if(e < O::KFF)
    {
      row_ptr[e] = ws.u[sel] + (size_t)(e - O::U) * Bp + b;
      row_stride[e] = (long long)NU * (long long)Bp;
    }
    else if(e < O::KFB)
    {
      row_ptr[e] = ws.kff + (size_t)(e - O::KFF) * Bp + b;
      row_stride[e] = (long long)NU * (long long)Bp;
    }
    else
    {
      row_ptr[e] = ws.kfb + (size_t)(e - O::KFB) * Bp + b;
      row_stride[e] = (long long)(NU * NX) * (long long)Bp;
    }
  }
  const int n_fills = (N + kSPS - 1) / kSPS;
  for(int f = 0; f < n_fills; f++)
  {
    const unsigned fg = in_base + (unsigned)f;
    const unsigned st = fg % kSplitIn;
    if(fg >= (unsigned)kSplitIn) mbarWait(&in_empty[st], ((fg / kSplitIn) - 1u) & 1u);
    S * dstp = in_ring + (size_t)st * kSPS * O::SIZE * kTile + lane;
#pragma unroll
    for(int q = 0; q < kSPS; q++)
    {
      const bool in_range = f * kSPS + q < N; // the slot past an odd horizon's last step is never read
#pragma unroll
      for(int e = 0; e < O::SIZE; e++)
      {
        if(in_range)
        {
          if constexpr(sizeof(S) == 8)
            cpAsync8(dstp + ((size_t)q * O::SIZE + e) * kTile, row_ptr[e]);
          else
            cpAsync4(dstp + ((size_t)q * O::SIZE + e) * kTile, row_ptr[e]);
        }
        row_ptr[e] += row_stride[e];
      }
    }
    cpAsyncArriveOn(&in_full[st]);
  }
}

/** Phase 1 of the line search: alpha_list[0] for every running instance of a 32-instance tile.
    Warp 0 rolls out, warp 1 evaluates costs / stores / decides, warp 2 loads. */
template<class M>
__global__ void __launch_bounds__(96) forward_first_split_kernel(const __grid_constant__ M model_in_constant_bank,
                                                                  const __grid_constant__ Workspace<typename M::Scalar> ws,
                                                                  const __grid_constant__ SolverParams<typename M::Scalar> prm,
                                                                  const __grid_constant__ FwdFanout<typename M::Scalar> fan,
                                                                  int iter)
{
  pdlPrologue();
  using S = typename M::Scalar;
  constexpr int NX = M::NX;
  using SL = SplitLayout<M>;
  using O = typename SL::O;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  S * in_ring = reinterpret_cast<S *>(smem_raw);
  S * out_ring = in_ring + SL::inElems(kTile);
  unsigned long long * in_full = reinterpret_cast<unsigned long long *>(out_ring + SL::outElems(kTile));
  unsigned long long * in_empty = in_full + kSplitIn;
  unsigned long long * out_full = in_empty + kSplitIn;
  unsigned long long * out_empty = out_full + kSplitOut;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if(threadIdx.x == 0)
  {
    for(int st = 0; st < kSplitIn; st++)
    {
      mbarInit(&in_full[st], 32);
      mbarInit(&in_empty[st], 32);
    }
    for(int st = 0; st < kSplitOut; st++)
    {
      mbarInit(&out_full[st], 32);
      mbarInit(&out_empty[st], 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  const int bg = blockIdx.x * kTile + lane;
  const int b = (bg < ws.B) ? bg : (ws.B - 1);
  const bool active = (bg < ws.B) && (ws.status[b] == 0);
  // all three warps see the same 32 verdicts: uniform exit; the barrier also publishes the mbarrier initialisation
  if(!__syncthreads_or(active)) return;
  const int sel = ws.sel[b];
  const int N = prm.N;

  if(warp == 2)
  {
    splitLoadTile<M>(ws, N, lane, b, sel, in_ring, in_full, in_empty);
    return;
  }
  using LM = typename LatencyOf<M>::type;
  const LM model(model_in_constant_bank);
  const S alpha = prm.alpha_list[0];
  if(warp == 0)
  {
    Matrix<S, NX, 1> x;
#pragma unroll
    for(int d = 0; d < NX; d++) x[d] = ws.x[sel][(size_t)d * ws.Bp + b];
    splitRollout<LM, O::SIZE * kTile, kTile>(model, prm.t0, N, alpha, x, in_ring, lane, in_full, in_empty, out_ring + lane,
                                            out_full, out_empty);
    return;
  }
  const bool work = active && prm.n_alpha > 0;
  const S cost_new = splitCost<LM>(model, prm.t0, N, out_ring + lane, out_full, out_empty, work, candidateBuffer<S>(ws, sel, b));
  if(!active) return;
  const S cost_cur = ws.cost_sum[b];
  S actual = S(0), expected = S(0), ratio = S(0);
  bool success = false;
  if(work) success = lineSearchTest<S>(prm, cost_cur, cost_new, alpha, ws.dV[b], ws.dV[(size_t)ws.Bp + b], actual,
                                       expected, ratio);
  if(success || prm.n_alpha <= 1)
  {
    lineSearchFinish<S>(ws, prm, b, iter, sel, success, work ? alpha : S(0), actual, expected, ratio, cost_cur,
                        cost_new, work ? 1 : 0);
    return;
  }
  const int slot = atomicAdd(fan.count, 1);
  fan.list[slot] = b;
}

/** K0 with the same three roles (solve()'s initial rollout, DDPSolver.hpp:36-38, :83-104): the given inputs as they
    are, costs and trajectory into buffer 0, lambda / dlambda / counters reset, iter-0 trace entry.  The loader streams
    the whole operand tile although only u_i is used (k, K rows hold whatever the previous solve left: never read by
    the INIT rollout). */
template<class M>
__global__ void __launch_bounds__(96) rollout_init_split_kernel(const __grid_constant__ M model_in_constant_bank,
                                                                 const __grid_constant__ Workspace<typename M::Scalar> ws,
                                                                 const __grid_constant__ SolverParams<typename M::Scalar> prm)
{
  pdlPrologue();
  using S = typename M::Scalar;
  constexpr int NX = M::NX;
  using SL = SplitLayout<M>;
  using O = typename SL::O;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  S * in_ring = reinterpret_cast<S *>(smem_raw);
  S * out_ring = in_ring + SL::inElems(kTile);
  unsigned long long * in_full = reinterpret_cast<unsigned long long *>(out_ring + SL::outElems(kTile));
  unsigned long long * in_empty = in_full + kSplitIn;
  unsigned long long * out_full = in_empty + kSplitIn;
  unsigned long long * out_empty = out_full + kSplitOut;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if(threadIdx.x == 0)
  {
    for(int st = 0; st < kSplitIn; st++)
    {
      mbarInit(&in_full[st], 32);
      mbarInit(&in_empty[st], 32);
    }
    for(int st = 0; st < kSplitOut; st++)
    {
      mbarInit(&out_full[st], 32);
      mbarInit(&out_empty[st], 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    if(blockIdx.x == 0) *ws.fan_count = 0;
  }
  __syncthreads();
  const int bg = blockIdx.x * kTile + lane;
  const int b = (bg < ws.B) ? bg : (ws.B - 1);
  const bool mine = bg < ws.B;
  const int N = prm.N;
  if(warp == 2)
  {
    splitLoadTile<M>(ws, N, lane, b, 0, in_ring, in_full, in_empty);
    return;
  }
  using LM = typename LatencyOf<M>::type;
  const LM model(model_in_constant_bank);
  if(warp == 0)
  {
    Matrix<S, NX, 1> x;
#pragma unroll
    for(int d = 0; d < NX; d++) x[d] = ws.x[0][(size_t)d * ws.Bp + b];
    splitRollout<LM, O::SIZE * kTile, kTile, true>(model, prm.t0, N, S(0), x, in_ring, lane, in_full, in_empty, out_ring + lane,
                                                  out_full, out_empty);
    return;
  }
  const FwdDest<S> dst{ws.x[0], ws.u[0], ws.cost[0], (size_t)ws.Bp, (size_t)b};
  const S csum = splitCost<LM>(model, prm.t0, N, out_ring + lane, out_full, out_empty, mine, dst);
  if(!mine) return;
  ws.lambda[b] = prm.initial_lambda;
  ws.dlambda[b] = prm.initial_dlambda;
  ws.cost_sum[b] = csum;
  ws.status[b] = 0;
  ws.sel[b] = 0;
  ws.iters[b] = 0;
  ws.n_fwd[b] = 0;
  ws.n_bwd[b] = 0;
  writeTrace<S>(ws, b, 0, S(0), csum, prm.initial_lambda, prm.initial_dlambda, S(0), S(0), S(0), S(0), S(0));
}

/** Shared-memory carve-up of phase 2: one broadcast in-ring for the CTA's eight listed instances and one out-ring per
    rollout / cost warp pair. */
template<class M, int W = kFanWarps>
struct FanSmem
{
  using S = typename M::Scalar;
  using SL = SplitLayout<M>;
  static constexpr int IPW = 32 / kFanLanes; //!< listed instances per rollout warp
  static constexpr int IPC = W * IPW; //!< ... per CTA and round
  static constexpr int ROWS = IPC * SL::O::SIZE; //!< in-ring rows of one step: [instance][operand]
  S * in_ring; //!< [kSplitIn][kSPS][ROWS]
  S * out_ring; //!< [W][kSplitOut][kSPS][OUT][32]
  unsigned long long * in_full;
  unsigned long long * in_empty;
  unsigned long long * out_bars; //!< per pair: kSplitOut full, kSplitOut empty
  static constexpr size_t bytes()
  {
    return ((sizeof(S) * ((size_t)kSplitIn * kSPS * ROWS + (size_t)W * SL::outElems(kTile))
             + sizeof(unsigned long long) * (2 * kSplitIn + 2 * kSplitOut * W) + 127)
            / 128)
           * 128;
  }
  __device__ __forceinline__ explicit FanSmem(unsigned char * base)
  {
    in_ring = reinterpret_cast<S *>(base);
    out_ring = in_ring + (size_t)kSplitIn * kSPS * ROWS;
    in_full = reinterpret_cast<unsigned long long *>(out_ring + (size_t)W * SL::outElems(kTile));
    in_empty = in_full + kSplitIn;
    out_bars = in_empty + kSplitIn;
  }
  /** One thread, before a CTA barrier. */
  __device__ __forceinline__ void initBarriers() const
  {
    for(int st = 0; st < kSplitIn; st++)
    {
      mbarInit(&in_full[st], 32);
      mbarInit(&in_empty[st], W * 32);
    }
    for(int st = 0; st < 2 * kSplitOut * W; st++) mbarInit(&out_bars[st], 32);
  }
  __device__ __forceinline__ S * outCol(int pair, int lane) const
  {
    return out_ring + (size_t)pair * SL::outElems(kTile) + lane;
  }
  __device__ __forceinline__ unsigned long long * outFull(int pair) const
  {
    return out_bars + (size_t)pair * 2 * kSplitOut;
  }
  __device__ __forceinline__ unsigned long long * outEmpty(int pair) const
  {
    return outFull(pair) + kSplitOut;
  }
};

/** One round of phase 2 for the listed instances list[slot0 .. slot0 + IPC) (slots >= count are idle): candidates
    1 .. n_alpha-1 of each at once, 16 lanes per instance.  Warps 0 .. W-1 roll out, warp W + w is
    the cost partner of warp w (same lane = same candidate), warp 2 W loads (the operands are read as a
    broadcast by the lanes of a group).  Candidate trajectories go to the scratch columns (item_slot0 + instance of
    the round) * 16 + candidate; the group then copies its winner into the instance's other buffer.  in_base /
    out_base: ring stages already passed (the rings' mbarriers keep their phase across calls). */
template<class M, int W = kFanWarps>
__device__ __forceinline__ void fanoutRound(const M & model_in_constant_bank,
                                            const Workspace<typename M::Scalar> & ws,
                                            const SolverParams<typename M::Scalar> & prm,
                                            const FwdFanout<typename M::Scalar> & fan,
                                            const FanSmem<M, W> & sm,
                                            int iter,
                                            const int * list,
                                            int count,
                                            int slot0,
                                            size_t item_slot0,
                                            int warp,
                                            int lane,
                                            unsigned in_base,
                                            unsigned out_base)
{
  using S = typename M::Scalar;
  constexpr int NX = M::NX, NU = M::NU;
  using O = typename SplitLayout<M>::O;
  constexpr int GA = kFanLanes;
  constexpr int IPW = FanSmem<M, W>::IPW;
  constexpr int ROWS = FanSmem<M, W>::ROWS;
  constexpr unsigned kFull = 0xffffffffu;
  constexpr unsigned kGroupMask = (1u << GA) - 1u;
  const size_t Bp = ws.Bp;
  const int N = prm.N;

  if(warp == 2 * W)
  {
    // loader: lane l streams in-ring rows l, l + 32, ... of every step; row = (instance of the round) * O::SIZE + operand
    constexpr int RPL = (ROWS + 31) / 32;
    const S * row_ptr[RPL];
    long long row_stride[RPL];
#pragma unroll
    for(int q = 0; q < RPL; q++)
    {
      const int row = q * 32 + lane;
      const int inst = (row < ROWS) ? row / O::SIZE : 0;
      const int e = (row < ROWS) ? row % O::SIZE : 0;
      const int slot = slot0 + inst;
      const int b = list[slot < count ? slot : slot0]; // surplus slots repeat a valid instance
      const int sel = ws.sel[b];
      if(e < O::U)
      {
        row_ptr[q] = ws.x[sel] + (size_t)(e - O::X) * Bp + b;
        row_stride[q] = (long long)NX * (long long)Bp;
      }
      else if(e < O::KFF)
      {
        row_ptr[q] = ws.u[sel] + (size_t)(e - O::U) * Bp + b;
        row_stride[q] = (long long)NU * (long long)Bp;
      }
      else if(e < O::KFB)
      {
        row_ptr[q] = ws.kff + (size_t)(e - O::KFF) * Bp + b;
        row_stride[q] = (long long)NU * (long long)Bp;
      }
      else
      {
        row_ptr[q] = ws.kfb + (size_t)(e - O::KFB) * Bp + b;
        row_stride[q] = (long long)(NU * NX) * (long long)Bp;
      }
    }
    const int n_fills = (int)splitInStages(N);
    for(int f = 0; f < n_fills; f++)
    {
      const unsigned fg = in_base + (unsigned)f;
      const unsigned st = fg % kSplitIn;
      if(fg >= (unsigned)kSplitIn) mbarWait(&sm.in_empty[st], ((fg / kSplitIn) - 1u) & 1u);
#pragma unroll
      for(int sq = 0; sq < kSPS; sq++)
      {
        const bool in_range = f * kSPS + sq < N;
#pragma unroll
        for(int q = 0; q < RPL; q++)
        {
          const int row = q * 32 + lane;
          if(row < ROWS && in_range)
          {
            if constexpr(sizeof(S) == 8)
              cpAsync8(sm.in_ring + ((size_t)st * kSPS + sq) * ROWS + row, row_ptr[q]);
            else
              cpAsync4(sm.in_ring + ((size_t)st * kSPS + sq) * ROWS + row, row_ptr[q]);
          }
          row_ptr[q] += row_stride[q];
        }
      }
      cpAsyncArriveOn(&sm.in_full[st]);
    }
    return;
  }
  if(warp > 2 * W) return;

  using LM = typename LatencyOf<M>::type;
  const LM model(model_in_constant_bank);
  const int pair = warp % W;
  const int g = lane / GA;
  const int a = lane % GA;
  const int inst = pair * IPW + g; // instance of the round
  const int slot = slot0 + inst;
  const bool valid = slot < count;
  const int b = list[valid ? slot : slot0];
  const int sel = ws.sel[b];
  const int rem = prm.n_alpha - 1;
  const bool work = valid && (a < rem);
  const S my_alpha = prm.alpha_list[work ? (1 + a) : 0];
  S * out_col = sm.outCol(pair, lane);

  if(warp < W)
  {
    Matrix<S, NX, 1> x;
#pragma unroll
    for(int d = 0; d < NX; d++) x[d] = ws.x[sel][(size_t)d * Bp + b];
    splitRollout<LM, ROWS, 1>(model, prm.t0, N, my_alpha, x, sm.in_ring, inst * O::SIZE, sm.in_full, sm.in_empty, out_col,
                             sm.outFull(pair), sm.outEmpty(pair), in_base, out_base);
    return;
  }

  // ------------------------------------------------------------------ cost warps
  // Scratch of this listed slot: ONE contiguous block [row][candidate] per part (x, u, cost), so that the winner's rows
  // are 128 bytes apart when the group copies them out -- with the batch-innermost layout of the main buffers every row
  // of a candidate sat 16 * Bp * 8 bytes from the next one (a different 2 MB page per few rows: the copy was 16 % of the
  // kernel's samples, one stalled store).
  const size_t scratch_slot = item_slot0 + (size_t)inst;
  const size_t rows_x_all = (size_t)(N + 1) * NX, rows_u_all = (size_t)N * NU, rows_c_all = (size_t)(N + 1);
  const FwdDest<S> dst{fan.sx + scratch_slot * rows_x_all * GA, fan.su + scratch_slot * rows_u_all * GA,
                       fan.sc + scratch_slot * rows_c_all * GA,
                       (size_t)GA, (size_t)a};
  const S my_cost = splitCost<LM>(model, prm.t0, N, out_col, sm.outFull(pair), sm.outEmpty(pair), work, dst, out_base);

  const S cost_cur = ws.cost_sum[b];
  S my_actual = S(0), my_expected = S(0), my_ratio = S(0);
  bool ok = false;
  if(work)
    ok = lineSearchTest<S>(prm, cost_cur, my_cost, my_alpha, ws.dV[b], ws.dV[(size_t)ws.Bp + b], my_actual, my_expected,
                           my_ratio);
  const unsigned ok_ballot = __ballot_sync(kFull, ok);
  const unsigned gm = (ok_ballot >> (g * GA)) & kGroupMask;
  const int pick = (gm != 0) ? (__ffs(gm) - 1) : (rem - 1); // first success, else the last candidate tried
  const int src_lane = g * GA + pick;
  const S r_actual = __shfl_sync(kFull, my_actual, src_lane);
  const S r_expected = __shfl_sync(kFull, my_expected, src_lane);
  const S r_ratio = __shfl_sync(kFull, my_ratio, src_lane);
  const S r_cost = __shfl_sync(kFull, my_cost, src_lane);
  const S r_alpha = __shfl_sync(kFull, my_alpha, src_lane);
  const bool success = valid && (gm != 0);
  if(valid && a == 0)
    lineSearchFinish<S>(ws, prm, b, iter, sel, success, r_alpha, r_actual, r_expected, r_ratio, cost_cur, r_cost,
                        success ? (2 + pick) : prm.n_alpha);
  // Phase 3, fused: the group copies its winner's scratch trajectory into the instance's other buffer (which
  // lineSearchFinish has just made the current one).  The winner lane's global stores are ordered before the
  // group's loads by the warp barrier.
  __syncwarp();
  if(success)
  {
    const int rows_x = (N + 1) * NX, rows_u = N * NU, rows_c = N + 1;
    S * dx = ws.x[sel ^ 1];
    S * du = ws.u[sel ^ 1];
    S * dc = ws.cost[sel ^ 1];
    const int rows = rows_x + rows_u + rows_c;
    auto src = [&](int r) -> const S * {
      return (r < rows_x) ? dst.x + (size_t)r * GA + pick
                          : (r < rows_x + rows_u) ? dst.u + (size_t)(r - rows_x) * GA + pick
                                                  : dst.c + (size_t)(r - rows_x - rows_u) * GA + pick;
    };
    auto dstp = [&](int r) -> S * {
      return (r < rows_x) ? dx + (size_t)r * Bp + b
                          : (r < rows_x + rows_u) ? du + (size_t)(r - rows_x) * Bp + b
                                                  : dc + (size_t)(r - rows_x - rows_u) * Bp + b;
    };
    // One round trip for a 100-step cart-pole trajectory (605 rows over 16 lanes).  In-kernel time stamps put this
    // copy at 6-8 us of the kernel's 38: it is bound by the LSU's sector rate -- every 8-byte element, read or written,
    // is alone in its 32-byte sector (4840 + 4840 sectors per CTA) -- not by instructions or latency: three rounds of
    // cheaply addressed loads (running pointers) were no faster (39.4 us) than this single round.
    constexpr int kInFlight = 40;
    for(int r0 = a; r0 < rows; r0 += GA * kInFlight)
    {
      S v[kInFlight];
#pragma unroll
      for(int q = 0; q < kInFlight; q++)
      {
        const int r = r0 + q * GA;
        if(r < rows) v[q] = *src(r);
      }
#pragma unroll
      for(int q = 0; q < kInFlight; q++)
      {
        const int r = r0 + q * GA;
        if(r < rows) *dstp(r) = v[q];
      }
    }
  }
}

/** Phase 2 as a kernel of its own: CTA c serves the slots c * IPC .. of the global work list. */
// 167 registers at W = 1 (the winner copy keeps 40 loads in flight): four CTAs per SM.  Capping the registers for more
// resident CTAs (6 / 10 per SM) spills and costs more than it gains: 43.0 / 54.1 us vs 33.8 us per launch.
template<class M, int W>
__global__ void __launch_bounds__((2 * W + 1) * 32)
    forward_fanout_split_kernel(const __grid_constant__ M model_in_constant_bank,
                                const __grid_constant__ Workspace<typename M::Scalar> ws,
                                const __grid_constant__ SolverParams<typename M::Scalar> prm,
                                const __grid_constant__ FwdFanout<typename M::Scalar> fan,
                                int iter)
{
  pdlPrologue();
  const int count = *fan.count;
  const int cta_slot0 = blockIdx.x * FanSmem<M, W>::IPC;
  if(cta_slot0 >= count) return; // CTA-uniform
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const FanSmem<M, W> sm(smem_raw);
  if(threadIdx.x == 0)
  {
    sm.initBarriers();
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  fanoutRound<M, W>(model_in_constant_bank, ws, prm, fan, sm, iter, fan.list, count, cta_slot0, (size_t)cta_slot0,
                 threadIdx.x >> 5, threadIdx.x & 31, 0u, 0u);
}
} // namespace ddp
} // namespace nmpc_b200
