/* nmpc_b200 -- K3 for small batches, in three phases (procOnce() Steps 3-4, DDPSolver.hpp:234-339).
 *
 * At 4096 instances the line search is bound by the latency of one rollout (100 dependent steps of
 * sincos -> reciprocal -> state update), not by throughput, so the number of SEQUENTIAL rollouts per
 * iteration is what matters:
 *
 *   phase 1  forward_first_kernel    every instance rolls out alpha_list[0] (one thread each, operands fed
 *                                    by a 4-deep cp.async ring) and stores the candidate; ~90 % of all line
 *                                    searches end here.  Instances whose first candidate fails are
 *                                    appended to a work list.
 *   phase 2  forward_fanout_kernel   ALL remaining candidates of ALL listed instances are rolled out
 *                                    concurrently, one lane per (instance, candidate), 16 lanes per
 *                                    instance sharing one operand ring; each lane writes its trajectory
 *                                    to scratch.  First success in list order wins -- identical to the
 *                                    reference's sequential backtracking because forwardPass(alpha) is a
 *                                    pure function of alpha.
 *   phase 3  (fused into phase 2)    each winner's scratch trajectory is copied into the (new) current buffer
 *                                    by the 16 lanes of its group.
 *
 * A late M-fixed iteration therefore costs two rollout latencies instead of up to eleven.
 */
#pragma once

#include "ddp_kernels.cuh"

namespace nmpc_b200
{
namespace ddp
{
constexpr int kFanLanes = 16; //!< lanes (candidate slots) per listed instance in phase 2

template<class S>
struct FwdFanout
{
  int * count; //!< [1] number of listed instances (reset by K1 / K0)
  int * list; //!< [Bp] instance index per slot
  int * commit_item; //!< [Bp] per slot: scratch column of the winning candidate, or -1
  S * sx; //!< scratch candidates [N+1][NX][items]
  S * su; //!< [N][NU][items]
  S * sc; //!< [N+1][items]
  size_t items; //!< scratch columns = Bp * kFanLanes
};

/** procOnce() Step 4 (DDPSolver.hpp:280-339) for one instance, after its line search is decided. */
template<class S>
__device__ __forceinline__ void lineSearchFinish(const Workspace<S> & ws,
                                                 const SolverParams<S> & prm,
                                                 int b,
                                                 int iter,
                                                 int sel,
                                                 bool success,
                                                 S alpha,
                                                 S actual,
                                                 S expected,
                                                 S ratio,
                                                 S cost_cur,
                                                 S cost_new,
                                                 int tried)
{
  S lambda = ws.lambda[b];
  S dlambda = ws.dlambda[b];
  const S k_rel_norm = ws.trace[((size_t)iter * kTraceFields + 5) * ws.Bp + b];
  int retval = 0;
  S cost_out = cost_cur;
  if(success)
  {
    ws.sel[b] = sel ^ 1;
    ws.cost_sum[b] = cost_new;
    cost_out = cost_new;
    if(actual < prm.cost_update_thre) retval = 1;
    dlambda = fmin(dlambda / prm.lambda_factor, S(1) / prm.lambda_factor);
    if(lambda >= prm.lambda_min)
      lambda *= dlambda;
    else
      lambda = S(0);
  }
  else
  {
    dlambda = fmax(dlambda * prm.lambda_factor, prm.lambda_factor);
    lambda = fmax(lambda * dlambda, prm.lambda_min);
    if(lambda > prm.lambda_max) retval = -1;
  }
  ws.lambda[b] = lambda;
  ws.dlambda[b] = dlambda;
  ws.n_fwd[b] += tried;
  ws.iters[b] = iter;
  if(retval != 0) ws.status[b] = retval;
  writeTrace<S>(ws, b, iter, S(iter), cost_out, lambda, dlambda, alpha, k_rel_norm, actual, expected, ratio);
}

constexpr int kFirstDepth = 4; //!< ring stages between the loader warp and the compute warp of phase 1

/** forwardPass(alpha) (DDPSolver.hpp:537-560) for one instance per lane with the operands {x_i, u_i, k_i, K_i} of the
    current trajectory delivered by the CTA's loader warp.  Same arithmetic, in the same order, as forwardRolloutRing /
    forwardRollout => the same costs bit for bit. */
template<class M>
__device__ __forceinline__ typename M::Scalar forwardRolloutFed(const M & model_in_constant_bank,
                                                                const Workspace<typename M::Scalar> & ws,
                                                                const SolverParams<typename M::Scalar> & prm,
                                                                const typename M::Scalar * __restrict__ ring,
                                                                unsigned long long * full,
                                                                unsigned long long * empty,
                                                                int lane,
                                                                int b,
                                                                int sel,
                                                                typename M::Scalar alpha,
                                                                bool work)
{
  using S = typename M::Scalar;
  constexpr int NX = M::NX, NU = M::NU;
  using O = FwdOperands<NX, NU>;
  const M model = model_in_constant_bank;
  const S t0 = prm.t0;
  const size_t Bp = ws.Bp;
  const int N = prm.N;
  S * __restrict__ xn = ws.x[sel ^ 1];
  S * __restrict__ un = ws.u[sel ^ 1];
  S * __restrict__ cn = ws.cost[sel ^ 1];

  Matrix<S, NX, 1> x;
#pragma unroll
  for(int d = 0; d < NX; d++) x[d] = ws.x[sel][(size_t)d * Bp + b];
  if(work)
  {
#pragma unroll
    for(int d = 0; d < NX; d++) xn[(size_t)d * Bp + b] = x[d]; // candidate x_list[0] (:540)
  }
  S * xs_ptr = xn + (size_t)NX * Bp + b;
  S * us_ptr = un + b;
  S * cs_ptr = cn + b;

  S csum = S(0);
  S fi = S(0); // == S(i) exactly: spares the int -> floating-point conversion of `i * dt` on every step
  for(int i = 0; i < N; i++, fi += S(1))
  {
    const int st = i % kFirstDepth;
    mbarWait(&full[st], (unsigned)(i / kFirstDepth) & 1u);
    S xr[NX], ur[NU], kr[NU], Kr[NU * NX];
    const S * op = ring + (size_t)st * O::SIZE * kTile + lane;
#pragma unroll
    for(int d = 0; d < NX; d++) xr[d] = op[(size_t)(O::X + d) * kTile];
#pragma unroll
    for(int d = 0; d < NU; d++) ur[d] = op[(size_t)(O::U + d) * kTile];
#pragma unroll
    for(int d = 0; d < NU; d++) kr[d] = op[(size_t)(O::KFF + d) * kTile];
#pragma unroll
    for(int d = 0; d < NU * NX; d++) Kr[d] = op[(size_t)(O::KFB + d) * kTile];
    mbarArrive(&empty[st]); // the operands are in registers: the loader may refill the stage

    if(work)
    {
      Matrix<S, NU, 1> u;
#pragma unroll
      for(int c = 0; c < NU; c++)
      {
        S acc = S(0);
#pragma unroll
        for(int j = 0; j < NX; j++) acc += Kr[c + j * NU] * (x[j] - xr[j]);
        u[c] = (ur[c] + alpha * kr[c]) + acc; // u' = u + alpha k + K (x' - x)   (:545-546)
        us_ptr[(size_t)c * Bp] = u[c];
      }
      const S t = t0 + fi * model.dt();
      const S c = model.runningCost(t, x, u);
      x = model.stateEq(t, x, u);
#pragma unroll
      for(int d = 0; d < NX; d++) xs_ptr[(size_t)d * Bp] = x[d];
      *cs_ptr = c;
      csum += c;
    }
    xs_ptr += (size_t)NX * Bp;
    us_ptr += (size_t)NU * Bp;
    cs_ptr += Bp;
  }
  if(work)
  {
    const S c = model.terminalCost(t0 + N * model.dt(), x);
    cn[(size_t)N * Bp + b] = c;
    csum += c;
  }
  return csum;
}

/** Phase 1: alpha_list[0] for every running instance.  CTA = one 32-instance tile: warp 0 rolls out, warp 1 loads. */
template<class M>
__global__ void __launch_bounds__(64) forward_first_kernel(const __grid_constant__ M model,
                                                           const __grid_constant__ Workspace<typename M::Scalar> ws,
                                                           const __grid_constant__ SolverParams<typename M::Scalar> prm,
                                                           const __grid_constant__ FwdFanout<typename M::Scalar> fan,
                                                           int iter)
{
  pdlPrologue();
  using S = typename M::Scalar;
  constexpr int NX = M::NX, NU = M::NU;
  using O = FwdOperands<NX, NU>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  S * ring = reinterpret_cast<S *>(smem_raw); // [kFirstDepth][O::SIZE][32]
  unsigned long long * full =
      reinterpret_cast<unsigned long long *>(smem_raw + sizeof(S) * (size_t)kFirstDepth * O::SIZE * kTile);
  unsigned long long * empty = full + kFirstDepth;
  if(threadIdx.x == 0)
  {
    for(int st = 0; st < kFirstDepth; st++)
    {
      mbarInit(&full[st], 32);
      mbarInit(&empty[st], 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }

  const int bg = blockIdx.x * kTile + lane;
  const int b = (bg < ws.B) ? bg : (ws.B - 1);
  const bool active = (bg < ws.B) && (ws.status[b] == 0);
  // both warps see the same 32 verdicts: uniform exit; the barrier also publishes the mbarrier initialisation
  if(!__syncthreads_or(active)) return;
  const int sel = ws.sel[b];
  const size_t Bp = ws.Bp;

  if(warp == 1)
  {
    // loader: {x_i, u_i, k_i, K_i} of this lane's instance for steps 0 .. N-1
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
    loaderLoop<S, O::SIZE, kFirstDepth>(ring, full, empty, lane, prm.N, row_ptr, row_stride);
    return;
  }

  const bool work = active && prm.n_alpha > 0;
  const S alpha = prm.alpha_list[0];
  const S cost_new = forwardRolloutFed<M>(model, ws, prm, ring, full, empty, lane, b, sel, alpha, work);
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

constexpr int kFanWarps = 4; //!< compute warps of a phase-2 CTA (one more warp loads)
constexpr int kFanDepth = 4; //!< ring stages between its loader warp and its compute warps

/** Phase 2: candidates 1 .. n_alpha-1 of every listed instance at once.  CTA = 4 compute warps (16 lanes per listed
    instance, two instances per warp) + 1 loader warp that streams {x_i, u_i, k_i, K_i} of the CTA's eight instances into
    a shared-memory ring (per-lane cp.async, full / empty mbarriers); every lane of a group reads its instance's
    operands from the ring as a broadcast. */
template<class M>
__global__ void __launch_bounds__((kFanWarps + 1) * 32)
    forward_fanout_kernel(const __grid_constant__ M model_in_constant_bank,
                          const __grid_constant__ Workspace<typename M::Scalar> ws,
                          const __grid_constant__ SolverParams<typename M::Scalar> prm,
                          const __grid_constant__ FwdFanout<typename M::Scalar> fan,
                          int iter)
{
  pdlPrologue();
  using S = typename M::Scalar;
  constexpr int NX = M::NX, NU = M::NU;
  using O = FwdOperands<NX, NU>;
  constexpr int GA = kFanLanes;
  constexpr int IPW = 32 / GA; // listed instances per compute warp
  constexpr int IPC = kFanWarps * IPW; // ... per CTA
  constexpr int ROWS = IPC * O::SIZE; // ring rows of one step: [instance of the CTA][operand]
  constexpr unsigned kFull = 0xffffffffu;
  constexpr unsigned kGroupMask = (1u << GA) - 1u;
  const int count = *fan.count;
  const int cta_slot0 = blockIdx.x * IPC;
  if(cta_slot0 >= count) return; // CTA-uniform
  extern __shared__ __align__(16) unsigned char smem_raw[];
  S * ring = reinterpret_cast<S *>(smem_raw); // [kFanDepth][ROWS]
  unsigned long long * full = reinterpret_cast<unsigned long long *>(smem_raw + sizeof(S) * (size_t)kFanDepth * ROWS);
  unsigned long long * empty = full + kFanDepth;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if(threadIdx.x == 0)
  {
    for(int st = 0; st < kFanDepth; st++)
    {
      mbarInit(&full[st], 32);
      mbarInit(&empty[st], kFanWarps * 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  const size_t Bp = ws.Bp;
  const int N = prm.N;

  if(warp == kFanWarps)
  {
    // loader: lane l streams ring rows l, l + 32, ...; row = (instance of the CTA) * O::SIZE + operand
    constexpr int RPL = (ROWS + 31) / 32;
    const S * row_ptr[RPL];
    long long row_stride[RPL];
#pragma unroll
    for(int q = 0; q < RPL; q++)
    {
      const int row = q * 32 + lane;
      const int inst = (row < ROWS) ? row / O::SIZE : 0;
      const int e = (row < ROWS) ? row % O::SIZE : 0;
      const int slot = cta_slot0 + inst;
      const int b = fan.list[slot < count ? slot : cta_slot0]; // surplus slots repeat a valid instance
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
    for(int f = 0; f < N; f++)
    {
      const int st = f % kFanDepth;
      if(f >= kFanDepth) mbarWait(&empty[st], (unsigned)((f / kFanDepth) - 1) & 1u);
#pragma unroll
      for(int q = 0; q < RPL; q++)
      {
        const int row = q * 32 + lane;
        if(row < ROWS)
        {
          if constexpr(sizeof(S) == 8)
            cpAsync8(ring + (size_t)st * ROWS + row, row_ptr[q]);
          else
            cpAsync4(ring + (size_t)st * ROWS + row, row_ptr[q]);
        }
        row_ptr[q] += row_stride[q];
      }
      cpAsyncArriveOn(&full[st]);
    }
    return;
  }

  // ------------------------------------------------------------------ compute warps
  const M model = model_in_constant_bank;
  const S t0 = prm.t0;
  const int g = lane / GA;
  const int a = lane % GA;
  const int inst = warp * IPW + g; // instance of the CTA
  const int slot = cta_slot0 + inst;
  const bool valid = slot < count;
  const int b = fan.list[valid ? slot : cta_slot0];
  const int sel = ws.sel[b];
  const int rem = prm.n_alpha - 1;
  const bool work = valid && (a < rem);
  const S my_alpha = prm.alpha_list[work ? (1 + a) : 0];
  const size_t item = (size_t)slot * GA + a; // scratch column of this candidate
  const size_t Bd = fan.items;

  Matrix<S, NX, 1> x;
#pragma unroll
  for(int d = 0; d < NX; d++) x[d] = ws.x[sel][(size_t)d * Bp + b];
  if(work)
  {
#pragma unroll
    for(int d = 0; d < NX; d++) fan.sx[(size_t)d * Bd + item] = x[d];
  }
  S * xs_ptr = fan.sx + (size_t)NX * Bd + item;
  S * us_ptr = fan.su + item;
  S * cs_ptr = fan.sc + item;
  S my_cost = S(0);
  S fi = S(0); // == S(i) exactly
  for(int i = 0; i < N; i++, fi += S(1))
  {
    const int st = i % kFanDepth;
    mbarWait(&full[st], (unsigned)(i / kFanDepth) & 1u);
    S xr[NX], ur[NU], kr[NU], Kr[NU * NX];
    const S * op = ring + (size_t)st * ROWS + (size_t)inst * O::SIZE; // same address in the whole group: broadcast
#pragma unroll
    for(int d = 0; d < NX; d++) xr[d] = op[O::X + d];
#pragma unroll
    for(int d = 0; d < NU; d++) ur[d] = op[O::U + d];
#pragma unroll
    for(int d = 0; d < NU; d++) kr[d] = op[O::KFF + d];
#pragma unroll
    for(int d = 0; d < NU * NX; d++) Kr[d] = op[O::KFB + d];
    mbarArrive(&empty[st]);

    if(work)
    {
      Matrix<S, NU, 1> u;
#pragma unroll
      for(int c = 0; c < NU; c++)
      {
        S acc = S(0);
#pragma unroll
        for(int j = 0; j < NX; j++) acc += Kr[c + j * NU] * (x[j] - xr[j]);
        u[c] = (ur[c] + my_alpha * kr[c]) + acc; // u' = u + alpha k + K (x' - x)   (:545-546)
        us_ptr[(size_t)c * Bd] = u[c];
      }
      const S t = t0 + fi * model.dt();
      const S c = model.runningCost(t, x, u);
      x = model.stateEq(t, x, u);
#pragma unroll
      for(int d = 0; d < NX; d++) xs_ptr[(size_t)d * Bd] = x[d];
      *cs_ptr = c;
      my_cost += c;
    }
    xs_ptr += (size_t)NX * Bd;
    us_ptr += (size_t)NU * Bd;
    cs_ptr += Bd;
  }
  if(work)
  {
    const S c = model.terminalCost(t0 + N * model.dt(), x);
    fan.sc[(size_t)N * Bd + item] = c;
    my_cost += c;
  }

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
  {
    fan.commit_item[slot] = success ? (int)((size_t)slot * GA + pick) : -1;
    lineSearchFinish<S>(ws, prm, b, iter, sel, success, r_alpha, r_actual, r_expected, r_ratio, cost_cur, r_cost,
                        success ? (2 + pick) : prm.n_alpha);
  }
  // Phase 3, fused: the group copies its winner's scratch trajectory into the instance's other buffer (which
  // lineSearchFinish has just made the current one).  The winner lane's global stores are ordered before the
  // group's loads by the warp barrier.
  __syncwarp();
  if(success)
  {
    const size_t win = (size_t)slot * GA + pick;
    const int rows_x = (N + 1) * NX, rows_u = N * NU, rows_c = N + 1;
    S * __restrict__ dx = ws.x[sel ^ 1];
    S * __restrict__ du = ws.u[sel ^ 1];
    S * __restrict__ dc = ws.cost[sel ^ 1];
    // one flat row index over (x | u | cost); 8 independent loads in flight per lane
    const int rows = rows_x + rows_u + rows_c;
    auto src = [&](int r) -> const S * {
      return (r < rows_x) ? fan.sx + (size_t)r * fan.items + win
                          : (r < rows_x + rows_u) ? fan.su + (size_t)(r - rows_x) * fan.items + win
                                                  : fan.sc + (size_t)(r - rows_x - rows_u) * fan.items + win;
    };
    auto dstp = [&](int r) -> S * {
      return (r < rows_x) ? dx + (size_t)r * Bp + b
                          : (r < rows_x + rows_u) ? du + (size_t)(r - rows_x) * Bp + b
                                                  : dc + (size_t)(r - rows_x - rows_u) * Bp + b;
    };
    for(int r0 = a; r0 < rows; r0 += GA * 8)
    {
      S v[8];
#pragma unroll
      for(int q = 0; q < 8; q++)
      {
        const int r = r0 + q * GA;
        if(r < rows) v[q] = *src(r);
      }
#pragma unroll
      for(int q = 0; q < 8; q++)
      {
        const int r = r0 + q * GA;
        if(r < rows) *dstp(r) = v[q];
      }
    }
  }
}

} // namespace ddp
} // namespace nmpc_b200
