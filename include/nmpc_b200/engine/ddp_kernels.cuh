/* nmpc_b200 -- DDP/iLQR stage kernels for sm_100a.
 *
 * One thread owns one problem instance; the batch index is the fastest-varying index of every
 * device array ("batch innermost"), so the 32 lanes of a warp read/write 32 consecutive scalars
 * (256 B for fp64) per access: every global access is fully coalesced and every per-instance
 * matrix lives in registers after unrolling (n_x, n_u are compile-time constants of the functor).
 *
 * Stages (reference: isri-aist/NMPC nmpc_ddp/include/nmpc_ddp/DDPSolver.hpp):
 *   K0 rollout_init_kernel   solve() initial rollout                         :83-104
 *   K1 linearize_kernel      procOnce() Step 1, parallel over (instance, step) :157-185
 *   K2 backward_kernel       procOnce() Step 2 + backwardPass() + termination  :188-231, :343-534
 *   K3 forward_kernel        procOnce() Step 3/4 + forwardPass()               :234-339, :537-560
 * Variants in their own headers: ddp_backward_fused.cuh (K1 + K2 in one kernel, the default for n_x < 8),
 * ddp_backward_quad.cuh / ddp_backward_coop.cuh (column-split K2 for n_x >= 8), ddp_forward_phased.cuh (K3 for
 * latency-bound batches), ddp_mpc.cuh (tick-to-tick kernel of the MPC loop).
 *
 * Device layout (S = scalar type, Bp = padded batch):
 *   x[2]    [N+1][NX][Bp]   current / candidate trajectories; sel[b] says which one is current
 *   u[2]    [N][NU][Bp]
 *   cost[2] [N+1][Bp]
 *   deriv   [N][tile][BLK][32]  BLK = {Fx, Fu, Lx, Lu, Lxx, Luu, Lxu} column-major; three-kernel pipeline only
 *   vterm   [NX+NX*NX][Bp]  terminal Vx, Vxx; three-kernel pipeline only
 *   kff     [N][NU][Bp], kfb [N][NU*NX][Bp]
 *   trace   [max_iter+1][9][Bp]
 *   per-instance scalars: lambda, dlambda, cost_sum, dV[2], status, sel, iters, n_fwd, n_bwd
 */
#pragma once

#include <cuda_runtime.h>

#include <type_traits>
#include <utility>

#include <nmpc_b200/matrix.h>

#include "boxqp.cuh"

namespace nmpc_b200
{
namespace ddp
{
constexpr int kTraceFields = 9;
constexpr int kMaxAlpha = 16;

/** Solver constants, passed by value to every kernel (mirror of DDPSolver::Configuration). */
template<class S>
struct SolverParams
{
  int N;
  int max_iter;
  int reg_type;
  int with_input_constraint;
  int n_alpha;
  S t0;
  S initial_lambda;
  S initial_dlambda;
  S lambda_factor;
  S lambda_min;
  S lambda_max;
  S k_rel_norm_thre;
  S lambda_thre;
  S cost_update_ratio_thre;
  S cost_update_thre;
  S alpha_list[kMaxAlpha];
};

template<class S>
struct Workspace
{
  int B; //!< live instances
  int Bp; //!< padded batch (allocation stride)
  S * x[2];
  S * u[2];
  S * cost[2];
  S * deriv;
  S * vterm;
  S * kff;
  S * kfb;
  S * trace;
  S * lambda;
  S * dlambda;
  S * cost_sum;
  S * dV;
  S * u_lo; //!< [N][NU] input limits of every horizon step (with_input_constraint; input_limits_func_(t_i))
  S * u_hi;
  int * status;
  int * sel;
  int * iters;
  int * n_fwd;
  int * n_bwd;
  int * fan_count; //!< [1] work-list length of the phased line search (reset by K0 / K1 / the fused K2)
};

/** The functor type a latency-bound kernel evaluates: `M::LatencyVariant` when the functor offers one (the same problem
    and the same values to within rounding, written for instruction latency rather than instruction count -- models/cartpole.h), else M.
    The kernels that run one rollout per lane on a few warps per SM (ddp_forward_split.cuh, ddp_backward_lanes.cuh)
    convert the functor they were launched with at entry; the throughput-bound kernels keep M. */
template<class M, class = void>
struct LatencyOf
{
  using type = M;
};
template<class M>
struct LatencyOf<M, std::void_t<typename M::LatencyVariant>>
{
  using type = typename M::LatencyVariant;
};

/** Does the functor split stateEq into statePrePair / stateEqPre (models/cartpole.h), so that a rollout can take the
    expensive functions of the state out of its step-to-step dependency chain? */
template<class M, class = void>
struct HasStatePre : std::false_type
{
};
template<class M>
struct HasStatePre<M, std::void_t<typename M::StatePre>> : std::true_type
{
};
/** What a rollout that advances two steps at a time carries besides the state: prepare(x) before the pair, then
    advance(.., 0) and advance(.., 1). */
template<class M, bool = HasStatePre<M>::value>
struct RolloutCarry
{
  __device__ __forceinline__ void prepare(const M &, typename M::Scalar, const typename M::StateDimVector &) {}
  __device__ __forceinline__ typename M::StateDimVector advance(const M & model,
                                                                typename M::Scalar t,
                                                                const typename M::StateDimVector & x,
                                                                const typename M::InputDimVector & u,
                                                                int)
  {
    return model.stateEq(t, x, u);
  }
};
template<class M>
struct RolloutCarry<M, true>
{
  typename M::StatePre pre[2];
  __device__ __forceinline__ void prepare(const M & model, typename M::Scalar t, const typename M::StateDimVector & x)
  {
    model.statePrePair(t, x, pre[0], pre[1]);
  }
  __device__ __forceinline__ typename M::StateDimVector advance(const M & model,
                                                                typename M::Scalar t,
                                                                const typename M::StateDimVector & x,
                                                                const typename M::InputDimVector & u,
                                                                int q)
  {
    return model.stateEqPre(t, x, u, pre[q]);
  }
};

/** Does the functor have a time-varying input dimension, `int inputDim(t)` <= NU (DDPProblem<StateDim, Eigen::Dynamic>,
    DDPProblem.h:61-85)?  Inputs a >= inputDim(t) are padding: kept at zero by K0 and decoupled by K1. */
template<class M, class = void>
struct HasInputDim : std::false_type
{
};
template<class M>
struct HasInputDim<M, std::void_t<decltype(std::declval<const M &>().inputDim(std::declval<typename M::Scalar>()))>>
: std::true_type
{
};

/** BoxQP warm start of step i (DDPSolver.hpp:452-467): k_list_[i + 1] is used only when it has the input dimension of
    step i; with compile-time sizes that is the case unless inputDim(t) changes between the two steps. */
template<class M>
__device__ __forceinline__ bool warmStartFromNextStep(const M & model, typename M::Scalar t0, int i, int N)
{
  if(i == N - 1) return false;
  if constexpr(HasInputDim<M>::value)
  {
    const typename M::Scalar dt = model.dt();
    return model.inputDim(t0 + i * dt) == model.inputDim(t0 + (i + 1) * dt);
  }
  return true;
}

template<int NX, int NU>
struct BlockLayout
{
  static constexpr int FX = 0;
  static constexpr int FU = FX + NX * NX;
  static constexpr int LX = FU + NX * NU;
  static constexpr int LU = LX + NX;
  static constexpr int LXX = LU + NU;
  static constexpr int LUU = LXX + NX * NX;
  static constexpr int LXU = LUU + NU * NU;
  static constexpr int SIZE = LXU + NX * NU;
};

/** Programmatic dependent launch (sm_90+): every stage kernel starts with pdlPrologue().  It lets the NEXT
    kernel of the stream be launched and made resident right away (launch_dependents) and then waits until the
    PREVIOUS kernel has completed and flushed its results (griddepcontrol.wait), so the launch latency between
    the ~45 dependent kernels of one solve overlaps with the tail of the predecessor.  Without the launch
    attribute both instructions are no-ops. */
__device__ __forceinline__ void pdlPrologue()
{
  asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
  asm volatile("griddepcontrol.wait;\n" ::: "memory");
}

template<class S>
__device__ __forceinline__ S ldStream(const S * p)
{
  return __ldg(p);
}

/** 8-byte asynchronous global->shared copy (LDGSTS); completion is tracked per thread. */
__device__ __forceinline__ void cpAsync8(void * smem_dst, const void * gmem_src)
{
  const unsigned sa = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cpAsync4(void * smem_dst, const void * gmem_src)
{
  const unsigned sa = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cpAsyncCommit()
{
  asm volatile("cp.async.commit_group;\n" ::: "memory");
}
template<int N>
__device__ __forceinline__ void cpAsyncWait()
{
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

/** Each thread stages its own column of one step's derivative block: smem[e][tid] <- deriv[step][e][b]. */
template<class S, int SIZE>
__device__ __forceinline__ void stageBlock(S * stage, const S * __restrict__ blk, size_t Bp, int tpb)
{
#pragma unroll
  for(int e = 0; e < SIZE; e++)
  {
    if constexpr(sizeof(S) == 8)
      cpAsync8(stage + (size_t)e * tpb, blk + (size_t)e * Bp);
    else
      cpAsync4(stage + (size_t)e * tpb, blk + (size_t)e * Bp);
  }
}

/* The derivative blocks are tiled by warp: [step][tile of 32 instances][BLK][32].  One warp's block of one
   step is a single contiguous BLK * 32 * sizeof(S) chunk, so K2 fetches it with ONE bulk (TMA) copy. */
constexpr int kTile = 32;

template<int SIZE>
__device__ __forceinline__ size_t derivTileOffset(int step, int b, int Bp)
{
  return ((size_t)step * (Bp / kTile) + (size_t)(b / kTile)) * SIZE * kTile + (size_t)(b % kTile);
}

__device__ __forceinline__ void mbarInit(unsigned long long * bar, unsigned count)
{
  const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarExpectTx(unsigned long long * bar, unsigned bytes)
{
  const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(unsigned long long * bar, unsigned parity)
{
  const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  asm volatile("{\n"
               ".reg .pred p;\n"
               "WAIT_LOOP:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
               "@p bra WAIT_DONE;\n"
               "bra WAIT_LOOP;\n"
               "WAIT_DONE:\n"
               "}\n" ::"r"(a),
               "r"(parity)
               : "memory");
}
/** One-dimensional bulk copy global -> shared through the TMA engine; completion lands on `bar`. */
__device__ __forceinline__ void bulkCopyG2S(void * smem_dst, const void * gmem_src, unsigned bytes, unsigned long long * bar)
{
  const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  const unsigned b = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(d),
               "l"(gmem_src), "r"(bytes), "r"(b)
               : "memory");
}

template<class S>
__device__ __forceinline__ void writeTrace(const Workspace<S> & ws,
                                           int b,
                                           int row,
                                           S iter,
                                           S cost,
                                           S lambda,
                                           S dlambda,
                                           S alpha,
                                           S k_rel_norm,
                                           S actual,
                                           S expected,
                                           S ratio)
{
  S * tr = ws.trace + (size_t)row * kTraceFields * ws.Bp + b;
  tr[0 * (size_t)ws.Bp] = iter;
  tr[1 * (size_t)ws.Bp] = cost;
  tr[2 * (size_t)ws.Bp] = lambda;
  tr[3 * (size_t)ws.Bp] = dlambda;
  tr[4 * (size_t)ws.Bp] = alpha;
  tr[5 * (size_t)ws.Bp] = k_rel_norm;
  tr[6 * (size_t)ws.Bp] = actual;
  tr[7 * (size_t)ws.Bp] = expected;
  tr[8 * (size_t)ws.Bp] = ratio;
}

/* ------------------------------------------------------------------------------------ K0 ---- */
/** solve(): reset lambda/dlambda, initial rollout and iter-0 trace entry (DDPSolver.hpp:36-38, :83-104).
    x[0][0] and u[0] were filled by the layout kernel. */
template<class M>
__global__ void rollout_init_kernel(const __grid_constant__ M model,
                                    const __grid_constant__ Workspace<typename M::Scalar> ws,
                                    const __grid_constant__ SolverParams<typename M::Scalar> prm)
{
  pdlPrologue();
  using S = typename M::Scalar;
  constexpr int NX = M::NX, NU = M::NU;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if(b == 0) *ws.fan_count = 0;
  if(b >= ws.B) return;
  const size_t Bp = ws.Bp;
  const int N = prm.N;

  Matrix<S, NX, 1> x;
#pragma unroll
  for(int d = 0; d < NX; d++) x[d] = ws.x[0][(size_t)d * Bp + b];

  S csum = S(0);
  // initial_u_list is streamed in chunks of kChunk steps, one chunk ahead of its use: the rollout is a dependent
  // chain, and a load issued in the step that needs it would put one HBM latency on every step
  constexpr int kChunk = 8;
  S ubuf[2][kChunk][NU];
  auto loadChunk = [&](int slot, int i0) {
#pragma unroll
    for(int q = 0; q < kChunk; q++)
    {
      const int i = (i0 + q < N) ? i0 + q : N - 1;
#pragma unroll
      for(int d = 0; d < NU; d++) ubuf[slot][q][d] = ws.u[0][((size_t)i * NU + d) * Bp + b];
    }
  };
  loadChunk(0, 0);
  for(int i0 = 0; i0 < N; i0 += 2 * kChunk)
  {
#pragma unroll
    for(int half = 0; half < 2; half++)
    {
      loadChunk(half ^ 1, i0 + (half + 1) * kChunk);
#pragma unroll
      for(int q = 0; q < kChunk; q++)
      {
        const int i = i0 + half * kChunk + q;
        if(i < N)
        {
    Matrix<S, NU, 1> u;
#pragma unroll
    for(int d = 0; d < NU; d++) u[d] = ubuf[half][q][d];
    const S t = prm.t0 + i * model.dt();
    if constexpr(HasInputDim<M>::value)
    {
      // padding entries of initial_u_list are forced to zero; the gains keep them there
      const int nu_act = model.inputDim(t);
#pragma unroll
      for(int d = 0; d < NU; d++)
      {
        if(d >= nu_act)
        {
          u[d] = S(0);
          ws.u[0][((size_t)i * NU + d) * Bp + b] = S(0);
        }
      }
    }
    const S c = model.runningCost(t, x, u);
    x = model.stateEq(t, x, u);
#pragma unroll
    for(int d = 0; d < NX; d++) ws.x[0][((size_t)(i + 1) * NX + d) * Bp + b] = x[d];
    ws.cost[0][(size_t)i * Bp + b] = c;
    csum += c;
        }
      }
    }
  }
  {
    const S t = prm.t0 + N * model.dt();
    const S c = model.terminalCost(t, x);
    ws.cost[0][(size_t)N * Bp + b] = c;
    csum += c;
  }

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

/** procOnce() Step 1 for one (instance, step) (DDPSolver.hpp:164-176): dynamics and running-cost derivatives, with the
    padding inputs of a time-varying input dimension decoupled. */
template<class M>
__device__ __forceinline__ void linearizeStep(const M & model,
                                              typename M::Scalar t,
                                              const Matrix<typename M::Scalar, M::NX, 1> & x,
                                              const Matrix<typename M::Scalar, M::NU, 1> & u,
                                              Matrix<typename M::Scalar, M::NX, M::NX> & Fx,
                                              Matrix<typename M::Scalar, M::NX, M::NU> & Fu,
                                              Matrix<typename M::Scalar, M::NX, 1> & Lx,
                                              Matrix<typename M::Scalar, M::NU, 1> & Lu,
                                              Matrix<typename M::Scalar, M::NX, M::NX> & Lxx,
                                              Matrix<typename M::Scalar, M::NU, M::NU> & Luu,
                                              Matrix<typename M::Scalar, M::NX, M::NU> & Lxu)
{
  using S = typename M::Scalar;
  constexpr int NX = M::NX, NU = M::NU;
  model.calcStateEqDeriv(t, x, u, Fx, Fu);
  model.calcRunningCostDeriv(t, x, u, Lx, Lu, Lxx, Luu, Lxu);
  if constexpr(HasInputDim<M>::value)
  {
    // decouple the padding inputs: Fu(:,a) = 0, Lu(a) = 0, Lxu(:,a) = 0, Luu(a,:) = Luu(:,a) = e_a.  Quu becomes
    // block diagonal [Quu_active, 1], so k(a) = 0, K(a,:) = 0 and every active quantity equals the reference's
    // reduced-dimension result (steps with inputDim 0 reduce to Vx = Qx, Vxx = Qxx, DDPSolver.hpp:513-517)
    const int nu_act = model.inputDim(t);
#pragma unroll
    for(int a = 0; a < NU; a++)
    {
      if(a >= nu_act)
      {
#pragma unroll
        for(int r = 0; r < NX; r++)
        {
          Fu(r, a) = S(0);
          Lxu(r, a) = S(0);
        }
        Lu[a] = S(0);
#pragma unroll
        for(int c = 0; c < NU; c++)
        {
          Luu(a, c) = S(0);
          Luu(c, a) = S(0);
        }
        Luu(a, a) = S(1);
      }
    }
  }
}

/* ------------------------------------------------------------------------------------ K1 ---- */
/** procOnce() Step 1 (DDPSolver.hpp:157-185): thread (b, i) differentiates dynamics and cost at
    (x_i, u_i) of the current trajectory; i == N evaluates the terminal cost derivatives. */
template<class M>
__global__ void linearize_kernel(const __grid_constant__ M model,
                                 const __grid_constant__ Workspace<typename M::Scalar> ws,
                                 const __grid_constant__ SolverParams<typename M::Scalar> prm)
{
  pdlPrologue();
  using S = typename M::Scalar;
  constexpr int NX = M::NX, NU = M::NU;
  using L = BlockLayout<NX, NU>;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if(b == 0 && i == 0) *ws.fan_count = 0; // the previous iteration's line-search work list is consumed
  if(b >= ws.B) return;
  if(ws.status[b] != 0) return;
  const size_t Bp = ws.Bp;
  const int sel = ws.sel[b];
  const S * xs = ws.x[sel];
  const S * us = ws.u[sel];

  Matrix<S, NX, 1> x;
#pragma unroll
  for(int d = 0; d < NX; d++) x[d] = xs[((size_t)i * NX + d) * Bp + b];
  const S t = prm.t0 + i * model.dt();

  if(i == prm.N)
  {
    Matrix<S, NX, 1> Vx;
    Matrix<S, NX, NX> Vxx;
    model.calcTerminalCostDeriv(t, x, Vx, Vxx);
#pragma unroll
    for(int d = 0; d < NX; d++) ws.vterm[(size_t)d * Bp + b] = Vx[d];
#pragma unroll
    for(int d = 0; d < NX * NX; d++) ws.vterm[(size_t)(NX + d) * Bp + b] = Vxx.d[d];
    return;
  }

  Matrix<S, NU, 1> u;
#pragma unroll
  for(int d = 0; d < NU; d++) u[d] = us[((size_t)i * NU + d) * Bp + b];

  Matrix<S, NX, NX> Fx, Lxx;
  Matrix<S, NX, NU> Fu, Lxu;
  Matrix<S, NX, 1> Lx;
  Matrix<S, NU, 1> Lu;
  Matrix<S, NU, NU> Luu;
  linearizeStep<M>(model, t, x, u, Fx, Fu, Lx, Lu, Lxx, Luu, Lxu);

  S * blk = ws.deriv + derivTileOffset<L::SIZE>(i, b, ws.Bp);
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
}

/* ------------------------------------------------------------------------------------ K2 ---- */
/** In-register Cholesky of an n x n matrix with Eigen::LLT's failure rule (pivot <= 0; a NaN pivot
    passes).  Lower triangle in/out, column-major. */
template<class S, int n>
__device__ __forceinline__ bool lltInPlace(S * a)
{
  bool ok = true;
#pragma unroll
  for(int k = 0; k < n; k++)
  {
    S x = a[k + k * n];
#pragma unroll
    for(int j = 0; j < k; j++) x -= a[k + j * n] * a[k + j * n];
    if(x <= S(0)) ok = false;
    x = sqrt(x);
    a[k + k * n] = x;
    const S inv = S(1) / x;
#pragma unroll
    for(int i = k + 1; i < n; i++)
    {
      S s = a[i + k * n];
#pragma unroll
      for(int j = 0; j < k; j++) s -= a[i + j * n] * a[k + j * n];
      a[i + k * n] = s * inv;
    }
  }
  return ok;
}

/** b <- (L L^T)^-1 b; `invd` holds 1 / L(i,i). */
template<class S, int n>
__device__ __forceinline__ void lltSolveInPlace(const S * l, const S * invd, S * b)
{
#pragma unroll
  for(int i = 0; i < n; i++)
  {
    S s = b[i];
#pragma unroll
    for(int j = 0; j < i; j++) s -= l[i + j * n] * b[j];
    b[i] = s * invd[i];
  }
#pragma unroll
  for(int i = n - 1; i >= 0; i--)
  {
    S s = b[i];
#pragma unroll
    for(int j = i + 1; j < n; j++) s -= l[j + i * n] * b[j];
    b[i] = s * invd[i];
  }
}

__device__ __forceinline__ void mbarArrive(unsigned long long * bar)
{
  const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(a) : "memory");
}

/* Loader-warp staging (used by the FMPC sweeps F2 / F3 and by phase 1 of the phased line search): a second warp of the
   CTA streams the rows a step needs into a shared-memory ring with per-lane cp.async from running pointers -- no
   registers, many loads in flight, running ahead by the ring depth -- and the copies' completion lands on a `full`
   mbarrier (32 arrivals: every lane's copies of its own column); the compute warp releases a stage on `empty`. */
__device__ __forceinline__ void cpAsyncArriveOn(unsigned long long * bar)
{
  const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(bar));
#ifdef NMPC_B200_LOADER_WAITS
  // diagnostic build (tools/sanitize_cases.py, profiles/r2_sanitizer.md): the loader waits for its copies and arrives
  // itself.  compute-sanitizer's racecheck follows an ordinary mbarrier arrive but not the arrive that the hardware
  // performs when a thread's cp.async operations complete; with this build its cp.async hazards disappear while the
  // results stay bit-identical -- i.e. they are a limitation of the tool, not races.
  asm volatile("cp.async.wait_all;\n" ::: "memory");
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(a) : "memory");
#else
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(a) : "memory");
#endif
}

/** Loader side: rows 0 .. ROWS-1 of `n_fills` consecutive steps.  row_ptr[r] = address of this lane's element of row r
    for the FIRST step, advanced by row_stride[r] (signed, in scalars) per step. */
template<class S, int ROWS, int DEPTH>
__device__ __forceinline__ void loaderLoop(S * ring,
                                           unsigned long long * full,
                                           unsigned long long * empty,
                                           int lane,
                                           int n_fills,
                                           const S * (&row_ptr)[ROWS],
                                           const long long (&row_stride)[ROWS])
{
  for(int f = 0; f < n_fills; f++)
  {
    const int st = f % DEPTH;
    if(f >= DEPTH) mbarWait(&empty[st], (unsigned)((f / DEPTH) - 1) & 1u); // the compute warp is done with it
    S * dst = ring + (size_t)st * ROWS * kTile + lane;
#pragma unroll
    for(int r = 0; r < ROWS; r++)
    {
      if constexpr(sizeof(S) == 8)
        cpAsync8(dst + (size_t)r * kTile, row_ptr[r]);
      else
        cpAsync4(dst + (size_t)r * kTile, row_ptr[r]);
      row_ptr[r] += row_stride[r];
    }
    cpAsyncArriveOn(&full[st]);
  }
}

/** Where a backward sweep gets the derivative tile of step i from.  TmaFeed: K1 wrote the tiles to HBM, lane 0 fetches
    step i-1 with one bulk (TMA) copy into a two-stage ring while step i computes. */
template<class S, int SIZE>
struct TmaFeed
{
  static constexpr bool kFused = false;
  S * ring; //!< this warp's [2][SIZE][32] ring
  unsigned long long * bars; //!< its two mbarriers
  unsigned & parity;
  const S * tile0; //!< this warp's tile of step 0 in ws.deriv
  size_t step_stride;
  int lane;
  int stage;

  __device__ __forceinline__ void stageStep(int st, int step)
  {
    if(lane == 0)
    {
      constexpr unsigned kStageBytes = (unsigned)(sizeof(S) * SIZE * kTile);
      mbarExpectTx(&bars[st], kStageBytes);
      bulkCopyG2S(ring + (size_t)st * SIZE * kTile, tile0 + (size_t)step * step_stride, kStageBytes, &bars[st]);
    }
  }
  __device__ __forceinline__ void begin(int N)
  {
    __syncwarp(); // the previous sweep's readers are done with the ring
    stageStep(0, N - 1);
    stage = 0;
  }
  /** Tile of step i for this lane (element e at blk[e * 32]); also starts the copy of step i-1. */
  __device__ __forceinline__ const S * acquire(int i)
  {
    const S * const blk = ring + (size_t)stage * SIZE * kTile + lane;
    __syncwarp(); // every lane has finished reading the other stage (step i+1)
    if(i > 0) stageStep(stage ^ 1, i - 1);
    mbarWait(&bars[stage], (parity >> stage) & 1u);
    parity ^= (1u << stage);
    stage ^= 1;
    return blk;
  }
  __device__ __forceinline__ void release() {}
};

/** ProducerFeed: no K1 and no HBM round trip -- a PRODUCER warp of the same CTA linearises step after step straight
    into a DEPTH-stage shared-memory ring (ddp_backward_fused.cuh); full/empty mbarriers with 32 arrivals each. */
template<class S, int SIZE, int DEPTH>
struct ProducerFeed
{
  static constexpr bool kFused = true;
  S * ring; //!< [DEPTH][SIZE][32]
  unsigned long long * full;
  unsigned long long * empty;
  unsigned & fill; //!< tiles consumed so far (continues across sweeps)
  int lane;

  __device__ __forceinline__ void begin(int) {}
  __device__ __forceinline__ const S * acquire(int)
  {
    const unsigned st = fill % DEPTH;
    mbarWait(&full[st], (fill / DEPTH) & 1u);
    return ring + (size_t)st * SIZE * kTile + lane;
  }
  __device__ __forceinline__ void release()
  {
    mbarArrive(&empty[fill % DEPTH]);
    fill++;
  }
};

/** One backwardPass() sweep (DDPSolver.hpp:343-534) with regularisation `lambda`, one thread per instance.
    All 32 lanes of the warp execute the loop (it contains the feed's barriers); only lanes with `work` compute.
    Returns false when the Cholesky factorisation of Quu_F failed at some step (LLT NumericalIssue, :500-508) --
    the caller then raises lambda and sweeps again. */
template<class M, bool CONSTRAINED, class Feed>
__device__ __forceinline__ bool backwardSweep(const M & model,
                                              const Workspace<typename M::Scalar> & ws,
                                              const SolverParams<typename M::Scalar> & prm,
                                              int b,
                                              int lane,
                                              const typename M::Scalar * __restrict__ us,
                                              const typename M::Scalar * __restrict__ xs,
                                              Feed & feed,
                                              bool work,
                                              typename M::Scalar lambda,
                                              typename M::Scalar & dV0,
                                              typename M::Scalar & dV1,
                                              typename M::Scalar & k_rel_norm)
{
  using S = typename M::Scalar;
  constexpr int NX = M::NX, NU = M::NU;
  using L = BlockLayout<NX, NU>;
  constexpr int tpb = kTile; // element stride inside a staged tile
  const size_t Bp = ws.Bp;
  const int N = prm.N;
  const int reg_type = prm.reg_type; // loop-invariant solver constants out of the constant bank, once

  S Vx[NX], Vxx[NX * NX];
  if constexpr(Feed::kFused)
  {
    // no K1: the terminal cost derivatives (DDPSolver.hpp:178-180) are evaluated here
    Matrix<S, NX, 1> xN, vx;
    Matrix<S, NX, NX> vxx;
#pragma unroll
    for(int d = 0; d < NX; d++) xN[d] = xs[((size_t)N * NX + d) * Bp + b];
    model.calcTerminalCostDeriv(prm.t0 + N * model.dt(), xN, vx, vxx);
#pragma unroll
    for(int d = 0; d < NX; d++) Vx[d] = vx[d];
#pragma unroll
    for(int d = 0; d < NX * NX; d++) Vxx[d] = vxx.d[d];
  }
  else
  {
#pragma unroll
    for(int d = 0; d < NX; d++) Vx[d] = ws.vterm[(size_t)d * Bp + b];
#pragma unroll
    for(int d = 0; d < NX * NX; d++) Vxx[d] = ws.vterm[(size_t)(NX + d) * Bp + b];
  }

  // accumulated locally and handed back only when this instance needed the sweep: an instance that is merely waiting
  // for its tile mates' lambda retry keeps the dV / k_rel_norm of its own successful sweep
  S dV0_acc = S(0), dV1_acc = S(0);
  // max_i |k_i| / (|u_i| + 1) is tracked as a (numerator, denominator) pair and divided once
  S krn_num = S(0), krn_den = S(1);

  // Shared-memory ring per warp (see the feeds above): the dependent Riccati chain never waits on HBM and no
  // address arithmetic is spent on the block entries -- they are read back with immediate-offset shared loads.
  feed.begin(N);
  S k_prev[NU]; // k_list_[i + 1], the BoxQP warm start (:452-467)
#pragma unroll
  for(int a = 0; a < NU; a++) k_prev[a] = S(0);
  S u_cur[NU], u_nxt[NU], u_nx2[NU]; // u_i for the termination test / input limits, fetched two steps ahead
#pragma unroll
  for(int a = 0; a < NU; a++)
  {
    u_cur[a] = us[((size_t)(N - 1) * NU + a) * Bp + b];
    u_nxt[a] = us[((size_t)(N > 1 ? N - 2 : 0) * NU + a) * Bp + b];
  }
  bool ok = true;
  // running pointers (step N-1 first, one step back per iteration) instead of per-step 64-bit address arithmetic
  S * kff_ptr = ws.kff + (size_t)(N - 1) * NU * Bp + b;
  S * kfb_ptr = ws.kfb + (size_t)(N - 1) * NU * NX * Bp + b;
  const S * u2_ptr = us + (size_t)(N > 2 ? N - 3 : 0) * NU * Bp + b; // u of step i - 2

  for(int i = N - 1; i >= 0; i--)
  {
    {
#pragma unroll
      for(int a = 0; a < NU; a++) u_nx2[a] = u2_ptr[(size_t)a * Bp];
      if(i > 2) u2_ptr -= (size_t)NU * Bp;
    }
    const S * const blk = feed.acquire(i);
    if(work && ok)
    {
      do
      {
    S Fx[NX * NX], Fu[NX * NU];
#pragma unroll
    for(int d = 0; d < NX * NX; d++) Fx[d] = blk[(size_t)(L::FX + d) * tpb];
#pragma unroll
    for(int d = 0; d < NX * NU; d++) Fu[d] = blk[(size_t)(L::FU + d) * tpb];

    // Qu = Lu + Fu^T Vx ; Qx = Lx + Fx^T Vx                                  (:386-388)
    S Qu[NU], Qx[NX];
#pragma unroll
    for(int a = 0; a < NU; a++)
    {
      S s = S(0);
#pragma unroll
      for(int r = 0; r < NX; r++) s += Fu[r + a * NX] * Vx[r];
      Qu[a] = blk[(size_t)(L::LU + a) * tpb] + s;
    }
#pragma unroll
    for(int j = 0; j < NX; j++)
    {
      S s = S(0);
#pragma unroll
      for(int r = 0; r < NX; r++) s += Fx[r + j * NX] * Vx[r];
      Qx[j] = blk[(size_t)(L::LX + j) * tpb] + s;
    }

    // Tu = Fu^T Vxx (NU x NX), Tx = Fx^T Vxx (NX x NX): products associate left to right as in Eigen
    S Tu[NU * NX], Tx[NX * NX];
#pragma unroll
    for(int j = 0; j < NX; j++)
    {
#pragma unroll
      for(int a = 0; a < NU; a++)
      {
        S s = S(0);
#pragma unroll
        for(int r = 0; r < NX; r++) s += Fu[r + a * NX] * Vxx[r + j * NX];
        Tu[a + j * NU] = s;
      }
#pragma unroll
      for(int c = 0; c < NX; c++)
      {
        S s = S(0);
#pragma unroll
        for(int r = 0; r < NX; r++) s += Fx[r + c * NX] * Vxx[r + j * NX];
        Tx[c + j * NX] = s;
      }
    }

    // Qux = Lxu^T + Tu Fx ; Quu = Luu + Tu Fu ; Qxx = Lxx + Tx Fx              (:390-408)
    S Qux[NU * NX], Quu[NU * NU], Qxx[NX * NX];
#pragma unroll
    for(int j = 0; j < NX; j++)
    {
#pragma unroll
      for(int a = 0; a < NU; a++)
      {
        S s = S(0);
#pragma unroll
        for(int r = 0; r < NX; r++) s += Tu[a + r * NU] * Fx[r + j * NX];
        Qux[a + j * NU] = blk[(size_t)(L::LXU + j + a * NX) * tpb] + s;
      }
#pragma unroll
      for(int c = 0; c < NX; c++)
      {
        S s = S(0);
#pragma unroll
        for(int r = 0; r < NX; r++) s += Tx[c + r * NX] * Fx[r + j * NX];
        Qxx[c + j * NX] = blk[(size_t)(L::LXX + c + j * NX) * tpb] + s;
      }
    }
#pragma unroll
    for(int c = 0; c < NU; c++)
#pragma unroll
      for(int a = 0; a < NU; a++)
      {
        S s = S(0);
#pragma unroll
        for(int r = 0; r < NX; r++) s += Tu[a + r * NU] * Fu[r + c * NX];
        Quu[a + c * NU] = blk[(size_t)(L::LUU + a + c * NU) * tpb] + s;
      }

    // regularisation (:421-441)
    S Qux_reg[NU * NX], Quu_F[NU * NU];
    if(reg_type == 2)
    {
      // Vxx_reg = Vxx + lambda I  =>  Tu_reg = Tu + lambda Fu^T
      S Tur[NU * NX];
#pragma unroll
      for(int j = 0; j < NX; j++)
#pragma unroll
        for(int a = 0; a < NU; a++) Tur[a + j * NU] = Tu[a + j * NU] + lambda * Fu[j + a * NX];
#pragma unroll
      for(int j = 0; j < NX; j++)
#pragma unroll
        for(int a = 0; a < NU; a++)
        {
          S s = S(0);
#pragma unroll
          for(int r = 0; r < NX; r++) s += Tur[a + r * NU] * Fx[r + j * NX];
          Qux_reg[a + j * NU] = blk[(size_t)(L::LXU + j + a * NX) * tpb] + s;
        }
#pragma unroll
      for(int c = 0; c < NU; c++)
#pragma unroll
        for(int a = 0; a < NU; a++)
        {
          S s = S(0);
#pragma unroll
          for(int r = 0; r < NX; r++) s += Tur[a + r * NU] * Fu[r + c * NX];
          Quu_F[a + c * NU] = blk[(size_t)(L::LUU + a + c * NU) * tpb] + s;
        }
    }
    else
    {
#pragma unroll
      for(int d = 0; d < NU * NX; d++) Qux_reg[d] = Qux[d];
#pragma unroll
      for(int d = 0; d < NU * NU; d++) Quu_F[d] = Quu[d];
      if(reg_type == 1)
      {
#pragma unroll
        for(int a = 0; a < NU; a++) Quu_F[a + a * NU] += lambda;
      }
    }

    // gains: LLT(Quu_F), k = -Quu_F^-1 Qu, K = -Quu_F^-1 Qux_reg           (:500-510)
    S k[NU], K[NU * NX];
    if constexpr(CONSTRAINED)
    {
      // control-limited gains (:450-497): k = argmin 1/2 k^T Quu_F k + Qu^T k, lo - u <= k <= up - u,
      // warm-started from the next step's k; K rows of clamped inputs are zero
      S lo[NU], hi[NU], init[NU];
#pragma unroll
      for(int a = 0; a < NU; a++)
      {
        const S uv = u_cur[a];
        lo[a] = ws.u_lo[(size_t)i * NU + a] - uv; // input_limits_func_(t_i) (:470)
        hi[a] = ws.u_hi[(size_t)i * NU + a] - uv;
        init[a] = warmStartFromNextStep<M>(model, prm.t0, i, N) ? k_prev[a] : S(0);
      }
      BoxQPResult<S, NU> qp;
      boxQpSolve<S, NU>(Quu_F, Qu, lo, hi, init, qp);
      if(qp.retval < 0)
      {
        ok = false;
        break;
      }
#pragma unroll
      for(int a = 0; a < NU; a++) k[a] = qp.x[a];
#pragma unroll
      for(int d = 0; d < NU * NX; d++) K[d] = S(0);
      const int nf = qp.n_free;
      for(int j = 0; j < NX; j++)
      {
        S rhs[NU];
        for(int r = 0; r < nf; r++) rhs[r] = Qux_reg[qp.free_idxs[r] + j * NU];
        for(int r = 0; r < nf; r++)
        {
          S s = rhs[r];
          for(int q = 0; q < r; q++) s -= qp.llt_free[r + q * nf] * rhs[q];
          rhs[r] = s / qp.llt_free[r + r * nf];
        }
        for(int r = nf - 1; r >= 0; r--)
        {
          S s = rhs[r];
          for(int q = r + 1; q < nf; q++) s -= qp.llt_free[q + r * nf] * rhs[q];
          rhs[r] = s / qp.llt_free[r + r * nf];
        }
        for(int r = 0; r < nf; r++) K[qp.free_idxs[r] + j * NU] = S(-1) * rhs[r];
      }
    }
    else if constexpr(NU == 1)
    {
      // 1x1: the LLT failure rule is "Quu_F <= 0"; L L^T solve == one reciprocal
      if(Quu_F[0] <= S(0))
      {
        ok = false;
        break;
      }
      const S inv = S(1) / Quu_F[0];
      k[0] = -(Qu[0] * inv);
#pragma unroll
      for(int j = 0; j < NX; j++) K[j] = -(Qux_reg[j] * inv);
    }
    else
    {
      if(!lltInPlace<S, NU>(Quu_F))
      {
        ok = false;
        break;
      }
      S invd[NU];
#pragma unroll
      for(int a = 0; a < NU; a++) invd[a] = S(1) / Quu_F[a + a * NU];
#pragma unroll
      for(int a = 0; a < NU; a++) k[a] = Qu[a];
      lltSolveInPlace<S, NU>(Quu_F, invd, k);
#pragma unroll
      for(int a = 0; a < NU; a++) k[a] = -k[a];
#pragma unroll
      for(int j = 0; j < NX; j++)
      {
        S col[NU];
#pragma unroll
        for(int a = 0; a < NU; a++) col[a] = Qux_reg[a + j * NU];
        lltSolveInPlace<S, NU>(Quu_F, invd, col);
#pragma unroll
        for(int a = 0; a < NU; a++) K[a + j * NU] = -col[a];
      }
    }

    // cost-to-go (:522-526)
    S Quuk[NU];
#pragma unroll
    for(int a = 0; a < NU; a++)
    {
      S s = S(0);
#pragma unroll
      for(int c = 0; c < NU; c++) s += Quu[a + c * NU] * k[c];
      Quuk[a] = s;
    }
    {
      S s0 = S(0), s1 = S(0);
#pragma unroll
      for(int a = 0; a < NU; a++)
      {
        s0 += k[a] * Qu[a];
        s1 += k[a] * Quuk[a];
      }
      dV0_acc += s0;
      dV1_acc += S(0.5) * s1;
    }
    // KtQuu = K^T Quu (NX x NU)
    S KtQuu[NX * NU];
#pragma unroll
    for(int c = 0; c < NU; c++)
#pragma unroll
      for(int j = 0; j < NX; j++)
      {
        S s = S(0);
#pragma unroll
        for(int a = 0; a < NU; a++) s += K[a + j * NU] * Quu[a + c * NU];
        KtQuu[j + c * NX] = s;
      }
    // Vx = Qx + K^T Quu k + K^T Qu + Qux^T k
#pragma unroll
    for(int j = 0; j < NX; j++)
    {
      S s1 = S(0), s2 = S(0), s3 = S(0);
#pragma unroll
      for(int a = 0; a < NU; a++)
      {
        s1 += KtQuu[j + a * NX] * k[a];
        s2 += K[a + j * NU] * Qu[a];
        s3 += Qux[a + j * NU] * k[a];
      }
      Vx[j] = ((Qx[j] + s1) + s2) + s3;
    }
    // Vxx = Qxx + K^T Quu K + K^T Qux + Qux^T K, then symmetrise
    S Vn[NX * NX];
#pragma unroll
    for(int j = 0; j < NX; j++)
#pragma unroll
      for(int r = 0; r < NX; r++)
      {
        S s1 = S(0), s2 = S(0), s3 = S(0);
#pragma unroll
        for(int a = 0; a < NU; a++)
        {
          s1 += KtQuu[r + a * NX] * K[a + j * NU];
          s2 += K[a + r * NU] * Qux[a + j * NU];
          s3 += Qux[a + r * NU] * K[a + j * NU];
        }
        Vn[r + j * NX] = ((Qxx[r + j * NX] + s1) + s2) + s3;
      }
#pragma unroll
    for(int j = 0; j < NX; j++)
#pragma unroll
      for(int r = 0; r < NX; r++) Vxx[r + j * NX] = S(0.5) * (Vn[r + j * NX] + Vn[j + r * NX]);

#pragma unroll
    for(int a = 0; a < NU; a++) k_prev[a] = k[a];

    // save gains (:529-530) and accumulate max_i |k_i| / (|u_i| + 1) (:217-221)
    S kn = S(0), un = S(0);
#pragma unroll
    for(int a = 0; a < NU; a++)
    {
      kff_ptr[(size_t)a * Bp] = k[a];
      kn += k[a] * k[a];
      const S uv = u_cur[a];
      un += uv * uv;
    }
#pragma unroll
    for(int d = 0; d < NU * NX; d++) kfb_ptr[(size_t)d * Bp] = K[d];
    {
      // |k| / (|u| + 1) > num / den  <=>  |k| * den > num * (|u| + 1)   (both denominators >= 1)
      const S a_num = (NU == 1) ? fabs(k[0]) : sqrt(kn);
      const S a_den = ((NU == 1) ? fabs(u_cur[0]) : sqrt(un)) + S(1);
      if(a_num * krn_den > krn_num * a_den)
      {
        krn_num = a_num;
        krn_den = a_den;
      }
    }
      } while(0);
    }
    feed.release();
    kff_ptr -= (size_t)NU * Bp;
    kfb_ptr -= (size_t)NU * NX * Bp;
#pragma unroll
    for(int a = 0; a < NU; a++)
    {
      u_cur[a] = u_nxt[a];
      u_nxt[a] = u_nx2[a];
    }
  }
  if(work)
  {
    dV0 = dV0_acc;
    dV1 = dV1_acc;
    k_rel_norm = krn_num / krn_den;
  }
  return ok;
}

/** procOnce() Step 2 (DDPSolver.hpp:188-231): retry the backward sweep with larger lambda until the
    factorisation succeeds, then the small-gradient termination test.  blockDim.x is a multiple of 32;
    each warp owns one 32-instance tile and a private two-stage TMA ring in shared memory. */
template<class M, bool CONSTRAINED>
__global__ void backward_kernel(const __grid_constant__ M model,
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
  const int n_warps = blockDim.x >> 5;
  // [n_warps][2 stages][BLK][32] tiles, then 2 mbarriers per warp
  S * ring = reinterpret_cast<S *>(smem_raw) + (size_t)warp * 2 * L::SIZE * kTile;
  unsigned long long * bars =
      reinterpret_cast<unsigned long long *>(smem_raw + sizeof(S) * (size_t)n_warps * 2 * L::SIZE * kTile) + 2 * warp;
  if(lane == 0)
  {
    mbarInit(&bars[0], 1);
    mbarInit(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncwarp();
  unsigned parity = 0u;

  const int bg = blockIdx.x * blockDim.x + threadIdx.x; // ws.Bp is a multiple of 128, so the tile is always in range
  const int b = (bg < ws.B) ? bg : bg; // padded instances read valid (padding) memory and never write
  const bool live = (bg < ws.B) && (ws.status[bg < ws.B ? bg : 0] == 0);

  S lambda = live ? ws.lambda[b] : S(0);
  S dlambda = live ? ws.dlambda[b] : S(0);
  const int sel = live ? ws.sel[b] : 0;
  const S * us = ws.u[sel];
  const S * xs = ws.x[sel];
  int n_bwd = live ? ws.n_bwd[b] : 0;
  S dV0 = S(0), dV1 = S(0), k_rel_norm = S(0);
  bool need = live;
  bool failed = false;
  TmaFeed<S, L::SIZE> feed{ring, bars, parity, ws.deriv + derivTileOffset<L::SIZE>(0, b - lane, ws.Bp),
                           (size_t)(ws.Bp / kTile) * L::SIZE * kTile, lane, 0};
  while(__any_sync(kFull, need))
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

/* ------------------------------------------------------------------------------------ K3 ---- */
/** forwardPass(alpha) (DDPSolver.hpp:537-560) for instance b from its current trajectory (buffer
    `sel`).  STORE: write the candidate trajectory into buffer sel^1; otherwise only the candidate's
    total cost is computed.  Both variants execute the same arithmetic in the same order, so a cost
    obtained without stores is reproduced bit for bit by the storing run. */
template<class M, bool STORE>
__device__ __forceinline__ typename M::Scalar forwardRollout(const M & model,
                                                             const Workspace<typename M::Scalar> & ws,
                                                             const SolverParams<typename M::Scalar> & prm,
                                                             int b,
                                                             int sel,
                                                             typename M::Scalar alpha)
{
  using S = typename M::Scalar;
  constexpr int NX = M::NX, NU = M::NU;
  const size_t Bp = ws.Bp;
  const int N = prm.N;
  const S * __restrict__ xc = ws.x[sel];
  const S * __restrict__ uc = ws.u[sel];
  S * __restrict__ xn = ws.x[sel ^ 1];
  S * __restrict__ un = ws.u[sel ^ 1];
  S * __restrict__ cn = ws.cost[sel ^ 1];

  Matrix<S, NX, 1> x;
#pragma unroll
  for(int d = 0; d < NX; d++) x[d] = xc[(size_t)d * Bp + b];
  if(STORE)
  {
#pragma unroll
    for(int d = 0; d < NX; d++) xn[(size_t)d * Bp + b] = x[d]; // candidate x_list[0] (:540)
  }

  S csum = S(0);
  // software prefetch: step i+1's operands are loaded while step i computes
  S xr[NX], ur[NU], kr[NU], Kr[NU * NX];
#pragma unroll
  for(int d = 0; d < NX; d++) xr[d] = x[d];
#pragma unroll
  for(int d = 0; d < NU; d++) ur[d] = ldStream(uc + (size_t)d * Bp + b);
#pragma unroll
  for(int d = 0; d < NU; d++) kr[d] = ldStream(ws.kff + (size_t)d * Bp + b);
#pragma unroll
  for(int d = 0; d < NU * NX; d++) Kr[d] = ldStream(ws.kfb + (size_t)d * Bp + b);
  for(int i = 0; i < N; i++)
  {
    S xr_n[NX], ur_n[NU], kr_n[NU], Kr_n[NU * NX];
    const int ip = (i + 1 < N) ? i + 1 : i;
#pragma unroll
    for(int d = 0; d < NX; d++) xr_n[d] = ldStream(xc + ((size_t)ip * NX + d) * Bp + b);
#pragma unroll
    for(int d = 0; d < NU; d++) ur_n[d] = ldStream(uc + ((size_t)ip * NU + d) * Bp + b);
#pragma unroll
    for(int d = 0; d < NU; d++) kr_n[d] = ldStream(ws.kff + ((size_t)ip * NU + d) * Bp + b);
#pragma unroll
    for(int d = 0; d < NU * NX; d++) Kr_n[d] = ldStream(ws.kfb + ((size_t)ip * NU * NX + d) * Bp + b);

    // u' = u + alpha k + K (x' - x)                                       (:545-546)
    Matrix<S, NU, 1> u;
#pragma unroll
    for(int a = 0; a < NU; a++)
    {
      S s = S(0);
#pragma unroll
      for(int j = 0; j < NX; j++) s += Kr[a + j * NU] * (x[j] - xr[j]);
      u[a] = (ur[a] + alpha * kr[a]) + s;
      if(STORE) un[((size_t)i * NU + a) * Bp + b] = u[a];
    }
    const S t = prm.t0 + i * model.dt();
    const S c = model.runningCost(t, x, u);
    x = model.stateEq(t, x, u);
    if(STORE)
    {
#pragma unroll
      for(int d = 0; d < NX; d++) xn[((size_t)(i + 1) * NX + d) * Bp + b] = x[d];
      cn[(size_t)i * Bp + b] = c;
    }
    csum += c;

#pragma unroll
    for(int d = 0; d < NX; d++) xr[d] = xr_n[d];
#pragma unroll
    for(int d = 0; d < NU; d++) ur[d] = ur_n[d];
#pragma unroll
    for(int d = 0; d < NU; d++) kr[d] = kr_n[d];
#pragma unroll
    for(int d = 0; d < NU * NX; d++) Kr[d] = Kr_n[d];
  }
  {
    const S t = prm.t0 + N * model.dt();
    const S c = model.terminalCost(t, x);
    if(STORE) cn[(size_t)N * Bp + b] = c;
    csum += c;
  }
  return csum;
}

/** Acceptance test of one line-search candidate (DDPSolver.hpp:248-264). */
template<class S>
__device__ __forceinline__ bool lineSearchTest(const SolverParams<S> & prm,
                                               S cost_cur,
                                               S cost_cand,
                                               S alpha,
                                               S dV0,
                                               S dV1,
                                               S & actual,
                                               S & expected,
                                               S & ratio)
{
  actual = cost_cur - cost_cand;
  expected = S(-1) * alpha * (dV0 + alpha * dV1);
  ratio = actual / expected;
  if(expected < S(0)) ratio = (actual >= S(0)) ? S(1) : S(-1);
  return ratio > prm.cost_update_ratio_thre;
}

/** procOnce() Steps 3-4 (DDPSolver.hpp:234-339): backtracking line search over alpha_list, then the
    accept/reject bookkeeping.  On success the two trajectory buffers swap roles (sel ^= 1) instead of
    the reference's three copies (:285-287).

    forwardPass(alpha) is a pure function of (current trajectory, k, K, alpha), so trying the
    candidates in parallel and keeping the first success in list order gives the reference's result.
    Every thread first tries alpha_list[0] for its own instance.  The (few) instances for which that
    fails are then served by the whole warp: the remaining n_alpha-1 candidates of up to
    32/(n_alpha-1) failed instances are rolled out concurrently, one candidate per lane, without
    stores; the winner (if any) is rolled out once more by the owning thread with stores.  A warp
    therefore spends 1 + ceil(F / 3) (+1) rollouts instead of up to 11 when F of its instances need
    backtracking. */
template<class M>
__global__ void forward_kernel(const __grid_constant__ M model,
                               const __grid_constant__ Workspace<typename M::Scalar> ws,
                               const __grid_constant__ SolverParams<typename M::Scalar> prm,
                               int iter)
{
  pdlPrologue();
  using S = typename M::Scalar;
  constexpr unsigned kFull = 0xffffffffu;
  const int bg = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int b = (bg < ws.B) ? bg : (ws.B - 1); // keep every lane of the warp alive for the shuffles
  const bool active = (bg < ws.B) && (ws.status[b] == 0);
  const size_t Bp = ws.Bp;

  const int sel = ws.sel[b];
  const S cost_cur = ws.cost_sum[b];
  const S dV0 = ws.dV[b];
  const S dV1 = ws.dV[Bp + b];

  bool success = false;
  S alpha = S(0), actual = S(0), expected = S(0), ratio = S(0), cost_new = S(0);
  int tried = 0;
  bool need_store_run = false;

  if(active && prm.n_alpha > 0)
  {
    alpha = prm.alpha_list[0];
    cost_new = forwardRollout<M, true>(model, ws, prm, b, sel, alpha);
    success = lineSearchTest<S>(prm, cost_cur, cost_new, alpha, dV0, dV1, actual, expected, ratio);
    tried = 1;
  }

  const int rem = prm.n_alpha - 1;
  unsigned pending = __ballot_sync(kFull, active && !success && rem > 0);
  if(rem > 0 && rem <= 32)
  {
    const int groups = 32 / rem; // failed instances served per round
    const int g = lane / rem; // this lane's group in a round
    const int ai = 1 + lane % rem; // and its candidate
    while(pending != 0)
    {
      // owners of this round: the `groups` lowest set bits of `pending`
      unsigned m = pending;
      int src = -1;
      for(int q = 0; q < groups && m != 0; q++)
      {
        const int l = __ffs(m) - 1;
        if(q == g) src = l;
        m &= m - 1;
      }
      const unsigned round_mask = pending & ~m;
      pending = m;
      const bool worker = (g < groups) && (src >= 0);
      const int s_lane = worker ? src : lane;
      const int w_b = __shfl_sync(kFull, b, s_lane);
      const int w_sel = __shfl_sync(kFull, sel, s_lane);
      const S w_cost = __shfl_sync(kFull, cost_cur, s_lane);
      const S w_dV0 = __shfl_sync(kFull, dV0, s_lane);
      const S w_dV1 = __shfl_sync(kFull, dV1, s_lane);

      S w_actual = S(0), w_expected = S(0), w_ratio = S(0), w_costn = S(0);
      bool w_ok = false;
      if(worker)
      {
        const S w_alpha = prm.alpha_list[ai];
        w_costn = forwardRollout<M, false>(model, ws, prm, w_b, w_sel, w_alpha);
        w_ok = lineSearchTest<S>(prm, w_cost, w_costn, w_alpha, w_dV0, w_dV1, w_actual, w_expected, w_ratio);
      }
      const unsigned ok_ballot = __ballot_sync(kFull, worker && w_ok);

      // each owner looks up its group's verdict: first success in list order, else the last candidate
      const bool owner = active && ((round_mask >> lane) & 1u);
      int res_lane = lane;
      int win = -1;
      if(owner)
      {
        const int q = __popc(round_mask & ((1u << lane) - 1u));
        const unsigned gm = (ok_ballot >> (q * rem)) & ((rem == 32) ? kFull : ((1u << rem) - 1u));
        win = (gm != 0) ? (__ffs(gm) - 1) : -1;
        res_lane = q * rem + ((win >= 0) ? win : (rem - 1));
      }
      const S r_actual = __shfl_sync(kFull, w_actual, res_lane);
      const S r_expected = __shfl_sync(kFull, w_expected, res_lane);
      const S r_ratio = __shfl_sync(kFull, w_ratio, res_lane);
      const S r_costn = __shfl_sync(kFull, w_costn, res_lane);
      if(owner)
      {
        actual = r_actual;
        expected = r_expected;
        ratio = r_ratio;
        if(win >= 0)
        {
          success = true;
          need_store_run = true;
          cost_new = r_costn;
          alpha = prm.alpha_list[1 + win];
          tried = 2 + win;
        }
        else
        {
          alpha = prm.alpha_list[rem];
          tried = 1 + rem;
        }
      }
    }
  }
  else if(active && !success)
  {
    // more candidates than lanes: plain serial backtracking
    for(int a = 1; a < prm.n_alpha; a++)
    {
      alpha = prm.alpha_list[a];
      cost_new = forwardRollout<M, true>(model, ws, prm, b, sel, alpha);
      tried++;
      success = lineSearchTest<S>(prm, cost_cur, cost_new, alpha, dV0, dV1, actual, expected, ratio);
      if(success) break;
    }
  }

  if(need_store_run)
  {
    // materialise the winning candidate (same arithmetic => same cost as the store-free rollout)
    cost_new = forwardRollout<M, true>(model, ws, prm, b, sel, alpha);
  }

  if(!active) return;

  // Step 4 (:280-333)
  S lambda = ws.lambda[b];
  S dlambda = ws.dlambda[b];
  const S k_rel_norm = ws.trace[((size_t)iter * kTraceFields + 5) * Bp + b];
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

/* --------------------------------------------------------------- K3, small-batch variant ---- */
/** Operands of one forward step for one instance, staged in shared memory:
    [ x_i (NX) | u_i (NU) | k_i (NU) | K_i (NU*NX) ] of the CURRENT trajectory. */
template<int NX, int NU>
struct FwdOperands
{
  static constexpr int X = 0;
  static constexpr int U = X + NX;
  static constexpr int KFF = U + NU;
  static constexpr int KFB = KFF + NU;
  static constexpr int SIZE = KFB + NU * NX;
};

/** forwardPass(alpha) with GA lanes per instance sharing one operand ring.  All 32 lanes execute the
    loop (it contains warp barriers); `work` lanes roll out their own candidate, `do_store` lanes also
    write the candidate trajectory, `gcopy` groups keep the ring fed.  Same arithmetic as
    forwardRollout => same costs. */
template<class S>
struct FwdDest
{
  S * x; //!< [N+1][NX][stride]
  S * u; //!< [N][NU][stride]
  S * c; //!< [N+1][stride]
  size_t stride;
  size_t col;
};

template<class S>
__device__ __forceinline__ FwdDest<S> candidateBuffer(const Workspace<S> & ws, int sel, int b)
{
  return FwdDest<S>{ws.x[sel ^ 1], ws.u[sel ^ 1], ws.cost[sel ^ 1], (size_t)ws.Bp, (size_t)b};
}

template<class M, int GA, int DEPTH>
__device__ __forceinline__ typename M::Scalar forwardRolloutRing(const M & model_in_constant_bank,
                                                                 const Workspace<typename M::Scalar> & ws,
                                                                 const SolverParams<typename M::Scalar> & prm,
                                                                 typename M::Scalar * __restrict__ ring,
                                                                 int g,
                                                                 int a,
                                                                 int b,
                                                                 int sel,
                                                                 typename M::Scalar alpha,
                                                                 bool work,
                                                                 bool do_store,
                                                                 bool gcopy,
                                                                 const FwdDest<typename M::Scalar> & dst)
{
  using S = typename M::Scalar;
  constexpr int NX = M::NX, NU = M::NU;
  using O = FwdOperands<NX, NU>;
  constexpr int IPW = 32 / GA;
  // the functor and the few solver constants of the step loop, copied out of the kernel-parameter constant bank once:
  // read in place they were re-fetched every step (ncu: LDC c[0x0][..] among the top stall sites of the rollout)
  const M model = model_in_constant_bank;
  const S t0 = prm.t0;
  const size_t Bp = ws.Bp;
  const int N = prm.N;
  const S * __restrict__ xc = ws.x[sel];
  const S * __restrict__ uc = ws.u[sel];
  S * __restrict__ xn = dst.x;
  S * __restrict__ un = dst.u;
  S * __restrict__ cn = dst.c;
  const size_t Bd = dst.stride, bd = dst.col;

  // Lane a of the group copies operands a, a+GA, ... of step `step` into ring slot `slot`.  The source of every
  // operand is a RUNNING pointer that advances by its per-step stride: issue() is called for steps 0, 1, 2, ... in
  // order, and ncu showed 18 % of this kernel's stall samples on address registers that the compiler recycled
  // between back-to-back LDGSTS when each address was recomputed from (step, element).
  constexpr int EPL = (O::SIZE + GA - 1) / GA; // operands per lane
  const S * src_ptr[EPL];
  size_t src_stride[EPL];
#pragma unroll
  for(int q = 0; q < EPL; q++)
  {
    const int e = q * GA + a;
    if(e < O::U)
    {
      src_ptr[q] = xc + (size_t)(e - O::X) * Bp + b;
      src_stride[q] = (size_t)NX * Bp;
    }
    else if(e < O::KFF)
    {
      src_ptr[q] = uc + (size_t)(e - O::U) * Bp + b;
      src_stride[q] = (size_t)NU * Bp;
    }
    else if(e < O::KFB)
    {
      src_ptr[q] = ws.kff + (size_t)(e - O::KFF) * Bp + b;
      src_stride[q] = (size_t)NU * Bp;
    }
    else
    {
      src_ptr[q] = ws.kfb + (size_t)((e < O::SIZE ? e : O::KFB) - O::KFB) * Bp + b;
      src_stride[q] = (size_t)NU * NX * Bp;
    }
  }
  auto issue = [&](int step, int slot) {
    if(gcopy && step < N)
    {
#pragma unroll
      for(int q = 0; q < EPL; q++)
      {
        const int e = q * GA + a;
        if(e < O::SIZE)
        {
          S * rdst = ring + ((size_t)slot * O::SIZE + e) * IPW + g;
          if constexpr(sizeof(S) == 8)
            cpAsync8(rdst, src_ptr[q]);
          else
            cpAsync4(rdst, src_ptr[q]);
        }
      }
    }
#pragma unroll
    for(int q = 0; q < EPL; q++) src_ptr[q] += src_stride[q];
    cpAsyncCommit();
  };

  Matrix<S, NX, 1> x;
#pragma unroll
  for(int d = 0; d < NX; d++) x[d] = xc[(size_t)d * Bp + b];
  if(do_store)
  {
#pragma unroll
    for(int d = 0; d < NX; d++) xn[(size_t)d * Bd + bd] = x[d];
  }
  // running store pointers: x_{i+1}, u_i, c_i of the candidate
  S * xs_ptr = xn + (size_t)NX * Bd + bd;
  S * us_ptr = un + bd;
  S * cs_ptr = cn + bd;

  // with GA == 1 every lane copies and reads only its own ring column: no cross-lane hand-over, no warp barriers
  if constexpr(GA > 1) __syncwarp(); // previous users of the ring are done
#pragma unroll
  for(int s = 0; s < DEPTH; s++) issue(s, s);

  S csum = S(0);
  int slot = 0;
  for(int i = 0; i < N; i++)
  {
    cpAsyncWait<DEPTH - 1>(); // this lane's copies of step i have landed
    if constexpr(GA > 1) __syncwarp(); // ... and so have the other lanes'
    S xr[NX], ur[NU], kr[NU], Kr[NU * NX];
    const S * op = ring + (size_t)slot * O::SIZE * IPW + g;
#pragma unroll
    for(int d = 0; d < NX; d++) xr[d] = op[(size_t)(O::X + d) * IPW];
#pragma unroll
    for(int d = 0; d < NU; d++) ur[d] = op[(size_t)(O::U + d) * IPW];
#pragma unroll
    for(int d = 0; d < NU; d++) kr[d] = op[(size_t)(O::KFF + d) * IPW];
#pragma unroll
    for(int d = 0; d < NU * NX; d++) Kr[d] = op[(size_t)(O::KFB + d) * IPW];
    if constexpr(GA > 1) __syncwarp(); // everyone has read the slot: refill it with step i + DEPTH
    issue(i + DEPTH, slot);
    slot = (slot + 1 == DEPTH) ? 0 : slot + 1;

    if(work)
    {
      Matrix<S, NU, 1> u;
#pragma unroll
      for(int c = 0; c < NU; c++)
      {
        S s = S(0);
#pragma unroll
        for(int j = 0; j < NX; j++) s += Kr[c + j * NU] * (x[j] - xr[j]);
        u[c] = (ur[c] + alpha * kr[c]) + s;
        if(do_store) us_ptr[(size_t)c * Bd] = u[c];
      }
      const S t = t0 + i * model.dt();
      const S c = model.runningCost(t, x, u);
      x = model.stateEq(t, x, u);
      if(do_store)
      {
#pragma unroll
        for(int d = 0; d < NX; d++) xs_ptr[(size_t)d * Bd] = x[d];
        *cs_ptr = c;
      }
      csum += c;
    }
    xs_ptr += (size_t)NX * Bd;
    us_ptr += (size_t)NU * Bd;
    cs_ptr += Bd;
  }
  cpAsyncWait<0>();
  if(work)
  {
    const S t = t0 + N * model.dt();
    const S c = model.terminalCost(t, x);
    if(do_store) cn[(size_t)N * Bd + bd] = c;
    csum += c;
  }
  return csum;
}

/** procOnce() Steps 3-4 for small batches: GA lanes per instance, lane a of a group rolls out
    candidate alpha_list[round * GA + a]; the first success in list order wins (identical to the
    reference's sequential backtracking because forwardPass is a pure function of alpha).  Candidate 0
    stores its trajectory speculatively (it wins in ~90 % of all line searches); any other winner is
    rolled out once more with stores.  With GA = 16 every candidate of the default 11-entry alpha_list
    is evaluated in one sweep, and a 4096-instance batch fills 2048 warps instead of 128. */
template<class M, int GA>
__global__ void forward_spec_kernel(const __grid_constant__ M model,
                                    const __grid_constant__ Workspace<typename M::Scalar> ws,
                                    const __grid_constant__ SolverParams<typename M::Scalar> prm,
                                    int iter)
{
  pdlPrologue();
  using S = typename M::Scalar;
  constexpr int NX = M::NX, NU = M::NU;
  constexpr int IPW = 32 / GA;
  constexpr int DEPTH = 4;
  constexpr unsigned kFull = 0xffffffffu;
  constexpr unsigned kGroupMask = (GA == 32) ? kFull : ((1u << GA) - 1u);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int g = lane / GA;
  const int a = lane % GA;
  S * ring = reinterpret_cast<S *>(smem_raw) + (size_t)warp * DEPTH * FwdOperands<NX, NU>::SIZE * IPW;

  const int bg = (blockIdx.x * (blockDim.x >> 5) + warp) * IPW + g;
  const int b = (bg < ws.B) ? bg : (ws.B - 1);
  const bool active = (bg < ws.B) && (ws.status[b] == 0);
  const size_t Bp = ws.Bp;
  const int sel = ws.sel[b];
  const S cost_cur = ws.cost_sum[b];
  const S dV0 = ws.dV[b];
  const S dV1 = ws.dV[Bp + b];

  bool success = false;
  int win_ai = -1;
  S alpha = S(0), actual = S(0), expected = S(0), ratio = S(0), cost_new = S(0);

  const int rounds = (prm.n_alpha + GA - 1) / GA;
  for(int round = 0; round < rounds; round++)
  {
    const int ai = round * GA + a;
    const bool work = active && !success && (ai < prm.n_alpha);
    const unsigned work_ballot = __ballot_sync(kFull, work);
    if(work_ballot == 0) break;
    const bool gcopy = ((work_ballot >> (g * GA)) & kGroupMask) != 0;
    const S my_alpha = prm.alpha_list[(ai < prm.n_alpha) ? ai : 0];
    const S my_cost = forwardRolloutRing<M, GA, DEPTH>(model, ws, prm, ring, g, a, b, sel, my_alpha, work,
                                                       work && (ai == 0), gcopy, candidateBuffer<S>(ws, sel, b));
    S my_actual = S(0), my_expected = S(0), my_ratio = S(0);
    bool ok = false;
    if(work) ok = lineSearchTest<S>(prm, cost_cur, my_cost, my_alpha, dV0, dV1, my_actual, my_expected, my_ratio);
    const unsigned ok_ballot = __ballot_sync(kFull, ok);
    const unsigned gm = (ok_ballot >> (g * GA)) & kGroupMask;
    // verdict of this round for the group: first success, else the last candidate tried
    const int n_here = min(GA, prm.n_alpha - round * GA);
    const int pick = (gm != 0) ? (__ffs(gm) - 1) : (n_here - 1);
    const bool take = active && !success;
    const int src = take ? (g * GA + pick) : lane;
    const S r_actual = __shfl_sync(kFull, my_actual, src);
    const S r_expected = __shfl_sync(kFull, my_expected, src);
    const S r_ratio = __shfl_sync(kFull, my_ratio, src);
    const S r_cost = __shfl_sync(kFull, my_cost, src);
    const S r_alpha = __shfl_sync(kFull, my_alpha, src);
    if(take)
    {
      actual = r_actual;
      expected = r_expected;
      ratio = r_ratio;
      alpha = r_alpha;
      if(gm != 0)
      {
        success = true;
        win_ai = round * GA + pick;
        cost_new = r_cost;
      }
    }
  }

  // a winner other than candidate 0 has not been stored yet
  const bool need = active && success && (win_ai != 0);
  const unsigned need_ballot = __ballot_sync(kFull, need && a == 0);
  if(need_ballot != 0)
  {
    const S c2 = forwardRolloutRing<M, GA, DEPTH>(model, ws, prm, ring, g, a, b, sel, alpha, need && a == 0,
                                                  need && a == 0, need, candidateBuffer<S>(ws, sel, b));
    if(need && a == 0) cost_new = c2; // bit-identical to the store-free rollout of the same candidate
  }

  if(!active || a != 0) return;

  // Step 4 (:280-333)
  const int tried = success ? (win_ai + 1) : prm.n_alpha;
  S lambda = ws.lambda[b];
  S dlambda = ws.dlambda[b];
  const S k_rel_norm = ws.trace[((size_t)iter * kTraceFields + 5) * Bp + b];
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
} // namespace ddp
} // namespace nmpc_b200
