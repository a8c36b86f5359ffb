/* nmpc_b200 -- FMPC (multiple shooting + primal-dual interior point + Riccati recursion) stage kernels.
 *
 * Reference: isri-aist/NMPC nmpc_fmpc/include/nmpc_fmpc/FmpcSolver.hpp.  Same mapping as the DDP
 * kernels: one thread per instance for the sequential sweeps, one thread per (instance, step) for
 * the embarrassingly parallel stages; batch index innermost in every array.
 *
 *   F0 fmpc_init_kernel      solve(): status/trace reset, optional init_complementary_variable  :166-188
 *   F1 fmpc_coeff_kernel     procOnce() Step 1 + per-step KKT residual terms                    :401-440, :496-521
 *   F2 fmpc_backward_kernel  barrier update, KKT test, backwardPass() (same Riccati sweep shape
 *                            as ddp::backward_kernel plus the C/D slack terms)                  :377-399, :443-448, :524-665
 *   F3 fmpc_forward_kernel   forwardPass() + fraction-to-boundary rule                          :668-708, :714-750
 *   F4 fmpc_update_kernel    updateVariables() element-wise part                                :802-831
 *
 * Device layout (Bp = padded batch):
 *   var   x[N+1][NX], u[N][NU], lam[N+1][NX], s[N][NG], nu[N][NG]     (each [..][Bp]), updated in place
 *   delta same shapes
 *   coeff [N][CSIZE][Bp]  {A, B, C, D, Lxx, Luu, Lxu, x_bar, g_bar, Lx_bar, Lu_bar}
 *   term  [NX + NX*NX + NX][Bp]   terminal {Lx, Lxx, Lx_bar}
 *   gains k[N][NU], K[N][NU*NX], sv[N+1][NX], P[N+1][NX*NX]
 *   kkt   [N+2][Bp]  squared residual terms: [0] initial-state, [1..N] per step, [N+1] terminal
 */
#pragma once

#include <cuda_runtime.h>

#include <nmpc_b200/matrix.h>

#include "ddp_kernels.cuh" // mbarrier / bulk-copy helpers, kTile
#include "fmpc_linalg.cuh"

namespace nmpc_b200
{
namespace fmpc
{
using ddp::kTile;

/* The sequential sweeps F2 / F3 read ~80 scalars per instance and horizon step.  Loaded where they are used, they put
   several dependent HBM latencies on every step of a chain that is already one warp's instruction stream; fetched with
   per-row bulk (TMA) copies they are ~80 small TMA requests per step, which is no faster (measured: 4.4 us per step
   either way).  Instead each 32-instance tile gets a LOADER warp next to its compute warp: the loader streams the rows
   of step after step into a shared-memory ring with per-lane cp.async (no registers, many loads in flight, running
   ahead by the ring depth), completion lands on a `full` mbarrier through cp.async.mbarrier.arrive.noinc (32 arrivals:
   every lane's copies of its own column); the compute warp releases a stage on an `empty` mbarrier.  Layouts in HBM
   stay the plain batch-innermost ones. */
using ddp::cpAsyncArriveOn;
using ddp::loaderLoop;

constexpr int kTraceFields = 5; // iter, kkt_error, barrier_eps, alpha_s, alpha_nu

// FmpcSolver::Status (FmpcSolver.h:92-114)
constexpr int kUninitialized = 0;
constexpr int kSucceeded = 1;
constexpr int kErrorInForward = 2;
constexpr int kErrorInBackward = 3;
constexpr int kErrorInUpdate = 4;
constexpr int kMaxIterationReached = 5;
constexpr int kIterationContinued = 6;

template<class S>
struct SolverParams
{
  int N;
  int max_iter;
  int check_nan;
  int init_complementary_variable;
  int update_barrier_eps;
  int break_if_llt_fails;
  int merit_const_scale_from_lagrange_multipliers;
  int keep_barrier_eps; //!< MPC loop, ticks after the first: barrier_eps_ persists across solve() calls (FmpcSolver.h:413-414)
  S t0;
  S kkt_error_thre;
  S initial_barrier_eps;
};

template<class S>
struct Workspace
{
  int B;
  int Bp;
  S * x0; //!< [NX][Bp] current_x
  S * x;
  S * u;
  S * lam;
  S * s;
  S * nu;
  S * dx;
  S * du;
  S * dlam;
  S * ds;
  S * dnu;
  S * coeff;
  S * term;
  S * kff;
  S * kfb;
  S * sv;
  S * P;
  S * kkt;
  S * trace; //!< [max_iter][5][Bp]
  S * barrier_eps; //!< [Bp]
  S * alpha; //!< [2][Bp] alpha_s, alpha_nu of the current iteration
  int * status;
  int * n_trace;
  int * bad_input; //!< [1] set when an initial s / nu is negative (checkVariable)
};

template<int NX, int NU, int NG>
struct CoeffLayout
{
  static constexpr int A = 0;
  static constexpr int B = A + NX * NX;
  static constexpr int C = B + NX * NU;
  static constexpr int D = C + NG * NX;
  static constexpr int LXX = D + NG * NU;
  static constexpr int LUU = LXX + NX * NX;
  static constexpr int LXU = LUU + NU * NU;
  static constexpr int XBAR = LXU + NX * NU;
  static constexpr int GBAR = XBAR + NX;
  static constexpr int LXBAR = GBAR + NG;
  static constexpr int LUBAR = LXBAR + NX;
  static constexpr int SIZE = LUBAR + NU;
  // terminal entry
  static constexpr int T_LX = 0;
  static constexpr int T_LXX = NX;
  static constexpr int T_LXBAR = NX + NX * NX;
  static constexpr int T_SIZE = NX + NX * NX + NX;
};

/** Rows staged per step by the backward sweep F2: the coefficient block, then s_i, then nu_i. */
template<int NX, int NU, int NG>
struct BackwardRows
{
  static constexpr int ROWS = CoeffLayout<NX, NU, NG>::SIZE + 2 * NG;
  static constexpr int kDepth = 3; //!< ring stages between the loader warp and the compute warp
  /** Shared memory of one CTA (one 32-instance tile: ring + full / empty mbarriers). */
  static size_t bytes(size_t scalar_bytes)
  {
    return scalar_bytes * (size_t)kDepth * ROWS * kTile + sizeof(unsigned long long) * 2 * kDepth + 16;
  }
};

/** Rows staged per step by the forward sweep F3: gains {P_i, s_i, K_i, k_i}, coefficients {A, B, C, D, x_bar, g_bar},
    then s_i, nu_i of the Variable. */
template<int NX, int NU, int NG>
struct ForwardRows
{
  using L = CoeffLayout<NX, NU, NG>;
  static constexpr int P = 0;
  static constexpr int SV = P + NX * NX;
  static constexpr int KFB = SV + NX;
  static constexpr int KFF = KFB + NU * NX;
  static constexpr int ABCD = KFF + NU; //!< rows L::A .. L::LXX-1 of the coefficient block
  static constexpr int N_ABCD = L::LXX;
  static constexpr int XG = ABCD + N_ABCD; //!< rows L::XBAR .. L::GBAR+NG-1
  static constexpr int N_XG = NX + NG;
  static constexpr int S_ = XG + N_XG;
  static constexpr int NU_ = S_ + NG;
  static constexpr int ROWS = NU_ + NG;
  static constexpr int kDepth = 4; //!< an F3 step is short: let the loader run further ahead
  /** Shared memory of one CTA (one 32-instance tile: ring + full / empty mbarriers). */
  static size_t bytes(size_t scalar_bytes)
  {
    return scalar_bytes * (size_t)kDepth * ROWS * kTile + sizeof(unsigned long long) * 2 * kDepth + 16;
  }
};

template<class S>
__device__ __forceinline__ bool finite(S v)
{
  return !(isnan(v) || isinf(v));
}

/** Does the functor have a time-varying inequality dimension, `int ineqDim(t)` <= NG (FmpcProblem<.., Eigen::Dynamic>,
    FmpcProblem.h:62-86)?  Rows j >= ineqDim(t_i) of step i are PADDING: F0 pins s = 1, nu = 0 there, F1 zeroes g + s, C
    and D, F2 counts only real rows in the barrier average, F3 keeps delta s = delta nu = 0 and the merit function
    ignores them.  Every active quantity then has the value the reference computes with vectors of size ineqDim(t)
    (tests/golden/reference_fmpc_dynamic.npz: the reference's own FmpcSolver<4, 1, Eigen::Dynamic>). */
template<class M, class = void>
struct HasIneqDim : std::false_type
{
};
template<class M>
struct HasIneqDim<M, std::void_t<decltype(std::declval<const M &>().ineqDim(std::declval<typename M::Scalar>()))>>
: std::true_type
{
};
template<class M>
__device__ __forceinline__ int ineqDimAt(const M & model, typename M::Scalar t)
{
  if constexpr(HasIneqDim<M>::value)
    return model.ineqDim(t);
  else
    return M::NG;
}

/* ------------------------------------------------------------------------------------ F0 ---- */
/** solve() prologue per (instance, step): optional init_complementary_variable (FmpcSolver.hpp:172-188)
    and the non-negativity part of checkVariable (:348-361); step 0 also resets the per-instance state. */
template<class M>
__global__ void fmpc_init_kernel(const __grid_constant__ M model,
                                 const __grid_constant__ Workspace<typename M::Scalar> ws,
                                 const __grid_constant__ SolverParams<typename M::Scalar> prm)
{
  using S = typename M::Scalar;
  constexpr int NX = M::NX, NU = M::NU, NG = M::NG;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if(b >= ws.B) return;
  const size_t Bp = ws.Bp;
  if(i == 0)
  {
    ws.status[b] = kIterationContinued;
    ws.n_trace[b] = 0;
    // init_complementary_variable resets the barrier parameter (:178); otherwise the member keeps its
    // value from the previous solve() -- the engine seeds it with initial_barrier_eps
    if(!prm.keep_barrier_eps) ws.barrier_eps[b] = prm.initial_barrier_eps;
  }
  const int ng_act = ineqDimAt<M>(model, prm.t0 + i * model.dt());
  if(prm.init_complementary_variable)
  {
    const S margin_rate = S(1e-2);
    const S var_min = S(1e-2);
    const S eps0 = S(1e-4);
    Matrix<S, NX, 1> x;
    Matrix<S, NU, 1> u;
#pragma unroll
    for(int d = 0; d < NX; d++) x[d] = ws.x[((size_t)i * NX + d) * Bp + b];
#pragma unroll
    for(int d = 0; d < NU; d++) u[d] = ws.u[((size_t)i * NU + d) * Bp + b];
    const S t = prm.t0 + i * model.dt();
    const Matrix<S, NG, 1> g = model.ineqConst(t, x, u);
#pragma unroll
    for(int j = 0; j < NG; j++)
    {
      const S sj = (S(1) + margin_rate) * fmax(S(-1) * g[j], var_min);
      const S nj = (S(1) + margin_rate) * fmax(eps0 * (S(1) / sj), var_min);
      ws.s[((size_t)i * NG + j) * Bp + b] = (j < ng_act) ? sj : S(1);
      ws.nu[((size_t)i * NG + j) * Bp + b] = (j < ng_act) ? nj : S(0);
    }
    if(i == 0) ws.barrier_eps[b] = eps0;
  }
  else
  {
    bool neg = false;
#pragma unroll
    for(int j = 0; j < NG; j++)
    {
      if(j < ng_act)
        neg = neg || (ws.s[((size_t)i * NG + j) * Bp + b] < S(0)) || (ws.nu[((size_t)i * NG + j) * Bp + b] < S(0));
      else
      {
        ws.s[((size_t)i * NG + j) * Bp + b] = S(1); // padding row of a time-varying inequality dimension
        ws.nu[((size_t)i * NG + j) * Bp + b] = S(0);
      }
    }
    if(neg) atomicExch(ws.bad_input, 1);
  }
}

/* ------------------------------------------------------------------------------------ F1 ---- */
/** procOnce() Step 1 (FmpcSolver.hpp:401-440) for (instance b, step i), i == N is the terminal entry;
    also the squared KKT residual terms of this step (calcKktError with barrier_eps = 0, :496-521). */
template<class M>
__global__ void fmpc_coeff_kernel(const __grid_constant__ M model,
                                  const __grid_constant__ Workspace<typename M::Scalar> ws,
                                  const __grid_constant__ SolverParams<typename M::Scalar> prm)
{
  using S = typename M::Scalar;
  constexpr int NX = M::NX, NU = M::NU, NG = M::NG;
  using L = CoeffLayout<NX, NU, NG>;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if(b >= ws.B) return;
  if(ws.status[b] != kIterationContinued) return;
  const size_t Bp = ws.Bp;
  const int N = prm.N;
  const S dt = model.dt();
  const S t = prm.t0 + i * dt;

  Matrix<S, NX, 1> x, lambda;
#pragma unroll
  for(int d = 0; d < NX; d++)
  {
    x[d] = ws.x[((size_t)i * NX + d) * Bp + b];
    lambda[d] = ws.lam[((size_t)i * NX + d) * Bp + b];
  }

  if(i == N)
  {
    Matrix<S, NX, 1> Lx;
    Matrix<S, NX, NX> Lxx;
    model.calcTerminalCostDeriv(t, x, Lx, Lxx);
    S kk = S(0);
#pragma unroll
    for(int d = 0; d < NX; d++)
    {
      const S lxb = Lx[d] - lambda[d]; // (2.25a)
      ws.term[(size_t)(L::T_LX + d) * Bp + b] = Lx[d];
      ws.term[(size_t)(L::T_LXBAR + d) * Bp + b] = lxb;
      kk += lxb * lxb;
    }
#pragma unroll
    for(int d = 0; d < NX * NX; d++) ws.term[(size_t)(L::T_LXX + d) * Bp + b] = Lxx.d[d];
    ws.kkt[(size_t)(N + 1) * Bp + b] = kk;
    return;
  }

  Matrix<S, NX, 1> next_x, next_lambda;
  Matrix<S, NU, 1> u;
  Matrix<S, NG, 1> s, nu;
#pragma unroll
  for(int d = 0; d < NX; d++)
  {
    next_x[d] = ws.x[((size_t)(i + 1) * NX + d) * Bp + b];
    next_lambda[d] = ws.lam[((size_t)(i + 1) * NX + d) * Bp + b];
  }
#pragma unroll
  for(int d = 0; d < NU; d++) u[d] = ws.u[((size_t)i * NU + d) * Bp + b];
#pragma unroll
  for(int d = 0; d < NG; d++)
  {
    s[d] = ws.s[((size_t)i * NG + d) * Bp + b];
    nu[d] = ws.nu[((size_t)i * NG + d) * Bp + b];
  }

  Matrix<S, NX, NX> A, Lxx;
  Matrix<S, NX, NU> Bm, Lxu;
  Matrix<S, NG, NX> C;
  Matrix<S, NG, NU> D;
  Matrix<S, NX, 1> Lx;
  Matrix<S, NU, 1> Lu;
  Matrix<S, NU, NU> Luu;
  model.calcStateEqDeriv(t, x, u, A, Bm);
  model.calcIneqConstDeriv(t, x, u, C, D);
  model.calcRunningCostDeriv(t, x, u, Lx, Lu, Lxx, Luu, Lxu);

  const Matrix<S, NX, 1> x_bar = model.stateEq(t, x, u) - next_x; // (2.23c)
  Matrix<S, NG, 1> g_bar = model.ineqConst(t, x, u) + s; // (2.23d)
  if constexpr(HasIneqDim<M>::value)
  {
    const int ng_act = model.ineqDim(t);
#pragma unroll
    for(int j = 0; j < NG; j++)
    {
      if(j >= ng_act)
      {
        g_bar[j] = S(0);
#pragma unroll
        for(int c = 0; c < NX; c++) C(j, c) = S(0);
#pragma unroll
        for(int c = 0; c < NU; c++) D(j, c) = S(0);
      }
    }
  }
  // (2.25b) Lx_bar = -lambda + dt Lx + A^T next_lambda + C^T nu
  Matrix<S, NX, 1> Lx_bar;
#pragma unroll
  for(int j = 0; j < NX; j++)
  {
    S atl = S(0), ctn = S(0);
#pragma unroll
    for(int r = 0; r < NX; r++) atl += A(r, j) * next_lambda[r];
#pragma unroll
    for(int r = 0; r < NG; r++) ctn += C(r, j) * nu[r];
    Lx_bar[j] = ((S(-1) * lambda[j] + dt * Lx[j]) + atl) + ctn;
  }
  // (2.25c) Lu_bar = dt Lu + B^T next_lambda + D^T nu
  Matrix<S, NU, 1> Lu_bar;
#pragma unroll
  for(int j = 0; j < NU; j++)
  {
    S btl = S(0), dtn = S(0);
#pragma unroll
    for(int r = 0; r < NX; r++) btl += Bm(r, j) * next_lambda[r];
#pragma unroll
    for(int r = 0; r < NG; r++) dtn += D(r, j) * nu[r];
    Lu_bar[j] = (dt * Lu[j] + btl) + dtn;
  }

  S * blk = ws.coeff + (size_t)i * L::SIZE * Bp + b;
#pragma unroll
  for(int d = 0; d < NX * NX; d++) blk[(size_t)(L::A + d) * Bp] = A.d[d];
#pragma unroll
  for(int d = 0; d < NX * NU; d++) blk[(size_t)(L::B + d) * Bp] = Bm.d[d];
#pragma unroll
  for(int d = 0; d < NG * NX; d++) blk[(size_t)(L::C + d) * Bp] = C.d[d];
#pragma unroll
  for(int d = 0; d < NG * NU; d++) blk[(size_t)(L::D + d) * Bp] = D.d[d];
#pragma unroll
  for(int d = 0; d < NX * NX; d++) blk[(size_t)(L::LXX + d) * Bp] = Lxx.d[d];
#pragma unroll
  for(int d = 0; d < NU * NU; d++) blk[(size_t)(L::LUU + d) * Bp] = Luu.d[d];
#pragma unroll
  for(int d = 0; d < NX * NU; d++) blk[(size_t)(L::LXU + d) * Bp] = Lxu.d[d];
#pragma unroll
  for(int d = 0; d < NX; d++) blk[(size_t)(L::XBAR + d) * Bp] = x_bar.d[d];
#pragma unroll
  for(int d = 0; d < NG; d++) blk[(size_t)(L::GBAR + d) * Bp] = g_bar.d[d];
#pragma unroll
  for(int d = 0; d < NX; d++) blk[(size_t)(L::LXBAR + d) * Bp] = Lx_bar.d[d];
#pragma unroll
  for(int d = 0; d < NU; d++) blk[(size_t)(L::LUBAR + d) * Bp] = Lu_bar.d[d];

  // calcKktError terms of this step, added in the reference's order (:506-511)
  S kk = S(0);
  kk += x_bar.squaredNorm();
  kk += g_bar.squaredNorm();
  kk += Lx_bar.squaredNorm();
  kk += Lu_bar.squaredNorm();
  {
    S comp = S(0);
#pragma unroll
    for(int j = 0; j < NG; j++)
    {
      const S v = fmax(s[j] * nu[j] - S(0), S(0));
      comp += v * v;
    }
    kk += comp;
  }
  ws.kkt[(size_t)(i + 1) * Bp + b] = kk;
  if(i == 0)
  {
    S e0 = S(0);
#pragma unroll
    for(int d = 0; d < NX; d++)
    {
      const S e = ws.x0[(size_t)d * Bp + b] - x[d];
      e0 += e * e;
    }
    ws.kkt[b] = e0; // (current_x - x_list[0]).squaredNorm()  (:501)
  }
}

/* ------------------------------------------------------------------------------------ F2 ---- */
/** Barrier update (FmpcSolver.hpp:377-399), KKT test (:443-448) and backwardPass() (:524-665). */
template<class M>
__global__ void fmpc_backward_kernel(const __grid_constant__ M model,
                                     const __grid_constant__ Workspace<typename M::Scalar> ws,
                                     const __grid_constant__ SolverParams<typename M::Scalar> prm,
                                     int iter)
{
  using S = typename M::Scalar;
  constexpr int NX = M::NX, NU = M::NU, NG = M::NG;
  using L = CoeffLayout<NX, NU, NG>;
  using R2 = BackwardRows<NX, NU, NG>;
  constexpr unsigned kFullMask = 0xffffffffu;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ int any_go;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31; // warp 0 computes, warp 1 loads
  S * ring = reinterpret_cast<S *>(smem_raw); // [kDepth][ROWS][32]
  unsigned long long * full = reinterpret_cast<unsigned long long *>(smem_raw + sizeof(S) * (size_t)R2::kDepth * R2::ROWS * kTile);
  unsigned long long * empty = full + R2::kDepth;
  if(threadIdx.x == 0)
  {
    for(int st = 0; st < R2::kDepth; st++)
    {
      ddp::mbarInit(&full[st], 32);
      ddp::mbarInit(&empty[st], 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }

  const int b = blockIdx.x * kTile + lane; // ws.Bp is a multiple of 128: padded lanes read valid memory
  const size_t Bp = ws.Bp;
  const int N = prm.N;
  const S dt = model.dt();
  bool go = (warp == 0) && (b < ws.B) && (ws.status[b < ws.B ? b : 0] == kIterationContinued);
  S barrier_eps = S(0);
  if(go)
  {
  // trace entry of this iteration (:373-375)
  ws.n_trace[b] = iter;
  S * tr = ws.trace + (size_t)(iter - 1) * kTraceFields * Bp + b;
  tr[0] = S(iter);
  tr[2 * Bp] = S(0);
  tr[3 * Bp] = S(0);
  tr[4 * Bp] = S(0);

  // barrier parameter (:378-399): eps = clamp(0.5 * mean(s . nu), 1e-8, 1e6)
  barrier_eps = ws.barrier_eps[b];
  if(prm.update_barrier_eps)
  {
    S s_nu_ave = S(0);
    int total_ineq_dim = 0;
    for(int i = 0; i < N; i++)
    {
      S dotv = S(0);
#pragma unroll
      for(int j = 0; j < NG; j++)
        dotv += ws.s[((size_t)i * NG + j) * Bp + b] * ws.nu[((size_t)i * NG + j) * Bp + b];
      s_nu_ave += dotv;
      total_ineq_dim += ineqDimAt<M>(model, prm.t0 + i * dt); // s_list[i].size() (:388)
    }
    s_nu_ave /= S(total_ineq_dim);
    barrier_eps = fmin(fmax(S(0.5) * s_nu_ave, S(1e-8)), S(1e6));
    ws.barrier_eps[b] = barrier_eps;
  }
  tr[2 * Bp] = barrier_eps;

  // KKT error (:443-448, :496-521)
  {
    S kkt = S(0);
    for(int i = 0; i < N + 2; i++) kkt += ws.kkt[(size_t)i * Bp + b];
    kkt = sqrt(kkt);
    tr[1 * Bp] = kkt;
    if(kkt <= prm.kkt_error_thre)
    {
      ws.status[b] = kSucceeded;
      go = false;
    }
  }
  }
  // does any instance of the tile sweep?  (also publishes the mbarrier initialisation to the loader warp)
  if(warp == 0)
  {
    const bool any = __any_sync(kFullMask, go);
    if(lane == 0) any_go = any ? 1 : 0;
  }
  __syncthreads();
  if(!any_go) return;
  if(warp == 1)
  {
    // loader: rows {coefficient block, s_i, nu_i} of steps N-1, N-2, ..., 0
    const S * row_ptr[R2::ROWS];
    long long row_stride[R2::ROWS];
#pragma unroll
    for(int r = 0; r < R2::ROWS; r++)
    {
      if(r < L::SIZE)
      {
        row_ptr[r] = ws.coeff + ((size_t)(N - 1) * L::SIZE + r) * Bp + b;
        row_stride[r] = -(long long)L::SIZE * (long long)Bp;
      }
      else if(r < L::SIZE + NG)
      {
        row_ptr[r] = ws.s + ((size_t)(N - 1) * NG + (r - L::SIZE)) * Bp + b;
        row_stride[r] = -(long long)NG * (long long)Bp;
      }
      else
      {
        row_ptr[r] = ws.nu + ((size_t)(N - 1) * NG + (r - L::SIZE - NG)) * Bp + b;
        row_stride[r] = -(long long)NG * (long long)Bp;
      }
    }
    loaderLoop<S, R2::ROWS, R2::kDepth>(ring, full, empty, lane, N, row_ptr, row_stride);
    return;
  }

  // backwardPass (:524-665)
  S sv[NX], P[NX * NX];
  S nan_probe = S(0); // becomes NaN as soon as any coefficient is NaN or infinite (x * 0)
  if(go)
  {
#pragma unroll
  for(int d = 0; d < NX; d++)
  {
    sv[d] = S(-1) * ws.term[(size_t)(L::T_LXBAR + d) * Bp + b]; // (2.34)
    ws.sv[((size_t)N * NX + d) * Bp + b] = sv[d];
    nan_probe += sv[d] * S(0) + ws.term[(size_t)(L::T_LX + d) * Bp + b] * S(0);
  }
#pragma unroll
  for(int d = 0; d < NX * NX; d++)
  {
    P[d] = ws.term[(size_t)(L::T_LXX + d) * Bp + b];
    ws.P[((size_t)N * NX * NX + d) * Bp + b] = P[d];
    nan_probe += P[d] * S(0);
  }
  }

  bool llt_failed = false;
  // running store pointers (step N-1 first) instead of per-step 64-bit address arithmetic
  S * kff_ptr = ws.kff + (size_t)(N - 1) * NU * Bp + b;
  S * kfb_ptr = ws.kfb + (size_t)(N - 1) * NU * NX * Bp + b;
  S * sv_ptr = ws.sv + (size_t)(N - 1) * NX * Bp + b;
  S * P_ptr = ws.P + (size_t)(N - 1) * NX * NX * Bp + b;
  for(int i = N - 1; i >= 0; i--, kff_ptr -= (size_t)NU * Bp, kfb_ptr -= (size_t)NU * NX * Bp, sv_ptr -= (size_t)NX * Bp,
          P_ptr -= (size_t)NX * NX * Bp)
  {
    const int st = (N - 1 - i) % R2::kDepth;
    const unsigned use = (unsigned)((N - 1 - i) / R2::kDepth);
    ddp::mbarWait(&full[st], use & 1u); // operands of step i = {coefficient block, s_i, nu_i} have landed
    if(!go)
    {
      ddp::mbarArrive(&empty[st]);
      continue;
    }
    const S * blk = ring + (size_t)st * R2::ROWS * kTile + lane;
    S A[NX * NX], Bm[NX * NU], C[NG * NX], D[NG * NU];
#pragma unroll
    for(int d = 0; d < NX * NX; d++) A[d] = blk[(size_t)(L::A + d) * kTile];
#pragma unroll
    for(int d = 0; d < NX * NU; d++) Bm[d] = blk[(size_t)(L::B + d) * kTile];
#pragma unroll
    for(int d = 0; d < NG * NX; d++) C[d] = blk[(size_t)(L::C + d) * kTile];
#pragma unroll
    for(int d = 0; d < NG * NU; d++) D[d] = blk[(size_t)(L::D + d) * kTile];
    S x_bar[NX], g_bar[NG];
#pragma unroll
    for(int d = 0; d < NX; d++) x_bar[d] = blk[(size_t)(L::XBAR + d) * kTile];
#pragma unroll
    for(int d = 0; d < NG; d++) g_bar[d] = blk[(size_t)(L::GBAR + d) * kTile];

    // pre-process (:572-583)
    S nu_s[NG], tilde_sub[NG];
#pragma unroll
    for(int j = 0; j < NG; j++)
    {
      const S sj = blk[(size_t)(L::SIZE + j) * kTile];
      const S nj = blk[(size_t)(L::SIZE + NG + j) * kTile];
      nu_s[j] = nj / sj;
      tilde_sub[j] = (nu_s[j] * g_bar[j] - nj) + barrier_eps * (S(1) / sj);
    }
    // Qxx~ = dt Lxx + C^T diag(nu/s) C ; Quu~ = dt Luu + D^T diag D ; Qxu~ = dt Lxu + C^T diag D   (2.28c-e)
    S F[NX * NX], H[NX * NU], G[NU * NU];
#pragma unroll
    for(int c = 0; c < NX; c++)
#pragma unroll
      for(int r = 0; r < NX; r++)
      {
        S acc = S(0);
#pragma unroll
        for(int j = 0; j < NG; j++) acc += (C[j + r * NG] * nu_s[j]) * C[j + c * NG];
        F[r + c * NX] = dt * blk[(size_t)(L::LXX + r + c * NX) * kTile] + acc;
      }
#pragma unroll
    for(int c = 0; c < NU; c++)
    {
#pragma unroll
      for(int r = 0; r < NU; r++)
      {
        S acc = S(0);
#pragma unroll
        for(int j = 0; j < NG; j++) acc += (D[j + r * NG] * nu_s[j]) * D[j + c * NG];
        G[r + c * NU] = dt * blk[(size_t)(L::LUU + r + c * NU) * kTile] + acc;
      }
#pragma unroll
      for(int r = 0; r < NX; r++)
      {
        S acc = S(0);
#pragma unroll
        for(int j = 0; j < NG; j++) acc += (C[j + r * NG] * nu_s[j]) * D[j + c * NG];
        H[r + c * NX] = dt * blk[(size_t)(L::LXU + r + c * NX) * kTile] + acc;
      }
    }
    // Lx~ = Lx_bar + C^T tilde_sub ; Lu~ = Lu_bar + D^T tilde_sub                         (2.28f-g)
    S Lx_t[NX], Lu_t[NU];
#pragma unroll
    for(int r = 0; r < NX; r++)
    {
      S acc = S(0);
#pragma unroll
      for(int j = 0; j < NG; j++) acc += C[j + r * NG] * tilde_sub[j];
      Lx_t[r] = blk[(size_t)(L::LXBAR + r) * kTile] + acc;
    }
#pragma unroll
    for(int r = 0; r < NU; r++)
    {
      S acc = S(0);
#pragma unroll
      for(int j = 0; j < NG; j++) acc += D[j + r * NG] * tilde_sub[j];
      Lu_t[r] = blk[(size_t)(L::LUBAR + r) * kTile] + acc;
    }
    // AtP = A^T P ; F += AtP A ; H += AtP B ; BtP = B^T P ; G += BtP B                   (2.35b-d)
    S AtP[NX * NX], BtP[NU * NX];
#pragma unroll
    for(int c = 0; c < NX; c++)
    {
#pragma unroll
      for(int r = 0; r < NX; r++)
      {
        S acc = S(0);
#pragma unroll
        for(int q = 0; q < NX; q++) acc += A[q + r * NX] * P[q + c * NX];
        AtP[r + c * NX] = acc;
      }
#pragma unroll
      for(int r = 0; r < NU; r++)
      {
        S acc = S(0);
#pragma unroll
        for(int q = 0; q < NX; q++) acc += Bm[q + r * NX] * P[q + c * NX];
        BtP[r + c * NU] = acc;
      }
    }
#pragma unroll
    for(int c = 0; c < NX; c++)
#pragma unroll
      for(int r = 0; r < NX; r++)
      {
        S acc = S(0);
#pragma unroll
        for(int q = 0; q < NX; q++) acc += AtP[r + q * NX] * A[q + c * NX];
        F[r + c * NX] += acc;
      }
#pragma unroll
    for(int c = 0; c < NU; c++)
    {
#pragma unroll
      for(int r = 0; r < NX; r++)
      {
        S acc = S(0);
#pragma unroll
        for(int q = 0; q < NX; q++) acc += AtP[r + q * NX] * Bm[q + c * NX];
        H[r + c * NX] += acc;
      }
#pragma unroll
      for(int r = 0; r < NU; r++)
      {
        S acc = S(0);
#pragma unroll
        for(int q = 0; q < NX; q++) acc += BtP[r + q * NU] * Bm[q + c * NX];
        G[r + c * NU] += acc;
      }
    }

    // gain solve (:592-624): k = -G^-1 (B^T (P x_bar - s) + Lu~), K = -G^-1 H^T   (2.35e)
    S Pxb_s[NX];
#pragma unroll
    for(int r = 0; r < NX; r++)
    {
      S acc = S(0);
#pragma unroll
      for(int q = 0; q < NX; q++) acc += P[r + q * NX] * x_bar[q];
      Pxb_s[r] = acc - sv[r];
    }
    S k[NU], K[NU * NX];
    if constexpr(NU == 1)
    {
      // Eigen::LDLT of a 1x1 matrix always reports Success; solve() divides by the pivot unless
      // |pivot| <= DBL_MIN, in which case the pseudo-inverse yields 0
      const S piv = G[0];
      const bool usable = fabs(piv) > S(2.2250738585072014e-308);
      S rhs = S(0);
#pragma unroll
      for(int q = 0; q < NX; q++) rhs += Bm[q] * Pxb_s[q];
      rhs += Lu_t[0];
      k[0] = usable ? S(-1) * (rhs / piv) : S(-0.0);
#pragma unroll
      for(int c = 0; c < NX; c++) K[c] = usable ? S(-1) * (H[c] / piv) : S(-0.0);
    }
    else
    {
      // general NU (:596-617): Eigen::LDLT with diagonal pivoting; if its info() is not Success either give up
      // (break_if_llt_fails) or solve with Eigen::FullPivLU (fmpc_linalg.cuh)
#pragma unroll
      for(int r = 0; r < NU; r++)
      {
        S acc = S(0);
#pragma unroll
        for(int q = 0; q < NX; q++) acc += Bm[q + r * NX] * Pxb_s[q];
        k[r] = acc + Lu_t[r];
      }
#pragma unroll
      for(int c = 0; c < NX; c++)
#pragma unroll
        for(int r = 0; r < NU; r++) K[r + c * NU] = H[c + r * NX]; // H^T
      LdltFactor<S, NU> ldlt;
      ldltCompute<S, NU>(G, ldlt);
      if(ldlt.success)
      {
        ldltSolveInPlace<S, NU>(ldlt, k);
        for(int c = 0; c < NX; c++) ldltSolveInPlace<S, NU>(ldlt, K + c * NU);
      }
      else if(prm.break_if_llt_fails)
      {
        llt_failed = true; // backwardPass() returns false (:608-611): ErrorInBackward
      }
      else
      {
        FullPivLuFactor<S, NU> lu;
        fullPivLuCompute<S, NU>(G, lu);
        fullPivLuSolveInPlace<S, NU>(lu, k);
        for(int c = 0; c < NX; c++) fullPivLuSolveInPlace<S, NU>(lu, K + c * NU);
      }
#pragma unroll
      for(int r = 0; r < NU; r++) k[r] = S(-1) * k[r];
#pragma unroll
      for(int d = 0; d < NU * NX; d++) K[d] = S(-1) * K[d];
    }

    // post-process (:633-637): s = A^T (s - P x_bar) - Lx~ - H k ; P = F - K^T G K, symmetrised   (2.35a)
    S s_new[NX];
#pragma unroll
    for(int r = 0; r < NX; r++)
    {
      S acc = S(0);
#pragma unroll
      for(int q = 0; q < NX; q++) acc += A[q + r * NX] * (S(-1) * Pxb_s[q]);
      S hk = S(0);
#pragma unroll
      for(int c = 0; c < NU; c++) hk += H[r + c * NX] * k[c];
      s_new[r] = (acc - Lx_t[r]) - hk;
    }
    S KtG[NX * NU];
#pragma unroll
    for(int c = 0; c < NU; c++)
#pragma unroll
      for(int r = 0; r < NX; r++)
      {
        S acc = S(0);
#pragma unroll
        for(int q = 0; q < NU; q++) acc += K[q + r * NU] * G[q + c * NU];
        KtG[r + c * NX] = acc;
      }
    S Pn[NX * NX];
#pragma unroll
    for(int c = 0; c < NX; c++)
#pragma unroll
      for(int r = 0; r < NX; r++)
      {
        S acc = S(0);
#pragma unroll
        for(int q = 0; q < NU; q++) acc += KtG[r + q * NX] * K[q + c * NU];
        Pn[r + c * NX] = F[r + c * NX] - acc;
      }
#pragma unroll
    for(int c = 0; c < NX; c++)
#pragma unroll
      for(int r = 0; r < NX; r++) P[r + c * NX] = S(0.5) * (Pn[r + c * NX] + Pn[c + r * NX]);
#pragma unroll
    for(int r = 0; r < NX; r++) sv[r] = s_new[r];

    // save gains (:643-646) and feed the NaN probe with everything containsNaN() inspects (:136-152)
#pragma unroll
    for(int d = 0; d < NU; d++)
    {
      kff_ptr[(size_t)d * Bp] = k[d];
      nan_probe += k[d] * S(0);
    }
#pragma unroll
    for(int d = 0; d < NU * NX; d++)
    {
      kfb_ptr[(size_t)d * Bp] = K[d];
      nan_probe += K[d] * S(0);
    }
#pragma unroll
    for(int d = 0; d < NX; d++)
    {
      sv_ptr[(size_t)d * Bp] = sv[d];
      nan_probe += sv[d] * S(0);
    }
#pragma unroll
    for(int d = 0; d < NX * NX; d++)
    {
      P_ptr[(size_t)d * Bp] = P[d];
      nan_probe += P[d] * S(0) + A[d] * S(0);
    }
#pragma unroll
    for(int d = 0; d < NX * NU; d++) nan_probe += Bm[d] * S(0) + H[d] * S(0);
#pragma unroll
    for(int d = 0; d < NG * NX; d++) nan_probe += C[d] * S(0);
#pragma unroll
    for(int d = 0; d < NG * NU; d++) nan_probe += D[d] * S(0);
#pragma unroll
    for(int d = 0; d < NX; d++) nan_probe += x_bar[d] * S(0) + Lx_t[d] * S(0);
#pragma unroll
    for(int d = 0; d < NG; d++) nan_probe += g_bar[d] * S(0);
#pragma unroll
    for(int d = 0; d < NU; d++) nan_probe += Lu_t[d] * S(0);
    ddp::mbarArrive(&empty[st]); // the loader may refill this stage
  }

  if(go && (llt_failed || (prm.check_nan && !finite(nan_probe))))
  {
    ws.status[b] = kErrorInBackward;
  }
}

/* ------------------------------------------------------------------------------------ F3 ---- */
/** forwardPass() (FmpcSolver.hpp:668-708) and the fraction-to-boundary rule (:714-750). */
template<class M>
__global__ void fmpc_forward_kernel(const __grid_constant__ M model,
                                    const __grid_constant__ Workspace<typename M::Scalar> ws,
                                    const __grid_constant__ SolverParams<typename M::Scalar> prm,
                                    int iter)
{
  using S = typename M::Scalar;
  constexpr int NX = M::NX, NU = M::NU, NG = M::NG;
  using L = CoeffLayout<NX, NU, NG>;
  using R3 = ForwardRows<NX, NU, NG>;
  constexpr unsigned kFullMask = 0xffffffffu;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31; // warp 0 computes, warp 1 loads
  S * ring = reinterpret_cast<S *>(smem_raw); // [kDepth][ROWS][32]
  unsigned long long * full = reinterpret_cast<unsigned long long *>(smem_raw + sizeof(S) * (size_t)R3::kDepth * R3::ROWS * kTile);
  unsigned long long * empty = full + R3::kDepth;
  if(threadIdx.x == 0)
  {
    for(int st = 0; st < R3::kDepth; st++)
    {
      ddp::mbarInit(&full[st], 32);
      ddp::mbarInit(&empty[st], 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }

  const int b = blockIdx.x * kTile + lane; // ws.Bp is a multiple of 128: padded lanes read valid memory
  const bool go = (b < ws.B) && (ws.status[b < ws.B ? b : 0] == kIterationContinued);
  // both warps see the same 32 verdicts: uniform exit; the barrier also publishes the mbarrier initialisation
  if(!__syncthreads_or(go)) return;
  const size_t Bp = ws.Bp;
  const int N = prm.N;
  if(warp == 1)
  {
    // loader: rows of ForwardRows for steps 0, 1, ..., N; step N needs only P_N and s_N (the other rows of that
    // stage repeat step N - 1: their pointers stop advancing there)
    const S * row_ptr[R3::ROWS];
    long long row_stride[R3::ROWS];
#pragma unroll
    for(int r = 0; r < R3::ROWS; r++)
    {
      if(r < R3::SV)
      {
        row_ptr[r] = ws.P + (size_t)(r - R3::P) * Bp + b;
        row_stride[r] = (long long)(NX * NX) * (long long)Bp;
      }
      else if(r < R3::KFB)
      {
        row_ptr[r] = ws.sv + (size_t)(r - R3::SV) * Bp + b;
        row_stride[r] = (long long)NX * (long long)Bp;
      }
      else if(r < R3::KFF)
      {
        row_ptr[r] = ws.kfb + (size_t)(r - R3::KFB) * Bp + b;
        row_stride[r] = (long long)(NU * NX) * (long long)Bp;
      }
      else if(r < R3::ABCD)
      {
        row_ptr[r] = ws.kff + (size_t)(r - R3::KFF) * Bp + b;
        row_stride[r] = (long long)NU * (long long)Bp;
      }
      else if(r < R3::XG)
      {
        row_ptr[r] = ws.coeff + (size_t)(L::A + (r - R3::ABCD)) * Bp + b;
        row_stride[r] = (long long)L::SIZE * (long long)Bp;
      }
      else if(r < R3::S_)
      {
        row_ptr[r] = ws.coeff + (size_t)(L::XBAR + (r - R3::XG)) * Bp + b;
        row_stride[r] = (long long)L::SIZE * (long long)Bp;
      }
      else if(r < R3::NU_)
      {
        row_ptr[r] = ws.s + (size_t)(r - R3::S_) * Bp + b;
        row_stride[r] = (long long)NG * (long long)Bp;
      }
      else
      {
        row_ptr[r] = ws.nu + (size_t)(r - R3::NU_) * Bp + b;
        row_stride[r] = (long long)NG * (long long)Bp;
      }
    }
    loaderLoop<S, R3::ROWS, R3::kDepth>(ring, full, empty, lane, N, row_ptr, row_stride); // steps 0 .. N-1
#pragma unroll
    for(int r = R3::KFB; r < R3::ROWS; r++) row_ptr[r] -= row_stride[r]; // step N: everything but P, s repeats N-1
    {
      const int f = N, st = f % R3::kDepth;
      if(f >= R3::kDepth) ddp::mbarWait(&empty[st], (unsigned)((f / R3::kDepth) - 1) & 1u);
      S * dst = ring + (size_t)st * R3::ROWS * kTile + lane;
#pragma unroll
      for(int r = 0; r < R3::ROWS; r++)
      {
        if constexpr(sizeof(S) == 8)
          ddp::cpAsync8(dst + (size_t)r * kTile, row_ptr[r]);
        else
          ddp::cpAsync4(dst + (size_t)r * kTile, row_ptr[r]);
      }
      cpAsyncArriveOn(&full[st]);
    }
    return;
  }
  const S barrier_eps = go ? ws.barrier_eps[b] : S(0);

  S dx[NX];
#pragma unroll
  for(int d = 0; d < NX; d++) dx[d] = ws.x0[(size_t)d * Bp + b] - ws.x[(size_t)d * Bp + b]; // (:670)

  S nan_probe = S(0);
  S alpha_s_max = S(1), alpha_nu_max = S(1);
  const S margin_ratio = S(0.995);
  // running store pointers instead of per-step 64-bit address arithmetic
  S * dlam_ptr = ws.dlam + b;
  S * dx_ptr = ws.dx + b;
  S * du_ptr = ws.du + b;
  S * ds_ptr = ws.ds + b;
  S * dnu_ptr = ws.dnu + b;
  for(int i = 0; i <= N; i++, dlam_ptr += (size_t)NX * Bp, dx_ptr += (size_t)NX * Bp, du_ptr += (size_t)NU * Bp,
          ds_ptr += (size_t)NG * Bp, dnu_ptr += (size_t)NG * Bp)
  {
    const int st = i % R3::kDepth;
    ddp::mbarWait(&full[st], (unsigned)(i / R3::kDepth) & 1u); // operands of step i (ForwardRows) have landed
    if(!go)
    {
      ddp::mbarArrive(&empty[st]);
      continue;
    }
    const S * op = ring + (size_t)st * R3::ROWS * kTile + lane;
    // dlambda_i = P_i dx_i - s_i                                                    (2.33)
#pragma unroll
    for(int r = 0; r < NX; r++)
    {
      S acc = S(0);
#pragma unroll
      for(int q = 0; q < NX; q++) acc += op[(size_t)(R3::P + r + q * NX) * kTile] * dx[q];
      const S dl = acc - op[(size_t)(R3::SV + r) * kTile];
      dlam_ptr[(size_t)r * Bp] = dl;
      dx_ptr[(size_t)r * Bp] = dx[r];
      nan_probe += dl * S(0) + dx[r] * S(0);
    }
    if(i == N) break;

    // du_i = K_i dx_i + k_i                                                         (2.36)
    S du[NU];
#pragma unroll
    for(int r = 0; r < NU; r++)
    {
      S acc = S(0);
#pragma unroll
      for(int q = 0; q < NX; q++) acc += op[(size_t)(R3::KFB + r + q * NU) * kTile] * dx[q];
      du[r] = acc + op[(size_t)(R3::KFF + r) * kTile];
      du_ptr[(size_t)r * Bp] = du[r];
      nan_probe += du[r] * S(0);
    }
    // ds_i = -(C dx + D du + g_bar) ; dnu_i = -(nu (ds + s) - eps) / s               (2.27a-b)
    const int ng_act = ineqDimAt<M>(model, prm.t0 + i * model.dt());
#pragma unroll
    for(int j = 0; j < NG; j++)
    {
      S cdx = S(0), ddu = S(0);
#pragma unroll
      for(int q = 0; q < NX; q++) cdx += op[(size_t)(R3::ABCD + L::C + j + q * NG) * kTile] * dx[q];
#pragma unroll
      for(int q = 0; q < NU; q++) ddu += op[(size_t)(R3::ABCD + L::D + j + q * NG) * kTile] * du[q];
      const S dsj = S(-1) * ((cdx + ddu) + op[(size_t)(R3::XG + NX + j) * kTile]);
      const S sj = op[(size_t)(R3::S_ + j) * kTile];
      const S nj = op[(size_t)(R3::NU_ + j) * kTile];
      S dnj = S(-1) * (nj * (dsj + sj) - barrier_eps) / sj;
      if constexpr(HasIneqDim<M>::value)
      {
        if(j >= ng_act) dnj = S(0); // padding row: nu stays 0 (delta s is already -(0 + 0 + 0))
      }
      ds_ptr[(size_t)j * Bp] = dsj;
      dnu_ptr[(size_t)j * Bp] = dnj;
      nan_probe += dsj * S(0) + dnj * S(0);
      // fraction-to-boundary (:729-736).  The quotients are formed unconditionally and selected afterwards: same
      // values where they are used, but the 2 x NG divisions of a step become independent instruction streams the
      // scheduler can interleave instead of NG x 2 branches with a dependent division each
      const S cand_s = S(-1) * margin_ratio * sj / dsj;
      const S cand_nu = S(-1) * margin_ratio * nj / dnj;
      alpha_s_max = (dsj < S(0)) ? fmin(alpha_s_max, cand_s) : alpha_s_max;
      alpha_nu_max = (dnj < S(0)) ? fmin(alpha_nu_max, cand_nu) : alpha_nu_max;
    }
    // dx_{i+1} = A dx_i + B du_i + x_bar                                             (2.26b)
    S dxn[NX];
#pragma unroll
    for(int r = 0; r < NX; r++)
    {
      S adx = S(0), bdu = S(0);
#pragma unroll
      for(int q = 0; q < NX; q++) adx += op[(size_t)(R3::ABCD + L::A + r + q * NX) * kTile] * dx[q];
#pragma unroll
      for(int q = 0; q < NU; q++) bdu += op[(size_t)(R3::ABCD + L::B + r + q * NX) * kTile] * du[q];
      dxn[r] = (adx + bdu) + op[(size_t)(R3::XG + r) * kTile];
    }
#pragma unroll
    for(int r = 0; r < NX; r++) dx[r] = dxn[r];
    ddp::mbarArrive(&empty[st]); // the loader may refill this stage
  }

  if(!go) return;
  S * tr = ws.trace + (size_t)(iter - 1) * kTraceFields * Bp + b;
  if(prm.check_nan && !finite(nan_probe))
  {
    ws.status[b] = kErrorInForward; // (:698-705)
    return;
  }
  if(!(alpha_s_max > S(0) && alpha_s_max <= S(1) && alpha_nu_max > S(0) && alpha_nu_max <= S(1)))
  {
    ws.status[b] = kErrorInUpdate; // (:739-747)
    return;
  }
  ws.alpha[b] = alpha_s_max;
  ws.alpha[Bp + b] = alpha_nu_max;
  tr[3 * Bp] = alpha_s_max;
  tr[4 * Bp] = alpha_nu_max;
}

/* ----------------------------------------------------------------------------------- F3b ---- */
/** (A.51) in Nocedal & Wright: directional derivative of |func|_1 along `dir` for one row of the Jacobian
    (MathUtils.h:17-38); `jd` = jac.row(i) . dir. */
template<class S>
__device__ __forceinline__ S l1DirDerivRow(S func, S jd)
{
  if(func > S(0)) return jd;
  if(func < S(0)) return S(-1) * jd;
  return fabs(jd);
}

/** calcMeritFunc() (FmpcSolver.hpp:936-982) at variable + alpha_s * delta (x, u, s only), one thread per instance:
    objective (running cost * dt, log barrier, terminal cost) + merit_const_scale * |constraints|_1. */
template<class M>
__device__ __forceinline__ typename M::Scalar fmpcMeritFunc(const M & model,
                                                            const Workspace<typename M::Scalar> & ws,
                                                            const SolverParams<typename M::Scalar> & prm,
                                                            int b,
                                                            typename M::Scalar alpha_s,
                                                            typename M::Scalar barrier_eps,
                                                            typename M::Scalar merit_const_scale)
{
  using S = typename M::Scalar;
  constexpr int NX = M::NX, NU = M::NU, NG = M::NG;
  const size_t Bp = ws.Bp;
  const int N = prm.N;
  const S dt = model.dt();
  S obj = S(0), con = S(0);
  Matrix<S, NX, 1> x, xn;
#pragma unroll
  for(int d = 0; d < NX; d++)
  {
    const size_t o = (size_t)d * Bp + b;
    x[d] = ws.x[o] + alpha_s * ws.dx[o];
    con += fabs(ws.x0[o] - x[d]);
  }
  for(int i = 0; i < N; i++)
  {
    const S t = prm.t0 + i * dt;
    Matrix<S, NU, 1> u;
    Matrix<S, NG, 1> sv;
#pragma unroll
    for(int d = 0; d < NU; d++)
    {
      const size_t o = ((size_t)i * NU + d) * Bp + b;
      u[d] = ws.u[o] + alpha_s * ws.du[o];
    }
#pragma unroll
    for(int d = 0; d < NG; d++)
    {
      const size_t o = ((size_t)i * NG + d) * Bp + b;
      sv[d] = ws.s[o] + alpha_s * ws.ds[o];
    }
#pragma unroll
    for(int d = 0; d < NX; d++)
    {
      const size_t o = ((size_t)(i + 1) * NX + d) * Bp + b;
      xn[d] = ws.x[o] + alpha_s * ws.dx[o];
    }
    obj += model.runningCost(t, x, u) * dt;
    {
      S log_sum = S(0);
#pragma unroll
      for(int d = 0; d < NG; d++) log_sum += log(sv[d]);
      obj += S(-1) * barrier_eps * log_sum;
    }
    {
      const Matrix<S, NX, 1> f = model.stateEq(t, x, u);
      S l1 = S(0);
#pragma unroll
      for(int d = 0; d < NX; d++) l1 += fabs(f[d] - xn[d]);
      con += l1;
    }
    {
      const Matrix<S, NG, 1> g = model.ineqConst(t, x, u);
      const int ng_act = ineqDimAt<M>(model, t);
      S l1 = S(0);
#pragma unroll
      for(int d = 0; d < NG; d++) l1 += (d < ng_act) ? fabs(g[d] + sv[d]) : S(0);
      con += l1;
    }
    x = xn;
  }
  obj += model.terminalCost(prm.t0 + N * dt, x);
  return obj + merit_const_scale * con;
}

/** The line-search part of updateVariables() (FmpcSolver.hpp:755-793) with setupMeritFunc() (:837-933): Armijo
    backtracking on alpha_s from the fraction-to-boundary value, merit function = objective + scale * l1 norm of
    the constraints.  One thread per instance; overwrites ws.alpha[b] (alpha_s) and the trace entry. */
template<class M>
__global__ void fmpc_linesearch_kernel(const __grid_constant__ M model,
                                       const __grid_constant__ Workspace<typename M::Scalar> ws,
                                       const __grid_constant__ SolverParams<typename M::Scalar> prm,
                                       int iter)
{
  using S = typename M::Scalar;
  constexpr int NX = M::NX, NU = M::NU, NG = M::NG;
  using L = CoeffLayout<NX, NU, NG>;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if(b >= ws.B) return;
  if(ws.status[b] != kIterationContinued) return;
  const size_t Bp = ws.Bp;
  const int N = prm.N;
  const S dt = model.dt();
  const S barrier_eps = ws.barrier_eps[b];

  // ---- setupMeritFunc
  S func_obj = S(0), func_con = S(0), deriv_obj = S(0), deriv_con = S(0);
  Matrix<S, NX, 1> x, xn;
  S dx[NX], dxn[NX];
#pragma unroll
  for(int d = 0; d < NX; d++)
  {
    const size_t o = (size_t)d * Bp + b;
    x[d] = ws.x[o];
    dx[d] = ws.dx[o];
  }
  {
    S l1 = S(0), dd = S(0);
#pragma unroll
    for(int d = 0; d < NX; d++)
    {
      const S cf = ws.x0[(size_t)d * Bp + b] - x[d];
      l1 += fabs(cf);
      dd += l1DirDerivRow<S>(cf, S(-1) * dx[d]); // jacobian -I
    }
    func_con += l1;
    deriv_con += dd;
  }
  for(int i = 0; i < N; i++)
  {
    const S t = prm.t0 + i * dt;
    const S * blk = ws.coeff + (size_t)i * L::SIZE * Bp + b;
    Matrix<S, NU, 1> u;
    Matrix<S, NG, 1> sv;
    S du[NU], ds[NG];
#pragma unroll
    for(int d = 0; d < NU; d++)
    {
      const size_t o = ((size_t)i * NU + d) * Bp + b;
      u[d] = ws.u[o];
      du[d] = ws.du[o];
    }
#pragma unroll
    for(int d = 0; d < NG; d++)
    {
      const size_t o = ((size_t)i * NG + d) * Bp + b;
      sv[d] = ws.s[o];
      ds[d] = ws.ds[o];
    }
#pragma unroll
    for(int d = 0; d < NX; d++)
    {
      const size_t o = ((size_t)(i + 1) * NX + d) * Bp + b;
      xn[d] = ws.x[o];
      dxn[d] = ws.dx[o];
    }
    {
      // coeff.Lx, coeff.Lu are the raw running-cost gradients (FmpcSolver.hpp:421-423)
      Matrix<S, NX, 1> Lx;
      Matrix<S, NU, 1> Lu;
      Matrix<S, NX, NX> Lxx;
      Matrix<S, NU, NU> Luu;
      Matrix<S, NX, NU> Lxu;
      model.calcRunningCostDeriv(t, x, u, Lx, Lu, Lxx, Luu, Lxu);
      func_obj += model.runningCost(t, x, u) * dt;
      S a = S(0), c = S(0);
#pragma unroll
      for(int d = 0; d < NX; d++) a += Lx[d] * dx[d];
#pragma unroll
      for(int d = 0; d < NU; d++) c += Lu[d] * du[d];
      deriv_obj += (a + c) * dt;
    }
    {
      S log_sum = S(0), inv_dot = S(0);
#pragma unroll
      for(int d = 0; d < NG; d++)
      {
        log_sum += log(sv[d]);
        inv_dot += (S(1) / sv[d]) * ds[d];
      }
      func_obj += S(-1) * barrier_eps * log_sum;
      deriv_obj += S(-1) * barrier_eps * inv_dot;
    }
    {
      const Matrix<S, NX, 1> f = model.stateEq(t, x, u);
      S l1 = S(0), da = S(0), db = S(0), dn = S(0);
#pragma unroll
      for(int r = 0; r < NX; r++)
      {
        const S cf = f[r] - xn[r];
        l1 += fabs(cf);
        S ja = S(0), jb = S(0);
#pragma unroll
        for(int q = 0; q < NX; q++) ja += __ldg(blk + (size_t)(L::A + r + q * NX) * Bp) * dx[q];
#pragma unroll
        for(int q = 0; q < NU; q++) jb += __ldg(blk + (size_t)(L::B + r + q * NX) * Bp) * du[q];
        da += l1DirDerivRow<S>(cf, ja);
        db += l1DirDerivRow<S>(cf, jb);
        dn += l1DirDerivRow<S>(cf, S(-1) * dxn[r]);
      }
      func_con += l1;
      deriv_con += da;
      deriv_con += db;
      deriv_con += dn;
    }
    {
      const Matrix<S, NG, 1> g = model.ineqConst(t, x, u);
      S l1 = S(0), dc = S(0), dd = S(0), dsl = S(0);
#pragma unroll
      const int ng_act = ineqDimAt<M>(model, t);
#pragma unroll
      for(int r = 0; r < NG; r++)
      {
        const S cf = (r < ng_act) ? g[r] + sv[r] : S(0); // padding rows: C = D = 0, delta s = 0
        l1 += fabs(cf);
        S jc = S(0), jd = S(0);
#pragma unroll
        for(int q = 0; q < NX; q++) jc += __ldg(blk + (size_t)(L::C + r + q * NG) * Bp) * dx[q];
#pragma unroll
        for(int q = 0; q < NU; q++) jd += __ldg(blk + (size_t)(L::D + r + q * NG) * Bp) * du[q];
        dc += l1DirDerivRow<S>(cf, jc);
        dd += l1DirDerivRow<S>(cf, jd);
        dsl += l1DirDerivRow<S>(cf, ds[r]);
      }
      func_con += l1;
      deriv_con += dc;
      deriv_con += dd;
      deriv_con += dsl;
    }
    x = xn;
#pragma unroll
    for(int d = 0; d < NX; d++) dx[d] = dxn[d];
  }
  {
    func_obj += model.terminalCost(prm.t0 + N * dt, x);
    S a = S(0);
#pragma unroll
    for(int d = 0; d < NX; d++) a += __ldg(ws.term + (size_t)(L::T_LX + d) * Bp + b) * dx[d];
    deriv_obj += a;
  }
  const S scale_min = S(1e-3);
  S scale = scale_min;
  if(prm.merit_const_scale_from_lagrange_multipliers)
  {
    // (18.32) in Nocedal & Wright
    for(int i = 0; i <= N; i++)
    {
#pragma unroll
      for(int d = 0; d < NX; d++) scale = fmax(scale, fabs(ws.lam[((size_t)i * NX + d) * Bp + b]));
      if(i < N)
      {
#pragma unroll
        for(int d = 0; d < NG; d++) scale = fmax(scale, fabs(ws.nu[((size_t)i * NG + d) * Bp + b]));
      }
    }
  }
  else
  {
    // (18.33) in Nocedal & Wright, rho = 0.5
    scale = fmax(deriv_obj / ((S(1) - S(0.5)) * func_con), scale_min);
  }
  const S merit_func = func_obj + scale * func_con;
  const S merit_deriv = deriv_obj + scale * deriv_con;

  // ---- Armijo backtracking (:759-793)
  const S armijo_scale = S(1e-3), update_ratio = S(0.5), alpha_s_min = S(1e-10);
  S alpha_s = ws.alpha[b];
  while(true)
  {
    if(alpha_s < alpha_s_min) break;
    const S merit_new = fmpcMeritFunc<M>(model, ws, prm, b, alpha_s, barrier_eps, scale);
    if(merit_new < merit_func + armijo_scale * alpha_s * merit_deriv) break;
    alpha_s *= update_ratio;
  }
  ws.alpha[b] = alpha_s;
  ws.trace[((size_t)(iter - 1) * kTraceFields + 3) * Bp + b] = alpha_s;
}

/* ------------------------------------------------------------------------------------ F4 ---- */
/** updateVariables() (FmpcSolver.hpp:802-831) for (instance b, step i): x, u, s += alpha_s * delta;
    lambda, nu += alpha_nu * delta.  The reference's clamp against numeric_limits<double>::lowest() is a
    no-op and is not reproduced. */
template<class M>
__global__ void fmpc_update_kernel(const __grid_constant__ Workspace<typename M::Scalar> ws,
                                   const __grid_constant__ SolverParams<typename M::Scalar> prm)
{
  using S = typename M::Scalar;
  constexpr int NX = M::NX, NU = M::NU, NG = M::NG;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if(b >= ws.B) return;
  if(ws.status[b] != kIterationContinued) return;
  const size_t Bp = ws.Bp;
  const S alpha_s = ws.alpha[b];
  const S alpha_nu = ws.alpha[Bp + b];
#pragma unroll
  for(int d = 0; d < NX; d++)
  {
    const size_t o = ((size_t)i * NX + d) * Bp + b;
    ws.x[o] += alpha_s * ws.dx[o];
    ws.lam[o] += alpha_nu * ws.dlam[o];
  }
  if(i < prm.N)
  {
#pragma unroll
    for(int d = 0; d < NU; d++)
    {
      const size_t o = ((size_t)i * NU + d) * Bp + b;
      ws.u[o] += alpha_s * ws.du[o];
    }
#pragma unroll
    for(int d = 0; d < NG; d++)
    {
      const size_t o = ((size_t)i * NG + d) * Bp + b;
      ws.s[o] += alpha_s * ws.ds[o];
      ws.nu[o] += alpha_nu * ws.dnu[o];
    }
  }
}

/** After the last iteration: IterationContinued => MaxIterationReached (FmpcSolver.hpp:241-244). */
static __global__ void fmpc_finalize_kernel(int * status, int B)
{
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if(b < B && status[b] == kIterationContinued) status[b] = kMaxIterationReached;
}
} // namespace fmpc
} // namespace nmpc_b200
