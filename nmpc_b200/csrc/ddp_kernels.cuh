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
 *
 * Device layout (S = scalar type, Bp = padded batch):
 *   x[2]    [N+1][NX][Bp]   current / candidate trajectories; sel[b] says which one is current
 *   u[2]    [N][NU][Bp]
 *   cost[2] [N+1][Bp]
 *   deriv   [N][BLK][Bp]    BLK = {Fx, Fu, Lx, Lu, Lxx, Luu, Lxu} column-major, in that order
 *   vterm   [NX+NX*NX][Bp]  terminal Vx, Vxx
 *   kff     [N][NU][Bp], kfb [N][NU*NX][Bp]
 *   trace   [max_iter+1][9][Bp]
 *   per-instance scalars: lambda, dlambda, cost_sum, dV[2], status, sel, iters, n_fwd, n_bwd
 */
#pragma once

#include <cuda_runtime.h>

#include <nmpc_b200/matrix.h>

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
  S * u_lo; //!< [NU] input limits (with_input_constraint)
  S * u_hi;
  int * status;
  int * sel;
  int * iters;
  int * n_fwd;
  int * n_bwd;
};

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

template<class S>
__device__ __forceinline__ S ldStream(const S * p)
{
  return __ldg(p);
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
  using S = typename M::Scalar;
  constexpr int NX = M::NX, NU = M::NU;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if(b >= ws.B) return;
  const size_t Bp = ws.Bp;
  const int N = prm.N;

  Matrix<S, NX, 1> x;
#pragma unroll
  for(int d = 0; d < NX; d++) x[d] = ws.x[0][(size_t)d * Bp + b];

  S csum = S(0);
  for(int i = 0; i < N; i++)
  {
    Matrix<S, NU, 1> u;
#pragma unroll
    for(int d = 0; d < NU; d++) u[d] = ws.u[0][((size_t)i * NU + d) * Bp + b];
    const S t = prm.t0 + i * model.dt();
    const S c = model.runningCost(t, x, u);
    x = model.stateEq(t, x, u);
#pragma unroll
    for(int d = 0; d < NX; d++) ws.x[0][((size_t)(i + 1) * NX + d) * Bp + b] = x[d];
    ws.cost[0][(size_t)i * Bp + b] = c;
    csum += c;
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

/* ------------------------------------------------------------------------------------ K1 ---- */
/** procOnce() Step 1 (DDPSolver.hpp:157-185): thread (b, i) differentiates dynamics and cost at
    (x_i, u_i) of the current trajectory; i == N evaluates the terminal cost derivatives. */
template<class M>
__global__ void linearize_kernel(const __grid_constant__ M model,
                                 const __grid_constant__ Workspace<typename M::Scalar> ws,
                                 const __grid_constant__ SolverParams<typename M::Scalar> prm)
{
  using S = typename M::Scalar;
  constexpr int NX = M::NX, NU = M::NU;
  using L = BlockLayout<NX, NU>;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
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
  model.calcStateEqDeriv(t, x, u, Fx, Fu);
  model.calcRunningCostDeriv(t, x, u, Lx, Lu, Lxx, Luu, Lxu);

  S * blk = ws.deriv + (size_t)i * L::SIZE * Bp + b;
#pragma unroll
  for(int d = 0; d < NX * NX; d++) blk[(size_t)(L::FX + d) * Bp] = Fx.d[d];
#pragma unroll
  for(int d = 0; d < NX * NU; d++) blk[(size_t)(L::FU + d) * Bp] = Fu.d[d];
#pragma unroll
  for(int d = 0; d < NX; d++) blk[(size_t)(L::LX + d) * Bp] = Lx.d[d];
#pragma unroll
  for(int d = 0; d < NU; d++) blk[(size_t)(L::LU + d) * Bp] = Lu.d[d];
#pragma unroll
  for(int d = 0; d < NX * NX; d++) blk[(size_t)(L::LXX + d) * Bp] = Lxx.d[d];
#pragma unroll
  for(int d = 0; d < NU * NU; d++) blk[(size_t)(L::LUU + d) * Bp] = Luu.d[d];
#pragma unroll
  for(int d = 0; d < NX * NU; d++) blk[(size_t)(L::LXU + d) * Bp] = Lxu.d[d];
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

/** One backwardPass() sweep (DDPSolver.hpp:343-534) with regularisation `lambda`.  Returns false as
    soon as the Cholesky factorisation of Quu_F fails at some step (LLT NumericalIssue, :500-508). */
template<class M>
__device__ __forceinline__ bool backwardSweep(const Workspace<typename M::Scalar> & ws,
                                              const SolverParams<typename M::Scalar> & prm,
                                              int b,
                                              const typename M::Scalar * __restrict__ us,
                                              typename M::Scalar lambda,
                                              typename M::Scalar & dV0,
                                              typename M::Scalar & dV1,
                                              typename M::Scalar & k_rel_norm)
{
  using S = typename M::Scalar;
  constexpr int NX = M::NX, NU = M::NU;
  using L = BlockLayout<NX, NU>;
  const size_t Bp = ws.Bp;
  const int N = prm.N;

  S Vx[NX], Vxx[NX * NX];
#pragma unroll
  for(int d = 0; d < NX; d++) Vx[d] = ws.vterm[(size_t)d * Bp + b];
#pragma unroll
  for(int d = 0; d < NX * NX; d++) Vxx[d] = ws.vterm[(size_t)(NX + d) * Bp + b];

  dV0 = S(0);
  dV1 = S(0);
  k_rel_norm = S(0);

  for(int i = N - 1; i >= 0; i--)
  {
    const S * blk = ws.deriv + (size_t)i * L::SIZE * Bp + b;
    S Fx[NX * NX], Fu[NX * NU];
#pragma unroll
    for(int d = 0; d < NX * NX; d++) Fx[d] = ldStream(blk + (size_t)(L::FX + d) * Bp);
#pragma unroll
    for(int d = 0; d < NX * NU; d++) Fu[d] = ldStream(blk + (size_t)(L::FU + d) * Bp);

    // Qu = Lu + Fu^T Vx ; Qx = Lx + Fx^T Vx                                  (:386-388)
    S Qu[NU], Qx[NX];
#pragma unroll
    for(int a = 0; a < NU; a++)
    {
      S s = S(0);
#pragma unroll
      for(int r = 0; r < NX; r++) s += Fu[r + a * NX] * Vx[r];
      Qu[a] = ldStream(blk + (size_t)(L::LU + a) * Bp) + s;
    }
#pragma unroll
    for(int j = 0; j < NX; j++)
    {
      S s = S(0);
#pragma unroll
      for(int r = 0; r < NX; r++) s += Fx[r + j * NX] * Vx[r];
      Qx[j] = ldStream(blk + (size_t)(L::LX + j) * Bp) + s;
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
        Qux[a + j * NU] = ldStream(blk + (size_t)(L::LXU + j + a * NX) * Bp) + s;
      }
#pragma unroll
      for(int c = 0; c < NX; c++)
      {
        S s = S(0);
#pragma unroll
        for(int r = 0; r < NX; r++) s += Tx[c + r * NX] * Fx[r + j * NX];
        Qxx[c + j * NX] = ldStream(blk + (size_t)(L::LXX + c + j * NX) * Bp) + s;
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
        Quu[a + c * NU] = ldStream(blk + (size_t)(L::LUU + a + c * NU) * Bp) + s;
      }

    // regularisation (:421-441)
    S Qux_reg[NU * NX], Quu_F[NU * NU];
    if(prm.reg_type == 2)
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
          Qux_reg[a + j * NU] = ldStream(blk + (size_t)(L::LXU + j + a * NX) * Bp) + s;
        }
#pragma unroll
      for(int c = 0; c < NU; c++)
#pragma unroll
        for(int a = 0; a < NU; a++)
        {
          S s = S(0);
#pragma unroll
          for(int r = 0; r < NX; r++) s += Tur[a + r * NU] * Fu[r + c * NX];
          Quu_F[a + c * NU] = ldStream(blk + (size_t)(L::LUU + a + c * NU) * Bp) + s;
        }
    }
    else
    {
#pragma unroll
      for(int d = 0; d < NU * NX; d++) Qux_reg[d] = Qux[d];
#pragma unroll
      for(int d = 0; d < NU * NU; d++) Quu_F[d] = Quu[d];
      if(prm.reg_type == 1)
      {
#pragma unroll
        for(int a = 0; a < NU; a++) Quu_F[a + a * NU] += lambda;
      }
    }

    // gains: LLT(Quu_F), k = -Quu_F^-1 Qu, K = -Quu_F^-1 Qux_reg           (:500-510)
    if(!lltInPlace<S, NU>(Quu_F)) return false;
    S invd[NU];
#pragma unroll
    for(int a = 0; a < NU; a++) invd[a] = S(1) / Quu_F[a + a * NU];
    S k[NU], K[NU * NX];
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
      dV0 += s0;
      dV1 += S(0.5) * s1;
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

    // save gains (:529-530) and accumulate max_i |k_i| / (|u_i| + 1) (:217-221)
    S kn = S(0), un = S(0);
#pragma unroll
    for(int a = 0; a < NU; a++)
    {
      ws.kff[((size_t)i * NU + a) * Bp + b] = k[a];
      kn += k[a] * k[a];
      const S uv = us[((size_t)i * NU + a) * Bp + b];
      un += uv * uv;
    }
#pragma unroll
    for(int d = 0; d < NU * NX; d++) ws.kfb[((size_t)i * NU * NX + d) * Bp + b] = K[d];
    k_rel_norm = fmax(k_rel_norm, sqrt(kn) / (sqrt(un) + S(1)));
  }
  return true;
}

/** procOnce() Step 2 (DDPSolver.hpp:188-231): retry the backward sweep with larger lambda until the
    factorisation succeeds, then the small-gradient termination test. */
template<class M>
__global__ void backward_kernel(const __grid_constant__ M model,
                                const __grid_constant__ Workspace<typename M::Scalar> ws,
                                const __grid_constant__ SolverParams<typename M::Scalar> prm,
                                int iter)
{
  using S = typename M::Scalar;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if(b >= ws.B) return;
  if(ws.status[b] != 0) return;

  S lambda = ws.lambda[b];
  S dlambda = ws.dlambda[b];
  const S * us = ws.u[ws.sel[b]];
  int n_bwd = ws.n_bwd[b];
  S dV0, dV1, k_rel_norm;
  bool failed = false;
  for(;;)
  {
    n_bwd++;
    if(backwardSweep<M>(ws, prm, b, us, lambda, dV0, dV1, k_rel_norm)) break;
    // increase lambda (:194-204)
    dlambda = fmax(dlambda * prm.lambda_factor, prm.lambda_factor);
    lambda = fmax(lambda * dlambda, prm.lambda_min);
    if(lambda > prm.lambda_max)
    {
      failed = true;
      break;
    }
  }
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
/** procOnce() Steps 3-4 (DDPSolver.hpp:234-339): backtracking line search over alpha_list with
    forwardPass(alpha) (:537-560) writing the candidate trajectory into the non-current buffer; on
    success the buffers swap roles (sel ^= 1) instead of the reference's three copies (:285-287). */
template<class M>
__global__ void forward_kernel(const __grid_constant__ M model,
                               const __grid_constant__ Workspace<typename M::Scalar> ws,
                               const __grid_constant__ SolverParams<typename M::Scalar> prm,
                               int iter)
{
  using S = typename M::Scalar;
  constexpr int NX = M::NX, NU = M::NU;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if(b >= ws.B) return;
  if(ws.status[b] != 0) return;
  const size_t Bp = ws.Bp;
  const int N = prm.N;

  const int sel = ws.sel[b];
  const S * __restrict__ xc = ws.x[sel];
  const S * __restrict__ uc = ws.u[sel];
  S * __restrict__ xn = ws.x[sel ^ 1];
  S * __restrict__ un = ws.u[sel ^ 1];
  S * __restrict__ cn = ws.cost[sel ^ 1];

  const S cost_cur = ws.cost_sum[b];
  const S dV0 = ws.dV[b];
  const S dV1 = ws.dV[Bp + b];
  S lambda = ws.lambda[b];
  S dlambda = ws.dlambda[b];
  const S k_rel_norm = ws.trace[((size_t)iter * kTraceFields + 5) * Bp + b];

  bool forward_pass_success = false;
  S alpha = S(0), cost_update_actual = S(0), cost_update_expected = S(0), cost_update_ratio = S(0);
  S cost_new = S(0);
  int n_fwd = ws.n_fwd[b];

  Matrix<S, NX, 1> x0;
#pragma unroll
  for(int d = 0; d < NX; d++) x0[d] = xc[(size_t)d * Bp + b];
#pragma unroll
  for(int d = 0; d < NX; d++) xn[(size_t)d * Bp + b] = x0[d]; // candidate x_list[0] (:540)

  for(int ai = 0; ai < prm.n_alpha; ai++)
  {
    alpha = prm.alpha_list[ai];
    n_fwd++;

    // forwardPass(alpha)
    Matrix<S, NX, 1> x = x0;
    S csum = S(0);
    // software prefetch of step i+1's operands while step i computes
    S xr[NX], ur[NU], kr[NU], Kr[NU * NX];
#pragma unroll
    for(int d = 0; d < NX; d++) xr[d] = x0[d];
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
        un[((size_t)i * NU + a) * Bp + b] = u[a];
      }
      const S t = prm.t0 + i * model.dt();
      const S c = model.runningCost(t, x, u);
      x = model.stateEq(t, x, u);
#pragma unroll
      for(int d = 0; d < NX; d++) xn[((size_t)(i + 1) * NX + d) * Bp + b] = x[d];
      cn[(size_t)i * Bp + b] = c;
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
      cn[(size_t)N * Bp + b] = c;
      csum += c;
    }

    // (:248-264)
    cost_new = csum;
    cost_update_actual = cost_cur - csum;
    cost_update_expected = S(-1) * alpha * (dV0 + alpha * dV1);
    cost_update_ratio = cost_update_actual / cost_update_expected;
    if(cost_update_expected < S(0))
    {
      cost_update_ratio = (cost_update_actual >= S(0)) ? S(1) : S(-1);
    }
    if(cost_update_ratio > prm.cost_update_ratio_thre)
    {
      forward_pass_success = true;
      break;
    }
  }

  // Step 4 (:280-333)
  int retval = 0;
  S cost_out = cost_cur;
  if(forward_pass_success)
  {
    ws.sel[b] = sel ^ 1;
    ws.cost_sum[b] = cost_new;
    cost_out = cost_new;
    if(cost_update_actual < prm.cost_update_thre) retval = 1;
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
  ws.n_fwd[b] = n_fwd;
  ws.iters[b] = iter;
  if(retval != 0) ws.status[b] = retval;
  writeTrace<S>(ws, b, iter, S(iter), cost_out, lambda, dlambda, alpha, k_rel_norm, cost_update_actual,
                cost_update_expected, cost_update_ratio);
}
} // namespace ddp
} // namespace nmpc_b200
