/* nmpc_b200 -- projected-Newton box-constrained QP for the control-limited backward pass.
 *
 * Device restatement of nmpc_ddp::BoxQP<VarDim>::solve (isri-aist/NMPC
 * nmpc_ddp/include/nmpc_ddp/BoxQP.h:141-347; Tassa, Mansard, Todorov, ICRA 2014) with the reference's
 * default configuration (BoxQP.h:33-55; DDPSolver constructs a fresh default BoxQP per step,
 * DDPSolver.hpp:469): max_iter 500, grad_thre 1e-8, rel_improve_thre 1e-8, step_factor 0.6,
 * min_step 1e-22, armijo_param 0.1.  One thread solves one QP; N (= n_u) is a small compile-time
 * constant, the free/clamped sets are bit masks.
 */
#pragma once

#include <cuda_runtime.h>

namespace nmpc_b200
{
namespace ddp
{
template<class S, int N>
struct BoxQPResult
{
  S x[N > 0 ? N : 1];
  S llt_free[(N > 0 ? N : 1) * (N > 0 ? N : 1)]; //!< lower Cholesky factor of H(free, free), leading dimension n_free
  int free_idxs[N > 0 ? N : 1];
  int n_free;
  int retval; //!< BoxQP.h:375-383: <0 failure; 4, 5, 6 regular exits; 1, 2 iteration limits
};

template<class S, int N>
__device__ __forceinline__ S boxQpObjective(const S * H, const S * g, const S * x)
{
  // x.dot(g) + 0.5 * x.dot(H * x)   (BoxQP.h:149)
  S xg = S(0), xHx = S(0);
#pragma unroll
  for(int i = 0; i < N; i++)
  {
    S hx = S(0);
#pragma unroll
    for(int j = 0; j < N; j++) hx += H[i + j * N] * x[j];
    xg += x[i] * g[i];
    xHx += x[i] * hx;
  }
  return xg + S(0.5) * xHx;
}

/** H (column-major N x N), g, lower, upper, initial_x -> res.  Mirrors BoxQP.h:141-347 statement by statement. */
template<class S, int N>
__device__ void boxQpSolve(const S * H, const S * g, const S * lower, const S * upper, const S * initial_x,
                           BoxQPResult<S, N> & res)
{
  const int max_iter = 500;
  const S grad_thre = S(1e-8);
  const S rel_improve_thre = S(1e-8);
  const S step_factor = S(0.6);
  const S min_step = S(1e-22);
  const S armijo_param = S(0.1);

  S x[N];
#pragma unroll
  for(int i = 0; i < N; i++) x[i] = fmax(fmin(initial_x[i], upper[i]), lower[i]); // :148
  S obj = boxQpObjective<S, N>(H, g, x);
  S old_obj = obj;

  res.retval = 0;
  res.n_free = 0;
  unsigned clamped = 0u, old_clamped = 0u;
  S grad[N];
  for(int iter = 1;; iter++)
  {
    // relative improvement (:176-181)
    if(iter > 1 && (old_obj - obj) < rel_improve_thre * fabs(old_obj))
    {
      res.retval = 4;
      break;
    }
    old_obj = obj;

    // gradient and clamped set (:184-191): exact equality with the bound, gradient pointing outwards
    old_clamped = clamped;
    clamped = 0u;
#pragma unroll
    for(int i = 0; i < N; i++)
    {
      S hx = S(0);
#pragma unroll
      for(int j = 0; j < N; j++) hx += H[i + j * N] * x[j];
      grad[i] = g[i] + hx;
      if((x[i] == lower[i] && grad[i] > S(0)) || (x[i] == upper[i] && grad[i] < S(0))) clamped |= (1u << i);
    }
    res.n_free = 0;
#pragma unroll
    for(int i = 0; i < N; i++)
      if(!((clamped >> i) & 1u)) res.free_idxs[res.n_free++] = i;
    if(res.n_free == 0) // all clamped (:209-213)
    {
      res.retval = 6;
      break;
    }

    // factorise H(free, free) when the clamped set changed (:216-241); LLT failure rule: pivot <= 0
    if(iter == 1 || clamped != old_clamped)
    {
      const int nf = res.n_free;
      for(int c = 0; c < nf; c++)
        for(int r = 0; r < nf; r++) res.llt_free[r + c * nf] = H[res.free_idxs[r] + res.free_idxs[c] * N];
      bool ok = true;
      for(int k = 0; k < nf; k++)
      {
        S d = res.llt_free[k + k * nf];
        for(int j = 0; j < k; j++) d -= res.llt_free[k + j * nf] * res.llt_free[k + j * nf];
        if(d <= S(0))
        {
          ok = false;
          break;
        }
        d = sqrt(d);
        res.llt_free[k + k * nf] = d;
        for(int r = k + 1; r < nf; r++)
        {
          S s = res.llt_free[r + k * nf];
          for(int j = 0; j < k; j++) s -= res.llt_free[r + j * nf] * res.llt_free[k + j * nf];
          res.llt_free[r + k * nf] = s / d;
        }
      }
      if(!ok)
      {
        res.retval = -1;
        break;
      }
    }

    // free-gradient norm (:244-253)
    S grad_norm = S(0);
    for(int i = 0; i < res.n_free; i++) grad_norm += grad[res.free_idxs[i]] * grad[res.free_idxs[i]];
    if(grad_norm < grad_thre * grad_thre)
    {
      res.retval = 5;
      break;
    }

    // search direction (:256-279): -H_ff^-1 (g_f + H_fc x_c) - x_f on the free set, 0 elsewhere
    S rhs[N];
    {
      const int nf = res.n_free;
      for(int i = 0; i < nf; i++)
      {
        S s = S(0);
#pragma unroll
        for(int j = 0; j < N; j++)
          if((clamped >> j) & 1u) s += H[res.free_idxs[i] + j * N] * x[j];
        rhs[i] = g[res.free_idxs[i]] + s;
      }
      for(int i = 0; i < nf; i++)
      {
        S s = rhs[i];
        for(int j = 0; j < i; j++) s -= res.llt_free[i + j * nf] * rhs[j];
        rhs[i] = s / res.llt_free[i + i * nf];
      }
      for(int i = nf - 1; i >= 0; i--)
      {
        S s = rhs[i];
        for(int j = i + 1; j < nf; j++) s -= res.llt_free[j + i * nf] * rhs[j];
        rhs[i] = s / res.llt_free[i + i * nf];
      }
    }
    S search_dir[N];
#pragma unroll
    for(int i = 0; i < N; i++) search_dir[i] = S(0);
    for(int i = 0; i < res.n_free; i++) search_dir[res.free_idxs[i]] = S(-1) * rhs[i] - x[res.free_idxs[i]];

    // descent check (:282-291)
    S search_dir_grad = S(0);
#pragma unroll
    for(int i = 0; i < N; i++) search_dir_grad += search_dir[i] * grad[i];
    if(search_dir_grad > S(1e-10))
    {
      res.retval = -2;
      break;
    }

    // Armijo line search (:294-309); retval 2 leaves only the inner loop
    S step = S(1);
    S x_candidate[N];
#pragma unroll
    for(int i = 0; i < N; i++) x_candidate[i] = fmax(fmin(x[i] + step * search_dir[i], upper[i]), lower[i]);
    S obj_candidate = boxQpObjective<S, N>(H, g, x_candidate);
    while((obj_candidate - old_obj) / (step * search_dir_grad) < armijo_param)
    {
      step = step * step_factor;
#pragma unroll
      for(int i = 0; i < N; i++) x_candidate[i] = fmax(fmin(x[i] + step * search_dir[i], upper[i]), lower[i]);
      obj_candidate = boxQpObjective<S, N>(H, g, x_candidate);
      if(step < min_step)
      {
        res.retval = 2;
        break;
      }
    }

    // accept (:328-329)
#pragma unroll
    for(int i = 0; i < N; i++) x[i] = x_candidate[i];
    obj = obj_candidate;

    if(iter == max_iter) // :332-336
    {
      res.retval = 1;
      break;
    }
  }
#pragma unroll
  for(int i = 0; i < N; i++) res.x[i] = x[i];
}
} // namespace ddp
} // namespace nmpc_b200
