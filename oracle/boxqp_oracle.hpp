// TEST INFRASTRUCTURE ONLY -- CPU oracle for nmpc_b200 (see oracle/README.md).
//
// Restatement of nmpc_ddp::BoxQP<VarDim>::solve (projected-Newton box QP, Tassa 2014):
//   /root/reference/nmpc_ddp/include/nmpc_ddp/BoxQP.h:141-347  (algorithm)
//   /root/reference/nmpc_ddp/include/nmpc_ddp/BoxQP.h:33-55    (defaults)
//   /root/reference/nmpc_ddp/include/nmpc_ddp/BoxQP.h:375-383  (exit codes)
// Pinned by the reference's own known-answer tests, nmpc_ddp/tests/src/TestBoxQP.cpp:39-55
// (tests/test_oracle_boxqp.py).
#pragma once

#include <cmath>

#include "linalg.hpp"

namespace oracle
{
template<int N>
struct BoxQP
{
  struct Configuration // BoxQP.h:33-55
  {
    int max_iter = 500;
    double grad_thre = 1e-8;
    double rel_improve_thre = 1e-8;
    double step_factor = 0.6;
    double min_step = 1e-22;
    double armijo_param = 0.1;
  };

  Configuration config;
  int retval = 0; // BoxQP.h:372
  int iter = 0;
  int factorization_num = 0;
  int n_free = 0;
  int free_idxs[N > 0 ? N : 1]; // BoxQP.h:389
  double llt_free[(N > 0 ? N : 1) * (N > 0 ? N : 1)]; // BoxQP.h:386, lower factor of H_free (ld = n_free)

  static double objective(const Mat<N, N> & H, const Vec<N> & g, const Vec<N> & x)
  {
    // x.dot(g) + 0.5 * x.dot(H * x)            (BoxQP.h:149,297,303)
    Vec<N> Hx = mul(H, x);
    return dot(x, g) + 0.5 * dot(x, Hx);
  }

  static Vec<N> clampVec(const Vec<N> & x, const Vec<N> & lower, const Vec<N> & upper)
  {
    // .cwiseMin(upper).cwiseMax(lower)         (BoxQP.h:148,296,302)
    Vec<N> y;
    for(int i = 0; i < N; i++) y[i] = std::fmax(std::fmin(x[i], upper[i]), lower[i]);
    return y;
  }

  Vec<N> solve(const Mat<N, N> & H,
               const Vec<N> & g,
               const Vec<N> & lower,
               const Vec<N> & upper,
               const Vec<N> & initial_x)
  {
    Vec<N> x = clampVec(initial_x, lower, upper); // :148
    double obj = objective(H, g, x); // :149
    double old_obj = obj; // :150

    retval = 0; // :161
    factorization_num = 0;
    Vec<N> grad = Vec<N>::Zero();
    bool clamped_flag[N], old_clamped_flag[N];
    for(int i = 0; i < N; i++) clamped_flag[i] = old_clamped_flag[i] = false;
    n_free = 0;
    iter = 1;
    for(;; iter++) // :168
    {
      // relative improvement, checked from the second pass on (:176-181)
      if(iter > 1 && (old_obj - obj) < config.rel_improve_thre * std::fabs(old_obj))
      {
        retval = 4;
        break;
      }
      old_obj = obj;

      // gradient (:184)
      Vec<N> Hx = mul(H, x);
      for(int i = 0; i < N; i++) grad[i] = g[i] + Hx[i];

      // clamped set: exact equality with the bound and gradient pointing outwards (:187-191)
      for(int i = 0; i < N; i++) old_clamped_flag[i] = clamped_flag[i];
      bool all_clamped = true;
      int clamped_idxs[N];
      int n_clamped = 0;
      n_free = 0;
      for(int i = 0; i < N; i++)
      {
        clamped_flag[i] = (x[i] == lower[i] && grad[i] > 0) || (x[i] == upper[i] && grad[i] < 0);
        if(clamped_flag[i])
          clamped_idxs[n_clamped++] = i;
        else
        {
          free_idxs[n_free++] = i;
          all_clamped = false;
        }
      }
      if(all_clamped) // :209-213
      {
        retval = 6;
        break;
      }

      // refactorise only when the clamped set changed (:216-241)
      bool changed = false;
      for(int i = 0; i < N; i++) changed = changed || (clamped_flag[i] != old_clamped_flag[i]);
      if(iter == 1 || changed)
      {
        for(int i = 0; i < n_free; i++)
          for(int j = 0; j < n_free; j++) llt_free[i + j * n_free] = H(free_idxs[i], free_idxs[j]);
        if(!lltInPlace(llt_free, n_free, n_free))
        {
          retval = -1;
          break;
        }
        factorization_num++;
      }

      // free-gradient norm (:244-253)
      double grad_norm = 0;
      for(int i = 0; i < n_free; i++) grad_norm += grad[free_idxs[i]] * grad[free_idxs[i]];
      if(grad_norm < config.grad_thre * config.grad_thre)
      {
        retval = 5;
        break;
      }

      // search direction (:256-279)
      double rhs[N];
      for(int i = 0; i < n_free; i++)
      {
        double s = 0.0;
        for(int j = 0; j < n_clamped; j++) s += H(free_idxs[i], clamped_idxs[j]) * x[clamped_idxs[j]];
        rhs[i] = g[free_idxs[i]] + s;
      }
      lltSolveInPlace(llt_free, n_free, n_free, rhs);
      Vec<N> search_dir = Vec<N>::Zero();
      for(int i = 0; i < n_free; i++) search_dir[free_idxs[i]] = -1 * rhs[i] - x[free_idxs[i]];

      // descent check (:282-291)
      double search_dir_grad = dot(search_dir, grad);
      if(search_dir_grad > 1e-10)
      {
        retval = -2;
        break;
      }

      // Armijo line search (:294-309); retval 2 only leaves the inner loop
      double step = 1;
      Vec<N> x_candidate;
      for(int i = 0; i < N; i++) x_candidate[i] = x[i] + step * search_dir[i];
      x_candidate = clampVec(x_candidate, lower, upper);
      double obj_candidate = objective(H, g, x_candidate);
      while((obj_candidate - old_obj) / (step * search_dir_grad) < config.armijo_param)
      {
        step = step * config.step_factor;
        for(int i = 0; i < N; i++) x_candidate[i] = x[i] + step * search_dir[i];
        x_candidate = clampVec(x_candidate, lower, upper);
        obj_candidate = objective(H, g, x_candidate);
        if(step < config.min_step)
        {
          retval = 2;
          break;
        }
      }

      // accept (:328-329)
      x = x_candidate;
      obj = obj_candidate;

      if(iter == config.max_iter) // :332-336
      {
        retval = 1;
        break;
      }
    }
    return x;
  }
};
} // namespace oracle
