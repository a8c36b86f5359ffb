/* nmpc_b200 -- the two dense factorisations FmpcSolver::backwardPass applies to G (n_u x n_u), restated for one
 * thread (FmpcSolver.hpp:592-624):
 *
 *   Eigen::LDLT<InputInputDimMatrix> llt_G(G);   if(llt_G.info() == Eigen::Success) k, K = -llt_G.solve(...)
 *   else if(break_if_llt_fails) return false;    else Eigen::FullPivLU<InputInputDimMatrix> lu_G(G); k, K = -lu_G.solve(...)
 *
 * Semantics follow Eigen 3.4 (the reference pins no version; 3.3.7 / 3.4.0 in its CI agree on these rules):
 *   LDLT       diagonal pivoting (largest |diagonal| of the trailing block first), unblocked lower in-place update;
 *              info() is NumericalIssue only when a zero pivot has a non-zero column below it or a non-zero pivot
 *              follows a zero one; a 1 x 1 matrix is always Success.  solve() uses the pseudo-inverse of D
 *              (|d| <= DBL_MIN => 0).
 *   FullPivLU  complete pivoting (first maximum of |a| in column-major order of the trailing block), in-place
 *              elimination; solve() keeps the leading `rank` pivots, rank = #{|pivot| > eps * n * max|pivot|}, and
 *              sets the other solution components to zero.
 * Pinned by tests/test_fmpc_extra.py against golden vectors of the reference's own FmpcSolver<6, 2, 4> (both the
 * LDLT and the FullPivLU branch).  n is a small compile-time constant; the pivot
 * permutations make the arrays dynamically indexed (local memory) -- only problems with n_u > 1 pay for it.
 */
#pragma once

#include <cuda_runtime.h>

namespace nmpc_b200
{
namespace fmpc
{
template<class S>
__device__ __forceinline__ S tinyPivot()
{
  return sizeof(S) == 8 ? S(2.2250738585072014e-308) : S(1.17549435e-38f);
}
template<class S>
__device__ __forceinline__ S machineEps()
{
  return sizeof(S) == 8 ? S(2.220446049250313e-16) : S(1.1920929e-7f);
}

template<class S, int n>
struct LdltFactor
{
  S a[n * n]; //!< unit lower factor below the diagonal, D on it (of the permuted matrix)
  int perm[n];
  bool success;
};

template<class S, int n>
__device__ inline void ldltCompute(const S * G, LdltFactor<S, n> & f)
{
  for(int j = 0; j < n; j++)
    for(int i = 0; i < n; i++) f.a[i + j * n] = G[i + j * n];
  for(int i = 0; i < n; i++) f.perm[i] = i;
  f.success = true;
  if(n <= 1) return;
  bool found_zero_pivot = false;
  for(int k = 0; k < n; k++)
  {
    int piv = k;
    S best = fabs(f.a[k + k * n]);
    for(int i = k + 1; i < n; i++)
    {
      const S v = fabs(f.a[i + i * n]);
      if(v > best)
      {
        best = v;
        piv = i;
      }
    }
    if(piv != k)
    {
      // symmetric row / column swap of the symmetric working copy (Eigen swaps inside the lower triangle: same result)
      for(int j = 0; j < n; j++)
      {
        const S t = f.a[k + j * n];
        f.a[k + j * n] = f.a[piv + j * n];
        f.a[piv + j * n] = t;
      }
      for(int i = 0; i < n; i++)
      {
        const S t = f.a[i + k * n];
        f.a[i + k * n] = f.a[i + piv * n];
        f.a[i + piv * n] = t;
      }
      const int t = f.perm[k];
      f.perm[k] = f.perm[piv];
      f.perm[piv] = t;
    }
    // A(k,k) -= A10 (D0 A10^T);  A21 -= A20 (D0 A10^T);  A21 /= A(k,k)
    S dk = f.a[k + k * n];
    for(int j = 0; j < k; j++) dk -= f.a[k + j * n] * (f.a[j + j * n] * f.a[k + j * n]);
    f.a[k + k * n] = dk;
    for(int i = k + 1; i < n; i++)
    {
      S s = f.a[i + k * n];
      for(int j = 0; j < k; j++) s -= f.a[i + j * n] * (f.a[j + j * n] * f.a[k + j * n]);
      f.a[i + k * n] = s;
    }
    const bool pivot_is_valid = fabs(dk) > S(0);
    if(pivot_is_valid)
    {
      for(int i = k + 1; i < n; i++) f.a[i + k * n] /= dk;
    }
    else
    {
      for(int i = k + 1; i < n; i++) f.success = f.success && (f.a[i + k * n] == S(0));
    }
    if(found_zero_pivot && pivot_is_valid)
      f.success = false;
    else if(!pivot_is_valid)
      found_zero_pivot = true;
  }
}

/** b <- A^-1 b (Eigen::LDLT::solve). */
template<class S, int n>
__device__ inline void ldltSolveInPlace(const LdltFactor<S, n> & f, S * b)
{
  S y[n];
  for(int i = 0; i < n; i++) y[i] = b[f.perm[i]];
  for(int i = 0; i < n; i++)
    for(int j = 0; j < i; j++) y[i] -= f.a[i + j * n] * y[j];
  const S tol = tinyPivot<S>();
  for(int i = 0; i < n; i++)
  {
    const S dk = f.a[i + i * n];
    y[i] = (fabs(dk) > tol) ? y[i] / dk : S(0);
  }
  for(int i = n - 1; i >= 0; i--)
    for(int j = i + 1; j < n; j++) y[i] -= f.a[j + i * n] * y[j];
  for(int i = 0; i < n; i++) b[f.perm[i]] = y[i];
}

template<class S, int n>
struct FullPivLuFactor
{
  S a[n * n]; //!< L (unit, below the diagonal) and U of P A Q
  int row_tr[n]; //!< row transposition applied at step k
  int col_perm[n]; //!< column of A that ended up at position i
  int rank;
};

template<class S, int n>
__device__ inline void fullPivLuCompute(const S * G, FullPivLuFactor<S, n> & f)
{
  for(int j = 0; j < n; j++)
    for(int i = 0; i < n; i++) f.a[i + j * n] = G[i + j * n];
  for(int i = 0; i < n; i++)
  {
    f.row_tr[i] = i;
    f.col_perm[i] = i;
  }
  S maxpivot = S(0);
  int nonzero = n;
  for(int k = 0; k < n; k++)
  {
    int pr = k, pc = k;
    S best = S(-1);
    for(int j = k; j < n; j++)
      for(int i = k; i < n; i++)
      {
        const S v = fabs(f.a[i + j * n]);
        if(v > best)
        {
          best = v;
          pr = i;
          pc = j;
        }
      }
    if(best == S(0))
    {
      nonzero = k; // the trailing block is exactly zero: no more pivots
      break;
    }
    if(best > maxpivot) maxpivot = best;
    f.row_tr[k] = pr;
    if(pr != k)
      for(int j = 0; j < n; j++)
      {
        const S t = f.a[k + j * n];
        f.a[k + j * n] = f.a[pr + j * n];
        f.a[pr + j * n] = t;
      }
    if(pc != k)
    {
      for(int i = 0; i < n; i++)
      {
        const S t = f.a[i + k * n];
        f.a[i + k * n] = f.a[i + pc * n];
        f.a[i + pc * n] = t;
      }
      const int t = f.col_perm[k];
      f.col_perm[k] = f.col_perm[pc];
      f.col_perm[pc] = t;
    }
    const S p = f.a[k + k * n];
    for(int i = k + 1; i < n; i++) f.a[i + k * n] /= p;
    for(int j = k + 1; j < n; j++)
      for(int i = k + 1; i < n; i++) f.a[i + j * n] -= f.a[i + k * n] * f.a[k + j * n];
  }
  // rank with Eigen's default threshold: epsilon * diagonal size, relative to the largest pivot
  const S thr = machineEps<S>() * S(n) * maxpivot;
  f.rank = 0;
  for(int i = 0; i < nonzero; i++)
    if(fabs(f.a[i + i * n]) > thr) f.rank++;
}

/** b <- the solution Eigen::FullPivLU::solve returns (components outside the leading rank x rank block are zero). */
template<class S, int n>
__device__ inline void fullPivLuSolveInPlace(const FullPivLuFactor<S, n> & f, S * b)
{
  S c[n];
  for(int i = 0; i < n; i++) c[i] = b[i];
  for(int k = 0; k < n; k++)
  {
    const int r = f.row_tr[k];
    if(r != k)
    {
      const S t = c[k];
      c[k] = c[r];
      c[r] = t;
    }
  }
  for(int k = 0; k < n; k++)
    for(int i = k + 1; i < n; i++) c[i] -= f.a[i + k * n] * c[k];
  for(int i = f.rank - 1; i >= 0; i--)
  {
    S s = c[i];
    for(int j = i + 1; j < f.rank; j++) s -= f.a[i + j * n] * c[j];
    c[i] = s / f.a[i + i * n];
  }
  for(int i = 0; i < n; i++) b[f.col_perm[i]] = (i < f.rank) ? c[i] : S(0);
}
} // namespace fmpc
} // namespace nmpc_b200
