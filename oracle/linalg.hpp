// TEST INFRASTRUCTURE ONLY -- CPU oracle for nmpc_b200 (see oracle/README.md).
// Nothing in the product path (nmpc_b200/, include/) may include or link this file.
//
// Minimal fixed-size dense algebra used by the oracle restatement of
// nmpc_ddp::DDPSolver / nmpc_fmpc::FmpcSolver.  The reference delegates all of this to
// Eigen 3 (third party, not vendored under /root/reference, version unpinned:
// nmpc_ddp/CMakeLists.txt:29-34, debian/control:12).  What is restated here is Eigen's
// *published* behaviour for the handful of operations the solvers use:
//   - column-major fixed-size matrices (Eigen default storage order),
//   - coefficient-wise lazy products: c(i,j) = sum_k a(i,k) * b(k,j), k ascending,
//   - LLT (unblocked right-looking Cholesky): at column k, x = A(k,k) - |L(k,0:k)|^2;
//     "x <= 0" => NumericalIssue (a NaN pivot therefore does NOT fail), L(k,k) = sqrt(x),
//     then forward/back substitution with division by the diagonal,
//   - LDLT with diagonal pivoting (largest |diagonal| first), used by FMPC.
// Sums are plain left-to-right; Eigen's packet reductions differ from this at the ulp level,
// which is why parity with anything Eigen-based is tolerance-based, never bitwise.
#pragma once

#include <cmath>
#include <cstring>

namespace oracle
{
template<int R, int C>
struct Mat
{
  double d[(R * C) > 0 ? (R * C) : 1];

  static constexpr int rows()
  {
    return R;
  }
  static constexpr int cols()
  {
    return C;
  }
  double & operator()(int i, int j)
  {
    return d[i + j * R];
  }
  const double & operator()(int i, int j) const
  {
    return d[i + j * R];
  }
  double & operator[](int i)
  {
    return d[i];
  }
  const double & operator[](int i) const
  {
    return d[i];
  }
  void setZero()
  {
    for(int i = 0; i < R * C; i++) d[i] = 0.0;
  }
  void setConstant(double v)
  {
    for(int i = 0; i < R * C; i++) d[i] = v;
  }
  static Mat Zero()
  {
    Mat m;
    m.setZero();
    return m;
  }
};

template<int N>
using Vec = Mat<N, 1>;

// c = a * b
template<int R, int K, int C>
inline Mat<R, C> mul(const Mat<R, K> & a, const Mat<K, C> & b)
{
  Mat<R, C> c;
  for(int j = 0; j < C; j++)
    for(int i = 0; i < R; i++)
    {
      double s = 0.0;
      for(int k = 0; k < K; k++) s += a(i, k) * b(k, j);
      c(i, j) = s;
    }
  return c;
}

// c = a^T * b
template<int K, int R, int C>
inline Mat<R, C> mulT(const Mat<K, R> & a, const Mat<K, C> & b)
{
  Mat<R, C> c;
  for(int j = 0; j < C; j++)
    for(int i = 0; i < R; i++)
    {
      double s = 0.0;
      for(int k = 0; k < K; k++) s += a(k, i) * b(k, j);
      c(i, j) = s;
    }
  return c;
}

template<int R, int C>
inline Mat<C, R> transpose(const Mat<R, C> & a)
{
  Mat<C, R> t;
  for(int j = 0; j < C; j++)
    for(int i = 0; i < R; i++) t(j, i) = a(i, j);
  return t;
}

template<int R, int C>
inline Mat<R, C> add(const Mat<R, C> & a, const Mat<R, C> & b)
{
  Mat<R, C> c;
  for(int i = 0; i < R * C; i++) c.d[i] = a.d[i] + b.d[i];
  return c;
}

template<int R, int C>
inline Mat<R, C> sub(const Mat<R, C> & a, const Mat<R, C> & b)
{
  Mat<R, C> c;
  for(int i = 0; i < R * C; i++) c.d[i] = a.d[i] - b.d[i];
  return c;
}

template<int R, int C>
inline Mat<R, C> scale(double s, const Mat<R, C> & a)
{
  Mat<R, C> c;
  for(int i = 0; i < R * C; i++) c.d[i] = s * a.d[i];
  return c;
}

template<int N>
inline double dot(const Vec<N> & a, const Vec<N> & b)
{
  double s = 0.0;
  for(int i = 0; i < N; i++) s += a[i] * b[i];
  return s;
}

template<int R, int C>
inline double squaredNorm(const Mat<R, C> & a)
{
  double s = 0.0;
  for(int i = 0; i < R * C; i++) s += a.d[i] * a.d[i];
  return s;
}

template<int N>
inline double norm(const Vec<N> & a)
{
  return std::sqrt(squaredNorm(a));
}

template<int R, int C>
inline bool hasNaNOrInf(const Mat<R, C> & a)
{
  for(int i = 0; i < R * C; i++)
    if(std::isnan(a.d[i]) || std::isinf(a.d[i])) return true;
  return false;
}

/** Cholesky A = L L^T of the leading n x n block of a column-major array with leading
    dimension ld; semantics of Eigen::LLT (see header comment).  Returns false on
    "NumericalIssue" (a pivot <= 0).  Only the lower triangle of `a` is read; L overwrites it. */
inline bool lltInPlace(double * a, int n, int ld)
{
  for(int k = 0; k < n; k++)
  {
    double x = a[k + k * ld];
    for(int j = 0; j < k; j++) x -= a[k + j * ld] * a[k + j * ld];
    if(x <= 0.0) return false;
    x = std::sqrt(x);
    a[k + k * ld] = x;
    for(int i = k + 1; i < n; i++)
    {
      double s = a[i + k * ld];
      for(int j = 0; j < k; j++) s -= a[i + j * ld] * a[k + j * ld];
      a[i + k * ld] = s / x;
    }
  }
  return true;
}

/** Solve L L^T x = b in place (b -> x), L from lltInPlace. */
inline void lltSolveInPlace(const double * l, int n, int ld, double * b)
{
  for(int i = 0; i < n; i++)
  {
    double s = b[i];
    for(int j = 0; j < i; j++) s -= l[i + j * ld] * b[j];
    b[i] = s / l[i + i * ld];
  }
  for(int i = n - 1; i >= 0; i--)
  {
    double s = b[i];
    for(int j = i + 1; j < n; j++) s -= l[j + i * ld] * b[j];
    b[i] = s / l[i + i * ld];
  }
}

/** LDLT with diagonal pivoting (P A P^T = L D L^T), semantics of Eigen::LDLT: at step k the
    largest remaining |diagonal| is swapped into place; info() is Success unless a non-finite
    value appears.  Solves A x = b for `nrhs` right-hand sides stored column-major in b (n x nrhs).
    Returns false when the factorisation hits a non-finite pivot (Eigen: NumericalIssue). */
inline bool ldltSolve(const double * a_in, int n, double * b, int nrhs)
{
  constexpr int MAXN = 32;
  double a[MAXN * MAXN];
  int perm[MAXN];
  for(int j = 0; j < n; j++)
    for(int i = 0; i < n; i++) a[i + j * n] = a_in[i + j * n];
  for(int i = 0; i < n; i++) perm[i] = i;

  for(int k = 0; k < n; k++)
  {
    // pivot: largest |a(i,i)|, i >= k
    int piv = k;
    double best = std::fabs(a[k + k * n]);
    for(int i = k + 1; i < n; i++)
    {
      double v = std::fabs(a[i + i * n]);
      if(v > best)
      {
        best = v;
        piv = i;
      }
    }
    if(piv != k)
    {
      // symmetric row/column swap of the (lower-triangular) working matrix
      for(int j = 0; j < n; j++)
      {
        double t = a[k + j * n];
        a[k + j * n] = a[piv + j * n];
        a[piv + j * n] = t;
      }
      for(int i = 0; i < n; i++)
      {
        double t = a[i + k * n];
        a[i + k * n] = a[i + piv * n];
        a[i + piv * n] = t;
      }
      int t = perm[k];
      perm[k] = perm[piv];
      perm[piv] = t;
    }
    // d_k = a(k,k) - sum_j L(k,j)^2 d_j ; stored: a(k,k) = d_k, a(i,k) = L(i,k)
    double dk = a[k + k * n];
    for(int j = 0; j < k; j++) dk -= a[k + j * n] * a[k + j * n] * a[j + j * n];
    a[k + k * n] = dk;
    if(!std::isfinite(dk)) return false;
    for(int i = k + 1; i < n; i++)
    {
      double s = a[i + k * n];
      for(int j = 0; j < k; j++) s -= a[i + j * n] * a[k + j * n] * a[j + j * n];
      // Eigen leaves the column untouched when the pivot is exactly zero
      a[i + k * n] = (dk != 0.0) ? s / dk : s;
    }
  }

  for(int r = 0; r < nrhs; r++)
  {
    double y[MAXN];
    for(int i = 0; i < n; i++) y[i] = b[perm[i] + r * n];
    for(int i = 0; i < n; i++)
      for(int j = 0; j < i; j++) y[i] -= a[i + j * n] * y[j];
    for(int i = 0; i < n; i++)
    {
      double dk = a[i + i * n];
      // Eigen::LDLT::solve uses a pseudo-inverse of D: tiny pivots give 0
      y[i] = (std::fabs(dk) > 2.2250738585072014e-308) ? y[i] / dk : 0.0;
    }
    for(int i = n - 1; i >= 0; i--)
      for(int j = i + 1; j < n; j++) y[i] -= a[j + i * n] * y[j];
    for(int i = 0; i < n; i++) b[perm[i] + r * n] = y[i];
  }
  return true;
}
} // namespace oracle
