/* nmpc_b200 -- fixed-size dense matrix usable in host and device code.
 *
 * The reference expresses every problem in Eigen fixed-size types (DDPProblem.h:20-35 of
 * isri-aist/NMPC).  Eigen is neither a dependency of this engine nor usable inside its kernels in
 * this form, so problem functors are written against this type instead.  It covers the idioms the
 * reference's problem bodies use (TestDDPCartPole.cpp:63-227): operator[] / operator()(i,j),
 * setZero(), setConstant(), *= s, + - and s * m, dot(), cwiseProduct(), cwiseAbs2(),
 * squaredNorm(), asDiagonal-style assignment (setDiagonal), addToDiagonal.
 * Storage is column-major like Eigen's default; everything is constexpr-sized and inlined so that
 * after unrolling the coefficients live in registers.
 */
#pragma once

#if defined(__CUDACC__)
#  define NMPC_HD __host__ __device__ __forceinline__
#  define NMPC_UNROLL _Pragma("unroll")
#else
#  define NMPC_HD inline
#  define NMPC_UNROLL
#endif

namespace nmpc_b200
{
template<class S, int R, int C>
struct Matrix
{
  S d[(R * C) > 0 ? (R * C) : 1];

  NMPC_HD static constexpr int rows()
  {
    return R;
  }
  NMPC_HD static constexpr int cols()
  {
    return C;
  }
  NMPC_HD static constexpr int size()
  {
    return R * C;
  }
  NMPC_HD S & operator()(int i, int j)
  {
    return d[i + j * R];
  }
  NMPC_HD const S & operator()(int i, int j) const
  {
    return d[i + j * R];
  }
  NMPC_HD S & operator[](int i)
  {
    return d[i];
  }
  NMPC_HD const S & operator[](int i) const
  {
    return d[i];
  }
  NMPC_HD void setZero()
  {
NMPC_UNROLL
    for(int i = 0; i < R * C; i++) d[i] = S(0);
  }
  NMPC_HD void setConstant(S v)
  {
NMPC_UNROLL
    for(int i = 0; i < R * C; i++) d[i] = v;
  }
  NMPC_HD void setIdentity()
  {
NMPC_UNROLL
    for(int j = 0; j < C; j++)
NMPC_UNROLL
      for(int i = 0; i < R; i++) d[i + j * R] = (i == j) ? S(1) : S(0);
  }
  /** m = v.asDiagonal() */
  NMPC_HD void setDiagonal(const Matrix<S, R, 1> & v)
  {
NMPC_UNROLL
    for(int j = 0; j < C; j++)
NMPC_UNROLL
      for(int i = 0; i < R; i++) d[i + j * R] = (i == j) ? v.d[i] : S(0);
  }
  /** m.diagonal().array() += s */
  NMPC_HD void addToDiagonal(S s)
  {
NMPC_UNROLL
    for(int i = 0; i < (R < C ? R : C); i++) d[i + i * R] += s;
  }
  NMPC_HD static Matrix Zero()
  {
    Matrix m;
    m.setZero();
    return m;
  }
  NMPC_HD Matrix & operator*=(S s)
  {
NMPC_UNROLL
    for(int i = 0; i < R * C; i++) d[i] *= s;
    return *this;
  }
  NMPC_HD Matrix & operator+=(const Matrix & o)
  {
NMPC_UNROLL
    for(int i = 0; i < R * C; i++) d[i] += o.d[i];
    return *this;
  }
  NMPC_HD S dot(const Matrix & o) const
  {
    S s = S(0);
NMPC_UNROLL
    for(int i = 0; i < R * C; i++) s += d[i] * o.d[i];
    return s;
  }
  NMPC_HD S squaredNorm() const
  {
    return dot(*this);
  }
  NMPC_HD Matrix cwiseProduct(const Matrix & o) const
  {
    Matrix m;
NMPC_UNROLL
    for(int i = 0; i < R * C; i++) m.d[i] = d[i] * o.d[i];
    return m;
  }
  NMPC_HD Matrix cwiseAbs2() const
  {
    Matrix m;
NMPC_UNROLL
    for(int i = 0; i < R * C; i++) m.d[i] = d[i] * d[i];
    return m;
  }
};

template<class S, int R, int C>
NMPC_HD Matrix<S, R, C> operator+(const Matrix<S, R, C> & a, const Matrix<S, R, C> & b)
{
  Matrix<S, R, C> m;
NMPC_UNROLL
  for(int i = 0; i < R * C; i++) m.d[i] = a.d[i] + b.d[i];
  return m;
}

template<class S, int R, int C>
NMPC_HD Matrix<S, R, C> operator-(const Matrix<S, R, C> & a, const Matrix<S, R, C> & b)
{
  Matrix<S, R, C> m;
NMPC_UNROLL
  for(int i = 0; i < R * C; i++) m.d[i] = a.d[i] - b.d[i];
  return m;
}

template<class S, int R, int C>
NMPC_HD Matrix<S, R, C> operator*(S s, const Matrix<S, R, C> & a)
{
  Matrix<S, R, C> m;
NMPC_UNROLL
  for(int i = 0; i < R * C; i++) m.d[i] = s * a.d[i];
  return m;
}

template<class S, int R, int K, int C>
NMPC_HD Matrix<S, R, C> operator*(const Matrix<S, R, K> & a, const Matrix<S, K, C> & b)
{
  Matrix<S, R, C> m;
NMPC_UNROLL
  for(int j = 0; j < C; j++)
NMPC_UNROLL
    for(int i = 0; i < R; i++)
    {
      S s = S(0);
NMPC_UNROLL
      for(int k = 0; k < K; k++) s += a(i, k) * b(k, j);
      m(i, j) = s;
    }
  return m;
}
} // namespace nmpc_b200
