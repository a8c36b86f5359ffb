/* nmpc_b200 -- K2 for problems with MANY inputs (n_u >= 8; centroidal motion: n_x = 9, n_u = 16).
 *
 * backwardPass() of the reference (isri-aist/NMPC nmpc_ddp/include/nmpc_ddp/DDPSolver.hpp:343-534, unconstrained
 * branch) with GS lanes per instance, like ddp_backward_coop.cuh -- but there every lane recomputes the n_u x n_u part
 * (Tu = Fu^T Vxx, Quu, its Cholesky factor) in registers, which for 9 x 16 is ~10^4 fp64 instructions per lane and
 * step and 36-70 KB of spills per thread (measured: 135 us per step).  Here the n_u side is spread over the lanes too
 * and every matrix lives in shared memory:
 *
 *   lane a < n_u   row a of Tu and of Qux, column a of Quu / Quu_F, Qu(a), (Quu k)(a), ROW a of the Cholesky factor
 *   lane c < n_x   column c of Qxx, Qx(c), the triangular solves for K(:, c), row c of K^T Quu, row c of Vxx'
 *   lane n_x       the triangular solves for k
 *
 * The Cholesky factorisation is left-looking and cooperative: at column step p every lane recomputes the pivot from
 * row p in shared memory (same verdict in every lane: Eigen::LLT's "pivot <= 0 => NumericalIssue"), lane i > p
 * finishes L(i, p) from its own row.  Each coefficient is the same expression, summed in the same order, as in
 * backward_kernel / backward_coop_kernel (lltInPlace, lltSolveInPlace), so results agree to the last bit except where
 * the compiler contracts differently.  One warp per CTA (two instances for GS = 16).
 */
#pragma once

#include "ddp_kernels.cuh"

namespace nmpc_b200
{
namespace ddp
{
template<class M, int GS>
struct WideLayout
{
  static constexpr int NX = M::NX, NU = M::NU;
  using L = BlockLayout<NX, NU>;
  static_assert(NU <= GS && NX < GS && GS <= 32, "one lane per input and per state column, one more for k");
  static constexpr int IPW = 32 / GS; //!< instances per warp
  static constexpr int STAGE = L::SIZE + NU; //!< derivative block + u_i
  static constexpr int DEPTH = 2; //!< ring slots: step i lives in slot i % DEPTH; one step (~15 us) of prefetch distance
                                  //!< hides the load latency, and 50 KB per CTA keeps four CTAs (one per scheduler) on an SM
  static constexpr int RING = 0;
  static constexpr int VXX = RING + DEPTH * STAGE;
  static constexpr int VX = VXX + NX * NX;
  static constexpr int TU = VX + NX; //!< Fu^T Vxx, [a + q * NU]
  static constexpr int QU = TU + NU * NX;
  static constexpr int QXX = QU + NU;
  static constexpr int QUU = QXX + NX * NX;
  static constexpr int QF = QUU + NU * NU; //!< regularised Quu
  static constexpr int LF = QF + NU * NU; //!< its Cholesky factor (lower)
  static constexpr int QUX = LF + NU * NU; //!< [a + c * NU]
  static constexpr int QUXR = QUX + NU * NX; //!< regularised Qux (reg_type 2)
  static constexpr int KFB = QUXR + NU * NX; //!< K, [a + c * NU]
  static constexpr int KFF = KFB + NU * NX; //!< k
  static constexpr int QUUK = KFF + NU;
  static constexpr int VN = QUUK + NU; //!< unsymmetrised Vxx'
  static constexpr int ELEMS = VN + NX * NX; //!< per instance
  static constexpr int WARP_ELEMS = ELEMS * IPW;
  static constexpr size_t bytes()
  {
    return sizeof(typename M::Scalar) * (size_t)WARP_ELEMS;
  }
};

/** One backwardPass() sweep, GS lanes per instance.  Every lane of the warp executes the loop (it contains warp
    barriers); only lanes with `work` compute.  Returns false (uniformly within the group) when the factorisation of
    Quu_F fails at some step.  dV0 / dV1 / k_rel_norm are valid in lane j == 0. */
template<class M, int GS>
__device__ __forceinline__ bool backwardSweepWide(const Workspace<typename M::Scalar> & ws,
                                                  const SolverParams<typename M::Scalar> & prm,
                                                  int b,
                                                  int j,
                                                  const typename M::Scalar * __restrict__ us,
                                                  typename M::Scalar * __restrict__ sm,
                                                  bool work,
                                                  typename M::Scalar lambda,
                                                  typename M::Scalar & dV0,
                                                  typename M::Scalar & dV1,
                                                  typename M::Scalar & k_rel_norm)
{
  using S = typename M::Scalar;
  constexpr int NX = M::NX, NU = M::NU;
  using L = BlockLayout<NX, NU>;
  using C = WideLayout<M, GS>;
  constexpr int IPW = C::IPW;
  const size_t Bp = ws.Bp;
  const int N = prm.N;
  const bool reg2 = prm.reg_type == 2;

  // element e of this instance's region `off` lives at sm[(off + e) * IPW]
  auto at = [&](int off, int e) -> S & { return sm[(size_t)(off + e) * IPW]; };

  auto stageStep = [&](int step) {
    const int stage = (step + C::DEPTH) % C::DEPTH;
    if(work && step >= 0)
    {
      for(int e = j; e < C::STAGE; e += GS)
      {
        const S * src = (e < L::SIZE) ? ws.deriv + derivTileOffset<L::SIZE>(step, b, ws.Bp) + (size_t)e * kTile
                                      : us + ((size_t)step * NU + (e - L::SIZE)) * Bp + b;
        S * dst = &at(C::RING + stage * C::STAGE, e);
        if constexpr(sizeof(S) == 8)
          cpAsync8(dst, src);
        else
          cpAsync4(dst, src);
      }
    }
    cpAsyncCommit();
  };

  // terminal value function
  if(work && j < NX)
  {
    at(C::VX, j) = ws.vterm[(size_t)j * Bp + b];
#pragma unroll
    for(int r = 0; r < NX; r++) at(C::VXX, r + j * NX) = ws.vterm[(size_t)(NX + r + j * NX) * Bp + b];
  }
#pragma unroll
  for(int d = 1; d < C::DEPTH; d++) stageStep(N - d);

  dV0 = S(0);
  dV1 = S(0);
  S krn_num = S(0), krn_den = S(1);
  bool ok = true;

  for(int i = N - 1; i >= 0; i--)
  {
    cpAsyncWait<C::DEPTH - 2>(); // this lane's share of block i has landed
    __syncwarp(); // [S1] ... and everyone else's; Vxx / Vx of the previous step visible; step i+1 fully retired
    const int blk = C::RING + (i % C::DEPTH) * C::STAGE;
    stageStep(i - (C::DEPTH - 1)); // refill the slot step i+1 just vacated

    const bool act = work && ok;
    S fu[NX], tu[NX]; // lane a < NU: Fu(:, a), row a of Tu = Fu^T Vxx
    S Qx_c = S(0); // lane c < NX
    if(act)
    {
      if(j < NU)
      {
        // Qu = Lu + Fu^T Vx, Tu = Fu^T Vxx                                             (:386, :399)
#pragma unroll
        for(int r = 0; r < NX; r++) fu[r] = at(blk, L::FU + r + j * NX);
        S s = S(0);
#pragma unroll
        for(int r = 0; r < NX; r++) s += fu[r] * at(C::VX, r);
        at(C::QU, j) = at(blk, L::LU + j) + s;
#pragma unroll
        for(int q = 0; q < NX; q++)
        {
          S t = S(0);
#pragma unroll
          for(int r = 0; r < NX; r++) t += fu[r] * at(C::VXX, r + q * NX);
          tu[q] = t;
          at(C::TU, j + q * NU) = t;
        }
      }
      if(j < NX)
      {
        // Qx = Lx + Fx^T Vx, Qxx = Lxx + Fx^T Vxx Fx: column j                         (:388, :404-408)
        S fxc[NX], W[NX];
#pragma unroll
        for(int r = 0; r < NX; r++) fxc[r] = at(blk, L::FX + r + j * NX);
        S s0 = S(0);
#pragma unroll
        for(int r = 0; r < NX; r++) s0 += fxc[r] * at(C::VX, r);
        Qx_c = at(blk, L::LX + j) + s0;
#pragma unroll
        for(int r = 0; r < NX; r++)
        {
          S s = S(0);
#pragma unroll
          for(int q = 0; q < NX; q++) s += at(C::VXX, r + q * NX) * fxc[q];
          W[r] = s;
        }
#pragma unroll
        for(int r = 0; r < NX; r++)
        {
          S s = S(0);
#pragma unroll
          for(int q = 0; q < NX; q++) s += at(blk, L::FX + q + r * NX) * W[q];
          at(C::QXX, r + j * NX) = at(blk, L::LXX + r + j * NX) + s;
        }
      }
    }
    __syncwarp(); // [Sa] Tu complete

    if(act && j < NU)
    {
      // Quu = Luu + Tu Fu: column j; Quu_F (:421-441): reg_type 1 Quu + lambda I, reg_type 2 Vxx + lambda I inside
#pragma unroll 4
      for(int a = 0; a < NU; a++)
      {
        S s = S(0), sr = S(0);
#pragma unroll
        for(int r = 0; r < NX; r++)
        {
          const S t = at(C::TU, a + r * NU);
          s += t * fu[r];
          if(reg2) sr += (t + lambda * at(blk, L::FU + r + a * NX)) * fu[r];
        }
        const S luu = at(blk, L::LUU + a + j * NU);
        at(C::QUU, a + j * NU) = luu + s;
        S f = reg2 ? (luu + sr) : (luu + s);
        if(!reg2 && a == j) f += lambda;
        at(C::QF, a + j * NU) = f;
      }
      // Qux = Lxu^T + Tu Fx: row j; Qux_reg likewise                                    (:402, :427)
#pragma unroll
      for(int c = 0; c < NX; c++)
      {
        S s = S(0), sr = S(0);
#pragma unroll
        for(int q = 0; q < NX; q++)
        {
          const S f = at(blk, L::FX + q + c * NX);
          s += tu[q] * f;
          if(reg2) sr += (tu[q] + lambda * fu[q]) * f;
        }
        const S lxu = at(blk, L::LXU + c + j * NX);
        at(C::QUX, j + c * NU) = lxu + s;
        at(C::QUXR, j + c * NU) = reg2 ? (lxu + sr) : (lxu + s);
      }
    }
    __syncwarp(); // [Sb] Quu_F, Qux complete

    // Cholesky factor of Quu_F, row j in this lane (lltInPlace's expressions)          (:500)
    S invd[NU];
    {
      S lrow[NU];
      if(act && j < NU)
      {
#pragma unroll
        for(int p = 0; p < NU; p++) lrow[p] = (p <= j) ? at(C::QF, j + p * NU) : S(0);
      }
#pragma unroll
      for(int p = 0; p < NU; p++)
      {
        if(act)
        {
          S lp[NU];
          S x = at(C::QF, p + p * NU);
#pragma unroll
          for(int q = 0; q < p; q++)
          {
            lp[q] = at(C::LF, p + q * NU);
            x -= lp[q] * lp[q];
          }
          if(x <= S(0)) ok = false;
          x = sqrt(x);
          const S inv = S(1) / x;
          invd[p] = inv;
          if(j == p)
          {
            at(C::LF, p + p * NU) = x;
          }
          else if(j > p && j < NU)
          {
            S s = lrow[p];
#pragma unroll
            for(int q = 0; q < p; q++) s -= lrow[q] * lp[q];
            s = s * inv;
            lrow[p] = s;
            at(C::LF, j + p * NU) = s;
          }
        }
        __syncwarp(); // column p of the factor visible
      }
    }

    // k = -Quu_F^-1 Qu (lane NX), K(:, c) = -Quu_F^-1 Qux_reg(:, c) (lane c < NX)     (:501-510)
    const bool act1 = work && ok;
    S Kc[NU];
    if(act1 && j <= NX)
    {
      const int src = (j < NX) ? (C::QUXR + j * NU) : C::QU;
#pragma unroll
      for(int a = 0; a < NU; a++) Kc[a] = at(src, a);
#pragma unroll
      for(int r = 0; r < NU; r++)
      {
        S s = Kc[r];
#pragma unroll
        for(int q = 0; q < r; q++) s -= at(C::LF, r + q * NU) * Kc[q];
        Kc[r] = s * invd[r];
      }
#pragma unroll
      for(int r = NU - 1; r >= 0; r--)
      {
        S s = Kc[r];
#pragma unroll
        for(int q = r + 1; q < NU; q++) s -= at(C::LF, q + r * NU) * Kc[q];
        Kc[r] = s * invd[r];
      }
      const int dst = (j < NX) ? (C::KFB + j * NU) : C::KFF;
      S * gdst = (j < NX) ? ws.kfb + ((size_t)i * NU * NX + (size_t)j * NU) * Bp + b : ws.kff + ((size_t)i * NU) * Bp + b;
#pragma unroll
      for(int a = 0; a < NU; a++)
      {
        Kc[a] = -Kc[a];
        at(dst, a) = Kc[a];
        gdst[(size_t)a * Bp] = Kc[a]; // gains of this step (:529-530)
      }
    }
    __syncwarp(); // [Sd] k, K complete

    S ktq[NU]; // lane c < NX: row c of K^T Quu
    if(act1)
    {
      if(j < NU)
      {
        S s = S(0);
#pragma unroll
        for(int c2 = 0; c2 < NU; c2++) s += at(C::QUU, j + c2 * NU) * at(C::KFF, c2);
        at(C::QUUK, j) = s;
      }
      if(j < NX)
      {
#pragma unroll
        for(int a = 0; a < NU; a++)
        {
          S s = S(0);
#pragma unroll
          for(int a2 = 0; a2 < NU; a2++) s += Kc[a2] * at(C::QUU, a2 + a * NU);
          ktq[a] = s;
        }
      }
    }
    __syncwarp(); // [Se] Quu k complete

    if(act1)
    {
      if(j == 0)
      {
        // expected cost change (:522) and the small-gradient measure (:219-221)
        S s0 = S(0), s1 = S(0), kn = S(0), un = S(0);
#pragma unroll
        for(int a = 0; a < NU; a++)
        {
          const S ka = at(C::KFF, a);
          s0 += ka * at(C::QU, a);
          s1 += ka * at(C::QUUK, a);
          kn += ka * ka;
          const S uv = at(blk, L::SIZE + a);
          un += uv * uv;
        }
        dV0 += s0;
        dV1 += S(0.5) * s1;
        const S a_num = sqrt(kn);
        const S a_den = sqrt(un) + S(1);
        if(a_num * krn_den > krn_num * a_den)
        {
          krn_num = a_num;
          krn_den = a_den;
        }
      }
      if(j < NX)
      {
        // Vx = Qx + K^T Quu k + K^T Qu + Qux^T k; Vxx = Qxx + K^T Quu K + K^T Qux + Qux^T K: ROW j   (:523-525)
        S quxc[NU];
#pragma unroll
        for(int a = 0; a < NU; a++) quxc[a] = at(C::QUX, a + j * NU);
        S s1 = S(0), s2 = S(0), s3 = S(0);
#pragma unroll
        for(int a = 0; a < NU; a++)
        {
          const S ka = at(C::KFF, a);
          s1 += ktq[a] * ka;
          s2 += Kc[a] * at(C::QU, a);
          s3 += quxc[a] * ka;
        }
        Qx_c = ((Qx_c + s1) + s2) + s3; // Vx'(j); written after [Sf]
#pragma unroll
        for(int r = 0; r < NX; r++)
        {
          S t1 = S(0), t2 = S(0), t3 = S(0);
#pragma unroll
          for(int a = 0; a < NU; a++)
          {
            const S kar = at(C::KFB, a + r * NU);
            t1 += ktq[a] * kar;
            t2 += Kc[a] * at(C::QUX, a + r * NU);
            t3 += quxc[a] * kar;
          }
          at(C::VN, j + r * NX) = ((at(C::QXX, j + r * NX) + t1) + t2) + t3;
        }
      }
    }
    __syncwarp(); // [Sf] unsymmetrised Vxx' complete

    if(act1 && j < NX)
    {
      at(C::VX, j) = Qx_c;
      // Vxx <- 0.5 (Vxx + Vxx^T)                                                       (:526)
#pragma unroll
      for(int r = 0; r < NX; r++) at(C::VXX, r + j * NX) = S(0.5) * (at(C::VN, r + j * NX) + at(C::VN, j + r * NX));
    }
  }
  cpAsyncWait<0>();
  __syncwarp();
  k_rel_norm = krn_num / krn_den;
  return ok;
}

/** procOnce() Step 2 (DDPSolver.hpp:188-231) for the wide sweep; same driver as backward_coop_kernel. */
template<class M, int GS>
__global__ void __launch_bounds__(32) backward_wide_kernel(const __grid_constant__ Workspace<typename M::Scalar> ws,
                                                           const __grid_constant__ SolverParams<typename M::Scalar> prm,
                                                           int iter)
{
  pdlPrologue();
  using S = typename M::Scalar;
  using C = WideLayout<M, GS>;
  constexpr unsigned kFull = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int g = lane / GS;
  const int j = lane % GS;
  S * sm = reinterpret_cast<S *>(smem_raw) + g;

  const int bg = blockIdx.x * C::IPW + g;
  const int b = (bg < ws.B) ? bg : (ws.B - 1);
  const bool live = (bg < ws.B) && (ws.status[b] == 0);

  S lambda = ws.lambda[b];
  S dlambda = ws.dlambda[b];
  const S * us = ws.u[ws.sel[b]];
  int n_bwd = ws.n_bwd[b];
  S dV0 = S(0), dV1 = S(0), k_rel_norm = S(0);
  bool need = live;
  bool failed = false;
  while(__any_sync(kFull, need))
  {
    if(need) n_bwd++;
    // results are kept only for instances that needed this sweep: one that is waiting for its tile mates' lambda retry
    // keeps the dV / k_rel_norm of its own successful sweep
    S sw_dV0 = S(0), sw_dV1 = S(0), sw_krn = S(0);
    const bool ok = backwardSweepWide<M, GS>(ws, prm, b, j, us, sm, need, lambda, sw_dV0, sw_dV1, sw_krn);
    if(need)
    {
      dV0 = sw_dV0;
      dV1 = sw_dV1;
      k_rel_norm = sw_krn;
    }
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
  if(!live || j != 0) return;

  ws.n_bwd[b] = n_bwd;
  ws.lambda[b] = lambda;
  ws.dlambda[b] = dlambda;
  if(failed)
  {
    ws.status[b] = -1;
    ws.iters[b] = iter;
    writeTrace<S>(ws, b, iter, S(iter), S(0), S(0), S(0), S(0), S(0), S(0), S(0), S(0));
    return;
  }
  ws.dV[b] = dV0;
  ws.dV[(size_t)ws.Bp + b] = dV1;
  if(k_rel_norm < prm.k_rel_norm_thre && lambda < prm.lambda_thre)
  {
    ws.status[b] = 1;
    ws.iters[b] = iter;
    writeTrace<S>(ws, b, iter, S(iter), S(0), S(0), S(0), S(0), k_rel_norm, S(0), S(0), S(0));
    return;
  }
  ws.trace[((size_t)iter * kTraceFields + 5) * ws.Bp + b] = k_rel_norm;
}
} // namespace ddp
} // namespace nmpc_b200
