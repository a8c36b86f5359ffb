/* nmpc_b200 -- K2 for small batches: GS lanes cooperate on one instance's backward pass.
 *
 * A 4096-instance batch gives the thread-per-instance kernel only 128 warps for 592 warp schedulers,
 * and each of those warps issues ~930 instructions per horizon step along one dependent chain.  Here
 * lane j of a group owns column j (and j+GS, ...) of the n_x x n_x matrices, so the per-lane stream
 * shrinks to the column work plus a few small redundant pieces (Quu, its factorisation, k):
 *
 *   - the step's derivative block is staged ONCE per instance in shared memory by the group's lanes
 *     (cp.async, two-stage ring) and read back with warp-broadcast loads;
 *   - Vxx/Vx live in shared memory between steps; W_c = Vxx Fx(:,c), Qxx(:,c) = Lxx(:,c) + Fx^T W_c and
 *     Qux(:,c), K(:,c) are lane-local; K/Qux columns and the unsymmetrised Vxx columns are exchanged
 *     through shared memory with three warp barriers per step.
 *
 * Same mathematics as ddp::backward_kernel (DDPSolver.hpp:188-231, :343-534); the product
 * Fx^T Vxx Fx is associated as Fx^T (Vxx Fx) here, which changes results at rounding level only.
 */
#pragma once

#include "ddp_kernels.cuh"

namespace nmpc_b200
{
namespace ddp
{
template<class M, int GS>
struct CoopLayout
{
  static constexpr int NX = M::NX, NU = M::NU;
  using L = BlockLayout<NX, NU>;
  static constexpr int IPW = 32 / GS; //!< instances per warp
  static constexpr int CPL = (NX + GS - 1) / GS; //!< matrix columns per lane
  static constexpr int STAGE = L::SIZE + NU; //!< derivative block + u_i
  static constexpr int DEPTH = 4; //!< ring slots: step i lives in slot i % DEPTH, prefetch distance DEPTH - 1 steps
  static constexpr int RING = 0;
  static constexpr int VXX = RING + DEPTH * STAGE;
  static constexpr int VX = VXX + NX * NX;
  static constexpr int KFB = VX + NX;
  static constexpr int QUX = KFB + NU * NX;
  static constexpr int VN = QUX + NU * NX;
  static constexpr int ELEMS = VN + NX * NX; //!< per instance
  static constexpr int WARP_ELEMS = ELEMS * IPW;
};

/** One cooperative backwardPass() sweep.  Every lane of the warp executes the loop (it contains warp
    barriers); only lanes with `work` compute.  Returns false (uniformly within the group) when the
    factorisation of Quu_F fails at some step. */
template<class M, int GS, bool CONSTRAINED>
__device__ __forceinline__ bool backwardSweepCoop(const M & model,
                                                  const Workspace<typename M::Scalar> & ws,
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
  using C = CoopLayout<M, GS>;
  constexpr int IPW = C::IPW, CPL = C::CPL;
  const size_t Bp = ws.Bp;
  const int N = prm.N;

  // element e of this instance's region `off` lives at sm[(off + e) * IPW]
  auto at = [&](int off, int e) -> S & { return sm[(size_t)(off + e) * IPW]; };

  auto stageStep = [&](int step) {
    const int stage = step % C::DEPTH;
    if(work && step >= 0)
    {
#pragma unroll
      for(int e0 = 0; e0 < C::STAGE; e0 += GS)
      {
        const int e = e0 + j;
        if(e < C::STAGE)
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
    }
    cpAsyncCommit();
  };

  // terminal value function: each lane brings in its own columns
  if(work)
  {
#pragma unroll
    for(int cc = 0; cc < CPL; cc++)
    {
      const int c = j + cc * GS;
      if(c < NX)
      {
        at(C::VX, c) = ws.vterm[(size_t)c * Bp + b];
#pragma unroll
        for(int r = 0; r < NX; r++) at(C::VXX, r + c * NX) = ws.vterm[(size_t)(NX + r + c * NX) * Bp + b];
      }
    }
  }
#pragma unroll
  for(int d = 1; d < C::DEPTH; d++) stageStep(N - d);

  dV0 = S(0);
  dV1 = S(0);
  S krn_num = S(0), krn_den = S(1);
  bool ok = true;
  S k_prev[NU];
#pragma unroll
  for(int a = 0; a < NU; a++) k_prev[a] = S(0);

  for(int i = N - 1; i >= 0; i--)
  {
    cpAsyncWait<C::DEPTH - 2>(); // this lane's share of block i has landed
    __syncwarp(); // [S1] ... and everyone else's; Vxx/Vx of the previous step visible; step i+1 fully retired
    const int blk = C::RING + (i % C::DEPTH) * C::STAGE;
    stageStep(i - (C::DEPTH - 1)); // refill the slot step i+1 just vacated

    const bool act = work && ok;
    // lane-local results that survive the barriers
    S Qu[NU], Quu[NU * NU], k[NU];
    S Qx_c[CPL], Qxx_c[CPL][NX], Qux_c[CPL][NU], K_c[CPL][NU];
    if(act)
    {
      S Fx[NX * NX], Fu[NX * NU], Vxx[NX * NX], Vx[NX];
#pragma unroll
      for(int d = 0; d < NX * NX; d++) Fx[d] = at(blk, L::FX + d);
#pragma unroll
      for(int d = 0; d < NX * NU; d++) Fu[d] = at(blk, L::FU + d);
#pragma unroll
      for(int d = 0; d < NX * NX; d++) Vxx[d] = at(C::VXX, d);
#pragma unroll
      for(int d = 0; d < NX; d++) Vx[d] = at(C::VX, d);

      // redundant small pieces: Qu, Tu = Fu^T Vxx, Quu                              (:386, :399)
#pragma unroll
      for(int a = 0; a < NU; a++)
      {
        S s = S(0);
#pragma unroll
        for(int r = 0; r < NX; r++) s += Fu[r + a * NX] * Vx[r];
        Qu[a] = at(blk, L::LU + a) + s;
      }
      S Tu[NU * NX];
#pragma unroll
      for(int q = 0; q < NX; q++)
#pragma unroll
        for(int a = 0; a < NU; a++)
        {
          S s = S(0);
#pragma unroll
          for(int r = 0; r < NX; r++) s += Fu[r + a * NX] * Vxx[r + q * NX];
          Tu[a + q * NU] = s;
        }
      S Quu_F[NU * NU];
#pragma unroll
      for(int c2 = 0; c2 < NU; c2++)
#pragma unroll
        for(int a = 0; a < NU; a++)
        {
          S s = S(0), sr = S(0);
#pragma unroll
          for(int r = 0; r < NX; r++)
          {
            s += Tu[a + r * NU] * Fu[r + c2 * NX];
            sr += (Tu[a + r * NU] + lambda * Fu[r + a * NX]) * Fu[r + c2 * NX];
          }
          const S luu = at(blk, L::LUU + a + c2 * NU);
          Quu[a + c2 * NU] = luu + s;
          // reg_type 2: Vxx_reg = Vxx + lambda I inside the product; reg_type 1: Quu + lambda I (:421-441)
          Quu_F[a + c2 * NU] = (prm.reg_type == 2) ? (luu + sr) : (luu + s);
        }
      if(prm.reg_type == 1)
      {
#pragma unroll
        for(int a = 0; a < NU; a++) Quu_F[a + a * NU] += lambda;
      }

      // column work: Qx_c, W_c = Vxx Fx(:,c), Qxx(:,c), Qux(:,c), Qux_reg(:,c)        (:388-408, :427)
      S Qux_reg_c[CPL][NU];
#pragma unroll
      for(int cc = 0; cc < CPL; cc++)
      {
        const int c = j + cc * GS;
        if(c < NX)
        {
          S fxc[NX];
#pragma unroll
          for(int r = 0; r < NX; r++) fxc[r] = at(blk, L::FX + r + c * NX); // lane-dependent column
          S s0 = S(0);
#pragma unroll
          for(int r = 0; r < NX; r++) s0 += fxc[r] * Vx[r];
          Qx_c[cc] = at(blk, L::LX + c) + s0;
          S W[NX];
#pragma unroll
          for(int r = 0; r < NX; r++)
          {
            S s = S(0);
#pragma unroll
            for(int q = 0; q < NX; q++) s += Vxx[r + q * NX] * fxc[q];
            W[r] = s;
          }
#pragma unroll
          for(int r = 0; r < NX; r++)
          {
            S s = S(0);
#pragma unroll
            for(int q = 0; q < NX; q++) s += Fx[q + r * NX] * W[q];
            Qxx_c[cc][r] = at(blk, L::LXX + r + c * NX) + s;
          }
#pragma unroll
          for(int a = 0; a < NU; a++)
          {
            S s = S(0), sr = S(0);
#pragma unroll
            for(int q = 0; q < NX; q++)
            {
              s += Tu[a + q * NU] * fxc[q];
              sr += (Tu[a + q * NU] + lambda * Fu[q + a * NX]) * fxc[q];
            }
            const S lxu = at(blk, L::LXU + c + a * NX);
            Qux_c[cc][a] = lxu + s;
            Qux_reg_c[cc][a] = (prm.reg_type == 2) ? (lxu + sr) : (lxu + s);
          }
        }
      }

      // gains (:448-517); every lane of the group factorises the same Quu_F => uniform verdict
      if constexpr(CONSTRAINED)
      {
        S lo[NU], hi[NU], init[NU];
#pragma unroll
        for(int a = 0; a < NU; a++)
        {
          const S uv = at(blk, L::SIZE + a);
          lo[a] = ws.u_lo[(size_t)i * NU + a] - uv; // input_limits_func_(t_i) (:470)
          hi[a] = ws.u_hi[(size_t)i * NU + a] - uv;
          init[a] = warmStartFromNextStep<M>(model, prm.t0, i, N) ? k_prev[a] : S(0);
        }
        BoxQPResult<S, NU> qp;
        boxQpSolve<S, NU>(Quu_F, Qu, lo, hi, init, qp);
        if(qp.retval < 0)
        {
          ok = false;
        }
        else
        {
#pragma unroll
          for(int a = 0; a < NU; a++) k[a] = qp.x[a];
          const int nf = qp.n_free;
#pragma unroll
          for(int cc = 0; cc < CPL; cc++)
          {
#pragma unroll
            for(int a = 0; a < NU; a++) K_c[cc][a] = S(0);
            S rhs[NU];
            for(int r = 0; r < nf; r++) rhs[r] = Qux_reg_c[cc][qp.free_idxs[r]];
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
            for(int r = 0; r < nf; r++) K_c[cc][qp.free_idxs[r]] = S(-1) * rhs[r];
          }
        }
      }
      else if constexpr(NU == 1)
      {
        if(Quu_F[0] <= S(0))
        {
          ok = false;
        }
        else
        {
          const S inv = S(1) / Quu_F[0];
          k[0] = -(Qu[0] * inv);
#pragma unroll
          for(int cc = 0; cc < CPL; cc++) K_c[cc][0] = -(Qux_reg_c[cc][0] * inv);
        }
      }
      else
      {
        if(!lltInPlace<S, NU>(Quu_F))
        {
          ok = false;
        }
        else
        {
          S invd[NU];
#pragma unroll
          for(int a = 0; a < NU; a++) invd[a] = S(1) / Quu_F[a + a * NU];
#pragma unroll
          for(int a = 0; a < NU; a++) k[a] = Qu[a];
          lltSolveInPlace<S, NU>(Quu_F, invd, k);
#pragma unroll
          for(int a = 0; a < NU; a++) k[a] = -k[a];
#pragma unroll
          for(int cc = 0; cc < CPL; cc++)
          {
#pragma unroll
            for(int a = 0; a < NU; a++) K_c[cc][a] = Qux_reg_c[cc][a];
            lltSolveInPlace<S, NU>(Quu_F, invd, K_c[cc]);
#pragma unroll
            for(int a = 0; a < NU; a++) K_c[cc][a] = -K_c[cc][a];
          }
        }
      }

      if(ok)
      {
#pragma unroll
        for(int cc = 0; cc < CPL; cc++)
        {
          const int c = j + cc * GS;
          if(c < NX)
          {
#pragma unroll
            for(int a = 0; a < NU; a++)
            {
              at(C::KFB, a + c * NU) = K_c[cc][a];
              at(C::QUX, a + c * NU) = Qux_c[cc][a];
            }
          }
        }
      }
    }
    __syncwarp(); // [S2] every column of K and Qux is in shared memory

    const bool act2 = work && ok;
    S Vx_new[CPL], Vn_c[CPL][NX];
    if(act2)
    {
      S K[NU * NX], Qux[NU * NX];
#pragma unroll
      for(int d = 0; d < NU * NX; d++) K[d] = at(C::KFB, d);
#pragma unroll
      for(int d = 0; d < NU * NX; d++) Qux[d] = at(C::QUX, d);

      // cost-to-go (:522-526)
      S Quuk[NU];
#pragma unroll
      for(int a = 0; a < NU; a++)
      {
        S s = S(0);
#pragma unroll
        for(int c2 = 0; c2 < NU; c2++) s += Quu[a + c2 * NU] * k[c2];
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
      S KtQuu[NX * NU];
#pragma unroll
      for(int c2 = 0; c2 < NU; c2++)
#pragma unroll
        for(int r = 0; r < NX; r++)
        {
          S s = S(0);
#pragma unroll
          for(int a = 0; a < NU; a++) s += K[a + r * NU] * Quu[a + c2 * NU];
          KtQuu[r + c2 * NX] = s;
        }
#pragma unroll
      for(int cc = 0; cc < CPL; cc++)
      {
        const int c = j + cc * GS;
        if(c < NX)
        {
          S s1 = S(0), s2 = S(0), s3 = S(0);
#pragma unroll
          for(int a = 0; a < NU; a++)
          {
            // row c of K^T Quu from the lane's own K column (c is lane dependent: no register indexing by it)
            S ktq = S(0);
#pragma unroll
            for(int a2 = 0; a2 < NU; a2++) ktq += K_c[cc][a2] * Quu[a2 + a * NU];
            s1 += ktq * k[a];
            s2 += K_c[cc][a] * Qu[a];
            s3 += Qux_c[cc][a] * k[a];
          }
          Vx_new[cc] = ((Qx_c[cc] + s1) + s2) + s3;
#pragma unroll
          for(int r = 0; r < NX; r++)
          {
            S t1 = S(0), t2 = S(0), t3 = S(0);
#pragma unroll
            for(int a = 0; a < NU; a++)
            {
              t1 += KtQuu[r + a * NX] * K_c[cc][a];
              t2 += K[a + r * NU] * Qux_c[cc][a];
              t3 += Qux[a + r * NU] * K_c[cc][a];
            }
            Vn_c[cc][r] = ((Qxx_c[cc][r] + t1) + t2) + t3;
            at(C::VN, r + c * NX) = Vn_c[cc][r];
          }
        }
      }
    }
    __syncwarp(); // [S3] unsymmetrised Vxx columns exchanged

    if(act2)
    {
#pragma unroll
      for(int cc = 0; cc < CPL; cc++)
      {
        const int c = j + cc * GS;
        if(c < NX)
        {
          at(C::VX, c) = Vx_new[cc];
#pragma unroll
          for(int r = 0; r < NX; r++) at(C::VXX, r + c * NX) = S(0.5) * (Vn_c[cc][r] + at(C::VN, c + r * NX));
          // gains of this step (:529-530)
#pragma unroll
          for(int a = 0; a < NU; a++) ws.kfb[((size_t)i * NU * NX + a + c * NU) * Bp + b] = K_c[cc][a];
        }
      }
      S kn = S(0), un = S(0);
#pragma unroll
      for(int a = 0; a < NU; a++)
      {
        if(j == 0) ws.kff[((size_t)i * NU + a) * Bp + b] = k[a];
        kn += k[a] * k[a];
        const S uv = at(blk, L::SIZE + a);
        un += uv * uv;
        k_prev[a] = k[a];
      }
      const S a_num = (NU == 1) ? fabs(k[0]) : sqrt(kn);
      const S a_den = ((NU == 1) ? fabs(at(blk, L::SIZE)) : sqrt(un)) + S(1);
      if(a_num * krn_den > krn_num * a_den)
      {
        krn_num = a_num;
        krn_den = a_den;
      }
    }
  }
  cpAsyncWait<0>();
  __syncwarp();
  k_rel_norm = krn_num / krn_den;
  return ok;
}

/** procOnce() Step 2 (DDPSolver.hpp:188-231) with GS lanes per instance. */
template<class M, int GS, bool CONSTRAINED>
__global__ void backward_coop_kernel(const __grid_constant__ M model,
                                     const __grid_constant__ Workspace<typename M::Scalar> ws,
                                     const __grid_constant__ SolverParams<typename M::Scalar> prm,
                                     int iter)
{
  pdlPrologue();
  using S = typename M::Scalar;
  using C = CoopLayout<M, GS>;
  constexpr unsigned kFull = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int g = lane / GS;
  const int j = lane % GS;
  S * sm = reinterpret_cast<S *>(smem_raw) + (size_t)warp * C::WARP_ELEMS + g;

  const int bg = (blockIdx.x * (blockDim.x >> 5) + warp) * C::IPW + g;
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
    const bool ok = backwardSweepCoop<M, GS, CONSTRAINED>(model, ws, prm, b, j, us, sm, need, lambda, sw_dV0, sw_dV1, sw_krn);
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
