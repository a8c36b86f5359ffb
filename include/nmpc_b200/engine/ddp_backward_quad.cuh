/* nmpc_b200 -- K2 for latency-bound batches: the W warps of a CTA (W = 4 ... 12, one matrix column each where possible)
 * split one 32-instance tile by COLUMNS.
 *
 * With one thread per instance a backward step is ~540 instructions (425 of them fp64, two issue cycles each)
 * along one warp's instruction stream, and a 4096-instance batch offers only 128 such warps to 592 warp
 * schedulers (ncu: 1500 cycles per horizon step, fp64 pipe of the one busy scheduler 55 % busy, the other three
 * schedulers of the SM idle).  Here lane l of EVERY warp of the CTA works on instance tile*32 + l, and warp w owns
 * columns w, w+W, ... of the n_x x n_x matrices (Vxx, Tx = Fx^T Vxx, Qxx, Vxx') plus the matching columns of
 * Tu = Fu^T Vxx, Qux, K.  The warps are spread over the four schedulers of the SM, each with its own fp64 pipe, and
 * exchange columns through shared memory at three block barriers per step (measured: 36 cycles per
 * STS / bar.sync / LDS round, tools/fp64_latency.cu):
 *
 *   phase A   Tx(:,c), Tu(:,c) for own columns c                     -> smem, barrier E1
 *   phase B   Qxx(:,c), Qux(:,c), Qx(c) for own c from ALL of Tx / Tu; Qu, Quu, its factorisation and k
 *             redundantly in every warp (same inputs, same arithmetic => same verdict); K(:,c), (K^T Quu)(c,:),
 *             Vx'(c)                                                   -> smem, barrier E2
 *   phase C   unsymmetrised Vxx'(:,c) from ALL of K, Qux, K^T Quu      -> smem, barrier E3
 *   then      Vxx(:,c) = 0.5 (Vxx'(:,c) + Vxx'(c,:)^T), Vx = all of Vx'
 *
 * Every scalar is computed by exactly the expression of ddp::backwardSweep (same operands, same order), so the
 * two variants agree bit for bit.  The step's derivative tile arrives by one bulk (TMA) copy per step into a
 * two-stage ring shared by the W warps; the wait on its mbarrier is issued one phase early so that its
 * latency overlaps phase C of the previous step.
 *
 * Reference: DDPSolver.hpp:188-231 (Step 2), :343-534 (backwardPass).
 */
#pragma once

#include "ddp_kernels.cuh"

namespace nmpc_b200
{
namespace ddp
{

template<class M>
struct QuadLayout
{
  static constexpr int NX = M::NX, NU = M::NU;
  using L = BlockLayout<NX, NU>;
  /** Warps that share one 32-instance tile: one matrix column per warp up to 12 warps (measured on the quadrotor,
      n_x = 12, B = 8192: 11.6 / 8.9 / 8.5 ms per 10 sweeps with 4 / 6 / 12 warps -- fewer instructions per warp and
      three warps per scheduler to hide each other's latency), never fewer than 4. */
  static constexpr int W = (NX >= 12) ? 12 : (NX < 4 ? 4 : NX);
  static constexpr int CPW = (NX + W - 1) / W; //!< matrix columns per warp
  // shared-memory regions, in elements of [32 lanes]
  static constexpr int RING = 0; //!< 2 stages of the derivative tile
  static constexpr int TX = RING + 2 * L::SIZE; //!< Tx [c + j*NX]
  static constexpr int TU = TX + NX * NX; //!< Tu [a + j*NU]
  static constexpr int QUX = TU + NU * NX; //!< Qux [a + j*NU]
  static constexpr int KFB = QUX + NU * NX; //!< K [a + j*NU]
  static constexpr int KTQ = KFB + NU * NX; //!< K^T Quu [j + c*NX]
  static constexpr int VN = KTQ + NU * NX; //!< unsymmetrised Vxx' [r + j*NX]
  static constexpr int VX = VN + NX * NX; //!< Vx' [j]
  /** Large tiles (n_x >= 8) keep the warp's own columns of Vxx and Qxx in shared memory instead of registers:
      Qxx(:,c) is parked in the VN region (its final home after the K terms are added), Vxx(:,c) in VXXC. */
  static constexpr bool kColsInSmem = (NX >= 8);
  static constexpr int VXXC = VX + NX; //!< own columns of Vxx [r + c*NX] (private to the owning warp)
  static constexpr int ELEMS = VXXC + (kColsInSmem ? NX * NX : 0);
  static constexpr size_t bytes()
  {
    return sizeof(typename M::Scalar) * (size_t)ELEMS * kTile + 2 * sizeof(unsigned long long) + 16;
  }
};

/** Non-blocking probe of an mbarrier phase: the predicate comes back as an int so that the (60-90 cycle) latency
    of the instruction overlaps whatever is issued before the result is consumed. */
__device__ __forceinline__ int mbarTryWait(unsigned long long * bar, unsigned parity)
{
  const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  int done;
  asm volatile("{\n"
               ".reg .pred p;\n"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
               "selp.b32 %0, 1, 0, p;\n"
               "}\n"
               : "=r"(done)
               : "r"(a), "r"(parity)
               : "memory");
  return done;
}

/** One column-split backwardPass() sweep.  All 128 threads execute every barrier; only lanes with `work` (and no
    factorisation failure so far) compute.  The return value is identical in all warps of a lane. */
template<class M, bool CONSTRAINED>
__device__ __forceinline__ bool backwardSweepQuad(const M & model,
                                                  const Workspace<typename M::Scalar> & ws,
                                                  const SolverParams<typename M::Scalar> & prm,
                                                  int b,
                                                  int lane,
                                                  int w,
                                                  const typename M::Scalar * __restrict__ us,
                                                  typename M::Scalar * __restrict__ sm,
                                                  unsigned long long * bars,
                                                  unsigned & parity,
                                                  bool work,
                                                  typename M::Scalar lambda,
                                                  typename M::Scalar & dV0,
                                                  typename M::Scalar & dV1,
                                                  typename M::Scalar & k_rel_norm)
{
  using S = typename M::Scalar;
  constexpr int NX = M::NX, NU = M::NU;
  using L = BlockLayout<NX, NU>;
  using Q = QuadLayout<M>;
  constexpr int CPW = Q::CPW;
  constexpr unsigned kStageBytes = (unsigned)(sizeof(S) * L::SIZE * kTile);
  const size_t Bp = ws.Bp;
  const int N = prm.N;

  // element e of region `off` for this lane's instance
  auto at = [&](int off, int e) -> S & { return sm[(size_t)(off + e) * kTile]; };

  // own columns of the terminal Vxx, all of Vx
  constexpr bool kSm = Q::kColsInSmem;
  S Vxx_c[kSm ? 1 : CPW][NX], Vx[NX];
#pragma unroll
  for(int d = 0; d < NX; d++) Vx[d] = ws.vterm[(size_t)d * Bp + b];
#pragma unroll
  for(int cc = 0; cc < CPW; cc++)
  {
    const int c = w + cc * Q::W;
#pragma unroll
    for(int r = 0; r < NX; r++)
    {
      const S v = (c < NX) ? ws.vterm[(size_t)(NX + r + c * NX) * Bp + b] : S(0);
      if constexpr(kSm)
      {
        if(c < NX) at(Q::VXXC, r + c * NX) = v;
      }
      else
        Vxx_c[cc][r] = v;
    }
  }

  dV0 = S(0);
  dV1 = S(0);
  S krn_num = S(0), krn_den = S(1);
  bool ok = true;
  S k_prev[NU];
#pragma unroll
  for(int a = 0; a < NU; a++) k_prev[a] = S(0);
  // u_i for the termination test / input limits, fetched TWO steps ahead: one step of this kernel is shorter than
  // an HBM round trip (ncu: 27 % of all stall samples sat on the consumer of a one-step-ahead load)
  S u_cur[NU], u_nxt[NU], u_nx2[NU];
#pragma unroll
  for(int a = 0; a < NU; a++)
  {
    u_cur[a] = us[((size_t)(N - 1) * NU + a) * Bp + b];
    u_nxt[a] = us[((size_t)(N > 1 ? N - 2 : 0) * NU + a) * Bp + b];
  }

  const S * const tile0 = ws.deriv + derivTileOffset<L::SIZE>(0, b - lane, ws.Bp); // this CTA's tile, step 0
  const size_t step_stride = (size_t)(ws.Bp / kTile) * L::SIZE * kTile;
  S * const ring0 = sm - lane; // lane-independent base of the CTA's shared memory
  auto stageStep = [&](int stage, int step) {
    if(threadIdx.x == 0)
    {
      mbarExpectTx(&bars[stage], kStageBytes);
      bulkCopyG2S(ring0 + (size_t)(Q::RING + stage * L::SIZE) * kTile, tile0 + (size_t)step * step_stride, kStageBytes,
                  &bars[stage]);
    }
  };

  __syncthreads(); // the previous sweep's readers are done with the ring and the exchange regions
  stageStep(0, N - 1);
  if(N > 1) stageStep(1, N - 2);
  int stage = 0;
  int landed = 0; // result of the early probe of the current stage's mbarrier

  for(int i = N - 1; i >= 0; i--)
  {
    const int blk = Q::RING + stage * L::SIZE;
    {
      const int ip = (i > 1) ? i - 2 : 0;
#pragma unroll
      for(int a = 0; a < NU; a++) u_nx2[a] = us[((size_t)ip * NU + a) * Bp + b];
    }
    if(!landed) mbarWait(&bars[stage], (parity >> stage) & 1u);
    parity ^= (1u << stage);

    const bool act = work && ok;
    S Qx_c[CPW], Qxx_c[kSm ? 1 : CPW][NX], Qux_c[CPW][NU], K_c[CPW][NU];
    S Qu[NU], Quu[NU * NU], k[NU];

    // ---------------------------------------------------------------- phase A: Tx(:,c), Tu(:,c)   (:386-408)
    if(act)
    {
      // small tiles hold Fx / Fu in registers; large ones read the operands straight from the staged tile
      S Fx[kSm ? 1 : NX * NX], Fu[kSm ? 1 : NX * NU];
      if constexpr(!kSm)
      {
#pragma unroll
        for(int d = 0; d < NX * NX; d++) Fx[d] = at(blk, L::FX + d);
#pragma unroll
        for(int d = 0; d < NX * NU; d++) Fu[d] = at(blk, L::FU + d);
      }
      auto fx = [&](int d) -> S { if constexpr(kSm) return at(blk, L::FX + d); else return Fx[d]; };
      auto fu = [&](int d) -> S { if constexpr(kSm) return at(blk, L::FU + d); else return Fu[d]; };
#pragma unroll
      for(int cc = 0; cc < CPW; cc++)
      {
        const int c = w + cc * Q::W;
        if(c < NX)
        {
          S vcol[NX];
#pragma unroll
          for(int r = 0; r < NX; r++)
          {
            if constexpr(kSm)
              vcol[r] = at(Q::VXXC, r + c * NX);
            else
              vcol[r] = Vxx_c[cc][r];
          }
#pragma unroll
          for(int a = 0; a < NU; a++)
          {
            S s = S(0);
#pragma unroll
            for(int r = 0; r < NX; r++) s += fu(r + a * NX) * vcol[r];
            at(Q::TU, a + c * NU) = s;
          }
#pragma unroll
          for(int q = 0; q < NX; q++)
          {
            S s = S(0);
#pragma unroll
            for(int r = 0; r < NX; r++) s += fx(r + q * NX) * vcol[r];
            at(Q::TX, q + c * NX) = s;
          }
        }
      }
    }
    __syncthreads(); // [E1] all of Tx, Tu visible; every warp is done with the other ring stage (step i+1)
    if(i > 0 && i < N - 1) stageStep(stage ^ 1, i - 1);

    // ---------------------------------------------------------------- phase B
    if(act)
    {
      S Fu_r[kSm ? 1 : NX * NU], Tu_r[kSm ? 1 : NU * NX];
      if constexpr(!kSm)
      {
#pragma unroll
        for(int d = 0; d < NX * NU; d++) Fu_r[d] = at(blk, L::FU + d);
#pragma unroll
        for(int d = 0; d < NU * NX; d++) Tu_r[d] = at(Q::TU, d);
      }
      auto Fu = [&](int d) -> S { if constexpr(kSm) return at(blk, L::FU + d); else return Fu_r[d]; };
      auto Tu = [&](int d) -> S { if constexpr(kSm) return at(Q::TU, d); else return Tu_r[d]; };
      // reg_type 2: Tu_reg = Tu + lambda Fu^T, formed on the fly
      auto Tur = [&](int a, int q) -> S { return Tu(a + q * NU) + lambda * Fu(q + a * NX); };

      // Qu = Lu + Fu^T Vx (redundant)                                                         (:386)
#pragma unroll
      for(int a = 0; a < NU; a++)
      {
        S s = S(0);
#pragma unroll
        for(int r = 0; r < NX; r++) s += Fu(r + a * NX) * Vx[r];
        Qu[a] = at(blk, L::LU + a) + s;
      }
      // Quu = Luu + Tu Fu (redundant)                                                         (:399)
#pragma unroll
      for(int c2 = 0; c2 < NU; c2++)
#pragma unroll
        for(int a = 0; a < NU; a++)
        {
          S s = S(0);
#pragma unroll
          for(int r = 0; r < NX; r++) s += Tu(a + r * NU) * Fu(r + c2 * NX);
          Quu[a + c2 * NU] = at(blk, L::LUU + a + c2 * NU) + s;
        }
      // regularisation (:421-441)
      S Quu_F[NU * NU];
      if(prm.reg_type == 2)
      {
#pragma unroll
        for(int c2 = 0; c2 < NU; c2++)
#pragma unroll
          for(int a = 0; a < NU; a++)
          {
            S s = S(0);
#pragma unroll
            for(int r = 0; r < NX; r++) s += Tur(a, r) * Fu(r + c2 * NX);
            Quu_F[a + c2 * NU] = at(blk, L::LUU + a + c2 * NU) + s;
          }
      }
      else
      {
#pragma unroll
        for(int d = 0; d < NU * NU; d++) Quu_F[d] = Quu[d];
        if(prm.reg_type == 1)
        {
#pragma unroll
          for(int a = 0; a < NU; a++) Quu_F[a + a * NU] += lambda;
        }
      }

      // own columns: Qx(c), Qxx(:,c), Qux(:,c), Qux_reg(:,c)                                  (:388-408)
      S Qux_reg_c[CPW][NU];
#pragma unroll
      for(int cc = 0; cc < CPW; cc++)
      {
        const int c = w + cc * Q::W;
        if(c < NX)
        {
          S fxc[NX];
#pragma unroll
          for(int r = 0; r < NX; r++) fxc[r] = at(blk, L::FX + r + c * NX);
          {
            S s = S(0);
#pragma unroll
            for(int r = 0; r < NX; r++) s += fxc[r] * Vx[r];
            Qx_c[cc] = at(blk, L::LX + c) + s;
          }
#pragma unroll
          for(int a = 0; a < NU; a++)
          {
            S s = S(0);
#pragma unroll
            for(int r = 0; r < NX; r++) s += Tu(a + r * NU) * fxc[r];
            const S lxu = at(blk, L::LXU + c + a * NX);
            Qux_c[cc][a] = lxu + s;
            if(prm.reg_type == 2)
            {
              S sr = S(0);
#pragma unroll
              for(int r = 0; r < NX; r++) sr += Tur(a, r) * fxc[r];
              Qux_reg_c[cc][a] = lxu + sr;
            }
            else
            {
              Qux_reg_c[cc][a] = Qux_c[cc][a];
            }
          }
#pragma unroll
          for(int q = 0; q < NX; q++)
          {
            S s = S(0);
#pragma unroll
            for(int r = 0; r < NX; r++) s += at(Q::TX, q + r * NX) * fxc[r];
            const S qxx = at(blk, L::LXX + q + c * NX) + s;
            if constexpr(kSm)
              at(Q::VN, q + c * NX) = qxx; // parked in its final home; nobody reads this column before [E3]
            else
              Qxx_c[cc][q] = qxx;
          }
        }
      }

      // gains (:448-517); every warp factorises the same Quu_F => the same verdict in all four
      if constexpr(CONSTRAINED)
      {
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
        }
        else
        {
#pragma unroll
          for(int a = 0; a < NU; a++) k[a] = qp.x[a];
          const int nf = qp.n_free;
#pragma unroll
          for(int cc = 0; cc < CPW; cc++)
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
          for(int cc = 0; cc < CPW; cc++) K_c[cc][0] = -(Qux_reg_c[cc][0] * inv);
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
          for(int cc = 0; cc < CPW; cc++)
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
        // cost-to-go, scalar part (:522); identical in every warp, warp 0's copy is the one written out
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
        // own rows of K^T Quu, own entries of Vx' = Qx + K^T Quu k + K^T Qu + Qux^T k            (:523)
#pragma unroll
        for(int cc = 0; cc < CPW; cc++)
        {
          const int c = w + cc * Q::W;
          if(c < NX)
          {
            S ktq[NU];
#pragma unroll
            for(int c2 = 0; c2 < NU; c2++)
            {
              S s = S(0);
#pragma unroll
              for(int a = 0; a < NU; a++) s += K_c[cc][a] * Quu[a + c2 * NU];
              ktq[c2] = s;
              at(Q::KTQ, c + c2 * NX) = s;
            }
            S s1 = S(0), s2 = S(0), s3 = S(0);
#pragma unroll
            for(int a = 0; a < NU; a++)
            {
              s1 += ktq[a] * k[a];
              s2 += K_c[cc][a] * Qu[a];
              s3 += Qux_c[cc][a] * k[a];
              at(Q::KFB, a + c * NU) = K_c[cc][a];
              at(Q::QUX, a + c * NU) = Qux_c[cc][a];
            }
            at(Q::VX, c) = ((Qx_c[cc] + s1) + s2) + s3;
          }
        }
      }
    }
    __syncthreads(); // [E2] all of K, Qux, K^T Quu, Vx' visible

    // ---------------------------------------------------------------- phase C: Vxx'(:,c)          (:524-526)
    const bool act2 = work && ok;
    // probe the next step's tile now: the answer is consumed at the top of the next iteration
    const int next_stage = stage ^ 1;
    landed = (i > 0) ? mbarTryWait(&bars[next_stage], (parity >> next_stage) & 1u) : 0;
    S Vn_c[kSm ? 1 : CPW][NX];
    if(act2)
    {
      S K[NU * NX], Qux[NU * NX], KtQuu[NX * NU];
#pragma unroll
      for(int d = 0; d < NU * NX; d++) K[d] = at(Q::KFB, d);
#pragma unroll
      for(int d = 0; d < NU * NX; d++) Qux[d] = at(Q::QUX, d);
#pragma unroll
      for(int d = 0; d < NU * NX; d++) KtQuu[d] = at(Q::KTQ, d);
#pragma unroll
      for(int cc = 0; cc < CPW; cc++)
      {
        const int c = w + cc * Q::W;
        if(c < NX)
        {
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
            S qxx;
            if constexpr(kSm)
              qxx = at(Q::VN, r + c * NX);
            else
              qxx = Qxx_c[cc][r];
            const S vn = ((qxx + t1) + t2) + t3;
            at(Q::VN, r + c * NX) = vn;
            if constexpr(!kSm) Vn_c[cc][r] = vn;
          }
          // gains of this step (:529-530)
#pragma unroll
          for(int a = 0; a < NU; a++) ws.kfb[((size_t)i * NU * NX + a + c * NU) * Bp + b] = K_c[cc][a];
        }
      }
      S kn = S(0), un = S(0);
#pragma unroll
      for(int a = 0; a < NU; a++)
      {
        if(w == 0) ws.kff[((size_t)i * NU + a) * Bp + b] = k[a];
        kn += k[a] * k[a];
        const S uv = u_cur[a];
        un += uv * uv;
        k_prev[a] = k[a];
      }
      // |k| / (|u| + 1) > num / den  <=>  |k| * den > num * (|u| + 1)
      const S a_num = (NU == 1) ? fabs(k[0]) : sqrt(kn);
      const S a_den = ((NU == 1) ? fabs(u_cur[0]) : sqrt(un)) + S(1);
      if(a_num * krn_den > krn_num * a_den)
      {
        krn_num = a_num;
        krn_den = a_den;
      }
    }
    __syncthreads(); // [E3] unsymmetrised columns exchanged

    if(act2)
    {
#pragma unroll
      for(int cc = 0; cc < CPW; cc++)
      {
        const int c = w + cc * Q::W;
        if(c < NX)
        {
#pragma unroll
          for(int r = 0; r < NX; r++)
          {
            if constexpr(kSm)
              at(Q::VXXC, r + c * NX) = S(0.5) * (at(Q::VN, r + c * NX) + at(Q::VN, c + r * NX));
            else
              Vxx_c[cc][r] = S(0.5) * (Vn_c[cc][r] + at(Q::VN, c + r * NX));
          }
        }
      }
#pragma unroll
      for(int d = 0; d < NX; d++) Vx[d] = at(Q::VX, d);
    }
#pragma unroll
    for(int a = 0; a < NU; a++)
    {
      u_cur[a] = u_nxt[a];
      u_nxt[a] = u_nx2[a];
    }
    stage ^= 1;
  }
  k_rel_norm = krn_num / krn_den;
  return ok;
}

/** procOnce() Step 2 (DDPSolver.hpp:188-231), W warps per 32-instance tile. */
template<class M, bool CONSTRAINED>
__global__ void __launch_bounds__(QuadLayout<M>::W * 32) backward_quad_kernel(const __grid_constant__ M model,
                                                                        const __grid_constant__ Workspace<typename M::Scalar> ws,
                                                                        const __grid_constant__ SolverParams<typename M::Scalar> prm,
                                                                        int iter)
{
  pdlPrologue();
  using S = typename M::Scalar;
  using Q = QuadLayout<M>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int w = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  S * sm = reinterpret_cast<S *>(smem_raw) + lane;
  unsigned long long * bars = reinterpret_cast<unsigned long long *>(smem_raw + sizeof(S) * (size_t)Q::ELEMS * kTile);
  if(threadIdx.x == 0)
  {
    mbarInit(&bars[0], 1);
    mbarInit(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  unsigned parity = 0u;

  const int b = blockIdx.x * kTile + lane; // ws.Bp is a multiple of 128: padded lanes read valid memory, never write
  const bool live = (b < ws.B) && (ws.status[b < ws.B ? b : 0] == 0);

  S lambda = live ? ws.lambda[b] : S(0);
  S dlambda = live ? ws.dlambda[b] : S(0);
  const S * us = ws.u[live ? ws.sel[b] : 0];
  int n_bwd = live ? ws.n_bwd[b] : 0;
  S dV0 = S(0), dV1 = S(0), k_rel_norm = S(0);
  bool need = live;
  bool failed = false;
  // `need` of a lane is the same in all warps, so the trip count is uniform over the CTA
  while(__syncthreads_or(need))
  {
    if(need) n_bwd++;
    // results are kept only for instances that needed this sweep: one that is waiting for its tile mates' lambda retry
    // keeps the dV / k_rel_norm of its own successful sweep
    S sw_dV0 = S(0), sw_dV1 = S(0), sw_krn = S(0);
    const bool ok = backwardSweepQuad<M, CONSTRAINED>(model, ws, prm, b, lane, w, us, sm, bars, parity, need, lambda, sw_dV0, sw_dV1,
                                                      sw_krn);
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
  if(!live || w != 0) return;
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
} // namespace ddp
} // namespace nmpc_b200
