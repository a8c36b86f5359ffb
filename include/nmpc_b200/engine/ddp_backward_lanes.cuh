/* nmpc_b200 -- K1 + K2 with G LANES PER INSTANCE (procOnce() Steps 1-2, DDPSolver.hpp:157-231, :343-534).
 *
 * A small batch (BASELINE.json configs[1]: 4096 cart-pole instances) is 128 warps for the 592 warp schedulers of a
 * B200 when one thread owns one instance, and the Riccati sweep is then bound by ONE warp's fp64 issue rate
 * (425 fp64 instructions per step at 2.25-2.5 cycles each, profiles/r1_v6_stage_kernels.md).  Here the G = 4 lanes
 * of a group share one instance (8 instances per warp, 512 consumer warps at B = 4096):
 *
 *   every lane (replicated, bit-identical in the group)   z = Vxx Fu, Quu = Luu + Fu^T z, Qu = Lu + Fu^T Vx,
 *                                                         regularisation, factorisation of Quu_F, k
 *   lane j (column j of the n_x x n_x matrices)           W = Vxx Fx(:,j), Qxx(:,j) = Lxx(:,j) + Fx^T W,
 *                                                         Qux(:,j) = Lxu(j,:)^T + z^T Fx(:,j), Qx(j), K(:,j)
 *   exchange 1 (all-gather in the group)                  Qux, K
 *   lane j                                                Vxx'(:,j) = Qxx(:,j) + K^T Quu K(:,j) + K^T Qux(:,j) + Qux^T K(:,j),
 *                                                         Vx'(j)
 *   exchange 2 (all-gather in the group)                  Vxx' (unsymmetrised), Vx'; every lane then symmetrises
 *
 * so a lane issues ~115 fp64 instructions per step instead of ~425.  The two exchanges per step are warp shuffles
 * (XchShfl: every lane reads the four group lanes in turn) or a round trip through a private shared-memory scratch
 * with one group barrier (XchSmem); both are compiled, the engine picks one (measured: profiles/r2_*).  The products
 * associate as Fx^T (Vxx Fx) instead of the reference's (Fx^T Vxx) Fx, and Vxx' is symmetric by construction of the
 * inputs: results agree with the reference to rounding (tolerances in tests/test_ddp_gpu.py), not bit for bit.
 *
 * As in ddp_backward_fused.cuh there is no K1: PRODUCER warps (one thread per instance of the CTA's 32-instance
 * tile, P warps taking every P-th step) evaluate the functor's derivatives straight into a shared-memory ring.  The
 * tile is pair-interleaved ([element / 2][instance][2]) so that both sides move 16 bytes per access, and it carries
 * u_i for the termination test and the input limits.
 */
#pragma once

#include "ddp_kernels.cuh"

namespace nmpc_b200
{
namespace ddp
{
constexpr int kLaneDepth = 4; //!< ring stages between the producer warps and the consumer warps
/** Elements between two pair-rows of a staged tile: 32 instances x 2 elements, plus one pair of padding.  Lane j of a
    group reads ITS column's rows (rows 2 j + q for n_x = 4): with an unpadded 512-byte row the four lanes of a group
    hit the same banks (ncu: 2.7 M of the 6.3 M shared-load wavefronts of a sweep were bank conflicts); with 528 bytes
    the eight lanes of a quarter-warp access touch eight different 16-byte bank groups. */
constexpr int kLaneRow = 2 * kTile + 2;
/** Elements of one ring stage for a tile of SIZE elements per instance. */
constexpr size_t laneStageElems(int size)
{
  return (size_t)(size / 2) * kLaneRow;
}

constexpr int evenUp(int v)
{
  return (v + 1) & ~1;
}

template<class S>
struct Vec2;
template<>
struct Vec2<double>
{
  using type = double2;
};
template<>
struct Vec2<float>
{
  using type = float2;
};

/** One step's tile for one instance; every block starts at an even element.  LXUT is Lxu transposed ([a + j * NU]) so
    that the n_u entries lane j needs are contiguous. */
template<int NX, int NU>
struct LaneTile
{
  static constexpr int FX = 0;
  static constexpr int LXX = FX + evenUp(NX * NX);
  static constexpr int FU = LXX + evenUp(NX * NX);
  static constexpr int LXUT = FU + evenUp(NX * NU);
  static constexpr int LX = LXUT + evenUp(NX * NU);
  static constexpr int LU = LX + evenUp(NX);
  static constexpr int LUU = LU + evenUp(NU);
  static constexpr int U = LUU + evenUp(NU * NU);
  static constexpr int SIZE = U + evenUp(NU);
};

/** Element e of the instance whose tile column starts at `tl` (= stage base + 2 * instance). */
template<class S>
__device__ __forceinline__ S tileElem(const S * tl, int e)
{
  return tl[(size_t)(e >> 1) * kLaneRow + (e & 1)];
}

/** CNT consecutive elements from e0; PAIRS: e0 is even, so pairs are read with one 2-element access. */
template<class S, int CNT, bool PAIRS>
__device__ __forceinline__ void tileLoad(const S * tl, int e0, S * out)
{
  if constexpr(PAIRS)
  {
    using V = typename Vec2<S>::type;
#pragma unroll
    for(int q = 0; q < CNT / 2; q++)
    {
      const V v = *reinterpret_cast<const V *>(tl + (size_t)((e0 >> 1) + q) * kLaneRow);
      out[2 * q] = v.x;
      out[2 * q + 1] = v.y;
    }
    if constexpr(CNT % 2 == 1) out[CNT - 1] = tileElem<S>(tl, e0 + CNT - 1);
  }
  else
  {
#pragma unroll
    for(int q = 0; q < CNT; q++) out[q] = tileElem<S>(tl, e0 + q);
  }
}

/** All-gather inside a group of G lanes through warp shuffles: all[c * CNT + q] = mine[q] of group lane c. */
template<class S, int G>
struct XchShfl
{
  static constexpr bool kSmem = false;
  template<int CNT>
  __device__ __forceinline__ static void gather(S *, unsigned gmask, int lane, int, const S * mine, S * all)
  {
    const int base = lane & ~(G - 1);
#pragma unroll
    for(int c = 0; c < G; c++)
#pragma unroll
      for(int q = 0; q < CNT; q++) all[c * CNT + q] = __shfl_sync(gmask, mine[q], base + c);
  }
};

/** The same through the group's private scratch in shared memory (one group barrier).  The caller alternates two
    scratch regions, so a region is never rewritten before every lane has read it (a barrier of the OTHER exchange
    lies between a read and the next write). */
template<class S, int G>
struct XchSmem
{
  static constexpr bool kSmem = true;
  template<int CNT>
  __device__ __forceinline__ static void gather(S * scratch, unsigned gmask, int, int j, const S * mine, S * all)
  {
    constexpr int C2 = evenUp(CNT); // per-lane chunk, kept even for the 2-element accesses
    using V = typename Vec2<S>::type;
    S * my = scratch + j * C2;
#pragma unroll
    for(int q = 0; q < CNT / 2; q++)
    {
      V v;
      v.x = mine[2 * q];
      v.y = mine[2 * q + 1];
      *reinterpret_cast<V *>(my + 2 * q) = v;
    }
    if constexpr(CNT % 2 == 1) my[CNT - 1] = mine[CNT - 1];
    __syncwarp(gmask); // gmask is the full warp in laneSweep: one WARPSYNC, no MATCH / REDUX
#pragma unroll
    for(int c = 0; c < G; c++)
    {
#pragma unroll
      for(int q = 0; q < CNT / 2; q++)
      {
        const V v = *reinterpret_cast<const V *>(scratch + c * C2 + 2 * q);
        all[c * CNT + 2 * q] = v.x;
        all[c * CNT + 2 * q + 1] = v.y;
      }
      if constexpr(CNT % 2 == 1) all[c * CNT + CNT - 1] = scratch[c * C2 + CNT - 1];
    }
  }
};

template<class M>
struct LaneLayout
{
  using S = typename M::Scalar;
  static constexpr int NX = M::NX, NU = M::NU;
  static constexpr int G = (NX <= 2) ? 2 : 4; //!< lanes per instance; lane j owns column j (n_x <= G)
  static constexpr int IPW = 32 / G; //!< instances per consumer warp
  static constexpr int CW = G; //!< consumer warps per 32-instance tile
  using T = LaneTile<NX, NU>;
  static constexpr int X1 = 2 * NU; //!< exchange 1 per lane: Qux(:,j), K(:,j)
  static constexpr int X2 = NX + 1; //!< exchange 2 per lane: Vxx'(:,j), Vx'(j)
  // scratch strides in elements: 8 modulo 16 (64 bytes modulo 128 for fp64), so that the two instances a quarter-warp
  // access covers land in different halves of the banks, for the lanes' writes and for the broadcast reads alike
  static constexpr int strideFor(int n)
  {
    return n <= 8 ? 8 : ((n - 8 + 15) / 16) * 16 + 8;
  }
  static constexpr int X1S = strideFor(G * evenUp(X1));
  static constexpr int X2S = strideFor(G * evenUp(X2));
  static constexpr size_t ringElems()
  {
    return (size_t)kLaneDepth * laneStageElems(T::SIZE);
  }
  static constexpr size_t scratchElems()
  {
    return (size_t)kTile * (X1S + X2S);
  }
  /** Shared memory of one 32-instance tile (ring, exchange scratch, mbarriers), a multiple of 128 bytes. */
  static constexpr size_t bytes()
  {
    return ((sizeof(S) * (ringElems() + scratchElems()) + sizeof(unsigned long long) * 2 * kLaneDepth + 127) / 128) * 128;
  }
};

/** Producer warp `p` of `P`: tiles of the steps whose global fill index is congruent to p, for one sweep. */
template<class M, int P>
__device__ __forceinline__ void produceSweepLanes(const M & model_in_constant_bank,
                                                  const Workspace<typename M::Scalar> & ws,
                                                  const SolverParams<typename M::Scalar> & prm,
                                                  int b,
                                                  int t,
                                                  int p,
                                                  const typename M::Scalar * xs,
                                                  const typename M::Scalar * us,
                                                  typename M::Scalar * ring,
                                                  unsigned long long * full,
                                                  unsigned long long * empty,
                                                  unsigned fill_base)
{
  using S = typename M::Scalar;
  constexpr int NX = M::NX, NU = M::NU;
  using T = LaneTile<NX, NU>;
  using V = typename Vec2<S>::type;
  using LM = typename LatencyOf<M>::type;
  const LM model(model_in_constant_bank);
  const S t0 = prm.t0;
  const size_t Bp = ws.Bp;
  const int N = prm.N;

  // this warp's first fill of the sweep
  int f = (int)((P + p - (int)(fill_base % P)) % P);
  // (x_i, u_i) are fetched two of this warp's steps ahead of their use
  S xb[3][NX], ub[3][NU];
  auto load = [&](int slot, int ff) {
    int i = N - 1 - ff;
    i = i > 0 ? i : 0;
#pragma unroll
    for(int d = 0; d < NX; d++) xb[slot][d] = xs[((size_t)i * NX + d) * Bp + b];
#pragma unroll
    for(int d = 0; d < NU; d++) ub[slot][d] = us[((size_t)i * NU + d) * Bp + b];
  };
  load(0, f);
  load(1, f + P);
  for(; f < N; f += P)
  {
    const int i = N - 1 - f;
    load(2, f + 2 * P);
    Matrix<S, NX, 1> x;
    Matrix<S, NU, 1> u;
#pragma unroll
    for(int d = 0; d < NX; d++) x[d] = xb[0][d];
#pragma unroll
    for(int d = 0; d < NU; d++) u[d] = ub[0][d];

    Matrix<S, NX, NX> Fx, Lxx;
    Matrix<S, NX, NU> Fu, Lxu;
    Matrix<S, NX, 1> Lx;
    Matrix<S, NU, 1> Lu;
    Matrix<S, NU, NU> Luu;
    linearizeStep<LM>(model, t0 + i * model.dt(), x, u, Fx, Fu, Lx, Lu, Lxx, Luu, Lxu);

    S v[T::SIZE];
#pragma unroll
    for(int e = 0; e < T::SIZE; e++) v[e] = S(0);
#pragma unroll
    for(int d = 0; d < NX * NX; d++)
    {
      v[T::FX + d] = Fx.d[d];
      v[T::LXX + d] = Lxx.d[d];
    }
#pragma unroll
    for(int d = 0; d < NX * NU; d++) v[T::FU + d] = Fu.d[d];
#pragma unroll
    for(int j = 0; j < NX; j++)
#pragma unroll
      for(int a = 0; a < NU; a++) v[T::LXUT + a + j * NU] = Lxu(j, a);
#pragma unroll
    for(int d = 0; d < NX; d++) v[T::LX + d] = Lx.d[d];
#pragma unroll
    for(int d = 0; d < NU; d++)
    {
      v[T::LU + d] = Lu.d[d];
      v[T::U + d] = u[d];
    }
#pragma unroll
    for(int d = 0; d < NU * NU; d++) v[T::LUU + d] = Luu.d[d];

    const unsigned fg = fill_base + (unsigned)f;
    const unsigned st = fg % kLaneDepth;
    if(fg >= (unsigned)kLaneDepth) mbarWait(&empty[st], ((fg / kLaneDepth) - 1u) & 1u); // the consumers are done with it
    S * tl = ring + (size_t)st * laneStageElems(T::SIZE) + 2 * t;
#pragma unroll
    for(int q = 0; q < T::SIZE / 2; q++)
    {
      V w;
      w.x = v[2 * q];
      w.y = v[2 * q + 1];
      *reinterpret_cast<V *>(tl + (size_t)q * kLaneRow) = w;
    }
    mbarArrive(&full[st]); // release: this lane's column of the tile

#pragma unroll
    for(int d = 0; d < NX; d++)
    {
      xb[0][d] = xb[1][d];
      xb[1][d] = xb[2][d];
    }
#pragma unroll
    for(int d = 0; d < NU; d++)
    {
      ub[0][d] = ub[1][d];
      ub[1][d] = ub[2][d];
    }
  }
}

/** Non-blocking probe of an mbarrier phase (the blocking wait is mbarWait). */
__device__ __forceinline__ bool mbarTestWait(unsigned long long * bar, unsigned parity)
{
  const unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(bar));
  unsigned done;
  asm volatile("{\n"
               ".reg .pred p;\n"
               "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
               "selp.u32 %0, 1, 0, p;\n"
               "}\n"
               : "=r"(done)
               : "r"(a), "r"(parity)
               : "memory");
  return done != 0;
}

/** What lane j keeps of one step's tile. */
template<class M>
struct LaneTileRegs
{
  using S = typename M::Scalar;
  S Fx[M::NX * M::NX], Fu[M::NX * M::NU], Fxj[M::NX], Lxxj[M::NX], Lxuj[M::NU], Lu[M::NU], Luu[M::NU * M::NU], u[M::NU];
  S Lxj;
};

template<class M>
__device__ __forceinline__ void loadLaneTile(const typename M::Scalar * tl, int jj, LaneTileRegs<M> & T)
{
  using S = typename M::Scalar;
  constexpr int NX = M::NX, NU = M::NU;
  using TL = LaneTile<NX, NU>;
  constexpr bool kPairs = (NX % 2 == 0); // lane j's blocks start at an even element
  constexpr bool kPairsU = (NU % 2 == 0);
  tileLoad<S, NX * NX, true>(tl, TL::FX, T.Fx);
  tileLoad<S, NX * NU, true>(tl, TL::FU, T.Fu);
  tileLoad<S, NX, kPairs>(tl, TL::FX + jj * NX, T.Fxj);
  tileLoad<S, NX, kPairs>(tl, TL::LXX + jj * NX, T.Lxxj);
  tileLoad<S, NU, kPairsU>(tl, TL::LXUT + jj * NU, T.Lxuj);
  tileLoad<S, NU, true>(tl, TL::LU, T.Lu);
  tileLoad<S, NU * NU, true>(tl, TL::LUU, T.Luu);
  tileLoad<S, NU, true>(tl, TL::U, T.u);
  T.Lxj = tileElem<S>(tl, TL::LX + jj);
}

/** One backwardPass() sweep (DDPSolver.hpp:343-534) for the instance this lane's group owns.  EVERY lane of the warp
    executes every step, exchange and barrier (full-mask warp barriers: a partial mask costs a MATCH / REDUX sequence
    per barrier, ~95 cycles in profiles/r2_lanes_a); an instance that does not need the sweep, or whose factorisation
    has failed, keeps computing on whatever it has and stores nothing.  Software pipeline: the tile of step i-1 is
    copied to registers between the two exchanges of step i, its mbarrier probed at the top of step i.
    Returns false when the factorisation of Quu_F failed (LLT NumericalIssue, :500-508) or BoxQP reported an error. */
template<class M, bool CONSTRAINED, bool REG2, class XCH>
__device__ __forceinline__ bool laneSweep(const M & model,
                                          const Workspace<typename M::Scalar> & ws,
                                          const SolverParams<typename M::Scalar> & prm,
                                          int b,
                                          int t,
                                          int lane,
                                          int j,
                                          int jj,
                                          const typename M::Scalar * xs,
                                          const typename M::Scalar * ring,
                                          typename M::Scalar * x1,
                                          typename M::Scalar * x2,
                                          unsigned long long * full,
                                          unsigned long long * empty,
                                          unsigned & fill,
                                          bool need,
                                          typename M::Scalar lambda,
                                          typename M::Scalar & dV0_out,
                                          typename M::Scalar & dV1_out,
                                          typename M::Scalar & k_rel_norm_out)
{
  using S = typename M::Scalar;
  constexpr int NX = M::NX, NU = M::NU;
  using LL = LaneLayout<M>;
  using TL = typename LL::T;
  constexpr int G = LL::G;
  constexpr unsigned kFull = 0xffffffffu;
  const size_t Bp = ws.Bp;
  const int N = prm.N;
  // reg_type 1: Quu_F = Quu + lambda I; reg_type 2 (REG2): Vxx_reg = Vxx + lambda I; anything else: no regularisation
  const S lambda_uu = (!REG2 && prm.reg_type == 1) ? lambda : S(0);

  S Vx[NX], Vxx[NX * NX];
  {
    Matrix<S, NX, 1> xN, vx;
    Matrix<S, NX, NX> vxx;
#pragma unroll
    for(int d = 0; d < NX; d++) xN[d] = xs[((size_t)N * NX + d) * Bp + b];
    model.calcTerminalCostDeriv(prm.t0 + N * model.dt(), xN, vx, vxx); // (:178-180)
#pragma unroll
    for(int d = 0; d < NX; d++) Vx[d] = vx[d];
#pragma unroll
    for(int d = 0; d < NX * NX; d++) Vxx[d] = vxx.d[d];
  }
  S dV0 = S(0), dV1 = S(0);
  S krn_num = S(0), krn_den = S(1); // max_i |k_i| / (|u_i| + 1) as a fraction, divided once
  S k_prev[NU]; // k_list_[i + 1], the BoxQP warm start (:452-467)
#pragma unroll
  for(int a = 0; a < NU; a++) k_prev[a] = S(0);
  bool ok = true;
  S * kff_ptr = ws.kff + (size_t)(N - 1) * NU * Bp + b;
  S * kfb_ptr = ws.kfb + ((size_t)(N - 1) * NU * NX + (size_t)jj * NU) * Bp + b;
  const S * tl0 = ring + 2 * t; // this instance's column of stage 0

  LaneTileRegs<M> T;
  {
    const unsigned st = fill % kLaneDepth;
    mbarWait(&full[st], (fill / kLaneDepth) & 1u);
    loadLaneTile<M>(tl0 + (size_t)st * laneStageElems(TL::SIZE), jj, T);
    mbarArrive(&empty[st]);
    fill++;
  }

  for(int i = N - 1; i >= 0; i--)
  {
    const unsigned stn = fill % kLaneDepth; // stage of step i - 1
    const unsigned parn = (fill / kLaneDepth) & 1u;
    const bool next_ready = (i > 0) ? mbarTestWait(&full[stn], parn) : true;
    const bool act = need && ok;
    S u_cur[NU];
#pragma unroll
    for(int a = 0; a < NU; a++) u_cur[a] = T.u[a];

    // ---- replicated in the group: z = Vxx Fu, Quu = Luu + Fu^T z, Qu = Lu + Fu^T Vx      (:386-408)
    S z[NX * NU], Quu[NU * NU], Qu[NU];
#pragma unroll
    for(int a = 0; a < NU; a++)
#pragma unroll
      for(int r = 0; r < NX; r++)
      {
        S s = S(0);
#pragma unroll
        for(int k = 0; k < NX; k++) s += Vxx[r + k * NX] * T.Fu[k + a * NX];
        z[r + a * NX] = s;
      }
#pragma unroll
    for(int c = 0; c < NU; c++)
#pragma unroll
      for(int a = 0; a < NU; a++)
      {
        S s = S(0);
#pragma unroll
        for(int r = 0; r < NX; r++) s += T.Fu[r + a * NX] * z[r + c * NX];
        Quu[a + c * NU] = T.Luu[a + c * NU] + s;
      }
#pragma unroll
    for(int a = 0; a < NU; a++)
    {
      S s = S(0);
#pragma unroll
      for(int r = 0; r < NX; r++) s += T.Fu[r + a * NX] * Vx[r];
      Qu[a] = T.Lu[a] + s;
    }

    // ---- regularisation (:421-441)
    S Quu_F[NU * NU], zr[NX * NU];
    if constexpr(REG2)
    {
      // Vxx_reg = Vxx + lambda I  =>  z_reg = z + lambda Fu
#pragma unroll
      for(int d = 0; d < NX * NU; d++) zr[d] = z[d] + lambda * T.Fu[d];
#pragma unroll
      for(int c = 0; c < NU; c++)
#pragma unroll
        for(int a = 0; a < NU; a++)
        {
          S s = S(0);
#pragma unroll
          for(int r = 0; r < NX; r++) s += T.Fu[r + a * NX] * zr[r + c * NX];
          Quu_F[a + c * NU] = T.Luu[a + c * NU] + s;
        }
    }
    else
    {
#pragma unroll
      for(int d = 0; d < NX * NU; d++) zr[d] = z[d];
#pragma unroll
      for(int d = 0; d < NU * NU; d++) Quu_F[d] = Quu[d];
#pragma unroll
      for(int a = 0; a < NU; a++) Quu_F[a + a * NU] += lambda_uu;
    }

    // ---- column j: W = Vxx Fx(:,j), Qxx(:,j), Qux(:,j), Qx(j)
    S W[NX], Qxxj[NX], Quxj[NU], Quxj_reg[NU];
#pragma unroll
    for(int r = 0; r < NX; r++)
    {
      S s = S(0);
#pragma unroll
      for(int k = 0; k < NX; k++) s += Vxx[r + k * NX] * T.Fxj[k];
      W[r] = s;
    }
#pragma unroll
    for(int r = 0; r < NX; r++)
    {
      S s = S(0);
#pragma unroll
      for(int k = 0; k < NX; k++) s += T.Fx[k + r * NX] * W[k];
      Qxxj[r] = T.Lxxj[r] + s;
    }
#pragma unroll
    for(int a = 0; a < NU; a++)
    {
      S s = S(0);
#pragma unroll
      for(int r = 0; r < NX; r++) s += z[r + a * NX] * T.Fxj[r];
      Quxj[a] = T.Lxuj[a] + s;
      if constexpr(REG2)
      {
        S sr = S(0);
#pragma unroll
        for(int r = 0; r < NX; r++) sr += zr[r + a * NX] * T.Fxj[r];
        Quxj_reg[a] = T.Lxuj[a] + sr;
      }
      else
        Quxj_reg[a] = Quxj[a];
    }
    S Qxj;
    {
      S s = S(0);
#pragma unroll
      for(int r = 0; r < NX; r++) s += T.Fxj[r] * Vx[r];
      Qxj = T.Lxj + s;
    }

    // ---- gains: k replicated, K(:,j) in lane j                                        (:450-510)
    S k[NU], Kj[NU];
    if constexpr(CONSTRAINED)
    {
#pragma unroll
      for(int a = 0; a < NU; a++)
      {
        k[a] = S(0);
        Kj[a] = S(0);
      }
      if(act)
      {
        S lo[NU], hi[NU], init[NU];
#pragma unroll
        for(int a = 0; a < NU; a++)
        {
          lo[a] = ws.u_lo[(size_t)i * NU + a] - u_cur[a]; // input_limits_func_(t_i) (:470)
          hi[a] = ws.u_hi[(size_t)i * NU + a] - u_cur[a];
          init[a] = warmStartFromNextStep<M>(model, prm.t0, i, N) ? k_prev[a] : S(0);
        }
        BoxQPResult<S, NU> qp;
        boxQpSolve<S, NU>(Quu_F, Qu, lo, hi, init, qp);
        if(qp.retval < 0)
          ok = false;
        else
        {
#pragma unroll
          for(int a = 0; a < NU; a++) k[a] = qp.x[a];
          const int nf = qp.n_free;
          S rhs[NU];
          for(int r = 0; r < nf; r++) rhs[r] = Quxj_reg[qp.free_idxs[r]];
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
          for(int r = 0; r < nf; r++) Kj[qp.free_idxs[r]] = S(-1) * rhs[r];
        }
      }
    }
    else if constexpr(NU == 1)
    {
      // 1x1: the LLT failure rule is "Quu_F <= 0"; the L L^T solve is one reciprocal
      if(Quu_F[0] <= S(0)) ok = false;
      const S inv = S(1) / Quu_F[0];
      k[0] = -(Qu[0] * inv);
      Kj[0] = -(Quxj_reg[0] * inv);
    }
    else
    {
      if(!lltInPlace<S, NU>(Quu_F)) ok = false;
      S invd[NU];
#pragma unroll
      for(int a = 0; a < NU; a++) invd[a] = S(1) / Quu_F[a + a * NU];
#pragma unroll
      for(int a = 0; a < NU; a++) k[a] = Qu[a];
      lltSolveInPlace<S, NU>(Quu_F, invd, k);
#pragma unroll
      for(int a = 0; a < NU; a++) k[a] = -k[a];
#pragma unroll
      for(int a = 0; a < NU; a++) Kj[a] = Quxj_reg[a];
      lltSolveInPlace<S, NU>(Quu_F, invd, Kj);
#pragma unroll
      for(int a = 0; a < NU; a++) Kj[a] = -Kj[a];
    }
    const bool store = act && ok; // this step's gains are valid

    // ---- exchange 1: every lane gets Qux and K of all columns
    S Qux[NU * G], K[NU * G];
    {
      S mine[LL::X1], all[G * LL::X1];
#pragma unroll
      for(int a = 0; a < NU; a++)
      {
        mine[a] = Quxj[a];
        mine[NU + a] = Kj[a];
      }
      XCH::template gather<LL::X1>(x1, kFull, lane, j, mine, all);
#pragma unroll
      for(int c = 0; c < G; c++)
#pragma unroll
        for(int a = 0; a < NU; a++)
        {
          Qux[a + c * NU] = all[c * LL::X1 + a];
          K[a + c * NU] = all[c * LL::X1 + NU + a];
        }
    }

    // ---- the tile of step i - 1 replaces the one just used (its registers are dead from here on)
    if(i > 0)
    {
      if(!next_ready) mbarWait(&full[stn], parn);
      loadLaneTile<M>(tl0 + (size_t)stn * laneStageElems(TL::SIZE), jj, T);
      mbarArrive(&empty[stn]);
      fill++;
    }

    // ---- cost-to-go (:522-526)
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
    S QuuKj[NU]; // Quu K(:,j)
#pragma unroll
    for(int a = 0; a < NU; a++)
    {
      S s = S(0);
#pragma unroll
      for(int c = 0; c < NU; c++) s += Quu[a + c * NU] * Kj[c];
      QuuKj[a] = s;
    }
    S mine2[LL::X2], all2[G * LL::X2];
    // Vxx'(r, j) = Qxx(r, j) + K(:,r)^T Quu K(:,j) + K(:,r)^T Qux(:,j) + Qux(:,r)^T K(:,j)
#pragma unroll
    for(int r = 0; r < NX; r++)
    {
      S s1 = S(0), s2 = S(0), s3 = S(0);
#pragma unroll
      for(int a = 0; a < NU; a++)
      {
        s1 += K[a + r * NU] * QuuKj[a];
        s2 += K[a + r * NU] * Quxj[a];
        s3 += Qux[a + r * NU] * Kj[a];
      }
      mine2[r] = ((Qxxj[r] + s1) + s2) + s3;
    }
    {
      // Vx'(j) = Qx(j) + K(:,j)^T Quu k + K(:,j)^T Qu + Qux(:,j)^T k
      S s1 = S(0), s2 = S(0), s3 = S(0);
#pragma unroll
      for(int a = 0; a < NU; a++)
      {
        s1 += Kj[a] * Quuk[a];
        s2 += Kj[a] * Qu[a];
        s3 += Quxj[a] * k[a];
      }
      mine2[NX] = ((Qxj + s1) + s2) + s3;
    }

    // ---- exchange 2: all columns of the unsymmetrised Vxx' and Vx'; symmetrise (:526)
    XCH::template gather<LL::X2>(x2, kFull, lane, j, mine2, all2);
#pragma unroll
    for(int c = 0; c < NX; c++)
    {
      Vx[c] = all2[c * LL::X2 + NX];
#pragma unroll
      for(int r = 0; r < NX; r++)
        Vxx[r + c * NX] = (r == c) ? all2[c * LL::X2 + r] : S(0.5) * (all2[c * LL::X2 + r] + all2[r * LL::X2 + c]);
    }

#pragma unroll
    for(int a = 0; a < NU; a++) k_prev[a] = k[a];

    // ---- save gains (:529-530), accumulate max_i |k_i| / (|u_i| + 1) (:217-221)
    S kn = S(0), un = S(0);
#pragma unroll
    for(int a = 0; a < NU; a++)
    {
      if(store && j == 0) kff_ptr[(size_t)a * Bp] = k[a];
      if(store && j < NX) kfb_ptr[(size_t)a * Bp] = Kj[a];
      kn += k[a] * k[a];
      un += u_cur[a] * u_cur[a];
    }
    {
      const S a_num = (NU == 1) ? fabs(k[0]) : sqrt(kn);
      const S a_den = ((NU == 1) ? fabs(u_cur[0]) : sqrt(un)) + S(1);
      if(a_num * krn_den > krn_num * a_den)
      {
        krn_num = a_num;
        krn_den = a_den;
      }
    }
    kff_ptr -= (size_t)NU * Bp;
    kfb_ptr -= (size_t)NU * NX * Bp;
  }
  if(need)
  {
    // an instance that is only waiting for its tile mates' retry keeps the results of its own successful sweep
    dV0_out = dV0;
    dV1_out = dV1;
    k_rel_norm_out = krn_num / krn_den;
  }
  return ok;
}

/** Shared-memory carve-up of one tile's backward pass. */
template<class M>
struct LaneSmem
{
  using S = typename M::Scalar;
  S * ring;
  S * scratch;
  unsigned long long * full;
  unsigned long long * empty;
  __device__ __forceinline__ explicit LaneSmem(unsigned char * base)
  {
    using LL = LaneLayout<M>;
    ring = reinterpret_cast<S *>(base);
    scratch = ring + LL::ringElems();
    full = reinterpret_cast<unsigned long long *>(scratch + LL::scratchElems());
    empty = full + kLaneDepth;
  }
  /** One thread, before a CTA barrier. */
  __device__ __forceinline__ void initBarriers() const
  {
    for(int st = 0; st < kLaneDepth; st++)
    {
      mbarInit(&full[st], 32);
      mbarInit(&empty[st], LaneLayout<M>::CW * 32);
    }
  }
};

enum LaneRole
{
  kLaneConsumer = 0, //!< a lane of the sweep: instance t of the tile, column lane % G
  kLaneProducer = 1, //!< linearises instance t for every P-th step
  kLaneIdle = 2 //!< a warp of the CTA without a part in the backward pass (persistent tile kernel): votes only
};

/** procOnce() Steps 1-2 (:157-231) for one 32-instance tile, called by EVERY warp of the CTA with its role: sweeps with
    growing lambda until the factorisation succeeds for every instance of the tile (the consumers vote after each sweep
    -- ONE barrier instruction for all roles), then the small-gradient termination test and the hand-over to the line
    search.  `fill` is the ring's running tile count (continues across calls). */
template<class M, bool CONSTRAINED, int P, class XCH>
__device__ __forceinline__ void laneBackward(const M & model_in_constant_bank,
                                             const Workspace<typename M::Scalar> & ws,
                                             const SolverParams<typename M::Scalar> & prm,
                                             const LaneSmem<M> & sm,
                                             int role,
                                             int b,
                                             int t,
                                             int lane,
                                             int p,
                                             bool live,
                                             const typename M::Scalar * xs,
                                             const typename M::Scalar * us,
                                             int iter,
                                             unsigned & fill)
{
  using S = typename M::Scalar;
  using LL = LaneLayout<M>;
  constexpr int NX = M::NX, G = LL::G;
  const int j = lane % G; // consumer: this lane's column
  const int jj = (j < NX) ? j : (NX - 1); // idle lanes (n_x < G) shadow the last column and never store
  // two scratch arrays (strides: LaneLayout::strideFor)
  S * x1 = sm.scratch + (size_t)t * LL::X1S;
  S * x2 = sm.scratch + (size_t)kTile * LL::X1S + (size_t)t * LL::X2S;
  const bool consumer = role == kLaneConsumer;
  S lambda = (consumer && live) ? ws.lambda[b] : S(0);
  S dlambda = (consumer && live) ? ws.dlambda[b] : S(0);
  int n_bwd = (consumer && live) ? ws.n_bwd[b] : 0;
  S dV0 = S(0), dV1 = S(0), k_rel_norm = S(0);
  bool need = consumer && live;
  bool failed = false;

  while(true)
  {
    if(role == kLaneProducer)
    {
      produceSweepLanes<M, P>(model_in_constant_bank, ws, prm, b, t, p, xs, us, sm.ring, sm.full, sm.empty, fill);
      fill += (unsigned)prm.N;
    }
    else if(role == kLaneIdle)
    {
      fill += (unsigned)prm.N;
    }
    else
    {
      using LM = typename LatencyOf<M>::type;
      const LM model(model_in_constant_bank);
      if(need) n_bwd++;
      const bool ok = (prm.reg_type == 2)
                          ? laneSweep<LM, CONSTRAINED, true, XCH>(model, ws, prm, b, t, lane, j, jj, xs, sm.ring, x1, x2, sm.full,
                                                                 sm.empty, fill, need, lambda, dV0, dV1, k_rel_norm)
                          : laneSweep<LM, CONSTRAINED, false, XCH>(model, ws, prm, b, t, lane, j, jj, xs, sm.ring, x1, x2, sm.full,
                                                                  sm.empty, fill, need, lambda, dV0, dV1, k_rel_norm);
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
    // does any instance of the tile need another sweep with a larger lambda?
    if(!__syncthreads_or(need ? 1 : 0)) break;
  }
  if(!consumer || !live || j != 0) return;
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

/** procOnce() Steps 1-2 as a kernel of its own.  A CTA holds TPC independent 32-instance tiles (each with its own ring,
    scratch and mbarriers); per tile, warps 0 .. G-1 run the sweep (G lanes per instance) and warps G .. G+P-1
    linearise. */
template<class M, bool CONSTRAINED, int P, class XCH, int TPC>
__global__ void __launch_bounds__((LaneLayout<M>::CW + P) * 32 * TPC)
    backward_lanes_kernel(const __grid_constant__ M model_in_constant_bank,
                          const __grid_constant__ Workspace<typename M::Scalar> ws,
                          const __grid_constant__ SolverParams<typename M::Scalar> prm,
                          int iter)
{
  pdlPrologue();
  using S = typename M::Scalar;
  using LL = LaneLayout<M>;
  constexpr int G = LL::G, IPW = LL::IPW, CW = LL::CW;
  static_assert(M::NX <= G, "one column per lane");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int sub = (threadIdx.x >> 5) / (CW + P); // tile of the CTA
  const int warp = (threadIdx.x >> 5) % (CW + P); // warp of the tile
  const int lane = threadIdx.x & 31;
  const LaneSmem<M> sm(smem_raw + (size_t)sub * LL::bytes());
  if(warp == 0 && lane == 0)
  {
    sm.initBarriers();
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    if(blockIdx.x == 0 && sub == 0) *ws.fan_count = 0; // the previous iteration's line-search work list is consumed
  }

  const bool producer = warp >= CW;
  const int t = producer ? lane : (warp * IPW + lane / G); // instance of the tile
  // ws.Bp is a multiple of 128 (four tiles): padded lanes read valid memory, never write
  const int b = (blockIdx.x * TPC + sub) * kTile + t;
  const bool live = (b < ws.B) && (ws.status[b < ws.B ? b : 0] == 0);
  // CTA-uniform exit; the barrier also publishes the mbarrier initialisation
  if(!__syncthreads_or(live)) return;

  const int sel = live ? ws.sel[b] : 0;
  const S * us = ws.u[sel];
  const S * xs = ws.x[sel];
  unsigned fill = 0;
  laneBackward<M, CONSTRAINED, P, XCH>(model_in_constant_bank, ws, prm, sm, producer ? kLaneProducer : kLaneConsumer, b, t, lane,
                                       warp - CW, live, xs, us, iter, fill);
}
} // namespace ddp
} // namespace nmpc_b200
