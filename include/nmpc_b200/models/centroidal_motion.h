/* nmpc_b200 -- centroidal-motion problem functor: n_x = 9, input dimension 16 or 0 along the horizon (device + host).
 *
 * Same problem as the reference's DDPProblemCentroidalMotion, a DDPProblem<9, Eigen::Dynamic> (isri-aist/NMPC
 * nmpc_ddp/tests/src/TestDDPCentroidalMotion.cpp:18-201): state [CoM position, linear momentum, angular momentum],
 * input = the 16 force scales along the friction-pyramid ridges of a rectangular contact (4 vertices x 4 ridges,
 * makeStanceDataFromRect :203-236).  The stance is a function of time (ref_stance_func of the test, :246-266: a
 * first rectangle, a flight phase WITHOUT contact -- input dimension 0 --, a second rectangle) and so is the
 * reference CoM position (ref_pos_func, :267-279).
 *
 * As for models/vertical_motion.h the functor declares NU = the largest dimension and `int inputDim(t)`; inputs
 * a >= inputDim(t) are padding that the engine keeps at zero and decouples in the linearisation.  During the flight
 * phase the stance tables of the second rectangle are used: every term they enter is multiplied by an input that is
 * exactly zero.
 *
 * Flat parameter layout (38): [dt, running_x(9), running_u, terminal_x(9), mass, stance1_end_t, stance2_start_t,
 * ref_switch_t, rect1(min_x, min_y, max_x, max_y), rect2(..), ref_pos1(3), ref_pos2(3)].
 */
#pragma once

#include <cmath>

#include <nmpc_b200/matrix.h>

namespace nmpc_b200
{
namespace models
{
template<class S = double>
struct CentroidalMotion
{
  static constexpr int NX = 9;
  static constexpr int NU = 16; //!< largest inputDim(t): 4 vertices x 4 ridges
  static constexpr int NUM_PARAMS = 38;

  using Scalar = S;
  using StateDimVector = Matrix<S, NX, 1>;
  using InputDimVector = Matrix<S, NU, 1>;
  using StateStateDimMatrix = Matrix<S, NX, NX>;
  using InputInputDimMatrix = Matrix<S, NU, NU>;
  using StateInputDimMatrix = Matrix<S, NX, NU>;

  S dt_ = S(0.03);
  S running_x[NX]; // CostWeight (:39-51)
  S running_u = S(1e-6);
  S terminal_x[NX];
  S mass_ = S(100.0); // [kg]
  S stance1_end_t = S(1.4), stance2_start_t = S(1.6), ref_switch_t = S(1.5);
  S rect[2][4]; //!< {min_x, min_y, max_x, max_y} of the two contact rectangles
  S ref_pos[2][3];
  S ridge[4][3]; //!< friction-pyramid ridges (:214-220), the same for every vertex; filled by fromParams

  static constexpr double g_ = 9.80665; // [m/s^2], along z

  static CentroidalMotion fromParams(const double * p)
  {
    CentroidalMotion m;
    m.dt_ = S(p[0]);
    for(int i = 0; i < NX; i++) m.running_x[i] = S(p[1 + i]);
    m.running_u = S(p[10]);
    for(int i = 0; i < NX; i++) m.terminal_x[i] = S(p[11 + i]);
    m.mass_ = S(p[20]);
    m.stance1_end_t = S(p[21]), m.stance2_start_t = S(p[22]), m.ref_switch_t = S(p[23]);
    for(int r = 0; r < 2; r++)
      for(int i = 0; i < 4; i++) m.rect[r][i] = S(p[24 + 4 * r + i]);
    for(int r = 0; r < 2; r++)
      for(int i = 0; i < 3; i++) m.ref_pos[r][i] = S(p[32 + 3 * r + i]);
    for(int i = 0; i < 4; i++)
    {
      // ridge_list[i] << 0.5 * cos(theta), 0.5 * sin(theta), 1; ridge_list[i].normalize();   (:216-220)
      const double theta = 2 * M_PI * (static_cast<double>(i) / 4);
      const double v[3] = {0.5 * std::cos(theta), 0.5 * std::sin(theta), 1.0};
      const double n = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
      for(int k = 0; k < 3; k++) m.ridge[i][k] = S(v[k] / n);
    }
    return m;
  }
  static void defaultParams(double * p)
  {
    const double d[NUM_PARAMS] = {0.03, 1, 1, 1, 0, 0, 0, 1, 1, 1, 1e-6, 1, 1, 1, 0, 0, 0, 1, 1, 1, 100.0, 1.4, 1.6, 1.5,
                                  -0.1, -0.1, 0.1, 0.1, 0.4, -0.1, 0.6, 0.1, 0.0, 0.0, 1.0, 0.5, 0.0, 1.0};
    for(int i = 0; i < NUM_PARAMS; i++) p[i] = d[i];
  }
  NMPC_HD S dt() const
  {
    return dt_;
  }

  /** Index of the contact rectangle at time t, -1 in the flight phase (ref_stance_func, :246-266). */
  NMPC_HD int stanceAt(S t) const
  {
    // Add small values to avoid numerical instability at inequality bounds
    t += S(1e-6);
    if(t < stance1_end_t) return 0;
    if(t < stance2_start_t) return -1;
    return 1;
  }

  /** DDPProblem::inputDim(t) (DDPProblem.h:72-85; TestDDPCentroidalMotion.cpp:65-69). */
  NMPC_HD int inputDim(S t) const
  {
    return stanceAt(t) < 0 ? 0 : NU;
  }

  /** ref_pos_func of the test (:267-279). */
  NMPC_HD const S * refPos(S t) const
  {
    t += S(1e-6);
    return (t < ref_switch_t) ? ref_pos[0] : ref_pos[1];
  }

  /** Column 4 * vertex + ridge of vertices_mat (makeStanceDataFromRect, :205-212 and :225-233). */
  NMPC_HD void vertexOf(int stance, int vertex, S v[3]) const
  {
    const S * r = rect[stance < 0 ? 1 : stance];
    v[0] = (vertex < 2) ? r[0] : r[2];
    v[1] = (vertex == 0 || vertex == 3) ? r[1] : r[3];
    v[2] = S(0);
  }

  /** (vertex - com).cross(ridge) */
  NMPC_HD static void armCrossRidge(const S v[3], const StateDimVector & x, const S r[3], S out[3])
  {
    const S a0 = v[0] - x[0], a1 = v[1] - x[1], a2 = v[2] - x[2];
    out[0] = a1 * r[2] - a2 * r[1];
    out[1] = a2 * r[0] - a0 * r[2];
    out[2] = a0 * r[1] - a1 * r[0];
  }

  NMPC_HD StateDimVector stateEq(S t, const StateDimVector & x, const InputDimVector & u) const
  {
    const int stance = stanceAt(t);
    S force[3] = {S(0), S(0), S(0)}; // ridges_mat * u
    S am_dot[3] = {S(0), S(0), S(0)};
NMPC_UNROLL
    for(int vi = 0; vi < 4; vi++)
    {
      S v[3];
      vertexOf(stance, vi, v);
NMPC_UNROLL
      for(int ri = 0; ri < 4; ri++)
      {
        const S ui = u[4 * vi + ri];
        S c[3];
        armCrossRidge(v, x, ridge[ri], c);
NMPC_UNROLL
        for(int k = 0; k < 3; k++)
        {
          force[k] += ridge[ri][k] * ui;
          am_dot[k] += ui * c[k];
        }
      }
    }
    StateDimVector out;
NMPC_UNROLL
    for(int k = 0; k < 3; k++)
    {
      const S gk = (k == 2) ? S(g_) : S(0);
      out[k] = x[k] + dt_ * (x[3 + k] / mass_); // com_dot = linear_momentum / mass_
      out[3 + k] = x[3 + k] + dt_ * (force[k] - mass_ * gk); // linear_momentum_dot = ridges_mat * u - mass_ * g_
      out[6 + k] = x[6 + k] + dt_ * am_dot[k];
    }
    return out;
  }

  NMPC_HD S runningCost(S t, const StateDimVector & x, const InputDimVector & u) const
  {
    const S * rp = refPos(t);
    S sx = S(0);
NMPC_UNROLL
    for(int k = 0; k < NX; k++)
    {
      const S e = (k < 3) ? (x[k] - rp[k]) : x[k];
      sx += running_x[k] * (e * e);
    }
    return S(0.5) * sx + S(0.5) * running_u * u.squaredNorm();
  }

  NMPC_HD S terminalCost(S t, const StateDimVector & x) const
  {
    const S * rp = refPos(t);
    S sx = S(0);
NMPC_UNROLL
    for(int k = 0; k < NX; k++)
    {
      const S e = (k < 3) ? (x[k] - rp[k]) : x[k];
      sx += terminal_x[k] * (e * e);
    }
    return S(0.5) * sx;
  }

  NMPC_HD void calcStateEqDeriv(S t, const StateDimVector & x, const InputDimVector & u, StateStateDimMatrix & Fx,
                                StateInputDimMatrix & Fu) const
  {
    const int stance = stanceAt(t);
    S force[3] = {S(0), S(0), S(0)};
    Fu.setZero();
NMPC_UNROLL
    for(int vi = 0; vi < 4; vi++)
    {
      S v[3];
      vertexOf(stance, vi, v);
NMPC_UNROLL
      for(int ri = 0; ri < 4; ri++)
      {
        const int i = 4 * vi + ri;
        S c[3];
        armCrossRidge(v, x, ridge[ri], c);
NMPC_UNROLL
        for(int k = 0; k < 3; k++)
        {
          force[k] += ridge[ri][k] * u[i];
          Fu(3 + k, i) = ridge[ri][k];
          Fu(6 + k, i) = c[k];
        }
      }
    }
    Fu *= dt_;

    Fx.setZero();
NMPC_UNROLL
    for(int k = 0; k < 3; k++) Fx(k, 3 + k) = S(1) / mass_;
    // block<3, 3>(6, 0) = crossMat(ridges_mat * u)
    Fx(6, 1) = -force[2], Fx(6, 2) = force[1];
    Fx(7, 0) = force[2], Fx(7, 2) = -force[0];
    Fx(8, 0) = -force[1], Fx(8, 1) = force[0];
    Fx *= dt_;
    Fx.addToDiagonal(S(1));
  }

  NMPC_HD void calcRunningCostDeriv(S t, const StateDimVector & x, const InputDimVector & u, StateDimVector & Lx,
                                    InputDimVector & Lu, StateStateDimMatrix & Lxx, InputInputDimMatrix & Luu,
                                    StateInputDimMatrix & Lxu) const
  {
    const S * rp = refPos(t);
    Lxx.setZero();
NMPC_UNROLL
    for(int k = 0; k < NX; k++)
    {
      Lx[k] = running_x[k] * ((k < 3) ? (x[k] - rp[k]) : x[k]);
      Lxx(k, k) = running_x[k];
    }
NMPC_UNROLL
    for(int i = 0; i < NU; i++) Lu[i] = running_u * u[i];
    Luu.setIdentity();
    Luu *= running_u;
    Lxu.setZero();
  }

  NMPC_HD void calcTerminalCostDeriv(S t, const StateDimVector & x, StateDimVector & Vx, StateStateDimMatrix & Vxx) const
  {
    const S * rp = refPos(t);
    Vxx.setZero();
NMPC_UNROLL
    for(int k = 0; k < NX; k++)
    {
      Vx[k] = terminal_x[k] * ((k < 3) ? (x[k] - rp[k]) : x[k]);
      Vxx(k, k) = terminal_x[k];
    }
  }
};
} // namespace models
} // namespace nmpc_b200
