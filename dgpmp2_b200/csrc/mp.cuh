// Mixed-precision Gauss-Newton step (fp32 I/O path): "node-owner" block cyclic reduction.
//
// Replaces, like bcr.cuh, the reference's dense normal equations + dense Cholesky + two dense inverses
// (plan_layer.py:214-234) -- but factorises in fp32 and recovers the accuracy of the fp64 path by
// iterative refinement with an fp64 residual:
//     x0 = M^-1 r,   x_{k+1} = x_k + M^-1 (r - Lambda x_k)        M = fp32 block Cholesky of Lambda
// r and Lambda x are evaluated in IEEE double from the factor definitions (factors.cuh), so the fixed point
// is the solution of the SAME system the fp64 kernels solve; the fp32 factorisation is only a preconditioner.
// A per-problem guard (correction norms, see mp_refine_decide) accepts after one refinement when the
// contraction is fast (cond * 2^-24 small: every configuration of the reference), iterates when it is slower
// and hands the problem to the fp64 block cyclic reduction (bcr.cuh) inside the same launch when fp32 breaks
// down (non-positive pivot, slow or no contraction).
//
// Work decomposition.  ONE THREAD PER TRAJECTORY STATE ("node").  Thread m of a problem owns the node in slot
// m of the level order of bcr.cuh (bcr_state_of_slot), so the nodes eliminated at one level are consecutive
// lanes.  The owner keeps its diagonal block D_t, right-hand side r_t and the couplings to its current left /
// right neighbours in REGISTERS through all levels; only what a neighbour needs travels through shared memory:
//   record of node j (written when j is eliminated at level l, stride s = 2^(l-1)):
//     E_j = L_j^-1 Lambda_{j-s,j}^T   (column-major)      F_j = L_j^-1 Lambda_{j,j+s}   (column-major)
//     W_j = -E_j^T F_j                (row-major; the new coupling Lambda'_{j-s,j+s})    g_j = L_j^-1 r_j  -> later x_j
// A kept node i updates  D_i -= F_l^T F_l + E_r^T E_r,  r_i -= F_l^T g_l + E_r^T g_r  from its eliminated
// neighbours l = i-s, r = i+s, and -- if it is eliminated at the next level -- picks up its new couplings W_l, W_r,
// then factors at once: ONE barrier per level.  Problems synchronise separately (named barrier per problem, or
// __syncwarp when a problem is one warp), so the problems of a CTA drift apart and hide each other's latency.
//
// Everything here is __host__ __device__: tests/host_emu runs this source on CPU threads (test infrastructure).
#pragma once
#include <string.h>
#include "factors.cuh"

namespace dgpmp2 {

#ifdef DGPMP2_MP_TIMING
#define MP_STAMP(i) cx.stamp(i)
#else
#define MP_STAMP(i) do { } while (0)
#endif

constexpr int kMpMaxRefine = 5;            // refinement iterations before a problem is handed to the fp64 path

template <int D>
struct MpRec {
  static constexpr int DD = D * D;
  static constexpr int oE = 0, oF = DD, oW = 2 * DD, oG = 3 * DD;
  // floats; D = 4: 52 (13 x 16 bytes, odd -> 128-bit accesses of consecutive records are bank-conflict free),
  // D = 6: 114 (57 x 8 bytes, odd -> 64-bit accesses conflict free)
  static constexpr int kStride = 3 * DD + D;
  __host__ __device__ static constexpr size_t problem_floats(int T) { return (size_t)T * kStride; }
};

template <int D>
struct MpNode {
  static constexpr int DS = D * (D + 1) / 2, DD = D * D;
  float Dl[DS];   // lower triangle of D_t (packed)            -> L_t, reciprocal diagonal, once eliminated
  float r[D];     // r_t                                        -> g_t; the residual / its forward sweep when refining
  float U[DD];    // Lambda_{t,t+s} column-major U[c*D+a]       -> F_t column-major
  float E[DD];    // Lambda_{t-s,t} row-major  E[a*D+c]         -> E_t column-major
  float x[D];     // solution (accumulated over the refinement)
};

DG_HD int dg_f2i(float x) {
#if DG_DEV
  return __float_as_int(x);
#else
  int i; memcpy(&i, &x, 4); return i;
#endif
}
DG_HD float dg_i2f(int i) {
#if DG_DEV
  return __int_as_float(i);
#else
  float x; memcpy(&x, &i, 4); return x;
#endif
}
DG_HD int dg_ffs(int x) {
#if DG_DEV
  return __ffs(x);
#else
  return __builtin_ffs(x);
#endif
}
// (host/device twins of bcr.cuh's slot maps; bcr.cuh is device-only)
DG_HD int mp_slot(int T, int t) {
  const int z = dg_ffs(t) - 1;
  const int s = (T - 1) - ((T - 1) >> z) + (t >> (z + 1));
  return (t == 0) ? (T - 1) : s;
}
DG_HD int mp_state_of_slot(int T, int m) {
  if (m == T - 1) return 0;
  int l = 1;
  while ((T - 1) - ((T - 1) >> l) <= m) ++l;
  const int off = (T - 1) - ((T - 1) >> (l - 1));
  return (2 * (m - off) + 1) << (l - 1);
}
// number of elimination levels: max over 1 <= t < T of ctz(t) + 1
DG_HD int mp_num_levels(int T) {
  int L = 0;
  while ((1 << L) < T) ++L;      // ceil(log2 T)
  return L;                       // T = 2: 1, T = 3,4: 2, T = 64: 6, T = 65: 7
}

// D contiguous floats <-> registers (16-byte aligned when D % 4 == 0, else 8-byte aligned)
template <int D>
DG_HD void mp_ld(const float* p, float* v) {
  if constexpr (D % 4 == 0) {
#pragma unroll
    for (int a = 0; a < D; a += 4) {
      const float4 t = *reinterpret_cast<const float4*>(p + a);
      v[a] = t.x; v[a + 1] = t.y; v[a + 2] = t.z; v[a + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int a = 0; a < D; a += 2) {
      const float2 t = *reinterpret_cast<const float2*>(p + a);
      v[a] = t.x; v[a + 1] = t.y;
    }
  }
}
template <int D>
DG_HD void mp_st(float* p, const float* v) {
  if constexpr (D % 4 == 0) {
#pragma unroll
    for (int a = 0; a < D; a += 4) *reinterpret_cast<float4*>(p + a) = make_float4(v[a], v[a + 1], v[a + 2], v[a + 3]);
  } else {
#pragma unroll
    for (int a = 0; a < D; a += 2) *reinterpret_cast<float2*>(p + a) = make_float2(v[a], v[a + 1]);
  }
}

template <int D>
DG_HD float mp_dot(const float* a, const float* b) {
  float s = a[0] * b[0];
#pragma unroll
  for (int k = 1; k < D; ++k) s = fmaf(a[k], b[k], s);
  return s;
}

// In-register Cholesky of a packed lower triangle (fp32); the diagonal is replaced by 1 / l_kk.
template <int D>
DG_HD bool mp_chol(float* L) {
  bool ok = true;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    const float akk = L[tri(k, k)];
    ok = ok && (akk > 0.0f);
    float rk = dg_rsqrtf(akk);
#ifdef DGPMP2_MP_RSQRT_NR
    rk = rk * fmaf(-0.5f * akk * rk, rk, 1.5f);
#endif
    L[tri(k, k)] = rk;
#pragma unroll
    for (int i = k + 1; i < D; ++i) L[tri(i, k)] *= rk;
#pragma unroll
    for (int j = k + 1; j < D; ++j)
#pragma unroll
      for (int i = j; i < D; ++i) L[tri(i, j)] = fmaf(-L[tri(i, k)], L[tri(j, k)], L[tri(i, j)]);
  }
  return ok;
}
template <int D>
DG_HD void mp_fwd(const float* L, float* v) {        // v <- L^-1 v
#pragma unroll
  for (int a = 0; a < D; ++a) {
    float s = v[a];
#pragma unroll
    for (int c = 0; c < a; ++c) s = fmaf(-L[tri(a, c)], v[c], s);
    v[a] = s * L[tri(a, a)];
  }
}
template <int D>
DG_HD void mp_bwd(const float* L, float* v) {        // v <- L^-T v
#pragma unroll
  for (int a = D - 1; a >= 0; --a) {
    float s = v[a];
#pragma unroll
    for (int c = a + 1; c < D; ++c) s = fmaf(-L[tri(c, a)], v[c], s);
    v[a] = s * L[tri(a, a)];
  }
}

// Eliminate the owner's node: L L^T = D, E = L^-1 Lambda_{j-s,j}^T, F = L^-1 Lambda_{j,j+s}, g = L^-1 r,
// W = -E^T F; publish E, F, W, g in the node's record.  has_right: node j + s exists.
template <int D>
DG_HD bool mp_eliminate(MpNode<D>& n, float* rec, bool has_right) {
  using R = MpRec<D>;
  const bool ok = mp_chol<D>(n.Dl);
#pragma unroll
  for (int c = 0; c < D; ++c) mp_fwd<D>(n.Dl, n.E + c * D);   // row c of Lambda_{j-s,j} = column c of its transpose, in place
  mp_fwd<D>(n.Dl, n.r);
#pragma unroll
  for (int c = 0; c < D; ++c) mp_st<D>(rec + R::oE + c * D, n.E + c * D);
  mp_st<D>(rec + R::oG, n.r);
  if (has_right) {
#pragma unroll
    for (int c = 0; c < D; ++c) mp_fwd<D>(n.Dl, n.U + c * D);
#pragma unroll
    for (int c = 0; c < D; ++c) mp_st<D>(rec + R::oF + c * D, n.U + c * D);
#pragma unroll
    for (int a = 0; a < D; ++a) {
      float w[D];
#pragma unroll
      for (int c = 0; c < D; ++c) w[c] = -mp_dot<D>(n.E + a * D, n.U + c * D);
      mp_st<D>(rec + R::oW + a * D, w);
    }
  }
  return ok;
}

// Schur update of the owner's (kept) node t at the level of stride s from its eliminated neighbours t - s, t + s.
// next: the node is eliminated at the following level and takes over its new couplings (stride 2s).
template <int D>
DG_HD void mp_schur(MpNode<D>& n, const float* recs, int T, int t, int s, bool next) {
  using R = MpRec<D>;
  constexpr int DD = D * D;
  if (t > 0) {
    const float* rl = recs + (size_t)mp_slot(T, t - s) * R::kStride;
    float F[DD], g[D];
#pragma unroll
    for (int c = 0; c < D; ++c) mp_ld<D>(rl + R::oF + c * D, F + c * D);
    mp_ld<D>(rl + R::oG, g);
#pragma unroll
    for (int a = 0; a < D; ++a) {
      n.r[a] -= mp_dot<D>(F + a * D, g);
#pragma unroll
      for (int c = 0; c <= a; ++c) n.Dl[tri(a, c)] -= mp_dot<D>(F + a * D, F + c * D);
    }
    if (next) {
#pragma unroll
      for (int a = 0; a < D; ++a) mp_ld<D>(rl + R::oW + a * D, n.E + a * D);   // Lambda'_{t-2s,t}, row-major
    }
  }
  if (t + s < T) {
    const float* rr = recs + (size_t)mp_slot(T, t + s) * R::kStride;
    float Em[DD], g[D];
#pragma unroll
    for (int c = 0; c < D; ++c) mp_ld<D>(rr + R::oE + c * D, Em + c * D);
    mp_ld<D>(rr + R::oG, g);
#pragma unroll
    for (int a = 0; a < D; ++a) {
      n.r[a] -= mp_dot<D>(Em + a * D, g);
#pragma unroll
      for (int c = 0; c <= a; ++c) n.Dl[tri(a, c)] -= mp_dot<D>(Em + a * D, Em + c * D);
    }
    if (next) {
      if (t + 2 * s < T) {
#pragma unroll
        for (int a = 0; a < D; ++a) {
          float w[D];
          mp_ld<D>(rr + R::oW + a * D, w);                                     // row a of Lambda'_{t,t+2s}
#pragma unroll
          for (int c = 0; c < D; ++c) n.U[c * D + a] = w[c];
        }
      } else {
#pragma unroll
        for (int k = 0; k < DD; ++k) n.U[k] = 0.0f;
      }
    }
  } else if (next) {
#pragma unroll
    for (int k = 0; k < DD; ++k) n.U[k] = 0.0f;
  }
}

// Forward sweep of a refinement solve, kept node: r'_t -= F_l^T g'_l + E_r^T g'_r
template <int D>
DG_HD void mp_schur_rhs(MpNode<D>& n, const float* recs, int T, int t, int s) {
  using R = MpRec<D>;
  if (t > 0) {
    const float* rl = recs + (size_t)mp_slot(T, t - s) * R::kStride;
    float g[D];
    mp_ld<D>(rl + R::oG, g);
#pragma unroll
    for (int a = 0; a < D; ++a) {
      float fa[D];
      mp_ld<D>(rl + R::oF + a * D, fa);
      n.r[a] -= mp_dot<D>(fa, g);
    }
  }
  if (t + s < T) {
    const float* rr = recs + (size_t)mp_slot(T, t + s) * R::kStride;
    float g[D];
    mp_ld<D>(rr + R::oG, g);
#pragma unroll
    for (int a = 0; a < D; ++a) {
      float ea[D];
      mp_ld<D>(rr + R::oE + a * D, ea);
      n.r[a] -= mp_dot<D>(ea, g);
    }
  }
}

// x_j = L_j^-T (g_j - E_j x_{j-s} - F_j x_{j+s});  g_j = n.r.  KEEP: E_j / F_j are still in the owner's registers,
// otherwise they are read back from its record.
template <int D, bool KEEP>
DG_HD void mp_backsub(const MpNode<D>& n, const float* rec, const float* xl_p, const float* xr_p, float* v) {
  using R = MpRec<D>;
  float xl[D];
  mp_ld<D>(xl_p, xl);
#pragma unroll
  for (int a = 0; a < D; ++a) v[a] = n.r[a];
#pragma unroll
  for (int c = 0; c < D; ++c) {
    float ec[D];
    if constexpr (KEEP) {
#pragma unroll
      for (int a = 0; a < D; ++a) ec[a] = n.E[c * D + a];
    } else {
      mp_ld<D>(rec + R::oE + c * D, ec);
    }
#pragma unroll
    for (int a = 0; a < D; ++a) v[a] = fmaf(-ec[a], xl[c], v[a]);
  }
  if (xr_p != nullptr) {
    float xr[D];
    mp_ld<D>(xr_p, xr);
#pragma unroll
    for (int c = 0; c < D; ++c) {
      float fc[D];
      if constexpr (KEEP) {
#pragma unroll
        for (int a = 0; a < D; ++a) fc[a] = n.U[c * D + a];
      } else {
        mp_ld<D>(rec + R::oF + c * D, fc);
      }
#pragma unroll
      for (int a = 0; a < D; ++a) v[a] = fmaf(-fc[a], xr[c], v[a]);
    }
  }
  mp_bwd<D>(n.Dl, v);
}

// ---------------------------------------------------------------------------------------------
// fp64 pieces: the owner's row of the normal equations (assemble_node, factors.cuh) and of the residual.
// ---------------------------------------------------------------------------------------------
// one state from global memory; vec (uniform): the trajectory is aligned for whole-state vector loads
template <int D>
DG_HD void mp_load_state(const float* p, double* out, bool vec) {
  float v[D];
#if DG_DEV
  if (!vec) {
#pragma unroll
    for (int a = 0; a < D; ++a) v[a] = __ldg(p + a);
  } else if constexpr (D % 4 == 0) {
#pragma unroll
    for (int a = 0; a < D; a += 4) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(p + a));
      v[a] = t.x; v[a + 1] = t.y; v[a + 2] = t.z; v[a + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int a = 0; a < D; a += 2) {
      const float2 t = __ldg(reinterpret_cast<const float2*>(p + a));
      v[a] = t.x; v[a + 1] = t.y;
    }
  }
#else
  (void)vec;
#pragma unroll
  for (int a = 0; a < D; ++a) v[a] = p[a];
#endif
#pragma unroll
  for (int a = 0; a < D; ++a) out[a] = (double)v[a];
}
template <int D>
DG_HD void mp_store_state(float* p, const float* v, bool vec) {
  if (vec) { mp_st<D>(p, v); return; }
#pragma unroll
  for (int a = 0; a < D; ++a) p[a] = v[a];
}

// What the owner keeps (registers) between the assembly and the residual evaluations.
template <int DOF>
struct MpAux {
  static constexpr int D = 2 * DOF;
  double r[D];        // r_t in double
  double hx, hy;      // obstacle Jacobian row (0 when the hinge is inactive)
  float th[D];        // the state itself (custom factors re-linearise nothing: their Jacobians depend on th only)
};

// res_t = r_t - (Lambda x)_t in double, with Lambda exactly as assemble_node defines it:
//   (Lambda x)_t = (reg + [t=0] ks + [t=T-1] kg) x_t + Phi^T Q_t (Phi x_t - x_{t+1}) + Q_{t-1} x_t - Q_{t-1}^T Phi x_{t-1}
//                  + w h (h^T x_t) + custom
template <int DOF, typename IO>
DG_HD void mp_residual_node(const KParams& P, const KWeights<IO>& Wt, int b, int t, const MpAux<DOF>& ax,
                            const float* xp, const float* xc, const float* xn, double* res) {
  constexpr int D = 2 * DOF;
  const int T = P.T;
  double x[D];
#pragma unroll
  for (int a = 0; a < D; ++a) x[a] = (double)xc[a];
  double diag = P.reg;
  if (t == 0) diag += P.ks;
  if (t == T - 1) diag += P.kg;
#pragma unroll
  for (int a = 0; a < D; ++a) res[a] = ax.r[a] - diag * x[a];
  if (t < T - 1) {
    // + Phi^T Q_t gx,  gx = x_{t+1} - Phi x_t
    double gx[D], q[D];
#pragma unroll
    for (int a = 0; a < DOF; ++a) {
      gx[a] = (double)xn[a] - (x[a] + P.dt * x[a + DOF]);
      gx[a + DOF] = (double)xn[a + DOF] - x[a + DOF];
    }
    if (P.static_gp) {
#pragma unroll
      for (int a = 0; a < D; ++a) {
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < D; ++c) s += P.Qs[a * D + c] * gx[c];
        q[a] = s;
      }
    } else {
      double Q[D][D];
      load_qinv<DOF, IO>(P, Wt, b, t, Q);
#pragma unroll
      for (int a = 0; a < D; ++a) {
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < D; ++c) s += Q[a][c] * gx[c];
        q[a] = s;
      }
    }
#pragma unroll
    for (int a = 0; a < DOF; ++a) {
      res[a] += q[a];
      res[a + DOF] += P.dt * q[a] + q[a + DOF];
    }
  }
  if (t > 0) {
    // - Q_{t-1} x_t + Q_{t-1}^T Phi x_{t-1}
    double u[D];
#pragma unroll
    for (int a = 0; a < DOF; ++a) {
      u[a] = (double)xp[a] + P.dt * (double)xp[a + DOF];
      u[a + DOF] = (double)xp[a + DOF];
    }
    if (P.static_gp) {
#pragma unroll
      for (int a = 0; a < D; ++a) {
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < D; ++c) s += P.Qs[a * D + c] * x[c] - P.Qs[c * D + a] * u[c];
        res[a] -= s;
      }
    } else {
      double Q[D][D];
      load_qinv<DOF, IO>(P, Wt, b, t - 1, Q);
#pragma unroll
      for (int a = 0; a < D; ++a) {
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < D; ++c) s += Q[a][c] * x[c] - Q[c][a] * u[c];
        res[a] -= s;
      }
    }
  }
  {
    const double w = load_state_weight<IO>(P, Wt.w, Wt.w_sb, Wt.w_st, b, t, P.w_const);
    const double s = w * (ax.hx * x[0] + ax.hy * x[1]);
    res[0] -= ax.hx * s;
    res[1] -= ax.hy * s;
  }
  if constexpr (DOF == 3) {
    if (P.flags & FLAG_NONHOLONOMIC) {
      double sh, ch;
      dg_sincos((double)ax.th[2], &sh, &ch);
      const double nn[D] = {0.0, 0.0, -(double)ax.th[4] * sh + (double)ax.th[3] * ch, -sh, ch, 0.0};
      double s = 0.0;
#pragma unroll
      for (int a = 0; a < D; ++a) s += nn[a] * x[a];
      s *= P.kd;
#pragma unroll
      for (int a = 0; a < D; ++a) res[a] -= nn[a] * s;
    }
  }
  if constexpr (DOF == 2) {
    if (P.flags & FLAG_VEL_LIMITS) {
      if (fabs((double)ax.th[2]) >= P.vx_lim) res[2] -= P.kv * x[2];
      if (fabs((double)ax.th[3]) >= P.vy_lim) res[3] -= P.kv * x[3];
    }
  }
}

// Lambda_{t-1,t} = -Phi^T Q_{t-1} (row-major, fp32) for the nodes eliminated at level 1
template <int DOF, typename IO>
DG_HD void mp_left_coupling(const KParams& P, const KWeights<IO>& Wt, int b, int t, float* Ul) {
  constexpr int D = 2 * DOF;
  if (P.static_gp) {
#pragma unroll
    for (int k = 0; k < D * D; ++k) Ul[k] = (float)(-P.PQs[k]);
    return;
  }
  double Q[D][D];
  load_qinv<DOF, IO>(P, Wt, b, t - 1, Q);
#pragma unroll
  for (int c = 0; c < D; ++c)
#pragma unroll
    for (int a = 0; a < DOF; ++a) {
      Ul[a * D + c] = (float)(-Q[a][c]);
      Ul[(a + DOF) * D + c] = (float)(-(P.dt * Q[a][c] + Q[a + DOF][c]));
    }
}

// Acceptance test of the refinement (inf-norms over the whole problem): nd = |dx_k|, nd_prev = |dx_{k-1}| (|x| for
// k = 1), nx = |x|.  The error left after adding dx_k is ~ rho * nd with rho the contraction of the iteration
// (~ cond(Lambda) * 2^-24: the Schur complements of the GP chain cancel ~3 bits per level).  rho is taken from the ratio
// of two successive corrections, floored at 2^-7; for k = 1 no such ratio exists yet (|dx_1| / |x| only says how x0
// happened to be aligned) and 2^-5 is assumed -- measured: 1e-2 at T = 64, 2.4e-2 at T = 128.
//   accept   rho * nd <= accept * |x|     (accept = 2^-17 = 7.6e-6: the fp32-I/O parity bar is 1e-5)
//   give up  nd > nd_prev / 4, NaN, or k == kMpMaxRefine without acceptance  -> fp64 path
// returns 1 accept, 0 iterate again, -1 fp64
DG_HD int mp_refine_decide(float nd, float nd_prev, float nx, int it, float accept) {
  if (nd == 0.0f) return 1;
  if (!(nd <= 0.25f * nd_prev)) return -1;
  const float rho = (it == 1) ? 0.03125f : fmaxf(nd / nd_prev, 0.0078125f);
  if (rho * nd <= accept * nx) return 1;
  return (it >= kMpMaxRefine) ? -1 : 0;
}

// ---------------------------------------------------------------------------------------------
// The program of ONE thread (one node).  Ctx supplies what differs between the GPU and the host emulator:
//   psync()                       barrier over the threads of this problem
//   sum2(a, b)                    a, b <- sums over the problem's threads (all threads receive them)
//   max2(a, b, it)                a, b <- integer maxima over the problem's threads (bit patterns of non-negative floats)
//   flag64()                      mark the problem for the fp64 path
// m = slot of the thread inside its problem; active = m < T.  recs = the problem's T records.
// ---------------------------------------------------------------------------------------------
template <int DOF, bool KEEP, typename Ctx>
DG_HD void mp_thread_program(Ctx& cx, const KParams& P, const KWeights<float>& Wt, const int b, const int m,
                             const bool active, const float* __restrict__ th, const float* __restrict__ start,
                             const float* __restrict__ goal, const float* __restrict__ sdf, float* recs,
                             float* __restrict__ dth, float* __restrict__ err, float* __restrict__ err_ext,
                             int* __restrict__ diag, const int force64) {
  constexpr int D = 2 * DOF;
  using R = MpRec<D>;
  // whole-state vector loads / stores need 16-byte (d = 4) or 8-byte (d = 6) aligned trajectories (uniform)
  constexpr uintptr_t amask = (D % 4 == 0) ? 15u : 7u;
  const bool vec_in = (reinterpret_cast<uintptr_t>(th) & amask) == 0, vec_out = (reinterpret_cast<uintptr_t>(dth) & amask) == 0;
  const int T = P.T;
  const int Lmax = mp_num_levels(T);
  int t = 0, lvl = 0;                         // lvl 0: takes part in nothing (inactive lane)
  MP_STAMP(0);
  MpNode<D> n;
  MpAux<DOF> ax;
  double e0 = 0.0, e1 = 0.0;
  float* rec = recs + (size_t)m * R::kStride;
  if (active) {
    t = mp_state_of_slot(T, m);
    lvl = (t == 0) ? (Lmax + 1) : dg_ffs(t);  // ctz(t) + 1
    double thp[D], thc[D], thn[D];
    const float* tp = th + ((size_t)b * T + t) * D;
    mp_load_state<D>(tp, thc, vec_in);
    if (t > 0) mp_load_state<D>(tp - D, thp, vec_in);
    if (t < T - 1) mp_load_state<D>(tp + D, thn, vec_in);
#pragma unroll
    for (int a = 0; a < D; ++a) {
      if (t == 0) thp[a] = 0.0;
      if (t == T - 1) thn[a] = 0.0;
      ax.th[a] = (float)thc[a];
    }
    NodeOut<DOF> o;
    assemble_node<DOF, float>(P, Wt, b, t, thp, thc, thn, start + (size_t)b * D, goal + (size_t)b * D,
                              sdf + (size_t)b * P.sdf_sb, o);
#pragma unroll
    for (int a = 0; a < D; ++a) {
      ax.r[a] = o.r[a];
      n.r[a] = (float)o.r[a];
      n.x[a] = 0.0f;
#pragma unroll
      for (int c = 0; c <= a; ++c) n.Dl[tri(a, c)] = (float)o.Dm[a][c];
#pragma unroll
      for (int c = 0; c < D; ++c) {
        n.U[c * D + a] = (float)o.Um[a][c];
        n.E[a * D + c] = 0.0f;
      }
    }
    ax.hx = o.obs_hx; ax.hy = o.obs_hy;
    e0 = o.err; e1 = o.err_ext;
    if (lvl == 1) mp_left_coupling<DOF, float>(P, Wt, b, t, n.E);
  }
  MP_STAMP(1);
  cx.sum2(e0, e1);
  MP_STAMP(2);
  if (m == 0) {
    const double invM = 1.0 / (double)P.M;
    err[b] = (float)(e0 * invM);
    err_ext[b] = (float)(e1 * invM);
  }

  // ------------------------------ factorisation + first solve ------------------------------
  bool ok = true;
  for (int l = 1; l <= Lmax; ++l) {
    const int s = 1 << (l - 1);
    if (lvl == l) ok = mp_eliminate<D>(n, rec, t + s < T) && ok;
    cx.psync();
    MP_STAMP(2 + l);
    if (lvl > l) mp_schur<D>(n, recs, T, t, s, lvl == l + 1);
  }
  MP_STAMP(12);
  if (lvl == Lmax + 1) {                      // root
    ok = mp_chol<D>(n.Dl) && ok;
    mp_fwd<D>(n.Dl, n.r);
    mp_bwd<D>(n.Dl, n.r);
#pragma unroll
    for (int a = 0; a < D; ++a) n.x[a] = n.r[a];
    mp_st<D>(rec + R::oG, n.x);
  }
  if (!ok || force64) cx.flag64();
  cx.psync();
  MP_STAMP(13);
  for (int l = Lmax; l >= 1; --l) {
    if (lvl == l) {
      const int s = 1 << (l - 1);
      const float* xl = recs + (size_t)mp_slot(T, t - s) * R::kStride + R::oG;
      const float* xr = (t + s < T) ? recs + (size_t)mp_slot(T, t + s) * R::kStride + R::oG : nullptr;
      mp_backsub<D, KEEP>(n, rec, xl, xr, n.x);
      mp_st<D>(rec + R::oG, n.x);
    }
    cx.psync();
  }

  MP_STAMP(14);
  // ------------------------------ iterative refinement ------------------------------
  int verdict = 0, it = 0;
  float nd_prev = 0.0f;
  while (verdict == 0) {
    ++it;
    // residual in double from the neighbours' x (published in the records), rounded to fp32
    if (active) {
      float xp[D], xn[D];
#pragma unroll
      for (int a = 0; a < D; ++a) { xp[a] = 0.0f; xn[a] = 0.0f; }
      if (t > 0) mp_ld<D>(recs + (size_t)mp_slot(T, t - 1) * R::kStride + R::oG, xp);
      if (t < T - 1) mp_ld<D>(recs + (size_t)mp_slot(T, t + 1) * R::kStride + R::oG, xn);
      double res[D];
      mp_residual_node<DOF, float>(P, Wt, b, t, ax, xp, n.x, xn, res);
#pragma unroll
      for (int a = 0; a < D; ++a) n.r[a] = (float)res[a];
    }
    cx.psync();                               // every x has been read before the records receive g'
    MP_STAMP(11 + 4 * it);
    for (int l = 1; l <= Lmax; ++l) {
      const int s = 1 << (l - 1);
      if (lvl == l) {
        mp_fwd<D>(n.Dl, n.r);
        mp_st<D>(rec + R::oG, n.r);
      }
      cx.psync();
      if (lvl > l) mp_schur_rhs<D>(n, recs, T, t, s);
    }
    float dx[D];
#pragma unroll
    for (int a = 0; a < D; ++a) dx[a] = 0.0f;
    if (lvl == Lmax + 1) {
      mp_fwd<D>(n.Dl, n.r);
      mp_bwd<D>(n.Dl, n.r);
#pragma unroll
      for (int a = 0; a < D; ++a) dx[a] = n.r[a];
      mp_st<D>(rec + R::oG, dx);
    }
    cx.psync();
    MP_STAMP(12 + 4 * it);
    for (int l = Lmax; l >= 1; --l) {
      if (lvl == l) {
        const int s = 1 << (l - 1);
        const float* xl = recs + (size_t)mp_slot(T, t - s) * R::kStride + R::oG;
        const float* xr = (t + s < T) ? recs + (size_t)mp_slot(T, t + s) * R::kStride + R::oG : nullptr;
        mp_backsub<D, KEEP>(n, rec, xl, xr, dx);
        mp_st<D>(rec + R::oG, dx);
      }
      cx.psync();
    }
    MP_STAMP(13 + 4 * it);
    int nd_b = 0, nx_b = 0;                   // bit patterns of non-negative floats order like integers; NaN sorts last
    if (active) {
#pragma unroll
      for (int a = 0; a < D; ++a) {
        n.x[a] += dx[a];
        nd_b = dg_max(nd_b, dg_f2i(fabsf(dx[a])));
        nx_b = dg_max(nx_b, dg_f2i(fabsf(n.x[a])));
      }
      mp_st<D>(rec + R::oG, n.x);             // publish x for the next residual
    }
    cx.max2(nd_b, nx_b, it);                  // includes a psync
    MP_STAMP(14 + 4 * it);
    const float nd = dg_i2f(nd_b), nx = dg_i2f(nx_b);
    verdict = mp_refine_decide(nd, (it == 1) ? nx : nd_prev, nx, it, P.mp_accept);
#ifdef DGPMP2_MP_TRACE
    if (m == 0) cx.trace(b, it, nd, nx);     // host emulator only
    if (cx.force_iters() > 0) verdict = (it >= cx.force_iters()) ? 1 : 0;
#endif
    nd_prev = nd;
  }
  MP_STAMP(40);
  if (verdict < 0) cx.flag64();
  if (diag != nullptr && m == 0) diag[b] = (verdict > 0) ? it : -it;
  if (active && verdict > 0) mp_store_state<D>(dth + ((size_t)b * T + t) * D, n.x, vec_out);
}

}  // namespace dgpmp2
