// Per-state factor evaluation shared by every kernel of the GN path (sm_100a).
//
// One call of assemble_node() evaluates, for trajectory state t of one problem,
// every factor that touches it -- start/goal prior, the two adjacent GP-prior
// factors, the SDF obstacle factor (bilinear lookup + hinge + Jacobian) and the
// optional nonholonomic / velocity-limit factor -- and returns that state's row
// of the block-tridiagonal normal equations (D_t, U_t, r_t) together with its
// contribution to the weighted and "external" errors.  The reference's dense
// A, b, K (plan_layer.py:152-200) are never formed.
//
// All arithmetic is IEEE double.  Where the reference takes a data-dependent
// branch (pixel floor, hinge test) the operation order of the reference is
// reproduced so the branch sees bit-identical numbers.
#pragma once
#include "hd.cuh"
#include "bcr_plan.cuh"

namespace dgpmp2 {

// Kernel-side constants (by-value kernel argument).
struct KParams {
  int B, T, H, W, flags, M;
  long long sdf_sb;
  double orig_x, orig_y, res;      // sdf_utils.py:57-58, obstacle_cost.py:34
  double inv_res;                  // 1 / res (gradient scaling only; pixel coordinates use the exact division)
  double dt, qa, qb, qc;           // 12 dt^-3, -6 dt^-2, 4 dt^-1   (gp_factor.py:66-68)
  double r_sphere, ks, kg, reg, kd, kv, vx_lim, vy_lim;
  double qc_const[9], qc_fix[9];
  double w_const, w_fix, eps_const;
  // Host-precomputed GP blocks for the static case (no per-(b,t) Qc^-1, not Q_FULL): row-major d x d.
  int static_gp;        // 1: use Qs / PQs / PQPs below instead of building them per state
  int ext_same;         // 1: err_ext == err (static weights equal to the constructor-time ones)
  int prefetch;         // L2 prefetch of the SDF rows: bit 0 ahead of the assembly arithmetic (kernels.cuh: assemble_cta), bit 1 at the top of gn_step_kernel instead
  int fuse1;            // 1: gn_step_kernel eliminates the level-1 nodes inside the assembly (kernels.cuh: assemble_cta)
  double Qs[36];        // Q^-1 from qc_const
  double PQs[36];       // Phi^T Q^-1
  double PQPs[36];      // Phi^T Q^-1 Phi
  double Qf[36];        // Q^-1 from qc_fix (err_ext)
  // BCR schedule (bcr.cuh: BcrPlan), computed on the host
  BcrPlan plan;
  // mixed-precision step (mp.cuh): refinement is accepted once the estimated remaining error is below mp_accept * |x|
  float mp_accept;
};

template <typename IO>
struct KWeights {
  const IO* qc; long long qc_sb, qc_st;
  const IO* w;  long long w_sb, w_st;
  const IO* eps; long long e_sb, e_st;
};

enum : int { FLAG_NONHOLONOMIC = 1, FLAG_VEL_LIMITS = 2, FLAG_Q_FULL = 4, FLAG_HEAD = 8, FLAG_HEAD_QC_VEC = 16 };

DG_HD constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }  // j <= i

template <typename IO> DG_HD double ldg_d(const IO* p) { return (double)dg_ldg(p); }

// ---------------------------------------------------------------------------
// Fused learned-covariance head (diff_gpmp2_planner.py:247-283).  With FLAG_HEAD the weight pointers hold
// the raw outputs of the learned module; every covariance entry is a product of two of them, rounded once in
// the I/O element type (what torch.mul gives on the reference's tensors) and then widened.
// ---------------------------------------------------------------------------
DG_HD double io_prod(float a, float b) { return (double)dg_fmul(a, b); }
DG_HD double io_prod(double a, double b) { return dg_dmul(a, b); }

// per-state scalar weight (w_obs, eps): the constant, the given value, or the square of the raw head output
template <typename IO>
DG_HD double load_state_weight(const KParams& P, const IO* p, long long sb, long long st,
                                                    int b, int t, double dflt) {
  if (p == nullptr) return dflt;
  const IO v = dg_ldg(p + (long long)b * sb + (long long)t * st);
  return (P.flags & FLAG_HEAD) ? io_prod(v, v) : (double)v;
}

// ---------------------------------------------------------------------------
// SDF bilinear lookup (utils/sdf_utils.py:57-94)
// ---------------------------------------------------------------------------
struct SdfSample { double dist, Jx, Jy; };   // J as returned by bilinear_interpolate

// EXACT_J: divide the gradient by res exactly as the reference does (bilinear_interpolate API, bit-exact J);
// otherwise multiply by 1/res (<= 1 ulp difference, no branch depends on J) -- used inside the GN kernels.
template <typename IO, bool EXACT_J = true>
DG_HD SdfSample sdf_bilinear(const IO* __restrict__ sdf, int H, int W,
                                                  double orig_x, double orig_y, double res,
                                                  double x, double y, double inv_res = 0.0) {
  // px = orig_x + x / res ; py = orig_y - y / res   (true divisions, reference order)
  const double px = dg_dadd(orig_x, dg_ddiv(x, res));
  const double py = dg_dsub(orig_y, dg_ddiv(y, res));
  // floor -> +1 -> clamp to the image (sdf_utils.py:64-72).  The clamps run on integers: cvt.rmi saturates, so
  // |px| >= 2^31 lands on the same border pixel as the reference's float clamp (a NaN coordinate gives NaN either way)
  const int ix = dg_d2i_rd(px), iy = dg_d2i_rd(py);
  const int x1 = dg_min(dg_max(ix, 0), W - 1), x2 = dg_min(dg_max(ix, -1), W - 2) + 1;   // clamp(ix + 1, 0, W - 1) without overflow
  const int y1 = dg_min(dg_max(iy, 0), H - 1), y2 = dg_min(dg_max(iy, -1), H - 2) + 1;
  const double x1d = (double)x1, x2d = (double)x2, y1d = (double)y1, y2d = (double)y2;
  const IO* r1 = sdf + (long long)y1 * W;
  const IO* r2 = sdf + (long long)y2 * W;
  const double v11 = ldg_d(r1 + x1), v21 = ldg_d(r1 + x2);
  const double v12 = ldg_d(r2 + x1), v22 = ldg_d(r2 + x2);
  // weights from the CLAMPED indices (:81-89)
  const double ax = dg_dsub(x2d, px), bx = dg_dsub(px, x1d);
  const double ay = dg_dsub(y2d, py), by = dg_dsub(py, y1d);
  // dist = wa*v11 + wb*v21 + wc*v12 + wd*v22, separately rounded like the tensor ops (:90)
  const double wa = dg_dmul(ax, ay), wb = dg_dmul(bx, ay), wc = dg_dmul(ax, by), wd = dg_dmul(bx, by);
  SdfSample s;
  s.dist = dg_dadd(dg_dadd(dg_dadd(dg_dmul(wa, v11), dg_dmul(wb, v21)), dg_dmul(wc, v12)),
                     dg_dmul(wd, v22));
  // J[:, :, 0] = -1*(wja*(v21-v11) + wjb*(v22-v12))/res ; J[:, :, 1] = (wjc*(v12-v11) + wjd*(v22-v21))/res  (:93-94)
  const double gx = dg_dadd(dg_dmul(ay, dg_dsub(v21, v11)), dg_dmul(by, dg_dsub(v22, v12)));
  const double gy = dg_dadd(dg_dmul(ax, dg_dsub(v12, v11)), dg_dmul(bx, dg_dsub(v22, v21)));
  if (EXACT_J) {
    s.Jx = dg_ddiv(-gx, res);
    s.Jy = dg_ddiv(gy, res);
  } else {
    s.Jx = -gx * inv_res;
    s.Jy = gy * inv_res;
  }
  return s;
}

// Hinge (obstacle_cost.py:30-37): c = (dist <= eps_tot) ? eps_tot - dist : 0 ; H_e = (dist <= eps_tot) ? -J : 0
struct ObsTerm { double c, hx, hy; };
DG_HD ObsTerm hinge(const SdfSample& s, double eps_tot) {
  ObsTerm o;
  const bool active = s.dist <= eps_tot;
  o.c = active ? (eps_tot - s.dist) : 0.0;
  o.hx = active ? -s.Jx : 0.0;
  o.hy = active ? -s.Jy : 0.0;
  return o;
}

// ---------------------------------------------------------------------------
// GP prior inverse covariance for factor i (gp_factor.py:65-73) or given in full (plan_layer.py:90)
// ---------------------------------------------------------------------------
template <int DOF, typename IO>
DG_HD void load_qinv(const KParams& P, const KWeights<IO>& Wt, int b, int i, double (&Q)[2 * DOF][2 * DOF]) {
  constexpr int D = 2 * DOF;
  if (P.flags & FLAG_Q_FULL) {
    const IO* q = Wt.qc + (long long)b * Wt.qc_sb + (long long)i * Wt.qc_st;
    if (P.flags & FLAG_HEAD) {          // 'q_full' head: Q^-1 = v v^T, v = d raw values (:274-278)
      IO v[D];
#pragma unroll
      for (int a = 0; a < D; ++a) v[a] = dg_ldg(q + a);
#pragma unroll
      for (int a = 0; a < D; ++a)
#pragma unroll
        for (int c = 0; c < D; ++c) Q[a][c] = io_prod(v[a], v[c]);
      return;
    }
#pragma unroll
    for (int a = 0; a < D; ++a)
#pragma unroll
      for (int c = 0; c < D; ++c) Q[a][c] = ldg_d(q + a * D + c);
    return;
  }
  double C[DOF][DOF];
  if (Wt.qc != nullptr && (P.flags & FLAG_HEAD)) {
    const IO* q = Wt.qc + (long long)b * Wt.qc_sb + (long long)i * Wt.qc_st;
    if (P.flags & FLAG_HEAD_QC_VEC) {   // 'qc_full' head: Qc^-1 = v v^T, v = dof raw values (:269-273)
      IO v[DOF];
#pragma unroll
      for (int a = 0; a < DOF; ++a) v[a] = dg_ldg(q + a);
#pragma unroll
      for (int a = 0; a < DOF; ++a)
#pragma unroll
        for (int c = 0; c < DOF; ++c) C[a][c] = io_prod(v[a], v[c]);
    } else {                            // 'diag_identity' head: Qc^-1 = q^2 I (:256-262)
      const IO v = dg_ldg(q);
      const double qq = io_prod(v, v);
#pragma unroll
      for (int a = 0; a < DOF; ++a)
#pragma unroll
        for (int c = 0; c < DOF; ++c) C[a][c] = (a == c) ? qq : 0.0;
    }
  } else if (Wt.qc != nullptr) {
    const IO* q = Wt.qc + (long long)b * Wt.qc_sb + (long long)i * Wt.qc_st;
#pragma unroll
    for (int a = 0; a < DOF; ++a)
#pragma unroll
      for (int c = 0; c < DOF; ++c) C[a][c] = ldg_d(q + a * DOF + c);
  } else {
#pragma unroll
    for (int a = 0; a < DOF; ++a)
#pragma unroll
      for (int c = 0; c < DOF; ++c) C[a][c] = P.qc_const[a * DOF + c];
  }
#pragma unroll
  for (int a = 0; a < DOF; ++a)
#pragma unroll
    for (int c = 0; c < DOF; ++c) {
      Q[a][c] = P.qa * C[a][c];
      Q[a][c + DOF] = P.qb * C[a][c];
      Q[a + DOF][c] = P.qb * C[a][c];
      Q[a + DOF][c + DOF] = P.qc * C[a][c];
    }
}

template <int DOF>
DG_HD void fixed_qinv(const KParams& P, double (&Q)[2 * DOF][2 * DOF]) {
#pragma unroll
  for (int a = 0; a < DOF; ++a)
#pragma unroll
    for (int c = 0; c < DOF; ++c) {
      const double v = P.qc_fix[a * DOF + c];
      Q[a][c] = P.qa * v;
      Q[a][c + DOF] = P.qb * v;
      Q[a + DOF][c] = P.qb * v;
      Q[a + DOF][c + DOF] = P.qc * v;
    }
}

// g = th_next - Phi th   (gp_factor.py:105), Phi = [[I, dt I],[0, I]]
template <int DOF>
DG_HD void gp_residual(const double (&th)[2 * DOF], const double (&thn)[2 * DOF], double dt,
                                            double (&g)[2 * DOF]) {
#pragma unroll
  for (int a = 0; a < DOF; ++a) {
    g[a] = thn[a] - (th[a] + dt * th[a + DOF]);
    g[a + DOF] = thn[a + DOF] - th[a + DOF];
  }
}

template <int D>
DG_HD double quad_form(const double (&Q)[D][D], const double (&g)[D]) {
  double s = 0.0;
#pragma unroll
  for (int a = 0; a < D; ++a) {
    double row = 0.0;
#pragma unroll
    for (int c = 0; c < D; ++c) row += Q[a][c] * g[c];
    s += g[a] * row;
  }
  return s;
}

// ---------------------------------------------------------------------------
// One state's share of the normal equations.
// ---------------------------------------------------------------------------
template <int DOF>
struct NodeOut {
  static constexpr int D = 2 * DOF;
  double Dm[D][D];   // full symmetric diagonal block Lambda_tt (includes reg)
  double Um[D][D];   // Lambda_{t,t+1} (zero for t == T-1)
  double r[D];
  double err, err_ext;        // this state's 0.5 e^T K e contributions (NOT yet divided by M)
  double e_sg, e_gp, e_obs;   // unweighted: 0.5|e_prior|^2, 0.5|g_t|^2 (t<T-1), 0.5 c_t^2
  double obs_c, obs_hx, obs_hy;
};

// th_prev / th_next are ignored at the trajectory ends.
template <int DOF, typename IO>
DG_HD void assemble_node(const KParams& P, const KWeights<IO>& Wt, int b, int t,
                                              const double (&thp)[2 * DOF], const double (&th)[2 * DOF],
                                              const double (&thn)[2 * DOF],
                                              const IO* __restrict__ start_b, const IO* __restrict__ goal_b,
                                              const IO* __restrict__ sdf_b, NodeOut<DOF>& o) {
  constexpr int D = 2 * DOF;
  const int T = P.T;
#pragma unroll
  for (int a = 0; a < D; ++a) {
    o.r[a] = 0.0;
#pragma unroll
    for (int c = 0; c < D; ++c) {
      o.Dm[a][c] = (a == c) ? P.reg : 0.0;
      o.Um[a][c] = 0.0;
    }
  }
  o.err = 0.0; o.err_ext = 0.0; o.e_sg = 0.0; o.e_gp = 0.0;

  // ---- start / goal priors (prior_factor.py:15-18; K = I/K^2, plan_layer.py:64-68) ----
  if (t == 0) {
    double s2 = 0.0;
#pragma unroll
    for (int a = 0; a < D; ++a) {
      const double e = ldg_d(start_b + a) - th[a];
      o.Dm[a][a] += P.ks;
      o.r[a] += P.ks * e;
      s2 += e * e;
    }
    o.err += 0.5 * P.ks * s2; o.err_ext += 0.5 * P.ks * s2; o.e_sg += 0.5 * s2;
  }
  if (t == T - 1) {
    double s2 = 0.0;
#pragma unroll
    for (int a = 0; a < D; ++a) {
      const double e = ldg_d(goal_b + a) - th[a];
      o.Dm[a][a] += P.kg;
      o.r[a] += P.kg * e;
      s2 += e * e;
    }
    o.err += 0.5 * P.kg * s2; o.err_ext += 0.5 * P.kg * s2; o.e_sg += 0.5 * s2;
  }

  // ---- GP factors t (between t and t+1; H1 = Phi on this state) and t-1 (H2 = -I on this state) ----
  if (P.static_gp) {
    // constant blocks precomputed on the host; only r and the errors depend on the trajectory
    if (t < T - 1) {
      double g[D], Qg[D];
      gp_residual<DOF>(th, thn, P.dt, g);
      double gQg = 0.0, g2 = 0.0;
#pragma unroll
      for (int a = 0; a < D; ++a) {
        double q = 0.0;
#pragma unroll
        for (int c = 0; c < D; ++c) {
          q += P.Qs[a * D + c] * g[c];
          o.Um[a][c] = -P.PQs[a * D + c];
          o.Dm[a][c] += P.PQPs[a * D + c];
        }
        Qg[a] = q;
        gQg += g[a] * q;
        g2 += g[a] * g[a];
      }
      // r += Phi^T (Q g): rows a < DOF: Qg[a]; rows a >= DOF: dt*Qg[a-DOF] + Qg[a]
#pragma unroll
      for (int a = 0; a < DOF; ++a) {
        o.r[a] += Qg[a];
        o.r[a + DOF] += P.dt * Qg[a] + Qg[a + DOF];
      }
      o.err += 0.5 * gQg;
      if (P.ext_same) {
        o.err_ext += 0.5 * gQg;
      } else {
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < D; ++a) {
          double q = 0.0;
#pragma unroll
          for (int c = 0; c < D; ++c) q += P.Qf[a * D + c] * g[c];
          s += g[a] * q;
        }
        o.err_ext += 0.5 * s;
      }
      o.e_gp = 0.5 * g2;
    }
    if (t > 0) {
      double g[D];
      gp_residual<DOF>(thp, th, P.dt, g);
#pragma unroll
      for (int a = 0; a < D; ++a) {
        double q = 0.0;
#pragma unroll
        for (int c = 0; c < D; ++c) {
          q += P.Qs[a * D + c] * g[c];
          o.Dm[a][c] += P.Qs[a * D + c];
        }
        o.r[a] -= q;
      }
    }
  } else {
  if (t < T - 1) {
    double Q[D][D], g[D];
    load_qinv<DOF, IO>(P, Wt, b, t, Q);
    gp_residual<DOF>(th, thn, P.dt, g);
    // PtQ = Phi^T Q : rows a<DOF: Q[a][:], rows a>=DOF: dt*Q[a-DOF][:] + Q[a][:]
    double PtQ[D][D];
#pragma unroll
    for (int c = 0; c < D; ++c)
#pragma unroll
      for (int a = 0; a < DOF; ++a) {
        PtQ[a][c] = Q[a][c];
        PtQ[a + DOF][c] = P.dt * Q[a][c] + Q[a + DOF][c];
      }
#pragma unroll
    for (int a = 0; a < D; ++a) {
      double ra = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        ra += PtQ[a][c] * g[c];
        o.Um[a][c] = -PtQ[a][c];
      }
      o.r[a] += ra;
      // (Phi^T Q Phi)[a][c] : c<DOF: PtQ[a][c] ; c>=DOF: dt*PtQ[a][c-DOF] + PtQ[a][c]
#pragma unroll
      for (int c = 0; c < DOF; ++c) {
        o.Dm[a][c] += PtQ[a][c];
        o.Dm[a][c + DOF] += P.dt * PtQ[a][c] + PtQ[a][c + DOF];
      }
    }
    o.err += 0.5 * quad_form<D>(Q, g);
    double Qf[D][D];
    fixed_qinv<DOF>(P, Qf);
    o.err_ext += 0.5 * quad_form<D>(Qf, g);
    double g2 = 0.0;
#pragma unroll
    for (int a = 0; a < D; ++a) g2 += g[a] * g[a];
    o.e_gp = 0.5 * g2;
  }
  if (t > 0) {
    double Q[D][D], g[D];
    load_qinv<DOF, IO>(P, Wt, b, t - 1, Q);
    gp_residual<DOF>(thp, th, P.dt, g);
#pragma unroll
    for (int a = 0; a < D; ++a) {
      double ra = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        ra += Q[a][c] * g[c];
        o.Dm[a][c] += Q[a][c];
      }
      o.r[a] -= ra;
    }
  }
  }

  // ---- obstacle factor (obstacle_factor.py:35-40; one sphere centred at (x, y), J_fk = [I2 0]) ----
  {
    const double eps = load_state_weight<IO>(P, Wt.eps, Wt.e_sb, Wt.e_st, b, t, P.eps_const);
    const double w = load_state_weight<IO>(P, Wt.w, Wt.w_sb, Wt.w_st, b, t, P.w_const);
    const double eps_tot = dg_dadd(eps, P.r_sphere);
    const SdfSample s = sdf_bilinear<IO, false>(sdf_b, P.H, P.W, P.orig_x, P.orig_y, P.res, th[0], th[1], P.inv_res);
    const ObsTerm ob = hinge(s, eps_tot);
    o.Dm[0][0] += w * ob.hx * ob.hx;
    o.Dm[0][1] += w * ob.hx * ob.hy;
    o.Dm[1][0] += w * ob.hx * ob.hy;
    o.Dm[1][1] += w * ob.hy * ob.hy;
    o.r[0] += w * ob.hx * ob.c;
    o.r[1] += w * ob.hy * ob.c;
    o.err += 0.5 * w * ob.c * ob.c;
    o.err_ext += 0.5 * P.w_fix * ob.c * ob.c;
    o.e_obs = 0.5 * ob.c * ob.c;
    o.obs_c = ob.c; o.obs_hx = ob.hx; o.obs_hy = ob.hy;
  }

  // ---- custom unary factors ----
  if constexpr (DOF == 3) {
    if (P.flags & FLAG_NONHOLONOMIC) {
      // nonholonomic_factor.py:16-30, state (x, y, h, vx, vy, w); Jacobian row reproduced literally
      double sh, ch;
      dg_sincos(th[2], &sh, &ch);
      const double e = th[4] * ch - th[3] * sh;
      double n[D] = {0.0, 0.0, -th[4] * sh + th[3] * ch, -sh, ch, 0.0};
#pragma unroll
      for (int a = 0; a < D; ++a) {
#pragma unroll
        for (int c = 0; c < D; ++c) o.Dm[a][c] += P.kd * n[a] * n[c];
        o.r[a] += P.kd * n[a] * e;
      }
      o.err += 0.5 * P.kd * e * e;
      o.err_ext += 0.5 * P.kd * e * e;
    }
  }
  if constexpr (DOF == 2) {
    if (P.flags & FLAG_VEL_LIMITS) {
      // velocity_limit_factor.py:17-29: active when |v| >= limit; H = -sign(v) on the velocity entry
      const double vx = th[2], vy = th[3];
      const bool ax = fabs(vx) >= P.vx_lim, ay = fabs(vy) >= P.vy_lim;
      const double cx = ax ? fabs(vx) - P.vx_lim : 0.0, cy = ay ? fabs(vy) - P.vy_lim : 0.0;
      const double sx = ax ? -((vx > 0.0) - (vx < 0.0)) : 0.0, sy = ay ? -((vy > 0.0) - (vy < 0.0)) : 0.0;
      o.Dm[2][2] += P.kv * sx * sx;
      o.Dm[3][3] += P.kv * sy * sy;
      o.r[2] += P.kv * sx * cx;
      o.r[3] += P.kv * sy * cy;
      o.err += 0.5 * P.kv * (cx * cx + cy * cy);
      o.err_ext += 0.5 * P.kv * (cx * cx + cy * cy);
    }
  }
}

}  // namespace dgpmp2

#ifdef __CUDACC__   // (device-only from here on; the forward part above also compiles for the host emulator of the tests)
// ===========================================================================
// Backward of one GN step (reverse-mode derivative of dtheta = Lambda^-1 R and of err_ext).
//
// With lambda = Lambda^-1 gbar (gbar = dL/d dtheta, solved with the same BCR) every factor f with
// Jacobian H_f, error e_f and weight K_f contributes, for a_f = H_f lambda (factor-space adjoint),
// rho_f = e_f - H_f dtheta (linearised residual after the step), alpha_f = K_f a_f, beta_f = K_f rho_f:
//     dL = alpha_f^T de_f + beta_f^T dH_f lambda - alpha_f^T dH_f dtheta + a_f^T dK_f rho_f
// (from dL = lambda^T dR - lambda^T dLambda dtheta with Lambda = sum H^T K H + reg I, R = sum H^T K e).
// err_ext = 0.5 sum e_f^T Kfix_f e_f / M adds  ghat Kfix_f e_f  to alpha_f in the de_f term only.
// The reference obtains the same numbers by autograd through its dense solve (plan_layer.py:214-234).
// ===========================================================================
namespace dgpmp2 {

struct SdfCell {       // everything the obstacle factor's backward needs
  int x1, x2, y1, y2;
  double ax, bx, ay, by, v11, v21, v12, v22, dist, hx, hy;   // h = grad dist = -J
};

template <typename IO>
__device__ __forceinline__ SdfCell sdf_cell(const IO* __restrict__ sdf, int H, int W, double orig_x, double orig_y,
                                            double res, double inv_res, double x, double y) {
  SdfCell c;
  const double px = __dadd_rn(orig_x, __ddiv_rn(x, res));
  const double py = __dsub_rn(orig_y, __ddiv_rn(y, res));
  const double fx = floor(px), fy = floor(py);
  const double wm = (double)(W - 1), hm = (double)(H - 1);
  const double x1d = fmin(fmax(fx, 0.0), wm), x2d = fmin(fmax(fx + 1.0, 0.0), wm);
  const double y1d = fmin(fmax(fy, 0.0), hm), y2d = fmin(fmax(fy + 1.0, 0.0), hm);
  c.x1 = (int)x1d; c.x2 = (int)x2d; c.y1 = (int)y1d; c.y2 = (int)y2d;
  c.v11 = ldg_d(sdf + (long long)c.y1 * W + c.x1); c.v21 = ldg_d(sdf + (long long)c.y1 * W + c.x2);
  c.v12 = ldg_d(sdf + (long long)c.y2 * W + c.x1); c.v22 = ldg_d(sdf + (long long)c.y2 * W + c.x2);
  c.ax = __dsub_rn(x2d, px); c.bx = __dsub_rn(px, x1d); c.ay = __dsub_rn(y2d, py); c.by = __dsub_rn(py, y1d);
  const double wa = __dmul_rn(c.ax, c.ay), wb = __dmul_rn(c.bx, c.ay), wc = __dmul_rn(c.ax, c.by), wd = __dmul_rn(c.bx, c.by);
  c.dist = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(wa, c.v11), __dmul_rn(wb, c.v21)), __dmul_rn(wc, c.v12)), __dmul_rn(wd, c.v22));
  const double gx = c.ay * (c.v21 - c.v11) + c.by * (c.v22 - c.v12);
  const double gy = c.ax * (c.v12 - c.v11) + c.bx * (c.v22 - c.v21);
  c.hx = gx * inv_res;      // = -J_x
  c.hy = -gy * inv_res;     // = -J_y
  return c;
}

template <int DOF>
struct NodeGrad {
  static constexpr int D = 2 * DOF;
  double g_th[D];        // dL/d th_t (all contributions of the factors touching state t)
  double g_prior[D];     // dL/d start (t == 0) or dL/d goal (t == T-1)
  double g_q[D][D];      // dL/d Q_t^-1 (GP factor t, t < T-1)
  double g_w, g_eps;     // dL/d w_t, dL/d eps_t
};

// Gradient contributions evaluated by the thread of state t.  lam*/dth* = lambda and dtheta of states
// t-1, t, t+1 (ignored where they do not exist).  ghat = dL/d err_ext / M.  If g_sdf != nullptr the
// SDF-tap gradients are accumulated there with atomics.
template <int DOF, typename IO>
__device__ __forceinline__ void backward_node(const KParams& P, const KWeights<IO>& Wt, int b, int t,
                                              const double (&thp)[2 * DOF], const double (&th)[2 * DOF], const double (&thn)[2 * DOF],
                                              const double (&lamp)[2 * DOF], const double (&lam)[2 * DOF], const double (&lamn)[2 * DOF],
                                              const double (&dtp)[2 * DOF], const double (&dt)[2 * DOF], const double (&dtn)[2 * DOF],
                                              const IO* __restrict__ start_b, const IO* __restrict__ goal_b,
                                              const IO* __restrict__ sdf_b, double ghat, IO* __restrict__ g_sdf_b,
                                              NodeGrad<DOF>& o) {
  constexpr int D = 2 * DOF;
  const int T = P.T;
#pragma unroll
  for (int a = 0; a < D; ++a) {
    o.g_th[a] = 0.0; o.g_prior[a] = 0.0;
#pragma unroll
    for (int c = 0; c < D; ++c) o.g_q[a][c] = 0.0;
  }
  o.g_w = 0.0; o.g_eps = 0.0;

  // ---- priors: e = mean - th, H = I, K = k I ----
  if (t == 0 || t == T - 1) {
    const double k = (t == 0) ? P.ks : P.kg;
    const IO* mean = (t == 0) ? start_b : goal_b;
#pragma unroll
    for (int a = 0; a < D; ++a) {
      const double e = ldg_d(mean + a) - th[a];
      const double alpha = k * lam[a] + ghat * k * e;
      o.g_prior[a] = alpha;
      o.g_th[a] -= alpha;
    }
  }
  // ---- GP factor t (this state is the "Phi" side) ----
  if (t < T - 1) {
    double Q[D][D], g[D], a[D], rho[D], u[D];
    load_qinv<DOF, IO>(P, Wt, b, t, Q);
    gp_residual<DOF>(th, thn, P.dt, g);
#pragma unroll
    for (int i = 0; i < DOF; ++i) {
      a[i] = lam[i] + P.dt * lam[i + DOF] - lamn[i];                 // Phi lam_t - lam_{t+1}
      a[i + DOF] = lam[i + DOF] - lamn[i + DOF];
      u[i] = dt[i] + P.dt * dt[i + DOF] - dtn[i];                     // H dtheta = Phi dth_t - dth_{t+1}
      u[i + DOF] = dt[i + DOF] - dtn[i + DOF];
      rho[i] = g[i] - u[i];
      rho[i + DOF] = g[i + DOF] - u[i + DOF];
    }
    double Qf[D][D];
    fixed_qinv<DOF>(P, Qf);
    double alpha[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
      double s = 0.0, sf = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) {
        s += Q[i][c] * a[c];
        sf += Qf[i][c] * g[c];
        // dL/dK = a e^T - sym(a u^T): the reference's Cholesky backward symmetrises dL/dLambda, which only
        // matters for a non-symmetric perturbation of K (the antisymmetric half of a u^T)
        o.g_q[i][c] = a[i] * rho[c] + 0.5 * (a[i] * u[c] - u[i] * a[c]);
      }
      alpha[i] = s + ghat * sf;
    }
    // de = d th_{t+1} - Phi d th_t : this thread owns the -Phi^T alpha part
#pragma unroll
    for (int i = 0; i < DOF; ++i) {
      o.g_th[i] -= alpha[i];
      o.g_th[i + DOF] -= P.dt * alpha[i] + alpha[i + DOF];
    }
  }
  // ---- GP factor t-1 (this state is the "-I" side): + alpha_{t-1} ----
  if (t > 0) {
    double Q[D][D], g[D], a[D];
    load_qinv<DOF, IO>(P, Wt, b, t - 1, Q);
    gp_residual<DOF>(thp, th, P.dt, g);
#pragma unroll
    for (int i = 0; i < DOF; ++i) {
      a[i] = lamp[i] + P.dt * lamp[i + DOF] - lam[i];
      a[i + DOF] = lamp[i + DOF] - lam[i + DOF];
    }
    double Qf[D][D];
    fixed_qinv<DOF>(P, Qf);
#pragma unroll
    for (int i = 0; i < D; ++i) {
      double s = 0.0, sf = 0.0;
#pragma unroll
      for (int c = 0; c < D; ++c) { s += Q[i][c] * a[c]; sf += Qf[i][c] * g[c]; }
      o.g_th[i] += s + ghat * sf;
    }
  }
  // ---- obstacle factor ----
  {
    const double eps = load_state_weight<IO>(P, Wt.eps, Wt.e_sb, Wt.e_st, b, t, P.eps_const);
    const double w = load_state_weight<IO>(P, Wt.w, Wt.w_sb, Wt.w_st, b, t, P.w_const);
    const double eps_tot = __dadd_rn(eps, P.r_sphere);
    const SdfCell c = sdf_cell<IO>(sdf_b, P.H, P.W, P.orig_x, P.orig_y, P.res, P.inv_res, th[0], th[1]);
    if (c.dist <= eps_tot) {
      const double cost = eps_tot - c.dist;
      const double a = c.hx * lam[0] + c.hy * lam[1];
      const double rho = cost - (c.hx * dt[0] + c.hy * dt[1]);
      const double alpha = w * a, beta = w * rho;
      const double alpha_tot = alpha + ghat * P.w_fix * cost;
      const double mx = beta * lam[0] - alpha * dt[0], my = beta * lam[1] - alpha * dt[1];
      const double kap = (c.v22 - c.v12 - c.v21 + c.v11) * P.inv_res * P.inv_res;   // -d hx/dy = -d hy/dx
      o.g_w = a * rho;
      o.g_eps = alpha_tot;
      o.g_th[0] += -alpha_tot * c.hx - my * kap;
      o.g_th[1] += -alpha_tot * c.hy - mx * kap;
      if (g_sdf_b != nullptr) {
        const double ir = P.inv_res;
        const double g11 = -alpha_tot * (c.ax * c.ay) - mx * c.ay * ir + my * c.ax * ir;
        const double g21 = -alpha_tot * (c.bx * c.ay) + mx * c.ay * ir + my * c.bx * ir;
        const double g12 = -alpha_tot * (c.ax * c.by) - mx * c.by * ir - my * c.ax * ir;
        const double g22 = -alpha_tot * (c.bx * c.by) + mx * c.by * ir - my * c.bx * ir;
        atomicAdd(g_sdf_b + (long long)c.y1 * P.W + c.x1, (IO)g11);
        atomicAdd(g_sdf_b + (long long)c.y1 * P.W + c.x2, (IO)g21);
        atomicAdd(g_sdf_b + (long long)c.y2 * P.W + c.x1, (IO)g12);
        atomicAdd(g_sdf_b + (long long)c.y2 * P.W + c.x2, (IO)g22);
      }
    }
  }
  // ---- custom unary factors ----
  if constexpr (DOF == 3) {
    if (P.flags & FLAG_NONHOLONOMIC) {
      double sh, ch;
      sincos(th[2], &sh, &ch);
      const double vx = th[3], vy = th[4];
      const double e = vy * ch - vx * sh;
      const double n2 = -vy * sh + vx * ch;
      const double a = n2 * lam[2] - sh * lam[3] + ch * lam[4];
      const double rho = e - (n2 * dt[2] - sh * dt[3] + ch * dt[4]);
      const double alpha = P.kd * a, beta = P.kd * rho, alpha_tot = alpha + ghat * P.kd * e;
      const double m2 = beta * lam[2] - alpha * dt[2], m3 = beta * lam[3] - alpha * dt[3], m4 = beta * lam[4] - alpha * dt[4];
      o.g_th[2] += alpha_tot * (-vy * sh - vx * ch) + m2 * (-vy * ch - vx * sh) - m3 * ch - m4 * sh;
      o.g_th[3] += -alpha_tot * sh + m2 * ch;
      o.g_th[4] += alpha_tot * ch - m2 * sh;
    }
  }
  if constexpr (DOF == 2) {
    if (P.flags & FLAG_VEL_LIMITS) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const double v = th[2 + k], lim = (k == 0) ? P.vx_lim : P.vy_lim;
        if (fabs(v) >= lim) {
          const double sg = (double)((v > 0.0) - (v < 0.0));
          const double cst = fabs(v) - lim;
          const double a = -sg * lam[2 + k];
          const double alpha_tot = P.kv * a + ghat * P.kv * cst;
          o.g_th[2 + k] += alpha_tot * sg;
        }
      }
    }
  }
}

}  // namespace dgpmp2
#endif  // __CUDACC__
