// Host side of the kernel arguments: dgpmp2_params / dgpmp2_weights (the C ABI, include/dgpmp2_b200.h) ->
// KParams / KWeights (factors.cuh).  Shared by c_abi.cu and the host emulator of the tests (tests/host_emu).
#pragma once
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "../../include/dgpmp2_b200.h"
#include "factors.cuh"

namespace dgpmp2 {

inline int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  if (e == nullptr) return dflt;
  const int v = atoi(e);
  return v > 0 ? v : dflt;
}

inline KParams make_kparams(const dgpmp2_params* p) {
  KParams k;
  memset(&k, 0, sizeof(k));
  k.B = p->B; k.T = p->T; k.H = p->H; k.W = p->W; k.flags = p->flags;
  const int d = 2 * p->dof;
  // plan_layer.py:39-45
  k.M = d * ((p->T - 1) + 2) + p->T;
  if (p->flags & DGPMP2_FLAG_NONHOLONOMIC) k.M += p->T;
  if (p->flags & DGPMP2_FLAG_VEL_LIMITS) k.M += p->dof * p->T;
  k.sdf_sb = p->sdf_stride_b;
  k.res = p->res;
  k.inv_res = 1.0 / p->res;
  k.orig_x = 0.0 - p->x_lo / p->res;          // sdf_utils.py:57
  k.orig_y = 0.0 - p->y_lo / p->res;          // sdf_utils.py:58
  k.dt = p->dt;
  k.qa = 12.0 * std::pow(p->dt, -3.0);        // gp_factor.py:66-68
  k.qb = -6.0 * std::pow(p->dt, -2.0);
  k.qc = 4.0 * std::pow(p->dt, -1.0);
  k.r_sphere = p->r_sphere; k.ks = p->ks_inv2; k.kg = p->kg_inv2; k.reg = p->reg;
  k.kd = p->kd_inv2; k.kv = p->kv_inv2; k.vx_lim = p->vx_lim; k.vy_lim = p->vy_lim;
  for (int i = 0; i < 9; ++i) { k.qc_const[i] = p->qc_inv[i]; k.qc_fix[i] = p->qc_inv_fix[i]; }
  k.w_const = p->w_obs; k.w_fix = p->w_obs_fix; k.eps_const = p->eps;
  // constant GP blocks for the static case: Q = [[qa C, qb C],[qb C, qc C]], Phi = [[I, dt I],[0, I]]
  const int dof = p->dof;
  auto kron = [&](const double* C, double* Q) {
    for (int a = 0; a < dof; ++a)
      for (int c = 0; c < dof; ++c) {
        const double v = C[a * dof + c];
        Q[a * d + c] = k.qa * v;
        Q[a * d + c + dof] = k.qb * v;
        Q[(a + dof) * d + c] = k.qb * v;
        Q[(a + dof) * d + c + dof] = k.qc * v;
      }
  };
  kron(k.qc_const, k.Qs);
  kron(k.qc_fix, k.Qf);
  for (int c = 0; c < d; ++c)
    for (int a = 0; a < dof; ++a) {
      k.PQs[a * d + c] = k.Qs[a * d + c];
      k.PQs[(a + dof) * d + c] = k.dt * k.Qs[a * d + c] + k.Qs[(a + dof) * d + c];
    }
  for (int a = 0; a < d; ++a)
    for (int c = 0; c < dof; ++c) {
      k.PQPs[a * d + c] = k.PQs[a * d + c];
      k.PQPs[a * d + c + dof] = k.dt * k.PQs[a * d + c] + k.PQs[a * d + c + dof];
    }
  k.static_gp = 0;   // set by the caller once the weights are known
  k.ext_same = 0;
  {  // SDF L2 prefetch (bit 0: inside the assembly, bit 1: gn_step_kernel issues it at its top instead).  DGPMP2_PREFETCH:
     // 1 / unset = default (early in the step kernel, in the assembly elsewhere), 2 = off, 3 = in the assembly everywhere (A/B)
    const int pf = env_int("DGPMP2_PREFETCH", 1);
    k.prefetch = (pf == 1) ? 3 : (pf == 3) ? 1 : 0;
  }
  k.mp_accept = ldexpf(1.0f, -env_int("DGPMP2_MP_ACCEPT_LOG2", 17));
  bcr_make_plan(p->T, env_int("DGPMP2_TAIL", kTailMaxDefault), env_int("DGPMP2_WIDE", kWideMinDefault), k.plan);
  return k;
}

// static_gp: no per-(b,t) Qc^-1 and not Q_FULL.  ext_same: additionally every weight equals its
// constructor-time value, so err_ext == err.
template <typename IO>
void finish_kparams(KParams& k, const KWeights<IO>& kw) {
  k.static_gp = (kw.qc == nullptr && !(k.flags & DGPMP2_FLAG_Q_FULL)) ? 1 : 0;
  bool same = k.static_gp && kw.w == nullptr && k.w_const == k.w_fix;
  for (int i = 0; i < 9 && same; ++i) same = (k.qc_const[i] == k.qc_fix[i]);
  k.ext_same = same ? 1 : 0;
  // level-1 elimination fused into the assembly: needs the host-known coupling -Phi^T Q^-1 (DGPMP2_FUSE1=2 disables it: A/B, tests)
  k.fuse1 = (k.static_gp && k.plan.nl >= 1 && env_int("DGPMP2_FUSE1", 1) == 1) ? 1 : 0;
}

template <typename IO>
KWeights<IO> make_kweights(const dgpmp2_weights* w) {
  KWeights<IO> k;
  memset(&k, 0, sizeof(k));
  if (w != nullptr) {
    k.qc = static_cast<const IO*>(w->qc_inv); k.qc_sb = w->qc_stride_b; k.qc_st = w->qc_stride_t;
    k.w = static_cast<const IO*>(w->w_obs); k.w_sb = w->w_stride_b; k.w_st = w->w_stride_t;
    k.eps = static_cast<const IO*>(w->eps); k.e_sb = w->eps_stride_b; k.e_st = w->eps_stride_t;
  }
  return k;
}

}  // namespace dgpmp2
