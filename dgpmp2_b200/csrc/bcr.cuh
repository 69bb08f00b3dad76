// Block cyclic reduction (BCR) of the SPD block-tridiagonal Gauss-Newton system,
// entirely in shared memory (sm_100a).
//
// Replaces the reference's dense normal equations + dense Cholesky + two dense
// inverses (plan_layer.py:214-234).  The T diagonal blocks D_t (d x d), the T-1
// couplings U_t = Lambda_{t,t+1} and the right-hand side r_t of every problem of
// the CTA live in shared memory as structure-of-arrays [element][node slot], with
// a COMPILE-TIME slot count NN so that every access is `base + slot*8 + immediate`.
//
// Levels l = 1..L, stride s = 2^(l-1).  At level l the nodes j = s(2q+1) are
// eliminated: L_j L_j^T = D_j, E_j = L_j^-1 U_{j-s}^T, F_j = L_j^-1 U_j,
// g_j = L_j^-1 r_j.  The kept neighbours i = j-s, k = j+s receive
//   D_i -= E_j^T E_j   r_i -= E_j^T g_j   U_i' = -E_j^T F_j
//   D_k -= F_j^T F_j   r_k -= F_j^T g_j
// Back substitution: x_j = L_j^-T (g_j - E_j x_{j-s} - F_j x_{j+s}).
// This is block Cholesky in nested-dissection order: backward stable for SPD
// systems, ceil(log2 T) dependent block steps instead of T.
//
// Nodes are stored in LEVEL ORDER: the nodes eliminated at level 1 first, then
// level 2, ..., the root (t = 0) last, so the work items of a level touch
// contiguous slots (bank-conflict free).
//
// Work decomposition: every problem owns a fixed group of TPP = LPN * ceil(T/2)
// threads; work item e of a level (one eliminated / kept node) is processed by LPN
// cooperating lanes that split the independent columns (elimination) or rows
// (Schur update) of the item.  The lanes never exchange registers; everything goes
// through the band in shared memory, bracketed by the two CTA barriers per level.
#pragma once
#include "factors.cuh"

namespace dgpmp2 {

constexpr int kMaxLevels = 16;

// number of nodes j = s(2q+1) < T
__host__ __device__ __forceinline__ int bcr_n_elim(int T, int s) { return (T + s - 1) / (2 * s); }
// number of nodes i = 2 s q < T
__host__ __device__ __forceinline__ int bcr_n_kept(int T, int s) { return (T + 2 * s - 1) / (2 * s); }

struct BcrLevels {
  int nlev;                  // number of elimination levels (strides 1, 2, ... < T)
  int off[kMaxLevels + 2];   // off[l] = first slot of level l (1-based); off[nlev+1] = T-1 = root slot
};

__host__ __device__ inline void bcr_make_levels(int T, BcrLevels& lv) {
  lv.off[0] = 0;
  lv.off[1] = 0;
  int l = 1;
  for (int s = 1; s < T; s <<= 1, ++l) lv.off[l + 1] = lv.off[l] + bcr_n_elim(T, s);
  lv.nlev = l - 1;
}

// slot of trajectory state t inside its problem
__device__ __forceinline__ int bcr_slot(const int* off, int T, int t) {
  if (t == 0) return T - 1;
  const int l = __ffs(t);          // ctz(t) + 1 = level at which t is eliminated
  return off[l] + (t >> l);
}
// inverse: trajectory state stored in slot m
__device__ __forceinline__ int bcr_state_of_slot(const int* off, int nlev, int T, int m) {
  if (m == T - 1) return 0;
  int l = 1;
  while (l < nlev && m >= off[l + 1]) ++l;
  return (2 * (m - off[l]) + 1) << (l - 1);
}

// Shared-memory band of all problems of the CTA: NN node slots, structure of arrays.
//   Dm [DS][NN]  lower triangle of D_t; after elimination: L_t with 1/l_kk on the diagonal
//   Um [DD][NN]  U_t (row-major a*D+b);  after elimination: F_t
//   Rm [D ][NN]  r_t; after elimination g_t; after back substitution x_t
//   Em [DD][NN]  E_t
template <int D, int NN>
struct Band {
  static constexpr int DS = D * (D + 1) / 2;
  static constexpr int DD = D * D;
  static constexpr int kDoublesPerNode = DS + DD + D + DD;
  static constexpr int oU = DS * NN, oR = (DS + DD) * NN, oE = (DS + DD + D) * NN;
  double* base;
  __device__ __forceinline__ double* Dp(int n) const { return base + n; }            // + k*NN
  __device__ __forceinline__ double* Up(int n) const { return base + oU + n; }       // + (a*D+b)*NN
  __device__ __forceinline__ double* Rp(int n) const { return base + oR + n; }       // + a*NN
  __device__ __forceinline__ double* Ep(int n) const { return base + oE + n; }       // + (a*D+b)*NN
};

// In-register Cholesky of a packed lower triangle; the diagonal is replaced by 1/l_kk.
// Returns false if a pivot is not strictly positive (incl. NaN).
template <int D>
__device__ __forceinline__ bool chol_packed(double (&L)[D * (D + 1) / 2]) {
  bool ok = true;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    const double akk = L[tri(k, k)];
    ok = ok && (akk > 0.0);
    const double rk = rsqrt(akk);
    L[tri(k, k)] = rk;
#pragma unroll
    for (int i = k + 1; i < D; ++i) L[tri(i, k)] *= rk;
#pragma unroll
    for (int j = k + 1; j < D; ++j)
#pragma unroll
      for (int i = j; i < D; ++i) L[tri(i, j)] -= L[tri(i, k)] * L[tri(j, k)];
  }
  return ok;
}

template <int D>
__device__ __forceinline__ void fwd_solve(const double (&L)[D * (D + 1) / 2], double (&v)[D]) {
#pragma unroll
  for (int a = 0; a < D; ++a) {
    double s = v[a];
#pragma unroll
    for (int c = 0; c < a; ++c) s -= L[tri(a, c)] * v[c];
    v[a] = s * L[tri(a, a)];
  }
}

template <int D>
__device__ __forceinline__ void bwd_solve(const double (&L)[D * (D + 1) / 2], double (&v)[D]) {
#pragma unroll
  for (int a = D - 1; a >= 0; --a) {
    double s = v[a];
#pragma unroll
    for (int c = a + 1; c < D; ++c) s -= L[tri(c, a)] * v[c];
    v[a] = s * L[tri(a, a)];
  }
}

// Factor + solve all problems of the CTA.  Thread u (< TPP) of problem p (< np) is lane
// (u % LPN) of work item (u / LPN).  On exit Rm holds x_t of every node (slot order).
// fail[p] (shared, pre-zeroed) receives t+1 of a node whose pivot was not positive.
// Must be called by all threads of the CTA (contains barriers).
template <int D, int NN, int LPN>
__device__ __forceinline__ void bcr_solve(const Band<D, NN>& bd, const int* __restrict__ lvl_off, int nlev,
                                          int T, bool active_problem, int p, int u, int EPP, int* fail) {
  constexpr int DS = D * (D + 1) / 2;
  static_assert((LPN & (LPN - 1)) == 0, "LPN must be a power of two");
  const int e0 = u / LPN, lane = u % LPN;
  const int nb = p * T;   // first slot of this problem
  // EPP work items are processed per pass by this problem's thread group

  // ------------------------------ forward elimination ------------------------------
  for (int l = 1; l <= nlev; ++l) {
    const int s = 1 << (l - 1);
    const int ne = (T + s - 1) >> l;          // bcr_n_elim(T, s) with 2s = 2^l
    const int off_l = lvl_off[l];
    // (a) factor the eliminated nodes: lanes split the columns of [U_i^T | U_j | r_j]
    for (int e = e0;; e += EPP) {
    const unsigned m_el = __ballot_sync(0xffffffffu, active_problem && e < ne);
    if (m_el == 0u) break;                    // warp-uniform
    if (active_problem && e < ne) {
      const int j = s * (2 * e + 1);
      const int pj = nb + off_l + e;
      const int pi = nb + bcr_slot(lvl_off, T, j - s);
      const bool has_right = (j + s) < T;
      double L[DS];
      {
        const double* dp = bd.Dp(pj);
#pragma unroll
        for (int k = 0; k < DS; ++k) L[k] = dp[k * NN];
      }
      // issue the column loads before the Cholesky chain so their latency overlaps it
      constexpr int NC = (D + LPN - 1) / LPN;   // columns per lane
      double ve[NC][D], vf[NC][D], vg[D];
#pragma unroll
      for (int cc = 0; cc < NC; ++cc) {
        const int c = lane + cc * LPN;
        if (c < D) {
          const double* up = bd.Up(pi) + c * D * NN;   // row c of U_i  = column c of U_i^T
          const double* fp = bd.Up(pj) + c * NN;       // column c of U_j
#pragma unroll
          for (int a = 0; a < D; ++a) {
            ve[cc][a] = up[a * NN];
            vf[cc][a] = has_right ? fp[a * D * NN] : 0.0;
          }
        }
      }
      {
        const double* rp = bd.Rp(pj);
#pragma unroll
        for (int a = 0; a < D; ++a) vg[a] = rp[a * NN];
      }
      __syncwarp(m_el);   // every lane of the item has read D_j, r_j before lane 0 overwrites them
      if (!chol_packed<D>(L)) atomicMax(&fail[p], j + 1);
      if (lane == 0) {
        double* dp = bd.Dp(pj);
#pragma unroll
        for (int k = 0; k < DS; ++k) dp[k * NN] = L[k];
      }
#pragma unroll
      for (int cc = 0; cc < NC; ++cc) {
        const int c = lane + cc * LPN;
        if (c < D) {
          fwd_solve<D>(L, ve[cc]);
          fwd_solve<D>(L, vf[cc]);
          double* ep = bd.Ep(pj) + c * NN;             // column c of E_j
          double* fp = bd.Up(pj) + c * NN;             // column c of F_j (in place)
#pragma unroll
          for (int a = 0; a < D; ++a) {
            ep[a * D * NN] = ve[cc][a];
            fp[a * D * NN] = vf[cc][a];
          }
        }
      }
      fwd_solve<D>(L, vg);
      if (lane == 0) {
        double* rp = bd.Rp(pj);
#pragma unroll
        for (int a = 0; a < D; ++a) rp[a * NN] = vg[a];
      }
    }
    }
    __syncthreads();
    // (b) Schur-complement update of the kept nodes: lanes split the rows of (D_i, r_i, U_i')
    const int nk = (T + 2 * s - 1) >> l;      // bcr_n_kept(T, s)
    for (int e = e0; e < nk; e += EPP) {
    if (active_problem) {
      const int i = 2 * s * e;
      const int pi = nb + bcr_slot(lvl_off, T, i);
      const bool has_l = e > 0, has_r = (i + s) < T, has_rr = (i + 2 * s) < T;
      const int pl = nb + off_l + (has_l ? e - 1 : 0);   // slot of j = i - s
      const int pr = nb + off_l + (has_r ? e : 0);       // slot of j = i + s
      constexpr int NR = (D + LPN - 1) / LPN;            // rows per lane
#pragma unroll
      for (int rr = 0; rr < NR; ++rr) {
        const int a = lane + rr * LPN;
        if (a < D) {
          double drow[D], unew[D], ra;
          {
            const double* dp = bd.Dp(pi) + (a * (a + 1) / 2) * NN;
#pragma unroll
            for (int c = 0; c < D; ++c) drow[c] = (c <= a) ? dp[c * NN] : 0.0;
            ra = bd.Rp(pi)[a * NN];
#pragma unroll
            for (int c = 0; c < D; ++c) unew[c] = 0.0;
          }
          if (has_l) {   // D_i -= F^T F, r_i -= F^T g  with F, g of j = i - s
            const double* fp = bd.Up(pl);
            const double* gp = bd.Rp(pl);
#pragma unroll
            for (int k = 0; k < D; ++k) {
              const double fka = fp[(k * D) * NN + a * NN];
              ra -= fka * gp[k * NN];
#pragma unroll
              for (int c = 0; c < D; ++c) drow[c] -= fka * fp[(k * D + c) * NN];
            }
          }
          if (has_r) {   // D_i -= E^T E, r_i -= E^T g, U_i' = -E^T F  with E, F, g of j = i + s
            const double* ep = bd.Ep(pr);
            const double* fp = bd.Up(pr);
            const double* gp = bd.Rp(pr);
#pragma unroll
            for (int k = 0; k < D; ++k) {
              const double eka = ep[(k * D) * NN + a * NN];
              ra -= eka * gp[k * NN];
#pragma unroll
              for (int c = 0; c < D; ++c) {
                drow[c] -= eka * ep[(k * D + c) * NN];
                unew[c] -= eka * fp[(k * D + c) * NN];
              }
            }
          }
          {
            double* dp = bd.Dp(pi) + (a * (a + 1) / 2) * NN;
#pragma unroll
            for (int c = 0; c < D; ++c)
              if (c <= a) dp[c * NN] = drow[c];
            bd.Rp(pi)[a * NN] = ra;
            if (has_rr) {
              double* up = bd.Up(pi) + a * D * NN;
#pragma unroll
              for (int c = 0; c < D; ++c) up[c * NN] = unew[c];
            }
          }
        }
      }
    }
    }
    __syncthreads();
  }

  // ------------------------------ root (t = 0) ------------------------------
  if (active_problem && u == 0) {
    const int p0 = nb + (T - 1);
    double L[DS], v[D];
    const double* dp = bd.Dp(p0);
#pragma unroll
    for (int k = 0; k < DS; ++k) L[k] = dp[k * NN];
    double* rp = bd.Rp(p0);
#pragma unroll
    for (int a = 0; a < D; ++a) v[a] = rp[a * NN];
    if (!chol_packed<D>(L)) atomicMax(&fail[p], 1);
    fwd_solve<D>(L, v);
    bwd_solve<D>(L, v);
#pragma unroll
    for (int a = 0; a < D; ++a) rp[a * NN] = v[a];
  }
  __syncthreads();

  // ------------------------------ back substitution ------------------------------
  for (int l = nlev; l >= 1; --l) {
    const int s = 1 << (l - 1);
    const int ne = (T + s - 1) >> l;
    const int off_l = lvl_off[l];
    for (int e = e0;; e += EPP) {
    const unsigned m_bs = __ballot_sync(0xffffffffu, active_problem && e < ne);
    if (m_bs == 0u) break;
    if (active_problem && e < ne) {
      const int j = s * (2 * e + 1);
      const int pj = nb + off_l + e;
      const int pi = nb + bcr_slot(lvl_off, T, j - s);
      const bool has_right = (j + s) < T;
      const int pk = has_right ? nb + bcr_slot(lvl_off, T, j + s) : pi;
      double xl[D], xr[D], v[D], L[DS];
      const double* rpi = bd.Rp(pi);
      const double* rpk = bd.Rp(pk);
      double* rpj = bd.Rp(pj);
#pragma unroll
      for (int a = 0; a < D; ++a) {
        xl[a] = rpi[a * NN];
        xr[a] = has_right ? rpk[a * NN] : 0.0;
        v[a] = rpj[a * NN];
      }
      const double* ep = bd.Ep(pj);
      const double* fp = bd.Up(pj);
      const double* dp = bd.Dp(pj);
#pragma unroll
      for (int k = 0; k < DS; ++k) L[k] = dp[k * NN];
#pragma unroll
      for (int a = 0; a < D; ++a) {
        double acc0 = v[a], acc1 = 0.0;
#pragma unroll
        for (int c = 0; c < D; ++c) {
          acc0 -= ep[(a * D + c) * NN] * xl[c];
          acc1 -= fp[(a * D + c) * NN] * xr[c];
        }
        v[a] = acc0 + acc1;
      }
      bwd_solve<D>(L, v);
      __syncwarp(m_bs);   // all lanes have read g_j before any lane overwrites it with x_j
      // all lanes hold the full x_j; lane ln stores entries ln, ln + LPN, ...
#pragma unroll
      for (int a = 0; a < D; ++a)
        if ((a % LPN) == lane) rpj[a * NN] = v[a];
    }
    }
    __syncthreads();
  }
}

}  // namespace dgpmp2
