// Block cyclic reduction (BCR) of the SPD block-tridiagonal Gauss-Newton system,
// entirely in shared memory, for NP problems per CTA (sm_100a).
//
// Replaces the reference's dense normal equations + dense Cholesky + two dense
// inverses (plan_layer.py:214-234).  The T diagonal blocks D_t (d x d), the T-1
// couplings U_t = Lambda_{t,t+1} and the right-hand side r_t of every problem
// live in shared memory as structure-of-arrays [element][node] so that a warp
// touching one element of 32 consecutive nodes is bank-conflict free.
//
// Levels l = 1..L, stride s = 2^(l-1).  At level l the nodes j = s(2q+1) are
// eliminated: L_j L_j^T = D_j, E_j = L_j^-1 U_{j-s}^T, F_j = L_j^-1 U_j,
// g_j = L_j^-1 r_j.  The kept neighbours i = j-s, k = j+s receive
//   D_i -= E_j^T E_j   r_i -= E_j^T g_j   U_i' = -E_j^T F_j
//   D_k -= F_j^T F_j   r_k -= F_j^T g_j
// Back substitution: x_j = L_j^-T (g_j - E_j x_{j-s} - F_j x_{j+s}).
// This is block Cholesky in nested-dissection order: backward stable for SPD
// systems, log2(T) dependent block steps instead of T.
//
// Nodes are stored in LEVEL ORDER: the nodes eliminated at level 1 come first,
// then level 2, ..., the root (t = 0) last, so that the work items of every level
// read and write contiguous node slots.
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

// Shared-memory view of the band of all problems of the CTA.  NN = NP*T slots.
template <int D>
struct BcrSmem {
  static constexpr int DS = D * (D + 1) / 2;
  static constexpr int DD = D * D;
  static constexpr int kDoublesPerNode = DS + DD + D + DD;
  double* Dm;   // [DS][NN]  lower triangle of D_t; after elimination: L_t with 1/l_kk on the diagonal
  double* Um;   // [DD][NN]  U_t (row-major a*D+b); after elimination: F_t
  double* Rm;   // [D ][NN]  r_t; after elimination g_t; after back substitution x_t
  double* Em;   // [DD][NN]  E_t
  int NN;
  __device__ __forceinline__ void carve(double* base, int nn) {
    NN = nn;
    Dm = base;
    Um = Dm + (size_t)DS * nn;
    Rm = Um + (size_t)DD * nn;
    Em = Rm + (size_t)D * nn;
  }
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

// Factor + solve all NP block-tridiagonal systems held in `sm`.  On exit Rm holds
// the solution x_t of every node (slot order).  fail[p] (shared, pre-zeroed) receives
// t+1 of a node whose pivot was not positive.  Must be called by all threads of the CTA.
template <int D>
__device__ __forceinline__ void bcr_solve(const BcrSmem<D>& sm, const int* __restrict__ lvl_off, int nlev,
                                          int NP, int T, int* fail) {
  constexpr int DS = D * (D + 1) / 2;
  const int NN = sm.NN;
  const int tid = threadIdx.x, nthr = blockDim.x;

  // ------------------------------ forward elimination ------------------------------
  for (int l = 1; l <= nlev; ++l) {
    const int s = 1 << (l - 1);
    const int ne = bcr_n_elim(T, s);
    const int off_l = lvl_off[l];
    // (a) factor the eliminated nodes
    for (int e = tid; e < NP * ne; e += nthr) {
      const int p = e / ne, q = e - p * ne;
      const int j = s * (2 * q + 1);
      const int pj = p * T + off_l + q;
      const int pi = p * T + bcr_slot(lvl_off, T, j - s);
      const bool has_right = (j + s) < T;
      double L[DS];
#pragma unroll
      for (int k = 0; k < DS; ++k) L[k] = sm.Dm[k * NN + pj];
      if (!chol_packed<D>(L)) atomicMax(&fail[p], j + 1);
#pragma unroll
      for (int k = 0; k < DS; ++k) sm.Dm[k * NN + pj] = L[k];
      // E_j = L^-1 U_i^T : column c of U_i^T is row c of U_i
#pragma unroll
      for (int c = 0; c < D; ++c) {
        double v[D];
#pragma unroll
        for (int a = 0; a < D; ++a) v[a] = sm.Um[(c * D + a) * NN + pi];
        fwd_solve<D>(L, v);
#pragma unroll
        for (int a = 0; a < D; ++a) sm.Em[(a * D + c) * NN + pj] = v[a];
      }
      // F_j = L^-1 U_j (zero when there is no right neighbour)
#pragma unroll
      for (int c = 0; c < D; ++c) {
        double v[D];
#pragma unroll
        for (int a = 0; a < D; ++a) v[a] = has_right ? sm.Um[(a * D + c) * NN + pj] : 0.0;
        fwd_solve<D>(L, v);
#pragma unroll
        for (int a = 0; a < D; ++a) sm.Um[(a * D + c) * NN + pj] = v[a];
      }
      {
        double v[D];
#pragma unroll
        for (int a = 0; a < D; ++a) v[a] = sm.Rm[a * NN + pj];
        fwd_solve<D>(L, v);
#pragma unroll
        for (int a = 0; a < D; ++a) sm.Rm[a * NN + pj] = v[a];
      }
    }
    __syncthreads();
    // (b) Schur-complement update of the kept nodes
    const int nk = bcr_n_kept(T, s);
    for (int e = tid; e < NP * nk; e += nthr) {
      const int p = e / nk, q = e - p * nk;
      const int i = 2 * s * q;
      const int pi = p * T + bcr_slot(lvl_off, T, i);
      const bool has_l = q > 0, has_r = (i + s) < T;
      const int pl = p * T + off_l + (q - 1);   // slot of j = i - s  (its q index is q-1)
      const int pr = p * T + off_l + q;         // slot of j = i + s
      double Dl[DS], r[D];
#pragma unroll
      for (int k = 0; k < DS; ++k) Dl[k] = sm.Dm[k * NN + pi];
#pragma unroll
      for (int a = 0; a < D; ++a) r[a] = sm.Rm[a * NN + pi];
      if (has_l) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
          double f[D];
#pragma unroll
          for (int a = 0; a < D; ++a) f[a] = sm.Um[(k * D + a) * NN + pl];
          const double gk = sm.Rm[k * NN + pl];
#pragma unroll
          for (int a = 0; a < D; ++a) {
            r[a] -= f[a] * gk;
#pragma unroll
            for (int c = 0; c <= a; ++c) Dl[tri(a, c)] -= f[a] * f[c];
          }
        }
      }
      if (has_r) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
          double ev[D];
#pragma unroll
          for (int a = 0; a < D; ++a) ev[a] = sm.Em[(k * D + a) * NN + pr];
          const double gk = sm.Rm[k * NN + pr];
#pragma unroll
          for (int a = 0; a < D; ++a) {
            r[a] -= ev[a] * gk;
#pragma unroll
            for (int c = 0; c <= a; ++c) Dl[tri(a, c)] -= ev[a] * ev[c];
          }
        }
      }
#pragma unroll
      for (int k = 0; k < DS; ++k) sm.Dm[k * NN + pi] = Dl[k];
#pragma unroll
      for (int a = 0; a < D; ++a) sm.Rm[a * NN + pi] = r[a];
      // new coupling to i + 2s:  U_i' = -E_j^T F_j  (j = i + s); only needed if i + 2s exists
      if ((i + 2 * s) < T) {
#pragma unroll
        for (int a = 0; a < D; ++a) {
          double acc[D];
#pragma unroll
          for (int c = 0; c < D; ++c) acc[c] = 0.0;
#pragma unroll
          for (int k = 0; k < D; ++k) {
            const double eka = sm.Em[(k * D + a) * NN + pr];
#pragma unroll
            for (int c = 0; c < D; ++c) acc[c] -= eka * sm.Um[(k * D + c) * NN + pr];
          }
#pragma unroll
          for (int c = 0; c < D; ++c) sm.Um[(a * D + c) * NN + pi] = acc[c];
        }
      }
    }
    __syncthreads();
  }

  // ------------------------------ root (t = 0) ------------------------------
  for (int p = tid; p < NP; p += nthr) {
    const int p0 = p * T + (T - 1);
    double L[DS], v[D];
#pragma unroll
    for (int k = 0; k < DS; ++k) L[k] = sm.Dm[k * NN + p0];
    if (!chol_packed<D>(L)) atomicMax(&fail[p], 1);
#pragma unroll
    for (int a = 0; a < D; ++a) v[a] = sm.Rm[a * NN + p0];
    fwd_solve<D>(L, v);
    bwd_solve<D>(L, v);
#pragma unroll
    for (int a = 0; a < D; ++a) sm.Rm[a * NN + p0] = v[a];
  }
  __syncthreads();

  // ------------------------------ back substitution ------------------------------
  for (int l = nlev; l >= 1; --l) {
    const int s = 1 << (l - 1);
    const int ne = bcr_n_elim(T, s);
    const int off_l = lvl_off[l];
    for (int e = tid; e < NP * ne; e += nthr) {
      const int p = e / ne, q = e - p * ne;
      const int j = s * (2 * q + 1);
      const int pj = p * T + off_l + q;
      const int pi = p * T + bcr_slot(lvl_off, T, j - s);
      const bool has_right = (j + s) < T;
      const int pk = has_right ? p * T + bcr_slot(lvl_off, T, j + s) : pi;
      double xl[D], xr[D], v[D], L[DS];
#pragma unroll
      for (int a = 0; a < D; ++a) {
        xl[a] = sm.Rm[a * NN + pi];
        xr[a] = has_right ? sm.Rm[a * NN + pk] : 0.0;
        v[a] = sm.Rm[a * NN + pj];
      }
#pragma unroll
      for (int a = 0; a < D; ++a) {
        double acc = v[a];
#pragma unroll
        for (int c = 0; c < D; ++c) {
          acc -= sm.Em[(a * D + c) * NN + pj] * xl[c];
          acc -= sm.Um[(a * D + c) * NN + pj] * xr[c];
        }
        v[a] = acc;
      }
#pragma unroll
      for (int k = 0; k < DS; ++k) L[k] = sm.Dm[k * NN + pj];
      bwd_solve<D>(L, v);
#pragma unroll
      for (int a = 0; a < D; ++a) sm.Rm[a * NN + pj] = v[a];
    }
    __syncthreads();
  }
}

}  // namespace dgpmp2
