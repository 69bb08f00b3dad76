// Block cyclic reduction (BCR) of the SPD block-tridiagonal Gauss-Newton system,
// entirely in shared memory (sm_100a).
//
// Replaces the reference's dense normal equations + dense Cholesky + two dense
// inverses (plan_layer.py:214-234).
//
// Algorithm.  Levels l = 1..L, stride s = 2^(l-1).  At level l the nodes j = s(2q+1)
// are eliminated: L_j L_j^T = D_j, E_j = L_j^-1 U_{j-s}^T, F_j = L_j^-1 U_j,
// g_j = L_j^-1 r_j.  The kept neighbours i = j-s, k = j+s receive
//   D_i -= E_j^T E_j   r_i -= E_j^T g_j   U_i' = -E_j^T F_j
//   D_k -= F_j^T F_j   r_k -= F_j^T g_j
// Back substitution: x_j = L_j^-T (g_j - E_j x_{j-s} - F_j x_{j+s}).
// This is block Cholesky in nested-dissection order: backward stable for SPD
// systems, ceil(log2 T) dependent block steps instead of T.
//
// Layout.  One NODE record per trajectory state, array-of-structures, in LEVEL
// ORDER (nodes eliminated at level 1 first, ..., the root t = 0 last) so the work
// items of a level touch consecutive records:
//   [oD, oD+D*D)  D_t row-major          -> after elimination: L_t packed lower (1/l_kk on the diagonal)
//   [oU, oU+D*D)  U_t row-major          -> after elimination: F_t column-major
//   [oR, oR+D)    r_t -> g_t -> x_t
//   [oE, oE+D*D)  E_t column-major
//   [oX, oX+2)    this state's error partials (err, err_ext)
// Every field starts on a 16-byte boundary and the record stride is == 12 (mod 32)
// words, so all traffic is 128-bit LDS/STS and a quarter-warp touching 8 consecutive
// records is bank-conflict free.
//
// Work decomposition.  Every problem owns TPP threads; work item e of a level (one
// node) is processed by LPN = 2 lanes that split its column pairs (elimination) or
// row pairs (Schur update).  Lanes never exchange registers: everything goes through
// the records, ordered by __syncwarp inside an item and by the two CTA barriers per
// level.
#pragma once
#include "factors.cuh"

#ifndef DGPMP2_BCR_STAMP
#define DGPMP2_BCR_STAMP(i) do { } while (0)
#endif

namespace dgpmp2 {

constexpr int kMaxLevels = 16;
constexpr int kLPN = 2;

__host__ __device__ __forceinline__ int bcr_n_elim(int T, int s) { return (T + s - 1) / (2 * s); }   // nodes j = s(2q+1) < T
__host__ __device__ __forceinline__ int bcr_n_kept(int T, int s) { return (T + 2 * s - 1) / (2 * s); } // nodes i = 2sq < T

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

template <int D>
struct Node {
  static_assert(D % 2 == 0, "state dimension must be even");
  static constexpr int DD = D * D;
  static constexpr int DS = D * (D + 1) / 2;
  static constexpr int oD = 0, oU = DD, oR = 2 * DD, oE = 2 * DD + D, oX = 3 * DD + D;
  static constexpr int kRaw = 3 * DD + D + 2;
  static constexpr int kStride = (kRaw % 4 == 2) ? kRaw : kRaw + 2;   // doubles; == 2 (mod 4) -> 12 (mod 32) words for D = 4, 6
};

__device__ __forceinline__ double2 lds2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void sts2(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }

// 1/sqrt(x) for positive normal x: hardware seed + Newton steps, no special-case branches
// (a non-positive pivot is reported through the status flag, its value is then irrelevant).
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#pragma unroll
  for (int it = 0; it < 2; ++it) {   // seed is good to ~2^-22: two Newton steps reach double precision
    const double t = x * y;
    const double h = 0.5 * y;
    const double r = fma(-t, h, 0.5);
    y = fma(y, r, y);
  }
  return y;
}

// In-register Cholesky of a packed lower triangle; the diagonal is replaced by 1/l_kk.
// Returns false if a pivot is not strictly positive (incl. NaN).
template <int D>
__device__ __forceinline__ bool chol_packed(double (&L)[D * (D + 1) / 2]) {
  bool ok = true;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    const double akk = L[tri(k, k)];
    ok = ok && (akk > 0.0);
    const double rk = fast_rsqrt(akk);
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

// D contiguous doubles (16-byte aligned) <-> registers
template <int D>
__device__ __forceinline__ void ld_vec(const double* p, double (&v)[D]) {
#pragma unroll
  for (int a = 0; a < D; a += 2) {
    const double2 t = lds2(p + a);
    v[a] = t.x;
    v[a + 1] = t.y;
  }
}
template <int D>
__device__ __forceinline__ void st_vec(double* p, const double (&v)[D]) {
#pragma unroll
  for (int a = 0; a < D; a += 2) sts2(p + a, v[a], v[a + 1]);
}

// lower triangle of a row-major D x D block -> packed L
template <int D>
__device__ __forceinline__ void ld_lower(const double* p, double (&L)[D * (D + 1) / 2]) {
#pragma unroll
  for (int a = 0; a < D; ++a) {
#pragma unroll
    for (int c = 0; c <= a; c += 2) {
      const double2 t = lds2(p + a * D + c);
      L[tri(a, c)] = t.x;
      if (c + 1 <= a) L[tri(a, c + 1)] = t.y;
    }
  }
}

// exact floor(e / n) for 0 <= e < 2^20, 1 <= n <= 2^12 with inv = 1.0f / n (error analysis: the
// quotient (e + 0.5) / n is at least 0.5 / n away from an integer, the float error is < 2^-22 * e / n)
__device__ __forceinline__ int fast_div(int e, float inv) { return __float2int_rz(((float)e + 0.5f) * inv); }

// Factor + solve the CTA's np problems.  `nodes` = first record of problem 0 (np * T records,
// problem-major).  The work items of a level are enumerated across ALL problems of the CTA
// (item m -> problem m / n_items, node m % n_items) and packed onto consecutive lane groups, so
// the sparse deep levels of several problems share warps.  Thread tid is lane (tid % LPN) of items
// tid / LPN, tid / LPN + blockDim / LPN, ...
// On exit every record's [oR, oR+D) holds x_t.  fail[p] (shared, pre-zeroed) receives t+1 of a node
// of problem p whose pivot was not positive.  Must be called by ALL threads of the CTA (barriers).
template <int D>
__device__ __forceinline__ void bcr_solve(double* __restrict__ nodes, const int* __restrict__ lvl_off, int nlev,
                                          int T, int np, int* fail) {
  using N = Node<D>;
  constexpr int DS = N::DS, S = N::kStride, LPN = kLPN, NP2 = D / 2;   // NP2 column / row pairs
  constexpr int NPL = (NP2 + LPN - 1) / LPN;                           // pairs per lane
  const int e0 = threadIdx.x / LPN, lane = threadIdx.x % LPN;
  const int EPP = blockDim.x / LPN;

  // ------------------------------ forward elimination ------------------------------
  int off_l = 0;
  for (int l = 1; l <= nlev; ++l) {
    const int s = 1 << (l - 1);
    const int ne = (T + s - 1) >> l;          // bcr_n_elim(T, s), 2s = 2^l
    // (a) factor the eliminated nodes; lanes split the column pairs of [U_i^T | U_j]
    const float inv_ne = 1.0f / (float)ne;
    for (int m = e0;; m += EPP) {
      const bool on = m < np * ne;
      const unsigned m_el = __ballot_sync(0xffffffffu, on);
      if (m_el == 0u) break;                  // warp-uniform
      if (on) {
        const int p = fast_div(m, inv_ne), e = m - p * ne;
        double* pn = nodes + (size_t)p * T * S;
        const int j = s * (2 * e + 1);
        double* nj = pn + (size_t)(off_l + e) * S;
        const double* ni = pn + (size_t)bcr_slot(lvl_off, T, j - s) * S;
        const bool has_right = (j + s) < T;
        double L[DS];
        ld_lower<D>(nj + N::oD, L);
        // this lane's column pairs: issue their loads before the Cholesky chain
        double ve[NPL][2][D], vf[NPL][2][D], vg[D];
#pragma unroll
        for (int q = 0; q < NPL; ++q) {
          const int cp = lane + q * LPN;
          if (cp < NP2) {
            ld_vec<D>(ni + N::oU + (2 * cp) * D, ve[q][0]);          // row c of U_i = column c of U_i^T
            ld_vec<D>(ni + N::oU + (2 * cp + 1) * D, ve[q][1]);
#pragma unroll
            for (int a = 0; a < D; ++a) {                             // columns (2cp, 2cp+1) of U_j, row-major
              const double2 t = lds2(nj + N::oU + a * D + 2 * cp);
              vf[q][0][a] = has_right ? t.x : 0.0;
              vf[q][1][a] = has_right ? t.y : 0.0;
            }
          }
        }
        ld_vec<D>(nj + N::oR, vg);
        __syncwarp(m_el);   // every lane of the item has read D_j, U_j, r_j before they are overwritten
        if (!chol_packed<D>(L)) atomicMax(&fail[p], j + 1);
        if (lane == 0) {
#pragma unroll
          for (int k = 0; k < DS; k += 2) sts2(nj + N::oD + k, L[k], (k + 1 < DS) ? L[k + 1] : 0.0);
        }
#pragma unroll
        for (int q = 0; q < NPL; ++q) {
          const int cp = lane + q * LPN;
          if (cp < NP2) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              fwd_solve<D>(L, ve[q][h]);
              fwd_solve<D>(L, vf[q][h]);
              st_vec<D>(nj + N::oE + (2 * cp + h) * D, ve[q][h]);    // column of E_j, column-major
              st_vec<D>(nj + N::oU + (2 * cp + h) * D, vf[q][h]);    // column of F_j, column-major, in place
            }
          }
        }
        if (lane == 0) {
          fwd_solve<D>(L, vg);
          st_vec<D>(nj + N::oR, vg);
        }
      }
    }
    __syncthreads();
    DGPMP2_BCR_STAMP(8 + 2 * l);
    // (b) Schur-complement update of the kept nodes; lanes split the row pairs of (D_i, r_i, U_i')
    const int nk = (T + 2 * s - 1) >> l;      // bcr_n_kept(T, s)
    const float inv_nk = 1.0f / (float)nk;
    for (int m = e0; m < np * nk; m += EPP) {
      {
        const int p = fast_div(m, inv_nk), e = m - p * nk;
        double* pn = nodes + (size_t)p * T * S;
        const int i = 2 * s * e;
        double* ni = pn + (size_t)bcr_slot(lvl_off, T, i) * S;
        const bool has_l = e > 0, has_r = (i + s) < T, has_rr = (i + 2 * s) < T;
        const double* nl = pn + (size_t)(off_l + (has_l ? e - 1 : 0)) * S;   // j = i - s
        const double* nr = pn + (size_t)(off_l + (has_r ? e : 0)) * S;       // j = i + s
#pragma unroll
        for (int q = 0; q < NPL; ++q) {
          const int rp = lane + q * LPN;
          if (rp < NP2) {
            const int a0 = 2 * rp;
            double d0[D], d1[D], r0, r1;
            ld_vec<D>(ni + N::oD + a0 * D, d0);
            ld_vec<D>(ni + N::oD + (a0 + 1) * D, d1);
            {
              const double2 t = lds2(ni + N::oR + a0);
              r0 = t.x; r1 = t.y;
            }
            if (has_l) {   // D_i -= F^T F, r_i -= F^T g with F (column-major), g of j = i - s
              double fa0[D], fa1[D], g[D];
              ld_vec<D>(nl + N::oU + a0 * D, fa0);
              ld_vec<D>(nl + N::oU + (a0 + 1) * D, fa1);
              ld_vec<D>(nl + N::oR, g);
#pragma unroll
              for (int k = 0; k < D; ++k) { r0 -= fa0[k] * g[k]; r1 -= fa1[k] * g[k]; }
#pragma unroll
              for (int c = 0; c < D; ++c) {
                double fc[D];
                ld_vec<D>(nl + N::oU + c * D, fc);
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int k = 0; k < D; ++k) { s0 += fa0[k] * fc[k]; s1 += fa1[k] * fc[k]; }
                d0[c] -= s0; d1[c] -= s1;
              }
            }
            double u0[D], u1[D];
#pragma unroll
            for (int c = 0; c < D; ++c) { u0[c] = 0.0; u1[c] = 0.0; }
            if (has_r) {   // D_i -= E^T E, r_i -= E^T g, U_i' = -E^T F with E, F, g of j = i + s
              double ea0[D], ea1[D], g[D];
              ld_vec<D>(nr + N::oE + a0 * D, ea0);
              ld_vec<D>(nr + N::oE + (a0 + 1) * D, ea1);
              ld_vec<D>(nr + N::oR, g);
#pragma unroll
              for (int k = 0; k < D; ++k) { r0 -= ea0[k] * g[k]; r1 -= ea1[k] * g[k]; }
#pragma unroll
              for (int c = 0; c < D; ++c) {
                double ec[D], fc[D];
                ld_vec<D>(nr + N::oE + c * D, ec);
                ld_vec<D>(nr + N::oU + c * D, fc);
                double s0 = 0.0, s1 = 0.0, t0 = 0.0, t1 = 0.0;
#pragma unroll
                for (int k = 0; k < D; ++k) {
                  s0 += ea0[k] * ec[k]; s1 += ea1[k] * ec[k];
                  t0 += ea0[k] * fc[k]; t1 += ea1[k] * fc[k];
                }
                d0[c] -= s0; d1[c] -= s1;
                u0[c] = -t0; u1[c] = -t1;
              }
            }
            st_vec<D>(ni + N::oD + a0 * D, d0);
            st_vec<D>(ni + N::oD + (a0 + 1) * D, d1);
            sts2(ni + N::oR + a0, r0, r1);
            if (has_rr) {
              st_vec<D>(ni + N::oU + a0 * D, u0);          // new coupling to i + 2s, row-major
              st_vec<D>(ni + N::oU + (a0 + 1) * D, u1);
            }
          }
        }
      }
    }
    __syncthreads();
    DGPMP2_BCR_STAMP(9 + 2 * l);
    off_l += ne;
  }

  // ------------------------------ root (t = 0) ------------------------------
  for (int p = threadIdx.x; p < np; p += blockDim.x) {
    double* n0 = nodes + ((size_t)p * T + (T - 1)) * S;
    double L[DS], v[D];
    ld_lower<D>(n0 + N::oD, L);
    ld_vec<D>(n0 + N::oR, v);
    if (!chol_packed<D>(L)) atomicMax(&fail[p], 1);
    fwd_solve<D>(L, v);
    bwd_solve<D>(L, v);
    st_vec<D>(n0 + N::oR, v);
  }
  __syncthreads();
  DGPMP2_BCR_STAMP(5);

  // ------------------------------ back substitution ------------------------------
  for (int l = nlev; l >= 1; --l) {
    const int s = 1 << (l - 1);
    const int ne = (T + s - 1) >> l;
    off_l -= ne;
    const float inv_ne = 1.0f / (float)ne;
    for (int m = e0;; m += EPP) {
      const bool on = m < np * ne;
      const unsigned m_bs = __ballot_sync(0xffffffffu, on);
      if (m_bs == 0u) break;
      if (on) {
        const int p = fast_div(m, inv_ne), e = m - p * ne;
        double* pn = nodes + (size_t)p * T * S;
        const int j = s * (2 * e + 1);
        double* nj = pn + (size_t)(off_l + e) * S;
        const double* ni = pn + (size_t)bcr_slot(lvl_off, T, j - s) * S;
        const bool has_right = (j + s) < T;
        const double* nk2 = has_right ? pn + (size_t)bcr_slot(lvl_off, T, j + s) * S : ni;
        double xl[D], xr[D], v[D], L[DS];
        ld_vec<D>(ni + N::oR, xl);
        ld_vec<D>(nk2 + N::oR, xr);
        ld_vec<D>(nj + N::oR, v);
#pragma unroll
        for (int k = 0; k < DS; k += 2) {
          const double2 t = lds2(nj + N::oD + k);
          L[k] = t.x;
          if (k + 1 < DS) L[k + 1] = t.y;
        }
#pragma unroll
        for (int c = 0; c < D; ++c) {
          double ec[D], fc[D];
          ld_vec<D>(nj + N::oE + c * D, ec);
          ld_vec<D>(nj + N::oU + c * D, fc);
          const double xrc = has_right ? xr[c] : 0.0;
#pragma unroll
          for (int a = 0; a < D; ++a) v[a] -= ec[a] * xl[c] + fc[a] * xrc;
        }
        bwd_solve<D>(L, v);
        __syncwarp(m_bs);   // all lanes have read g_j before any lane overwrites it with x_j
        // every lane holds the full x_j; lane ln stores the pairs ln, ln + LPN, ...
#pragma unroll
        for (int rp = 0; rp < NP2; ++rp)
          if ((rp % LPN) == lane) sts2(nj + N::oR + 2 * rp, v[2 * rp], v[2 * rp + 1]);
      }
    }
    __syncthreads();
    DGPMP2_BCR_STAMP(40 + l);
  }
}

}  // namespace dgpmp2
