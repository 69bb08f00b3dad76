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
// Work decomposition.  The work items of a level (one node each) are enumerated across
// all problems of the CTA.  On the wide levels (>= plan.wide_min items in the CTA) one lane
// handles one item: those levels are bound by the shared-memory pipe and one lane per item
// halves the wavefronts.  On the narrow levels an item is processed by kLPN = 4 adjacent lanes
// that split its columns (elimination), rows (Schur update) or rows of the right-hand side
// (back substitution): those levels are bound by the dependent chain of one item.  Lanes never
// exchange registers: everything goes through the records, ordered by __syncwarp inside an
// item and by the two CTA barriers per level.  Once <= plan.tail_nc nodes per problem are
// left, the chain is finished by sequential block Cholesky (bcr_tail) without CTA barriers.
// Every multiply-add is an explicit fma / __dmul_rn / __dadd_rn, so all instantiations round
// identically and a problem's result does not depend on its position in the batch.
#pragma once
#include "factors.cuh"

#ifndef DGPMP2_BCR_STAMP
#define DGPMP2_BCR_STAMP(i) do { } while (0)
#endif

// profiling aid (scratch/rep_prof.py): -DDGPMP2_REPEAT=n -DDGPMP2_REPEAT_WHAT=1|2|3 -DDGPMP2_REPEAT_LEVEL=l repeats
// one phase of one level n times (results become meaningless) so that ncu's per-kernel counters describe that phase
#if defined(DGPMP2_REPEAT)
#define DGPMP2_REP(what, l) for (int rep_ = 0; rep_ < (((what) == DGPMP2_REPEAT_WHAT && (l) == DGPMP2_REPEAT_LEVEL) ? DGPMP2_REPEAT : 1); ++rep_)
#else
#define DGPMP2_REP(what, l)
#endif

namespace dgpmp2 {

// slot of trajectory state t inside its problem.  Closed form of off[l] + (t >> l) with
// l = ctz(t) + 1 and off[l] = #{1 <= u < T : ctz(u) + 1 < l} = (T-1) - ((T-1) >> (l-1)).
__device__ __forceinline__ int bcr_slot(int T, int t) {
  const int z = __ffs(t) - 1;      // ctz(t); -1 for t == 0
  const int s = (T - 1) - ((T - 1) >> z) + (t >> (z + 1));
  return (t == 0) ? (T - 1) : s;
}
// inverse: trajectory state stored in slot m.  The level is the first l with off[l+1] > m.
__device__ __forceinline__ int bcr_state_of_slot(int T, int m) {
  if (m == T - 1) return 0;
  int l = 1;
  while ((T - 1) - ((T - 1) >> l) <= m) ++l;
  const int off = (T - 1) - ((T - 1) >> (l - 1));
  return (2 * (m - off) + 1) << (l - 1);
}

template <int D>
struct Node {
  static_assert(D % 2 == 0, "state dimension must be even");
  static constexpr int DD = D * D;
  static constexpr int DS = D * (D + 1) / 2;
  static constexpr int oD = 0, oU = DD, oR = 2 * DD, oE = 2 * DD + D, oX = 3 * DD + D;
  static constexpr int kRaw = 3 * DD + D + 2;
  static constexpr int kStride = (kRaw % 4 == 2) ? kRaw : kRaw + 2;   // doubles; == 2 (mod 4) -> 12 (mod 32) words for D = 4, 6
  // Records of consecutive problems are staggered by 16 bytes (4 banks): when T * kStride is a multiple of 32 words
  // (every T that is a multiple of 8), lanes of different problems that read the same field of the same node would
  // otherwise hit the same banks (the tail and the deepest levels mix problems inside a quarter-warp).
  static constexpr int kProblemPad = 2;
  __host__ __device__ static constexpr size_t problem_stride(int T) { return (size_t)T * kStride + kProblemPad; }
};

__device__ __forceinline__ double2 lds2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void sts2(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }

// 1/sqrt(x) for positive normal x: hardware seed + one cubically convergent correction, no special-case
// branches (a non-positive pivot is reported through the status flag, its value is then irrelevant).
// With e = 1 - x y0^2:  1/sqrt(x) = y0 (1 - e)^(-1/2) = y0 (1 + e/2 + 3e^2/8 + O(e^3)); the seed is good to
// ~2^-20, so the O(e^3) remainder is below 2^-60.  Four dependent operations after the MUFU instead of the
// six of two Newton steps - this sits on the critical path of every Cholesky pivot.
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
  const double t = __dmul_rn(x, y0);
  const double e = fma(-t, y0, 1.0);
  const double a = __dmul_rn(y0, e);
  const double p = fma(e, 0.375, 0.5);
  return fma(a, p, y0);
}

// c - a * b with a single rounding.  Every multiply-add of the factorisation is written with explicit
// fma / __dmul_rn / __dadd_rn so that the compiler's contraction choices cannot differ between the
// LPN = 1 and LPN = 4 instantiations: a problem's result is bit-identical wherever it sits in a CTA.
__device__ __forceinline__ double fnma(double a, double b, double c) { return fma(-a, b, c); }
// a . b over D entries, accumulated left to right
template <int D>
__device__ __forceinline__ double dot(const double (&a)[D], const double (&b)[D]) {
  double s = __dmul_rn(a[0], b[0]);
#pragma unroll
  for (int k = 1; k < D; ++k) s = fma(a[k], b[k], s);
  return s;
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
    for (int i = k + 1; i < D; ++i) L[tri(i, k)] = __dmul_rn(L[tri(i, k)], rk);
#pragma unroll
    for (int j = k + 1; j < D; ++j)
#pragma unroll
      for (int i = j; i < D; ++i) L[tri(i, j)] = fnma(L[tri(i, k)], L[tri(j, k)], L[tri(i, j)]);
  }
  return ok;
}

template <int D>
__device__ __forceinline__ void fwd_solve(const double (&L)[D * (D + 1) / 2], double (&v)[D]) {
#pragma unroll
  for (int a = 0; a < D; ++a) {
    double s = v[a];
#pragma unroll
    for (int c = 0; c < a; ++c) s = fnma(L[tri(a, c)], v[c], s);
    v[a] = __dmul_rn(s, L[tri(a, a)]);
  }
}

template <int D>
__device__ __forceinline__ void bwd_solve(const double (&L)[D * (D + 1) / 2], double (&v)[D]) {
#pragma unroll
  for (int a = D - 1; a >= 0; --a) {
    double s = v[a];
#pragma unroll
    for (int c = a + 1; c < D; ++c) s = fnma(L[tri(c, a)], v[c], s);
    v[a] = __dmul_rn(s, L[tri(a, a)]);
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
// m / n and m % n for a divisor n that is the same for the whole level: a shift when n is a power of two
struct LevelDiv {
  int n, sh; float inv;
  __device__ __forceinline__ void split(int m, int& q, int& r) const {
    q = (sh >= 0) ? (m >> sh) : fast_div(m, inv);
    r = m - q * n;
  }
};
// ---------------------------------------------------------------------------------------------
// The three per-level phases, templated on LPN = lanes per work item.
//   LPN = 4: the lanes of an item split its columns / rows -> shortest dependent chain; used for the
//            narrow (deep) levels, which are latency bound.
//   LPN = 1: one lane per item, every access is a 128-bit LDS/STS at the same field offset of
//            consecutive records (bank-conflict free by the record stride) and nothing is loaded
//            twice -> about half the shared-memory wavefronts per item; used for the wide levels,
//            which are bound by the shared-memory pipe (profiles/README.md, isolated-phase runs).
// All functions must be called by every thread of the CTA with uniform arguments.
// ---------------------------------------------------------------------------------------------

// (a) factor the eliminated nodes j = s(2e+1): L_j, E_j = L^-1 U_{j-s}^T, F_j = L^-1 U_j, g_j = L^-1 r_j
template <int D, int LPN>
__device__ __forceinline__ void bcr_elim_level(double* __restrict__ nodes, int T, int np,
                                               int s, const LevelDiv dv_e, int off_l, int* fail) {
  using N = Node<D>;
  constexpr int DS = N::DS, S = N::kStride;
  constexpr int NCL = (D + LPN - 1) / LPN;                             // columns per lane
  const int e0 = threadIdx.x / LPN, lane = threadIdx.x % LPN;
  const int EPP = blockDim.x / LPN;
  const int ne = dv_e.n;
  for (int base = 0; base < np * ne; base += EPP) {   // uniform trip count for the whole CTA
    const int m = base + e0;
    const bool on = m < np * ne;
    const unsigned m_el = (LPN > 1) ? __ballot_sync(0xffffffffu, on) : 0u;
    if (on) {
      int p, e;
      dv_e.split(m, p, e);
      double* pn = nodes + (size_t)p * N::problem_stride(T);
      const int j = s * (2 * e + 1);
      double* nj = pn + (size_t)(off_l + e) * S;
      const double* ni = pn + (size_t)bcr_slot(T, j - s) * S;
      const bool has_right = (j + s) < T;
      double L[DS];
      ld_lower<D>(nj + N::oD, L);
      if constexpr (LPN == 1) {
        // two batches ([U_j | r_j], then U_i^T) keep the live set at L + D*D + D doubles: no spills
        double vf[D][D], vg[D];
#pragma unroll
        for (int a = 0; a < D; ++a) {
          double row[D];
          ld_vec<D>(nj + N::oU + a * D, row);                         // row a of U_j -> transposed in registers
#pragma unroll
          for (int c = 0; c < D; ++c) vf[c][a] = has_right ? row[c] : 0.0;
        }
        ld_vec<D>(nj + N::oR, vg);
        if (!chol_packed<D>(L)) atomicMax(&fail[p], j + 1);
#pragma unroll
        for (int k = 0; k < DS; k += 2) sts2(nj + N::oD + k, L[k], (k + 1 < DS) ? L[k + 1] : 0.0);
#pragma unroll
        for (int c = 0; c < D; ++c) fwd_solve<D>(L, vf[c]);
        fwd_solve<D>(L, vg);
#pragma unroll
        for (int c = 0; c < D; ++c) st_vec<D>(nj + N::oU + c * D, vf[c]);   // F_j column-major, in place
        st_vec<D>(nj + N::oR, vg);
#pragma unroll
        for (int c = 0; c < D; ++c) ld_vec<D>(ni + N::oU + c * D, vf[c]);   // row c of U_i = column c of U_i^T
#pragma unroll
        for (int c = 0; c < D; ++c) fwd_solve<D>(L, vf[c]);
#pragma unroll
        for (int c = 0; c < D; ++c) st_vec<D>(nj + N::oE + c * D, vf[c]);   // E_j column-major
      } else {
        // this lane's columns: issue their loads before the Cholesky chain
        double ve[NCL][D], vf[NCL][D], vg[D];
#pragma unroll
        for (int q = 0; q < NCL; ++q) {
          const int c = lane + q * LPN;
          if (c < D) {
            ld_vec<D>(ni + N::oU + c * D, ve[q]);                     // row c of U_i = column c of U_i^T
#pragma unroll
            for (int a = 0; a < D; ++a) vf[q][a] = has_right ? nj[N::oU + a * D + c] : 0.0;   // column c of U_j (row-major)
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
        for (int q = 0; q < NCL; ++q) {
          const int c = lane + q * LPN;
          if (c < D) {
            fwd_solve<D>(L, ve[q]);
            fwd_solve<D>(L, vf[q]);
            st_vec<D>(nj + N::oE + c * D, ve[q]);                     // column c of E_j, column-major
            st_vec<D>(nj + N::oU + c * D, vf[q]);                     // column c of F_j, column-major, in place
          }
        }
        fwd_solve<D>(L, vg);
        if (lane == 0) st_vec<D>(nj + N::oR, vg);
      }
    }
  }
}

// (b) Schur-complement update of the kept nodes i = 2se:
//   D_i -= F_l^T F_l + E_r^T E_r,  r_i -= F_l^T g_l + E_r^T g_r,  U_i' = -E_r^T F_r   (l: j = i - s, r: j = i + s)
template <int D, int LPN>
__device__ __forceinline__ void bcr_kept_level(double* __restrict__ nodes, int T, int np,
                                               int s, const LevelDiv dv_k, int off_l) {
  using N = Node<D>;
  constexpr int S = N::kStride;
  constexpr int NCL = (D + LPN - 1) / LPN;                             // rows per lane
  const int e0 = threadIdx.x / LPN, lane = threadIdx.x % LPN;
  const int EPP = blockDim.x / LPN;
  const int nk = dv_k.n;
  for (int m = e0; m < np * nk; m += EPP) {
    int p, e;
    dv_k.split(m, p, e);
    double* pn = nodes + (size_t)p * N::problem_stride(T);
    const int i = 2 * s * e;
    double* ni = pn + (size_t)bcr_slot(T, i) * S;
    const bool has_l = e > 0, has_r = (i + s) < T, has_rr = (i + 2 * s) < T;
    const double* nl = pn + (size_t)(off_l + (has_l ? e - 1 : 0)) * S;   // j = i - s
    const double* nr = pn + (size_t)(off_l + (has_r ? e : 0)) * S;       // j = i + s
    if constexpr (LPN == 1) {
      // only the lower triangle of D_i is ever read (ld_lower); the strict upper part is written as its mirror
      double Dl[N::DS], r[D];
      ld_lower<D>(ni + N::oD, Dl);
      ld_vec<D>(ni + N::oR, r);
      if (has_l) {
        double F[D][D], g[D];                                          // F[c][k] = F_l(k, c)
#pragma unroll
        for (int c = 0; c < D; ++c) ld_vec<D>(nl + N::oU + c * D, F[c]);
        ld_vec<D>(nl + N::oR, g);
#pragma unroll
        for (int a = 0; a < D; ++a) {
          r[a] = __dsub_rn(r[a], dot<D>(F[a], g));
#pragma unroll
          for (int c = 0; c <= a; ++c) Dl[tri(a, c)] = __dsub_rn(Dl[tri(a, c)], dot<D>(F[a], F[c]));
        }
      }
      double Em[D][D];                                                 // Em[c][k] = E_r(k, c)
      if (has_r) {
        double g[D];
#pragma unroll
        for (int c = 0; c < D; ++c) ld_vec<D>(nr + N::oE + c * D, Em[c]);
        ld_vec<D>(nr + N::oR, g);
#pragma unroll
        for (int a = 0; a < D; ++a) {
          r[a] = __dsub_rn(r[a], dot<D>(Em[a], g));
#pragma unroll
          for (int c = 0; c <= a; ++c) Dl[tri(a, c)] = __dsub_rn(Dl[tri(a, c)], dot<D>(Em[a], Em[c]));
        }
      }
#pragma unroll
      for (int a = 0; a < D; ++a)
#pragma unroll
        for (int c = 0; c < D; c += 2)
          sts2(ni + N::oD + a * D + c, (c <= a) ? Dl[tri(a, c)] : Dl[tri(c, a)], (c + 1 <= a) ? Dl[tri(a, c + 1)] : Dl[tri(c + 1, a)]);
      st_vec<D>(ni + N::oR, r);
      if (has_rr) {   // has_rr implies has_r
        double un[D][D];
#pragma unroll
        for (int c = 0; c < D; ++c) {
          double fc[D];
          ld_vec<D>(nr + N::oU + c * D, fc);
#pragma unroll
          for (int a = 0; a < D; ++a) un[a][c] = -dot<D>(Em[a], fc);
        }
#pragma unroll
        for (int a = 0; a < D; ++a) st_vec<D>(ni + N::oU + a * D, un[a]);   // new coupling to i + 2s, row-major
      }
    } else {
#pragma unroll
      for (int q = 0; q < NCL; ++q) {
        const int a = lane + q * LPN;
        if (a < D) {
          double drow[D], unew[D], ra;
          ld_vec<D>(ni + N::oD + a * D, drow);
          ra = ni[N::oR + a];
#pragma unroll
          for (int c = 0; c < D; ++c) unew[c] = 0.0;
          if (has_l) {   // D_i -= F^T F, r_i -= F^T g with F (column-major), g of j = i - s
            double fa[D], g[D];
            ld_vec<D>(nl + N::oU + a * D, fa);
            ld_vec<D>(nl + N::oR, g);
            ra = __dsub_rn(ra, dot<D>(fa, g));
#pragma unroll
            for (int c = 0; c < D; ++c) {
              double fc[D];
              ld_vec<D>(nl + N::oU + c * D, fc);
              drow[c] = __dsub_rn(drow[c], dot<D>(fa, fc));
            }
          }
          if (has_r) {   // D_i -= E^T E, r_i -= E^T g, U_i' = -E^T F with E, F, g of j = i + s
            double ea[D], g[D];
            ld_vec<D>(nr + N::oE + a * D, ea);
            ld_vec<D>(nr + N::oR, g);
            ra = __dsub_rn(ra, dot<D>(ea, g));
#pragma unroll
            for (int c = 0; c < D; ++c) {
              double ec[D], fc[D];
              ld_vec<D>(nr + N::oE + c * D, ec);
              ld_vec<D>(nr + N::oU + c * D, fc);
              drow[c] = __dsub_rn(drow[c], dot<D>(ea, ec));
              unew[c] = -dot<D>(ea, fc);
            }
          }
          st_vec<D>(ni + N::oD + a * D, drow);
          ni[N::oR + a] = ra;
          if (has_rr) st_vec<D>(ni + N::oU + a * D, unew);           // new coupling to i + 2s, row-major
        }
      }
    }
  }
}

// (c) back substitution of the nodes eliminated at this level: x_j = L_j^-T (g_j - E_j x_{j-s} - F_j x_{j+s})
template <int D, int LPN>
__device__ __forceinline__ void bcr_back_level(double* __restrict__ nodes, int T, int np,
                                               int s, const LevelDiv dv_e, int off_l) {
  using N = Node<D>;
  constexpr int DS = N::DS, S = N::kStride;
  constexpr int NCL = (D + LPN - 1) / LPN;                             // rows per lane
  const int e0 = threadIdx.x / LPN, lane = threadIdx.x % LPN;
  const int EPP = blockDim.x / LPN;
  const int ne = dv_e.n;
  for (int base = 0; base < np * ne; base += EPP) {
    const int m = base + e0;
    const bool on = m < np * ne;
    const unsigned m_bs = (LPN > 1) ? __ballot_sync(0xffffffffu, on) : 0u;
    if (on) {
      int p, e;
      dv_e.split(m, p, e);
      double* pn = nodes + (size_t)p * N::problem_stride(T);
      const int j = s * (2 * e + 1);
      double* nj = pn + (size_t)(off_l + e) * S;
      const double* ni = pn + (size_t)bcr_slot(T, j - s) * S;
      const bool has_right = (j + s) < T;
      const double* nk2 = has_right ? pn + (size_t)bcr_slot(T, j + s) * S : ni;
      double xl[D], xr[D], L[DS];
      ld_vec<D>(ni + N::oR, xl);
      ld_vec<D>(nk2 + N::oR, xr);
#pragma unroll
      for (int c = 0; c < D; ++c) xr[c] = has_right ? xr[c] : 0.0;
#pragma unroll
      for (int k = 0; k < DS; k += 2) {
        const double2 t = lds2(nj + N::oD + k);
        L[k] = t.x;
        if (k + 1 < DS) L[k + 1] = t.y;
      }
      if constexpr (LPN == 1) {
        double v[D], w[D];
        ld_vec<D>(nj + N::oR, v);
#pragma unroll
        for (int a = 0; a < D; ++a) w[a] = 0.0;
#pragma unroll
        for (int c = 0; c < D; ++c) {
          double ec[D], fc[D];
          ld_vec<D>(nj + N::oE + c * D, ec);
          ld_vec<D>(nj + N::oU + c * D, fc);
#pragma unroll
          for (int a = 0; a < D; ++a) {
            v[a] = fnma(ec[a], xl[c], v[a]);
            w[a] = fnma(fc[a], xr[c], w[a]);
          }
        }
#pragma unroll
        for (int a = 0; a < D; ++a) v[a] = __dadd_rn(v[a], w[a]);
        bwd_solve<D>(L, v);
        st_vec<D>(nj + N::oR, v);
      } else {
        // each lane forms its rows of v = g_j - E_j x_{j-s} - F_j x_{j+s} and publishes them in place of g_j
#pragma unroll
        for (int q = 0; q < NCL; ++q) {
          const int a = lane + q * LPN;
          if (a < D) {
            double va = nj[N::oR + a], vb = 0.0;
#pragma unroll
            for (int c = 0; c < D; ++c) {
              va = fnma(nj[N::oE + c * D + a], xl[c], va);
              vb = fnma(nj[N::oU + c * D + a], xr[c], vb);
            }
            nj[N::oR + a] = __dadd_rn(va, vb);
          }
        }
        __syncwarp(m_bs);
        double v[D];
        ld_vec<D>(nj + N::oR, v);
        bwd_solve<D>(L, v);
        __syncwarp(m_bs);   // all lanes have read v before it is overwritten with x_j
        if (lane == 0) st_vec<D>(nj + N::oR, v);   // (indexing v[] by the lane would put it in local memory)
      }
    }
  }
}

// (d) the chain that is left once the kept nodes are few: nodes t = S e, e = 0..nc-1, coupled through U.
// Solved by sequential block Cholesky (block Thomas) with kLPN lanes per problem: a step costs one
// Cholesky chain, one __syncwarp and no CTA barrier, which beats another elimination level (three
// phases, two CTA barriers) as soon as nc is small.
//   forward:  D_e' = D_e - F_{e-1}^T F_{e-1},  L_e L_e^T = D_e',  F_e = L_e^-1 U_e,  g_e = L_e^-1 (r_e - F_{e-1}^T g_{e-1})
//   backward: x_e = L_e^-T (g_e - F_e x_{e+1})
// Every lane forms D_e', its Cholesky factor, g_e and the whole back substitution redundantly (no
// exchange needed); the lanes only split the columns of F_e, which travel through the record.
// With nc == 1 this is the classic root solve of cyclic reduction.
template <int D>
__device__ __forceinline__ void bcr_tail(double* __restrict__ nodes, int T, int np,
                                         int S_t, int nc, int* fail) {
  using N = Node<D>;
  constexpr int DS = N::DS, S = N::kStride, LPN = kLPN;
  constexpr int NCL = (D + LPN - 1) / LPN;
  const int e0 = threadIdx.x / LPN, lane = threadIdx.x % LPN;
  const int EPP = blockDim.x / LPN;
  for (int base = 0; base < np; base += EPP) {          // uniform trip count (one pass unless blockDim < 4 np)
    const int p = base + e0;
    const bool on = p < np;
    const unsigned m_t = __ballot_sync(0xffffffffu, on);
    if (on) {
      double* pn = nodes + (size_t)p * N::problem_stride(T);
      double gp[D];
      const double* prev = pn;
#pragma unroll 1
      for (int e = 0; e < nc; ++e) {
        double* nd = pn + (size_t)bcr_slot(T, S_t * e) * S;
        const bool has_next = e + 1 < nc;
        double L[DS], r[D], vf[NCL][D];
#pragma unroll
        for (int q = 0; q < NCL; ++q) {
          const int c = lane + q * LPN;
          if (c < D) {
#pragma unroll
            for (int a = 0; a < D; ++a) vf[q][a] = has_next ? nd[N::oU + a * D + c] : 0.0;   // column c of U_e (row-major)
          }
        }
        if (e > 0) {
          // Schur update, split by rows over the lanes of the problem (it is fp64-issue bound: 14 dot products
          // in one lane cost more than the Cholesky chain) and exchanged through the record itself
#pragma unroll
          for (int q = 0; q < NCL; ++q) {
            const int a = lane + q * LPN;
            if (a < D) {
              double drow[D], fa[D];
              ld_vec<D>(nd + N::oD + a * D, drow);
              double ra = nd[N::oR + a];
              ld_vec<D>(prev + N::oU + a * D, fa);                     // column a of F_{e-1}
              ra = __dsub_rn(ra, dot<D>(fa, gp));
#pragma unroll
              for (int c = 0; c < D; ++c) {
                double fc[D];
                ld_vec<D>(prev + N::oU + c * D, fc);
                drow[c] = __dsub_rn(drow[c], dot<D>(fa, fc));
              }
              st_vec<D>(nd + N::oD + a * D, drow);
              nd[N::oR + a] = ra;
            }
          }
        }
        __syncwarp(m_t);   // rows of D_e', r_e' are published; every lane has read its column of U_e
        ld_lower<D>(nd + N::oD, L);
        ld_vec<D>(nd + N::oR, r);
        __syncwarp(m_t);   // every lane holds D_e', r_e' before they are overwritten by L_e, g_e / F_e
        if (!chol_packed<D>(L)) atomicMax(&fail[p], S_t * e + 1);
        fwd_solve<D>(L, r);
#pragma unroll
        for (int q = 0; q < NCL; ++q) {
          const int c = lane + q * LPN;
          if (c < D && has_next) {
            fwd_solve<D>(L, vf[q]);
            st_vec<D>(nd + N::oU + c * D, vf[q]);                      // column c of F_e, column-major, in place
          }
        }
        if (lane == 0) {
#pragma unroll
          for (int k = 0; k < DS; k += 2) sts2(nd + N::oD + k, L[k], (k + 1 < DS) ? L[k + 1] : 0.0);
          st_vec<D>(nd + N::oR, r);
        }
#pragma unroll
        for (int a = 0; a < D; ++a) gp[a] = r[a];
        prev = nd;
        __syncwarp(m_t);   // F_e, L_e, g_e are visible to the other lanes of the problem
      }
      DGPMP2_BCR_STAMP(6);
      double x[D];
#pragma unroll
      for (int a = 0; a < D; ++a) x[a] = 0.0;
#pragma unroll 1
      for (int e = nc - 1; e >= 0; --e) {
        double* nd = pn + (size_t)bcr_slot(T, S_t * e) * S;
        double L[DS], v[D];
#pragma unroll
        for (int k = 0; k < DS; k += 2) {
          const double2 t = lds2(nd + N::oD + k);
          L[k] = t.x;
          if (k + 1 < DS) L[k + 1] = t.y;
        }
        ld_vec<D>(nd + N::oR, v);
        if (e + 1 < nc) {
#pragma unroll
          for (int c = 0; c < D; ++c) {
            double fc[D];
            ld_vec<D>(nd + N::oU + c * D, fc);
#pragma unroll
            for (int a = 0; a < D; ++a) v[a] = fnma(fc[a], x[c], v[a]);
          }
        }
        bwd_solve<D>(L, v);
#pragma unroll
        for (int a = 0; a < D; ++a) x[a] = v[a];
        __syncwarp(m_t);   // every lane has read g_e before lane 0 replaces it with x_e
        if (lane == 0) st_vec<D>(nd + N::oR, v);
      }
    }
  }
}

// Factor + solve the CTA's np problems.  `nodes` = first record of problem 0 (np * T records,
// problem-major).  The work items of a level are enumerated across ALL problems of the CTA
// (item m -> problem m / n_items, node m % n_items) and packed onto consecutive lane groups, so
// the sparse deep levels of several problems share warps.  A level with at least plan.wide_min items
// in the CTA runs one lane per item, the others kLPN lanes per item.  Elimination stops after
// plan.nl levels; the remaining chain of plan.tail_nc nodes is solved sequentially (bcr_tail).
// level1_eliminated (uniform): the records of the level-1 nodes already hold L, E, F, g (skip that phase).
// On exit every record's [oR, oR+D) holds x_t.  fail[p] (shared, pre-zeroed) receives t+1 of a node
// of problem p whose pivot was not positive.  Must be called by ALL threads of the CTA (barriers).
// side_work(first_idle_warp) is called by every thread between the sequential tail and its barrier: the tail keeps only
// the first ceil(kLPN * np / 32) warps busy, so the caller can give the others something that does not depend on the
// solve (gn_step_kernel: the per-problem error reductions).
struct BcrNoSideWork { __device__ __forceinline__ void operator()(int) const {} };
template <int D, typename SideWork = BcrNoSideWork>
__device__ __forceinline__ void bcr_solve(double* __restrict__ nodes, const BcrPlan& plan, int T, int np, int* fail,
                                          bool level1_eliminated = false, SideWork side_work = SideWork()) {
  const int nl = plan.nl, wide_min = plan.wide_min;

  // ------------------------------ forward elimination ------------------------------
  int off_l = 0;
  for (int l = 1; l <= nl; ++l) {
    const int s = 1 << (l - 1);
    const BcrLevelPlan& lv = plan.lv[l - 1];
    const LevelDiv dv_e = {lv.ne, lv.e_sh, lv.e_inv}, dv_k = {lv.nk, lv.k_sh, lv.k_inv};
    const bool wide = np * lv.ne >= wide_min;    // uniform
    if (!(level1_eliminated && l == 1)) {   // (the assembly already wrote the level-1 nodes in eliminated form: kernels.cuh fuse1)
      DGPMP2_REP(1, l)
      if (wide) bcr_elim_level<D, 1>(nodes, T, np, s, dv_e, off_l, fail);
      else      bcr_elim_level<D, kLPN>(nodes, T, np, s, dv_e, off_l, fail);
      __syncthreads();
    }
    DGPMP2_BCR_STAMP(8 + 2 * l);
    DGPMP2_REP(2, l)
    if (wide) bcr_kept_level<D, 1>(nodes, T, np, s, dv_k, off_l);
    else      bcr_kept_level<D, kLPN>(nodes, T, np, s, dv_k, off_l);
    __syncthreads();
    DGPMP2_BCR_STAMP(9 + 2 * l);
    off_l += lv.ne;
  }

  // ------------------------------ remaining chain (root when tail_max == 1) ------------------------------
  bcr_tail<D>(nodes, T, np, plan.tail_stride, plan.tail_nc, fail);
  side_work((kLPN * np + 31) >> 5);
  __syncthreads();
  DGPMP2_BCR_STAMP(5);

  // ------------------------------ back substitution ------------------------------
  for (int l = nl; l >= 1; --l) {
    const int s = 1 << (l - 1);
    const BcrLevelPlan& lv = plan.lv[l - 1];
    const LevelDiv dv_e = {lv.ne, lv.e_sh, lv.e_inv};
    off_l -= lv.ne;
    DGPMP2_REP(3, l)
    if (np * lv.ne >= wide_min) bcr_back_level<D, 1>(nodes, T, np, s, dv_e, off_l);
    else                        bcr_back_level<D, kLPN>(nodes, T, np, s, dv_e, off_l);
    __syncthreads();
    DGPMP2_BCR_STAMP(40 + l);
  }
}

}  // namespace dgpmp2
