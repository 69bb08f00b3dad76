// Elimination schedule of the block cyclic reduction (bcr.cuh): everything that depends only on the
// trajectory length T is computed once on the host and travels in the kernel arguments (constant
// bank), so the device never divides or searches a level table.  Plain host/device structs.
#pragma once
#include <cuda_runtime.h>

namespace dgpmp2 {

constexpr int kMaxLevels = 16;
constexpr int kLPN = 4;               // lanes per work item on the narrow levels
constexpr int kWideMinDefault = 56;   // work items in the CTA from which a level runs one lane per item
constexpr int kTailMaxDefault = 4;    // elimination stops at this many nodes per problem; the rest is solved sequentially

__host__ __device__ __forceinline__ int bcr_n_elim(int T, int s) { return (T + s - 1) / (2 * s); }   // nodes j = s(2q+1) < T
__host__ __device__ __forceinline__ int bcr_n_kept(int T, int s) { return (T + 2 * s - 1) / (2 * s); } // nodes i = 2sq < T

struct BcrLevelPlan {
  int ne, nk;            // eliminated / kept nodes of the level (per problem)
  int e_sh, k_sh;        // log2 when the count is a power of two, else -1
  float e_inv, k_inv;    // 1.0f / count (see fast_div in bcr.cuh)
};
struct BcrPlan {
  int nl;                // elimination levels run (strides 1, 2, ..., 2^(nl-1))
  int tail_stride;       // 2^nl: stride of the chain solved sequentially afterwards
  int tail_nc;           // nodes of that chain (<= tail_max)
  int wide_min;          // a level with at least this many items in the CTA runs one lane per item
  float inv_T;           // 1.0f / T (fast_div of node indices by the trajectory length)
  BcrLevelPlan lv[kMaxLevels];   // lv[l - 1] for level l
};

inline int bcr_log2_exact(int n) {
  int sh = 0;
  while ((1 << sh) < n) ++sh;
  return ((1 << sh) == n) ? sh : -1;
}

inline void bcr_make_plan(int T, int tail_max, int wide_min, BcrPlan& pl) {
  if (tail_max < 1) tail_max = 1;
  int nl = 0;
  while (((T + (1 << nl) - 1) >> nl) > tail_max) ++nl;
  pl.nl = nl;
  pl.tail_stride = 1 << nl;
  pl.tail_nc = (T + (1 << nl) - 1) >> nl;
  pl.wide_min = wide_min;
  pl.inv_T = 1.0f / (float)T;
  for (int l = 1; l <= kMaxLevels; ++l) {
    BcrLevelPlan& v = pl.lv[l - 1];
    const int s = 1 << (l - 1);
    v.ne = (l <= nl) ? bcr_n_elim(T, s) : 1;
    v.nk = (l <= nl) ? bcr_n_kept(T, s) : 1;
    v.e_sh = bcr_log2_exact(v.ne);
    v.k_sh = bcr_log2_exact(v.nk);
    v.e_inv = 1.0f / (float)v.ne;
    v.k_inv = 1.0f / (float)v.nk;
  }
}

}  // namespace dgpmp2
