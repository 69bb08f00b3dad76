// Global kernels of the dGPMP2 Gauss-Newton path (sm_100a).
//
//   gn_step_kernel    one fused GN iteration: stage th -> assemble band in smem ->
//                     block cyclic reduction -> dth, err, err_ext.  HBM traffic is
//                     th in, dth out, 4 SDF taps per state, 2 scalars per problem.
//   gn_solve_kernel   the same iteration looped to convergence with th resident
//                     in shared memory (DiffGPMP2Planner.forward).
//   errors_kernel     factor sweep only (error_batch / error_ext_batch / unweighted errors).
//   band_kernel       writes the information band (D, U, r) in double (inspection / parity).
//   factors_kernel    stand-alone factor outputs (GPFactor / ObstacleFactor / custom factors).
//   sdf_lookup_kernel bilinear_interpolate.
#pragma once
#include <type_traits>
#ifdef DGPMP2_TIMING
#define DGPMP2_BCR_STAMP(i) do { if (blockIdx.x == (DGPMP2_TIMING - 1) && threadIdx.x == 0) dgpmp2::g_phase_clock_fwd(i); } while (0)
namespace dgpmp2 { __device__ void g_phase_clock_fwd(int i); }
#endif
#include "bcr.cuh"
#include "mp.cuh"

namespace dgpmp2 {

// Shared-memory carve-up of the step / solve kernels: NP problems x T node records (bcr.cuh),
// the staged trajectory, per-node |dth|^2 (solve kernel) and per-problem flags.
template <int D, typename IO>
struct StepSmem {
  double* nodes;    // [NP*T] records of Node<D>::kStride doubles
  double* nrm;      // [NP*T] per-node |dth|^2 (solve kernel only)
  IO* th;           // [NP*T][D] staged trajectory, natural (problem, t, a) order
  IO* dth;          // [NP*T][D] staged forward step (backward kernel only)
  int* fail;        // [NP]
  int* flags;       // [2*NP] solve kernel: converged flag per problem, then iteration count
  // mode: 0 = step, 1 = solve (adds nrm), 2 = backward (adds the dth stage)
  __host__ __device__ static size_t bytes(int NP, int T, int mode) {
    const size_t NN = (size_t)NP * T;
    size_t b = (size_t)NP * Node<D>::problem_stride(T) * 8;
    if (mode == 1) b += ((NN + 1) & ~(size_t)1) * 8;   // keeps the staged trajectory 16-byte aligned
    b += NN * D * sizeof(IO) * (mode == 2 ? 2 : 1);
    b = (b + 15) & ~(size_t)15;
    b += (size_t)NP * 4 * 3 + 16;
    return b;
  }
  __device__ __forceinline__ void carve(unsigned char* raw, int NP, int T, int mode) {
    const size_t NN = (size_t)NP * T;
    nodes = reinterpret_cast<double*>(raw);
    double* nxt = nodes + (size_t)NP * Node<D>::problem_stride(T);
    nrm = nxt;
    if (mode == 1) nxt += (NN + 1) & ~(size_t)1;
    th = reinterpret_cast<IO*>(nxt);
    dth = th + NN * D;
    size_t off = (reinterpret_cast<unsigned char*>(th + NN * D * (mode == 2 ? 2 : 1)) - raw + 15) & ~(size_t)15;
    fail = reinterpret_cast<int*>(raw + off);
    flags = fail + NP;
  }
};

// Assemble the CTA's np * T nodes (one thread per node record, records enumerated problem-major
// in slot order) from the staged trajectory.
//
// fuse1 (uniform; requires P.static_gp and at least one elimination level; gn_step_kernel only): the thread of a level-1 node (odd state j) eliminates it straight from its registers
// -- L_j, F_j = L^-1 U_j, g_j = L^-1 r_j and E_j = L^-1 U_{j-1}^T with U_{j-1} = -Phi^T Q^-1, a host-known constant in
// the static-GP case -- and writes the record once in its eliminated form; bcr_solve then skips the level-1
// elimination phase and its barrier.  Same arithmetic on the same doubles as bcr_elim_level: bit-identical results.
template <int DOF, typename IO>
__device__ __forceinline__ void assemble_cta(const KParams& P, const KWeights<IO>& Wt, const StepSmem<2 * DOF, IO>& S,
                                             int b0, int np,
                                             const IO* __restrict__ start, const IO* __restrict__ goal,
                                             const IO* __restrict__ sdf, bool fuse1 = false, bool prefetch = true,
                                             const IO* __restrict__ rhs = nullptr) {
  constexpr int D = 2 * DOF;
  using N = Node<D>;
  const int T = P.T;
  const float inv_T = P.plan.inv_T;
  prefetch = prefetch && (P.prefetch & 1);
  for (int m = threadIdx.x; m < np * T; m += blockDim.x) {
    const int p = fast_div(m, inv_T), slot = m - p * T;
    const int t = bcr_state_of_slot(T, slot);
    const int b = b0 + p;
    double thp[D], thc[D], thn[D];
    const IO* tp = S.th + ((size_t)p * T + t) * D;
#pragma unroll
    for (int a = 0; a < D; ++a) {
      thc[a] = (double)tp[a];
      thp[a] = (t > 0) ? (double)tp[a - D] : 0.0;
      thn[a] = (t < T - 1) ? (double)tp[a + D] : 0.0;
    }
    if (prefetch) {
      // L2 prefetch of the two SDF rows this state's obstacle factor gathers from, issued before the prior / GP arithmetic
      // so that the DRAM latency of the (DRAM-cold) gather hides behind it.  A hint only: the pixel is located in float
      // arithmetic (the exact, branch-deciding location is computed in double inside assemble_node); no register is
      // held across the arithmetic, unlike an early load.
      const float fx = (float)P.orig_x + (float)thc[0] * (float)P.inv_res;
      const float fy = (float)P.orig_y - (float)thc[1] * (float)P.inv_res;
      const int ix = min(max(__float2int_rd(fx), 0), P.W - 1), iy = min(max(__float2int_rd(fy), 0), P.H - 1);
      const IO* q = sdf + (size_t)b * P.sdf_sb + (size_t)iy * P.W + ix;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(q + ((iy + 1 < P.H) ? P.W : 0)));
    }
    NodeOut<DOF> o;
    assemble_node<DOF, IO>(P, Wt, b, t, thp, thc, thn, start + (size_t)b * D, goal + (size_t)b * D,
                           sdf + (size_t)b * P.sdf_sb, o);
    if (rhs != nullptr) {                          // backward kernel: the system is solved for a given right-hand side
      const IO* gp = rhs + ((size_t)b * T + t) * D;
#pragma unroll
      for (int a = 0; a < D; ++a) o.r[a] = ldg_d(gp + a);
    }
    double* nd = S.nodes + (size_t)p * N::problem_stride(T) + (size_t)slot * N::kStride;
    if (fuse1 && slot < P.plan.lv[0].ne) {          // level-1 node: the slots [0, ne) of level order
      constexpr int DS = N::DS;
      double L[DS];
#pragma unroll
      for (int a = 0; a < D; ++a)
#pragma unroll
        for (int c = 0; c <= a; ++c) L[tri(a, c)] = o.Dm[a][c];
      if (!chol_packed<D>(L)) atomicMax(&S.fail[p], t + 1);
#pragma unroll
      for (int k = 0; k < DS; k += 2) sts2(nd + N::oD + k, L[k], (k + 1 < DS) ? L[k + 1] : 0.0);
      double v[D];
#pragma unroll
      for (int c = 0; c < D; ++c) {                 // F_j = L^-1 U_j, column-major (U_j = 0 for the last state)
#pragma unroll
        for (int a = 0; a < D; ++a) v[a] = o.Um[a][c];
        fwd_solve<D>(L, v);
        st_vec<D>(nd + N::oU + c * D, v);
      }
      fwd_solve<D>(L, o.r);
      st_vec<D>(nd + N::oR, o.r);
#pragma unroll
      for (int c = 0; c < D; ++c) {                 // E_j = L^-1 U_{j-1}^T: column c = row c of U_{j-1} = -Phi^T Q^-1
#pragma unroll
        for (int a = 0; a < D; ++a) v[a] = -P.PQs[c * D + a];
        fwd_solve<D>(L, v);
        st_vec<D>(nd + N::oE + c * D, v);
      }
      sts2(nd + N::oX, o.err, o.err_ext);
      continue;
    }
#pragma unroll
    for (int a = 0; a < D; ++a) {
      st_vec<D>(nd + N::oD + a * D, o.Dm[a]);
      st_vec<D>(nd + N::oU + a * D, o.Um[a]);
    }
    st_vec<D>(nd + N::oR, o.r);
    sts2(nd + N::oX, o.err, o.err_ext);
  }
}

// Deterministic per-problem reduction by warp w for problems w, w+nwarps, ...: sum over the T
// values `base[p * pstride + i * stride]`.  `f(p, sum)` is called by lane 0.
template <typename F>
__device__ __forceinline__ void reduce_per_problem(const double* base, int stride, size_t pstride, int np, int T, F f) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int p = warp; p < np; p += nwarps) {
    double s = 0.0;
    for (int i = lane; i < T; i += 32) s += base[(size_t)p * pstride + (size_t)i * stride];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) f(p, s);
  }
}

// Same for two interleaved quantities stored as adjacent doubles (16-byte aligned): `base[p * pstride + i * stride + {0, 1}]`.
// Each component is summed in exactly the order reduce_per_problem uses.  `f(p, sum0, sum1)` is called by lane 0.
// wbase: the reductions are done by the warps [wbase, nwarps) only (warp wbase + k takes problems k, k + nwarps - wbase, ...)
template <typename F>
__device__ __forceinline__ void reduce2_per_problem(const double* base, int stride, size_t pstride, int np, int T, F f,
                                                    int wbase = 0) {
  const int warp = (threadIdx.x >> 5) - wbase, lane = threadIdx.x & 31, nwarps = (blockDim.x >> 5) - wbase;
  if (warp < 0) return;
  for (int p = warp; p < np; p += nwarps) {
    double s0 = 0.0, s1 = 0.0;
    for (int i = lane; i < T; i += 32) {
      const double2 v = lds2(base + (size_t)p * pstride + (size_t)i * stride);
      s0 += v.x;
      s1 += v.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if (lane == 0) f(p, s0, s1);
  }
}

// one state (D values) to global memory; `vec` (uniform): the array base is 16-byte aligned -> wide stores
template <int D, typename IO>
__device__ __forceinline__ void store_state(IO* __restrict__ dst, const double (&x)[D], bool vec) {
  if (!vec) {
#pragma unroll
    for (int a = 0; a < D; ++a) dst[a] = (IO)x[a];
    return;
  }
  if constexpr (sizeof(IO) == 4 && D % 4 == 0) {
#pragma unroll
    for (int a = 0; a < D; a += 4)
      *reinterpret_cast<float4*>(dst + a) = make_float4((float)x[a], (float)x[a + 1], (float)x[a + 2], (float)x[a + 3]);
  } else if constexpr (sizeof(IO) == 4) {
#pragma unroll
    for (int a = 0; a < D; a += 2) *reinterpret_cast<float2*>(dst + a) = make_float2((float)x[a], (float)x[a + 1]);
  } else {
#pragma unroll
    for (int a = 0; a < D; a += 2) *reinterpret_cast<double2*>(dst + a) = make_double2(x[a], x[a + 1]);
  }
}

template <int D, typename IO>
__device__ __forceinline__ void cta_prologue(const KParams& P, const StepSmem<D, IO>& S, int T, int NP, int np,
                                             const IO* __restrict__ th_src, bool solve) {
  for (int p = threadIdx.x; p < NP; p += blockDim.x) {
    S.fail[p] = 0;
    if (solve) { S.flags[p] = 0; S.flags[NP + p] = 0; }
  }
  const int n = np * T * D;   // contiguous in HBM -> coalesced
  for (int i = threadIdx.x; i < n; i += blockDim.x) S.th[i] = __ldg(th_src + i);
}

#ifdef DGPMP2_TIMING
// Debug builds only (-DDGPMP2_TIMING): per-phase clock64() stamps of CTA 0 / thread 0.
__device__ long long g_phase_clock[64];
__device__ void g_phase_clock_fwd(int i) { g_phase_clock[i] = clock64(); }
#define DGPMP2_STAMP(i) do { if (blockIdx.x == (DGPMP2_TIMING - 1) && threadIdx.x == 0) g_phase_clock[i] = clock64(); } while (0)
#else
#define DGPMP2_STAMP(i) do { } while (0)
#endif

// err / err_ext of the CTA's problems: one deterministic warp reduction per problem over the states' partials (which the
// solver never touches), done by the warps [wbase, nwarps)
template <int DOF, typename IO>
__device__ __forceinline__ void step_errors(const KParams& P, const StepSmem<2 * DOF, IO>& S, int T, int b0, int np,
                                            IO* __restrict__ err, IO* __restrict__ err_ext, int wbase) {
  using N = Node<2 * DOF>;
  const double invM = 1.0 / (double)P.M;
  reduce2_per_problem(S.nodes + N::oX, N::kStride, N::problem_stride(T), np, T, [&](int p, double s0, double s1) {
    err[b0 + p] = (IO)(s0 * invM);
    err_ext[b0 + p] = (IO)(s1 * invM);
  }, wbase);
}

// dth (natural order -> coalesced stores), err / err_ext (one deterministic warp reduction per problem), status
template <int DOF, typename IO>
__device__ __forceinline__ void step_epilogue(const KParams& P, const StepSmem<2 * DOF, IO>& S, int T, int b0, int np,
                                              IO* __restrict__ dth, IO* __restrict__ err, IO* __restrict__ err_ext,
                                              int* __restrict__ status, bool errors_done = false) {
  constexpr int D = 2 * DOF;
  using N = Node<D>;
  {
    IO* dst = dth + (size_t)b0 * T * D;
    const int n = np * T;
    const float inv_T = P.plan.inv_T;
    const bool vec = (reinterpret_cast<unsigned long long>(dst) & 15ull) == 0ull;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int p = fast_div(i, inv_T), t = i - p * T;
      double x[D];
      ld_vec<D>(S.nodes + (size_t)p * N::problem_stride(T) + (size_t)bcr_slot(T, t) * N::kStride + N::oR, x);
      store_state<D, IO>(dst + (size_t)i * D, x, vec);
    }
  }
  if (!errors_done) step_errors<DOF, IO>(P, S, T, b0, np, err, err_ext, 0);
  if (status != nullptr)
    for (int p = threadIdx.x; p < np; p += blockDim.x) status[b0 + p] = S.fail[p];
}

// Early SDF prefetch (measured DRAM-cold, B=1024: 19.16 -> 18.21 us at T=64, 38.0 -> 36.4 us at T=128): every thread reads
// the position of one state straight from global memory (natural order, coalesced; the staging loads that follow then
// hit L1 / L2) and requests the two SDF rows of its obstacle factor before the trajectory is staged, so the DRAM-cold
// gather starts one barrier earlier.  A hint only: the pixel is located in float arithmetic.
template <int D, typename IO>
__device__ __forceinline__ void early_sdf_prefetch(const KParams& P, int T, int b0, int np, const IO* __restrict__ th,
                                                   const IO* __restrict__ sdf) {
  const IO* src = th + (size_t)b0 * T * D;
  const float inv_T = P.plan.inv_T;
  for (int m = threadIdx.x; m < np * T; m += blockDim.x) {
    const int p = fast_div(m, inv_T);
    const float fx = (float)P.orig_x + (float)__ldg(src + (size_t)m * D) * (float)P.inv_res;
    const float fy = (float)P.orig_y - (float)__ldg(src + (size_t)m * D + 1) * (float)P.inv_res;
    const int ix = min(max(__float2int_rd(fx), 0), P.W - 1), iy = min(max(__float2int_rd(fy), 0), P.H - 1);
    const IO* q = sdf + (size_t)(b0 + p) * P.sdf_sb + (size_t)iy * P.W + ix;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(q + ((iy + 1 < P.H) ? P.W : 0)));
  }
}

// Staging of the CTA's trajectories fused with the early SDF prefetch: returns false (nothing done) when the global
// trajectory is not aligned for the vector loads.  V = 16 bytes when a state is a multiple of 16 bytes, else 8.
template <int D, typename IO>
__device__ __forceinline__ bool stage_and_prefetch(const KParams& P, const StepSmem<D, IO>& S, int T, int NP, int np, int b0,
                                                   const IO* __restrict__ th, const IO* __restrict__ sdf) {
  constexpr int SB = D * (int)sizeof(IO);                       // bytes per state
  using V = typename std::conditional<SB % 16 == 0, typename std::conditional<sizeof(IO) == 4, float4, double2>::type,
                                      typename std::conditional<sizeof(IO) == 4, float2, double>::type>::type;
  constexpr int NV = SB / (int)sizeof(V), EV = (int)(sizeof(V) / sizeof(IO));
  if ((reinterpret_cast<unsigned long long>(th) & (sizeof(V) - 1)) != 0ull) return false;   // uniform
  for (int p = threadIdx.x; p < NP; p += blockDim.x) S.fail[p] = 0;
  const V* src = reinterpret_cast<const V*>(th + (size_t)b0 * T * D);
  V* dst = reinterpret_cast<V*>(S.th);
  const float inv_T = P.plan.inv_T;
  for (int m = threadIdx.x; m < np * T; m += blockDim.x) {
    V v[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = __ldg(src + (size_t)m * NV + k);
#pragma unroll
    for (int k = 0; k < NV; ++k) dst[(size_t)m * NV + k] = v[k];
    const IO* e = reinterpret_cast<const IO*>(&v[0]);           // x, y are the first two elements of the state
    static_assert(EV >= 2, "a vector holds at least x and y");
    const float x = (float)e[0], y = (float)e[1];
    const int p = fast_div(m, inv_T);
    const float fx = (float)P.orig_x + x * (float)P.inv_res, fy = (float)P.orig_y - y * (float)P.inv_res;
    const int ix = min(max(__float2int_rd(fx), 0), P.W - 1), iy = min(max(__float2int_rd(fy), 0), P.H - 1);
    const IO* q = sdf + (size_t)(b0 + p) * P.sdf_sb + (size_t)iy * P.W + ix;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(q + ((iy + 1 < P.H) ? P.W : 0)));
  }
  return true;
}

// One fused Gauss-Newton iteration.  grid = ceil(B / NP), block = NP * TPP threads (rounded to a warp).
template <int DOF, typename IO>
__global__ void __launch_bounds__(DOF == 2 ? 512 : 256)
gn_step_kernel(const KParams P, const KWeights<IO> Wt, const IO* __restrict__ th, const IO* __restrict__ start,
               const IO* __restrict__ goal, const IO* __restrict__ sdf, IO* __restrict__ dth,
               IO* __restrict__ err, IO* __restrict__ err_ext, int* __restrict__ status, const int NP,
               const int n_big) {
  constexpr int D = 2 * DOF;
  using N = Node<D>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  StepSmem<D, IO> S;
  const int T = P.T;
  S.carve(smem_raw, NP, T, 0);
  // CTAs [0, n_big) take NP problems each, the later ones NP - 1 (c_abi.cu: choose_shape, balanced waves)
  const int small = max((int)blockIdx.x - n_big, 0);
  const int b0 = blockIdx.x * NP - small;
  const int np = min(NP - ((int)blockIdx.x >= n_big ? 1 : 0), P.B - b0);
  // Programmatic dependent launch (c_abi.cu: launch_step): let the next launch of the stream be scheduled onto the
  // SMs as soon as this grid's CTAs leave them, and wait here -- before the first global access -- until the
  // previous grid of the stream has completed and its writes are visible.  Both are no-ops for a plain launch.
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  DGPMP2_STAMP(0);

  const bool early = (P.prefetch & 2) != 0;
  // One pass (when the trajectory is vector-aligned): every thread loads whole states in natural order, stages them and
  // requests the SDF rows of their obstacle factors -- one global round trip instead of the prefetch pass followed by the
  // staging pass (measured, B=1024, T=64: 16.59 -> 16.17 us L2-warm, 17.96 -> 17.84 us DRAM-cold; same bits).
  if (early && stage_and_prefetch<D, IO>(P, S, T, NP, np, b0, th, sdf)) {
  } else {
    if (early) early_sdf_prefetch<D, IO>(P, T, b0, np, th, sdf);
    cta_prologue<D, IO>(P, S, T, NP, np, th + (size_t)b0 * T * D, false);
  }
  __syncthreads();
  DGPMP2_STAMP(1);

  const bool fuse1 = P.fuse1 != 0;      // uniform (host: static GP blocks and at least one elimination level)
  assemble_cta<DOF, IO>(P, Wt, S, b0, np, start, goal, sdf, fuse1, !early);
  __syncthreads();
  DGPMP2_STAMP(2);

  // the error reductions ride in the warps the sequential tail leaves idle (when there are any)
  const bool side = ((kLPN * np + 31) >> 5) < (int)(blockDim.x >> 5);
  bcr_solve<D>(S.nodes, P.plan, T, np, S.fail, fuse1, [&](int first_idle_warp) {
    if (side) step_errors<DOF, IO>(P, S, T, b0, np, err, err_ext, first_idle_warp);
  });   // ends with a barrier
  DGPMP2_STAMP(3);

  step_epilogue<DOF, IO>(P, S, T, b0, np, dth, err, err_ext, status, side);
  DGPMP2_STAMP(4);
}

// ---------------------------------------------------------------------------
// Mixed-precision GN step (mp.cuh): fp32 node-owner block cyclic reduction + fp64 residual refinement, one thread per
// trajectory state, TPP = ceil32(T) threads per problem, NP problems per CTA, each problem on its own named barrier.
// Problems whose fp32 factorisation breaks down or whose refinement does not contract are redone, inside this
// launch, by the fp64 path of gn_step_kernel (one at a time: rare by construction).
// ---------------------------------------------------------------------------
constexpr int kMpMaxNP = 32;

struct MpDevCtx {
  int bar_id, tpp, lane, wip, nwp;
  double* errw;   // [2 * nwp], this problem's per-warp error partials
  int* norm;      // [4], this problem's double-buffered norm maxima
  int* need64;    // this problem's flag
  __device__ __forceinline__ void psync() const {
    if (tpp == 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(tpp) : "memory");
  }
  __device__ __forceinline__ void sum2(double& a, double& b) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (nwp > 1) {
      if (lane == 0) { errw[2 * wip] = a; errw[2 * wip + 1] = b; }
      psync();
      a = 0.0; b = 0.0;
      for (int w = 0; w < nwp; ++w) { a += errw[2 * w]; b += errw[2 * w + 1]; }
    }
  }
  __device__ __forceinline__ void max2(int& a, int& b, int it) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a = max(a, __shfl_xor_sync(0xffffffffu, a, o));
      b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
    }
    if (nwp == 1) { psync(); return; }
    int* buf = norm + 2 * (it & 1);
    if (lane == 0) { atomicMax(buf, a); atomicMax(buf + 1, b); }
    psync();
    a = buf[0]; b = buf[1];
    if (wip == 0 && lane == 0) { int* o = norm + 2 * ((it + 1) & 1); o[0] = 0; o[1] = 0; }
  }
  __device__ __forceinline__ void flag64() const { *need64 = 1; }
#ifdef DGPMP2_MP_TIMING
  __device__ __forceinline__ void stamp(int i) const;
#endif
};
#ifdef DGPMP2_MP_TIMING
// Debug builds only (-DDGPMP2_MP_TIMING=cta+1): clock64() stamps of the phases of mp_thread_program, thread 0 of one CTA.
__device__ long long g_mp_clock[64];
__device__ __forceinline__ void MpDevCtx::stamp(int i) const {
  if (blockIdx.x == (DGPMP2_MP_TIMING - 1) && threadIdx.x == 0 && i < 64) g_mp_clock[i] = clock64();
}
#endif

template <int DOF, int MAXT>
__global__ void __launch_bounds__(MAXT)
gn_step_mp_kernel(const KParams P, const KWeights<float> Wt, const float* __restrict__ th, const float* __restrict__ start,
                  const float* __restrict__ goal, const float* __restrict__ sdf, float* __restrict__ dth,
                  float* __restrict__ err, float* __restrict__ err_ext, int* __restrict__ status, int* __restrict__ diag,
                  const int NP, const int TPP, const int force64) {
  constexpr int D = 2 * DOF;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_need64[kMpMaxNP];
  __shared__ int s_norm[kMpMaxNP][4];
  __shared__ double s_errw[32][2];
  const int T = P.T;
  const int p = threadIdx.x / TPP, m = threadIdx.x - p * TPP;
  const int b0 = blockIdx.x * NP;
  const int np = min(NP, P.B - b0);
  if (threadIdx.x < kMpMaxNP) {
    s_need64[threadIdx.x] = 0;
    s_norm[threadIdx.x][0] = 0; s_norm[threadIdx.x][1] = 0; s_norm[threadIdx.x][2] = 0; s_norm[threadIdx.x][3] = 0;
  }
  __syncthreads();
  if (p < np) {
    MpDevCtx cx;
    cx.bar_id = 1 + p; cx.tpp = TPP; cx.lane = threadIdx.x & 31; cx.wip = m >> 5; cx.nwp = TPP >> 5;
    cx.errw = &s_errw[(p * TPP) >> 5][0];
    cx.norm = s_norm[p];
    cx.need64 = &s_need64[p];
    float* recs = reinterpret_cast<float*>(smem_raw) + (size_t)p * MpRec<D>::problem_floats(T);
    // Which of the problem's warps holds which 32 slots rotates with the problem index: the LAST slots are the deep
    // levels -- the warp that works through most of the solve -- and warp w of a CTA issues on SM sub-partition w % 4,
    // so without the rotation the deep warps of all problems pile up on one or two of the four schedulers.
    const int rot = (cx.nwp == 2) ? ((p >> 1) & 1) : (p % cx.nwp);
    int ms = m + 32 * rot;
    if (ms >= TPP) ms -= TPP;
    cx.wip = ms >> 5;   // error partials are summed in slot order, wherever the problem sits in the CTA
    mp_thread_program<DOF, (DOF == 2), MpDevCtx>(cx, P, Wt, b0 + p, ms, ms < T, th, start, goal, sdf, recs, dth, err, err_ext,
                                                 diag, force64);
  }
  __syncthreads();
  if (status != nullptr && threadIdx.x < np && !s_need64[threadIdx.x]) status[b0 + threadIdx.x] = 0;
  // fp64 path for the flagged problems (uniform: every thread reads the same flags)
  for (int q = 0; q < np; ++q) {
    if (!s_need64[q]) continue;
    StepSmem<D, float> S;
    S.carve(smem_raw, 1, T, 0);
    const int b = b0 + q;
    cta_prologue<D, float>(P, S, T, 1, 1, th + (size_t)b * T * D, false);
    __syncthreads();
    assemble_cta<DOF, float>(P, Wt, S, b, 1, start, goal, sdf);
    __syncthreads();
    bcr_solve<D>(S.nodes, P.plan, T, 1, S.fail);   // ends with a barrier
    step_epilogue<DOF, float>(P, S, T, b, 1, dth, err, err_ext, status);
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// Persistent solve-to-convergence (DiffGPMP2Planner.forward, diff_gpmp2_planner.py:104-165)
// ---------------------------------------------------------------------------
template <int DOF, typename IO>
__global__ void __launch_bounds__(DOF == 2 ? 512 : 256)
gn_solve_kernel(const KParams P, const KWeights<IO> Wt, const IO* __restrict__ th_init, const IO* __restrict__ start,
                const IO* __restrict__ goal, const IO* __restrict__ sdf, const int max_iters, const double tol_delta,
                IO* __restrict__ th_final, int* __restrict__ iters, IO* __restrict__ err_pi, IO* __restrict__ err_ext_pi,
                IO* __restrict__ err_final, IO* __restrict__ err_ext_final, int* __restrict__ status, const int NP) {
  constexpr int D = 2 * DOF;
  using N = Node<D>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  StepSmem<D, IO> S;
  const int T = P.T;
  S.carve(smem_raw, NP, T, 1);
  int* done = S.flags;          // [NP]
  int* nit = S.flags + NP;      // [NP]
  const int b0 = blockIdx.x * NP;
  const int np = min(NP, P.B - b0);

  cta_prologue<D, IO>(P, S, T, NP, np, th_init + (size_t)b0 * T * D, true);
  __syncthreads();
  const double invM = 1.0 / (double)P.M;
  const bool fuse1 = P.fuse1 != 0;      // uniform (host: static GP blocks and at least one elimination level)

  for (int j = 0;; ++j) {
    // assemble at the current iterate (level-1 elimination fused in, as in gn_step_kernel); the errors at iterate j
    // are a by-product
    assemble_cta<DOF, IO>(P, Wt, S, b0, np, start, goal, sdf, fuse1);
    __syncthreads();
    const bool last = (j >= max_iters);
    reduce2_per_problem(S.nodes + N::oX, N::kStride, N::problem_stride(T), np, T, [&](int p, double s0, double s1) {
      if (!done[p] && !last) {
        if (err_pi != nullptr) err_pi[(size_t)(b0 + p) * max_iters + j] = (IO)(s0 * invM);
        if (err_ext_pi != nullptr) err_ext_pi[(size_t)(b0 + p) * max_iters + j] = (IO)(s1 * invM);
      }
      if (done[p] == 1 || last) {
        if (err_final != nullptr) err_final[b0 + p] = (IO)(s0 * invM);
        if (err_ext_final != nullptr) err_ext_final[b0 + p] = (IO)(s1 * invM);
      }
    });
    __syncthreads();
    // problems that converged at the previous iteration have now had their final error recorded
    for (int p = threadIdx.x; p < np; p += blockDim.x)
      if (done[p] == 1) done[p] = 2;
    __syncthreads();
    bool all_done = true;
    for (int p = 0; p < np; ++p) all_done = all_done && (done[p] == 2);
    if (all_done || last) break;

    bcr_solve<D>(S.nodes, P.plan, T, np, S.fail, fuse1);

    // th <- th + dth for problems still running; |dth|^2 partials
    for (int m = threadIdx.x; m < np * T; m += blockDim.x) {
      const int p = fast_div(m, P.plan.inv_T), t = m - p * T;
      double s2 = 0.0;
      if (!done[p]) {
        double x[D];
        ld_vec<D>(S.nodes + (size_t)p * N::problem_stride(T) + (size_t)bcr_slot(T, t) * N::kStride + N::oR, x);
#pragma unroll
        for (int a = 0; a < D; ++a) {
          // the reference adds dtheta (I/O dtype) to th (I/O dtype): round dth first, then add
          const IO dxi = (IO)x[a];
          s2 += (double)dxi * (double)dxi;
          IO* q = S.th + ((size_t)p * T + t) * D + a;
          *q = (IO)(*q + dxi);
        }
      }
      S.nrm[m] = s2;
    }
    __syncthreads();
    reduce_per_problem(S.nrm, 1, (size_t)T, np, T, [&](int p, double s) {
      if (!done[p]) {
        nit[p] = j + 1;
        // check_convergence (planner_utils.py:3-16): ||dtheta|| < tol_delta  or  j+1 >= max_iters
        if (sqrt(s) < tol_delta || (j + 1) >= max_iters) done[p] = 1;
      }
    });
    __syncthreads();
  }

  {
    IO* dst = th_final + (size_t)b0 * T * D;
    const int n = np * T * D;
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = S.th[i];
  }
  for (int p = threadIdx.x; p < np; p += blockDim.x) {
    iters[b0 + p] = nit[p];
    if (status != nullptr) status[b0 + p] = S.fail[p];
  }
}

// ---------------------------------------------------------------------------
// Backward of one GN step: lambda = Lambda^-1 gbar with the same assembly + BCR, then the factor VJPs.
// Replaces autograd through the reference's dense solve (plan_layer.py:214-234).
// ---------------------------------------------------------------------------
template <int DOF, typename IO>
__global__ void __launch_bounds__(DOF == 2 ? 512 : 256)
gn_step_bwd_kernel(const KParams P, const KWeights<IO> Wt, const IO* __restrict__ th, const IO* __restrict__ start,
                   const IO* __restrict__ goal, const IO* __restrict__ sdf, const IO* __restrict__ dth,
                   const IO* __restrict__ g_dth, const IO* __restrict__ g_err_ext,
                   IO* __restrict__ g_th, IO* __restrict__ g_start, IO* __restrict__ g_goal, IO* __restrict__ g_qc,
                   IO* __restrict__ g_w, IO* __restrict__ g_eps, IO* __restrict__ g_sdf, const long long g_sdf_sb,
                   const int NP) {
  constexpr int D = 2 * DOF;
  using N = Node<D>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  StepSmem<D, IO> S;
  const int T = P.T;
  S.carve(smem_raw, NP, T, 2);
  const int b0 = blockIdx.x * NP;
  const int np = min(NP, P.B - b0);

  const bool early = (P.prefetch & 2) != 0;
  if (early && stage_and_prefetch<D, IO>(P, S, T, NP, np, b0, th, sdf)) {
  } else {
    if (early) early_sdf_prefetch<D, IO>(P, T, b0, np, th, sdf);
    cta_prologue<D, IO>(P, S, T, NP, np, th + (size_t)b0 * T * D, false);
  }
  {
    const IO* src = dth + (size_t)b0 * T * D;
    const int n = np * T * D;
    for (int i = threadIdx.x; i < n; i += blockDim.x) S.dth[i] = __ldg(src + i);
  }
  __syncthreads();

  // the band of the forward step with gbar as its right-hand side (static-GP blocks and the level-1 elimination fused
  // into the assembly exactly as in gn_step_kernel; backward_node below evaluates Q^-1 itself)
  const bool fuse1 = P.fuse1 != 0;
  assemble_cta<DOF, IO>(P, Wt, S, b0, np, start, goal, sdf, fuse1, !early, g_dth);
  __syncthreads();

  bcr_solve<D>(S.nodes, P.plan, T, np, S.fail, fuse1);   // lambda in every record's [oR, oR+D)

  const double invM = 1.0 / (double)P.M;
  const float inv_T = P.plan.inv_T;
  const int blk = (P.flags & FLAG_Q_FULL) ? D : DOF;
  for (int m = threadIdx.x; m < np * T; m += blockDim.x) {
    const int p = fast_div(m, inv_T), t = m - p * T;
    const int b = b0 + p;
    double thp[D], thc[D], thn[D], lp[D], lc[D], ln[D], dp[D], dc[D], dn[D];
    const IO* tp = S.th + ((size_t)p * T + t) * D;
    const IO* xp = S.dth + ((size_t)p * T + t) * D;
    const double* nb = S.nodes + (size_t)p * N::problem_stride(T);
    ld_vec<D>(nb + (size_t)bcr_slot(T, t) * N::kStride + N::oR, lc);
    if (t > 0) ld_vec<D>(nb + (size_t)bcr_slot(T, t - 1) * N::kStride + N::oR, lp);
    if (t < T - 1) ld_vec<D>(nb + (size_t)bcr_slot(T, t + 1) * N::kStride + N::oR, ln);
#pragma unroll
    for (int a = 0; a < D; ++a) {
      thc[a] = (double)tp[a]; dc[a] = (double)xp[a];
      thp[a] = (t > 0) ? (double)tp[a - D] : 0.0;      dp[a] = (t > 0) ? (double)xp[a - D] : 0.0;
      thn[a] = (t < T - 1) ? (double)tp[a + D] : 0.0;  dn[a] = (t < T - 1) ? (double)xp[a + D] : 0.0;
      if (t == 0) lp[a] = 0.0;
      if (t == T - 1) ln[a] = 0.0;
    }
    const double ghat = (g_err_ext != nullptr) ? ldg_d(g_err_ext + b) * invM : 0.0;
    NodeGrad<DOF> o;
    backward_node<DOF, IO>(P, Wt, b, t, thp, thc, thn, lp, lc, ln, dp, dc, dn, start + (size_t)b * D, goal + (size_t)b * D,
                           sdf + (size_t)b * P.sdf_sb, ghat, (g_sdf != nullptr) ? g_sdf + (size_t)b * g_sdf_sb : nullptr, o);
    if (g_th != nullptr) {
#pragma unroll
      for (int a = 0; a < D; ++a) g_th[((size_t)b * T + t) * D + a] = (IO)o.g_th[a];
    }
    if (t == 0 && g_start != nullptr) {
#pragma unroll
      for (int a = 0; a < D; ++a) g_start[(size_t)b * D + a] = (IO)o.g_prior[a];
    }
    if (t == T - 1 && g_goal != nullptr) {
#pragma unroll
      for (int a = 0; a < D; ++a) g_goal[(size_t)b * D + a] = (IO)o.g_prior[a];
    }
    if (g_w != nullptr) g_w[(size_t)b * T + t] = (IO)o.g_w;
    if (g_eps != nullptr) g_eps[(size_t)b * T + t] = (IO)o.g_eps;
    if (g_qc != nullptr && t < T - 1) {
      IO* q = g_qc + ((size_t)b * (T - 1) + t) * blk * blk;
      if (P.flags & FLAG_Q_FULL) {
#pragma unroll
        for (int a = 0; a < D; ++a)
#pragma unroll
          for (int c = 0; c < D; ++c) q[a * D + c] = (IO)o.g_q[a][c];
      } else {
        // Q = [[qa C, qb C],[qb C, qc C]]  =>  dL/dC = qa G11 + qb (G12 + G21) + qc G22
#pragma unroll
        for (int a = 0; a < DOF; ++a)
#pragma unroll
          for (int c = 0; c < DOF; ++c)
            q[a * DOF + c] = (IO)(P.qa * o.g_q[a][c] + P.qb * (o.g_q[a][c + DOF] + o.g_q[a + DOF][c]) + P.qc * o.g_q[a + DOF][c + DOF]);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Factor sweep only: one CTA per problem.
// ---------------------------------------------------------------------------
template <int DOF, typename IO>
__device__ __forceinline__ void load_state3(const IO* __restrict__ th_b, int T, int t, double (&thp)[2 * DOF],
                                            double (&thc)[2 * DOF], double (&thn)[2 * DOF]) {
  constexpr int D = 2 * DOF;
  const IO* tp = th_b + (size_t)t * D;
#pragma unroll
  for (int a = 0; a < D; ++a) {
    thc[a] = ldg_d(tp + a);
    thp[a] = (t > 0) ? ldg_d(tp + a - D) : 0.0;
    thn[a] = (t < T - 1) ? ldg_d(tp + a + D) : 0.0;
  }
}

template <int DOF, typename IO>
__global__ void __launch_bounds__(256)
errors_kernel(const KParams P, const KWeights<IO> Wt, const IO* __restrict__ th, const IO* __restrict__ start,
              const IO* __restrict__ goal, const IO* __restrict__ sdf, IO* __restrict__ err, IO* __restrict__ err_ext,
              IO* __restrict__ err_sg, IO* __restrict__ err_gp, IO* __restrict__ err_obs) {
  constexpr int D = 2 * DOF;
  __shared__ double part[5][8];
  const int b = blockIdx.x, T = P.T;
  double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    double thp[D], thc[D], thn[D];
    load_state3<DOF, IO>(th + (size_t)b * T * D, T, t, thp, thc, thn);
    NodeOut<DOF> o;
    assemble_node<DOF, IO>(P, Wt, b, t, thp, thc, thn, start + (size_t)b * D, goal + (size_t)b * D,
                           sdf + (size_t)b * P.sdf_sb, o);
    acc[0] += o.err; acc[1] += o.err_ext; acc[2] += o.e_sg; acc[3] += o.e_gp; acc[4] += o.e_obs;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    double s = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) part[k][warp] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s[5];
    for (int k = 0; k < 5; ++k) {
      s[k] = 0.0;
      for (int w = 0; w < nwarps; ++w) s[k] += part[k][w];
    }
    if (err) err[b] = (IO)(s[0] / (double)P.M);
    if (err_ext) err_ext[b] = (IO)(s[1] / (double)P.M);
    if (err_sg) err_sg[b] = (IO)s[2];                       // plan_layer.py:384-388 (mean over a size-1 dim)
    if (err_gp) err_gp[b] = (IO)(s[3] / (double)(T - 1));   // :374-377 mean over GP factors
    if (err_obs) err_obs[b] = (IO)(s[4] / (double)T);       // :379-382 mean over states
  }
}

// ---------------------------------------------------------------------------
// Backward of the factor sweep: d(err_ext, err_sg, err_gp, err_obs) / d th for given upstream gradients (one scalar
// per problem and error; NULL = not requested).  The reference gets these from autograd through error_ext_batch /
// gp_error / obs_error / start_goal_error (plan_layer.py:310-388) -- its training loss is built from them
// (learning/train_planner.py:327-346).  err itself is computed under no_grad there (:275) and has no gradient.
// One thread per state; every factor touching state t contributes (priors, GP factors t-1 and t, obstacle, custom).
// ---------------------------------------------------------------------------
template <int DOF, typename IO>
__global__ void __launch_bounds__(256)
errors_bwd_kernel(const KParams P, const KWeights<IO> Wt, const IO* __restrict__ th, const IO* __restrict__ start,
                  const IO* __restrict__ goal, const IO* __restrict__ sdf, const IO* __restrict__ g_ext,
                  const IO* __restrict__ g_sg, const IO* __restrict__ g_gp, const IO* __restrict__ g_obs,
                  IO* __restrict__ g_th) {
  constexpr int D = 2 * DOF;
  const int T = P.T;
  const long long n = (long long)P.B * T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / T), t = (int)(i - (long long)b * T);
    double thp[D], thc[D], thn[D], g[D];
    load_state3<DOF, IO>(th + (size_t)b * T * D, T, t, thp, thc, thn);
#pragma unroll
    for (int a = 0; a < D; ++a) g[a] = 0.0;
    const double invM = 1.0 / (double)P.M;
    const double ge = (g_ext != nullptr) ? ldg_d(g_ext + b) * invM : 0.0;     // err_ext = sum / M
    const double gs = (g_sg != nullptr) ? ldg_d(g_sg + b) : 0.0;
    const double gg = (g_gp != nullptr) ? ldg_d(g_gp + b) / (double)(T - 1) : 0.0;
    const double go = (g_obs != nullptr) ? ldg_d(g_obs + b) / (double)T : 0.0;
    // priors: e = mean - th  ->  d(0.5 k |e|^2) / d th = -k e
    if (t == 0 || t == T - 1) {
      const double k = (t == 0) ? P.ks : P.kg;
      const IO* mean = ((t == 0) ? start : goal) + (size_t)b * D;
#pragma unroll
      for (int a = 0; a < D; ++a) {
        const double e = ldg_d(mean + a) - thc[a];
        g[a] -= (ge * k + gs) * e;
      }
    }
    // GP factors: g_i = th_{i+1} - Phi th_i ; err_ext uses the constructor-time Q^-1, err_gp the identity
    double Qf[D][D];
    fixed_qinv<DOF>(P, Qf);
    if (t < T - 1) {          // factor t: d g_t / d th_t = -Phi
      double r[D], u[D];
      gp_residual<DOF>(thc, thn, P.dt, r);
#pragma unroll
      for (int a = 0; a < D; ++a) {
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < D; ++c) s += 0.5 * (Qf[a][c] + Qf[c][a]) * r[c];
        u[a] = ge * s + gg * r[a];
      }
#pragma unroll
      for (int a = 0; a < DOF; ++a) {
        g[a] -= u[a];
        g[a + DOF] -= P.dt * u[a] + u[a + DOF];
      }
    }
    if (t > 0) {              // factor t-1: d g_{t-1} / d th_t = I
      double r[D];
      gp_residual<DOF>(thp, thc, P.dt, r);
#pragma unroll
      for (int a = 0; a < D; ++a) {
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < D; ++c) s += 0.5 * (Qf[a][c] + Qf[c][a]) * r[c];
        g[a] += ge * s + gg * r[a];
      }
    }
    // obstacle: c = eps_tot - dist when active, d c / d (x, y) = -grad dist = -(hx, hy)
    {
      const double eps = load_state_weight<IO>(P, Wt.eps, Wt.e_sb, Wt.e_st, b, t, P.eps_const);
      const SdfSample sm = sdf_bilinear<IO, false>(sdf + (size_t)b * P.sdf_sb, P.H, P.W, P.orig_x, P.orig_y, P.res, thc[0],
                                                   thc[1], P.inv_res);
      const ObsTerm ob = hinge(sm, __dadd_rn(eps, P.r_sphere));
      const double k = (ge * P.w_fix + go) * ob.c;
      g[0] -= k * ob.hx;
      g[1] -= k * ob.hy;
    }
    if constexpr (DOF == 3) {
      if (P.flags & FLAG_NONHOLONOMIC) {       // e = vy cos h - vx sin h (weighted error only)
        double sh, ch;
        sincos(thc[2], &sh, &ch);
        const double e = thc[4] * ch - thc[3] * sh, k = ge * P.kd * e;
        g[2] += k * (-thc[4] * sh - thc[3] * ch);
        g[3] += k * (-sh);
        g[4] += k * ch;
      }
    }
    if constexpr (DOF == 2) {
      if (P.flags & FLAG_VEL_LIMITS) {          // c = |v| - limit when |v| >= limit
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const double v = thc[2 + q], lim = (q == 0) ? P.vx_lim : P.vy_lim;
          if (fabs(v) >= lim) g[2 + q] += ge * P.kv * (fabs(v) - lim) * (double)((v > 0.0) - (v < 0.0));
        }
      }
    }
#pragma unroll
    for (int a = 0; a < D; ++a) g_th[(size_t)i * D + a] = (IO)g[a];
  }
}

// ---------------------------------------------------------------------------
// Information band to HBM in double (inspection / parity).
// ---------------------------------------------------------------------------
template <int DOF, typename IO>
__global__ void __launch_bounds__(128)
band_kernel(const KParams P, const KWeights<IO> Wt, const IO* __restrict__ th, const IO* __restrict__ start,
            const IO* __restrict__ goal, const IO* __restrict__ sdf, double* __restrict__ Dg, double* __restrict__ Ug,
            double* __restrict__ rg) {
  constexpr int D = 2 * DOF;
  const int T = P.T;
  const long long n = (long long)P.B * T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / T), t = (int)(i - (long long)b * T);
    double thp[D], thc[D], thn[D];
    load_state3<DOF, IO>(th + (size_t)b * T * D, T, t, thp, thc, thn);
    NodeOut<DOF> o;
    assemble_node<DOF, IO>(P, Wt, b, t, thp, thc, thn, start + (size_t)b * D, goal + (size_t)b * D,
                           sdf + (size_t)b * P.sdf_sb, o);
#pragma unroll
    for (int a = 0; a < D; ++a) {
#pragma unroll
      for (int c = 0; c < D; ++c) {
        Dg[(size_t)i * D * D + a * D + c] = o.Dm[a][c];
        if (t < T - 1) Ug[((size_t)b * (T - 1) + t) * D * D + a * D + c] = o.Um[a][c];
      }
      rg[(size_t)i * D + a] = o.r[a];
    }
  }
}

// ---------------------------------------------------------------------------
// Stand-alone factor outputs.
// ---------------------------------------------------------------------------
template <int DOF, typename IO>
__global__ void __launch_bounds__(256)
factors_kernel(const KParams P, const KWeights<IO> Wt, const IO* __restrict__ th, const IO* __restrict__ sdf,
               IO* __restrict__ gp_err, IO* __restrict__ obs_cost, IO* __restrict__ obs_H, IO* __restrict__ cust_err,
               IO* __restrict__ cust_H) {
  constexpr int D = 2 * DOF;
  const int T = P.T;
  const long long n = (long long)P.B * T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / T), t = (int)(i - (long long)b * T);
    double thp[D], thc[D], thn[D];
    load_state3<DOF, IO>(th + (size_t)b * T * D, T, t, thp, thc, thn);
    if (gp_err != nullptr && t < T - 1) {
      double g[D];
      gp_residual<DOF>(thc, thn, P.dt, g);
#pragma unroll
      for (int a = 0; a < D; ++a) gp_err[((size_t)b * (T - 1) + t) * D + a] = (IO)g[a];
    }
    if (obs_cost != nullptr || obs_H != nullptr) {
      const double eps = load_state_weight<IO>(P, Wt.eps, Wt.e_sb, Wt.e_st, b, t, P.eps_const);
      const SdfSample s = sdf_bilinear<IO>(sdf + (size_t)b * P.sdf_sb, P.H, P.W, P.orig_x, P.orig_y, P.res, thc[0], thc[1]);
      const ObsTerm ob = hinge(s, __dadd_rn(eps, P.r_sphere));
      if (obs_cost) obs_cost[i] = (IO)ob.c;
      if (obs_H) {
        obs_H[(size_t)i * D + 0] = (IO)ob.hx;
        obs_H[(size_t)i * D + 1] = (IO)ob.hy;
#pragma unroll
        for (int a = 2; a < D; ++a) obs_H[(size_t)i * D + a] = (IO)0;
      }
    }
    if constexpr (DOF == 3) {
      if ((P.flags & FLAG_NONHOLONOMIC) && (cust_err != nullptr || cust_H != nullptr)) {
        double sh, ch;
        sincos(thc[2], &sh, &ch);
        if (cust_err) cust_err[i] = (IO)(thc[4] * ch - thc[3] * sh);
        if (cust_H) {
          IO* h = cust_H + (size_t)i * D;
          h[0] = (IO)0; h[1] = (IO)0; h[2] = (IO)(-thc[4] * sh + thc[3] * ch);
          h[3] = (IO)(-sh); h[4] = (IO)ch; h[5] = (IO)0;
        }
      }
    }
    if constexpr (DOF == 2) {
      if ((P.flags & FLAG_VEL_LIMITS) && (cust_err != nullptr || cust_H != nullptr)) {
        const double vx = thc[2], vy = thc[3];
        const bool ax = fabs(vx) >= P.vx_lim, ay = fabs(vy) >= P.vy_lim;
        if (cust_err) {
          cust_err[(size_t)i * 2 + 0] = (IO)(ax ? fabs(vx) - P.vx_lim : 0.0);
          cust_err[(size_t)i * 2 + 1] = (IO)(ay ? fabs(vy) - P.vy_lim : 0.0);
        }
        if (cust_H) {
          IO* h = cust_H + (size_t)i * 2 * D;
#pragma unroll
          for (int a = 0; a < 2 * D; ++a) h[a] = (IO)0;
          h[2] = (IO)(ax ? -(double)((vx > 0.0) - (vx < 0.0)) : 0.0);
          h[D + 3] = (IO)(ay ? -(double)((vy > 0.0) - (vy < 0.0)) : 0.0);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Obstacle factor alone (the streaming fast path of factors_kernel): fused SDF bilinear lookup +
// hinge + Jacobian, one thread per state.  Reads only the 2-D position of each state (one vector
// load), the four SDF taps through the read-only path, and writes cost and the d-wide Jacobian row
// with vector stores.  HBM-bound at >= 1e6 states per launch (profiles/README.md).
// ---------------------------------------------------------------------------
template <typename IO> struct Vec2;
template <> struct Vec2<float> { using type = float2; };
template <> struct Vec2<double> { using type = double2; };

template <int DOF, typename IO>
__global__ void __launch_bounds__(256)
obstacle_kernel(const KParams P, const KWeights<IO> Wt, const IO* __restrict__ th, const IO* __restrict__ sdf,
                IO* __restrict__ obs_cost, IO* __restrict__ obs_H) {
  constexpr int D = 2 * DOF;
  using V2 = typename Vec2<IO>::type;
  const int T = P.T;
  const long long n = (long long)P.B * T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / T), t = (int)(i - (long long)b * T);
    const V2 pos = __ldg(reinterpret_cast<const V2*>(th + (size_t)i * D));
    const double eps = load_state_weight<IO>(P, Wt.eps, Wt.e_sb, Wt.e_st, b, t, P.eps_const);
    const SdfSample s = sdf_bilinear<IO, false>(sdf + (size_t)b * P.sdf_sb, P.H, P.W, P.orig_x, P.orig_y, P.res,
                                                (double)pos.x, (double)pos.y, P.inv_res);
    const ObsTerm ob = hinge(s, __dadd_rn(eps, P.r_sphere));
    if (obs_cost != nullptr) obs_cost[i] = (IO)ob.c;
    if (obs_H != nullptr) {
      V2* h = reinterpret_cast<V2*>(obs_H + (size_t)i * D);
      V2 h0; h0.x = (IO)ob.hx; h0.y = (IO)ob.hy;
      V2 z; z.x = (IO)0; z.y = (IO)0;
      h[0] = h0;
#pragma unroll
      for (int k = 1; k < DOF; ++k) h[k] = z;
    }
  }
}

// ---------------------------------------------------------------------------
// HingeLossObstacleCost.hinge_loss_signed_batch (obstacle_cost.py:29-38): sphere centres in, hinge cost and its 2-wide
// gradient out -- the fused SDF bilinear lookup + obstacle cost + Jacobian in its leanest form: per point 8 B of
// position, four 4-byte taps, 4 + 8 B of output = the 36 algorithmic bytes of SURVEY 8(d).  Two consecutive points per
// thread: one 16-byte load of both positions, one 8-byte store of both costs, one 16-byte store of both gradients
// (fp32 I/O), eight independent tap loads in flight.  HBM-bound; what it can reach is set by the DRAM fetch granularity,
// which is a whole 128-byte line per miss on B200 whatever the load flavour (scratch/ubench8.cu): a 2 x 2 tap patch
// touches 2 image rows = 2 lines, shared with the neighbouring states of the same trajectory only.
// ---------------------------------------------------------------------------
#ifndef DGPMP2_K1_MINBLOCKS
#define DGPMP2_K1_MINBLOCKS 6     // <= 40 registers: 6 CTAs = 48 warps per SM (measured 50.3 -> 46.9 us against 44 registers / 5 CTAs)
#endif
#ifndef DGPMP2_K1_PFDIST
#define DGPMP2_K1_PFDIST 1
#endif
#ifndef DGPMP2_K1_NPT
#define DGPMP2_K1_NPT 2           // consecutive points per thread (even; 4 measured slower: 53-55 us)
#endif
template <typename IO>
__global__ void __launch_bounds__(256, DGPMP2_K1_MINBLOCKS)
hinge_kernel(const IO* __restrict__ sdf, int B, int H, int W, long long sdf_sb, const IO* __restrict__ pts, int N,
             double res, double inv_res, double orig_x, double orig_y, const IO* __restrict__ eps, long long e_sb,
             long long e_sn, double eps_const, double r_sphere, IO* __restrict__ cost, IO* __restrict__ He) {
  using V2 = typename Vec2<IO>::type;
  constexpr int NPT = DGPMP2_K1_NPT;
  const long long n = (long long)B * N;
  const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * NPT;
  if (i0 >= n) return;
  const int cnt = (int)((n - i0 < NPT) ? (n - i0) : NPT);
#if DGPMP2_K1_PFDIST > 0
  {  // L2 prefetch of the positions a CTA DGPMP2_K1_PFDIST waves of resident CTAs later will read (one request per 128-byte
     // line): the first of a thread's two dependent DRAM round trips then hits L2 (measured 47.0 -> 45.0 us; prefetching the
     // tap rows of later CTAs the same way was slower, 51-61 us)
    const long long ip = i0 + (long long)DGPMP2_K1_PFDIST * 148 * DGPMP2_K1_MINBLOCKS * 256 * NPT;
    if (ip < n && (threadIdx.x & (128 / (NPT * 2 * (int)sizeof(IO)) - 1)) == 0)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const V2*>(pts) + ip));
  }
#endif
  V2 p[NPT];
#pragma unroll
  for (int k = 0; k < NPT; ++k) p[k] = __ldg(reinterpret_cast<const V2*>(pts) + i0 + ((k < cnt) ? k : 0));
  ObsTerm ob[NPT];
#pragma unroll
  for (int k = 0; k < NPT; ++k) {
    const long long i = i0 + ((k < cnt) ? k : 0);
    // the point's trajectory: 32-bit division when the point count allows it
    const int b = (n <= 0xffffffffLL) ? (int)((unsigned)i / (unsigned)N) : (int)(i / N);
    const int j = (int)(i - (long long)b * N);
    const double e = (eps != nullptr) ? ldg_d(eps + (long long)b * e_sb + (long long)j * e_sn) : eps_const;
    const SdfSample sm = sdf_bilinear<IO, false>(sdf + (size_t)b * sdf_sb, H, W, orig_x, orig_y, res, (double)p[k].x,
                                                 (double)p[k].y, inv_res);
    ob[k] = hinge(sm, __dadd_rn(e, r_sphere));
  }
  if (cnt == NPT && ((reinterpret_cast<unsigned long long>(cost) | (reinterpret_cast<unsigned long long>(He) >> 1)) & (2 * sizeof(IO) - 1)) == 0) {
#pragma unroll
    for (int k = 0; k < NPT; k += 2) {
      V2 c; c.x = (IO)ob[k].c; c.y = (IO)ob[k + 1].c;
      *reinterpret_cast<V2*>(cost + i0 + k) = c;             // i0 + k is even: 2-element aligned when the base is
      V2 h0, h1; h0.x = (IO)ob[k].hx; h0.y = (IO)ob[k].hy; h1.x = (IO)ob[k + 1].hx; h1.y = (IO)ob[k + 1].hy;
      reinterpret_cast<V2*>(He)[i0 + k] = h0;
      reinterpret_cast<V2*>(He)[i0 + k + 1] = h1;
    }
  } else {
#pragma unroll
    for (int k = 0; k < NPT; ++k) {      // (unrolled: a dynamic index would put ob[] in local memory)
      if (k < cnt) {
        cost[i0 + k] = (IO)ob[k].c;
        He[2 * (i0 + k)] = (IO)ob[k].hx;
        He[2 * (i0 + k) + 1] = (IO)ob[k].hy;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// bilinear_interpolate (utils/sdf_utils.py:38-107)
// ---------------------------------------------------------------------------
template <typename IO>
__global__ void __launch_bounds__(256)
sdf_lookup_kernel(const IO* __restrict__ sdf, int B, int H, int W, long long sdf_sb, const IO* __restrict__ pts, int N,
                  double res, double orig_x, double orig_y, IO* __restrict__ dist, IO* __restrict__ J) {
  const long long n = (long long)B * N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / N);
    const double x = ldg_d(pts + 2 * i), y = ldg_d(pts + 2 * i + 1);
    const SdfSample s = sdf_bilinear<IO>(sdf + (size_t)b * sdf_sb, H, W, orig_x, orig_y, res, x, y);
    if (dist) dist[i] = (IO)s.dist;
    if (J) { J[2 * i] = (IO)s.Jx; J[2 * i + 1] = (IO)s.Jy; }
  }
}

// ---------------------------------------------------------------------------
// Signed Euclidean distance field of an occupancy image -- replaces sdf_2d (utils/sdf_utils.py:6-21,
// datasets/utils.py:4-18), i.e. two scipy.ndimage.distance_transform_edt calls per map.  Exact EDT, separable:
//   A. one warp per image row: the row's free / obstacle pixels as bit masks (__ballot_sync over coalesced loads);
//   B. per pixel and polarity, the distance ALONG THE ROW to the nearest pixel of that polarity: nearest set bit of the
//      row mask to the left / right (shift + ffs / clz, no scan) -> one byte per pixel and polarity (255 = none);
//   C. per column and polarity (one thread each), the exact lower envelope of the parabolas (y - i)^2 + g(i, x)^2 by the
//      integer stack algorithm of Meijster et al. (a forward sweep that builds the envelope, a backward sweep that
//      evaluates it; O(H) per column whatever the distances are) -- the backward sweep runs over the rows in lock step,
//      so the lanes of a warp write consecutive pixels of one row; sqrt in double.
// A pixel is background of one of the two transforms (distance 0 there), so each pixel needs ONE polarity: free pixels
// look for the nearest obstacle, obstacle pixels for the nearest free pixel.  Bit-identical to the scipy path, including
// scipy's behaviour for an image WITHOUT any background pixel (distance to a virtual pixel at row -1, column 0).
// One CTA per image; 1 byte + 2 mask bits of shared memory per pixel and one stack byte per (task, row): 52 KB for a
// 128 x 128 map, 4 images resident per SM (round 2, session 2: was 102 KB / 2 per SM; 0.355 -> 0.345 ms per 1024 maps -- the
// kernel is bound by the divergent, serial envelope sweeps, not by occupancy).
// ---------------------------------------------------------------------------
__device__ __forceinline__ int edt_nearest_bit(const unsigned* __restrict__ m, int NW, int x) {
  const int w = x >> 5, bp = x & 31;
  int dr = 255, dl = 255;
  unsigned v = m[w] >> bp;                                   // bits at positions >= x
  if (v != 0u) {
    dr = __ffs(v) - 1;
  } else {
    for (int q = w + 1; q < NW; ++q) {
      const unsigned u = m[q];
      if (u != 0u) { dr = (q << 5) + __ffs(u) - 1 - x; break; }
    }
  }
  v = m[w] << (31 - bp);                                     // bits at positions <= x, the bit at x in the msb
  if (v != 0u) {
    dl = __clz(v);
  } else {
    for (int q = w - 1; q >= 0; --q) {
      const unsigned u = m[q];
      if (u != 0u) { dl = x - ((q << 5) + 31 - __clz(u)); break; }
    }
  }
  return min(min(dl, dr), 255);
}

// bit-packed occupancy input: row-major, every row padded to 32-bit words, bit (x & 31) of word x >> 5 set = pixel FREE
struct OccBits { unsigned w; };

struct EdtSmem {
  unsigned* m_obst;      // [Hp][NW] bit x of row y: pixel (y, x) is an obstacle
  unsigned* m_free;      // [Hp][NW] ... is free
  unsigned char* g;      // [Hp*Wp] distance along the row to the nearest pixel of the OTHER polarity (255: none in this row);
                         //         a pixel is at distance 0 from its own polarity, which the mask bit tells
  unsigned char* st_s;   // [Hp][threads] envelope stacks, entry k of the thread's task at [k * threads + tid]: parabola index (row).
                         //         The first row at which a parabola is the lowest is recomputed from its predecessor on a
                         //         pop instead of being stored: half the stack bytes, 4 instead of 2 CTAs per SM.
  __host__ __device__ static size_t bytes(int Hp, int Wp, int threads) {
    const size_t NW = (size_t)(Wp + 31) / 32;
    const size_t stacks = (size_t)Hp * threads, tmp = (size_t)Hp * Wp;      // (the stacks' area first holds one byte per pixel)
    return 2 * (size_t)Hp * NW * 4 + (((size_t)Hp * Wp + 15) & ~(size_t)15) + (stacks > tmp ? stacks : tmp);
  }
};

template <typename IN, typename IO>
__global__ void __launch_bounds__(256)
sdf_from_occupancy_kernel(const IN* __restrict__ im, int H, int W, int pad, double thresh, double res,
                          IO* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int Hp = H + 2 * pad, Wp = W + 2 * pad, NW = (Wp + 31) >> 5;
  EdtSmem S;
  S.m_obst = reinterpret_cast<unsigned*>(smem_raw);
  S.m_free = S.m_obst + (size_t)Hp * NW;
  S.g = reinterpret_cast<unsigned char*>(S.m_free + (size_t)Hp * NW);
  unsigned char* const after_g = S.g + (((size_t)Hp * Wp + 15) & ~(size_t)15);
  const IN* src = im + (size_t)blockIdx.x * H * W;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  // ---- A: row masks.  All pixels are fetched first (independent, coalesced loads; one DRAM latency for the image,
  // not one per row) as one byte each into the area the envelope stacks use later, then one warp per row ballots them.
  int any_obst = 0, any_free = 0;
  if constexpr (sizeof(IN) == sizeof(OccBits) && !std::is_arithmetic<IN>::value) {
    // bit-packed input (pad == 0, checked by the host): the row masks ARE the input words
    const unsigned* bits = reinterpret_cast<const unsigned*>(im) + (size_t)blockIdx.x * H * NW;
    for (int i = threadIdx.x; i < Hp * NW; i += blockDim.x) {
      const int c = i % NW;
      const unsigned valid = (c == NW - 1 && (Wp & 31)) ? ((1u << (Wp & 31)) - 1u) : 0xffffffffu;
      const unsigned bf = __ldg(bits + i) & valid, bo = ~bf & valid;
      S.m_free[i] = bf; S.m_obst[i] = bo;
      any_obst |= (bo != 0u);
      any_free |= (bf != 0u);
    }
  } else {
    unsigned char* tmp = after_g;
#pragma unroll 8
    for (int i = threadIdx.x; i < Hp * Wp; i += blockDim.x) {
      const int y = i / Wp, x = i - y * Wp, yy = y - pad, xx = x - pad;
      const bool inside = yy >= 0 && yy < H && xx >= 0 && xx < W;
      bool free_px = true;                                                                     // padding is free space
      if constexpr (std::is_arithmetic<IN>::value) {
        if (inside) free_px = (double)__ldg(src + (size_t)yy * W + xx) > thresh;
      }
      tmp[i] = free_px ? 1 : 0;
    }
    __syncthreads();
    for (int y = warp; y < Hp; y += nwarps) {
      for (int c = 0; c < NW; ++c) {
        const int x = (c << 5) + lane;
        const bool valid = x < Wp;
        const bool free_px = valid && tmp[(size_t)y * Wp + (valid ? x : 0)] != 0;
        const unsigned bf = __ballot_sync(0xffffffffu, free_px);
        const unsigned bo = __ballot_sync(0xffffffffu, valid && !free_px);
        if (lane == 0) { S.m_free[y * NW + c] = bf; S.m_obst[y * NW + c] = bo; }
        any_obst |= (bo != 0u);
        any_free |= (bf != 0u);
      }
    }
  }
  const int has_obst = __syncthreads_or(any_obst);
  const int has_free = __syncthreads_or(any_free);
  // ---- B: distance along the row, both polarities ----
  for (int i = threadIdx.x; i < Hp * Wp; i += blockDim.x) {
    const int y = i / Wp, x = i - y * Wp;
    // a pixel is at distance 0 from its own polarity: one search per pixel, for the other polarity
    const bool is_free = (S.m_free[y * NW + (x >> 5)] >> (x & 31)) & 1u;
    S.g[i] = (unsigned char)edt_nearest_bit((is_free ? S.m_obst : S.m_free) + y * NW, NW, x);
  }
  __syncthreads();
  // ---- C: lower envelope per (column, polarity) ----
  S.st_s = after_g;
  IO* dst = out + (size_t)blockIdx.x * Hp * Wp;
  const int NT = 2 * Wp;                                       // tasks: polarity * Wp + x
  constexpr int BIG = 1 << 20;                                 // "no pixel of that polarity in this row": above every real value
  for (int base = 0; base < NT; base += blockDim.x) {          // uniform trip count (one pass for Wp <= 128)
    const int j = base + threadIdx.x;
    const bool on = j < NT;
    const int pol = on ? j / Wp : 0, x = on ? j - pol * Wp : 0;
    const unsigned char* g = S.g + x;                          // column x of the row distances
    const unsigned* mf = S.m_free + (x >> 5);                  // ... and of the free-pixel mask
    const int xb = x & 31;
    unsigned char* ss = S.st_s + threadIdx.x;                  // (the stacks are reused by every batch of tasks)
    const int SN = blockDim.x;
    // squared distance along row i from (i, x) to the nearest pixel of polarity `pol` (0: obstacle, 1: free): 0 for a
    // pixel of that polarity itself
    auto G = [&](int i) {
      const int is_free = (int)((mf[(size_t)i * NW] >> xb) & 1u);
      const int a = (is_free != pol) ? (int)g[(size_t)i * Wp] : 0;
      return (a == 255) ? BIG : a * a;
    };
    // 1 + Sep(s, u) = first row at which parabola u (> s) is lower than parabola s:
    // Sep(s, u) = floor((u^2 - s^2 + Gu - Gs) / (2 (u - s)))
    auto first_row = [&](int s_, int Gs, int u_, int Gu) {
      const int num = u_ * u_ - s_ * s_ + Gu - Gs, den = 2 * (u_ - s_);
      int sep = (int)__fdividef((float)num, (float)den);
      sep += ((sep + 1) * den <= num) ? 1 : 0;
      sep -= (sep * den > num) ? 1 : 0;
      return 1 + sep;
    };
    // top of the stack cached in registers: parabola sq, lowest from row tq on, Gsq = G(sq)
    auto reload_top = [&](int q, int& sq, int& tq, int& Gsq) {
      sq = ss[(size_t)q * SN];
      Gsq = G(sq);
      if (q == 0) { tq = 0; return; }
      const int sp = ss[(size_t)(q - 1) * SN];
      tq = first_row(sp, G(sp), sq, Gsq);
    };
    int q = 0, sq = 0, tq = 0, Gsq = 0;
    if (on) {
      Gsq = G(0);
      ss[0] = 0;
      for (int u = 1; u < Hp; ++u) {
        const int Gu = G(u);
        // pop parabolas that the one at u beats at the start of their interval
        while (q >= 0 && (tq - sq) * (tq - sq) + Gsq > (tq - u) * (tq - u) + Gu) {
          --q;
          if (q >= 0) reload_top(q, sq, tq, Gsq);
        }
        if (q < 0) {
          q = 0; sq = u; tq = 0; Gsq = Gu;
          ss[0] = (unsigned char)u;
        } else {
          const int w = first_row(sq, Gsq, u, Gu);             // >= tq >= 0
          if (w < Hp) {
            ++q; sq = u; tq = w; Gsq = Gu;
            ss[(size_t)q * SN] = (unsigned char)u;
          }
        }
      }
    }
    // backward sweep, all tasks at the same row u: the lanes of a warp write consecutive pixels of that row
    for (int u = Hp - 1; u >= 0; --u) {
      if (on) {
        const int d2 = (u - sq) * (u - sq) + Gsq;
        const bool is_free = ((mf[(size_t)u * NW] >> xb) & 1u) != 0u;
        if ((pol == 0) == is_free) {          // free pixels take the distance to the obstacles, obstacle pixels to free space
          // EDT(im) for free pixels (-> nearest obstacle), EDT(1 - im) for obstacle pixels (-> nearest free pixel).  No
          // background pixel at all (uniform): scipy (1.18) measures from a virtual pixel at row -1, column 0.
          const bool have = is_free ? (has_obst != 0) : (has_free != 0);
          const long long sq2 = have ? (long long)d2 : ((long long)(u + 1) * (u + 1) + (long long)x * x);
          const double d = sqrt((double)sq2);
          dst[(size_t)u * Wp + x] = (IO)((is_free ? d : (0.0 - d)) * res);
        }
        if (u == tq && q > 0) {
          --q;
          reload_top(q, sq, tq, Gsq);
        }
      }
    }
  }
}

}  // namespace dgpmp2
