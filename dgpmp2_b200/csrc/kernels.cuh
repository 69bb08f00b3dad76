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
#include "bcr.cuh"

namespace dgpmp2 {

// Launch-shape constants of the step / solve kernels for NN node slots and LPN lanes per BCR item.
template <int D, int NN, int LPN>
struct StepShape {
  static constexpr int kCap = (D == 4) ? 512 : 384;                       // threads per CTA upper bound
  static constexpr int kMaxThreads = (LPN * NN / 2 < kCap) ? ((LPN * NN / 2 + 31) / 32 * 32) : kCap;
  // CTAs per SM that shared memory allows (band = kDoublesPerNode doubles per slot); the register
  // budget is capped so that registers never limit residency below that.
  static constexpr int kSmemPerCta = (Band<D, NN>::kDoublesPerNode + 3) * NN * 8 + NN * D * 8 + 1024;
  static constexpr int kMinBlocks = (232448 / kSmemPerCta) < 1 ? 1 : ((232448 / kSmemPerCta) > 8 ? 8 : (232448 / kSmemPerCta));
};

// Shared-memory carve-up of the step / solve kernels (NN node slots, compile time).
template <int D, int NN, typename IO>
struct StepSmem {
  Band<D, NN> band;
  double* errp;     // [2][NN] per-node error partials (err, err_ext), slot order
  double* nrm;      // [NN]    per-node |dth|^2 (solve kernel only)
  IO* th;           // [NN][D] staged trajectory, natural (problem, t, a) order
  int* lvl_off;     // [kMaxLevels + 2]
  int* fail;        // [NPmax]
  int* flags;       // [2*NPmax] solve kernel: converged flag per problem, then iteration count
  static constexpr int kMaxNP = NN / 2;
  __host__ __device__ static constexpr size_t bytes(bool solve) {
    size_t b = (size_t)Band<D, NN>::kDoublesPerNode * NN * 8 + 2 * (size_t)NN * 8;
    if (solve) b += (size_t)NN * 8;
    b += (size_t)NN * D * sizeof(IO);
    b = (b + 7) & ~(size_t)7;
    b += (kMaxLevels + 2) * 4 + (size_t)kMaxNP * 4 * 3 + 16;
    return b;
  }
  __device__ __forceinline__ void carve(unsigned char* raw, bool solve) {
    double* base = reinterpret_cast<double*>(raw);
    band.base = base;
    errp = base + (size_t)Band<D, NN>::kDoublesPerNode * NN;
    double* nxt = errp + 2 * (size_t)NN;
    nrm = nxt;
    if (solve) nxt += NN;
    th = reinterpret_cast<IO*>(nxt);
    size_t off = (reinterpret_cast<unsigned char*>(th + (size_t)NN * D) - raw + 7) & ~(size_t)7;
    lvl_off = reinterpret_cast<int*>(raw + off);
    fail = lvl_off + (kMaxLevels + 2);
    flags = fail + kMaxNP;
  }
};

// Thread geometry shared by the step and solve kernels: each problem of the CTA owns TPP
// consecutive threads; thread u of problem p handles node slot u during assembly (u < T) and is
// lane (u % LPN) of BCR work item (u / LPN).
struct CtaGeom { int p, u; bool active; };
__device__ __forceinline__ CtaGeom cta_geom(int TPP, int np) {
  CtaGeom g;
  g.p = threadIdx.x / TPP;
  g.u = threadIdx.x - g.p * TPP;
  g.active = g.p < np;
  return g;
}

// Assemble this thread's nodes (slots u, u + TPP, ... of problem p) from the staged trajectory into the band.
template <int DOF, int NN, typename IO>
__device__ __forceinline__ void assemble_cta(const KParams& P, const KWeights<IO>& Wt, const StepSmem<2 * DOF, NN, IO>& S,
                                             int b0, const CtaGeom& g, int TPP, int nlev,
                                             const IO* __restrict__ start, const IO* __restrict__ goal,
                                             const IO* __restrict__ sdf) {
  constexpr int D = 2 * DOF;
  const int T = P.T;
  if (!g.active) return;
  for (int slot = g.u; slot < T; slot += TPP) {
  const int t = bcr_state_of_slot(S.lvl_off, nlev, T, slot);
  const int b = b0 + g.p;
  double thp[D], thc[D], thn[D];
  const IO* tp = S.th + ((size_t)g.p * T + t) * D;
#pragma unroll
  for (int a = 0; a < D; ++a) {
    thc[a] = (double)tp[a];
    thp[a] = (t > 0) ? (double)tp[a - D] : 0.0;
    thn[a] = (t < T - 1) ? (double)tp[a + D] : 0.0;
  }
  NodeOut<DOF> o;
  assemble_node<DOF, IO>(P, Wt, b, t, thp, thc, thn, start + (size_t)b * D, goal + (size_t)b * D,
                         sdf + (size_t)b * P.sdf_sb, o);
  const int n = g.p * T + slot;
  double* dp = S.band.Dp(n);
  double* up = S.band.Up(n);
  double* rp = S.band.Rp(n);
#pragma unroll
  for (int a = 0; a < D; ++a) {
#pragma unroll
    for (int c = 0; c <= a; ++c) dp[tri(a, c) * NN] = o.Dm[a][c];
#pragma unroll
    for (int c = 0; c < D; ++c) up[(a * D + c) * NN] = o.Um[a][c];
    rp[a * NN] = o.r[a];
  }
  S.errp[n] = o.err;
  S.errp[NN + n] = o.err_ext;
  }
}

// Deterministic per-problem sum of `vals[p*T .. p*T+T)` by warp w for problems w, w+nwarps, ...
// `f(p, sum)` is called by lane 0.
template <typename F>
__device__ __forceinline__ void reduce_per_problem(const double* vals, int np, int T, F f) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int p = warp; p < np; p += nwarps) {
    double s = 0.0;
    for (int i = lane; i < T; i += 32) s += vals[p * T + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) f(p, s);
  }
}

template <int D, int NN, typename IO>
__device__ __forceinline__ void cta_prologue(const StepSmem<D, NN, IO>& S, int T, int NP, int np,
                                             const IO* __restrict__ th_src, bool solve) {
  if (threadIdx.x == 0) {
    BcrLevels lv;
    bcr_make_levels(T, lv);
    for (int l = 0; l <= lv.nlev + 1; ++l) S.lvl_off[l] = lv.off[l];
    S.lvl_off[kMaxLevels + 1] = lv.nlev;
  }
  for (int p = threadIdx.x; p < NP; p += blockDim.x) {
    S.fail[p] = 0;
    if (solve) { S.flags[p] = 0; S.flags[NP + p] = 0; }
  }
  const int n = np * T * D;   // contiguous in HBM -> coalesced
  for (int i = threadIdx.x; i < n; i += blockDim.x) S.th[i] = __ldg(th_src + i);
}

// One fused Gauss-Newton iteration.  grid = ceil(B / NP), block = NP * TPP threads (rounded to a warp).
template <int DOF, int NN, int LPN, typename IO>
__global__ void __launch_bounds__((StepShape<2 * DOF, NN, LPN>::kMaxThreads), (StepShape<2 * DOF, NN, LPN>::kMinBlocks))
gn_step_kernel(const KParams P, const KWeights<IO> Wt, const IO* __restrict__ th, const IO* __restrict__ start,
               const IO* __restrict__ goal, const IO* __restrict__ sdf, IO* __restrict__ dth,
               IO* __restrict__ err, IO* __restrict__ err_ext, int* __restrict__ status, const int NP, const int TPP) {
  constexpr int D = 2 * DOF;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  StepSmem<D, NN, IO> S;
  const int T = P.T;
  S.carve(smem_raw, false);
  const int b0 = blockIdx.x * NP;
  const int np = min(NP, P.B - b0);
  const CtaGeom g = cta_geom(TPP, np);

  cta_prologue<D, NN, IO>(S, T, NP, np, th + (size_t)b0 * T * D, false);
  __syncthreads();
  const int nlev = S.lvl_off[kMaxLevels + 1];

  assemble_cta<DOF, NN, IO>(P, Wt, S, b0, g, TPP, nlev, start, goal, sdf);
  __syncthreads();

  bcr_solve<D, NN, LPN>(S.band, S.lvl_off, nlev, T, g.active, g.p, g.u, TPP / LPN, S.fail);   // ends with a barrier

  {  // dth, natural order -> coalesced stores
    IO* dst = dth + (size_t)b0 * T * D;
    const int n = np * T * D;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int a = i % D, pt = i / D;
      const int p = pt / T, t = pt - p * T;
      dst[i] = (IO)S.band.Rp(p * T + bcr_slot(S.lvl_off, T, t))[a * NN];
    }
  }
  const double invM = 1.0 / (double)P.M;
  reduce_per_problem(S.errp, np, T, [&](int p, double s) { err[b0 + p] = (IO)(s * invM); });
  reduce_per_problem(S.errp + NN, np, T, [&](int p, double s) { err_ext[b0 + p] = (IO)(s * invM); });
  if (status != nullptr)
    for (int p = threadIdx.x; p < np; p += blockDim.x) status[b0 + p] = S.fail[p];
}

// ---------------------------------------------------------------------------
// Persistent solve-to-convergence (DiffGPMP2Planner.forward, diff_gpmp2_planner.py:104-165)
// ---------------------------------------------------------------------------
template <int DOF, int NN, int LPN, typename IO>
__global__ void __launch_bounds__((StepShape<2 * DOF, NN, LPN>::kMaxThreads), (StepShape<2 * DOF, NN, LPN>::kMinBlocks))
gn_solve_kernel(const KParams P, const KWeights<IO> Wt, const IO* __restrict__ th_init, const IO* __restrict__ start,
                const IO* __restrict__ goal, const IO* __restrict__ sdf, const int max_iters, const double tol_delta,
                IO* __restrict__ th_final, int* __restrict__ iters, IO* __restrict__ err_pi, IO* __restrict__ err_ext_pi,
                IO* __restrict__ err_final, IO* __restrict__ err_ext_final, int* __restrict__ status, const int NP,
                const int TPP) {
  constexpr int D = 2 * DOF;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  StepSmem<D, NN, IO> S;
  const int T = P.T;
  S.carve(smem_raw, true);
  int* done = S.flags;          // [NP]
  int* nit = S.flags + NP;      // [NP]
  const int b0 = blockIdx.x * NP;
  const int np = min(NP, P.B - b0);
  const CtaGeom g = cta_geom(TPP, np);

  cta_prologue<D, NN, IO>(S, T, NP, np, th_init + (size_t)b0 * T * D, true);
  __syncthreads();
  const int nlev = S.lvl_off[kMaxLevels + 1];
  const double invM = 1.0 / (double)P.M;

  for (int j = 0;; ++j) {
    // assemble at the current iterate; the errors at iterate j are a by-product
    assemble_cta<DOF, NN, IO>(P, Wt, S, b0, g, TPP, nlev, start, goal, sdf);
    __syncthreads();
    const bool last = (j >= max_iters);
    reduce_per_problem(S.errp, np, T, [&](int p, double s) {
      if (!done[p] && !last && err_pi != nullptr) err_pi[(size_t)(b0 + p) * max_iters + j] = (IO)(s * invM);
      if ((done[p] == 1 || last) && err_final != nullptr) err_final[b0 + p] = (IO)(s * invM);
    });
    reduce_per_problem(S.errp + NN, np, T, [&](int p, double s) {
      if (!done[p] && !last && err_ext_pi != nullptr) err_ext_pi[(size_t)(b0 + p) * max_iters + j] = (IO)(s * invM);
      if ((done[p] == 1 || last) && err_ext_final != nullptr) err_ext_final[b0 + p] = (IO)(s * invM);
    });
    __syncthreads();
    // problems that converged at the previous iteration have now had their final error recorded
    for (int p = threadIdx.x; p < np; p += blockDim.x)
      if (done[p] == 1) done[p] = 2;
    __syncthreads();
    bool all_done = true;
    for (int p = 0; p < np; ++p) all_done = all_done && (done[p] == 2);
    if (all_done || last) break;

    bcr_solve<D, NN, LPN>(S.band, S.lvl_off, nlev, T, g.active, g.p, g.u, TPP / LPN, S.fail);

    // th <- th + dth for problems still running; |dth|^2 partials
    for (int m = threadIdx.x; m < np * T; m += blockDim.x) {
      const int p = m / T, t = m - p * T;
      const int n = p * T + bcr_slot(S.lvl_off, T, t);
      double s2 = 0.0;
      if (!done[p]) {
#pragma unroll
        for (int a = 0; a < D; ++a) {
          const double dx = S.band.Rp(n)[a * NN];
          // the reference adds dtheta (I/O dtype) to th (I/O dtype): round dth first, then add
          const IO dxi = (IO)dx;
          s2 += (double)dxi * (double)dxi;
          IO* q = S.th + ((size_t)p * T + t) * D + a;
          *q = (IO)(*q + dxi);
        }
      }
      S.nrm[m] = s2;
    }
    __syncthreads();
    reduce_per_problem(S.nrm, np, T, [&](int p, double s) {
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
      const double eps = (Wt.eps != nullptr) ? ldg_d(Wt.eps + (long long)b * Wt.e_sb + (long long)t * Wt.e_st) : P.eps_const;
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

}  // namespace dgpmp2
