// extern "C" entry points of libdgpmp2_b200.so (declared in include/dgpmp2_b200.h).
// Host-side argument checking, derived constants, launch-shape selection, launches.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>
#include <utility>
#include <vector>

#include "../../include/dgpmp2_b200.h"
#include "kernels.cuh"
#include "host_params.h"

using namespace dgpmp2;

namespace {

thread_local char g_cuda_err[256] = {0};

int cuda_fail(cudaError_t e) {
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", cudaGetErrorName(e), cudaGetErrorString(e));
  return DGPMP2_ERR_CUDA;
}
#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return cuda_fail(e_); } while (0)

constexpr int kPdlDefault = 2;       // DGPMP2_PDL: 1 = plain launches, 2 = programmatic dependent launch of gn_step (default)
constexpr int kSmemLimit = 232448;   // 227 KB opt-in dynamic shared memory per CTA on sm_100
constexpr int kMpSmemLimit = kSmemLimit - 2048;   // gn_step_mp_kernel also holds ~1.2 KB of static shared memory

int check_params(const dgpmp2_params* p, const dgpmp2_weights* w) {
  if (p == nullptr) return DGPMP2_ERR_ARG;
  if (p->B < 0 || p->T < 2 || p->H < 1 || p->W < 1) return DGPMP2_ERR_ARG;
  if (p->dof != 2 && p->dof != 3) return DGPMP2_ERR_UNSUPPORTED;
  if (p->flags & ~(DGPMP2_FLAG_NONHOLONOMIC | DGPMP2_FLAG_VEL_LIMITS | DGPMP2_FLAG_Q_FULL | DGPMP2_FLAG_HEAD |
                   DGPMP2_FLAG_HEAD_QC_VEC)) return DGPMP2_ERR_ARG;
  if ((p->flags & DGPMP2_FLAG_HEAD_QC_VEC) && !(p->flags & DGPMP2_FLAG_HEAD)) return DGPMP2_ERR_ARG;
  if ((p->flags & DGPMP2_FLAG_HEAD) && w == nullptr) return DGPMP2_ERR_ARG;   // a head without outputs
  if ((p->flags & DGPMP2_FLAG_NONHOLONOMIC) && (p->flags & DGPMP2_FLAG_VEL_LIMITS)) return DGPMP2_ERR_ARG;
  if ((p->flags & DGPMP2_FLAG_NONHOLONOMIC) && p->dof != 3) return DGPMP2_ERR_ARG;
  if ((p->flags & DGPMP2_FLAG_VEL_LIMITS) && p->dof != 2) return DGPMP2_ERR_ARG;
  if ((p->flags & DGPMP2_FLAG_Q_FULL) && (w == nullptr || w->qc_inv == nullptr)) return DGPMP2_ERR_ARG;
  if (!(p->res > 0.0) || !(p->dt > 0.0)) return DGPMP2_ERR_ARG;
  if (p->sdf_stride_b < 0) return DGPMP2_ERR_ARG;
  return DGPMP2_OK;
}

struct LaunchShape { int np, tpp, threads, smem, grid, n_big; };   // CTAs [0, n_big) take np problems, the others np - 1

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// Problems per CTA (NP) and CTA size.  All problems an SM has to process are put in ONE CTA when
// they fit (NP = ceil(B / #SMs), bounded by shared memory and the thread limit): the BCR work items
// of all of them are packed onto consecutive lanes, so the sparse deep levels of several problems
// share warps, and one CTA per SM launches without a ramp.  Threads = kLPN lanes per level-1 item
// when that fits the CTA, else one thread per node record.
//
// balanced (gn_step_kernel): when an SM's share q = ceil(B / #SMs) does not fit one CTA, it is cut into
// w = ceil(q / np_max) CTAs of as equal a size as possible -- np = ceil(q / w) problems in the first CTA(s) an SM
// receives, np - 1 in the later ones (the hardware hands out CTAs in index order) -- instead of w CTAs of np_max with
// a ragged last wave: T = 128, B = 1024 runs 148 x 4 + 144 x 3 problems instead of 256 x 4 in waves of 148 + 108.
template <int D, typename IO>
int choose_shape(int B, int T, int mode, LaunchShape& s, bool balanced = false) {
  const int max_threads = (D == 4) ? 512 : 256;
  const int items = (T + 1) / 2;                 // level-1 work items (and assembly needs T threads ~ 2 * items)
  const int q = (B + sm_count() - 1) / sm_count();
  int np = env_int("DGPMP2_NP", q);
  if (np > B) np = B;
  if (np < 1) np = 1;
  while (np > 1 && StepSmem<D, IO>::bytes(np, T, mode) > (size_t)kSmemLimit) --np;
  if (balanced) { const int cap = env_int("DGPMP2_NP_MAX", np); if (cap < np) np = cap; }   // experiment: several smaller co-resident CTAs per SM
  int big_waves = -1;                            // -1: every CTA takes np problems
  if (balanced && np < q && env_int("DGPMP2_BALANCED", 1) == 1 && getenv("DGPMP2_NP") == nullptr) {
    const int w = (q + np - 1) / np;
    np = (q + w - 1) / w;                        // <= the shared-memory bound found above
    big_waves = q - w * (np - 1);                // q = big_waves * np + (w - big_waves) * (np - 1)
    if (np == 1 || big_waves >= w) big_waves = -1;
  }
  const size_t bytes = StepSmem<D, IO>::bytes(np, T, mode);
  if (bytes > (size_t)kSmemLimit) return DGPMP2_ERR_UNSUPPORTED;
  int threads = kLPN * items * np;
  threads = (threads + 31) / 32 * 32;
  if (threads > max_threads) {
    // cannot give every level-1 item its kLPN lanes: one thread per node record (assembly, one-lane levels)
    // is then enough; more warps would only wait at the barriers
    threads = (np * T + 31) / 32 * 32;
    if (threads > max_threads) threads = max_threads;
  }
  if (threads < 64) threads = 64;
  threads = env_int("DGPMP2_THREADS", threads);
  if (threads > max_threads) threads = max_threads;
  s.np = np; s.tpp = threads / np;
  s.threads = threads;
  s.smem = (int)bytes;
  s.grid = (B + np - 1) / np;
  s.n_big = s.grid;
  if (big_waves >= 0) {
    const long long cap_big = (long long)sm_count() * big_waves * np;
    if (cap_big < B) {
      s.n_big = sm_count() * big_waves;
      s.grid = s.n_big + (int)((B - cap_big + np - 2) / (np - 1));
    }
  }
  return DGPMP2_OK;
}

// Opt a kernel in to > 48 KB of dynamic shared memory, once per (kernel, device).
template <typename K>
int allow_smem(K kernel, int bytes, int limit = kSmemLimit) {
  if (bytes <= 48 * 1024) return DGPMP2_OK;
  static std::mutex mu;
  static std::vector<std::pair<const void*, int>> done;
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  const void* key = reinterpret_cast<const void*>(kernel);
  {
    std::lock_guard<std::mutex> lk(mu);
    for (const auto& e : done)
      if (e.first == key && e.second == dev) return DGPMP2_OK;
  }
  CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, limit));
  std::lock_guard<std::mutex> lk(mu);
  done.emplace_back(key, dev);
  return DGPMP2_OK;
}

template <int DOF, typename IO>
int launch_step(const KParams& k, const KWeights<IO>& kw, const IO* th, const IO* start, const IO* goal, const IO* sdf,
                IO* dth, IO* err, IO* err_ext, int32_t* status, cudaStream_t st) {
  LaunchShape s;
  int rc = choose_shape<2 * DOF, IO>(k.B, k.T, 0, s, true);
  if (rc != DGPMP2_OK) return rc;
  auto kern = gn_step_kernel<DOF, IO>;
  rc = allow_smem(kern, s.smem);
  if (rc != DGPMP2_OK) return rc;
  if (env_int("DGPMP2_PDL", kPdlDefault) == 2) {
    // programmatic dependent launch: the launch latency of step n + 1 hides behind step n (the kernel waits at
    // griddepcontrol.wait before it touches global memory, so stream order is preserved for every access)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)s.grid);
    cfg.blockDim = dim3((unsigned)s.threads);
    cfg.dynamicSmemBytes = (size_t)s.smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, k, kw, th, start, goal, sdf, dth, err, err_ext, status, s.np, s.n_big));
    return DGPMP2_OK;
  }
  kern<<<s.grid, s.threads, s.smem, st>>>(k, kw, th, start, goal, sdf, dth, err, err_ext, status, s.np, s.n_big);
  CUDA_TRY(cudaGetLastError());
  return DGPMP2_OK;
}

// ---- mixed-precision step (mp.cuh) --------------------------------------------------------------------------------
// Shape: TPP = ceil32(T) threads per problem (one per node), NP problems per CTA = the SM's share of the batch, bounded
// by the named barriers (15 problems of more than one warp), the CTA size and shared memory.  CTAs of <= 512 threads
// run the 128-register instantiation, larger ones the 64-register one.
struct MpShape { int np, tpp, threads, smem, grid, maxt; };

template <int D>
int choose_shape_mp(int B, int T, MpShape& s) {
  const int tpp = (T + 31) / 32 * 32;
  if (tpp > 1024) return DGPMP2_ERR_UNSUPPORTED;
  const size_t legacy1 = StepSmem<D, float>::bytes(1, T, 0);      // the in-kernel fp64 path needs its own carve-up
  if (legacy1 > (size_t)kMpSmemLimit) return DGPMP2_ERR_UNSUPPORTED;
  const int maxt = env_int("DGPMP2_MP_MAXT", 1024) <= 512 ? 512 : 1024;
  int np = (B + sm_count() - 1) / sm_count();
  np = env_int("DGPMP2_NP", np);
  if (np > B) np = B;
  const int cap_bar = (tpp == 32) ? kMpMaxNP : 15;
  if (np > cap_bar) np = cap_bar;
  if (np * tpp > maxt) np = maxt / tpp;
  if (np < 1) {
    if (tpp > maxt) return DGPMP2_ERR_UNSUPPORTED;
    np = 1;
  }
  const size_t per = MpRec<D>::problem_floats(T) * sizeof(float);
  while (np > 1 && (size_t)np * per > (size_t)kMpSmemLimit) --np;
  size_t bytes = (size_t)np * per;
  if (bytes > (size_t)kMpSmemLimit) return DGPMP2_ERR_UNSUPPORTED;
  if (bytes < legacy1) bytes = legacy1;
  s.np = np; s.tpp = tpp; s.threads = np * tpp; s.smem = (int)bytes; s.grid = (B + np - 1) / np;
  s.maxt = (s.threads <= 512) ? 512 : 1024;
  return DGPMP2_OK;
}

// DGPMP2_PRECISION=32 routes the fp32-I/O step through the mixed-precision kernel (mp.cuh).  It is opt-in: measured on
// B200 it is SLOWER than the all-double kernel at every BASELINE shape (profiles/r02_mp_experiment.md) -- fp64 issues at
// half the fp32 rate on this part, the step is bound by dependent-instruction latency, and the two refinement sweeps
// the 1e-5 parity bar needs cost more chain length than the fp32 factorisation saves.
bool mp_enabled() { return env_int("DGPMP2_PRECISION", 64) == 32; }

template <int DOF>
int launch_step_mp(const KParams& k, const KWeights<float>& kw, const float* th, const float* start, const float* goal,
                   const float* sdf, float* dth, float* err, float* err_ext, int32_t* status, int32_t* diag,
                   cudaStream_t st, bool& launched) {
  constexpr int D = 2 * DOF;
  launched = false;
  MpShape s;
  if (choose_shape_mp<D>(k.B, k.T, s) != DGPMP2_OK) return DGPMP2_OK;   // not applicable: the caller uses the fp64 kernel
  const int force64 = env_int("DGPMP2_MP_FORCE64", 0) == 1 ? 1 : 0;
  int rc;
  if (s.maxt == 512) {
    auto kern = gn_step_mp_kernel<DOF, 512>;
    rc = allow_smem(kern, s.smem, kMpSmemLimit);
    if (rc != DGPMP2_OK) return rc;
    kern<<<s.grid, s.threads, s.smem, st>>>(k, kw, th, start, goal, sdf, dth, err, err_ext, status, diag, s.np, s.tpp, force64);
  } else {
    auto kern = gn_step_mp_kernel<DOF, 1024>;
    rc = allow_smem(kern, s.smem, kMpSmemLimit);
    if (rc != DGPMP2_OK) return rc;
    kern<<<s.grid, s.threads, s.smem, st>>>(k, kw, th, start, goal, sdf, dth, err, err_ext, status, diag, s.np, s.tpp, force64);
  }
  CUDA_TRY(cudaGetLastError());
  launched = true;
  return DGPMP2_OK;
}

template <typename IO>
int gn_step_impl(const dgpmp2_params* p, const IO* th, const IO* start, const IO* goal, const IO* sdf,
                 const dgpmp2_weights* w, IO* dth, IO* err, IO* err_ext, int32_t* status, void* stream,
                 int32_t* diag = nullptr) {
  int rc = check_params(p, w);
  if (rc != DGPMP2_OK) return rc;
  if (p->B == 0) return DGPMP2_OK;
  if (!th || !start || !goal || !sdf || !dth || !err || !err_ext) return DGPMP2_ERR_ARG;
  KParams k = make_kparams(p);
  const KWeights<IO> kw = make_kweights<IO>(w);
  finish_kparams(k, kw);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if constexpr (sizeof(IO) == 4) {
    if (mp_enabled()) {
      bool launched = false;
      rc = (p->dof == 2) ? launch_step_mp<2>(k, kw, th, start, goal, sdf, dth, err, err_ext, status, diag, st, launched)
                         : launch_step_mp<3>(k, kw, th, start, goal, sdf, dth, err, err_ext, status, diag, st, launched);
      if (rc != DGPMP2_OK || launched) return rc;
    }
  }
  if (diag != nullptr) CUDA_TRY(cudaMemsetAsync(diag, 0, sizeof(int32_t) * (size_t)p->B, st));   // 0 = all-double kernel
  if (p->dof == 2) return launch_step<2, IO>(k, kw, th, start, goal, sdf, dth, err, err_ext, status, st);
  return launch_step<3, IO>(k, kw, th, start, goal, sdf, dth, err, err_ext, status, st);
}

template <int DOF, typename IO>
int launch_solve(const KParams& k, const KWeights<IO>& kw, const IO* th, const IO* start, const IO* goal, const IO* sdf,
                 int max_iters, double tol, IO* th_final, int32_t* iters, IO* epi, IO* eepi, IO* ef, IO* eef,
                 int32_t* status, cudaStream_t st) {
  LaunchShape s;
  int rc = choose_shape<2 * DOF, IO>(k.B, k.T, 1, s);
  if (rc != DGPMP2_OK) return rc;
  auto kern = gn_solve_kernel<DOF, IO>;
  rc = allow_smem(kern, s.smem);
  if (rc != DGPMP2_OK) return rc;
  kern<<<s.grid, s.threads, s.smem, st>>>(k, kw, th, start, goal, sdf, max_iters, tol, th_final, iters, epi, eepi, ef,
                                          eef, status, s.np);
  CUDA_TRY(cudaGetLastError());
  return DGPMP2_OK;
}

template <typename IO>
int gn_solve_impl(const dgpmp2_params* p, const IO* th, const IO* start, const IO* goal, const IO* sdf,
                  const dgpmp2_weights* w, int32_t max_iters, double tol, IO* th_final, int32_t* iters, IO* epi,
                  IO* eepi, IO* ef, IO* eef, int32_t* status, void* stream) {
  int rc = check_params(p, w);
  if (rc != DGPMP2_OK) return rc;
  if (max_iters < 1) return DGPMP2_ERR_ARG;
  if (p->B == 0) return DGPMP2_OK;
  if (!th || !start || !goal || !sdf || !th_final || !iters) return DGPMP2_ERR_ARG;
  KParams k = make_kparams(p);
  const KWeights<IO> kw = make_kweights<IO>(w);
  finish_kparams(k, kw);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p->dof == 2)
    return launch_solve<2, IO>(k, kw, th, start, goal, sdf, max_iters, tol, th_final, iters, epi, eepi, ef, eef, status, st);
  return launch_solve<3, IO>(k, kw, th, start, goal, sdf, max_iters, tol, th_final, iters, epi, eepi, ef, eef, status, st);
}

template <int DOF, typename IO>
int launch_bwd(const KParams& k, const KWeights<IO>& kw, const IO* th, const IO* start, const IO* goal, const IO* sdf,
               const IO* dth, const IO* g_dth, const IO* g_err_ext, IO* g_th, IO* g_start, IO* g_goal, IO* g_qc, IO* g_w,
               IO* g_eps, IO* g_sdf, long long g_sdf_sb, cudaStream_t st) {
  LaunchShape s;
  int rc = choose_shape<2 * DOF, IO>(k.B, k.T, 2, s);
  if (rc != DGPMP2_OK) return rc;
  auto kern = gn_step_bwd_kernel<DOF, IO>;
  rc = allow_smem(kern, s.smem);
  if (rc != DGPMP2_OK) return rc;
  kern<<<s.grid, s.threads, s.smem, st>>>(k, kw, th, start, goal, sdf, dth, g_dth, g_err_ext, g_th, g_start, g_goal, g_qc,
                                          g_w, g_eps, g_sdf, g_sdf_sb, s.np);
  CUDA_TRY(cudaGetLastError());
  return DGPMP2_OK;
}

template <typename IO>
int gn_step_backward_impl(const dgpmp2_params* p, const IO* th, const IO* start, const IO* goal, const IO* sdf,
                          const dgpmp2_weights* w, const IO* dth, const IO* g_dth, const IO* g_err_ext, IO* g_th,
                          IO* g_start, IO* g_goal, IO* g_qc, IO* g_w, IO* g_eps, IO* g_sdf, void* stream) {
  int rc = check_params(p, w);
  if (rc != DGPMP2_OK) return rc;
  if (p->B == 0) return DGPMP2_OK;
  if (!th || !start || !goal || !sdf || !dth || !g_dth) return DGPMP2_ERR_ARG;
  KParams k = make_kparams(p);
  const KWeights<IO> kw = make_kweights<IO>(w);
  finish_kparams(k, kw);   // static-GP blocks + fused level 1 for the band assembly, as in the forward step
  if (env_int("DGPMP2_BWD_STATIC", 1) != 1) { k.static_gp = 0; k.fuse1 = 0; }   // A/B: the generic assembly of round 1
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p->dof == 2)
    return launch_bwd<2, IO>(k, kw, th, start, goal, sdf, dth, g_dth, g_err_ext, g_th, g_start, g_goal, g_qc, g_w, g_eps,
                             g_sdf, p->sdf_stride_b, st);
  return launch_bwd<3, IO>(k, kw, th, start, goal, sdf, dth, g_dth, g_err_ext, g_th, g_start, g_goal, g_qc, g_w, g_eps,
                           g_sdf, p->sdf_stride_b, st);
}

int grid_for(long long n, int threads) {
  long long g = (n + threads - 1) / threads;
  const long long cap = 148LL * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

template <typename IO>
int errors_impl(const dgpmp2_params* p, const IO* th, const IO* start, const IO* goal, const IO* sdf,
                const dgpmp2_weights* w, IO* err, IO* err_ext, IO* err_sg, IO* err_gp, IO* err_obs, void* stream) {
  int rc = check_params(p, w);
  if (rc != DGPMP2_OK) return rc;
  if (p->B == 0) return DGPMP2_OK;
  if (!th || !start || !goal || !sdf) return DGPMP2_ERR_ARG;
  KParams k = make_kparams(p);
  const KWeights<IO> kw = make_kweights<IO>(w);
  finish_kparams(k, kw);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int threads = ((p->T + 31) / 32) * 32;
  if (threads > 256) threads = 256;
  if (p->dof == 2)
    errors_kernel<2, IO><<<p->B, threads, 0, st>>>(k, kw, th, start, goal, sdf, err, err_ext, err_sg, err_gp, err_obs);
  else
    errors_kernel<3, IO><<<p->B, threads, 0, st>>>(k, kw, th, start, goal, sdf, err, err_ext, err_sg, err_gp, err_obs);
  CUDA_TRY(cudaGetLastError());
  return DGPMP2_OK;
}


template <typename IO>
int errors_backward_impl(const dgpmp2_params* p, const IO* th, const IO* start, const IO* goal, const IO* sdf,
                         const dgpmp2_weights* w, const IO* g_ext, const IO* g_sg, const IO* g_gp, const IO* g_obs,
                         IO* g_th, void* stream) {
  dgpmp2_params q = *p;
  q.flags &= ~DGPMP2_FLAG_Q_FULL;   // only the constructor-time GP covariance (err_ext) and eps are read
  int rc = check_params(&q, w);
  if (rc != DGPMP2_OK) return rc;
  if (p->B == 0) return DGPMP2_OK;
  if (!th || !start || !goal || !sdf || !g_th) return DGPMP2_ERR_ARG;
  const KParams k = make_kparams(&q);
  const KWeights<IO> kw = make_kweights<IO>(w);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int g = grid_for((long long)p->B * p->T, 256);
  if (p->dof == 2) errors_bwd_kernel<2, IO><<<g, 256, 0, st>>>(k, kw, th, start, goal, sdf, g_ext, g_sg, g_gp, g_obs, g_th);
  else errors_bwd_kernel<3, IO><<<g, 256, 0, st>>>(k, kw, th, start, goal, sdf, g_ext, g_sg, g_gp, g_obs, g_th);
  CUDA_TRY(cudaGetLastError());
  return DGPMP2_OK;
}

template <typename IO>
int factors_impl(const dgpmp2_params* p, const IO* th, const IO* sdf, const dgpmp2_weights* w, IO* gp_err,
                 IO* obs_cost, IO* obs_H, IO* cust_err, IO* cust_H, void* stream) {
  dgpmp2_params q = *p;
  q.flags &= ~DGPMP2_FLAG_Q_FULL;   // GP covariance is not read here
  int rc = check_params(&q, w);
  if (rc != DGPMP2_OK) return rc;
  if (p->B == 0) return DGPMP2_OK;
  if (!th) return DGPMP2_ERR_ARG;
  if ((obs_cost || obs_H) && !sdf) return DGPMP2_ERR_ARG;
  const KParams k = make_kparams(&q);
  const KWeights<IO> kw = make_kweights<IO>(w);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int g = grid_for((long long)p->B * p->T, 256);
  if (gp_err == nullptr && cust_err == nullptr && cust_H == nullptr) {   // obstacle factor only: streaming fast path
    if (p->dof == 2) obstacle_kernel<2, IO><<<g, 256, 0, st>>>(k, kw, th, sdf, obs_cost, obs_H);
    else obstacle_kernel<3, IO><<<g, 256, 0, st>>>(k, kw, th, sdf, obs_cost, obs_H);
    CUDA_TRY(cudaGetLastError());
    return DGPMP2_OK;
  }
  if (p->dof == 2)
    factors_kernel<2, IO><<<g, 256, 0, st>>>(k, kw, th, sdf, gp_err, obs_cost, obs_H, cust_err, cust_H);
  else
    factors_kernel<3, IO><<<g, 256, 0, st>>>(k, kw, th, sdf, gp_err, obs_cost, obs_H, cust_err, cust_H);
  CUDA_TRY(cudaGetLastError());
  return DGPMP2_OK;
}

template <typename IO>
int band_impl(const dgpmp2_params* p, const IO* th, const IO* start, const IO* goal, const IO* sdf,
              const dgpmp2_weights* w, double* D, double* U, double* r, void* stream) {
  int rc = check_params(p, w);
  if (rc != DGPMP2_OK) return rc;
  if (p->B == 0) return DGPMP2_OK;
  if (!th || !start || !goal || !sdf || !D || !U || !r) return DGPMP2_ERR_ARG;
  KParams k = make_kparams(p);
  const KWeights<IO> kw = make_kweights<IO>(w);
  finish_kparams(k, kw);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int g = grid_for((long long)p->B * p->T, 128);
  if (p->dof == 2) band_kernel<2, IO><<<g, 128, 0, st>>>(k, kw, th, start, goal, sdf, D, U, r);
  else band_kernel<3, IO><<<g, 128, 0, st>>>(k, kw, th, start, goal, sdf, D, U, r);
  CUDA_TRY(cudaGetLastError());
  return DGPMP2_OK;
}

template <typename IO>
int sdf_lookup_impl(const IO* sdf, int32_t B, int32_t H, int32_t W, int64_t sdf_sb, const IO* pts, int32_t N,
                    double res, double x_lo, double y_lo, IO* dist, IO* J, void* stream) {
  if (B < 0 || N < 0 || H < 1 || W < 1 || !(res > 0.0) || sdf_sb < 0) return DGPMP2_ERR_ARG;
  if (B == 0 || N == 0) return DGPMP2_OK;
  if (!sdf || !pts) return DGPMP2_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const double ox = 0.0 - x_lo / res, oy = 0.0 - y_lo / res;
  sdf_lookup_kernel<IO><<<grid_for((long long)B * N, 256), 256, 0, st>>>(sdf, B, H, W, sdf_sb, pts, N, res, ox, oy, dist, J);
  CUDA_TRY(cudaGetLastError());
  return DGPMP2_OK;
}

template <typename IO>
int hinge_batch_impl(const IO* sdf, int32_t B, int32_t H, int32_t W, int64_t sdf_sb, const IO* pts, int32_t N, double res,
                     double x_lo, double y_lo, const IO* eps, int64_t eps_sb, int64_t eps_sn, double eps_const,
                     double r_sphere, IO* cost, IO* He, void* stream) {
  if (B < 0 || N < 0 || H < 1 || W < 1 || !(res > 0.0) || sdf_sb < 0) return DGPMP2_ERR_ARG;
  if (B == 0 || N == 0) return DGPMP2_OK;
  if (!sdf || !pts || !cost || !He) return DGPMP2_ERR_ARG;
  if (reinterpret_cast<unsigned long long>(pts) & (2 * sizeof(IO) - 1)) return DGPMP2_ERR_ARG;   // (x, y) pairs are loaded as one vector
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const double ox = 0.0 - x_lo / res, oy = 0.0 - y_lo / res;
  const long long n = (long long)B * N;
  const long long blocks = (n + 256 * DGPMP2_K1_NPT - 1) / (256 * DGPMP2_K1_NPT);
  if (blocks > 0x7fffffffLL) return DGPMP2_ERR_UNSUPPORTED;
  hinge_kernel<IO><<<(unsigned)blocks, 256, 0, st>>>(sdf, B, H, W, sdf_sb, pts, N, res, 1.0 / res, ox, oy, eps, eps_sb, eps_sn,
                                                     eps_const, r_sphere, cost, He);
  CUDA_TRY(cudaGetLastError());
  return DGPMP2_OK;
}

template <typename IN, typename IO>
int sdf_from_occupancy_impl(const IN* im, int32_t B, int32_t H, int32_t W, int32_t pad, double thresh, double res, IO* out,
                            void* stream) {
  if (B < 0 || H < 1 || W < 1 || pad < 0 || !(res > 0.0)) return DGPMP2_ERR_ARG;
  if (B == 0) return DGPMP2_OK;
  if (!im || !out) return DGPMP2_ERR_ARG;
  const size_t Hp = (size_t)H + 2 * pad, Wp = (size_t)W + 2 * pad;
  if (Hp > 254 || Wp > 254) return DGPMP2_ERR_UNSUPPORTED;          // one-byte row distances, 255 = none
  if (!std::is_arithmetic<IN>::value && pad != 0) return DGPMP2_ERR_UNSUPPORTED;   // bit-packed maps: padlen 0 only
  // one thread per (column, polarity) envelope: 256 threads when their stacks fit beside the row distances, else fewer
  int threads = 256;
  while (threads > 32 && EdtSmem::bytes((int)Hp, (int)Wp, threads) > (size_t)kSmemLimit) threads >>= 1;
  const size_t bytes = EdtSmem::bytes((int)Hp, (int)Wp, threads);
  if (bytes > (size_t)kSmemLimit) return DGPMP2_ERR_UNSUPPORTED;
  auto kern = sdf_from_occupancy_kernel<IN, IO>;
  int rc = allow_smem(kern, (int)bytes);
  if (rc != DGPMP2_OK) return rc;
  kern<<<B, threads, bytes, static_cast<cudaStream_t>(stream)>>>(im, H, W, pad, thresh, res, out);
  CUDA_TRY(cudaGetLastError());
  return DGPMP2_OK;
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct HostWs { size_t th, start, goal, sdf, dth, err, err_ext, status, total; };

HostWs host_ws_layout(const dgpmp2_params* p, size_t es) {
  const size_t B = p->B, T = p->T, d = 2 * p->dof;
  const size_t sdf_elems = (p->sdf_stride_b == 0) ? (size_t)p->H * p->W : (size_t)p->sdf_stride_b * B;
  HostWs w;
  size_t o = 0;
  w.th = o; o += align256(B * T * d * es);
  w.start = o; o += align256(B * d * es);
  w.goal = o; o += align256(B * d * es);
  w.sdf = o; o += align256(sdf_elems * es);
  w.dth = o; o += align256(B * T * d * es);
  w.err = o; o += align256(B * es);
  w.err_ext = o; o += align256(B * es);
  w.status = o; o += align256(B * 4);
  w.total = o;
  return w;
}

// Two helper streams per device for the chunked host step (created once, kept for the life of the process).
struct HostPipe { cudaStream_t s[2]; cudaEvent_t ready, done[2]; bool ok; std::mutex busy; };

HostPipe* host_pipe() {
  static std::mutex mu;
  static HostPipe pipes[64];
  static bool made[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lk(mu);
  if (!made[dev]) {
    HostPipe& h = pipes[dev];
    h.ok = cudaStreamCreateWithFlags(&h.s[0], cudaStreamNonBlocking) == cudaSuccess &&
           cudaStreamCreateWithFlags(&h.s[1], cudaStreamNonBlocking) == cudaSuccess &&
           cudaEventCreateWithFlags(&h.ready, cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&h.done[0], cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&h.done[1], cudaEventDisableTiming) == cudaSuccess;
    made[dev] = true;
  }
  return pipes[dev].ok ? &pipes[dev] : nullptr;
}

// Device-side alias of a host buffer that is pinned and mapped (cudaHostAlloc / cudaHostRegister under unified addressing),
// nullptr for pageable memory.
template <typename IO>
const IO* mapped_alias(const IO* host) {
  if (host == nullptr) return nullptr;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, host) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  if (at.type != cudaMemoryTypeHost || at.devicePointer == nullptr) return nullptr;
  return static_cast<const IO*>(at.devicePointer);
}

template <typename IO>
int gn_step_host_impl(const dgpmp2_params* p, const IO* th, const IO* start, const IO* goal, const IO* sdf, IO* dth,
                      IO* err, IO* err_ext, int32_t* status, void* dev_ws, size_t dev_ws_bytes, int32_t sdf_resident,
                      void* stream) {
  int rc = check_params(p, nullptr);
  if (rc != DGPMP2_OK) return rc;
  if (p->B == 0) return DGPMP2_OK;
  if (!th || !start || !goal || !dth || !err || !err_ext || !dev_ws) return DGPMP2_ERR_ARG;
  if (sdf_resident < 0 || sdf_resident > 2) return DGPMP2_ERR_ARG;
  if (sdf_resident != DGPMP2_SDF_RESIDENT && !sdf) return DGPMP2_ERR_ARG;
  const HostWs L = host_ws_layout(p, sizeof(IO));
  if (dev_ws_bytes < L.total) return DGPMP2_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned char* ws = static_cast<unsigned char*>(dev_ws);
  const size_t B = p->B, T = p->T, d = 2 * p->dof;
  const size_t sdf_elems = (p->sdf_stride_b == 0) ? (size_t)p->H * p->W : (size_t)p->sdf_stride_b * B;
  // Chunked pipeline (per-problem SDFs copied every step, enough problems): the batch is cut into chunks that alternate
  // between two helper streams, so chunk c's kernel and device-to-host copies overlap chunk c+1's host-to-device copy
  // (the two directions use different copy engines).  A problem's result does not depend on the batch it is launched
  // in, so the output is bit-identical to the one-launch path.  Opt-in (DGPMP2_HOST_CHUNKS=n > 1): measured on B200 it
  // does not pay -- the 64 MiB SDF copy alone saturates PCIe (1024 problems: 763 k problem-iters/s in one piece, 741 k
  // in 4 chunks, 683 k in 8; profiles/r02_host_chunks_ab.txt), the kernel and the 1 MB of results are < 3 % of the step.
  const int want_chunks = env_int("DGPMP2_HOST_CHUNKS", 1);
  HostPipe* hp = (want_chunks > 1 && sdf_resident == DGPMP2_SDF_COPY && p->sdf_stride_b > 0 && B >= 256) ? host_pipe() : nullptr;
  if (hp != nullptr) {
    std::lock_guard<std::mutex> in_use(hp->busy);      // the helper streams / events serve one (synchronous) call at a time
    const size_t nch = (size_t)want_chunks;
    CUDA_TRY(cudaEventRecord(hp->ready, st));
    for (int k = 0; k < 2; ++k) CUDA_TRY(cudaStreamWaitEvent(hp->s[k], hp->ready, 0));
    for (size_t c = 0; c < nch; ++c) {
      const size_t lo = B * c / nch, hi = B * (c + 1) / nch, nb = hi - lo;
      if (nb == 0) continue;
      cudaStream_t cs = hp->s[c & 1];
      IO* d_th = reinterpret_cast<IO*>(ws + L.th) + lo * T * d;
      IO* d_start = reinterpret_cast<IO*>(ws + L.start) + lo * d;
      IO* d_goal = reinterpret_cast<IO*>(ws + L.goal) + lo * d;
      IO* d_sdf = reinterpret_cast<IO*>(ws + L.sdf) + lo * (size_t)p->sdf_stride_b;
      IO* d_dth = reinterpret_cast<IO*>(ws + L.dth) + lo * T * d;
      IO* d_err = reinterpret_cast<IO*>(ws + L.err) + lo;
      IO* d_ee = reinterpret_cast<IO*>(ws + L.err_ext) + lo;
      int32_t* d_st = reinterpret_cast<int32_t*>(ws + L.status) + lo;
      CUDA_TRY(cudaMemcpyAsync(d_th, th + lo * T * d, nb * T * d * sizeof(IO), cudaMemcpyHostToDevice, cs));
      CUDA_TRY(cudaMemcpyAsync(d_start, start + lo * d, nb * d * sizeof(IO), cudaMemcpyHostToDevice, cs));
      CUDA_TRY(cudaMemcpyAsync(d_goal, goal + lo * d, nb * d * sizeof(IO), cudaMemcpyHostToDevice, cs));
      CUDA_TRY(cudaMemcpyAsync(d_sdf, sdf + lo * (size_t)p->sdf_stride_b, nb * (size_t)p->sdf_stride_b * sizeof(IO),
                               cudaMemcpyHostToDevice, cs));
      dgpmp2_params q = *p;
      q.B = (int32_t)nb;
      rc = gn_step_impl<IO>(&q, d_th, d_start, d_goal, d_sdf, nullptr, d_dth, d_err, d_ee, d_st, cs);
      if (rc != DGPMP2_OK) return rc;
      CUDA_TRY(cudaMemcpyAsync(dth + lo * T * d, d_dth, nb * T * d * sizeof(IO), cudaMemcpyDeviceToHost, cs));
      CUDA_TRY(cudaMemcpyAsync(err + lo, d_err, nb * sizeof(IO), cudaMemcpyDeviceToHost, cs));
      CUDA_TRY(cudaMemcpyAsync(err_ext + lo, d_ee, nb * sizeof(IO), cudaMemcpyDeviceToHost, cs));
      if (status) CUDA_TRY(cudaMemcpyAsync(status + lo, d_st, nb * 4, cudaMemcpyDeviceToHost, cs));
    }
    for (int k = 0; k < 2; ++k) {
      CUDA_TRY(cudaEventRecord(hp->done[k], hp->s[k]));
      CUDA_TRY(cudaStreamWaitEvent(st, hp->done[k], 0));
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    return DGPMP2_OK;
  }
  // Where each operand lives for the launch: the workspace copy, or -- mode DGPMP2_SDF_IN_PLACE -- the caller's own
  // buffer when it is pinned (mapped) host memory.  The SDF is the case that matters: a step reads 4 taps per state, so
  // the kernel pulls ~0.15 M 32-byte sectors over PCIe instead of the whole 64 MiB field crossing it first
  // (measured, B=1024, T=64, 128x128 maps: 0.32 ms per step instead of 1.42 ms; same bits).
  const bool in_place = (sdf_resident == DGPMP2_SDF_IN_PLACE);
  // The small operands (trajectories in, dtheta / errors / status out) are read / written in place in every mode when
  // they are pinned (DGPMP2_HOST_INPLACE_IO=2 copies them instead, A/B): with every operand in pinned memory the call is
  // one kernel launch and one stream synchronisation (2.57 -> 3.13 M problem-iters/s with DGPMP2_SDF_IN_PLACE).
  const bool io_in_place = env_int("DGPMP2_HOST_INPLACE_IO", 1) == 1;
  const IO* k_sdf = reinterpret_cast<const IO*>(ws + L.sdf);
  bool copy_sdf = (sdf_resident == DGPMP2_SDF_COPY);
  if (in_place) {
    const IO* alias = mapped_alias(sdf);
    if (alias != nullptr) k_sdf = alias; else copy_sdf = true;
  }
  const IO* a_th = io_in_place ? mapped_alias(th) : nullptr;
  const IO* a_start = io_in_place ? mapped_alias(start) : nullptr;
  const IO* a_goal = io_in_place ? mapped_alias(goal) : nullptr;
  IO* a_dth = io_in_place ? const_cast<IO*>(mapped_alias(dth)) : nullptr;
  IO* a_err = io_in_place ? const_cast<IO*>(mapped_alias(err)) : nullptr;
  IO* a_ee = io_in_place ? const_cast<IO*>(mapped_alias(err_ext)) : nullptr;
  int32_t* a_status = (io_in_place && status) ? const_cast<int32_t*>(mapped_alias(status)) : nullptr;
  if (!a_th) CUDA_TRY(cudaMemcpyAsync(ws + L.th, th, B * T * d * sizeof(IO), cudaMemcpyHostToDevice, st));
  if (!a_start) CUDA_TRY(cudaMemcpyAsync(ws + L.start, start, B * d * sizeof(IO), cudaMemcpyHostToDevice, st));
  if (!a_goal) CUDA_TRY(cudaMemcpyAsync(ws + L.goal, goal, B * d * sizeof(IO), cudaMemcpyHostToDevice, st));
  if (copy_sdf) CUDA_TRY(cudaMemcpyAsync(ws + L.sdf, sdf, sdf_elems * sizeof(IO), cudaMemcpyHostToDevice, st));
  rc = gn_step_impl<IO>(p, a_th ? a_th : reinterpret_cast<const IO*>(ws + L.th),
                        a_start ? a_start : reinterpret_cast<const IO*>(ws + L.start),
                        a_goal ? a_goal : reinterpret_cast<const IO*>(ws + L.goal), k_sdf, nullptr,
                        a_dth ? a_dth : reinterpret_cast<IO*>(ws + L.dth), a_err ? a_err : reinterpret_cast<IO*>(ws + L.err),
                        a_ee ? a_ee : reinterpret_cast<IO*>(ws + L.err_ext),
                        a_status ? a_status : reinterpret_cast<int32_t*>(ws + L.status), stream);
  if (rc != DGPMP2_OK) return rc;
  if (!a_dth) CUDA_TRY(cudaMemcpyAsync(dth, ws + L.dth, B * T * d * sizeof(IO), cudaMemcpyDeviceToHost, st));
  if (!a_err) CUDA_TRY(cudaMemcpyAsync(err, ws + L.err, B * sizeof(IO), cudaMemcpyDeviceToHost, st));
  if (!a_ee) CUDA_TRY(cudaMemcpyAsync(err_ext, ws + L.err_ext, B * sizeof(IO), cudaMemcpyDeviceToHost, st));
  if (status && !a_status) CUDA_TRY(cudaMemcpyAsync(status, ws + L.status, B * 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return DGPMP2_OK;
}

// End to end from bit-packed occupancy maps: H2D of the trajectories and of B*H*ceil(W/32) words of map bits (1/32 of
// the float SDF), exact EDT on the device into the workspace's SDF region, GN step, D2H of the results.
size_t occ_words(const dgpmp2_params* p) { return (size_t)p->B * p->H * ((size_t)(p->W + 31) / 32); }

int gn_step_host_occ_impl(const dgpmp2_params* p, const float* th, const float* start, const float* goal,
                          const uint32_t* occ_bits, float* dth, float* err, float* err_ext, int32_t* status, void* dev_ws,
                          size_t dev_ws_bytes, void* stream) {
  int rc = check_params(p, nullptr);
  if (rc != DGPMP2_OK) return rc;
  if (p->B == 0) return DGPMP2_OK;
  if (!th || !start || !goal || !occ_bits || !dth || !err || !err_ext || !dev_ws) return DGPMP2_ERR_ARG;
  if (p->sdf_stride_b != (int64_t)p->H * p->W) return DGPMP2_ERR_ARG;          // one map per problem
  const HostWs L = host_ws_layout(p, sizeof(float));
  const size_t bits_bytes = occ_words(p) * 4;
  if (dev_ws_bytes < L.total + align256(bits_bytes)) return DGPMP2_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned char* ws = static_cast<unsigned char*>(dev_ws);
  const size_t B = p->B, T = p->T, d = 2 * p->dof;
  CUDA_TRY(cudaMemcpyAsync(ws + L.total, occ_bits, bits_bytes, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(ws + L.th, th, B * T * d * sizeof(float), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(ws + L.start, start, B * d * sizeof(float), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(ws + L.goal, goal, B * d * sizeof(float), cudaMemcpyHostToDevice, st));
  rc = sdf_from_occupancy_impl<OccBits, float>(reinterpret_cast<const OccBits*>(ws + L.total), p->B, p->H, p->W, 0, 0.0, p->res,
                                               reinterpret_cast<float*>(ws + L.sdf), stream);
  if (rc != DGPMP2_OK) return rc;
  rc = gn_step_impl<float>(p, reinterpret_cast<float*>(ws + L.th), reinterpret_cast<float*>(ws + L.start),
                           reinterpret_cast<float*>(ws + L.goal), reinterpret_cast<float*>(ws + L.sdf), nullptr,
                           reinterpret_cast<float*>(ws + L.dth), reinterpret_cast<float*>(ws + L.err),
                           reinterpret_cast<float*>(ws + L.err_ext), reinterpret_cast<int32_t*>(ws + L.status), stream);
  if (rc != DGPMP2_OK) return rc;
  CUDA_TRY(cudaMemcpyAsync(dth, ws + L.dth, B * T * d * sizeof(float), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(err, ws + L.err, B * sizeof(float), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(err_ext, ws + L.err_ext, B * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (status) CUDA_TRY(cudaMemcpyAsync(status, ws + L.status, B * 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return DGPMP2_OK;
}

}  // namespace

#ifdef DGPMP2_MP_TIMING
extern "C" int dgpmp2_debug_mp_clocks(long long* out) {
  return cudaMemcpyFromSymbol(out, dgpmp2::g_mp_clock, sizeof(long long) * 64) == cudaSuccess ? 0 : -3;
}
#endif

#ifdef DGPMP2_TIMING
extern "C" int dgpmp2_debug_phase_clocks(long long* out) {
  return cudaMemcpyFromSymbol(out, dgpmp2::g_phase_clock, sizeof(long long) * 64) == cudaSuccess ? 0 : -3;
}
#endif

extern "C" {

int dgpmp2_abi_version(void) { return DGPMP2_ABI_VERSION; }

int dgpmp2_host_pointer_is_mapped(const void* host_ptr) {
  return mapped_alias(static_cast<const unsigned char*>(host_ptr)) != nullptr ? 1 : 0;
}

const char* dgpmp2_status_string(int code) {
  switch (code) {
    case DGPMP2_OK: return "ok";
    case DGPMP2_ERR_ARG: return "invalid argument";
    case DGPMP2_ERR_UNSUPPORTED: return "unsupported configuration (dof not in {2,3} or trajectory too long for on-chip band)";
    case DGPMP2_ERR_CUDA: return "CUDA runtime error";
    default: return "unknown status";
  }
}

const char* dgpmp2_last_cuda_error(void) { return g_cuda_err; }

int dgpmp2_gn_step_f32(const dgpmp2_params* p, const float* th, const float* start, const float* goal, const float* sdf,
                       const dgpmp2_weights* w, float* dth, float* err, float* err_ext, int32_t* status, void* stream) {
  return gn_step_impl<float>(p, th, start, goal, sdf, w, dth, err, err_ext, status, stream);
}
int dgpmp2_gn_step_f64(const dgpmp2_params* p, const double* th, const double* start, const double* goal,
                       const double* sdf, const dgpmp2_weights* w, double* dth, double* err, double* err_ext,
                       int32_t* status, void* stream) {
  return gn_step_impl<double>(p, th, start, goal, sdf, w, dth, err, err_ext, status, stream);
}
int dgpmp2_gn_step_diag_f32(const dgpmp2_params* p, const float* th, const float* start, const float* goal,
                            const float* sdf, const dgpmp2_weights* w, float* dth, float* err, float* err_ext,
                            int32_t* status, int32_t* refine, void* stream) {
  return gn_step_impl<float>(p, th, start, goal, sdf, w, dth, err, err_ext, status, stream, refine);
}

int dgpmp2_gn_step_backward_f32(const dgpmp2_params* p, const float* th, const float* start, const float* goal,
                                const float* sdf, const dgpmp2_weights* w, const float* dth, const float* g_dth,
                                const float* g_err_ext, float* g_th, float* g_start, float* g_goal, float* g_qc,
                                float* g_w, float* g_eps, float* g_sdf, void* stream) {
  return gn_step_backward_impl<float>(p, th, start, goal, sdf, w, dth, g_dth, g_err_ext, g_th, g_start, g_goal, g_qc, g_w,
                                      g_eps, g_sdf, stream);
}
int dgpmp2_gn_step_backward_f64(const dgpmp2_params* p, const double* th, const double* start, const double* goal,
                                const double* sdf, const dgpmp2_weights* w, const double* dth, const double* g_dth,
                                const double* g_err_ext, double* g_th, double* g_start, double* g_goal, double* g_qc,
                                double* g_w, double* g_eps, double* g_sdf, void* stream) {
  return gn_step_backward_impl<double>(p, th, start, goal, sdf, w, dth, g_dth, g_err_ext, g_th, g_start, g_goal, g_qc, g_w,
                                       g_eps, g_sdf, stream);
}

int dgpmp2_gn_solve_f32(const dgpmp2_params* p, const float* th_init, const float* start, const float* goal,
                        const float* sdf, const dgpmp2_weights* w, int32_t max_iters, double tol_delta, float* th_final,
                        int32_t* iters, float* err_per_iter, float* err_ext_per_iter, float* err_final,
                        float* err_ext_final, int32_t* status, void* stream) {
  return gn_solve_impl<float>(p, th_init, start, goal, sdf, w, max_iters, tol_delta, th_final, iters, err_per_iter,
                              err_ext_per_iter, err_final, err_ext_final, status, stream);
}
int dgpmp2_gn_solve_f64(const dgpmp2_params* p, const double* th_init, const double* start, const double* goal,
                        const double* sdf, const dgpmp2_weights* w, int32_t max_iters, double tol_delta,
                        double* th_final, int32_t* iters, double* err_per_iter, double* err_ext_per_iter,
                        double* err_final, double* err_ext_final, int32_t* status, void* stream) {
  return gn_solve_impl<double>(p, th_init, start, goal, sdf, w, max_iters, tol_delta, th_final, iters, err_per_iter,
                               err_ext_per_iter, err_final, err_ext_final, status, stream);
}

int dgpmp2_errors_f32(const dgpmp2_params* p, const float* th, const float* start, const float* goal, const float* sdf,
                      const dgpmp2_weights* w, float* err, float* err_ext, float* err_sg, float* err_gp, float* err_obs,
                      void* stream) {
  return errors_impl<float>(p, th, start, goal, sdf, w, err, err_ext, err_sg, err_gp, err_obs, stream);
}
int dgpmp2_errors_f64(const dgpmp2_params* p, const double* th, const double* start, const double* goal,
                      const double* sdf, const dgpmp2_weights* w, double* err, double* err_ext, double* err_sg,
                      double* err_gp, double* err_obs, void* stream) {
  return errors_impl<double>(p, th, start, goal, sdf, w, err, err_ext, err_sg, err_gp, err_obs, stream);
}

int dgpmp2_errors_backward_f32(const dgpmp2_params* p, const float* th, const float* start, const float* goal,
                               const float* sdf, const dgpmp2_weights* w, const float* g_err_ext, const float* g_err_sg,
                               const float* g_err_gp, const float* g_err_obs, float* g_th, void* stream) {
  if (p == nullptr) return DGPMP2_ERR_ARG;
  return errors_backward_impl<float>(p, th, start, goal, sdf, w, g_err_ext, g_err_sg, g_err_gp, g_err_obs, g_th, stream);
}
int dgpmp2_errors_backward_f64(const dgpmp2_params* p, const double* th, const double* start, const double* goal,
                               const double* sdf, const dgpmp2_weights* w, const double* g_err_ext,
                               const double* g_err_sg, const double* g_err_gp, const double* g_err_obs, double* g_th,
                               void* stream) {
  if (p == nullptr) return DGPMP2_ERR_ARG;
  return errors_backward_impl<double>(p, th, start, goal, sdf, w, g_err_ext, g_err_sg, g_err_gp, g_err_obs, g_th, stream);
}

int dgpmp2_factors_f32(const dgpmp2_params* p, const float* th, const float* sdf, const dgpmp2_weights* w, float* gp_err,
                       float* obs_cost, float* obs_H, float* cust_err, float* cust_H, void* stream) {
  if (p == nullptr) return DGPMP2_ERR_ARG;
  return factors_impl<float>(p, th, sdf, w, gp_err, obs_cost, obs_H, cust_err, cust_H, stream);
}
int dgpmp2_factors_f64(const dgpmp2_params* p, const double* th, const double* sdf, const dgpmp2_weights* w,
                       double* gp_err, double* obs_cost, double* obs_H, double* cust_err, double* cust_H, void* stream) {
  if (p == nullptr) return DGPMP2_ERR_ARG;
  return factors_impl<double>(p, th, sdf, w, gp_err, obs_cost, obs_H, cust_err, cust_H, stream);
}

int dgpmp2_sdf_lookup_f32(const float* sdf, int32_t B, int32_t H, int32_t W, int64_t sdf_stride_b, const float* pts,
                          int32_t N, double res, double x_lo, double y_lo, float* dist, float* J, void* stream) {
  return sdf_lookup_impl<float>(sdf, B, H, W, sdf_stride_b, pts, N, res, x_lo, y_lo, dist, J, stream);
}
int dgpmp2_sdf_lookup_f64(const double* sdf, int32_t B, int32_t H, int32_t W, int64_t sdf_stride_b, const double* pts,
                          int32_t N, double res, double x_lo, double y_lo, double* dist, double* J, void* stream) {
  return sdf_lookup_impl<double>(sdf, B, H, W, sdf_stride_b, pts, N, res, x_lo, y_lo, dist, J, stream);
}

int dgpmp2_hinge_batch_f32(const float* sdf, int32_t B, int32_t H, int32_t W, int64_t sdf_stride_b, const float* pts,
                           int32_t N, double res, double x_lo, double y_lo, const float* eps, int64_t eps_stride_b,
                           int64_t eps_stride_n, double eps_const, double r_sphere, float* cost, float* He, void* stream) {
  return hinge_batch_impl<float>(sdf, B, H, W, sdf_stride_b, pts, N, res, x_lo, y_lo, eps, eps_stride_b, eps_stride_n,
                                 eps_const, r_sphere, cost, He, stream);
}
int dgpmp2_hinge_batch_f64(const double* sdf, int32_t B, int32_t H, int32_t W, int64_t sdf_stride_b, const double* pts,
                           int32_t N, double res, double x_lo, double y_lo, const double* eps, int64_t eps_stride_b,
                           int64_t eps_stride_n, double eps_const, double r_sphere, double* cost, double* He, void* stream) {
  return hinge_batch_impl<double>(sdf, B, H, W, sdf_stride_b, pts, N, res, x_lo, y_lo, eps, eps_stride_b, eps_stride_n,
                                  eps_const, r_sphere, cost, He, stream);
}

int dgpmp2_sdf_from_occupancy_f32(const float* im, int32_t B, int32_t H, int32_t W, int32_t padlen, double thresh,
                                  double res, float* sdf_out, void* stream) {
  return sdf_from_occupancy_impl<float, float>(im, B, H, W, padlen, thresh, res, sdf_out, stream);
}
int dgpmp2_sdf_from_occupancy_f64(const double* im, int32_t B, int32_t H, int32_t W, int32_t padlen, double thresh,
                                  double res, double* sdf_out, void* stream) {
  return sdf_from_occupancy_impl<double, double>(im, B, H, W, padlen, thresh, res, sdf_out, stream);
}
int dgpmp2_sdf_from_occupancy_u8_f32(const uint8_t* im, int32_t B, int32_t H, int32_t W, int32_t padlen, double thresh,
                                     double res, float* sdf_out, void* stream) {
  return sdf_from_occupancy_impl<uint8_t, float>(im, B, H, W, padlen, thresh, res, sdf_out, stream);
}

int dgpmp2_sdf_from_occupancy_bits_f32(const uint32_t* im_bits, int32_t B, int32_t H, int32_t W, double res, float* sdf_out,
                                       void* stream) {
  return sdf_from_occupancy_impl<OccBits, float>(reinterpret_cast<const OccBits*>(im_bits), B, H, W, 0, 0.0, res, sdf_out, stream);
}
int dgpmp2_host_step_occ_workspace_bytes(const dgpmp2_params* p, size_t* bytes) {
  if (p == nullptr || bytes == nullptr) return DGPMP2_ERR_ARG;
  int rc = check_params(p, nullptr);
  if (rc != DGPMP2_OK) return rc;
  *bytes = host_ws_layout(p, sizeof(float)).total + align256(occ_words(p) * 4);
  return DGPMP2_OK;
}
int dgpmp2_gn_step_host_occ_f32(const dgpmp2_params* p, const float* th, const float* start, const float* goal,
                                const uint32_t* occ_bits, float* dth, float* err, float* err_ext, int32_t* status,
                                void* dev_ws, size_t dev_ws_bytes, void* stream) {
  return gn_step_host_occ_impl(p, th, start, goal, occ_bits, dth, err, err_ext, status, dev_ws, dev_ws_bytes, stream);
}

int dgpmp2_band_f32(const dgpmp2_params* p, const float* th, const float* start, const float* goal, const float* sdf,
                    const dgpmp2_weights* w, double* D, double* U, double* r, void* stream) {
  return band_impl<float>(p, th, start, goal, sdf, w, D, U, r, stream);
}
int dgpmp2_band_f64(const dgpmp2_params* p, const double* th, const double* start, const double* goal,
                    const double* sdf, const dgpmp2_weights* w, double* D, double* U, double* r, void* stream) {
  return band_impl<double>(p, th, start, goal, sdf, w, D, U, r, stream);
}

int dgpmp2_host_step_workspace_bytes(const dgpmp2_params* p, int32_t elem_size, size_t* bytes) {
  if (p == nullptr || bytes == nullptr || (elem_size != 4 && elem_size != 8)) return DGPMP2_ERR_ARG;
  int rc = check_params(p, nullptr);
  if (rc != DGPMP2_OK) return rc;
  *bytes = host_ws_layout(p, (size_t)elem_size).total;
  return DGPMP2_OK;
}
int dgpmp2_gn_step_host_f32(const dgpmp2_params* p, const float* th, const float* start, const float* goal,
                            const float* sdf, float* dth, float* err, float* err_ext, int32_t* status, void* dev_ws,
                            size_t dev_ws_bytes, int32_t sdf_resident, void* stream) {
  return gn_step_host_impl<float>(p, th, start, goal, sdf, dth, err, err_ext, status, dev_ws, dev_ws_bytes,
                                  sdf_resident, stream);
}
int dgpmp2_gn_step_host_f64(const dgpmp2_params* p, const double* th, const double* start, const double* goal,
                            const double* sdf, double* dth, double* err, double* err_ext, int32_t* status, void* dev_ws,
                            size_t dev_ws_bytes, int32_t sdf_resident, void* stream) {
  return gn_step_host_impl<double>(p, th, start, goal, sdf, dth, err, err_ext, status, dev_ws, dev_ws_bytes,
                                   sdf_resident, stream);
}

int dgpmp2_gn_step_launch_shape(const dgpmp2_params* p, int32_t elem_size, int32_t* problems_per_cta, int32_t* threads,
                                int32_t* smem_bytes, int32_t* grid) {
  if (p == nullptr || (elem_size != 4 && elem_size != 8)) return DGPMP2_ERR_ARG;
  dgpmp2_params q = *p;
  q.flags &= ~DGPMP2_FLAG_Q_FULL;
  int rc = check_params(&q, nullptr);
  if (rc != DGPMP2_OK) return rc;
  LaunchShape s{0, 0, 0, 0, 0};
  const int B = p->B > 0 ? p->B : 1;
  if (elem_size == 4 && mp_enabled()) {
    MpShape m;
    rc = (p->dof == 2) ? choose_shape_mp<4>(B, p->T, m) : choose_shape_mp<6>(B, p->T, m);
    if (rc == DGPMP2_OK) {
      if (problems_per_cta) *problems_per_cta = m.np;
      if (threads) *threads = m.threads;
      if (smem_bytes) *smem_bytes = m.smem;
      if (grid) *grid = m.grid;
      return DGPMP2_OK;
    }
  }
  if (p->dof == 2) rc = (elem_size == 4) ? choose_shape<4, float>(B, p->T, 0, s, true) : choose_shape<4, double>(B, p->T, 0, s, true);
  else rc = (elem_size == 4) ? choose_shape<6, float>(B, p->T, 0, s, true) : choose_shape<6, double>(B, p->T, 0, s, true);
  if (rc != DGPMP2_OK) return rc;
  if (problems_per_cta) *problems_per_cta = s.np;
  if (threads) *threads = s.threads;
  if (smem_bytes) *smem_bytes = s.smem;
  if (grid) *grid = s.grid;
  return DGPMP2_OK;
}

}  // extern "C"
