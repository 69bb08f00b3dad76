// TEST INFRASTRUCTURE -- host emulation of the mixed-precision GN step.
//
// Runs dgpmp2_b200/csrc/mp.cuh (the very source the GPU kernel gn_step_mp_kernel executes: it is written as
// __host__ __device__ code) with one CPU thread per trajectory state and a pthread barrier in place of the
// per-problem named barrier.  This box has no GPU; the emulator lets the CPU test-suite check the algorithm,
// its indexing and its synchronisation against the golden vectors of the live reference before any GPU time is
// spent.  It is NOT part of the product: nothing under dgpmp2_b200/ loads it, bench.py never calls it.
//
// Build (tests/test_mp_emu_cpu.py does this):  g++ -O1 -std=c++17 -shared -fPIC -pthread -I/usr/local/cuda/include
#include <pthread.h>

#include <algorithm>
#include <cstdio>
#include <mutex>
#include <thread>
#include <vector>

#include "../../dgpmp2_b200/csrc/host_params.h"
#include "../../dgpmp2_b200/csrc/mp.cuh"

using namespace dgpmp2;

namespace {

struct Shared {
  pthread_barrier_t bar;
  std::mutex mu;
  double sa = 0.0, sb = 0.0;
  int mx[4] = {0, 0, 0, 0};
  int need64 = 0;
};

struct HostCtx {
  Shared* sh;
  int m;
  void psync() { pthread_barrier_wait(&sh->bar); }
  void sum2(double& a, double& b) {
    {
      std::lock_guard<std::mutex> lk(sh->mu);
      sh->sa += a; sh->sb += b;
    }
    psync();
    a = sh->sa; b = sh->sb;
  }
  void max2(int& a, int& b, int it) {
    int* buf = sh->mx + 2 * (it & 1);
    {
      std::lock_guard<std::mutex> lk(sh->mu);
      buf[0] = std::max(buf[0], a); buf[1] = std::max(buf[1], b);
    }
    psync();
    a = buf[0]; b = buf[1];
    if (m == 0) { int* o = sh->mx + 2 * ((it + 1) & 1); o[0] = 0; o[1] = 0; }
  }
  void flag64() { sh->need64 = 1; }
  void stamp(int) {}
  int force_iters() const { const char* e = getenv("MP_EMU_ITERS"); return e ? atoi(e) : 0; }
  void trace(int b, int it, float nd, float nx) { if (getenv("MP_EMU_TRACE")) printf("trace b=%d it=%d nd/nx=%.3e\n", b, it, (double)(nd / nx)); }
};

template <int DOF>
int run(const dgpmp2_params* p, const float* th, const float* start, const float* goal, const float* sdf,
        const dgpmp2_weights* w, float* dth, float* err, float* err_ext, int* diag, int* need64, int force64) {
  constexpr int D = 2 * DOF;
  KParams k = make_kparams(p);
  const KWeights<float> kw = make_kweights<float>(w);
  finish_kparams(k, kw);
  const int T = k.T, TPP = (T + 31) / 32 * 32;
  for (int b = 0; b < k.B; ++b) {
    Shared sh;
    pthread_barrier_init(&sh.bar, nullptr, (unsigned)TPP);
    std::vector<float> recs_raw(MpRec<D>::problem_floats(T) + 8);
    float* recs = recs_raw.data();
    while (reinterpret_cast<uintptr_t>(recs) & 15u) ++recs;
    std::vector<std::thread> th_;
    for (int m = 0; m < TPP; ++m) {
      th_.emplace_back([&, m]() {
        HostCtx cx{&sh, m};
        if (DOF == 2)
          mp_thread_program<DOF, true, HostCtx>(cx, k, kw, b, m, m < T, th, start, goal, sdf, recs, dth, err, err_ext,
                                                diag, force64);
        else
          mp_thread_program<DOF, false, HostCtx>(cx, k, kw, b, m, m < T, th, start, goal, sdf, recs, dth, err, err_ext,
                                                 diag, force64);
      });
    }
    for (auto& t : th_) t.join();
    pthread_barrier_destroy(&sh.bar);
    if (need64) need64[b] = sh.need64;
  }
  return 0;
}

}  // namespace

extern "C" int mp_emu_step_f32(const dgpmp2_params* p, const float* th, const float* start, const float* goal,
                               const float* sdf, const dgpmp2_weights* w, float* dth, float* err, float* err_ext,
                               int* diag, int* need64, int force64) {
  if (p->dof == 2) return run<2>(p, th, start, goal, sdf, w, dth, err, err_ext, diag, need64, force64);
  if (p->dof == 3) return run<3>(p, th, start, goal, sdf, w, dth, err, err_ext, diag, need64, force64);
  return -2;
}
