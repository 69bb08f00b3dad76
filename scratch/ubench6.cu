// where does a block-Thomas forward step spend its cycles?  (copy of bcr_tail's forward loop with clock stamps)
#include <cstdio>
#include <cuda_runtime.h>
#include "../dgpmp2_b200/csrc/bcr.cuh"
using namespace dgpmp2;
#define FENCE(x) asm volatile("" :: "d"(x) : "memory")
__global__ void k(int T, int nc, long long* cyc, double* out) {
  extern __shared__ double sm[];
  using N = Node<4>;
  constexpr int D = 4, DS = 10, S = N::kStride;
  for (int i = threadIdx.x; i < T * S; i += blockDim.x) {
    const int f = i % S;
    double v = 0.01 * ((i * 7) % 13);
    if (f < 16 && (f / 4 == f % 4)) v += 8.0;
    sm[i] = v;
  }
  __syncthreads();
  const int lane = threadIdx.x % 4;
  if (threadIdx.x >= 4) return;
  long long acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  double gp[D] = {0, 0, 0, 0};
  const double* prev = sm;
#pragma unroll 1
  for (int e = 0; e < nc; ++e) {
    long long t0 = clock64();
    double* nd = sm + (size_t)bcr_slot(T, e) * S;
    double L[DS], r[D], vf[D];
    ld_lower<D>(nd + N::oD, L);
    ld_vec<D>(nd + N::oR, r);
#pragma unroll
    for (int a = 0; a < D; ++a) vf[a] = nd[N::oU + a * D + lane];
    double F[D][D];
#pragma unroll
    for (int c = 0; c < D; ++c) ld_vec<D>(prev + N::oU + c * D, F[c]);
    FENCE(L[0] + L[9] + r[3] + vf[3] + F[3][3] + F[0][0]);
    long long t1 = clock64();
    if (e > 0) {
#pragma unroll
      for (int a = 0; a < D; ++a) {
        r[a] = __dsub_rn(r[a], dot<D>(F[a], gp));
#pragma unroll
        for (int c = 0; c <= a; ++c) L[tri(a, c)] = __dsub_rn(L[tri(a, c)], dot<D>(F[a], F[c]));
      }
    }
    FENCE(L[0] + L[9] + r[3] + L[5]);
    long long t2 = clock64();
    __syncwarp(0xf);
    bool ok = chol_packed<D>(L);
    FENCE(L[9] + L[0] + (ok ? 1.0 : 0.0));
    long long t3 = clock64();
    fwd_solve<D>(L, r);
    fwd_solve<D>(L, vf);
    FENCE(r[3] + vf[3]);
    long long t4 = clock64();
    st_vec<D>(nd + N::oU + lane * D, vf);
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < DS; q += 2) sts2(nd + N::oD + q, L[q], L[q + 1]);
      st_vec<D>(nd + N::oR, r);
    }
#pragma unroll
    for (int a = 0; a < D; ++a) gp[a] = r[a];
    prev = nd;
    __syncwarp(0xf);
    long long t5 = clock64();
    acc[0] += t1 - t0; acc[1] += t2 - t1; acc[2] += t3 - t2; acc[3] += t4 - t3; acc[4] += t5 - t4;
  }
  if (threadIdx.x == 0) { for (int i = 0; i < 5; ++i) cyc[i] = acc[i]; out[0] = gp[0]; }
}
// pure dependent chains for reference
__global__ void chains(long long* cyc, double* out, double x) {
  long long t0 = clock64();
  double y = x;
#pragma unroll
  for (int i = 0; i < 64; ++i) y = fast_rsqrt(y + 1.5);
  FENCE(y);
  long long t1 = clock64();
  double z = x;
#pragma unroll
  for (int i = 0; i < 64; ++i) { double s; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(s) : "d"(z)); z = s + 1.5; }
  FENCE(z);
  long long t2 = clock64();
  double w = x;
#pragma unroll
  for (int i = 0; i < 64; ++i) w = __dmul_rn(w, 1.0000001);
  FENCE(w);
  long long t3 = clock64();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; out[0] = y + z + w; }
}
int main() {
  long long* cyc; double* out;
  cudaMalloc(&cyc, 128); cudaMalloc(&out, 64);
  const int T = 64, nc = 32;
  const int smem = T * Node<4>::kStride * 8;
  long long c[8];
  for (int it = 0; it < 2; ++it) { k<<<1, 32, smem>>>(T, nc, cyc, out); cudaDeviceSynchronize(); cudaMemcpy(c, cyc, 64, cudaMemcpyDeviceToHost); }
  printf("per step: addr+loads %.0f  update %.0f  chol %.0f  solves %.0f  store+sync %.0f\n", (double)c[0] / nc, (double)c[1] / nc, (double)c[2] / nc, (double)c[3] / nc, (double)c[4] / nc);
  for (int it = 0; it < 2; ++it) { chains<<<1, 32>>>(cyc, out, 2.0); cudaDeviceSynchronize(); cudaMemcpy(c, cyc, 64, cudaMemcpyDeviceToHost); }
  printf("per op: fast_rsqrt+add %.1f  mufu.rsq64+add %.1f  dmul %.1f cycles\n", c[0] / 64.0, c[1] / 64.0, c[2] / 64.0);
  return 0;
}
