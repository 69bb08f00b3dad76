#include <cstdio>
#include <cuda_runtime.h>
#include "../dgpmp2_b200/csrc/bcr.cuh"
using namespace dgpmp2;
__global__ void k(double* io, long long* cyc) {
  __shared__ double sm[64 * 54];
  for (int i = threadIdx.x; i < 64 * 54; i += blockDim.x) sm[i] = io[i % 256] + (i % 7 == 0 ? 8.0 : 0.1);
  __syncthreads();
  double* nd = sm + (threadIdx.x / 4) * 54;
  long long t0, t1, t2, t3, t4;
  double L[10], v[4], w[4];
  t0 = clock64();
  ld_lower<4>(nd, L);
  ld_vec<4>(nd + 16, v);
  ld_vec<4>(nd + 32, w);
  // force loads complete
  double s = L[0] + L[9] + v[0] + w[3];
  asm volatile("" :: "d"(s));
  t1 = clock64();
  bool ok = chol_packed<4>(L);
  asm volatile("" :: "d"(L[9]), "d"(L[0]));
  t2 = clock64();
  fwd_solve<4>(L, v);
  fwd_solve<4>(L, w);
  asm volatile("" :: "d"(v[3]), "d"(w[3]));
  t3 = clock64();
  bwd_solve<4>(L, v);
  asm volatile("" :: "d"(v[0]));
  t4 = clock64();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; }
  io[threadIdx.x] = v[0] + v[1] + w[2] + L[5] + (ok ? 1.0 : 0.0);
}
int main() {
  double* io; long long* cyc;
  cudaMalloc(&io, 4096 * 8); cudaMalloc(&cyc, 64);
  double h[256]; for (int i = 0; i < 256; ++i) h[i] = 0.01 * (i % 13);
  cudaMemcpy(io, h, sizeof(h), cudaMemcpyHostToDevice);
  for (int it = 0; it < 2; ++it) {
    k<<<1, 32>>>(io, cyc); cudaDeviceSynchronize();
    long long c[8]; cudaMemcpy(c, cyc, 32, cudaMemcpyDeviceToHost);
    printf("loads %lld  chol %lld  2x fwd_solve %lld  bwd_solve %lld cycles\n", c[0], c[1], c[2], c[3]);
  }
  return 0;
}
