// latency of the sequential tail (bcr_tail) on a synthetic SPD chain: nvcc -arch=sm_100a -O3 -o ubench4 ubench4.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../dgpmp2_b200/csrc/bcr.cuh"
using namespace dgpmp2;
template <int NT>
__global__ void k(int T, int nc, long long* cyc, double* out) {
  extern __shared__ double sm[];
  __shared__ int fail[8];
  using N = Node<4>;
  for (int i = threadIdx.x; i < T * N::kStride; i += blockDim.x) {
    const int f = i % N::kStride;
    double v = 0.01 * ((i * 7) % 13);
    if (f < 16 && (f / 4 == f % 4)) v += 8.0;
    sm[i] = v;
  }
  if (threadIdx.x < 8) fail[threadIdx.x] = 0;
  __syncthreads();
  long long t0 = clock64();
  bcr_tail<4>(sm, T, 1, 1, nc, fail);
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; out[0] = sm[N::oR] + fail[0]; }
}
int main() {
  long long* cyc; double* out;
  cudaMalloc(&cyc, 64); cudaMalloc(&out, 64);
  const int T = 64;
  const int smem = T * Node<4>::kStride * 8;
  cudaFuncSetAttribute(k<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int threads : {32, 128, 512})
    for (int nc : {1, 2, 8, 16, 64}) {
      long long c = 0;
      for (int it = 0; it < 2; ++it) { k<32><<<1, threads, smem>>>(T, nc, cyc, out); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); }
      printf("threads %d nc %d: %lld cycles (%.0f per node)\n", threads, nc, c, (double)c / nc);
    }
  return 0;
}
