// fp64 issue throughput / latency per warp and per SM: nvcc -arch=sm_100a -O3 -o ubench5 ubench5.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(long long* cyc, double* out, double a, double b, int active_lanes) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
  __syncthreads();
  long long t0 = clock64();
  if ((threadIdx.x & 31) < active_lanes) {
#pragma unroll 1
    for (int it = 0; it < 64; ++it) {
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ILP>
void run(int threads, int lanes, long long* cyc, double* out) {
  long long c = 0;
  for (int it = 0; it < 2; ++it) { k<ILP><<<1, threads>>>(cyc, out, 1.0000001, 1e-9, lanes); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); }
  const double n = 64.0 * 8 * ILP;
  printf("threads %4d lanes %2d ILP %2d: %7lld cycles, %.2f cycles per DFMA per warp, %.2f DFMA warp-instr/cycle/SM\n", threads, lanes, ILP, c, c / n, n * (threads / 32) / c);
}
int main() {
  long long* cyc; double* out;
  cudaMalloc(&cyc, 64); cudaMalloc(&out, 8 * 1024 * 8);
  for (int lanes : {32, 4}) {
    run<1>(32, lanes, cyc, out); run<2>(32, lanes, cyc, out); run<4>(32, lanes, cyc, out); run<8>(32, lanes, cyc, out); run<16>(32, lanes, cyc, out);
  }
  run<8>(128, 32, cyc, out); run<8>(256, 32, cyc, out); run<8>(512, 32, cyc, out); run<8>(1024, 32, cyc, out);
  run<1>(128, 32, cyc, out); run<1>(512, 32, cyc, out); run<1>(1024, 32, cyc, out);
  return 0;
}
