// fp64 issue rate with three distinct register operands per DFMA (as in real code) vs shared operands
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(long long* cyc, double* out, const double* in) {
  constexpr int ILP = 8;
  double acc[ILP], b[ILP], c[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { acc[i] = in[threadIdx.x + i]; b[i] = in[64 + threadIdx.x + i]; c[i] = in[128 + threadIdx.x + i]; }
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 64; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        if (MODE == 0) acc[i] = fma(acc[i], b[0], c[0]);                 // shared operands
        if (MODE == 1) acc[i] = fma(acc[i], b[i], c[i]);                 // 3 distinct registers
        if (MODE == 2) acc[i] = fma(b[i], c[(i + r) % ILP], acc[i]);     // dot-product style accumulate
        if (MODE == 3) acc[i] = __dmul_rn(acc[i], b[i]);                 // DMUL 2 operands
        if (MODE == 4) acc[i] = __dadd_rn(acc[i], b[i]);                 // DADD
      }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int MODE>
void run(const char* name, long long* cyc, double* out, double* in) {
  for (int threads : {32, 128, 256, 512}) {
    long long c = 0;
    for (int it = 0; it < 2; ++it) { k<MODE><<<1, threads>>>(cyc, out, in); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); }
    const double n = 64.0 * 8 * 8;
    printf("%-28s threads %4d: %.2f cycles per instr per warp, %.2f warp-instr/cycle/SM\n", name, threads, c / n, n * (threads / 32) / c);
  }
}
int main() {
  long long* cyc; double *out, *in;
  cudaMalloc(&cyc, 64); cudaMalloc(&out, 8 * 1024 * 8); cudaMalloc(&in, 2048 * 8);
  double h[2048]; for (int i = 0; i < 2048; ++i) h[i] = 1.0 + 1e-9 * i;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  run<0>("DFMA shared operands", cyc, out, in);
  run<1>("DFMA 3 distinct regs", cyc, out, in);
  run<2>("DFMA dot-style", cyc, out, in);
  run<3>("DMUL", cyc, out, in);
  run<4>("DADD", cyc, out, in);
  return 0;
}
