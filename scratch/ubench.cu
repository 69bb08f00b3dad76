// micro-benchmarks: dependent-chain latency of fp64 ops, fp32 ops, LDS on this GPU
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double a, double b, float fa, float fb) {
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i * 1e-3;
  __syncthreads();
  double x = a; float fx = fa;
  long long t0, t1;
  // DFMA dependent chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) x = fma(x, b, a);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  // DMUL chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) x = x * b;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[1] = t1 - t0;
  // DADD chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) x = x + b;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[2] = t1 - t0;
  // FFMA chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; ++i) fx = fmaf(fx, fb, fa);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[3] = t1 - t0;
  // rsqrt.approx.f64 chain
  double y = a + 1.5;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) { double r; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(y)); y = r + 1.0; }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[4] = t1 - t0;   // includes one DADD per iter
  // rcp.approx.f64 chain
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(y)); y = r + 1.0; }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[5] = t1 - t0;
  // LDS.64 pointer chase
  int idx = threadIdx.x & 1023;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) { double v = sm[idx]; idx = ((int)(v * 1000.0 + 0.5) + 7) & 1023; }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[6] = t1 - t0;   // includes conversion chain
  // independent DFMA throughput: 8 chains
  double z0=a,z1=a+1,z2=a+2,z3=a+3,z4=a+4,z5=a+5,z6=a+6,z7=a+7;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) { z0=fma(z0,b,a); z1=fma(z1,b,a); z2=fma(z2,b,a); z3=fma(z3,b,a); z4=fma(z4,b,a); z5=fma(z5,b,a); z6=fma(z6,b,a); z7=fma(z7,b,a); }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[7] = t1 - t0;
  // sqrt / full-precision rsqrt / division chains
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 32; ++i) y = rsqrt(y) + 1.0;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[8] = t1 - t0;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 32; ++i) y = a / y + 1.0;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[9] = t1 - t0;
  // 128-bit LDS latency (dependent via index)
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) { double2 v = *reinterpret_cast<double2*>(&sm[idx & 1022]); idx = (__double2int_rn(v.x * 1000.0) + 6) & 1023; }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[10] = t1 - t0;
  out[threadIdx.x] = x + fx + y + idx + z0+z1+z2+z3+z4+z5+z6+z7;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 16 * 8);
  for (int warps = 1; warps <= 4; warps *= 2) {
    k<<<1, 32 * warps>>>(out, cyc, 1.0000001, 0.9999999, 1.0f, 0.5f);
    cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("warps/SM-CTA=%d: DFMA %.1f  DMUL %.1f  DADD %.1f  FFMA %.1f  rsqrt.approx+dadd %.1f  rcp.approx+dadd %.1f  LDS64+cvt chain %.1f  8xDFMA(indep)/8 %.2f  rsqrt()+dadd %.1f  div+dadd %.1f LDS128+cvt %.1f\n",
           warps, h[0] / 256.0, h[1] / 256.0, h[2] / 256.0, h[3] / 256.0, h[4] / 64.0, h[5] / 64.0, h[6] / 64.0, h[7] / 512.0, h[8] / 32.0, h[9] / 32.0, h[10] / 64.0);
  }
  return 0;
}
