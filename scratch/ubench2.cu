#include <cstdio>
#include <cuda_runtime.h>
#define REP16(x) x x x x x x x x x x x x x x x x
#define REP64(x) REP16(x) REP16(x) REP16(x) REP16(x)
__global__ void k(double* out, long long* cyc, double a, double b, float fa, float fb) {
  double x = a, y = a + 0.5; float fx = fa;
  long long t0, t1;
  t0 = clock64();
  REP64(asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x) : "d"(b), "d"(a));)
  t1 = clock64(); if (threadIdx.x == 0) cyc[0] = t1 - t0;
  t0 = clock64();
  REP64(asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(x) : "d"(b));)
  t1 = clock64(); if (threadIdx.x == 0) cyc[1] = t1 - t0;
  t0 = clock64();
  REP64(asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x) : "d"(b));)
  t1 = clock64(); if (threadIdx.x == 0) cyc[2] = t1 - t0;
  t0 = clock64();
  REP64(asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(fx) : "f"(fb), "f"(fa));)
  t1 = clock64(); if (threadIdx.x == 0) cyc[3] = t1 - t0;
  t0 = clock64();
  REP64(asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(y));)
  t1 = clock64(); if (threadIdx.x == 0) cyc[4] = t1 - t0;
  t0 = clock64();
  REP64(asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(y));)
  t1 = clock64(); if (threadIdx.x == 0) cyc[5] = t1 - t0;
  // two interleaved independent DFMA chains
  double z = a + 2;
  t0 = clock64();
  REP64(asm volatile("fma.rn.f64 %0, %0, %2, %3;\n\tfma.rn.f64 %1, %1, %2, %3;" : "+d"(x), "+d"(z) : "d"(b), "d"(a));)
  t1 = clock64(); if (threadIdx.x == 0) cyc[6] = t1 - t0;
  double z2 = a + 3, z3 = a + 4;
  t0 = clock64();
  REP64(asm volatile("fma.rn.f64 %0, %0, %4, %5;\n\tfma.rn.f64 %1, %1, %4, %5;\n\tfma.rn.f64 %2, %2, %4, %5;\n\tfma.rn.f64 %3, %3, %4, %5;" : "+d"(x), "+d"(z), "+d"(z2), "+d"(z3) : "d"(b), "d"(a));)
  t1 = clock64(); if (threadIdx.x == 0) cyc[7] = t1 - t0;
  float f1 = fa + 1, f2 = fa + 2, f3 = fa + 3;
  t0 = clock64();
  REP64(asm volatile("fma.rn.f32 %0, %0, %4, %5;\n\tfma.rn.f32 %1, %1, %4, %5;\n\tfma.rn.f32 %2, %2, %4, %5;\n\tfma.rn.f32 %3, %3, %4, %5;" : "+f"(fx), "+f"(f1), "+f"(f2), "+f"(f3) : "f"(fb), "f"(fa));)
  t1 = clock64(); if (threadIdx.x == 0) cyc[8] = t1 - t0;
  float fr = fa + 0.5f;
  t0 = clock64();
  REP64(asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(fr));)
  t1 = clock64(); if (threadIdx.x == 0) cyc[9] = t1 - t0;
  out[threadIdx.x] = x + fx + y + z + z2 + z3 + f1 + f2 + f3 + fr;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 16 * 8);
  for (int warps = 1; warps <= 8; warps *= 2) {
    k<<<1, 32 * warps>>>(out, cyc, 1.0000001, 0.9999999, 1.0f, 0.5f);
    cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("warps=%d: DFMA %.1f DMUL %.1f DADD %.1f FFMA %.1f rsqrt64 %.1f rcp64 %.1f | 2xDFMA %.1f (per pair) 4xDFMA %.1f (per quad) 4xFFMA %.1f rsqrt32 %.1f\n",
           warps, h[0] / 64.0, h[1] / 64.0, h[2] / 64.0, h[3] / 64.0, h[4] / 64.0, h[5] / 64.0, h[6] / 64.0, h[7] / 64.0, h[8] / 64.0, h[9] / 64.0);
  }
  return 0;
}
