// Host/device portability shims.
//
// The factor arithmetic (factors.cuh) and the mixed-precision solver (mp.cuh) are written as
// __host__ __device__ code so that tests/host_emu can run the SAME source on CPU threads (this
// repository is developed on a box without a GPU; the emulator is test infrastructure only and is
// never linked into libdgpmp2_b200.so's callers).  In device code every wrapper below is exactly the
// intrinsic it names.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__)
#define DG_DEV 1
#else
#define DG_DEV 0
#endif

#define DG_HD __host__ __device__ __forceinline__

namespace dgpmp2 {

// separately rounded IEEE double operations (never contracted into FMAs)
DG_HD double dg_dadd(double a, double b) {
#if DG_DEV
  return __dadd_rn(a, b);
#else
  volatile double r = a + b; return r;
#endif
}
DG_HD double dg_dsub(double a, double b) {
#if DG_DEV
  return __dsub_rn(a, b);
#else
  volatile double r = a - b; return r;
#endif
}
DG_HD double dg_dmul(double a, double b) {
#if DG_DEV
  return __dmul_rn(a, b);
#else
  volatile double r = a * b; return r;
#endif
}
DG_HD double dg_ddiv(double a, double b) {
#if DG_DEV
  return __ddiv_rn(a, b);
#else
  volatile double r = a / b; return r;
#endif
}
DG_HD float dg_fmul(float a, float b) {
#if DG_DEV
  return __fmul_rn(a, b);
#else
  volatile float r = a * b; return r;
#endif
}
// floor to int, saturating (cvt.rmi.s32.f64)
DG_HD int dg_d2i_rd(double x) {
#if DG_DEV
  return __double2int_rd(x);
#else
  const double f = floor(x);
  if (!(f == f)) return 0;
  if (f >= 2147483647.0) return 2147483647;
  if (f <= -2147483648.0) return (-2147483647 - 1);
  return (int)f;
#endif
}
// read-only global load
template <typename T> DG_HD T dg_ldg(const T* p) {
#if DG_DEV
  return __ldg(p);
#else
  return *p;
#endif
}
// 1/sqrt(x), fp32, approximate on the device (MUFU.RSQ, ~2^-22.4 relative)
DG_HD float dg_rsqrtf(float x) {
#if DG_DEV
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#else
  return 1.0f / sqrtf(x);
#endif
}
DG_HD void dg_sincos(double x, double* s, double* c) {
#if DG_DEV
  sincos(x, s, c);
#else
  *s = sin(x); *c = cos(x);
#endif
}
DG_HD int dg_min(int a, int b) {
#if DG_DEV
  return min(a, b);
#else
  return a < b ? a : b;
#endif
}
DG_HD int dg_max(int a, int b) {
#if DG_DEV
  return max(a, b);
#else
  return a > b ? a : b;
#endif
}

}  // namespace dgpmp2
