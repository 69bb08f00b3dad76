// (ubench9: the same gathers from PINNED HOST memory over PCIe)  DRAM fetch granularity of isolated 4-byte gathers on B200: N threads each read ONE float from a distinct pseudo-random
// 32-byte sector of a 2 GiB buffer, with different load flavours and cudaLimitMaxL2FetchGranularity settings.
// If a miss fetches 128 B from DRAM the time is ~4x that of 32-byte fetches.   nvcc -arch=sm_100a -O3 -o ubench8 ubench8.cu
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
template <int MODE>
__device__ __forceinline__ float ld(const float* p) {
  float v;
  if (MODE == 0) v = __ldg(p);
  else if (MODE == 1) v = *p;
  else if (MODE == 2) v = __ldcg(p);
  else if (MODE == 3) v = __ldcs(p);
  else if (MODE == 4) v = __ldlu(p);
  else if (MODE == 5) v = __ldcv(p);
  else if (MODE == 6) asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  else if (MODE == 7) asm volatile("ld.global.nc.L2::64B.f32 %0, [%1];" : "=f"(v) : "l"(p));
  else if (MODE == 8) asm volatile("ld.global.nc.L2::128B.f32 %0, [%1];" : "=f"(v) : "l"(p));
  else if (MODE == 9) asm volatile("ld.global.nc.L1::evict_first.f32 %0, [%1];" : "=f"(v) : "l"(p));
  else asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
template <int MODE>
__global__ void gather(const float* __restrict__ buf, unsigned nsec_mask, unsigned n, unsigned salt, float* out) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned s = ((i + salt) * 2654435761u) & nsec_mask;       // odd multiplier: a bijection on 2^k sectors
  const float v = ld<MODE>(buf + (size_t)s * 8 + (i & 7));
  if (v == 123456.0f) out[0] = v;
}
template <int MODE>
float run(const float* buf, unsigned mask, unsigned n, float* out) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  gather<MODE><<<(n + 255) / 256, 256>>>(buf, mask, n, 1u << 20, out);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(e0);
    gather<MODE><<<(n + 255) / 256, 256>>>(buf, mask, n, (r + 2) * 40000003u, out);   // fresh sectors every run
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}
int main() {
  const size_t bytes = 1ull << 30; const unsigned nsec = (unsigned)(bytes / 32), n = 1u << 20;
  float *buf, *out; cudaHostAlloc(&buf, bytes, cudaHostAllocMapped); cudaMalloc(&out, 4); memset(buf, 0, bytes);
  const char* names[] = {"ld.global.nc (__ldg)", "ld.global", "ld.global.cg", "ld.global.cs", "ld.global.lu", "ld.global.cv", "nc.L1::no_allocate", "nc.L2::64B", "nc.L2::128B", "nc.L1::evict_first", "ld.relaxed.gpu"};
  const size_t grans[] = {0, 32};
  for (size_t g : grans) {
    if (g) { cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g); if (e != cudaSuccess) printf("set limit %zu: %s\n", g, cudaGetErrorString(e)); }
    size_t cur = 0; cudaDeviceGetLimit(&cur, cudaLimitMaxL2FetchGranularity);
    printf("cudaLimitMaxL2FetchGranularity = %zu\n", cur);
    float t[11];
    t[0] = run<0>(buf, nsec - 1, n, out); t[1] = run<1>(buf, nsec - 1, n, out); t[2] = run<2>(buf, nsec - 1, n, out);
    t[3] = run<3>(buf, nsec - 1, n, out); t[4] = run<4>(buf, nsec - 1, n, out); t[5] = run<5>(buf, nsec - 1, n, out);
    t[6] = run<6>(buf, nsec - 1, n, out); t[7] = run<7>(buf, nsec - 1, n, out); t[8] = run<8>(buf, nsec - 1, n, out);
    t[9] = run<9>(buf, nsec - 1, n, out); t[10] = run<10>(buf, nsec - 1, n, out);
    for (int m = 0; m < 11; ++m)
      printf("  %-24s %8.3f ms  = %7.1f GB/s of 32-byte sectors (%.2f G sectors/s)\n", names[m], t[m], n * 32.0 / (t[m] * 1e-3) / 1e9, n / (t[m] * 1e-3) / 1e9);
  }
  return 0;
}
