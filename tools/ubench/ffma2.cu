// microbenchmark: FFMA vs packed fma.rn.f32x2 throughput on sm_100a (register-resident)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
  unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
  const unsigned long long aa = *reinterpret_cast<const unsigned long long*>(&a);
  const unsigned long long bb = *reinterpret_cast<const unsigned long long*>(&b);
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
  d = *reinterpret_cast<float2*>(&dd);
}
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
  float2 acc[16];
  for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x + i, i);
  const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.9999f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) { acc[i].x = fmaf(acc[i].x, a2.x, b2.x); acc[i].y = fmaf(acc[i].y, a2.y, b2.y); }
      else { float2 t = acc[i]; unsigned long long dd = *reinterpret_cast<unsigned long long*>(&t);
             const unsigned long long aa = *reinterpret_cast<const unsigned long long*>(&a2);
             const unsigned long long bb = *reinterpret_cast<const unsigned long long*>(&b2);
             asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(dd) : "l"(aa), "l"(bb));
             acc[i] = *reinterpret_cast<float2*>(&dd); }
    }
  }
  float s = 0; for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 2; ++mode) for (int warps = 4; warps <= 32; warps *= 2) {
    const int iters = 20000; float ms;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148, warps * 32>>>(out, iters, 1.0001f, 0.5f); else k<1><<<148, warps * 32>>>(out, iters, 1.0001f, 0.5f);
      cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    }
    const double flop = 2.0 * 32 * iters * 148.0 * warps * 32;
    printf("mode %s warps/SM %2d: %.3f ms  %.1f TFLOP/s  err=%s\n", mode ? "FFMA2" : "FFMA ", warps, ms, flop / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
