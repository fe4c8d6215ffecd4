// microbenchmark: how fast can ONE SM (all 148 concurrently) stream an L2-resident 835 KB buffer with 12 warps?
//   mode 0: LDG.128 with U loads in flight per thread;  mode 1: cp.async 16 B into shared memory (U groups in flight)
#include <cstdio>
#include <cuda_runtime.h>
template <int U>
__global__ void __launch_bounds__(384, 1) k_ldg(const uint4* __restrict__ w, int n16, int reps, uint4* out) {
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (int r = 0; r < reps; ++r)
    for (int i0 = threadIdx.x; i0 < n16; i0 += 384 * U) {
      uint4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) { const int i = i0 + u * 384; v[u] = (i < n16) ? __ldg(w + i) : make_uint4(0, 0, 0, 0); }
#pragma unroll
      for (int u = 0; u < U; ++u) { acc.x ^= v[u].x; acc.y += v[u].y; acc.z ^= v[u].z; acc.w += v[u].w; }
    }
  if (acc.x == 0x12345678u) out[blockIdx.x * 384 + threadIdx.x] = acc;
}
template <int U>
__global__ void __launch_bounds__(384, 1) k_cpasync(const uint4* __restrict__ w, int n16, int reps, uint4* out) {
  extern __shared__ uint4 sm[];        // U stages x 384 x 16 B
  uint4 acc = make_uint4(0, 0, 0, 0);
  const unsigned sbase = static_cast<unsigned>(__cvta_generic_to_shared(sm));
  for (int r = 0; r < reps; ++r) {
    int issued = 0, done = 0;
    const int total = (n16 + 383) / 384;
    for (; issued < U && issued < total; ++issued) {
      const int i = issued * 384 + threadIdx.x;
      if (i < n16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + ((issued % U) * 384 + threadIdx.x) * 16), "l"(w + i));
      asm volatile("cp.async.commit_group;");
    }
    for (; done < total; ++done) {
      asm volatile("cp.async.wait_group %0;" ::"n"(U - 1));
      const uint4 v = sm[(done % U) * 384 + threadIdx.x];
      acc.x ^= v.x; acc.y += v.y; acc.z ^= v.z; acc.w += v.w;
      const int i = issued * 384 + threadIdx.x;
      if (issued < total && i < n16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + ((issued % U) * 384 + threadIdx.x) * 16), "l"(w + i));
      asm volatile("cp.async.commit_group;");
      ++issued;
    }
    asm volatile("cp.async.wait_group 0;");
  }
  if (acc.x == 0x12345678u) out[blockIdx.x * 384 + threadIdx.x] = acc;
}
int main() {
  const int bytes = 835584, n16 = bytes / 16, reps = 200;
  uint4 *w, *out; cudaMalloc(&w, bytes); cudaMemset(w, 1, bytes); cudaMalloc(&out, 148 * 384 * 16);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  auto report = [&](const char* name, float ms) {
    const double cyc = ms * 1e-3 * clk * 1e3;
    printf("%-22s %.3f ms  %.1f B/clk/SM  (%.2f TB/s chip)\n", name, ms, double(bytes) * reps / cyc, 148.0 * bytes * reps / (ms * 1e-3) / 1e12);
  };
#define RUN_LDG(U) { cudaFuncSetAttribute(k_ldg<U>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); k_ldg<U><<<148, 384, 200 * 1024>>>(w, n16, 2, out); cudaEventRecord(e0); k_ldg<U><<<148, 384, 200 * 1024>>>(w, n16, reps, out); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); report("LDG.128 (200 KB smem carve-out) U=" #U, ms); }
#define RUN_CPA(U) { cudaFuncSetAttribute(k_cpasync<U>, cudaFuncAttributeMaxDynamicSharedMemorySize, U * 384 * 16); k_cpasync<U><<<148, 384, U * 384 * 16>>>(w, n16, 2, out); cudaEventRecord(e0); k_cpasync<U><<<148, 384, U * 384 * 16>>>(w, n16, reps, out); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); report("cp.async16 U=" #U, ms); }
  RUN_LDG(2) RUN_LDG(4) RUN_LDG(8) RUN_LDG(16)
  RUN_CPA(4) RUN_CPA(8) RUN_CPA(16)
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
