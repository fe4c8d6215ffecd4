// reduce.cuh -- the second stage of every cross-CTA reduction of the training path: per-CTA partials in a ctx-owned workspace
// are added in a FIXED order with a double accumulator (no floating-point atomics: two runs are bit-identical).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace coskad {

// out[i] += sum_j part[j*stride + i] in a FIXED order (double accumulator): the second stage of every cross-CTA reduction.
// Few partials: one thread per element.  Many partials (hundreds of CTAs): a 32 x kPsRows block owns 32 consecutive elements,
// row y sums the partials y, y + kPsRows, .. (coalesced across x), the rows meet in shared memory in a fixed order.
template <typename TOut>
__global__ void partial_sum_kernel(const float* __restrict__ part, int nparts, int64_t stride, int64_t n, TOut* out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  int j = 0;
  for (; j + 8 <= nparts; j += 8) {               // 8 loads in flight, added in ascending order
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = part[static_cast<int64_t>(j + u) * stride + i];
#pragma unroll
    for (int u = 0; u < 8; ++u) s += static_cast<double>(v[u]);
  }
  for (; j < nparts; ++j) s += static_cast<double>(part[static_cast<int64_t>(j) * stride + i]);
  out[i] += static_cast<TOut>(s);
}
constexpr int kPsRows = 16;
// row y of a 32 x kPsRows block sums the partials y, y + kPsRows, .. of element i (ascending j: a fixed order); the loads go
// out in batches of 8 (as a load -> add loop every partial cost one exposed L2 round trip: the second-stage kernels of a
// training step took 5-13 us each for a few KB of data); the rows then meet in shared memory, again in a fixed order
template <typename TIn>
__device__ __forceinline__ double partial_sum_block(const TIn* __restrict__ part, int nparts, int64_t stride, int64_t i, bool ok,
                                                    double (*sh)[33]) {
  double s = 0.0;
  if (ok) {
    int j = threadIdx.y;
    for (; j + 7 * kPsRows < nparts; j += 8 * kPsRows) {
      TIn v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = part[static_cast<int64_t>(j + u * kPsRows) * stride + i];
#pragma unroll
      for (int u = 0; u < 8; ++u) s += static_cast<double>(v[u]);
    }
    for (; j < nparts; j += kPsRows) s += static_cast<double>(part[static_cast<int64_t>(j) * stride + i]);
  }
  sh[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.y == 0)
#pragma unroll
    for (int y = 0; y < kPsRows; ++y) t += sh[y][threadIdx.x];
  return t;
}
template <typename TOut>
__global__ void partial_sum_wide_kernel(const float* __restrict__ part, int nparts, int64_t stride, int64_t n, TOut* out) {
  __shared__ double sh[kPsRows][33];
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 32 + threadIdx.x;
  const double t = partial_sum_block(part, nparts, stride, i, i < n, sh);
  if (threadIdx.y == 0 && i < n) out[i] += static_cast<TOut>(t);
}


}  // namespace coskad
