// tc_test.cuh -- stand-alone check of the tcgen05 plumbing (tests only): out[128,N] = A[128,K] * W[N,K]^T with the
// 3xTF32 split, A staged into TMEM by the CTA, W_hi / W_lo given as canonical K-major shared-memory images.
#pragma once
#include "common.cuh"
#include "tc.cuh"

namespace coskad {

__global__ void __launch_bounds__(128, 1) tc_mix_test_kernel(const float* __restrict__ A, const float* __restrict__ Bhi,
                                                             const float* __restrict__ Blo, int K, int N, int swap_strides,
                                                             float* __restrict__ out) {
  extern __shared__ __align__(128) float sm[];
  float* bh = sm;
  float* bl = sm + N * K;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int stage = swap_strides >> 4;   // debug bisection: 1 = alloc only, 2 = + st, 3 = + mma, 0 = everything
  swap_strides &= 1;
  for (int i = tid; i < N * K; i += blockDim.x) { bh[i] = Bhi[i]; bl[i] = Blo[i]; }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = tmem_base_s;
  const uint32_t lane_base = tbase + (static_cast<uint32_t>(warp * 32) << 16);
  // stage A: row m = tid; hi -> cols [128, 128+K), lo -> cols [128+K, 128+2K); D -> cols [0, N)
  const uint32_t colA = 128;
  if (stage != 1)
  for (int k0 = 0; k0 < K; k0 += 16) {
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float a = (k0 + j < K) ? A[tid * K + k0 + j] : 0.f;
      tc::split_tf32(a, hi[j], lo[j]);
    }
    tc::tmem_st16(lane_base + colA + k0, hi);
    tc::tmem_st16(lane_base + colA + K + k0, lo);
  }
  tc::wait_st();
  tc::fence_before_sync();
  __syncthreads();
  if (tid == 0 && (stage == 0 || stage >= 3)) {
    tc::fence_after_sync();
    const uint32_t idesc = tc::make_idesc_tf32(128, N);
    const uint32_t lbo = (N / 8) * 128, sbo = 128;
    for (int kb = 0; kb < K / 8; ++kb) {
      const uint32_t offs = kb * 2 * lbo;
      const uint64_t dh = swap_strides ? tc::make_smem_desc(tc::smem_u32(bh) + offs, sbo, lbo) : tc::make_smem_desc(tc::smem_u32(bh) + offs, lbo, sbo);
      const uint64_t dl = swap_strides ? tc::make_smem_desc(tc::smem_u32(bl) + offs, sbo, lbo) : tc::make_smem_desc(tc::smem_u32(bl) + offs, lbo, sbo);
      tc::mma_tf32_ts(tbase, tbase + colA + kb * 8, dh, idesc, kb > 0 ? 1u : 0u);
      tc::mma_tf32_ts(tbase, tbase + colA + K + kb * 8, dh, idesc, 1u);
      tc::mma_tf32_ts(tbase, tbase + colA + kb * 8, dl, idesc, 1u);
    }
    tc::mma_commit(&bar);
  }
  if (stage == 0 || stage >= 3) tc::mbar_wait(&bar, 0);
  tc::fence_after_sync();
  if (stage == 0 || stage >= 4)
  for (int n0 = 0; n0 < N; n0 += 16) {
    uint32_t v[16];
    tc::tmem_ld16(lane_base + n0, v);
    tc::wait_ld();
#pragma unroll
    for (int j = 0; j < 16; ++j) out[tid * N + n0 + j] = __uint_as_float(v[j]);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tbase, 512);
}

// Sustained tcgen05 kind::tf32 rate of the device: every SM issues `iters` back-to-back M128 x N256 x K8 MMAs (A from TMEM,
// B from shared memory -- the operand form of the channel-mixing stages) onto one accumulator.  This is the measured
// denominator of the tensor-pipe roofline ("measure it the same way before quoting a tensor fraction").
__global__ void __launch_bounds__(128, 1) tf32_peak_kernel(int iters, float* __restrict__ sink) {
  __shared__ __align__(128) float bsm[256 * 8];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 256 * 8; i += blockDim.x) bsm[i] = __uint_as_float(__float_as_uint(1e-3f * static_cast<float>((i * 37) % 101 - 50)) & 0xFFFFE000u);
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  tc::fence_proxy_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = tmem_base_s;
  const uint32_t lane_base = tbase + (static_cast<uint32_t>(warp * 32) << 16);
  uint32_t a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = __float_as_uint(1e-3f * static_cast<float>((tid + 3 * j) % 17 - 8)) & 0xFFFFE000u;
  tc::tmem_st8(lane_base + 256, a);
  tc::wait_st();
  tc::fence_before_sync();
  __syncthreads();
  if (tid == 0) {
    tc::fence_after_sync();
    const uint32_t idesc = tc::make_idesc_tf32(128, 256);
    const uint64_t bd = tc::make_smem_desc(tc::smem_u32(bsm), (256 / 8) * 128, 128);
    for (int i = 0; i < iters; ++i) tc::mma_tf32_ts(tbase, tbase + 256, bd, idesc, i > 0 ? 1u : 0u);
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::fence_after_sync();
  uint32_t v[16];
  tc::tmem_ld16(lane_base, v);
  tc::wait_ld();
  if (__uint_as_float(v[0]) == 12345.678f) sink[0] = __uint_as_float(v[1]);     // never true; keeps the result live
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tbase, 512);
}

}  // namespace coskad
