// common.cuh -- shared constants and small device helpers for libcoskad_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace coskad {

// Shapes every reference config uses (config/*/*.yaml: dataset_seg_len 12, 17 joints,
// channels [32,16,32], h_dim 64, num_coords 2).
constexpr int kT = 12;              // frames per window
constexpr int kV = 17;              // joints
constexpr int kP = kT * kV;         // 204 positions
constexpr int kC0 = 2, kC1 = 32, kC2 = 16, kC3 = 32, kC4 = 64;
constexpr int kF = kC4 * kP;        // 13056 flattened features (c,t,v order, models/sts/ae.py:96-100)
constexpr int kDP = 16;             // head rows padded to 16 (latent 16, 8, or 8+1 for the VAE head)

// Fused-kernel tiling
constexpr int kNW = 3;              // windows per CTA tile
constexpr int kCS = 205;            // smem row stride (floats) of one channel plane: odd => a warp whose lanes
                                    // walk rows (n,c) at a fixed position hits 32 distinct banks
constexpr int kAW = 20;             // A[t][v][.] rows padded 17 -> 20 floats so a row is 5 x LDS.128
constexpr int kThreads = 448;       // 14 warps: 7 position chunks x 2 channel chunks in the mixing stages
constexpr int kWarps = kThreads / 32;
constexpr int kPCH = (kP + 31) / 32;   // 7 position chunks of 32

constexpr int kTwFloats = kV * kT * kT;     // 2448
constexpr int kAwFloats = kT * kV * kAW;    // 4080

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// sum over aligned groups of 16 lanes
__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// packed FP32 FMA (sm_100a FFMA2): d.{lo,hi} += a.{lo,hi} * b.{lo,hi}; one issue slot for two FMAs (measured 92 % of
// the scalar FFMA rate in FLOP terms, tools/ubench/ffma2.cu) -- used where the FP32 stages are issue-bound
__device__ __forceinline__ void ffma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ unsigned long long dup2(float x) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "r"(__float_as_uint(x)));
  return r;
}
__device__ __forceinline__ float lo2(unsigned long long v) { return __uint_as_float(static_cast<uint32_t>(v & 0xffffffffull)); }
__device__ __forceinline__ float hi2(unsigned long long v) { return __uint_as_float(static_cast<uint32_t>(v >> 32)); }

__device__ __forceinline__ float prelu(float v, float a) { return v >= 0.f ? v : a * v; }

}  // namespace coskad
