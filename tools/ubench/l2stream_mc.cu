// microbenchmark: can a thread-block cluster cut the per-SM L2 fetch of the eval kernel's head-weight stream?
//
// The head stage of fused_eval_tc_kernel re-streams the same L2-resident 835 KB of btlnk weights on every SM for every
// 3-window tile and runs at the per-SM L2 fetch rate (l2stream.cu: 68 B/clk per SM with LDG.128, all 148 SMs streaming).
// VERDICT r1 item 4 proposes a cluster in which each CTA fetches 1/k of every weight slab and multicasts it
// (cp.async.bulk ... .multicast::cluster) into a shared ring of all k CTAs.  This measures the DELIVERED bytes per clock
// and SM of exactly that pattern, next to the unicast bulk-copy ring and the LDG.128 stream, with all 148 SMs running:
//   mode ring<1>: every CTA bulk-copies every 8 KB slab itself (unicast), ring of 8 slabs, dedicated producer warp
//   mode ring<2>, ring<4>: cluster of 2 / 4 CTAs, CTA r copies part r of each slab and multicasts it to the whole cluster
// Consumers (12 warps) wait on the slab's full barrier, touch it with LDS.128 and release it to every CTA of the cluster.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2stream_mc l2stream_mc.cu && timeout 60 ./l2stream_mc
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kConsumers = 384, kThreads = kConsumers + 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  const uint32_t a = smem_u32(b);
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* b, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(b)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(r) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int CS, int kSlab, int kNB>
__global__ void __launch_bounds__(kThreads, 1) k_ring(const char* __restrict__ w, int nslab, int reps, uint4* out) {
  extern __shared__ __align__(128) char ring[];          // kNB x kSlab
  __shared__ __align__(8) uint64_t full[kNB], empty[kNB];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t rank = 0;
  if (CS > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (tid == 0) {
    for (int b = 0; b < kNB; ++b) { mbar_init(full + b, 1); mbar_init(empty + b, CS * (kConsumers / 32)); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (CS > 1) cluster_sync_all();
  const int total = nslab * reps;
  uint4 acc = make_uint4(0, 0, 0, 0);
  if (warp == kConsumers / 32) {                          // producer warp
    if (lane == 0) {
      constexpr uint32_t part = kSlab / CS;
      for (int s = 0; s < total; ++s) {
        const int b = s % kNB;
        if (s >= kNB) mbar_wait(empty + b, ((s / kNB) - 1) & 1);      // every CTA of the cluster released the slot
        mbar_expect_tx(full + b, kSlab);                               // the CS parts land here from CS producers
        const char* src = w + static_cast<size_t>(s % nslab) * kSlab + rank * part;
        const uint32_t dst = smem_u32(ring + b * kSlab + rank * part), bar = smem_u32(full + b);
        if (CS == 1)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(dst), "l"(src), "r"(part), "r"(bar) : "memory");
        else
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                       ::"r"(dst), "l"(src), "r"(part), "r"(bar), "h"(static_cast<uint16_t>((1u << CS) - 1)) : "memory");
      }
    }
  } else {                                                // 12 consumer warps
    for (int s = 0; s < total; ++s) {
      const int b = s % kNB;
      mbar_wait(full + b, (s / kNB) & 1);
      const uint4* p = reinterpret_cast<const uint4*>(ring + b * kSlab);
      for (int i = tid; i < kSlab / 16; i += kConsumers) { const uint4 v = p[i]; acc.x ^= v.x; acc.y += v.y; acc.z ^= v.z; acc.w += v.w; }
      __syncwarp();
      if (lane == 0)
        for (int c = 0; c < CS; ++c) mbar_arrive_cluster(empty + b, c);
    }
  }
  if (acc.x == 0x12345678u) out[blockIdx.x * kThreads + tid] = acc;
  __syncthreads();
  if (CS > 1) cluster_sync_all();                         // nobody exits while a peer may still signal its barriers
}

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

template <int CS, int kSlab, int kNB>
static float run_ring(const char* w, int bytes, int reps, uint4* out, int grid) {
  const int nslab = bytes / kSlab;
  cudaFuncSetAttribute(k_ring<CS, kSlab, kNB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = 200 * 1024;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaLaunchKernelEx(&cfg, k_ring<CS, kSlab, kNB>, w, nslab, 2, out);
  cudaEventRecord(e0);
  cudaLaunchKernelEx(&cfg, k_ring<CS, kSlab, kNB>, w, nslab, reps, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  const int bytes = 835584 / 32768 * 32768, reps = 200;      // 25 x 32 KB = 819 200 B (the ring walks whole slabs)
  char* w; uint4* out; cudaMalloc(&w, bytes); cudaMemset(w, 1, bytes); cudaMalloc(&out, 148 * kThreads * 16);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  auto report = [&](const char* name, float ms, int sms, double l2_fraction) {
    const double cyc = ms * 1e-3 * clk * 1e3;
    printf("%-44s %8.3f ms  delivered %6.1f B/clk/SM  (L2 read %6.1f B/clk/SM)  %s\n", name, ms, double(bytes) * reps / cyc,
           double(bytes) * reps / cyc * l2_fraction, cudaGetErrorString(cudaGetLastError()));
    (void)sms;
  };
  {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaFuncSetAttribute(k_ldg<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    k_ldg<8><<<148, 384, 200 * 1024>>>(reinterpret_cast<const uint4*>(w), bytes / 16, 2, out);
    cudaEventRecord(e0);
    k_ldg<8><<<148, 384, 200 * 1024>>>(reinterpret_cast<const uint4*>(w), bytes / 16, reps, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    report("LDG.128 U=8, 148 CTAs (today's head path)", ms, 148, 1.0);
  }
  report("bulk ring 8 x 8 KB, unicast", run_ring<1, 8192, 8>(w, bytes, reps, out, 148), 148, 1.0);
  report("bulk ring 8 x 8 KB, cluster 2 multicast", run_ring<2, 8192, 8>(w, bytes, reps, out, 148), 148, 0.5);
  report("bulk ring 8 x 8 KB, cluster 4 multicast", run_ring<4, 8192, 8>(w, bytes, reps, out, 148), 148, 0.25);
  report("bulk ring 6 x 32 KB, unicast", run_ring<1, 32768, 6>(w, bytes, reps, out, 148), 148, 1.0);
  report("bulk ring 6 x 32 KB, cluster 2 multicast", run_ring<2, 32768, 6>(w, bytes, reps, out, 148), 148, 0.5);
  report("bulk ring 6 x 32 KB, cluster 4 multicast", run_ring<4, 32768, 6>(w, bytes, reps, out, 148), 148, 0.25);
  cudaDeviceSynchronize();
  printf("final: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
