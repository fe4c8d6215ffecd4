// train_ops.cuh -- kernels of the training path.
#pragma once
#include "common.cuh"
#include "geometry.cuh"
#include "latent_ops.cuh"
#include "reduce.cuh"

namespace coskad {

// Backward of  s(z) = gmath.dist(c, x),  x = [project](expmap0(z)),  k = -1
// (training loss, models/hyperbolic_encoder.py:147,157; plain autograd through the geoopt formulas:
// clamps pass zero gradient outside their range).  One warp per row, D <= 32.
__global__ void poincare_score_bwd_kernel(const float* __restrict__ z, const float* __restrict__ center,
                                          const float* __restrict__ dscore, int64_t B, int D, int with_project,
                                          float* __restrict__ dz) {
  const int lane = threadIdx.x & 31;
  const float R = 1.f - 4e-3f;
  const float cd = (lane < D) ? center[lane] : 0.f;
  const float a = -cd;                                  // mobius_add(-c, x)
  const float a2 = warp_sum(a * a);
  const int64_t wpg = static_cast<int64_t>(gridDim.x) * kRowWarps;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * kRowWarps + (threadIdx.x >> 5); r < B; r += wpg) {
    const float zd = (lane < D) ? z[r * D + lane] : 0.f;
    // forward
    const float nraw = sqrtf(warp_sum(zd * zd));
    const float n = fmaxf(nraw, 1e-15f);
    const bool n_live = nraw > 1e-15f;
    const float t = clamped_tanh(n);
    const float e = t * (zd / n);
    float x = e;
    const float neraw = sqrtf(warp_sum(e * e));
    const float ne = fmaxf(neraw, 1e-15f);
    const bool clipped = with_project && (ne > R);
    if (clipped) x = e / ne * R;
    const float x2 = warp_sum(x * x), ax = warp_sum(a * x);
    const float A = 1.f + 2.f * ax + x2, Bc = 1.f - a2;
    const float den_raw = 1.f + 2.f * ax + a2 * x2;
    const float den = fmaxf(den_raw, 1e-15f);
    const float num = A * a + Bc * x;
    const float rr = num / den;
    const float rn = sqrtf(warp_sum(rr * rr));
    // backward
    const float hi = 1.f - 1e-7f;
    const float g_s = dscore[r];
    const float d_rn = (rn < hi && rn > -hi) ? g_s * 2.f / (1.f - rn * rn) : 0.f;
    const float d_r = (rn > 0.f) ? d_rn * rr / rn : 0.f;
    const float d_num = d_r / den;
    const float d_den = (den_raw > 1e-15f) ? -warp_sum(d_r * rr) / den : 0.f;
    const float d_A = warp_sum(d_num * a);
    const float d_ax = 2.f * d_A + 2.f * d_den;
    const float d_x2 = d_A + a2 * d_den;
    const float d_x = Bc * d_num + d_ax * a + 2.f * d_x2 * x;
    float d_e = d_x;
    if (clipped) {
      const float dot = warp_sum(d_x * e);
      d_e = R * (d_x / ne - e * dot / (ne * ne * ne));
    }
    // e = f(n) z, f = tanh(n)/n
    const float f = t / n;
    const float tp = (n < 15.f) ? 1.f - t * t : 0.f;
    const float fp = n_live ? (tp * n - t) / (n * n) : 0.f;
    const float dot2 = warp_sum(d_e * zd);
    const float g = f * d_e + (n_live ? fp * (zd / n) * dot2 : 0.f);
    if (lane < D) dz[r * D + lane] = g;
  }
}

}  // namespace coskad

// =================================================================================================
// Training path: per-layer kernels with train-mode BatchNorm (batch statistics, per GPU -- the
// reference has no SyncBN).  Activations are [B, C, 204] float32 in HBM between kernels: batch
// statistics over (B, T, V) sit between the convolution and the activation, so a layer cannot be
// fused end to end in training.  Reference: models/graph_layers/stsgcn.py:94-156
// forward; the backward kernels implement the analytic gradients of the same graph (autograd
// semantics of nn.Conv2d 1x1, nn.BatchNorm2d(train), nn.PReLU, torch.einsum).
// =================================================================================================
namespace coskad {

constexpr int kTrainThreads = 256;

// ---- graph contraction forward / backward over rows (row = one (window, channel) plane of 204 positions) ------------
// Persistent blocks of 384 threads; one iteration = kCRows = 96 consecutive rows held as planes [row][205] in shared
// memory and pushed through the register-blocked FFMA2 contraction stages of the eval kernel (fused_eval.cuh:
// temporal_stage_c32 / spatial_stage_c32: lane = row % 32, three 32-row groups per lane, weights broadcast from shared
// memory).  The backward pass runs the same stages with transposed weights and accumulates dA / dT as per-thread register
// tiles over the block's rows (one partial block per CTA, summed in a fixed order afterwards).
constexpr int kCRows = kNW * 32;                 // 96
constexpr int kCThreads = 384;
constexpr int kCWarps = kCThreads / 32;
// Two block sizes: NR = 96 rows (one 182 KB block per SM, the C = 32 stages: lane = row % 32) and NR = 48 rows (two 102 KB
// blocks per SM, the C = 16 stages: lane = (row % 16, output half)): with two resident blocks the load -> compute -> store
// phases of one overlap the other's and the FP32 stages see 24 instead of 12 warps per SM.
constexpr int kCRowsSmall = kNW * 16;            // 48
// plane stride (floats): 205 (odd: the 32 lanes = rows of the C = 32 stages hit 32 banks) for the 96-row kernels; 206 for the 48-row
// kernels: the 16 lanes = channels of the C = 16 stages are 14 c mod 32 apart -- 16 distinct banks, the two output halves
// broadcast -- and an EVEN stride makes every row 8-byte aligned, so rows move with 8-byte cp.async / LDS.64 / STG.64: half the
// load / store instructions of these MIO-bound phases (ncu: mio_throttle is their top stall)
__host__ __device__ constexpr int contract_stride(int nr) { return nr == 3 * 16 ? 206 : kCS; }
__host__ __device__ constexpr int contract_smem_bytes(int nr) { return (2 * nr * contract_stride(nr) + kTwFloats + kAwFloats) * 4; }
constexpr int kCSmemBytes = contract_smem_bytes(kCRows);
static_assert(kCSmemBytes <= 227 * 1024, "contraction kernels: shared memory plan exceeds 227 KB");
static_assert(2 * (contract_smem_bytes(kCRowsSmall) + 1024) <= 227 * 1024, "two 48-row blocks per SM");
// the C = 16 contraction stages of fused_eval.cuh (lane = (channel, output half), 3 row groups per lane, FFMA2) with the plane
// stride as a template parameter; src == dst is allowed
template <int NWARPS, int CS>
__device__ __forceinline__ void temporal_stage_c16_cs(const float* src, float* dst, const float* Tw, int warp, int lane) {
  const int c = lane & 15, q0 = (lane >> 4) * 6;
  for (int v = warp; v < kV; v += NWARPS) {
    unsigned long long acc[kNW][3];
#pragma unroll
    for (int n = 0; n < kNW; ++n)
#pragma unroll
      for (int q = 0; q < 3; ++q) acc[n][q] = 0ull;
    const float* s_ = src + c * CS + v;
    const float* w = Tw + v * (kT * kT) + q0;
#pragma unroll 4
    for (int t = 0; t < kT; ++t) {
      unsigned long long x[kNW];
#pragma unroll
      for (int n = 0; n < kNW; ++n) x[n] = dup2(s_[n * 16 * CS + t * kV]);
      const unsigned long long w01 = *reinterpret_cast<const unsigned long long*>(w + t * kT);
      const unsigned long long w23 = *reinterpret_cast<const unsigned long long*>(w + t * kT + 2);
      const unsigned long long w45 = *reinterpret_cast<const unsigned long long*>(w + t * kT + 4);
#pragma unroll
      for (int n = 0; n < kNW; ++n) {
        ffma2(acc[n][0], x[n], w01); ffma2(acc[n][1], x[n], w23); ffma2(acc[n][2], x[n], w45);
      }
    }
    __syncwarp();
#pragma unroll
    for (int n = 0; n < kNW; ++n) {
      float* d = dst + (n * 16 + c) * CS + v;
#pragma unroll
      for (int q = 0; q < 3; ++q) { d[(q0 + 2 * q) * kV] = lo2(acc[n][q]); d[(q0 + 2 * q + 1) * kV] = hi2(acc[n][q]); }
    }
  }
}
template <int NWARPS, int CS>
__device__ __forceinline__ void spatial_stage_c16_cs(float* buf, const float* Aw, int warp, int lane) {
  const int c = lane & 15, h = lane >> 4;
  for (int t = warp; t < kT; t += NWARPS) {
    unsigned long long acc[kNW][4];
    float acc8[kNW];
#pragma unroll
    for (int n = 0; n < kNW; ++n) {
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[n][j] = 0ull;
      acc8[n] = 0.f;
    }
    const float* s_ = buf + c * CS + t * kV;
    const float* a = Aw + t * (kV * kAW) + 8 * h;
#pragma unroll 1
    for (int v = 0; v < kV; ++v) {
      float g[kNW];
#pragma unroll
      for (int n = 0; n < kNW; ++n) g[n] = s_[n * 16 * CS + v];
      const ulonglong2 w0 = *reinterpret_cast<const ulonglong2*>(a + v * kAW);
      const ulonglong2 w1 = *reinterpret_cast<const ulonglong2*>(a + v * kAW + 4);
      const float w8 = a[v * kAW + 8];
#pragma unroll
      for (int n = 0; n < kNW; ++n) {
        const unsigned long long g2 = dup2(g[n]);
        ffma2(acc[n][0], g2, w0.x); ffma2(acc[n][1], g2, w0.y);
        ffma2(acc[n][2], g2, w1.x); ffma2(acc[n][3], g2, w1.y);
        acc8[n] = fmaf(g[n], w8, acc8[n]);
      }
    }
    __syncwarp();
#pragma unroll
    for (int n = 0; n < kNW; ++n) {
      float* d = buf + (n * 16 + c) * CS + t * kV;
#pragma unroll
      for (int j = 0; j < 4; ++j) { d[8 * h + 2 * j] = lo2(acc[n][j]); d[8 * h + 2 * j + 1] = hi2(acc[n][j]); }
      if (h == 1) d[16] = acc8[n];
    }
  }
}
template <int NR> struct ContractStages;
template <> struct ContractStages<kCRows> {
  static __device__ __forceinline__ void temporal(const float* src, float* dst, const float* Tw, int warp, int lane) {
    temporal_stage_c32<kCThreads / 32>(src, dst, Tw, warp, lane);
  }
  static __device__ __forceinline__ void spatial(float* buf, const float* Aw, int warp, int lane) {
    spatial_stage_c32<kCThreads / 32>(buf, Aw, warp, lane);
  }
};
template <> struct ContractStages<kCRowsSmall> {
  static __device__ __forceinline__ void temporal(const float* src, float* dst, const float* Tw, int warp, int lane) {
    temporal_stage_c16_cs<kCThreads / 32, contract_stride(kCRowsSmall)>(src, dst, Tw, warp, lane);
  }
  static __device__ __forceinline__ void spatial(float* buf, const float* Aw, int warp, int lane) {
    spatial_stage_c16_cs<kCThreads / 32, contract_stride(kCRowsSmall)>(buf, Aw, warp, lane);
  }
};
constexpr int kContractPart = kT * kV * kV + kV * kT * kT;     // floats of one block's [dA | dT] partial

// Row-block copies, one warp per row (rows warp, warp + 12, ..), lanes = positions p = lane + 32 k: one pointer per row and
// a few instructions per element (a flat element index cost ~20 integer instructions per 4-byte copy).
constexpr int kRowIters = (kP + 31) / 32;                         // 7
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  unsigned s_ = static_cast<unsigned>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s_), "l"(gmem) : "memory");
}
template <int NR>
__device__ __forceinline__ void contract_load_rows(float* dst, const float* __restrict__ src, int64_t r0, int nr, int tid) {
  constexpr int CS = contract_stride(NR);
  const int warp = tid >> 5, lane = tid & 31;
  for (int row = warp; row < NR; row += kCWarps) {
    const float* s = src + (r0 + (row < nr ? row : nr - 1)) * kP;   // ragged last block: replicate the last row, never stored
    float* d = dst + row * CS;
    if constexpr (CS % 2 == 0) {                                    // 8-byte copies: rows are 8-byte aligned on both sides
#pragma unroll
      for (int k = 0; k < (kP / 2 + 31) / 32; ++k) {
        const int p = 2 * (lane + 32 * k);
        if (p < kP) cp_async8(d + p, s + p);
      }
    } else {
#pragma unroll
      for (int k = 0; k < kRowIters; ++k) {
        const int p = lane + 32 * k;
        if (p < kP) cp_async4(d + p, s + p);
      }
    }
  }
}
// L2 prefetch of rows [r0, r0 + nr) of a [R][204] array (one 128-byte line per request): issued one block ahead, so that the
// later cp.async / LDG of these rows find them in L2 (the kernels have no shared memory or registers left for a deeper pipeline)
__device__ __forceinline__ void contract_prefetch_rows(const float* __restrict__ src, int64_t r0, int nr, int tid) {
  const char* base = reinterpret_cast<const char*>(src + r0 * kP);
  const int nlines = (nr * kP * 4 + 127) / 128;
  for (int i = tid; i < nlines; i += kCThreads) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + static_cast<int64_t>(i) * 128));
}
// the T / A operator tables -> shared memory with all loads of a thread in flight together (as a load -> store loop the
// prologue exposed one L2 / HBM round trip per element: 12-17 % of the kernels' time, ncu source page)
template <class F>
__device__ __forceinline__ void contract_fill_table(float* dst, int n, int tid, F value_ptr) {
  constexpr int kBatch = 6;
  for (int i0 = tid; i0 < n; i0 += kCThreads * kBatch) {
    float v[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int i = i0 + u * kCThreads;
      const float* p = value_ptr(i < n ? i : n - 1);
      v[u] = p ? __ldg(p) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int i = i0 + u * kCThreads;
      if (i < n) dst[i] = v[u];
    }
  }
}
// dst rows [r0, r0+nr) = planes (+ add, nullable); the loads of `add` are batched four rows at a time
template <int NR>
__device__ __forceinline__ void contract_store_rows(float* __restrict__ dst, const float* planes, const float* __restrict__ add,
                                                    int64_t r0, int nr, int tid) {
  constexpr int CS = contract_stride(NR);
  const int warp = tid >> 5, lane = tid & 31;
  constexpr int kRB = 4;
  static_assert(NR % (kCWarps * kRB) == 0, "row-block store plan");
  if constexpr (CS % 2 == 0) {                                      // 8-byte path: LDS.64 + (LDG.64) + STG.64, 4 iterations per row
    constexpr int kIt = (kP / 2 + 31) / 32;
    for (int rb = warp; rb < NR; rb += kCWarps * kRB) {
      float2 a[kRB][kIt];
#pragma unroll
      for (int j = 0; j < kRB; ++j) {
        const int row = rb + j * kCWarps;
#pragma unroll
        for (int k = 0; k < kIt; ++k) {
          const int p = 2 * (lane + 32 * k);
          a[j][k] = (add != nullptr && row < nr && p < kP) ? __ldg(reinterpret_cast<const float2*>(add + (r0 + row) * kP + p)) : make_float2(0.f, 0.f);
        }
      }
#pragma unroll
      for (int j = 0; j < kRB; ++j) {
        const int row = rb + j * kCWarps;
        if (row < nr) {
          float* d = dst + (r0 + row) * kP;
          const float* s = planes + row * CS;
#pragma unroll
          for (int k = 0; k < kIt; ++k) {
            const int p = 2 * (lane + 32 * k);
            if (p < kP) {
              const float2 v = *reinterpret_cast<const float2*>(s + p);
              *reinterpret_cast<float2*>(d + p) = make_float2(v.x + a[j][k].x, v.y + a[j][k].y);
            }
          }
        }
      }
    }
    return;
  }
  for (int rb = warp; rb < NR; rb += kCWarps * kRB) {
    float a[kRB][kRowIters];
#pragma unroll
    for (int j = 0; j < kRB; ++j) {
      const int row = rb + j * kCWarps;
#pragma unroll
      for (int k = 0; k < kRowIters; ++k) {
        const int p = lane + 32 * k;
        a[j][k] = (add != nullptr && row < nr && p < kP) ? __ldg(add + (r0 + row) * kP + p) : 0.f;
      }
    }
#pragma unroll
    for (int j = 0; j < kRB; ++j) {
      const int row = rb + j * kCWarps;
      if (row < nr) {
        float* d = dst + (r0 + row) * kP;
        const float* s = planes + row * CS;
#pragma unroll
        for (int k = 0; k < kRowIters; ++k) {
          const int p = lane + 32 * k;
          if (p < kP) d[p] = s[p] + a[j][k];
        }
      }
    }
  }
}

// G1 = einsum('nctv,vtq->ncqv', X, T); G = einsum('nctv,tvw->nctw', G1, A)      (stsgcn.py:154-155)
template <int NR>
__global__ void __launch_bounds__(kCThreads, NR == kCRows ? 1 : 2) train_contract_fwd_kernel(const float* __restrict__ X, const float* __restrict__ A,
                                                                         const float* __restrict__ T, int64_t R,
                                                                         float* __restrict__ G1, float* __restrict__ G) {
  extern __shared__ __align__(128) float csm[];
  constexpr int CS = contract_stride(NR);
  float* Xs = csm;
  float* Gs = Xs + NR * CS;
  float* Ts = Gs + NR * CS;
  float* As = Ts + kTwFloats;                      // rows padded 17 -> kAW
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t nblk = (R + NR - 1) / NR;
  if (static_cast<int64_t>(blockIdx.x) < nblk) {                  // the first block's rows fly while the tables are filled
    const int64_t r0 = static_cast<int64_t>(blockIdx.x) * NR;
    contract_load_rows<NR>(Xs, X, r0, static_cast<int>(R - r0 < NR ? R - r0 : NR), tid);
  }
  cp_async_commit();
  contract_fill_table(Ts, kTwFloats, tid, [&](int i) { return T + i; });
  contract_fill_table(As, kAwFloats, tid, [&](int i) -> const float* { const int w = i % kAW, tv = i / kAW; return (w < kV) ? A + tv * kV + w : nullptr; });
  for (int64_t blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const int64_t r0 = blk * NR;
    const int nr = static_cast<int>(R - r0 < NR ? R - r0 : NR);
    cp_async_wait_all();
    __syncthreads();
    ContractStages<NR>::temporal(Xs, Gs, Ts, warp, lane);
    __syncthreads();
    {   // Xs is dead: prefetch the next block's rows while this one finishes
      const int64_t nb = blk + gridDim.x;
      if (nb < nblk) { const int64_t n0 = nb * NR; contract_load_rows<NR>(Xs, X, n0, static_cast<int>(R - n0 < NR ? R - n0 : NR), tid); }
      cp_async_commit();
    }
    contract_store_rows<NR>(G1, Gs, nullptr, r0, nr, tid);
    __syncthreads();
    ContractStages<NR>::spatial(Gs, As, warp, lane);
    __syncthreads();
    contract_store_rows<NR>(G, Gs, nullptr, r0, nr, tid);
  }
  cp_async_wait_all();
}

// dG1[t,v] = sum_w dG[t,w] A[t,v,w];  dX[t,v] = dXres[t,v] + sum_q dG1[q,v] T[v,t,q]
// dA[t,v,w] += G1[t,v] dG[t,w];       dT[v,t,q] += X[t,v] dG1[q,v]      (summed over rows)
template <int NR>
__global__ void __launch_bounds__(kCThreads, NR == kCRows ? 1 : 2) train_contract_bwd_kernel(
    const float* __restrict__ dG, const float* __restrict__ dXres, const float* __restrict__ X, const float* __restrict__ G1,
    const float* __restrict__ A, const float* __restrict__ T, int64_t R, float* __restrict__ dX, float* __restrict__ part) {
  extern __shared__ __align__(128) float csm[];
  constexpr int CS = contract_stride(NR);
  float* P0 = csm;                                 // dG -> dG1 (in place)
  float* P1 = P0 + NR * CS;                        // G1, then X, then dX
  float* Tt = P1 + NR * CS;                        // Tt[v][q][t] = T[v][t][q]
  float* At = Tt + kTwFloats;                      // At[t][w][v] = A[t][v][w], rows padded 17 -> kAW
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  contract_fill_table(Tt, kTwFloats, tid, [&](int i) { const int v = i / (kT * kT), q = (i / kT) % kT, t = i % kT; return T + v * (kT * kT) + t * kT + q; });
  contract_fill_table(At, kAwFloats, tid, [&](int i) -> const float* {
    const int v = i % kAW, tw = i / kAW, t = tw / kV, w = tw % kV;
    return (v < kV) ? A + (t * kV + v) * kV + w : nullptr;
  });
  // dA tiles: thread < 300 owns (t, 4 v, 4 w); dT tiles: thread < 306 owns (v, 4 t, 2 q)
  const bool hasA = tid < kT * 25, hasT = tid < kV * 18;
  const int at = tid / 25, avg = (tid % 25) / 5, awg = tid % 5;
  const int tv = tid / 18, ttg = (tid % 18) / 6, tqg = tid % 6;
  int aiv[4], aiw[4], tit[4], tiq[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    aiv[i] = at * kV + (4 * avg + i < kV ? 4 * avg + i : kV - 1);
    aiw[i] = at * kV + (4 * awg + i < kV ? 4 * awg + i : kV - 1);
    tit[i] = (4 * ttg + i) * kV + tv;
  }
  tiq[0] = (2 * tqg) * kV + tv; tiq[1] = (2 * tqg + 1) * kV + tv;
  float accA[4][4], accT[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    accT[i][0] = accT[i][1] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) accA[i][j] = 0.f;
  }
  const int64_t nblk = (R + NR - 1) / NR;
  for (int64_t blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const int64_t r0 = blk * NR;
    const int nr = static_cast<int>(R - r0 < NR ? R - r0 : NR);
    __syncthreads();
    contract_load_rows<NR>(P0, dG, r0, nr, tid);
    contract_load_rows<NR>(P1, G1, r0, nr, tid);
    cp_async_commit();
    {   // this block's late operands (X, dXres) and the next block's early ones (dG, G1): into L2 now, used 1-4 stages later
      contract_prefetch_rows(X, r0, nr, tid);
      if (dXres != nullptr) contract_prefetch_rows(dXres, r0, nr, tid);
      const int64_t nb = blk + gridDim.x;
      if (nb < nblk) {
        const int64_t n0 = nb * NR;
        const int nn = static_cast<int>(R - n0 < NR ? R - n0 : NR);
        contract_prefetch_rows(dG, n0, nn, tid);
        contract_prefetch_rows(G1, n0, nn, tid);
      }
    }
    cp_async_wait_all();
    __syncthreads();
    if (hasA) {
      for (int r = 0; r < nr; ++r) {
        float g[4], d[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { g[i] = P1[r * CS + aiv[i]]; d[i] = P0[r * CS + aiw[i]]; }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) accA[i][j] = fmaf(g[i], d[j], accA[i][j]);
      }
    }
    __syncthreads();
    contract_load_rows<NR>(P1, X, r0, nr, tid);          // G1 is dead
    cp_async_commit();
    ContractStages<NR>::spatial(P0, At, warp, lane);  // dG -> dG1 in place
    cp_async_wait_all();
    __syncthreads();
    if (hasT) {
      for (int r = 0; r < nr; ++r) {
        float x[4], d[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = P1[r * CS + tit[i]];
        d[0] = P0[r * CS + tiq[0]]; d[1] = P0[r * CS + tiq[1]];
#pragma unroll
        for (int i = 0; i < 4; ++i) { accT[i][0] = fmaf(x[i], d[0], accT[i][0]); accT[i][1] = fmaf(x[i], d[1], accT[i][1]); }
      }
    }
    __syncthreads();
    ContractStages<NR>::temporal(P0, P1, Tt, warp, lane);      // dG1 -> temporal^T -> P1 (X is dead)
    __syncthreads();
    contract_store_rows<NR>(dX, P1, dXres, r0, nr, tid);
  }
  // per-block partial sums [dA (T*V*V) | dT (V*T*T)]: every element has exactly one owner thread, so these are plain stores;
  // the blocks' partials are added in a fixed order by partial_sum_kernel (no floating-point atomics: bit-reproducible)
  float* pA = part + static_cast<int64_t>(blockIdx.x) * kContractPart;
  float* pT = pA + kT * kV * kV;
  if (hasA) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (4 * avg + i < kV && 4 * awg + j < kV) pA[(at * kV + 4 * avg + i) * kV + 4 * awg + j] = accA[i][j];
  }
  if (hasT) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) pT[tv * (kT * kT) + (4 * ttg + i) * kT + 2 * tqg + j] = accT[i][j];
  }
}

// mean / invstd from the sums, running-statistics update (momentum 0.1, unbiased variance), nn.BatchNorm2d semantics
__global__ void train_bn_finalize_kernel(const double* __restrict__ stats, double N, int CO, float eps, float momentum,
                                         float* rm1, float* rv1, float* rm2, float* rv2, float* __restrict__ mi,
                                         int64_t* nbt1, int64_t* nbt2) {
  const int co = blockIdx.x * blockDim.x + threadIdx.x;
  if (co == 0) {                                   // nn.BatchNorm2d.num_batches_tracked += 1 (was two one-element torch kernels per layer)
    if (nbt1) *nbt1 += 1;
    if (nbt2) *nbt2 += 1;
  }
  if (co >= CO) return;
  for (int br = 0; br < 2; ++br) {
    const double mean = stats[(2 * br) * CO + co] / N;
    double var = stats[(2 * br + 1) * CO + co] / N - mean * mean;
    if (var < 0.0) var = 0.0;
    mi[(2 * br) * CO + co] = static_cast<float>(mean);
    mi[(2 * br + 1) * CO + co] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    float* rm = br ? rm2 : rm1;
    float* rv = br ? rv2 : rv1;
    if (rm) rm[co] = (1.f - momentum) * rm[co] + momentum * static_cast<float>(mean);
    if (rv) rv[co] = (1.f - momentum) * rv[co] + momentum * static_cast<float>(var * N / (N - 1.0));
  }
}

// second stage of the BatchNorm statistics (per-CTA partials [nparts][4*CO] of tc_mix_fwd_kernel, fixed order) fused with
// train_bn_finalize_kernel: one launch instead of partial_sum + finalize (+ the zero fill of the intermediate sums); the
// arithmetic is the same as the finalize kernel's.  Block x owns 8 channels: lane = quantity q * 8 + channel (q = sum y1, sum y1^2,
// sum y2, sum y2^2), 32 rows of threads share the partials (row y adds partials y, y + 32, .. in ascending order, 8 loads in flight),
// the rows meet in shared memory in a fixed order, and the quantities of a channel meet through shuffles.  (A first version with
// 32 channels per block walked the four quantities one after the other on 1-2 blocks: 13.7 us per launch.)
constexpr int kBsfRows = 32;
__global__ void __launch_bounds__(32 * kBsfRows) train_bn_stats_finalize_kernel(
    const float* __restrict__ part, int nparts, double N, int CO, float eps, float momentum, float* rm1, float* rv1, float* rm2,
    float* rv2, float* __restrict__ mi, int64_t* nbt1, int64_t* nbt2) {
  __shared__ double sh[kBsfRows][33];
  const int lane = threadIdx.x, y = threadIdx.y;
  const int q = lane >> 3, co = blockIdx.x * 8 + (lane & 7);
  const bool ok = co < CO;
  const int64_t i = static_cast<int64_t>(q) * CO + co, stride = 4 * CO;
  double s = 0.0;
  if (ok) {
    int j = y;
    for (; j + 7 * kBsfRows < nparts; j += 8 * kBsfRows) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = part[static_cast<int64_t>(j + u * kBsfRows) * stride + i];
#pragma unroll
      for (int u = 0; u < 8; ++u) s += static_cast<double>(v[u]);
    }
    for (; j < nparts; j += kBsfRows) s += static_cast<double>(part[static_cast<int64_t>(j) * stride + i]);
  }
  sh[y][lane] = s;
  __syncthreads();
  if (y != 0) return;
  double t = 0.0;
#pragma unroll
  for (int r = 0; r < kBsfRows; ++r) t += sh[r][lane];
  // lane (q, c): q even holds a sum, q odd the matching sum of squares, 8 lanes further
  const double sq = __shfl_down_sync(0xffffffffu, t, 8);
  if (blockIdx.x == 0 && lane == 0) {
    if (nbt1) *nbt1 += 1;
    if (nbt2) *nbt2 += 1;
  }
  if (!ok || (q & 1)) return;
  const int br = q >> 1;
  const double mean = t / N;
  double var = sq / N - mean * mean;
  if (var < 0.0) var = 0.0;
  mi[(2 * br) * CO + co] = static_cast<float>(mean);
  mi[(2 * br + 1) * CO + co] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  float* rm = br ? rm2 : rm1;
  float* rv = br ? rv2 : rv1;
  if (rm) rm[co] = (1.f - momentum) * rm[co] + momentum * static_cast<float>(mean);
  if (rv) rv[co] = (1.f - momentum) * rv[co] + momentum * static_cast<float>(var * N / (N - 1.0));
}

// BN1(y1) + BN2(y2) with a fixed operation order (explicit rounding intrinsics: no re-contraction), so that the
// forward and the two backward kernels see bit-identical pre-activations and hence the same PReLU branch.
__device__ __forceinline__ float bn_pre(float y1, float y2, float m1, float i1, float m2, float i2, float g1, float be1,
                                        float g2, float be2, float& h1, float& h2) {
  h1 = __fmul_rn(__fsub_rn(y1, m1), i1);
  h2 = __fmul_rn(__fsub_rn(y2, m2), i2);
  return __fadd_rn(__fmaf_rn(h1, g1, be1), __fmaf_rn(h2, g2, be2));
}
// nn.PReLU: positive branch iff x > 0 (ATen prelu backward uses the same strict comparison)
__device__ __forceinline__ float prelu_strict(float v, float a) { return v > 0.f ? v : a * v; }

// out = PReLU(BN1(y1) + BN2(y2)).  One warp per row (b, co) of 204 positions = 51 float4: the per-channel constants are
// fetched once per row instead of once per element, 16-byte loads and stores.
constexpr int kRowV4 = kP / 4;                                   // 51
__global__ void train_bn_prelu_fwd_kernel(const float* __restrict__ y1, const float* __restrict__ y2,
                                          const float* __restrict__ mi, const float* __restrict__ g1,
                                          const float* __restrict__ be1, const float* __restrict__ g2,
                                          const float* __restrict__ be2, const float* __restrict__ slope, int64_t B,
                                          int CO, float* __restrict__ out) {
  const int64_t rows = B * CO;
  const float a = slope[0];
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * wpb + (threadIdx.x >> 5); r < rows; r += static_cast<int64_t>(gridDim.x) * wpb) {
    const int co = static_cast<int>(r % CO);
    const float m1 = mi[co], i1 = mi[CO + co], m2 = mi[2 * CO + co], i2 = mi[3 * CO + co];
    const float ga1 = g1[co], bb1 = be1[co], ga2 = g2[co], bb2 = be2[co];
    const float4* a1 = reinterpret_cast<const float4*>(y1 + r * kP);
    const float4* a2 = reinterpret_cast<const float4*>(y2 + r * kP);
    float4* o = reinterpret_cast<float4*>(out + r * kP);
    for (int i = lane; i < kRowV4; i += 32) {
      const float4 u = a1[i], v = a2[i];
      float h1, h2;
      float4 w;
      w.x = prelu_strict(bn_pre(u.x, v.x, m1, i1, m2, i2, ga1, bb1, ga2, bb2, h1, h2), a);
      w.y = prelu_strict(bn_pre(u.y, v.y, m1, i1, m2, i2, ga1, bb1, ga2, bb2, h1, h2), a);
      w.z = prelu_strict(bn_pre(u.z, v.z, m1, i1, m2, i2, ga1, bb1, ga2, bb2, h1, h2), a);
      w.w = prelu_strict(bn_pre(u.w, v.w, m1, i1, m2, i2, ga1, bb1, ga2, bb2, h1, h2), a);
      o[i] = w;
    }
  }
}

// backward reductions: red (double) [3*CO + 1] = sum ds, sum ds*yhat1, sum ds*yhat2 per channel, then d slope
// grid (CO, NB): block (co, j) walks the rows (b, co) of its batch slice
__global__ void train_bn_prelu_bwd_reduce_kernel(const float* __restrict__ dout, const float* __restrict__ y1,
                                                 const float* __restrict__ y2, const float* __restrict__ mi,
                                                 const float* __restrict__ g1, const float* __restrict__ be1,
                                                 const float* __restrict__ g2, const float* __restrict__ be2,
                                                 const float* __restrict__ slope, int64_t B, int CO,
                                                 float* __restrict__ part /*[gridDim.y][4][CO]*/) {
  const int co = blockIdx.x;
  const float a = slope[0];
  const float m1 = mi[co], i1 = mi[CO + co], m2 = mi[2 * CO + co], i2 = mi[3 * CO + co];
  const float ga1 = g1[co], ga2 = g2[co], bb1 = be1[co], bb2 = be2[co];
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, sa = 0.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int64_t b = static_cast<int64_t>(blockIdx.y) * nwarp + warp; b < B; b += static_cast<int64_t>(gridDim.y) * nwarp) {
    const int64_t base = (b * CO + co) * kP;
    const float4* a1 = reinterpret_cast<const float4*>(y1 + base);
    const float4* a2 = reinterpret_cast<const float4*>(y2 + base);
    const float4* dd = reinterpret_cast<const float4*>(dout + base);
    for (int i = lane; i < kRowV4; i += 32) {
      const float4 u = a1[i], v = a2[i], d4 = dd[i];
      const float uu[4] = {u.x, u.y, u.z, u.w}, vv[4] = {v.x, v.y, v.z, v.w}, dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float h1, h2;
        const float pre = bn_pre(uu[k], vv[k], m1, i1, m2, i2, ga1, bb1, ga2, bb2, h1, h2);
        const float d = dv[k];
        const float ds = pre > 0.f ? d : a * d;
        s0 += ds; s1 = fmaf(ds, h1, s1); s2 = fmaf(ds, h2, s2);
        if (!(pre > 0.f)) sa = fmaf(d, pre, sa);
      }
    }
  }
  __shared__ float sh[4][8];
  s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2); sa = warp_sum(sa);
  if (lane == 0) { sh[0][warp] = s0; sh[1][warp] = s1; sh[2][warp] = s2; sh[3][warp] = sa; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
    for (int w = 0; w < nwarp; ++w) t += static_cast<double>(sh[threadIdx.x][w]);
    part[(static_cast<int64_t>(blockIdx.y) * 4 + threadIdx.x) * CO + co] = static_cast<float>(t);
  }
}
// second stage of the reduction above, fixed order (partial_sum_block: 32 x kPsRows threads per 32 elements).  Blocks
// 0 .. ceil(3CO/32)-1: red[i] += sum over the batch slices of part[y][q][co], i = q*CO + co < 3CO.  Last block: the PReLU-slope
// sums, first per channel over the slices, then over the channels in ascending order -> red[3CO].
// GRADS = false: red += the sums (the documented contract of coskad_train_bn_prelu_bwd: red is zeroed by the caller).
// GRADS = true (coskad_train_bn_prelu_bwd_grads): red = the sums (no zero fill needed) and the parameter gradients are
// accumulated in the same launch: d beta1 = d beta2 += red[0..CO), d gamma1 += red[CO..2CO), d gamma2 += red[2CO..3CO),
// d slope += red[3CO] (what train_bn_param_grads_kernel does as a launch of its own).
template <bool GRADS>
__global__ void train_bn_prelu_bwd_reduce_final_kernel(const float* __restrict__ part, int nb, int CO, double* red, float* dg1,
                                                       float* dbe1, float* dg2, float* dbe2, float* dslope) {
  __shared__ double sh[kPsRows][33];
  const int nmain = (3 * CO + 31) / 32;
  if (static_cast<int>(blockIdx.x) < nmain) {
    const int i = blockIdx.x * 32 + threadIdx.x;
    const double t = partial_sum_block(part, nb, 4 * CO, i, i < 3 * CO, sh);
    if (threadIdx.y == 0 && i < 3 * CO) {
      if (GRADS) {
        red[i] = t;
        const float f = static_cast<float>(t);
        const int q = i / CO, co = i - q * CO;
        if (q == 0) { if (dbe1) dbe1[co] += f; if (dbe2) dbe2[co] += f; }
        else if (q == 1) { if (dg1) dg1[co] += f; }
        else if (dg2) dg2[co] += f;
      } else {
        red[i] += t;
      }
    }
    return;
  }
  __shared__ double ch[64];
  for (int c0 = 0; c0 < CO; c0 += 32) {
    const int co = c0 + threadIdx.x;
    const double t = partial_sum_block(part, nb, 4 * CO, 3 * CO + co, co < CO, sh);
    if (threadIdx.y == 0 && co < CO) ch[co] = t;
    __syncthreads();
  }
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    double s = 0.0;
    for (int co = 0; co < CO; ++co) s += ch[co];
    if (GRADS) { red[3 * CO] = s; if (dslope) dslope[0] += static_cast<float>(s); }
    else red[3 * CO] += s;
  }
}

// parameter gradients of the two BatchNorms and the PReLU from the float64 sums `red` (train_bn_prelu_bwd_reduce_*):
// d beta1 = d beta2 += red[0..CO), d gamma1 += red[CO..2CO), d gamma2 += red[2CO..3CO), d slope += red[3CO]
// (accumulating: the destinations are the zeroed .grad views of the flat gradient bucket, or zeroed temporaries)
__global__ void train_bn_param_grads_kernel(const double* __restrict__ red, int CO, float* dg1, float* dbe1, float* dg2,
                                            float* dbe2, float* dslope) {
  const int co = blockIdx.x * blockDim.x + threadIdx.x;
  if (co < CO) {
    const float db = static_cast<float>(red[co]);
    if (dbe1) dbe1[co] += db;
    if (dbe2) dbe2[co] += db;
    if (dg1) dg1[co] += static_cast<float>(red[CO + co]);
    if (dg2) dg2[co] += static_cast<float>(red[2 * CO + co]);
  }
  if (co == 0 && dslope) dslope[0] += static_cast<float>(red[3 * CO]);
}

// dy1 = g1*is1*(ds - mean(ds) - yhat1*mean(ds*yhat1)), dy2 likewise (BatchNorm train backward); one warp per row, float4
__global__ void train_bn_prelu_bwd_apply_kernel(const float* __restrict__ dout, const float* __restrict__ y1,
                                                const float* __restrict__ y2, const float* __restrict__ mi,
                                                const float* __restrict__ g1, const float* __restrict__ be1,
                                                const float* __restrict__ g2, const float* __restrict__ be2,
                                                const float* __restrict__ slope, const double* __restrict__ red,
                                                int64_t B, int CO, float* __restrict__ dy1, float* __restrict__ dy2) {
  const int64_t rows = B * CO;
  const float a = slope[0];
  const double N = static_cast<double>(B) * kP;
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * wpb + (threadIdx.x >> 5); r < rows; r += static_cast<int64_t>(gridDim.x) * wpb) {
    const int co = static_cast<int>(r % CO);
    const float m1 = mi[co], i1 = mi[CO + co], m2 = mi[2 * CO + co], i2 = mi[3 * CO + co];
    const float ga1 = g1[co], bb1 = be1[co], ga2 = g2[co], bb2 = be2[co];
    const float mds = static_cast<float>(red[co] / N);
    const float mh1 = static_cast<float>(red[CO + co] / N), mh2 = static_cast<float>(red[2 * CO + co] / N);
    const float sc1 = ga1 * i1, sc2 = ga2 * i2;
    const float4* a1 = reinterpret_cast<const float4*>(y1 + r * kP);
    const float4* a2 = reinterpret_cast<const float4*>(y2 + r * kP);
    const float4* dd = reinterpret_cast<const float4*>(dout + r * kP);
    float4* o1 = reinterpret_cast<float4*>(dy1 + r * kP);
    float4* o2 = reinterpret_cast<float4*>(dy2 + r * kP);
    for (int i = lane; i < kRowV4; i += 32) {
      const float4 u = a1[i], v = a2[i], d4 = dd[i];
      const float uu[4] = {u.x, u.y, u.z, u.w}, vv[4] = {v.x, v.y, v.z, v.w}, dv[4] = {d4.x, d4.y, d4.z, d4.w};
      float r1[4], r2[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float h1, h2;
        const float pre = bn_pre(uu[k], vv[k], m1, i1, m2, i2, ga1, bb1, ga2, bb2, h1, h2);
        const float ds = pre > 0.f ? dv[k] : a * dv[k];
        r1[k] = sc1 * (ds - mds - h1 * mh1);
        r2[k] = sc2 * (ds - mds - h2 * mh2);
      }
      o1[i] = make_float4(r1[0], r1[1], r1[2], r1[3]);
      o2[i] = make_float4(r2[0], r2[1], r2[2], r2[3]);
    }
  }
}

// dW1[co,ci] += sum_e dy1[e,co] G[e,ci]; db1[co] += sum_e dy1[e,co]; same for the residual branch (e = (b,p)).
// A split-K "GEMM" with tiny M x N (<= 64 x 64) and K = B*204: every block walks chunks of WC elements e of all channels,
// staged in shared memory as rows [c][WC + 4] (row stride = 4 mod 32 banks: LDS.128 of 8 consecutive rows is conflict free)
// with 16-byte asynchronous copies (4 consecutive positions never straddle a window: 204 = 4 * 51) into two buffers, the
// next chunk in flight while this one is consumed.  Every thread owns a 4 x 4 tile of (co, ci) pairs with STRIDED members
// (co = cot + i*NOT, ci = cit + j*NCT, so that the lanes of a quarter warp read consecutive rows) and, when there are fewer
// than 256 tiles, one of the k-slices of the chunk; 16 LDS.128 feed 128 FMAs.  The bias gradients (row sums of dy) are a
// separate cooperative pass over the staged rows (inside the tile loop every thread of a tile row repeated them: 32 extra
// FADD per 64 FFMA2).  At the end the k-slices of a block are summed through shared memory and the block issues ONE
// atomicAdd per weight (with one per thread the 2->32 layer sent 1.2 M atomics to 128 addresses: 178 us, ncu).
// WC = 128 for <= 96 staged rows, 64 for the 64-channel layers: two buffers of either fit two blocks per SM.
template <int WC>
__global__ void __launch_bounds__(kTrainThreads) train_mix_bwd_weight_kernel(
    const float* __restrict__ dy1, const float* __restrict__ dy2, const float* __restrict__ G, const float* __restrict__ X,
    int64_t B, int CI, int CO, float* dW1, float* db1, float* dW2, float* db2) {
  extern __shared__ __align__(128) float sm[];
  constexpr int WCS = WC + 4;
  constexpr int kV4 = WC / 4;                              // float4 per staged row
  const int buf_floats = 2 * (CO + CI) * WCS;              // rows: [dy1 CO | dy2 CO | G CI | X CI]
  const int TCO = CO < 4 ? CO : 4, TCI = CI < 4 ? CI : 4;
  const int NOT = CO / TCO, NCT = CI / TCI;              // tiles along co / ci
  const int ntile = NOT * NCT;                           // 8 .. 128 for every configured layer
  const int nks = kTrainThreads / ntile;                 // k-slices (>= 2)
  const int kslice = WC / nks;                           // multiple of 4 (host-checked)
  const int tile = threadIdx.x % ntile, ks = threadIdx.x / ntile;
  const int cit = tile % NCT, cot = tile / NCT;
  // accumulators packed as (sum over even k, sum over odd k): both FFMA2 operands are natural halves of the float4 loads
  unsigned long long q1[4][4], q2[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { q1[i][j] = 0ull; q2[i][j] = 0ull; }
  // bias pass: thread -> (row of [dy1 rows | dy2 rows], every tpr-th float4 of the chunk); 2*CO <= 128 rows
  const int tpr = kTrainThreads / (2 * CO) < 32 ? kTrainThreads / (2 * CO) : 32;       // power of two
  const int brow = threadIdx.x / tpr, bk4 = threadIdx.x % tpr;                        // rows >= 2*CO: idle threads
  const bool want_bias = (db1 != nullptr || db2 != nullptr) && brow < 2 * CO;
  float bsum = 0.f;
  const int64_t E = B * kP;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * WC;
  // elements e = e0 + 4 k4 .. + 3 of channel row c live at ((b*C + c)*204 + p): one division per thread and chunk
  auto stage = [&](int64_t e0, float* buf) {
    const int k4 = threadIdx.x % kV4;
    const int64_t e = e0 + 4 * k4;
    const int64_t b = e / kP;
    const int p = static_cast<int>(e - b * kP);
    const bool valid = e < E;                              // E % 4 == 0: a float4 is entirely in or out
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = threadIdx.x / kV4; c < CO + CI; c += kTrainThreads / kV4) {
      const bool is_dy = c < CO;
      const int r = is_dy ? c : c - CO;
      const int64_t o = (b * (is_dy ? CO : CI) + r) * kP + p;
      float* s1 = buf + (is_dy ? r : 2 * CO + r) * WCS + 4 * k4;
      float* s2 = s1 + (is_dy ? CO : CI) * WCS;
      if (valid) { cp_async16(s1, (is_dy ? dy1 : G) + o); cp_async16(s2, (is_dy ? dy2 : X) + o); }
      else { *reinterpret_cast<float4*>(s1) = zero4; *reinterpret_cast<float4*>(s2) = zero4; }
    }
    cp_async_commit();
  };
  int64_t e0 = static_cast<int64_t>(blockIdx.x) * WC;
  if (e0 < E) stage(e0, sm);
  for (int it = 0; e0 < E; e0 += stride, ++it) {
    float* buf = sm + (it & 1) * buf_floats;
    if (e0 + stride < E) {
      stage(e0 + stride, sm + ((it + 1) & 1) * buf_floats);
      asm volatile("cp.async.wait_group 1;\n" ::: "memory");
    } else {
      cp_async_wait_all();
    }
    __syncthreads();
    const float* d1s = buf;                      // [CO][WCS]
    const float* d2s = d1s + CO * WCS;           // [CO][WCS]
    const float* gs = d2s + CO * WCS;            // [CI][WCS]
    const float* xs = gs + CI * WCS;             // [CI][WCS]
    if (want_bias) {
      const float* row = d1s + brow * WCS;       // d2s follows d1s: rows CO.. are the residual branch
      for (int q = bk4; q < kV4; q += tpr) {
        const float4 v = *reinterpret_cast<const float4*>(row + 4 * q);
        bsum += (v.x + v.y) + (v.z + v.w);
      }
    }
    for (int k = ks * kslice; k < (ks + 1) * kslice; k += 4) {
      ulonglong2 d1[4], d2[4], g[4], x[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int co = (i < TCO) ? cot + i * NOT : cot;
        const int ci = (i < TCI) ? cit + i * NCT : cit;
        d1[i] = *reinterpret_cast<const ulonglong2*>(d1s + co * WCS + k);
        d2[i] = *reinterpret_cast<const ulonglong2*>(d2s + co * WCS + k);
        g[i] = *reinterpret_cast<const ulonglong2*>(gs + ci * WCS + k);
        x[i] = *reinterpret_cast<const ulonglong2*>(xs + ci * WCS + k);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          ffma2(q1[i][j], d1[i].x, g[j].x); ffma2(q1[i][j], d1[i].y, g[j].y);
          ffma2(q2[i][j], d2[i].x, x[j].x); ffma2(q2[i][j], d2[i].y, x[j].y);
        }
      }
    }
    __syncthreads();                             // this buffer is restaged by the next iteration
  }
  // block reduction over the k-slices: red[v][thread], v = (i*4 + j)*2 + branch; then one atomic per weight and block
  float* red = sm;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      red[((i * 4 + j) * 2 + 0) * kTrainThreads + threadIdx.x] = lo2(q1[i][j]) + hi2(q1[i][j]);
      red[((i * 4 + j) * 2 + 1) * kTrainThreads + threadIdx.x] = lo2(q2[i][j]) + hi2(q2[i][j]);
    }
  __syncthreads();
  for (int o = threadIdx.x; o < ntile * 32; o += kTrainThreads) {
    const int t = o % ntile, v = o / ntile;
    const int i = v >> 3, j = (v >> 1) & 3, branch = v & 1;
    if (i >= TCO || j >= TCI) continue;
    float acc = 0.f;
    for (int s2 = 0; s2 < nks; ++s2) acc += red[v * kTrainThreads + s2 * ntile + t];
    const int co = t / NCT + i * NOT, ci = t % NCT + j * NCT;
    atomicAdd((branch ? dW2 : dW1) + co * CI + ci, acc);
  }
  // bias gradients: the tpr threads of a row are neighbouring lanes
  for (int o = tpr >> 1; o > 0; o >>= 1) bsum += __shfl_down_sync(0xffffffffu, bsum, o, 32);
  if (bk4 == 0 && brow < 2 * CO) {
    if (brow < CO) { if (db1) atomicAdd(db1 + brow, bsum); }
    else if (db2) atomicAdd(db2 + (brow - CO), bsum);
  }
}

// ---- 1x1 convolutions as a register-tiled GEMM over positions ---------------------------------------------------------
// out1[m, e] = sum_k Wa(m,k) in1[k, e] (+ bias1[m]);  out2 likewise from in2 / Wb / bias2;  e = (b, p) flattened, tensors
// stored [b][C][204].  Forward: in = (G, X), W = (tcn conv, residual conv) [M = c_out][K = c_in], optional BatchNorm
// statistics (double) [4*M] += sum out1, sum out1^2, sum out2, sum out2^2.  Backward data: in = (dy1, dy2), W used
// transposed (dG[ci] = sum_co W1[co,ci] dy1[co]): w_is_km = 1, no bias, no statistics.
// A block stages the weights once as [k][m] and walks chunks of CE positions: the two input tiles live in shared memory as
// [k][CE + 4]; a warp owns a slab of 128 positions (lane = 4 consecutive positions) and up to two 4-row output tiles, so
// that one broadcast LDS.128 of weights and one LDS.128 of inputs feed 16 FMAs (the per-position kernels this replaces
// paid one shared-memory load per FMA, and 20 shuffles per output channel for the statistics).
constexpr int kGThreads = 256;
constexpr int kGWarps = kGThreads / 32;
template <bool STATS>
__global__ void __launch_bounds__(kGThreads) chan_gemm_kernel(const float* __restrict__ in1, const float* __restrict__ in2,
                                                              const float* __restrict__ Wa, const float* __restrict__ Wb,
                                                              int w_is_km, const float* __restrict__ bias1,
                                                              const float* __restrict__ bias2, int64_t B, int K, int M, int CE,
                                                              float* __restrict__ out1, float* __restrict__ out2, double* stats) {
  extern __shared__ __align__(128) float sm[];
  const int CES = CE + 4;
  float* wa = sm;                        // [K][M]
  float* wb = wa + K * M;
  float* ia = wb + K * M;                // [K][CES]
  float* ib = ia + K * CES;
  float* red = ib + K * CES;             // [4][M]  (STATS)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < K * M; i += kGThreads) {
    const int k = i / M, m = i % M;
    wa[i] = w_is_km ? Wa[k * M + m] : Wa[m * K + k];
    wb[i] = w_is_km ? Wb[k * M + m] : Wb[m * K + k];
  }
  if (STATS) for (int i = tid; i < 4 * M; i += kGThreads) red[i] = 0.f;
  const int TM = M < 4 ? M : 4;
  const int n_ct = M / TM;               // output-row tiles: 16, 8, 4 or 1
  const int nslab = CE / 128;            // 1 or 2
  // warp -> (slab, first tile, number of tiles): n_ct >= 8: slab 0, tiles warp, warp + 8; else one tile, slab = warp / n_ct
  const int slab = n_ct >= kGWarps ? 0 : warp / n_ct;
  const int ct0 = n_ct >= kGWarps ? warp : warp % n_ct;
  const int nct = n_ct >= kGWarps ? n_ct / kGWarps : (slab < nslab ? 1 : 0);
  const int64_t E = B * kP;
  const int v4_per_row = CE / 4;
  for (int64_t e0 = static_cast<int64_t>(blockIdx.x) * CE; e0 < E; e0 += static_cast<int64_t>(gridDim.x) * CE) {
    __syncthreads();
    // stage the input tiles with 16-byte loads: 4 consecutive positions never straddle a window (204 = 4 * 51, e0 % 4 == 0)
    for (int i0 = tid; i0 < K * v4_per_row; i0 += kGThreads * 4) {
      float4 va[4], vb[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = i0 + j * kGThreads;
        va[j] = make_float4(0.f, 0.f, 0.f, 0.f); vb[j] = va[j];
        if (i < K * v4_per_row) {
          const int k = i / v4_per_row, c4 = i - k * v4_per_row;
          const int64_t e = e0 + 4 * c4;
          if (e < E) {
            const int64_t b = e / kP;
            const int p = static_cast<int>(e - b * kP);
            va[j] = __ldg(reinterpret_cast<const float4*>(in1 + (b * K + k) * kP + p));
            vb[j] = __ldg(reinterpret_cast<const float4*>(in2 + (b * K + k) * kP + p));
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = i0 + j * kGThreads;
        if (i < K * v4_per_row) {
          const int k = i / v4_per_row, c4 = i - k * v4_per_row;
          *reinterpret_cast<float4*>(ia + k * CES + 4 * c4) = va[j];
          *reinterpret_cast<float4*>(ib + k * CES + 4 * c4) = vb[j];
        }
      }
    }
    __syncthreads();
    if (nct > 0) {
      const int el = slab * 128 + lane * 4;                 // this lane's 4 positions inside the chunk
      // accumulators packed over position pairs (FFMA2: one issue slot per two FMAs; the weight is the broadcast operand)
      unsigned long long p1[2][4][2], p2[2][4][2];
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 4; ++i) { p1[c][i][0] = p1[c][i][1] = 0ull; p2[c][i][0] = p2[c][i][1] = 0ull; }
      for (int k = 0; k < K; ++k) {
        const ulonglong2 xa = *reinterpret_cast<const ulonglong2*>(ia + k * CES + el);
        const ulonglong2 xb = *reinterpret_cast<const ulonglong2*>(ib + k * CES + el);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (c < nct) {
            const int m0 = (ct0 + c * kGWarps) * TM;
            float wav[4], wbv[4];
            if (TM == 4) {
              const float4 w1 = *reinterpret_cast<const float4*>(wa + k * M + m0);
              const float4 w2 = *reinterpret_cast<const float4*>(wb + k * M + m0);
              wav[0] = w1.x; wav[1] = w1.y; wav[2] = w1.z; wav[3] = w1.w;
              wbv[0] = w2.x; wbv[1] = w2.y; wbv[2] = w2.z; wbv[3] = w2.w;
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i) { wav[i] = i < TM ? wa[k * M + m0 + i] : 0.f; wbv[i] = i < TM ? wb[k * M + m0 + i] : 0.f; }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const unsigned long long w1d = dup2(wav[i]), w2d = dup2(wbv[i]);
              ffma2(p1[c][i][0], w1d, xa.x); ffma2(p1[c][i][1], w1d, xa.y);
              ffma2(p2[c][i][0], w2d, xb.x); ffma2(p2[c][i][1], w2d, xb.y);
            }
          }
        }
      }
      float a1[2][4][4], a2[2][4][4];
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          a1[c][i][0] = lo2(p1[c][i][0]); a1[c][i][1] = hi2(p1[c][i][0]); a1[c][i][2] = lo2(p1[c][i][1]); a1[c][i][3] = hi2(p1[c][i][1]);
          a2[c][i][0] = lo2(p2[c][i][0]); a2[c][i][1] = hi2(p2[c][i][0]); a2[c][i][2] = lo2(p2[c][i][1]); a2[c][i][3] = hi2(p2[c][i][1]);
        }
      const int64_t e = e0 + el;
      const bool valid = e < E;                             // E % 4 == 0: the 4 positions are valid together
      const int64_t b = valid ? e / kP : 0;
      const int p = valid ? static_cast<int>(e - b * kP) : 0;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (c < nct) {
          const int m0 = (ct0 + c * kGWarps) * TM;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (i < TM) {
              const int m = m0 + i;
              const float bb1 = bias1 ? bias1[m] : 0.f, bb2 = bias2 ? bias2[m] : 0.f;
              const float4 o1 = make_float4(a1[c][i][0] + bb1, a1[c][i][1] + bb1, a1[c][i][2] + bb1, a1[c][i][3] + bb1);
              const float4 o2 = make_float4(a2[c][i][0] + bb2, a2[c][i][1] + bb2, a2[c][i][2] + bb2, a2[c][i][3] + bb2);
              if (valid) {
                *reinterpret_cast<float4*>(out1 + (b * M + m) * kP + p) = o1;
                *reinterpret_cast<float4*>(out2 + (b * M + m) * kP + p) = o2;
              }
              if (STATS) {
                float s1 = 0.f, q1 = 0.f, s2 = 0.f, q2 = 0.f;
                if (valid) {
                  s1 = (o1.x + o1.y) + (o1.z + o1.w); q1 = fmaf(o1.x, o1.x, fmaf(o1.y, o1.y, fmaf(o1.z, o1.z, o1.w * o1.w)));
                  s2 = (o2.x + o2.y) + (o2.z + o2.w); q2 = fmaf(o2.x, o2.x, fmaf(o2.y, o2.y, fmaf(o2.z, o2.z, o2.w * o2.w)));
                }
                s1 = warp_sum(s1); q1 = warp_sum(q1); s2 = warp_sum(s2); q2 = warp_sum(q2);
                if (lane == 0) { atomicAdd(red + m, s1); atomicAdd(red + M + m, q1); atomicAdd(red + 2 * M + m, s2); atomicAdd(red + 3 * M + m, q2); }
              }
            }
          }
        }
      }
    }
  }
  if (STATS) {
    __syncthreads();
    for (int i = tid; i < 4 * M; i += kGThreads) atomicAdd(stats + i, static_cast<double>(red[i]));
  }
}

// ---- linear layers over the flattened features ------------------------------------------------------
// W(d, f) = W[d*sd + f*sf]:  btlnk / fc_* weight [D,F]: sd = F, sf = 1;  rev_btlnk weight [F,D]: sd = 1, sf = D
// out[b,d] += sum_{f in slice} A[b,f] W(d,f) (+ bias[d] from slice 0)   (wide-in: head forward, rev_btlnk input gradient)
// grid (ceil(B / 32), kLinSlices): a block owns 32 rows (4 per warp) and one slice of the features; it stages W chunks of
// [DMAX][kLinFC] in shared memory once for all its rows (one row per block re-read all 835 KB of W per row from L2) and
// writes its partial sums to its slice of `out` ([slices][B][D]); partial_sum_kernel adds the slices in a fixed order.  A lane takes 4 consecutive features: 16-byte loads
// of the activations (4 in flight per lane) and of the staged weights; F % 4 == 0 (host-checked).
constexpr int kLinRows = 4;            // rows per warp
constexpr int kLinFC = 512;            // features per staged chunk (544 = a third of a 1 632-feature slice was measured slower:
                                       // 4.25 lane strides per chunk leave most lanes idle in the fifth)
constexpr int kLinSlices = 8;
template <int DMAX>
__global__ void lin_reduce_f_kernel(const float* __restrict__ A, const float* __restrict__ W, int64_t sd, int64_t sf,
                                    const float* __restrict__ bias, int64_t B, int F, int D, float* __restrict__ out) {
  __shared__ __align__(16) float ws[DMAX][kLinFC + 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const int64_t b0 = (static_cast<int64_t>(blockIdx.x) * nwarp + warp) * kLinRows;
  const int fslice = ((F + kLinSlices - 1) / kLinSlices + 3) / 4 * 4;
  const int f_lo = blockIdx.y * fslice, f_hi = min(F, f_lo + fslice);
  float acc[kLinRows][DMAX];
#pragma unroll
  for (int r = 0; r < kLinRows; ++r)
#pragma unroll
    for (int d = 0; d < DMAX; ++d) acc[r][d] = 0.f;
  for (int f0 = f_lo; f0 < f_hi; f0 += kLinFC) {
    const int nf = min(kLinFC, f_hi - f0);
    __syncthreads();
    if (sf == 1 && sd % 4 == 0) {
      // rows of W are contiguous: 16-byte asynchronous copies, all in flight at once (as a load -> store loop the staging
      // exposed the L2 latency 8 times per chunk and was most of the kernel's time); nf % 4 == 0
      for (int i = threadIdx.x; i < DMAX * (kLinFC / 4); i += blockDim.x) {
        const int d = i / (kLinFC / 4), f = 4 * (i - d * (kLinFC / 4));
        if (d < D && f < nf) cp_async16(&ws[d][f], W + d * sd + f0 + f);
        else *reinterpret_cast<float4*>(&ws[d][f]) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      cp_async_commit();
      cp_async_wait_all();
    } else {
      for (int i = threadIdx.x; i < DMAX * kLinFC; i += blockDim.x) {
        const int d = i / kLinFC, f = i - d * kLinFC;
        ws[d][f] = (d < D && f < nf) ? W[d * sd + (f0 + f) * sf] : 0.f;
      }
    }
    __syncthreads();
    for (int f = 4 * lane; f < nf; f += 128) {
      float4 a[kLinRows];
#pragma unroll
      for (int r = 0; r < kLinRows; ++r)
        a[r] = (b0 + r < B) ? __ldg(reinterpret_cast<const float4*>(A + (b0 + r) * F + f0 + f)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int d = 0; d < DMAX; ++d) {
        const float4 w = *reinterpret_cast<const float4*>(&ws[d][f]);
#pragma unroll
        for (int r = 0; r < kLinRows; ++r)
          acc[r][d] = fmaf(a[r].w, w.w, fmaf(a[r].z, w.z, fmaf(a[r].y, w.y, fmaf(a[r].x, w.x, acc[r][d]))));
      }
    }
  }
#pragma unroll
  for (int r = 0; r < kLinRows; ++r)
#pragma unroll
    for (int d = 0; d < DMAX; ++d) {
      const float s = warp_sum(acc[r][d]);
      if (lane == 0 && d < D && b0 + r < B)      // one partial per feature slice: out is [kLinSlices][B][D] here
        out[(static_cast<int64_t>(blockIdx.y) * B + b0 + r) * D + d] = s + ((blockIdx.y == 0 && bias) ? bias[d] : 0.f);
    }
}
// out[b,f] = sum_d a[b,d] W(d,f) + bias[f]           (wide-out: rev_btlnk forward, head input gradient)
// grid (ceil(F/256), ceil(B/kExpRows)): a thread keeps the D weights of its feature f in registers for kExpRows rows b
constexpr int kExpRows = 128;
template <int DMAX>
__global__ void lin_expand_f_kernel(const float* __restrict__ a, const float* __restrict__ W, int64_t sd, int64_t sf,
                                    const float* __restrict__ bias, int64_t B, int F, int D, float* __restrict__ out) {
  static_assert(DMAX % 4 == 0, "rows of the staged latents are read as float4");
  __shared__ __align__(16) float as[kExpRows][DMAX];
  const int64_t b0 = static_cast<int64_t>(blockIdx.y) * kExpRows;
  for (int i = threadIdx.x; i < kExpRows * DMAX; i += blockDim.x) {
    const int r = i / DMAX, d = i % DMAX;
    as[r][d] = (d < D && b0 + r < B) ? a[(b0 + r) * D + d] : 0.f;
  }
  __syncthreads();
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  float w[DMAX];
#pragma unroll
  for (int d = 0; d < DMAX; ++d) w[d] = (d < D) ? W[d * sd + f * sf] : 0.f;
  const float bf = bias ? bias[f] : 0.f;
  for (int r = 0; r < kExpRows && b0 + r < B; ++r) {
    float s = bf;
#pragma unroll
    for (int d4 = 0; d4 < DMAX / 4; ++d4) {        // broadcast LDS.128: one shared-memory load per 4 FMAs
      const float4 v = *reinterpret_cast<const float4*>(&as[r][4 * d4]);
      s = fmaf(v.w, w[4 * d4 + 3], fmaf(v.z, w[4 * d4 + 2], fmaf(v.y, w[4 * d4 + 1], fmaf(v.x, w[4 * d4], s))));
    }
    out[(b0 + r) * F + f] = s;
  }
}
// dW(d,f) += sum_b a[b,d] A[b,f].  grid (ceil(F/128), NB): the 8 warps of a block share 128 features (a lane owns 4
// consecutive ones, 16-byte loads) and split the block's rows; their partial sums meet in shared memory and the block
// writes one partial per weight and row slice (summed in a fixed order by partial_sum_kernel).
constexpr int kWgF = 128;
template <int DMAX>
__global__ void __launch_bounds__(kTrainThreads) lin_wgrad_kernel(const float* __restrict__ a, const float* __restrict__ A,
                                                                  int64_t sd, int64_t sf, int64_t B, int F, int D, float* dW) {
  static_assert(DMAX % 4 == 0, "latent rows are read as float4");
  constexpr int kHalf = kTrainThreads / 64;              // the warps meet in two rounds: 33 KB of static shared memory
  __shared__ float red[kHalf][DMAX][kWgF + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = kTrainThreads / 32;
  const int f = blockIdx.x * kWgF + 4 * lane;
  float acc[DMAX][4];
#pragma unroll
  for (int d = 0; d < DMAX; ++d) acc[d][0] = acc[d][1] = acc[d][2] = acc[d][3] = 0.f;
  if (f < F) {
    const bool a_vec = (D == DMAX);                      // rows of `a` are 16-byte aligned only for the full width
    const int64_t bstep = static_cast<int64_t>(gridDim.y) * nwarp;
    for (int64_t bb = static_cast<int64_t>(blockIdx.y) * nwarp + warp; bb < B; bb += 4 * bstep) {
     // the wide rows of 4 steps are requested together (the latent rows that follow are L1 hits)
     float4 vq[4];
#pragma unroll
     for (int u = 0; u < 4; ++u)
       vq[u] = (bb + u * bstep < B) ? __ldg(reinterpret_cast<const float4*>(A + (bb + u * bstep) * F + f)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
     for (int u = 0; u < 4; ++u) {
      const int64_t b = bb + u * bstep;
      if (b >= B) break;
      const float4 v = vq[u];
      float ar[DMAX];
      if (a_vec) {
#pragma unroll
        for (int d4 = 0; d4 < DMAX / 4; ++d4) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(a + b * D) + d4);
          ar[4 * d4] = t.x; ar[4 * d4 + 1] = t.y; ar[4 * d4 + 2] = t.z; ar[4 * d4 + 3] = t.w;
        }
      } else {
#pragma unroll
        for (int d = 0; d < DMAX; ++d) ar[d] = (d < D) ? __ldg(a + b * D + d) : 0.f;
      }
#pragma unroll
      for (int d = 0; d < DMAX; ++d) {
        acc[d][0] = fmaf(ar[d], v.x, acc[d][0]); acc[d][1] = fmaf(ar[d], v.y, acc[d][1]);
        acc[d][2] = fmaf(ar[d], v.z, acc[d][2]); acc[d][3] = fmaf(ar[d], v.w, acc[d][3]);
      }
     }
    }
  }
  if (warp >= kHalf) {
#pragma unroll
    for (int d = 0; d < DMAX; ++d)
#pragma unroll
      for (int q = 0; q < 4; ++q) red[warp - kHalf][d][4 * lane + q] = acc[d][q];
  }
  __syncthreads();
  if (warp < kHalf) {
#pragma unroll
    for (int d = 0; d < DMAX; ++d)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[d][q] += red[warp][d][4 * lane + q];
  }
  __syncthreads();
  if (warp < kHalf) {
#pragma unroll
    for (int d = 0; d < DMAX; ++d)
#pragma unroll
      for (int q = 0; q < 4; ++q) red[warp][d][4 * lane + q] = acc[d][q];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D * kWgF; i += kTrainThreads) {
    const int d = i / kWgF, fl = i - d * kWgF;
    const int fo = blockIdx.x * kWgF + fl;
    if (fo >= F) continue;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kHalf; ++w) s += red[w][d][fl];
    dW[static_cast<int64_t>(blockIdx.y) * D * F + d * sd + fo * sf] = s;       // partial of this row slice: [gridDim.y][D*F]
  }
}
// column sums: partial[y][j] = sum over the slice's rows of a[b, j]   (bias gradients); grid (ceil(N/32), row slices): a block owns 32 columns and
// every gridDim.y-th group of 8 rows
__global__ void col_sum_kernel(const float* __restrict__ a, int64_t B, int N, float* out) {
  const int j = blockIdx.x * 32 + (threadIdx.x & 31);
  const int nrow = blockDim.x >> 5;
  const int row0 = blockIdx.y * nrow + (threadIdx.x >> 5);
  float s = 0.f;
  if (j < N) for (int64_t b = row0; b < B; b += static_cast<int64_t>(nrow) * gridDim.y) s += a[b * N + j];
  __shared__ float sh[8][33];
  sh[threadIdx.x >> 5][threadIdx.x & 31] = s;
  __syncthreads();
  if ((threadIdx.x >> 5) == 0 && j < N) {
    float t = 0.f;
    for (int r = 0; r < nrow; ++r) t += sh[r][threadIdx.x & 31];
    out[static_cast<int64_t>(blockIdx.y) * N + j] = t;                          // partial of this row slice: [gridDim.y][N]
  }
}

// ---- Adam over ONE flat parameter buffer (dist.FlatGradBucket gradients, optim.FlatAdam) -------------------------------------
// torch.optim.Adam(params, lr, betas, eps) with its defaults (no weight decay, no amsgrad) -- models/hyperbolic_encoder.py:199 --
// in the arithmetic of torch's fused / capturable kernel: m = lerp(m, g, 1 - b1); v = b2 v + (1 - b2) g^2;
// p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps).  torch's multi-tensor kernel takes two 22 us launches for the
// encoder's 62 small tensors (240 k parameters); one pass over the flat buffers is a few microseconds.
// tick: t += 1, sc[0] = 1 / (1 - b1^t), sc[1] = 1 / sqrt(1 - b2^t)  (one thread; ordered before the update by the stream)
__global__ void adam_tick_kernel(int64_t* step, float* sc, float beta1, float beta2) {
  const int64_t t = *step + 1;
  *step = t;
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), static_cast<double>(t));
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), static_cast<double>(t));
  sc[0] = static_cast<float>(1.0 / bc1);
  sc[1] = static_cast<float>(1.0 / sqrt(bc2));
}
__global__ void adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                 int64_t n4, const float* __restrict__ lr, const float* __restrict__ sc, float beta1, float beta2,
                                 float eps) {
  const float step_size = lr[0] * sc[0], inv_bc2s = sc[1];
  const float w1 = 1.f - beta1, w2 = 1.f - beta2;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float* pa = &pp.x; float* ma = &mm.x; float* va = &vv.x; const float* ga = &gg.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      ma[k] = fmaf(w1, ga[k] - ma[k], ma[k]);
      va[k] = fmaf(w2 * ga[k], ga[k], beta2 * va[k]);
      const float denom = fmaf(sqrtf(va[k]), inv_bc2s, eps);
      pa[k] = pa[k] - step_size * (ma[k] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
  }
}

}  // namespace coskad
