// train_tc.cuh -- the GEMM-shaped parts of a training step on the 5th-generation tensor cores (tcgen05, TMEM).
//
// A train-mode ST_GCNN layer (models/graph_layers/stsgcn.py:94-116 under autograd) holds three GEMMs over the
// E = B*204 positions of a batch, all with tiny channel extents (2..64):
//   forward        y1[e,co] = sum_ci W1[co,ci] G[e,ci] + b1[co],  y2 likewise from X / W2     (M = e, N = co, K = ci)
//   backward data  dG[e,ci] = sum_co W1[co,ci] dy1[e,co],         dXres likewise              (M = e, N = ci, K = co)
//   weight grad    dW1[co,ci] = sum_e dy1[e,co] G[e,ci], db1[co] = sum_e dy1[e,co]            (M = co, N = ci, K = e)
// Single-pass TF32 breaks the 1e-4 tolerance, so every product is the 3xTF32 split of tc.cuh:
// a = a_hi + a_lo, D += A_hi B_hi + A_lo B_hi + A_hi B_lo.  With the tensor pipe doing the MACs these kernels are bound by
// HBM (activations are [B, C, 204] float32 between kernels) instead of by the FP32 pipe.
//
//  * forward / backward data: M-tile = 128 consecutive positions, one thread per position (= TMEM lane).  The thread loads
//    its position's channel values (coalesced across the warp), splits them and writes them as the A operand straight into
//    TMEM (tcgen05.st); the weights are canonical K-major shared-memory images built once per CTA; D comes back with
//    tcgen05.ld and is stored coalesced.  The backward-data kernel computes dy1 / dy2 = BatchNorm-train + PReLU backward of
//    (dout, y1, y2) on the fly (the former element-wise "apply" kernel is gone) and writes them once for the weight gradient.
//  * weight gradient: both operands are K-major in global memory already (positions are contiguous per channel row), so
//    16-byte loads go through registers (split) into padded canonical images (LBO = 144 B: conflict-free STS.128); M = 64
//    rows of dy^T, N = c_in rows of G^T plus one row of ones (its column of D is the bias gradient); the accumulators stay
//    in TMEM for all tiles of a CTA (split-K over the grid) and leave as one partial block per CTA.
//  * every cross-CTA sum (BatchNorm statistics, dW, db) goes through per-CTA partials and a fixed-order second stage
//    (partial_sum_kernel): no floating-point atomics, two runs are bit-identical.
#pragma once
#include "common.cuh"
#include "reduce.cuh"
#include "tc.cuh"

namespace coskad {

constexpr int kTcT = 128;                       // threads per CTA = positions per M-tile = TMEM lanes

__host__ __device__ constexpr int tmem_alloc_cols(int need) {
  return need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;
}
// canonical K-major, no swizzle: element (n, k) of an [N][K] operand; LBO = (N/8)*128 B, SBO = 128 B (fold.cuh, tc_test.cuh)
__device__ __forceinline__ int kmaj_idx(int n, int k, int N) { return ((k >> 2) * (N >> 3) + (n >> 3)) * 32 + (n & 7) * 4 + (k & 3); }

// hi / lo K-major images of a (zero-padded) weight matrix pair, built with all global loads of a thread in flight together: as a
// load -> split -> store loop the prologue exposed one L2 round trip per element (8-10 % of the kernels, ncu source page)
template <int NROWS, int KDIM, class F>
__device__ __forceinline__ void build_weight_images(float (*wimg)[NROWS * KDIM], int tid, F elem) {
  constexpr int kBatch = 8;
  for (int i0 = tid; i0 < NROWS * KDIM; i0 += kTcT * kBatch) {
    float a[kBatch], b[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int i = i0 + u * kTcT;
      const int ii = i < NROWS * KDIM ? i : NROWS * KDIM - 1;
      elem(ii / KDIM, ii % KDIM, a[u], b[u]);
    }
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int i = i0 + u * kTcT;
      if (i < NROWS * KDIM) {
        const int n = i / KDIM, k = i - n * KDIM;
        uint32_t h, l;
        tc::split_tf32(a[u], h, l);
        wimg[0][kmaj_idx(n, k, NROWS)] = __uint_as_float(h); wimg[1][kmaj_idx(n, k, NROWS)] = __uint_as_float(l);
        tc::split_tf32(b[u], h, l);
        wimg[2][kmaj_idx(n, k, NROWS)] = __uint_as_float(h); wimg[3][kmaj_idx(n, k, NROWS)] = __uint_as_float(l);
      }
    }
  }
}

// ---- forward: y1 = conv1x1(G; W1, b1), y2 = conv1x1(X; W2, b2) + per-CTA BatchNorm statistics ---------------------------
// part [gridDim.x][4*CO] = sum y1, sum y1^2, sum y2, sum y2^2 per channel of the positions the CTA owned.
// CIP = c_in padded to a multiple of 8 (K step of kind::tf32), COP = c_out padded to a multiple of 16 (N of an M = 128 MMA).
template <int CIP, int CO, int COP>
__global__ void __launch_bounds__(kTcT) tc_mix_fwd_kernel(const float* __restrict__ G, const float* __restrict__ X,
                                                         const float* __restrict__ W1, const float* __restrict__ b1,
                                                         const float* __restrict__ W2, const float* __restrict__ b2, int64_t E,
                                                         int CI, float* __restrict__ y1, float* __restrict__ y2,
                                                         float* __restrict__ part) {
  constexpr int kColD = 4 * CIP;                                  // A: G_hi, G_lo, X_hi, X_lo; D: y1 | y2
  constexpr int kAlloc = tmem_alloc_cols(4 * CIP + 2 * COP);
  constexpr int NQ = 2 * COP / 16;                                // 16-column chunks of [y1 | y2]
  __shared__ __align__(128) float wimg[4][COP * CIP];             // W1 hi, W1 lo, W2 hi, W2 lo
  __shared__ float bias_s[2 * COP];
  __shared__ float tr[4][32][17];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  build_weight_images<COP, CIP>(wimg, tid, [&](int n, int k, float& a, float& b) {
    const bool in = n < CO && k < CI;
    a = in ? __ldg(W1 + n * CI + k) : 0.f;
    b = in ? __ldg(W2 + n * CI + k) : 0.f;
  });
  for (int i = tid; i < 2 * COP; i += kTcT) {
    const int br = i / COP, co = i % COP;
    const float* bp = br ? b2 : b1;
    bias_s[i] = (co < CO && bp) ? bp[co] : 0.f;
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, kAlloc);
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  tc::fence_proxy_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = tmem_base_s;
  const uint32_t lane_base = tbase + (static_cast<uint32_t>(warp * 32) << 16);
  const int64_t ntiles = (E + kTcT - 1) / kTcT;
  float sacc[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) sacc[q] = 0.f;
  float g[CIP], x[CIP];
  auto load_tile = [&](int64_t t) {
    // out-of-range threads / tiles read a valid address (element 0 of their channel row) and are masked afterwards, so that
    // the whole batch is unconditional loads in flight together
    const int64_t e = t * kTcT + tid;
    const bool ok = t < ntiles && e < E;
    const int64_t b = ok ? e / kP : 0;
    const int64_t base = (b * CI) * kP + (ok ? e - b * kP : 0);
#pragma unroll
    for (int c = 0; c < CIP; ++c) {
      if (c < CI) { g[c] = tc::ldg_stay(G + base + static_cast<int64_t>(c) * kP); x[c] = tc::ldg_stay(X + base + static_cast<int64_t>(c) * kP); }
      else { g[c] = 0.f; x[c] = 0.f; }
    }
    if (!ok) {
#pragma unroll
      for (int c = 0; c < CIP; ++c) { g[c] = 0.f; x[c] = 0.f; }
    }
  };
  uint32_t phase = 0;
  load_tile(blockIdx.x);
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    // registers -> split -> A operand in TMEM
#pragma unroll
    for (int c0 = 0; c0 < CIP; c0 += 8) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) tc::split_tf32(g[c0 + j], hi[j], lo[j]);
      tc::tmem_st8(lane_base + c0, hi);
      tc::tmem_st8(lane_base + CIP + c0, lo);
#pragma unroll
      for (int j = 0; j < 8; ++j) tc::split_tf32(x[c0 + j], hi[j], lo[j]);
      tc::tmem_st8(lane_base + 2 * CIP + c0, hi);
      tc::tmem_st8(lane_base + 3 * CIP + c0, lo);
    }
    tc::wait_st();
    tc::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_sync();
      const uint32_t idesc = tc::make_idesc_tf32(128, COP);
      const uint32_t lbo = (COP / 8) * 128, sbo = 128;
#pragma unroll
      for (int kb = 0; kb < CIP / 8; ++kb) {
        const uint32_t offs = kb * 2 * lbo;
        const uint64_t d1h = tc::make_smem_desc(tc::smem_u32(wimg[0]) + offs, lbo, sbo);
        const uint64_t d1l = tc::make_smem_desc(tc::smem_u32(wimg[1]) + offs, lbo, sbo);
        const uint64_t d2h = tc::make_smem_desc(tc::smem_u32(wimg[2]) + offs, lbo, sbo);
        const uint64_t d2l = tc::make_smem_desc(tc::smem_u32(wimg[3]) + offs, lbo, sbo);
        tc::mma_tf32_ts(tbase + kColD, tbase + kb * 8, d1h, idesc, kb > 0 ? 1u : 0u);
        tc::mma_tf32_ts(tbase + kColD, tbase + CIP + kb * 8, d1h, idesc, 1u);
        tc::mma_tf32_ts(tbase + kColD, tbase + kb * 8, d1l, idesc, 1u);
        tc::mma_tf32_ts(tbase + kColD + COP, tbase + 2 * CIP + kb * 8, d2h, idesc, kb > 0 ? 1u : 0u);
        tc::mma_tf32_ts(tbase + kColD + COP, tbase + 3 * CIP + kb * 8, d2h, idesc, 1u);
        tc::mma_tf32_ts(tbase + kColD + COP, tbase + 2 * CIP + kb * 8, d2l, idesc, 1u);
      }
      tc::mma_commit(&bar);
    }
    const int64_t e = t * kTcT + tid;
    const bool ok = e < E;
    const int64_t b = ok ? e / kP : 0;
    const int64_t obase = (b * CO) * kP + (ok ? e - b * kP : 0);
    load_tile(t + gridDim.x);                       // the next tile's loads fly while the tensor pipe and the epilogue run
    tc::mbar_wait(&bar, phase);
    phase ^= 1;
    tc::fence_after_sync();
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      uint32_t v[16];
      tc::tmem_ld16(lane_base + kColD + q * 16, v);
      tc::wait_ld();
      float* out = (q * 16 < COP) ? y1 : y2;
      const int co0 = (q * 16) % COP;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float val = __uint_as_float(v[j]) + bias_s[q * 16 + j];
        if (ok && co0 + j < CO) out[obase + static_cast<int64_t>(co0 + j) * kP] = val;
        tr[warp][lane][j] = ok ? val : 0.f;
      }
      __syncwarp();
      float s = 0.f;
#pragma unroll 8
      for (int r = 0; r < 32; ++r) { const float a = tr[warp][r][lane & 15]; s += (lane < 16) ? a : a * a; }
      sacc[q] += s;
      __syncwarp();
    }
  }
  {
    // the 4 warps' sums meet in shared memory in a fixed order: one partial [4*CO] per CTA
    float* wsum = &tr[0][0][0];                      // 4 * 32 * 17 floats >= 4 warps x 4*CO
    static_assert(4 * 32 * 17 >= 4 * 4 * CO, "stats scratch");
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int c = q * 16 + (lane & 15), branch = c / COP, co = c % COP;
      if (co < CO) wsum[warp * (4 * CO) + (2 * branch + (lane >= 16 ? 1 : 0)) * CO + co] = sacc[q];
    }
    __syncthreads();
    float* dst = part + static_cast<int64_t>(blockIdx.x) * (4 * CO);
    for (int i = tid; i < 4 * CO; i += kTcT) dst[i] = ((wsum[i] + wsum[4 * CO + i]) + wsum[8 * CO + i]) + wsum[12 * CO + i];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tbase, kAlloc);
}

// ---- backward data with the BatchNorm-train + PReLU backward fused in -----------------------------------------------------
// dy1 = g1*is1*(ds - mean(ds) - yhat1*mean(ds*yhat1)), ds = dout * PReLU'(BN1(y1) + BN2(y2)); dy2 likewise
// (train_bn_prelu_bwd_apply_kernel's expressions, bit for bit); dG = W1^T dy1, dXres = W2^T dy2.
// NP = max(16, c_in) output columns per branch; the K = COK = max(16, c_out) channels go through TMEM in chunks of KC = 16
// (few TMEM columns per CTA: the resident CTAs, not a deeper pipeline inside one, hide the HBM latency).
constexpr int kBwdKC = 16;
// ASYNC: the (dout, y1, y2) values of a 16-channel pass are staged through a dynamic shared-memory buffer [3][16][128] with
// 16-byte cp.async copies (12 per thread instead of 48 scalar loads; 4 consecutive positions never straddle a window:
// 204 = 4 * 51) that are issued one pass AHEAD -- right after every thread has copied the current pass out of the buffer --
// so the HBM latency of pass i + 1 overlaps the BatchNorm / PReLU arithmetic, the dy stores and the TMEM staging of pass i
// without holding the values in registers (a register prefetch cost a resident CTA per SM and was slower).
constexpr int kBwdStageFloats = 3 * kBwdKC * kTcT;                // 6144 floats = 24 KB
template <int CO, int COK, int NP, bool ASYNC>
__global__ void __launch_bounds__(kTcT) tc_mix_bwd_data_kernel(
    const float* __restrict__ dout, const float* __restrict__ y1, const float* __restrict__ y2, const float* __restrict__ mi,
    const float* __restrict__ g1, const float* __restrict__ be1, const float* __restrict__ g2, const float* __restrict__ be2,
    const float* __restrict__ slope, const double* __restrict__ red, const float* __restrict__ W1, const float* __restrict__ W2,
    int64_t E, int CI, float* __restrict__ dy1, float* __restrict__ dy2, float* __restrict__ dG, float* __restrict__ dXres) {
  constexpr int KC = kBwdKC;                                      // 16 channels per pass: 64 + 2 NP columns -> 4 CTAs per SM
  constexpr int NCH = COK / KC;
  constexpr int kColD = 4 * KC;                                   // A: dy1_hi, dy1_lo, dy2_hi, dy2_lo; D: dG | dXres
  constexpr int kAlloc = tmem_alloc_cols(4 * KC + 2 * NP);
  __shared__ __align__(128) float wimg[4][NP * COK];              // (W1^T) hi, lo, (W2^T) hi, lo: [N = ci][K = co]
  __shared__ __align__(16) float cst[COK][16];                    // 13 per-channel constants, one 64-byte row per channel: 4 LDS.128
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  extern __shared__ __align__(16) float bwd_stage[];              // ASYNC: [3][16][128] (dout | y1 | y2 of one pass)
  const int tid = threadIdx.x, warp = tid >> 5;
  const int64_t ntiles = (E + kTcT - 1) / kTcT;
  // ASYNC: thread tid copies the position quad (tid % 32) of the rows (array, channel) = warp + 4 k, k = 0..11
  auto issue_async = [&](int64_t t, int ch) {
    if (t >= ntiles) return;
    int64_t e = t * kTcT + 4 * (tid & 31);
    if (e >= E) e = 0;                                            // ragged last tile: any valid quad, masked by `ok` later
    const int64_t b = e / kP;
    const int64_t base = (b * CO) * kP + (e - b * kP);
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      const int row = warp + 4 * k, arr = row >> 4;
      int co = ch * kBwdKC + (row & 15);
      if (co >= CO) co = CO - 1;                                  // padded channels (c_out < 16): masked by `in` later
      const float* src = (arr == 0 ? dout : arr == 1 ? y1 : y2) + base + static_cast<int64_t>(co) * kP;
      cp_async16(bwd_stage + row * kTcT + 4 * (tid & 31), src);
    }
  };
  if (ASYNC) { issue_async(blockIdx.x, 0); cp_async_commit(); }
  build_weight_images<NP, COK>(wimg, tid, [&](int n, int k, float& a_, float& b_) {       // n = ci, k = co
    const bool in = n < CI && k < CO;
    a_ = in ? __ldg(W1 + k * CI + n) : 0.f;
    b_ = in ? __ldg(W2 + k * CI + n) : 0.f;
  });
  const double Nd = static_cast<double>(E);
  for (int co = tid; co < COK; co += kTcT) {
    if (co >= CO) {
#pragma unroll
      for (int q = 0; q < 16; ++q) cst[co][q] = 0.f;
      continue;
    }
    const float i1 = mi[CO + co], i2 = mi[3 * CO + co], ga1 = g1[co], ga2 = g2[co];
    cst[co][0] = mi[co]; cst[co][1] = i1; cst[co][2] = mi[2 * CO + co]; cst[co][3] = i2;
    cst[co][4] = ga1; cst[co][5] = be1[co]; cst[co][6] = ga2; cst[co][7] = be2[co];
    cst[co][8] = static_cast<float>(red[co] / Nd);
    cst[co][9] = static_cast<float>(red[CO + co] / Nd);
    cst[co][10] = static_cast<float>(red[2 * CO + co] / Nd);
    cst[co][11] = ga1 * i1; cst[co][12] = ga2 * i2;
    cst[co][13] = cst[co][14] = cst[co][15] = 0.f;
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, kAlloc);
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  tc::fence_proxy_async_smem();
  tc::fence_before_sync();
  if (ASYNC) cp_async_wait_all();                                 // the first pass has landed when the barrier opens
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = tmem_base_s;
  const uint32_t lane_base = tbase + (static_cast<uint32_t>(warp * 32) << 16);
  const float a = slope[0];
  uint32_t phase = 0;
  static_assert(KC == 16, "one 16-channel group per pass");
  // the (dout, y1, y2) values of a 16-channel pass: 48 unconditional loads in flight together (an out-of-range thread reads
  // row 0 and is masked later).  Loading one pass AHEAD was measured slower: +70 registers cost a resident CTA per SM, and
  // it is the resident CTAs (4 per SM) that hide the HBM latency here
  float nd[16], nu[16], nv[16];
  auto issue = [&](int64_t t, int ch) {
    const int64_t e = t * kTcT + tid;
    const bool ok = t < ntiles && e < E;
    const int64_t b = ok ? e / kP : 0;
    const int64_t ybase = (b * CO) * kP + (ok ? e - b * kP : 0);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (ch * KC + j < CO) {
        const int64_t o = ybase + static_cast<int64_t>(ch * KC + j) * kP;
        nd[j] = tc::ldg_stay(dout + o);
        nu[j] = tc::ldg_stay(y1 + o);
        nv[j] = tc::ldg_stay(y2 + o);
      } else { nd[j] = 0.f; nu[j] = 0.f; nv[j] = 0.f; }
    }
  };
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t e = t * kTcT + tid;
    const bool ok = e < E;
    const int64_t b = ok ? e / kP : 0;
    const int p = ok ? static_cast<int>(e - b * kP) : 0;
    const int64_t ybase = (b * CO) * kP + p;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      float dv[16], uv[16], vv[16];
      if (ASYNC) {
        // this pass is in the staging buffer (waited for before the previous barrier): copy it out, and as soon as every
        // thread has done so, put the next pass in flight
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          dv[j] = bwd_stage[j * kTcT + tid]; uv[j] = bwd_stage[(16 + j) * kTcT + tid]; vv[j] = bwd_stage[(32 + j) * kTcT + tid];
        }
        __syncthreads();
        if (ch + 1 < NCH) issue_async(t, ch + 1); else issue_async(t + gridDim.x, 0);
        cp_async_commit();
      } else {
        issue(t, ch);
#pragma unroll
        for (int j = 0; j < 16; ++j) { dv[j] = nd[j]; uv[j] = nu[j]; vv[j] = nv[j]; }
      }
      uint32_t h1[16], l1[16], h2[16], l2[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int co = ch * KC + j;
        float hh1, hh2;
        const float4 k0 = *reinterpret_cast<const float4*>(&cst[co][0]), k1 = *reinterpret_cast<const float4*>(&cst[co][4]);
        const float4 k2 = *reinterpret_cast<const float4*>(&cst[co][8]), k3 = *reinterpret_cast<const float4*>(&cst[co][12]);
        const float pre = bn_pre(uv[j], vv[j], k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w, hh1, hh2);
        const float ds = pre > 0.f ? dv[j] : a * dv[j];
        const bool in = ok && co < CO;
        const float r1 = in ? k2.w * (ds - k2.x - hh1 * k2.y) : 0.f;
        const float r2 = in ? k3.x * (ds - k2.x - hh2 * k2.z) : 0.f;
        if (in) {
          const int64_t o = ybase + static_cast<int64_t>(co) * kP;
          dy1[o] = r1;
          dy2[o] = r2;
        }
        tc::split_tf32(r1, h1[j], l1[j]);
        tc::split_tf32(r2, h2[j], l2[j]);
      }
      if (ch > 0) {                                               // the previous pass's MMAs still read the A columns
        tc::mbar_wait(&bar, phase);
        phase ^= 1;
        tc::fence_after_sync();
      }
      tc::tmem_st16(lane_base, h1);
      tc::tmem_st16(lane_base + KC, l1);
      tc::tmem_st16(lane_base + 2 * KC, h2);
      tc::tmem_st16(lane_base + 3 * KC, l2);
      tc::wait_st();
      tc::fence_before_sync();
      if (ASYNC) cp_async_wait_all();                             // the next pass (in flight since the top of this one)
      __syncthreads();
      if (tid == 0) {
        tc::fence_after_sync();
        const uint32_t idesc = tc::make_idesc_tf32(128, NP);
        const uint32_t lbo = (NP / 8) * 128, sbo = 128;
#pragma unroll
        for (int kb = 0; kb < KC / 8; ++kb) {
          const uint32_t offs = (ch * (KC / 8) + kb) * 2 * lbo;
          const uint64_t d1h = tc::make_smem_desc(tc::smem_u32(wimg[0]) + offs, lbo, sbo);
          const uint64_t d1l = tc::make_smem_desc(tc::smem_u32(wimg[1]) + offs, lbo, sbo);
          const uint64_t d2h = tc::make_smem_desc(tc::smem_u32(wimg[2]) + offs, lbo, sbo);
          const uint64_t d2l = tc::make_smem_desc(tc::smem_u32(wimg[3]) + offs, lbo, sbo);
          const uint32_t acc = (ch > 0 || kb > 0) ? 1u : 0u;
          tc::mma_tf32_ts(tbase + kColD, tbase + kb * 8, d1h, idesc, acc);
          tc::mma_tf32_ts(tbase + kColD, tbase + KC + kb * 8, d1h, idesc, 1u);
          tc::mma_tf32_ts(tbase + kColD, tbase + kb * 8, d1l, idesc, 1u);
          tc::mma_tf32_ts(tbase + kColD + NP, tbase + 2 * KC + kb * 8, d2h, idesc, acc);
          tc::mma_tf32_ts(tbase + kColD + NP, tbase + 3 * KC + kb * 8, d2h, idesc, 1u);
          tc::mma_tf32_ts(tbase + kColD + NP, tbase + 2 * KC + kb * 8, d2l, idesc, 1u);
        }
        tc::mma_commit(&bar);
      }
    }
    tc::mbar_wait(&bar, phase);
    phase ^= 1;
    tc::fence_after_sync();
    const int64_t xbase = (b * CI) * kP + p;
#pragma unroll
    for (int q = 0; q < 2 * NP / 16; ++q) {
      uint32_t v[16];
      tc::tmem_ld16(lane_base + kColD + q * 16, v);
      tc::wait_ld();
      float* out = (q * 16 < NP) ? dG : dXres;
      const int ci0 = (q * 16) % NP;
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (ok && ci0 + j < CI) out[xbase + static_cast<int64_t>(ci0 + j) * kP] = __uint_as_float(v[j]);
    }
    // the next tile's tcgen05.st must not overtake this tile's tcgen05.ld of other warps' MMAs: ordered by the barrier of the
    // next chunk (every thread finishes its epilogue before it arrives there)
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tbase, kAlloc);
}

// ---- weight gradient: dW1 = dy1^T G, db1 = dy1^T 1, dW2 = dy2^T X, db2 = dy2^T 1, split-K over the grid ----------------------
// part [gridDim.x][2][CO][CI + 1] (column CI = bias gradient).  KT positions per tile (32 for the 64-channel layer, 64
// otherwise: the per-tile barrier + MMA round trip is amortised over more bytes where shared memory allows 3 CTAs per SM);
// images are padded canonical K-major: element (r, k) at (r/8)*sbo + (k/4)*kWgLbo + (r%8)*4 + (k%4) floats.  An A image holds
// only the c_out rows that exist; the M = 64 MMA reads on into the images that follow (finite data), which only fills
// accumulator rows >= c_out that nobody reads.
constexpr int kWgLbo = 36;                      // floats (144 B): the 8 lanes of an STS.128 wavefront hit 32 distinct banks
__host__ __device__ constexpr int wg_n(int ci8) { return ci8 + 8; }
__host__ __device__ constexpr int wg_sbo(int kt) { return (kt / 4) * kWgLbo; }            // floats per 8-row group
__host__ __device__ constexpr int wg_agroups(int co) { return co < 8 ? 1 : co / 8; }
__host__ __device__ constexpr int wg_smem_floats(int co, int ci8, int kt) { return 4 * (wg_agroups(co) + wg_n(ci8) / 8) * wg_sbo(kt); }

template <int CO, int CI8, int KT>
__global__ void __launch_bounds__(kTcT) tc_mix_bwd_weight_kernel(const float* __restrict__ dy1, const float* __restrict__ dy2,
                                                                const float* __restrict__ G, const float* __restrict__ X,
                                                                int64_t E, int CI, float* __restrict__ part) {
  constexpr int N = wg_n(CI8);
  constexpr int kCh = KT / 4;                                     // 16-byte chunks per row and tile
  constexpr int kSbo = wg_sbo(KT);
  constexpr int kAimg = wg_agroups(CO) * kSbo;
  constexpr int kBimg = (N / 8) * kSbo;
  static_assert(wg_agroups(CO) + 4 * (N / 8) >= 8, "the M = 64 read of the last A image must stay inside the allocation");
  constexpr int kAlloc = tmem_alloc_cols(2 * N);
  constexpr int kItemsMax = ((2 * CO + 2 * CI8) * kCh + kTcT - 1) / kTcT;   // float4 per thread and tile
  extern __shared__ __align__(128) float wsm[];
  float* Aimg = wsm;                                              // dy1 hi, dy1 lo, dy2 hi, dy2 lo
  float* Bimg = wsm + 4 * kAimg;                                  // G hi, G lo, X hi, X lo (+ the row of ones at row CI8 of the hi images)
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 4 * (kAimg + kBimg); i += kTcT) wsm[i] = 0.f;
  __syncthreads();
  if (tid < KT) {                                                 // ones: row CI8 (first row of the last group), every k
    const int o = (CI8 / 8) * kSbo + (tid >> 2) * kWgLbo + (tid & 3);
    Bimg[0 * kBimg + o] = 1.f;
    Bimg[2 * kBimg + o] = 1.f;
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, kAlloc);
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  tc::fence_proxy_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tbase = tmem_base_s;
  const int rows = 2 * CO + 2 * CI;
  const int items = rows * kCh;
  const int64_t ntiles = (E + KT - 1) / KT;
  uint32_t phase = 0;
  bool first = true;
  // the 16-byte loads of a tile: unconditional, all in flight together while the previous tile's MMAs run; an item past the
  // end of the batch (E % 4 == 0: entirely in or out) or past the row list reads a valid address and is zeroed when consumed
  float4 vn[kItemsMax];
  // item u of thread tid is the 16-byte chunk kc = tid % kCh of row r = tid / kCh + u * kRS.  When the class boundaries
  // (c_out, 2 c_out, 2 c_out + c_in) are multiples of kRS -- every layer but the 2-channel input layer -- the class of an item
  // (dy1 | dy2 | G | X), its channel and its image offset depend on u alone: resolved at compile time, and the tile position
  // (window, offset) is computed once per tile instead of once per item.  (As generic per-item integer arithmetic this loop was
  // 64 % of the kernel's time: 49 M of its 54 M warp instructions, ncu source page.)
  constexpr int kRS = kTcT / kCh;                                  // rows between a thread's consecutive items
  static_assert(kTcT % kCh == 0 && kRS % 8 == 0, "row step of the item mapping");
  constexpr bool kAligned = (CO % kRS == 0) && (CI8 % kRS == 0);
  const bool fast = kAligned && CI == CI8;
  const int kcT = tid % kCh, rbT = tid / kCh;
  const int oT = (rbT >> 3) * kSbo + kcT * kWgLbo + (rbT & 7) * 4;
  auto issue = [&](int64_t t, bool PF = false) {
    const int64_t e0 = t * KT;
    const int64_t b0 = e0 / kP;
    const int p0 = static_cast<int>(e0 - b0 * kP);
    if (fast) {
      int p = p0 + 4 * kcT;
      int64_t b = b0;
      if (p >= kP) { p -= kP; b += 1; }
      if (e0 + 4 * kcT >= E) { b = 0; p = 0; }
      const int64_t off_co = (b * CO + rbT) * kP + p, off_ci = (b * CI8 + rbT) * kP + p;
#pragma unroll
      for (int u = 0; u < kItemsMax; ++u) {
        constexpr int kUA = CO / kRS, kUB = CI8 / kRS;              // items per class
        const int m = u < kUA ? u : u < 2 * kUA ? u - kUA : u < 2 * kUA + kUB ? u - 2 * kUA : u - 2 * kUA - kUB;
        const float* src = u < kUA ? dy1 + off_co : u < 2 * kUA ? dy2 + off_co : u < 2 * kUA + kUB ? G + off_ci : X + off_ci;
        if (u < 2 * kUA + 2 * kUB) {
          if (PF) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + static_cast<int64_t>(m * kRS) * kP));
          else vn[u] = tc::ldg_stay4(src + static_cast<int64_t>(m * kRS) * kP);
        }
      }
      return;
    }
#pragma unroll
    for (int u = 0; u < kItemsMax; ++u) {
      int it = tid + u * kTcT;
      if (it >= items) it = items - 1;
      const int r = it / kCh, kc = it % kCh;
      int p = p0 + 4 * kc;
      int64_t b = b0;
      if (p >= kP) { p -= kP; b += 1; }
      if (e0 + 4 * kc >= E) { b = 0; p = 0; }
      const float* src;
      int C, c;
      if (r < CO) { src = dy1; C = CO; c = r; }
      else if (r < 2 * CO) { src = dy2; C = CO; c = r - CO; }
      else if (r < 2 * CO + CI) { src = G; C = CI; c = r - 2 * CO; }
      else { src = X; C = CI; c = r - 2 * CO - CI; }
      vn[u] = tc::ldg_stay4(src + (b * C + c) * kP + p);
    }
  };
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t e0 = t * KT;
    issue(t);                                                       // overlaps the previous tile's MMAs (waited for below)
    // the tile after the next one: into L2 now (the loads of a tile only have the previous tile's MMA time to land: since the
    // integer overhead is gone this kernel waits on them -- long scoreboard is its top stall -- and neither registers nor shared
    // memory are left for a second tile in flight)
    if (fast && t + 2 * static_cast<int64_t>(gridDim.x) < ntiles) issue(t + 2 * static_cast<int64_t>(gridDim.x), true);
    if (!first) {                                                   // the previous tile's MMAs still read the images
      tc::mbar_wait(&bar, phase);
      phase ^= 1;
      tc::fence_after_sync();
    }
    if (fast) {
      const bool oob = e0 + 4 * kcT >= E;
#pragma unroll
      for (int u = 0; u < kItemsMax; ++u) {
        constexpr int kUA = CO / kRS, kUB = CI8 / kRS;
        if (u < 2 * kUA + 2 * kUB) {
          const int m = u < kUA ? u : u < 2 * kUA ? u - kUA : u < 2 * kUA + kUB ? u - 2 * kUA : u - 2 * kUA - kUB;
          float* hi_img = u < kUA ? Aimg : u < 2 * kUA ? Aimg + 2 * kAimg : u < 2 * kUA + kUB ? Bimg : Bimg + 2 * kBimg;
          const int img = u < 2 * kUA ? kAimg : kBimg;
          const int o = oT + m * (kRS / 8) * kSbo;
          const float4 v = oob ? make_float4(0.f, 0.f, 0.f, 0.f) : vn[u];
          uint32_t h[4], l[4];
          tc::split_tf32(v.x, h[0], l[0]); tc::split_tf32(v.y, h[1], l[1]);
          tc::split_tf32(v.z, h[2], l[2]); tc::split_tf32(v.w, h[3], l[3]);
          *reinterpret_cast<uint4*>(hi_img + o) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4*>(hi_img + img + o) = make_uint4(l[0], l[1], l[2], l[3]);
        }
      }
    } else
#pragma unroll
    for (int u = 0; u < kItemsMax; ++u) {
      const int it = tid + u * kTcT;
      if (it < items) {
        const int r = it / kCh, kc = it % kCh;
        const float4 v = (e0 + 4 * kc >= E) ? make_float4(0.f, 0.f, 0.f, 0.f) : vn[u];
        float* hi_img;
        int rr;
        if (r < CO) { hi_img = Aimg; rr = r; }
        else if (r < 2 * CO) { hi_img = Aimg + 2 * kAimg; rr = r - CO; }
        else if (r < 2 * CO + CI) { hi_img = Bimg; rr = r - 2 * CO; }
        else { hi_img = Bimg + 2 * kBimg; rr = r - 2 * CO - CI; }
        const int img = (r < 2 * CO) ? kAimg : kBimg;
        const int o = (rr >> 3) * kSbo + kc * kWgLbo + (rr & 7) * 4;
        uint32_t h[4], l[4];
        tc::split_tf32(v.x, h[0], l[0]); tc::split_tf32(v.y, h[1], l[1]);
        tc::split_tf32(v.z, h[2], l[2]); tc::split_tf32(v.w, h[3], l[3]);
        *reinterpret_cast<uint4*>(hi_img + o) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(hi_img + img + o) = make_uint4(l[0], l[1], l[2], l[3]);
      }
    }
    tc::fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_sync();
      const uint32_t idesc = tc::make_idesc_tf32(64, N);
      const uint32_t lbo = kWgLbo * 4, sbo = kSbo * 4;
      const uint32_t a0 = tc::smem_u32(Aimg), bb = tc::smem_u32(Bimg);
#pragma unroll
      for (int ks = 0; ks < KT / 8; ++ks) {
        const uint32_t offs = ks * 2 * lbo;
        const uint32_t acc = (!first || ks > 0) ? 1u : 0u;
        const uint64_t a1h = tc::make_smem_desc(a0 + offs, lbo, sbo), a1l = tc::make_smem_desc(a0 + kAimg * 4 + offs, lbo, sbo);
        const uint64_t a2h = tc::make_smem_desc(a0 + 2 * kAimg * 4 + offs, lbo, sbo), a2l = tc::make_smem_desc(a0 + 3 * kAimg * 4 + offs, lbo, sbo);
        const uint64_t b1h = tc::make_smem_desc(bb + offs, lbo, sbo), b1l = tc::make_smem_desc(bb + kBimg * 4 + offs, lbo, sbo);
        const uint64_t b2h = tc::make_smem_desc(bb + 2 * kBimg * 4 + offs, lbo, sbo), b2l = tc::make_smem_desc(bb + 3 * kBimg * 4 + offs, lbo, sbo);
        tc::mma_tf32_ss(tbase, a1h, b1h, idesc, acc);
        tc::mma_tf32_ss(tbase, a1l, b1h, idesc, 1u);
        tc::mma_tf32_ss(tbase, a1h, b1l, idesc, 1u);
        tc::mma_tf32_ss(tbase + N, a2h, b2h, idesc, acc);
        tc::mma_tf32_ss(tbase + N, a2l, b2h, idesc, 1u);
        tc::mma_tf32_ss(tbase + N, a2h, b2l, idesc, 1u);
      }
      tc::mma_commit(&bar);
    }
    first = false;
  }
  // every CTA owns at least one tile (grid <= ntiles): wait for the last commit, then write the partial block
  tc::mbar_wait(&bar, phase);
  tc::fence_after_sync();
  {
    // M = 64 accumulator layout: row m lives in TMEM lane (m % 16) + 32 * (m / 16): warp w holds rows 16 w .. 16 w + 15 in its
    // lanes 0..15
    const uint32_t lane_base = tbase + (static_cast<uint32_t>(warp * 32) << 16);
    const int row = 16 * warp + lane;
    float* dst = part + static_cast<int64_t>(blockIdx.x) * 2 * CO * (CI + 1);
#pragma unroll
    for (int q = 0; q < (2 * N + 15) / 16; ++q) {
      uint32_t v[16];
      tc::tmem_ld16(lane_base + q * 16, v);
      tc::wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int c = q * 16 + j;
        if (c < 2 * N && lane < 16 && row < CO) {
          const int branch = c / N, n = c % N;
          if (n < CI) dst[(branch * CO + row) * (CI + 1) + n] = __uint_as_float(v[j]);
          else if (n == CI8) dst[(branch * CO + row) * (CI + 1) + CI] = __uint_as_float(v[j]);
        }
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tbase, kAlloc);
}

// second stage: dW[co,ci] += sum_blocks part, db[co] += ...; 32 x 8 blocks over the (branch, co, n) elements, fixed order
__global__ void tc_wgrad_reduce_kernel(const float* __restrict__ part, int nblk, int CO, int CI, float* dW1, float* db1,
                                       float* dW2, float* db2) {
  __shared__ double sh[kPsRows][33];
  const int per = CO * (CI + 1);
  const int i = blockIdx.x * 32 + threadIdx.x;
  const double s = partial_sum_block(part, nblk, 2 * per, i, i < 2 * per, sh);
  if (threadIdx.y != 0 || i >= 2 * per) return;
  const int branch = i / per, r = i % per, co = r / (CI + 1), n = r % (CI + 1);
  if (n < CI) { float* dW = branch ? dW2 : dW1; dW[co * CI + n] += static_cast<float>(s); }
  else { float* db = branch ? db2 : db1; if (db) db[co] += static_cast<float>(s); }
}

}  // namespace coskad
