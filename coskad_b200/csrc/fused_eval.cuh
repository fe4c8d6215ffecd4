// fused_eval.cuh -- the fused eval hot path: one persistent sm_100a kernel that takes a tile of
// kNW pose windows from HBM, runs the four ST_GCNN layers, the linear head, the latent geometry
// and the distance-to-center score with every activation resident in shared memory, and (for
// the auto-encoder) continues through the decoder and the reconstruction score.
//
// Reference arithmetic reproduced (paths relative to the COSKAD tree):
//   ST_GCNN_layer.forward           models/graph_layers/stsgcn.py:94-116
//   ConvTemporalGraphical.forward   models/graph_layers/stsgcn.py:143-156
//   Encoder/Decoder.forward         models/common/components.py:94-105,168-179
//   STSE.encode / STSAE.decode      models/sts/ae.py:76-105,210-230
//   scores                          utils/eval_utils.py:57-106, eval_COSKAD.py:186-199
//
// Eval-mode algebra used (DESIGN.md "Kernel 1"):
//   * BatchNorm folded into the 1x1 convs:  out = PReLU(W1' G(X) + W2' X + b')
//   * for layers with c_out < c_in the channel mixing is applied BEFORE the graph contraction
//     (G acts on (t,v), W1' on c: they commute), halving the contraction work;
//   * the first decoder layer is linear in z, so rev_btlnk + that layer's linear part collapse
//     into one [32*204, latent] matrix M (built in float64 at set_decoder time).
#pragma once
#include "common.cuh"
#include "geometry.cuh"

namespace coskad {

// Offsets (floats) of the mixing blob of one layer: [K*COUT weights][COUT bias][slope,0,0,0]
__host__ __device__ constexpr int mix_blob_floats(int K, int COUT) { return K * COUT + COUT + 4; }

struct FusedParams {
  // encoder layers 0..3
  const float* eTw[4];
  const float* eAw[4];
  const float* eWm[4];
  const float* head_w;   // [16][kF] zero padded rows
  const float* head_b;   // [16]
  // decoder: folded first layer + layers 1..3
  const float* dM;       // [32*204][DL]
  const float* dm0;      // [32*204]
  float d_slope0;
  int DL;                // latent dim of the decoder input (8 or 16)
  const float* dTw[3];
  const float* dAw[3];
  const float* dWm[3];
  // io
  const float* x;        // [B,2,12,17]
  const float* center;   // [D]
  float* z;              // [B, head_rows] or null
  float* score;          // [B] or null        (latent score)
  float* xhat;           // [B,2,12,17] or null
  float* rec_score;      // [B] or null
  int64_t B;
  int head_rows;         // rows of the head written to z
  int D;                 // latent dim used by the geometry (<= head_rows)
  int flavour;
  float* dbg;            // debug: dump R0,R1,GB,GB2 of tile 0 after stage dbg_stage, then exit
  int dbg_stage;
};

// ------------------------------------------------------------------------------------------
// shared memory plan (floats)
constexpr int kRBig = kNW * 32 * kCS;                 // 96 rows
constexpr int kRSmall = kNW * 2 * kCS;                // 6 rows
constexpr int kWMsFloats = mix_blob_floats(32, 32);   // 1060: L1, L3, D2, D4
constexpr int kWMbFloats = mix_blob_floats(64, 64);   // 4164: L2, L4, D3
constexpr int kSmemFloats = 2 * kRBig          // R0, R1
                            + 2 * kRSmall      // XB[2]
                            + 2 * kRSmall      // GB, GB2
                            + kTwFloats + kAwFloats + kWMsFloats + kWMbFloats
                            + kWarps * kNW * kDP   // zpart
                            + kNW * kDP            // zfin
                            + 32;                  // center + pad
constexpr int kSmemBytes = kSmemFloats * 4;
static_assert(kSmemBytes <= 227 * 1024, "shared memory plan exceeds 227 KB");

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void async_copy_floats(float* dst, const float* src, int nfloats, int tid) {
  // nfloats % 4 == 0, both 16B aligned
  for (int i = tid * 4; i < nfloats; i += kThreads * 4) cp_async16(dst + i, src + i);
}

// temporal contraction: dst[row][q][v] = sum_t src[row][t][v] * Tw[v][t][q]   (stsgcn.py:154)
// lanes walk rows (= window*C + channel), v is warp-uniform.  src == dst is allowed.
template <int ROWS, int NWARPS = kWarps>
__device__ __forceinline__ void temporal_stage(const float* src, float* dst, const float* Tw, int warp, int lane) {
  constexpr int CHUNKS = (ROWS + 31) / 32;
  constexpr int NTASK = kV * CHUNKS;
  for (int task = warp; task < NTASK; task += NWARPS) {
    const int v = task % kV;
    const int row = (task / kV) * 32 + lane;
    if (row < ROWS) {
      const float* s = src + row * kCS + v;
      float x[kT], acc[kT];
#pragma unroll
      for (int t = 0; t < kT; ++t) { x[t] = s[t * kV]; acc[t] = 0.f; }
      const float4* w4 = reinterpret_cast<const float4*>(Tw + v * (kT * kT));
#pragma unroll
      for (int t = 0; t < kT; ++t) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float4 w = w4[t * 3 + j];
          acc[4 * j + 0] = fmaf(x[t], w.x, acc[4 * j + 0]);
          acc[4 * j + 1] = fmaf(x[t], w.y, acc[4 * j + 1]);
          acc[4 * j + 2] = fmaf(x[t], w.z, acc[4 * j + 2]);
          acc[4 * j + 3] = fmaf(x[t], w.w, acc[4 * j + 3]);
        }
      }
      float* d = dst + row * kCS + v;
#pragma unroll
      for (int q = 0; q < kT; ++q) d[q * kV] = acc[q];
    }
  }
}

struct EpiIdentity {
  __device__ __forceinline__ float operator()(float v, int, int) const { return v; }
};
// out = PReLU(G(U) + Rsd) for the mix-first layers
struct EpiAddResPrelu {
  const float* rsd;
  float slope;
  __device__ __forceinline__ float operator()(float v, int row, int p) const {
    return prelu(v + rsd[row * kCS + p], slope);
  }
};

// spatial contraction in place: buf[row][t][w] = epi(sum_v buf[row][t][v] * Aw[t][v][w])   (stsgcn.py:155)
template <int ROWS, class Epi, int NWARPS = kWarps>
__device__ __forceinline__ void spatial_stage(float* buf, const float* Aw, const Epi epi, int warp, int lane) {
  constexpr int CHUNKS = (ROWS + 31) / 32;
  constexpr int NTASK = kT * CHUNKS;
  for (int task = warp; task < NTASK; task += NWARPS) {
    const int t = task % kT;
    const int row = (task / kT) * 32 + lane;
    if (row < ROWS) {
      float* s = buf + row * kCS + t * kV;
      float g[kV], acc[kV];
#pragma unroll
      for (int v = 0; v < kV; ++v) { g[v] = s[v]; acc[v] = 0.f; }
      const float4* a4 = reinterpret_cast<const float4*>(Aw + t * (kV * kAW));
#pragma unroll
      for (int v = 0; v < kV; ++v) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 w = a4[v * 5 + j];
          acc[4 * j + 0] = fmaf(g[v], w.x, acc[4 * j + 0]);
          acc[4 * j + 1] = fmaf(g[v], w.y, acc[4 * j + 1]);
          acc[4 * j + 2] = fmaf(g[v], w.z, acc[4 * j + 2]);
          acc[4 * j + 3] = fmaf(g[v], w.w, acc[4 * j + 3]);
        }
        acc[16] = fmaf(g[v], Aw[t * (kV * kAW) + v * kAW + 16], acc[16]);
      }
#pragma unroll
      for (int w = 0; w < kV; ++w) s[w] = epi(acc[w], row, t * kV + w);
    }
  }
}

// channel mixing: acc[n][j] = sum_k in_k[n][p] * Wm[k][co0+j]; lanes walk positions, each thread
// keeps kNW windows x RCO output channels in registers.  Inputs k < K1 come from src1 (rows
// n*K1+k), the rest from src2 (rows n*K2+k-K1).
template <int K1, int K2, int COUT, int RCO, class Epi, int NWARPS = kWarps>
__device__ __forceinline__ void mix_stage(const float* src1, const float* src2, const float* Wm, Epi& epi,
                                          int warp, int lane) {
  static_assert(RCO % 4 == 0 && COUT % RCO == 0, "bad register tile");
  constexpr int NCO = COUT / RCO;
  constexpr int NTASK = kPCH * NCO;
  for (int task = warp; task < NTASK; task += NWARPS) {
    const int p = (task % kPCH) * 32 + lane;
    const int co0 = (task / kPCH) * RCO;
    const bool valid = p < kP;
    const int pc = valid ? p : kP - 1;
    float acc[kNW][RCO];
#pragma unroll
    for (int n = 0; n < kNW; ++n)
#pragma unroll
      for (int j = 0; j < RCO; ++j) acc[n][j] = 0.f;
#pragma unroll 2
    for (int k = 0; k < K1; ++k) {
      float in[kNW];
#pragma unroll
      for (int n = 0; n < kNW; ++n) in[n] = src1[(n * K1 + k) * kCS + pc];
      const float4* w4 = reinterpret_cast<const float4*>(Wm + k * COUT + co0);
#pragma unroll
      for (int j = 0; j < RCO / 4; ++j) {
        const float4 w = w4[j];
#pragma unroll
        for (int n = 0; n < kNW; ++n) {
          acc[n][4 * j + 0] = fmaf(in[n], w.x, acc[n][4 * j + 0]);
          acc[n][4 * j + 1] = fmaf(in[n], w.y, acc[n][4 * j + 1]);
          acc[n][4 * j + 2] = fmaf(in[n], w.z, acc[n][4 * j + 2]);
          acc[n][4 * j + 3] = fmaf(in[n], w.w, acc[n][4 * j + 3]);
        }
      }
    }
    if constexpr (K2 > 0) {
#pragma unroll 2
      for (int k = 0; k < K2; ++k) {
        float in[kNW];
#pragma unroll
        for (int n = 0; n < kNW; ++n) in[n] = src2[(n * K2 + k) * kCS + pc];
        const float4* w4 = reinterpret_cast<const float4*>(Wm + (K1 + k) * COUT + co0);
#pragma unroll
        for (int j = 0; j < RCO / 4; ++j) {
          const float4 w = w4[j];
#pragma unroll
          for (int n = 0; n < kNW; ++n) {
            acc[n][4 * j + 0] = fmaf(in[n], w.x, acc[n][4 * j + 0]);
            acc[n][4 * j + 1] = fmaf(in[n], w.y, acc[n][4 * j + 1]);
            acc[n][4 * j + 2] = fmaf(in[n], w.z, acc[n][4 * j + 2]);
            acc[n][4 * j + 3] = fmaf(in[n], w.w, acc[n][4 * j + 3]);
          }
        }
      }
    }
    epi.template apply<RCO>(acc, co0, p, valid);
  }
}

// dst[(n*COUT+co)][p] = PReLU(acc + bias[co])
template <int COUT>
struct EpiStorePrelu {
  float* dst;
  const float* bias;
  float slope;
  template <int RCO>
  __device__ __forceinline__ void apply(float (&acc)[kNW][RCO], int co0, int p, bool valid) {
    if (!valid) return;
#pragma unroll
    for (int j = 0; j < RCO; ++j) {
      const float b = bias[co0 + j];
#pragma unroll
      for (int n = 0; n < kNW; ++n) dst[(n * COUT + co0 + j) * kCS + p] = prelu(acc[n][j] + b, slope);
    }
  }
};
// mix-first layers: columns [0,CO) -> U rows, columns [CO,2CO) -> Rsd rows (+ folded bias)
template <int CO>
struct EpiSplit {
  float* dstU;
  float* dstR;
  const float* bias;   // [2*CO], zero for the U half
  template <int RCO>
  __device__ __forceinline__ void apply(float (&acc)[kNW][RCO], int co0, int p, bool valid) {
    if (!valid) return;
#pragma unroll
    for (int j = 0; j < RCO; ++j) {
      const int co = co0 + j;
      const float b = bias[co];
      float* d = (co < CO) ? dstU : dstR;
      const int c = (co < CO) ? co : co - CO;
#pragma unroll
      for (int n = 0; n < kNW; ++n) d[(n * CO + c) * kCS + p] = acc[n][j] + b;
    }
  }
};
// last encoder layer: PReLU then the linear head straight from registers (no H4 in memory):
// z[n][d] += h[n][co][p] * head_w[d][co*204 + p]      (models/sts/ae.py:96-101 flatten order (c,t,v))
struct EpiHead {
  const float* bias;
  float slope;
  const float* head_w;
  float z[kNW][kDP];    // per-thread partial sums of the head
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int n = 0; n < kNW; ++n)
#pragma unroll
      for (int d = 0; d < kDP; ++d) z[n][d] = 0.f;
  }
  template <int RCO>
  __device__ __forceinline__ void apply(float (&acc)[kNW][RCO], int co0, int p, bool valid) {
    if (!valid) return;
#pragma unroll
    for (int j = 0; j < RCO; ++j) {
      const float b = bias[co0 + j];
      float h[kNW];
#pragma unroll
      for (int n = 0; n < kNW; ++n) h[n] = prelu(acc[n][j] + b, slope);
      const float* wp = head_w + (co0 + j) * kP + p;
      float w[kDP];
#pragma unroll
      for (int d = 0; d < kDP; ++d) w[d] = __ldg(wp + d * kF);
#pragma unroll
      for (int d = 0; d < kDP; ++d)
#pragma unroll
        for (int n = 0; n < kNW; ++n) z[n][d] = fmaf(h[n], w[d], z[n][d]);
    }
  }
};

template <bool kDec>
__global__ void __launch_bounds__(kThreads, 1) fused_eval_kernel(const __grid_constant__ FusedParams P) {
  extern __shared__ __align__(128) float smem[];
  float* R0 = smem;
  float* R1 = R0 + kRBig;
  float* XB = R1 + kRBig;            // 2 x kRSmall
  float* GB = XB + 2 * kRSmall;
  float* GB2 = GB + kRSmall;
  float* TB = GB2 + kRSmall;
  float* AB = TB + kTwFloats;
  float* WMs = AB + kAwFloats;
  float* WMb = WMs + kWMsFloats;
  float* zpart = WMb + kWMbFloats;   // [kWarps][kNW*kDP]
  float* zfin = zpart + kWarps * kNW * kDP;
  float* cen = zfin + kNW * kDP;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (P.B + kNW - 1) / kNW;
  if (static_cast<int64_t>(blockIdx.x) >= ntiles) return;

  auto load_x = [&](float* dst, int64_t tile) {
    // rows r = n*2+c of the tile are contiguous in HBM: x[(w0+n)*408 + c*204 + p]
    const int64_t w0 = tile * kNW;
    for (int i = tid; i < kNW * 2 * kP; i += kThreads) {
      const int r = i / kP, p = i - r * kP;
      int64_t w = w0 + (r >> 1);
      if (w >= P.B) w = P.B - 1;     // ragged last tile: replicate the last window, never stored
      cp_async4(dst + r * kCS + p, P.x + w * (2 * kP) + (r & 1) * kP + p);
    }
  };
  auto boundary = [&]() { cp_async_wait_all(); __syncthreads(); };
  // debug aid (tests only): after stage k of the first tile, CTA 0 copies its activation buffers out
  auto dbg_dump = [&](int k) -> bool {
    if (P.dbg == nullptr || P.dbg_stage != k) return false;
    __syncthreads();
    if (blockIdx.x == 0) {
      for (int i = tid; i < 2 * kRBig; i += kThreads) P.dbg[i] = R0[i];
      for (int i = tid; i < 2 * kRSmall; i += kThreads) P.dbg[2 * kRBig + i] = GB[i];
      for (int i = tid; i < kNW * kDP; i += kThreads) P.dbg[2 * kRBig + 2 * kRSmall + i] = zfin[i];
    }
    return true;
  };
#define COSKAD_DBG(k) if (dbg_dump(k)) { cp_async_wait_all(); return; }

  // prologue
  if (tid < 32) cen[tid] = (P.center != nullptr && tid < P.D) ? P.center[tid] : 0.f;
  load_x(XB, blockIdx.x);
  async_copy_floats(TB, P.eTw[0], kTwFloats, tid);
  async_copy_floats(AB, P.eAw[0], kAwFloats, tid);
  async_copy_floats(WMs, P.eWm[0], mix_blob_floats(4, 32), tid);
  async_copy_floats(WMb, P.eWm[1], mix_blob_floats(32, 32), tid);
  cp_async_commit();

  int cur = 0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, cur ^= 1) {
    float* X0 = XB + cur * kRSmall;
    const int64_t next_tile = tile + gridDim.x;

    // ---- S0: L1 temporal  X0 -> GB ------------------------------------------------------
    boundary();
    temporal_stage<kNW * kC0>(X0, GB, TB, warp, lane);
    COSKAD_DBG(0);
    // ---- S1: L1 spatial in place on GB ---------------------------------------------------
    boundary();
    async_copy_floats(TB, P.eTw[1], kTwFloats, tid);
    if (next_tile < ntiles) load_x(XB + (cur ^ 1) * kRSmall, next_tile);
    cp_async_commit();
    spatial_stage<kNW * kC0>(GB, AB, EpiIdentity{}, warp, lane);
    COSKAD_DBG(1);
    // ---- S2: L1 mix (G=GB, X=X0) -> R0 (32 ch) -------------------------------------------
    boundary();
    async_copy_floats(AB, P.eAw[1], kAwFloats, tid);
    cp_async_commit();
    {
      EpiStorePrelu<kC1> epi{R0, WMs + 4 * kC1, WMs[4 * kC1 + kC1]};
      mix_stage<kC0, kC0, kC1, 16>(GB, X0, WMs, epi, warp, lane);
    }
    COSKAD_DBG(2);
    // ---- S3: L2 mix-first: R0 -> U (R1 rows 0..47), Rsd (R1 rows 48..95) ----------------
    boundary();
    async_copy_floats(WMs, P.eWm[2], mix_blob_floats(32, 32), tid);
    cp_async_commit();
    float* U2 = R1;
    float* Rsd2 = R1 + kNW * kC2 * kCS;
    const float slope2 = WMb[kC1 * 2 * kC2 + 2 * kC2];
    {
      EpiSplit<kC2> epi{U2, Rsd2, WMb + kC1 * 2 * kC2};
      mix_stage<kC1, 0, 2 * kC2, 16>(R0, nullptr, WMb, epi, warp, lane);
    }
    COSKAD_DBG(3);
    // ---- S4: L2 temporal in place on U ---------------------------------------------------
    boundary();
    async_copy_floats(WMb, P.eWm[3], mix_blob_floats(64, 64), tid);
    cp_async_commit();
    temporal_stage<kNW * kC2>(U2, U2, TB, warp, lane);
    COSKAD_DBG(4);
    // ---- S5: L2 spatial in place + residual + PReLU -> H2 = R1 rows 0..47 ---------------
    boundary();
    async_copy_floats(TB, P.eTw[2], kTwFloats, tid);
    cp_async_commit();
    spatial_stage<kNW * kC2>(U2, AB, EpiAddResPrelu{Rsd2, slope2}, warp, lane);
    COSKAD_DBG(5);
    // ---- S6: L3 temporal: H2 -> G3 = R1 rows 48..95 --------------------------------------
    boundary();
    async_copy_floats(AB, P.eAw[2], kAwFloats, tid);
    cp_async_commit();
    float* H2 = R1;
    float* G3 = R1 + kNW * kC2 * kCS;
    temporal_stage<kNW * kC2>(H2, G3, TB, warp, lane);
    COSKAD_DBG(6);
    // ---- S7: L3 spatial in place on G3 ---------------------------------------------------
    boundary();
    async_copy_floats(TB, P.eTw[3], kTwFloats, tid);
    cp_async_commit();
    spatial_stage<kNW * kC2>(G3, AB, EpiIdentity{}, warp, lane);
    COSKAD_DBG(7);
    // ---- S8: L3 mix (G3, H2) -> H3 = R0 (32 ch) ------------------------------------------
    boundary();
    async_copy_floats(AB, P.eAw[3], kAwFloats, tid);
    cp_async_commit();
    {
      EpiStorePrelu<kC3> epi{R0, WMs + 2 * kC2 * kC3, WMs[2 * kC2 * kC3 + kC3]};
      mix_stage<kC2, kC2, kC3, 16>(G3, H2, WMs, epi, warp, lane);
    }
    COSKAD_DBG(8);
    // ---- S9: L4 temporal: H3 (R0) -> G4 (R1) ---------------------------------------------
    boundary();
    if (kDec) async_copy_floats(WMs, P.dWm[0], mix_blob_floats(32, 32), tid);
    else async_copy_floats(WMs, P.eWm[0], mix_blob_floats(4, 32), tid);
    cp_async_commit();
    temporal_stage<kNW * kC3>(R0, R1, TB, warp, lane);
    COSKAD_DBG(9);
    // ---- S10: L4 spatial in place on R1 --------------------------------------------------
    boundary();
    async_copy_floats(TB, kDec ? P.dTw[0] : P.eTw[0], kTwFloats, tid);
    cp_async_commit();
    spatial_stage<kNW * kC3>(R1, AB, EpiIdentity{}, warp, lane);
    COSKAD_DBG(10);
    // ---- S11: L4 mix (G4=R1, H3=R0) + PReLU + head, H4 never stored ----------------------
    boundary();
    async_copy_floats(AB, kDec ? P.dAw[0] : P.eAw[0], kAwFloats, tid);
    cp_async_commit();
    {
      EpiHead epi;
      epi.bias = WMb + 2 * kC3 * kC4;
      epi.slope = WMb[2 * kC3 * kC4 + kC4];
      epi.head_w = P.head_w;
      epi.clear();
      mix_stage<kC3, kC3, kC4, 16>(R1, R0, WMb, epi, warp, lane);
#pragma unroll
      for (int n = 0; n < kNW; ++n)
#pragma unroll
        for (int d = 0; d < kDP; ++d) {
          const float s = warp_sum(epi.z[n][d]);
          if (lane == 0) zpart[warp * (kNW * kDP) + n * kDP + d] = s;
        }
    }
    COSKAD_DBG(11);
    // ---- S12: head reduce, geometry, score -----------------------------------------------
    boundary();
    if (kDec) async_copy_floats(WMb, P.dWm[1], mix_blob_floats(32, 32), tid);
    else async_copy_floats(WMb, P.eWm[1], mix_blob_floats(32, 32), tid);
    cp_async_commit();
    if (tid < kNW * kDP) {
      float s = __ldg(P.head_b + (tid % kDP));
#pragma unroll
      for (int w = 0; w < kWarps; ++w) s += zpart[w * (kNW * kDP) + tid];
      zfin[tid] = s;
    }
    __syncthreads();
    if (warp < kNW) {
      const int64_t w = tile * kNW + warp;
      if (w < P.B) {
        float u[1] = {lane < kDP ? zfin[warp * kDP + lane] : 0.f};
        if (P.z != nullptr && lane < P.head_rows) P.z[w * P.head_rows + lane] = u[0];
        if (P.score != nullptr) {
          if (lane >= P.D) u[0] = 0.f;
          const float c[1] = {cen[lane]};
          const float sc = score_from_latent<1>(P.flavour, u, c, P.D);
          if (lane == 0) P.score[w] = sc;
        }
      }
    }

    if constexpr (!kDec) { COSKAD_DBG(12); }
    if constexpr (kDec) {
      COSKAD_DBG(12);
      // ---- S13: folded first decoder layer: H = PReLU(M z + m0) -> R0 (32 ch) -----------
      // (rev_btlnk models/sts/ae.py:222 + decoder layer 0 linear part, collapsed at set_decoder)
      {
        const int DL = P.DL;
        for (int i = tid; i < kC1 * kP; i += kThreads) {
          const int co = i / kP, p = i - co * kP;
          const float m0 = __ldg(P.dm0 + i);
          float o[kNW];
#pragma unroll
          for (int n = 0; n < kNW; ++n) o[n] = m0;
          const float4* m4 = reinterpret_cast<const float4*>(P.dM + static_cast<size_t>(i) * DL);
          for (int d4 = 0; d4 < DL / 4; ++d4) {
            const float4 m = __ldg(m4 + d4);
#pragma unroll
            for (int n = 0; n < kNW; ++n) {
              const float* zz = zfin + n * kDP + d4 * 4;
              o[n] = fmaf(m.x, zz[0], o[n]);
              o[n] = fmaf(m.y, zz[1], o[n]);
              o[n] = fmaf(m.z, zz[2], o[n]);
              o[n] = fmaf(m.w, zz[3], o[n]);
            }
          }
#pragma unroll
          for (int n = 0; n < kNW; ++n) R0[(n * kC1 + co) * kCS + p] = prelu(o[n], P.d_slope0);
        }
      }
      COSKAD_DBG(13);
      // ---- S14: D2 (32->16) mix-first: R0 -> U (R1 rows 0..47), Rsd (rows 48..95) --------
      boundary();
      float* Ud = R1;
      float* Rsdd = R1 + kNW * kC2 * kCS;
      const float dslope1 = WMs[kC1 * 2 * kC2 + 2 * kC2];
      {
        EpiSplit<kC2> epi{Ud, Rsdd, WMs + kC1 * 2 * kC2};
        mix_stage<kC1, 0, 2 * kC2, 16>(R0, nullptr, WMs, epi, warp, lane);
      }
      COSKAD_DBG(14);
      // ---- S15: D2 temporal in place ------------------------------------------------------
      boundary();
      async_copy_floats(WMs, P.dWm[2], mix_blob_floats(32, 4), tid);
      cp_async_commit();
      temporal_stage<kNW * kC2>(Ud, Ud, TB, warp, lane);
      COSKAD_DBG(15);
      // ---- S16: D2 spatial + residual + PReLU -> R1 rows 0..47 ----------------------------
      boundary();
      async_copy_floats(TB, P.dTw[1], kTwFloats, tid);
      cp_async_commit();
      spatial_stage<kNW * kC2>(Ud, AB, EpiAddResPrelu{Rsdd, dslope1}, warp, lane);
      COSKAD_DBG(16);
      // ---- S17: D3 (16->32) temporal: R1 rows 0..47 -> rows 48..95 -----------------------
      boundary();
      async_copy_floats(AB, P.dAw[1], kAwFloats, tid);
      cp_async_commit();
      temporal_stage<kNW * kC2>(R1, R1 + kNW * kC2 * kCS, TB, warp, lane);
      COSKAD_DBG(17);
      // ---- S18: D3 spatial in place --------------------------------------------------------
      boundary();
      async_copy_floats(TB, P.dTw[2], kTwFloats, tid);
      cp_async_commit();
      spatial_stage<kNW * kC2>(R1 + kNW * kC2 * kCS, AB, EpiIdentity{}, warp, lane);
      COSKAD_DBG(18);
      // ---- S19: D3 mix -> R0 (32 ch) -------------------------------------------------------
      boundary();
      async_copy_floats(AB, P.dAw[2], kAwFloats, tid);
      cp_async_commit();
      {
        EpiStorePrelu<kC3> epi{R0, WMb + 2 * kC2 * kC3, WMb[2 * kC2 * kC3 + kC3]};
        mix_stage<kC2, kC2, kC3, 16>(R1 + kNW * kC2 * kCS, R1, WMb, epi, warp, lane);
      }
      COSKAD_DBG(19);
      // ---- S20: D4 (32->2) mix-first: R0 -> U (GB), Rsd (GB2) ------------------------------
      boundary();
      async_copy_floats(WMb, P.eWm[1], mix_blob_floats(32, 32), tid);
      cp_async_commit();
      const float dslope3 = WMs[kC3 * 2 * kC0 + 2 * kC0];
      {
        EpiSplit<kC0> epi{GB, GB2, WMs + kC3 * 2 * kC0};
        mix_stage<kC3, 0, 2 * kC0, 4>(R0, nullptr, WMs, epi, warp, lane);
      }
      COSKAD_DBG(20);
      // ---- S21: D4 temporal in place on GB -------------------------------------------------
      boundary();
      async_copy_floats(WMs, P.eWm[0], mix_blob_floats(4, 32), tid);
      cp_async_commit();
      temporal_stage<kNW * kC0>(GB, GB, TB, warp, lane);
      COSKAD_DBG(21);
      // ---- S22: D4 spatial + residual + PReLU -> xhat in GB --------------------------------
      boundary();
      async_copy_floats(TB, P.eTw[0], kTwFloats, tid);
      cp_async_commit();
      spatial_stage<kNW * kC0>(GB, AB, EpiAddResPrelu{GB2, dslope3}, warp, lane);
      COSKAD_DBG(22);
      // ---- S23: reconstruction score mean_{c,t,v}(x - xhat)^2, optional xhat store --------
      boundary();
      async_copy_floats(AB, P.eAw[0], kAwFloats, tid);
      cp_async_commit();
      if (warp < kNW) {
        const int64_t w = tile * kNW + warp;
        if (w < P.B) {
          float s = 0.f;
          for (int i = lane; i < 2 * kP; i += 32) {
            const int c = i / kP, p = i - c * kP;
            const float xh = GB[(warp * 2 + c) * kCS + p];
            const float d = X0[(warp * 2 + c) * kCS + p] - xh;
            s = fmaf(d, d, s);
            if (P.xhat != nullptr) P.xhat[w * (2 * kP) + i] = xh;
          }
          s = warp_sum(s);
          if (lane == 0 && P.rec_score != nullptr) P.rec_score[w] = s / static_cast<float>(2 * kP);
        }
      }
    }
  }
  cp_async_wait_all();
}

}  // namespace coskad

namespace coskad {
// ---- register-blocked contraction stages for C = 32 channel layers (rows = kNW * 32) ---------------------------------
// lane = channel, each lane carries the kNW windows of its channel: every broadcast weight load (2 LSU wavefronts per
// LDS.128) feeds kNW x 4 FFMAs instead of 4 -- the R = 1 stages above are LSU-bound by 2.6x (profiles/r01_v2*).
// The contraction index (t or v) is a real loop (not unrolled): each warp runs a task exactly once per tile, so fully
// unrolled bodies are straight-line code with no reuse and the warps starve on instruction fetch
// (profiles/r01_v3*: stall_no_inst 25-35 % of the samples of these stages).
// one task (v, q-half) of the C = 32 temporal contraction: 6 of the 12 outputs q, accumulated as 3 packed pairs (FFMA2)
__device__ __forceinline__ void temporal_task_c32(const float* src, float* dst, const float* Tw, int task, int lane) {
  const int v = task >> 1, q0 = (task & 1) * 6;
  unsigned long long acc[kNW][3];
#pragma unroll
  for (int n = 0; n < kNW; ++n)
#pragma unroll
    for (int q = 0; q < 3; ++q) acc[n][q] = 0ull;
  const float* s = src + lane * kCS + v;
  const float* w = Tw + v * (kT * kT) + q0;
#pragma unroll 4
  for (int t = 0; t < kT; ++t) {
    unsigned long long x[kNW];
#pragma unroll
    for (int n = 0; n < kNW; ++n) x[n] = dup2(s[n * 32 * kCS + t * kV]);
    const unsigned long long w01 = *reinterpret_cast<const unsigned long long*>(w + t * kT);
    const unsigned long long w23 = *reinterpret_cast<const unsigned long long*>(w + t * kT + 2);
    const unsigned long long w45 = *reinterpret_cast<const unsigned long long*>(w + t * kT + 4);
#pragma unroll
    for (int n = 0; n < kNW; ++n) {
      ffma2(acc[n][0], x[n], w01); ffma2(acc[n][1], x[n], w23); ffma2(acc[n][2], x[n], w45);
    }
  }
#pragma unroll
  for (int n = 0; n < kNW; ++n) {
    float* d = dst + (n * 32 + lane) * kCS + v;
#pragma unroll
    for (int q = 0; q < 3; ++q) { d[(q0 + 2 * q) * kV] = lo2(acc[n][q]); d[(q0 + 2 * q + 1) * kV] = hi2(acc[n][q]); }
  }
}
// next task from a shared-memory counter (zeroed by the caller before the preceding barrier); warp-uniform result
__device__ __forceinline__ int next_task(int* ctr, int lane) {
  int tk = 0;
  if (lane == 0) tk = atomicAdd(ctr, 1);
  return __shfl_sync(0xffffffffu, tk, 0);
}
// `ctr` != nullptr: tasks are handed out dynamically -- used where the warps enter the stage at different times
template <int NWARPS>
__device__ __forceinline__ void temporal_stage_c32(const float* src, float* dst, const float* Tw, int warp, int lane,
                                                   int* ctr = nullptr) {
  for (int task = warp;; task += NWARPS) {        // 34 tasks (v, q-half)
    if (ctr != nullptr) task = next_task(ctr, lane);
    if (task >= 2 * kV) break;
    temporal_task_c32(src, dst, Tw, task, lane);
  }
}

template <int NWARPS>
__device__ __forceinline__ void spatial_stage_c32(float* buf, const float* Aw, int warp, int lane) {
  for (int t = warp; t < kT; t += NWARPS) {
    unsigned long long acc[kNW][8];     // outputs w = 0..15 as packed pairs (FFMA2)
    float acc16[kNW];                   // w = 16
#pragma unroll
    for (int n = 0; n < kNW; ++n) {
#pragma unroll
      for (int w = 0; w < 8; ++w) acc[n][w] = 0ull;
      acc16[n] = 0.f;
    }
    float* s = buf + lane * kCS + t * kV;
    const float* a = Aw + t * (kV * kAW);
#pragma unroll 1
    for (int v = 0; v < kV; ++v) {
      float g[kNW];
#pragma unroll
      for (int n = 0; n < kNW; ++n) g[n] = s[n * 32 * kCS + v];
      const ulonglong2* a4 = reinterpret_cast<const ulonglong2*>(a + v * kAW);
      const float w16 = a[v * kAW + 16];
      unsigned long long g2[kNW];
#pragma unroll
      for (int n = 0; n < kNW; ++n) g2[n] = dup2(g[n]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const ulonglong2 w = a4[j];
#pragma unroll
        for (int n = 0; n < kNW; ++n) { ffma2(acc[n][2 * j], g2[n], w.x); ffma2(acc[n][2 * j + 1], g2[n], w.y); }
      }
#pragma unroll
      for (int n = 0; n < kNW; ++n) acc16[n] = fmaf(g[n], w16, acc16[n]);
    }
#pragma unroll
    for (int n = 0; n < kNW; ++n) {
#pragma unroll
      for (int w = 0; w < 8; ++w) { s[n * 32 * kCS + 2 * w] = lo2(acc[n][w]); s[n * 32 * kCS + 2 * w + 1] = hi2(acc[n][w]); }
      s[n * 32 * kCS + 16] = acc16[n];
    }
  }
}
}  // namespace coskad

namespace coskad {
// ---- layer-1 contraction stages (rows = kNW * 2 = 6): lanes = (row, output group) instead of rows only --------------
// With 2 input channels the generic stages keep 6 of 32 lanes busy; here a warp task covers one joint v (temporal) or
// one frame t (spatial) for all 6 rows, each lane producing 3 (4) of the 12 (17) outputs.
template <int NWARPS>
__device__ __forceinline__ void temporal_stage_l1(const float* src, float* dst, const float* Tw, int warp, int lane) {
  constexpr int ROWS = kNW * kC0;                 // 6
  const int row = lane % ROWS, qg = lane / ROWS;  // qg 0..3 -> outputs q = 3 qg .. 3 qg + 2 (lanes >= 24 idle)
  for (int v = warp; v < kV; v += NWARPS) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    if (qg < 4) {
      const float* s = src + row * kCS + v;
      const float* w = Tw + v * (kT * kT) + 3 * qg;
#pragma unroll
      for (int t = 0; t < kT; ++t) {
        const float x = s[t * kV];
        a0 = fmaf(x, w[t * kT + 0], a0); a1 = fmaf(x, w[t * kT + 1], a1); a2 = fmaf(x, w[t * kT + 2], a2);
      }
    }
    __syncwarp();      // src == dst is allowed: every lane of the row has read the column before anyone overwrites it
    if (qg < 4) {
      float* d = dst + row * kCS + v;
      d[(3 * qg + 0) * kV] = a0; d[(3 * qg + 1) * kV] = a1; d[(3 * qg + 2) * kV] = a2;
    }
  }
}
template <int NWARPS, class Epi = EpiIdentity>
__device__ __forceinline__ void spatial_stage_l1(float* buf, const float* Aw, int warp, int lane, const Epi epi = Epi{}) {
  constexpr int ROWS = kNW * kC0;                 // 6
  const int row = lane % ROWS, wg = lane / ROWS;  // wg 0..4 -> outputs w = 4 wg .. 4 wg + 3 (w < 17; lanes >= 30 idle)
  for (int t = warp; t < kT; t += NWARPS) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    float* s = buf + row * kCS + t * kV;
    if (wg < 5) {
      const float4* a4 = reinterpret_cast<const float4*>(Aw + t * (kV * kAW)) + wg;
#pragma unroll
      for (int v = 0; v < kV; ++v) {
        const float g = s[v];
        const float4 w = a4[v * 5];
        acc[0] = fmaf(g, w.x, acc[0]); acc[1] = fmaf(g, w.y, acc[1]); acc[2] = fmaf(g, w.z, acc[2]); acc[3] = fmaf(g, w.w, acc[3]);
      }
    }
    __syncwarp();      // every lane of the row has read its 17 inputs before anyone overwrites them (in place)
    if (wg < 5) {
#pragma unroll
      for (int j = 0; j < 4; ++j) if (4 * wg + j < kV) s[4 * wg + j] = epi(acc[j], row, t * kV + 4 * wg + j);
    }
  }
}
}  // namespace coskad

namespace coskad {
// ---- contraction stages for C = 16 channel layers (rows = kNW * 16): lane = (channel, output half), kNW windows per lane
template <int NWARPS>
__device__ __forceinline__ void temporal_stage_c16(const float* src, float* dst, const float* Tw, int warp, int lane) {
  const int c = lane & 15, q0 = (lane >> 4) * 6;          // half 0: q 0..5, half 1: q 6..11
  for (int v = warp; v < kV; v += NWARPS) {
    unsigned long long acc[kNW][3];
#pragma unroll
    for (int n = 0; n < kNW; ++n)
#pragma unroll
      for (int q = 0; q < 3; ++q) acc[n][q] = 0ull;
    const float* s = src + c * kCS + v;
    const float* w = Tw + v * (kT * kT) + q0;
#pragma unroll 4
    for (int t = 0; t < kT; ++t) {
      unsigned long long x[kNW];
#pragma unroll
      for (int n = 0; n < kNW; ++n) x[n] = dup2(s[n * 16 * kCS + t * kV]);
      const unsigned long long w01 = *reinterpret_cast<const unsigned long long*>(w + t * kT);
      const unsigned long long w23 = *reinterpret_cast<const unsigned long long*>(w + t * kT + 2);
      const unsigned long long w45 = *reinterpret_cast<const unsigned long long*>(w + t * kT + 4);
#pragma unroll
      for (int n = 0; n < kNW; ++n) {
        ffma2(acc[n][0], x[n], w01); ffma2(acc[n][1], x[n], w23); ffma2(acc[n][2], x[n], w45);
      }
    }
    __syncwarp();      // in-place use: both halves have read the column before either writes it
#pragma unroll
    for (int n = 0; n < kNW; ++n) {
      float* d = dst + (n * 16 + c) * kCS + v;
#pragma unroll
      for (int q = 0; q < 3; ++q) { d[(q0 + 2 * q) * kV] = lo2(acc[n][q]); d[(q0 + 2 * q + 1) * kV] = hi2(acc[n][q]); }
    }
  }
}

// half 0: outputs w 0..7 (two LDS.128 per v), half 1: outputs w 8..16 (two LDS.128 + one LDS.32 per v)
template <class Epi, int NWARPS>
__device__ __forceinline__ void spatial_stage_c16(float* buf, const float* Aw, const Epi epi, int warp, int lane) {
  const int c = lane & 15, h = lane >> 4;
  for (int t = warp; t < kT; t += NWARPS) {
    unsigned long long acc[kNW][4];     // outputs 8 h .. 8 h + 7 as packed pairs (FFMA2)
    float acc8[kNW];
#pragma unroll
    for (int n = 0; n < kNW; ++n) {
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[n][j] = 0ull;
      acc8[n] = 0.f;
    }
    const float* s = buf + c * kCS + t * kV;
    const float* a = Aw + t * (kV * kAW) + 8 * h;
#pragma unroll 1
    for (int v = 0; v < kV; ++v) {
      float g[kNW];
#pragma unroll
      for (int n = 0; n < kNW; ++n) g[n] = s[n * 16 * kCS + v];
      const ulonglong2 w0 = *reinterpret_cast<const ulonglong2*>(a + v * kAW);
      const ulonglong2 w1 = *reinterpret_cast<const ulonglong2*>(a + v * kAW + 4);
      const float w8 = a[v * kAW + 8];                  // half 1: w = 16; half 0: w = 8 (belongs to half 1, discarded)
#pragma unroll
      for (int n = 0; n < kNW; ++n) {
        const unsigned long long g2 = dup2(g[n]);
        ffma2(acc[n][0], g2, w0.x); ffma2(acc[n][1], g2, w0.y);
        ffma2(acc[n][2], g2, w1.x); ffma2(acc[n][3], g2, w1.y);
        acc8[n] = fmaf(g[n], w8, acc8[n]);
      }
    }
    __syncwarp();      // in place: both halves have read the 17 inputs of their rows
#pragma unroll
    for (int n = 0; n < kNW; ++n) {
      const int row = n * 16 + c;
      float* d = buf + row * kCS + t * kV;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        d[8 * h + 2 * j] = epi(lo2(acc[n][j]), row, t * kV + 8 * h + 2 * j);
        d[8 * h + 2 * j + 1] = epi(hi2(acc[n][j]), row, t * kV + 8 * h + 2 * j + 1);
      }
      if (h == 1) d[16] = epi(acc8[n], row, t * kV + 16);
    }
  }
}
}  // namespace coskad
