// fold.cuh -- weight packing for the fused eval kernel (runs once per set_encoder/set_decoder).
//
// Eval-mode BatchNorm2d (eps 1e-5, models/graph_layers/stsgcn.py:65,77) is folded into the 1x1
// convolutions in float64:   BN1(W1 g + b1) + BN2(W2 x + b2) = W1' g + W2' x + b'.
// The packed blobs are what fused_eval.cuh stages through shared memory:
//   Tw [V][T][T]            copy of gcn.T                      (stsgcn.py:138)
//   Aw [T][V][kAW]          gcn.A with rows padded 17 -> 20    (stsgcn.py:134)
//   Wm [K][COUT] + bias[COUT] + {slope,0,0,0}
//        normal layer   (c_out >= c_in): K = 2 c_in (G rows then X rows), COUT = c_out
//        mix-first layer (c_out <  c_in): K = c_in, COUT = 2 c_out (U columns then Rsd columns)
#pragma once
#include "common.cuh"
#include "../../include/coskad_b200.h"

namespace coskad {

__device__ __forceinline__ double bn_scale(const float* w, const float* rv, int c) {
  return static_cast<double>(w[c]) / sqrt(static_cast<double>(rv[c]) + 1e-5);
}
// folded W1'[co][ci], W2'[co][ci], b'[co] of one layer, float64
__device__ __forceinline__ double fold_w1(const coskad_layer_params& L, int co, int ci) {
  return static_cast<double>(L.w1[co * L.c_in + ci]) * bn_scale(L.bn1_w, L.bn1_rv, co);
}
__device__ __forceinline__ double fold_w2(const coskad_layer_params& L, int co, int ci) {
  if (L.w2 == nullptr) return co == ci ? 1.0 : 0.0;       // nn.Identity residual (stsgcn.py:80)
  return static_cast<double>(L.w2[co * L.c_in + ci]) * bn_scale(L.bn2_w, L.bn2_rv, co);
}
__device__ __forceinline__ double fold_b(const coskad_layer_params& L, int co) {
  double b = ((L.b1 ? static_cast<double>(L.b1[co]) : 0.0) - static_cast<double>(L.bn1_rm[co])) *
                 bn_scale(L.bn1_w, L.bn1_rv, co) + static_cast<double>(L.bn1_b[co]);
  if (L.w2 != nullptr)
    b += ((L.b2 ? static_cast<double>(L.b2[co]) : 0.0) - static_cast<double>(L.bn2_rm[co])) *
             bn_scale(L.bn2_w, L.bn2_rv, co) + static_cast<double>(L.bn2_b[co]);
  return b;
}

__global__ void fold_layer_kernel(coskad_layer_params L, int mix_first, float* Tw, float* Aw, float* Wm) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  const int ci_n = L.c_in, co_n = L.c_out;
  for (int i = tid; i < kTwFloats; i += nth) Tw[i] = L.T[i];
  for (int i = tid; i < kAwFloats; i += nth) {
    const int w = i % kAW, tv = i / kAW;
    Aw[i] = (w < kV) ? L.A[tv * kV + w] : 0.f;
  }
  const int K = mix_first ? ci_n : 2 * ci_n;
  const int COUT = mix_first ? 2 * co_n : co_n;
  for (int i = tid; i < K * COUT; i += nth) {
    const int k = i / COUT, j = i % COUT;
    double v;
    if (mix_first) v = (j < co_n) ? fold_w1(L, j, k) : fold_w2(L, j - co_n, k);
    else v = (k < ci_n) ? fold_w1(L, j, k) : fold_w2(L, j, k - ci_n);
    Wm[i] = static_cast<float>(v);
  }
  for (int j = tid; j < COUT; j += nth) {
    double b;
    if (mix_first) b = (j < co_n) ? 0.0 : fold_b(L, j - co_n);
    else b = fold_b(L, j);
    Wm[K * COUT + j] = static_cast<float>(b);
  }
  if (tid < 4) Wm[K * COUT + COUT + tid] = (tid == 0) ? L.prelu[0] : 0.f;
}

// ---- tensor-core blobs: [B_hi image Kp*N][B_lo image Kp*N][bias N][slope,0,0,0] -----------------------------------
// image = canonical K-major, no swizzle: float index ((k/4)*(N/8) + co/8)*32 + (co%8)*4 + (k%4); hi = value with the low
// 13 mantissa bits cleared (exact TF32), lo = TF32-truncated remainder.
// mode 0: K = [G (c_in) | X (c_in)] padded to Kp, N = c_out        (normal layer: L1, L3)
// mode 1: K = X (c_in), N = [U (c_out) | Rsd (c_out)]               (mix-first layer: L2)
// mode 2: K = X (c_in), N = c_out, residual conv only, no bias       (L4 residual half)
// mode 3: K = G (c_in), N = c_out, tcn conv, bias + slope            (L4 graph half)
__global__ void fold_layer_tc_kernel(coskad_layer_params L, int mode, int Kp, int N, float* blob) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  const int ci_n = L.c_in, co_n = L.c_out;
  for (int i = tid; i < Kp * N; i += nth) {
    const int k = i / N, co = i % N;
    double v = 0.0;
    if (mode == 0) { if (k < ci_n) v = fold_w1(L, co, k); else if (k < 2 * ci_n) v = fold_w2(L, co, k - ci_n); }
    else if (mode == 1) { if (k < ci_n) v = (co < co_n) ? fold_w1(L, co, k) : fold_w2(L, co - co_n, k); }
    else if (mode == 2) { if (k < ci_n) v = fold_w2(L, co, k); }
    else { if (k < ci_n) v = fold_w1(L, co, k); }
    const float f = static_cast<float>(v);
    const float hi = __uint_as_float(__float_as_uint(f) & 0xFFFFE000u);
    const float lo = __uint_as_float(__float_as_uint(f - hi) & 0xFFFFE000u);
    const int idx = ((k / 4) * (N / 8) + co / 8) * 32 + (co % 8) * 4 + (k % 4);
    blob[idx] = hi;
    blob[Kp * N + idx] = lo;
  }
  for (int j = tid; j < N; j += nth) {
    double b = 0.0;
    if (mode == 0 || mode == 3) b = fold_b(L, j);
    else if (mode == 1) b = (j < co_n) ? 0.0 : fold_b(L, j - co_n);
    blob[2 * Kp * N + j] = static_cast<float>(b);
  }
  if (tid < 4) blob[2 * Kp * N + N + tid] = (tid == 0) ? L.prelu[0] : 0.f;
}

// head rows padded to kDP with zeros
__global__ void pack_head_kernel(const float* w, const float* b, int rows, float* wp, float* bp) {
  const int64_t tid = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t nth = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = tid; i < static_cast<int64_t>(kDP) * kF; i += nth) {
    const int r = static_cast<int>(i / kF);
    wp[i] = (r < rows) ? w[i] : 0.f;
  }
  for (int64_t i = tid; i < kDP; i += nth) bp[i] = (i < rows && b != nullptr) ? b[i] : 0.f;
}

// head weights for the tensor-core kernel's head stage: wq[c][dq][p][4] = w[4 dq + i][c*204 + p] (rows >= `rows` are 0), so
// that a lane (= position p) fetches 4 latent rows of its (c, p) feature with one LDG.128 and a warp's request covers
// 512 contiguous bytes
__global__ void pack_head4_kernel(const float* w, int rows, float* wq) {
  const int64_t tid = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t nth = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = tid; i < static_cast<int64_t>(kDP) * kF; i += nth) {
    const int e = static_cast<int>(i & 3);
    const int64_t r = i >> 2;
    const int p = static_cast<int>(r % kP);
    const int dq = static_cast<int>((r / kP) % (kDP / 4));
    const int c = static_cast<int>(r / (kP * (kDP / 4)));
    const int d = 4 * dq + e;
    wq[i] = (d < rows) ? w[static_cast<int64_t>(d) * kF + c * kP + p] : 0.f;
  }
}

// ---- decoder first layer collapse (float64) -----------------------------------------------------
// basis e < DL: In_e[ci][p] = rev_w[(ci*204+p)*DL + e];  e == DL: rev_b[ci*204+p]
__global__ void dec_basis_kernel(const float* rev_w, const float* rev_b, int DL, int CI, double* In) {
  const int n = (DL + 1) * CI * kP;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int e = i / (CI * kP), f = i % (CI * kP);
    In[i] = (e < DL) ? static_cast<double>(rev_w[static_cast<size_t>(f) * DL + e]) : static_cast<double>(rev_b[f]);
  }
}
// G1[r][q][v] = sum_t In[r][t][v] T[v][t][q]
__global__ void dec_temporal_kernel(const double* In, const float* T, int rows, double* G1) {
  const int n = rows * kP;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int r = i / kP, p = i % kP, q = p / kV, v = p % kV;
    double s = 0.0;
    for (int t = 0; t < kT; ++t) s += In[r * kP + t * kV + v] * static_cast<double>(T[v * kT * kT + t * kT + q]);
    G1[i] = s;
  }
}
// G[r][t][w] = sum_v G1[r][t][v] A[t][v][w]
__global__ void dec_spatial_kernel(const double* G1, const float* A, int rows, double* G) {
  const int n = rows * kP;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int r = i / kP, p = i % kP, t = p / kV, w = p % kV;
    double s = 0.0;
    for (int v = 0; v < kV; ++v) s += G1[r * kP + t * kV + v] * static_cast<double>(A[t * kV * kV + v * kV + w]);
    G[i] = s;
  }
}
// M[(co*204+p)*DL + e] = sum_ci W1'[co][ci] G[e][ci][p] + W2'[co][ci] In[e][ci][p];  m0 likewise (+ b')
__global__ void dec_mix_kernel(coskad_layer_params L, const double* In, const double* G, int DL, float* M, float* m0) {
  const int CI = L.c_in, CO = L.c_out;
  const int n = (DL + 1) * CO * kP;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int e = i / (CO * kP), rem = i % (CO * kP), co = rem / kP, p = rem % kP;
    double s = 0.0;
    for (int ci = 0; ci < CI; ++ci) {
      const size_t idx = (static_cast<size_t>(e) * CI + ci) * kP + p;
      s += fold_w1(L, co, ci) * G[idx] + fold_w2(L, co, ci) * In[idx];
    }
    if (e < DL) M[static_cast<size_t>(co * kP + p) * DL + e] = static_cast<float>(s);
    else m0[co * kP + p] = static_cast<float>(s + fold_b(L, co));
  }
}

}  // namespace coskad
