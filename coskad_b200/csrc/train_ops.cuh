// train_ops.cuh -- kernels of the training path.
#pragma once
#include "common.cuh"
#include "geometry.cuh"
#include "latent_ops.cuh"

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
