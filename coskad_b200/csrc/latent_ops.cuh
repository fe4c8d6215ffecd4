// latent_ops.cuh -- standalone kernels on latents [B, D]: the gmath namespace, the scores, and the
// center update.  One warp per latent row, D <= 32*E with E elements per lane.
#pragma once
#include "common.cuh"
#include "geometry.cuh"
#include "../../include/coskad_b200.h"

namespace coskad {

constexpr int kRowWarps = 8;   // warps (rows) per CTA in the row kernels

template <int E>
__device__ __forceinline__ void load_row(const float* p, int D, int lane, float (&x)[E]) {
#pragma unroll
  for (int e = 0; e < E; ++e) { const int i = lane + 32 * e; x[e] = (i < D) ? p[i] : 0.f; }
}
template <int E>
__device__ __forceinline__ void store_row(float* p, int D, int lane, const float (&x)[E]) {
#pragma unroll
  for (int e = 0; e < E; ++e) { const int i = lane + 32 * e; if (i < D) p[i] = x[e]; }
}

// gmath.expmap0 / project / expmap0+project, hyper_math flavours, L2 normalise
template <int E>
__global__ void geom_map_kernel(int op, const float* __restrict__ in, int64_t B, int D, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t wpg = static_cast<int64_t>(gridDim.x) * kRowWarps;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * kRowWarps + (threadIdx.x >> 5); r < B; r += wpg) {
    float x[E];
    load_row<E>(in + r * D, D, lane, x);
    switch (op) {
      case COSKAD_MAP_EXPMAP0: expmap0(x, geo_geoopt(), false); break;
      case COSKAD_MAP_PROJECT: project(x, geo_geoopt()); break;
      case COSKAD_MAP_EXPMAP0_PROJECT: expmap0(x, geo_geoopt(), false); project(x, geo_geoopt()); break;
      case COSKAD_MAP_EXPMAP0_HM: expmap0(x, geo_hm(), true); break;
      case COSKAD_MAP_PROJECT_HM: project(x, geo_hm()); break;
      case COSKAD_MAP_L2NORMALIZE: {     // z / ||z||  (models/sts/vae.py:81, no eps)
        const float n = sqrtf(vec_sumsq(x));
#pragma unroll
        for (int e = 0; e < E; ++e) x[e] = x[e] / n;
      } break;
      default: break;
    }
    store_row<E>(out + r * D, D, lane, x);
  }
}

template <int E>
__global__ void dist_kernel(int flavour, const float* __restrict__ a, const float* __restrict__ b, int b_bcast,
                            int64_t B, int D, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t wpg = static_cast<int64_t>(gridDim.x) * kRowWarps;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * kRowWarps + (threadIdx.x >> 5); r < B; r += wpg) {
    float x[E], y[E];
    load_row<E>(a + r * D, D, lane, x);
    load_row<E>(b_bcast ? b : b + r * D, D, lane, y);
    float s;
    switch (flavour) {
      case COSKAD_SCORE_POINCARE:
      case COSKAD_SCORE_POINCARE_NOPROJ: s = poincare_dist(x, y, geo_geoopt()); break;
      case COSKAD_SCORE_POINCARE_HM: s = poincare_dist(x, y, geo_hm()); break;
      case COSKAD_SCORE_EUCLID: s = euclid_score(x, y, D); break;
      case COSKAD_SCORE_COSINE: s = cosine_score(x, y); break;
      default: s = 0.f; break;
    }
    if (lane == 0) out[r] = s;
  }
}

template <int E>
__global__ void dist0_kernel(const float* __restrict__ x, int64_t B, int D, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t wpg = static_cast<int64_t>(gridDim.x) * kRowWarps;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * kRowWarps + (threadIdx.x >> 5); r < B; r += wpg) {
    float v[E];
    load_row<E>(x + r * D, D, lane, v);
    const float s = 2.f * clamped_artanh(sqrtf(vec_sumsq(v)), 1e-7f);
    if (lane == 0) out[r] = s;
  }
}

// ---- backward of the element-wise gmath ops ----------------------------------------------------------------------
// The reference's own training_step differentiates through gmath.expmap0 / gmath.project / gmath.dist one call at a
// time (models/hyperbolic_encoder.py:147,157); these kernels are the analytic vector-Jacobian products of exactly the
// forward expressions above (clamps pass zero gradient outside their range, like autograd).
template <int E>
__device__ __forceinline__ void expmap0_vjp(const float (&u)[E], float (&g)[E]) {      // g: in = dL/dy, out = dL/du
  const float nraw = sqrtf(vec_sumsq(u));
  const float n = fmaxf(nraw, 1e-15f);
  const bool live = nraw > 1e-15f;
  const float t = clamped_tanh(n);
  const float f = t / n;
  const float tp = (n < 15.f) ? 1.f - t * t : 0.f;
  const float fp = live ? (tp * n - t) / (n * n) : 0.f;
  const float dot = vec_dot(g, u);
#pragma unroll
  for (int e = 0; e < E; ++e) g[e] = f * g[e] + (live ? fp * (u[e] / n) * dot : 0.f);
}
// y = x / n * R with n = max(|x|, floor); radial part of the gradient removed
template <int E>
__device__ __forceinline__ void rescale_vjp(const float (&x)[E], float (&g)[E], float R, float floor_) {
  const float nraw = sqrtf(vec_sumsq(x));
  const float n = fmaxf(nraw, floor_);
  const float dot = (nraw > floor_) ? vec_dot(g, x) : 0.f;
#pragma unroll
  for (int e = 0; e < E; ++e) g[e] = R * (g[e] / n - x[e] * dot / (n * n * n));
}
template <int E>
__device__ __forceinline__ void project_vjp(const float (&x)[E], float (&g)[E]) {
  const float n = fmaxf(sqrtf(vec_sumsq(x)), 1e-15f);
  const float maxnorm = 1.f - 4e-3f;
  if (n > maxnorm) rescale_vjp(x, g, maxnorm, 1e-15f);
}

template <int E>
__global__ void geom_map_bwd_kernel(int op, const float* __restrict__ in, const float* __restrict__ gout, int64_t B, int D,
                                    float* __restrict__ gin) {
  const int lane = threadIdx.x & 31;
  const int64_t wpg = static_cast<int64_t>(gridDim.x) * kRowWarps;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * kRowWarps + (threadIdx.x >> 5); r < B; r += wpg) {
    float x[E], g[E];
    load_row<E>(in + r * D, D, lane, x);
    load_row<E>(gout + r * D, D, lane, g);
    switch (op) {
      case COSKAD_MAP_EXPMAP0: expmap0_vjp(x, g); break;
      case COSKAD_MAP_PROJECT: project_vjp(x, g); break;
      case COSKAD_MAP_EXPMAP0_PROJECT: {
        float e_[E];
#pragma unroll
        for (int e = 0; e < E; ++e) e_[e] = x[e];
        expmap0(e_, geo_geoopt(), false);
        project_vjp(e_, g);
        expmap0_vjp(x, g);
      } break;
      case COSKAD_MAP_L2NORMALIZE: rescale_vjp(x, g, 1.f, 0.f); break;
      default: break;
    }
    store_row<E>(gin + r * D, D, lane, g);
  }
}

// ga[B,D] (nullable) = gs * d f(a,b)/da, gb[B,D] (nullable) = gs * d f(a,b)/db per row (a broadcast b gets per-row rows too)
template <int E>
__global__ void dist_bwd_kernel(int flavour, const float* __restrict__ a, const float* __restrict__ b, int b_bcast,
                                const float* __restrict__ gs, int64_t B, int D, float* __restrict__ ga,
                                float* __restrict__ gb) {
  const int lane = threadIdx.x & 31;
  const int64_t wpg = static_cast<int64_t>(gridDim.x) * kRowWarps;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * kRowWarps + (threadIdx.x >> 5); r < B; r += wpg) {
    float x[E], y[E], dx[E], dy[E];
    load_row<E>(a + r * D, D, lane, x);
    load_row<E>(b_bcast ? b : b + r * D, D, lane, y);
    const float g = gs[r];
    if (flavour == COSKAD_SCORE_POINCARE || flavour == COSKAD_SCORE_POINCARE_NOPROJ) {
#pragma unroll
      for (int e = 0; e < E; ++e) x[e] = -x[e];                   // mobius_add(-a, b)
      const float x2 = vec_sumsq(x), y2 = vec_sumsq(y), xy = vec_dot(x, y);
      const float ca = 1.f + 2.f * xy + y2, cb = 1.f - x2;
      const float den_raw = 1.f + 2.f * xy + x2 * y2;
      const float den = fmaxf(den_raw, 1e-15f);
      float rr[E], dnum[E];
#pragma unroll
      for (int e = 0; e < E; ++e) rr[e] = (ca * x[e] + cb * y[e]) / den;
      const float rn = sqrtf(vec_sumsq(rr));
      const float hi = 1.f - 1e-7f;
      const float d_rn = (rn < hi && rn > -hi) ? g * 2.f / (1.f - rn * rn) : 0.f;
      float s_rr = 0.f;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const float d_r = (rn > 0.f) ? d_rn * rr[e] / rn : 0.f;
        dnum[e] = d_r / den;
        s_rr = fmaf(d_r, rr[e], s_rr);
      }
      const float d_den = (den_raw > 1e-15f) ? -warp_sum(s_rr) / den : 0.f;
      const float d_ca = vec_dot(dnum, x), d_cb = vec_dot(dnum, y);
      const float d_xy = 2.f * d_ca + 2.f * d_den;
      const float d_y2 = d_ca + x2 * d_den;
      const float d_x2 = -d_cb + y2 * d_den;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        dx[e] = -(ca * dnum[e] + d_xy * y[e] + 2.f * d_x2 * x[e]);   // d/da = -d/dx
        dy[e] = cb * dnum[e] + d_xy * x[e] + 2.f * d_y2 * y[e];
      }
    } else if (flavour == COSKAD_SCORE_EUCLID) {
#pragma unroll
      for (int e = 0; e < E; ++e) { const float d = 2.f * (y[e] - x[e]) / static_cast<float>(D) * g; dx[e] = -d; dy[e] = d; }
    } else {   // COSINE: 1 - <b/|b|, a/|a|>, norms clamped at 1e-8
      const float nar = sqrtf(vec_sumsq(x)), nbr = sqrtf(vec_sumsq(y));
      const float na = fmaxf(nar, 1e-8f), nb = fmaxf(nbr, 1e-8f);
      float ah[E], bh[E];
#pragma unroll
      for (int e = 0; e < E; ++e) { ah[e] = x[e] / na; bh[e] = y[e] / nb; }
      const float c = vec_dot(ah, bh);
#pragma unroll
      for (int e = 0; e < E; ++e) {
        dx[e] = -g * (bh[e] - (nar > 1e-8f ? c * ah[e] : 0.f)) / na;
        dy[e] = -g * (ah[e] - (nbr > 1e-8f ? c * bh[e] : 0.f)) / nb;
      }
    }
    if (ga) store_row<E>(ga + r * D, D, lane, dx);
    if (gb) store_row<E>(gb + r * D, D, lane, dy);
  }
}

// PowerSpherical reparameterised sample from explicit noise (models/sts/vae.py:110,129; power_spherical's
// _TTransform + _HouseholderRotationTransform): y = [t, sqrt(clamp(1-t^2,1e-7)) v]; u = (e1 - mu)/(|e1 - mu| + 1e-5);
// z = y - 2 <y,u> u.   mu [B,d], t [B], v [B,d-1] (unit vectors), d <= 32, one warp per row.
__global__ void ps_sample_kernel(const float* __restrict__ mu, const float* __restrict__ t, const float* __restrict__ v,
                                 int64_t B, int d, float* __restrict__ z) {
  const int lane = threadIdx.x & 31;
  const int64_t wpg = static_cast<int64_t>(gridDim.x) * kRowWarps;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * kRowWarps + (threadIdx.x >> 5); r < B; r += wpg) {
    const float tt = t[r];
    float y = 0.f, u = 0.f;
    if (lane < d) {
      y = (lane == 0) ? tt : v[r * (d - 1) + lane - 1] * sqrtf(fmaxf(1.f - tt * tt, 1e-7f));
      u = ((lane == 0) ? 1.f : 0.f) - mu[r * d + lane];
    }
    const float un = sqrtf(warp_sum(u * u)) + 1e-5f;
    u = u / un;
    const float dot = warp_sum(y * u);
    if (lane < d) z[r * d + lane] = y - 2.f * dot * u;
  }
}

// ---- Mahalanobis distance to the center (distance: 'mahalanobis') -----------------------------------------------
// utils/eval_utils.py:28-38 mahalanobis(u, v, VI): sqrt((u - v)^T VI (u - v)) per row, evaluated like the reference's two
// matmuls: t = (u - v)^T VI first, then t (u - v).  One warp per row, lane = component (D <= 32); VI staged in shared memory.
constexpr int kMahD = 32;
__global__ void mahalanobis_kernel(const float* __restrict__ z, const float* __restrict__ c, const float* __restrict__ VI,
                                   int64_t B, int D, float* __restrict__ out) {
  __shared__ float vi[kMahD][kMahD + 1];
  for (int i = threadIdx.x; i < D * D; i += blockDim.x) vi[i / D][i % D] = VI[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const float cl = lane < D ? c[lane] : 0.f;
  const int64_t wpg = static_cast<int64_t>(gridDim.x) * kRowWarps;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * kRowWarps + (threadIdx.x >> 5); r < B; r += wpg) {
    const float d = lane < D ? z[r * D + lane] - cl : 0.f;
    float t = 0.f;
    for (int i = 0; i < D; ++i) t = fmaf(__shfl_sync(0xffffffffu, d, i), lane < D ? vi[i][lane] : 0.f, t);
    const float q = warp_sum(t * d);
    if (lane == 0) out[r] = sqrtf(q);
  }
}
// gz[r, :] = gs[r] * (VI + VI^T)(z_r - c) / (2 sqrt(q_r))   (autograd of the expression above w.r.t. u; q = 0: zero gradient)
__global__ void mahalanobis_bwd_kernel(const float* __restrict__ z, const float* __restrict__ c, const float* __restrict__ VI,
                                       const float* __restrict__ gs, int64_t B, int D, float* __restrict__ gz) {
  __shared__ float vi[kMahD][kMahD + 1];
  for (int i = threadIdx.x; i < D * D; i += blockDim.x) vi[i / D][i % D] = VI[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const float cl = lane < D ? c[lane] : 0.f;
  const int64_t wpg = static_cast<int64_t>(gridDim.x) * kRowWarps;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * kRowWarps + (threadIdx.x >> 5); r < B; r += wpg) {
    const float d = lane < D ? z[r * D + lane] - cl : 0.f;
    float t = 0.f, s = 0.f;                      // t = (d^T VI)[lane], s = (VI d)[lane]
    for (int i = 0; i < D; ++i) {
      const float di = __shfl_sync(0xffffffffu, d, i);
      if (lane < D) { t = fmaf(di, vi[i][lane], t); s = fmaf(di, vi[lane][i], s); }
    }
    const float q = warp_sum(t * d);
    const float f = q > 0.f ? gs[r] / (2.f * sqrtf(q)) : 0.f;
    if (lane < D) gz[r * D + lane] = f * (t + s);
  }
}
// sum over rows of (x - mu)(x - mu)^T (models/euclidean_encoder_staticCenter.py:40-46 batch_cov_mat_step), shard-additive:
// one float64 partial [D*D + 1] (the last entry counts the rows) per CTA, added in a fixed order by cov_partial_final_kernel.
// Lane i of a row's warp owns row i of the outer product; the products are float32 like the reference's matmul.
template <int DMAX>
__global__ void cov_partial_kernel(const float* __restrict__ z, const float* __restrict__ mu, int64_t B, int D,
                                   double* __restrict__ part) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float ml = lane < D ? mu[lane] : 0.f;
  double acc[DMAX];
#pragma unroll
  for (int j = 0; j < DMAX; ++j) acc[j] = 0.0;
  double cnt = 0.0;
  const int64_t wpg = static_cast<int64_t>(gridDim.x) * kRowWarps;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * kRowWarps + warp; r < B; r += wpg) {
    const float d = lane < D ? z[r * D + lane] - ml : 0.f;
#pragma unroll
    for (int j = 0; j < DMAX; ++j) acc[j] += static_cast<double>(d * __shfl_sync(0xffffffffu, d, j));
    cnt += 1.0;
  }
  // CTA reduction in warp order (fixed: the partial is bit-reproducible); one [DMAX][33] buffer, the warps add in turn
  __shared__ double red[DMAX][33];
  __shared__ double rc;
  for (int w = 0; w < kRowWarps; ++w) {
    if (warp == w) {
#pragma unroll
      for (int j = 0; j < DMAX; ++j) red[j][lane] = (w == 0 ? 0.0 : red[j][lane]) + acc[j];
      if (lane == 0) rc = (w == 0 ? 0.0 : rc) + cnt;
    }
    __syncthreads();
  }
  double* out = part + static_cast<int64_t>(blockIdx.x) * (D * D + 1);
  for (int e = threadIdx.x; e < D * D; e += blockDim.x) out[e] = red[e % D][e / D];
  if (threadIdx.x == 0) out[D * D] = rc;
}
__global__ void cov_partial_final_kernel(const double* __restrict__ part, int nblk, int n, double* acc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int b = 0; b < nblk; ++b) s += part[static_cast<int64_t>(b) * n + i];
  acc[i] += s;
}

// ---- center partial sums ----------------------------------------------------------------------
// POINCARE (gmath.weighted_midpoint, weights=None): gamma_i = lambda_x(x_i) = 2 / max(1 - |x_i|^2, 1e-15)
// evaluated in float32 like the reference; the sums over windows run in float64.
template <int E>
__global__ void center_partial_kernel(int flavour, const float* __restrict__ z, int64_t B, int D, double* __restrict__ part) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double sx[E], sg = 0.0, sn = 0.0;
#pragma unroll
  for (int e = 0; e < E; ++e) sx[e] = 0.0;
  const int64_t wpg = static_cast<int64_t>(gridDim.x) * kRowWarps;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * kRowWarps + warp; r < B; r += wpg) {
    float x[E];
    load_row<E>(z + r * D, D, lane, x);
    if (flavour == COSKAD_SCORE_POINCARE || flavour == COSKAD_SCORE_POINCARE_NOPROJ) {
      const float ss = vec_sumsq(x);
      const float gamma = 2.f / fmaxf(1.f - ss, 1e-15f);
#pragma unroll
      for (int e = 0; e < E; ++e) sx[e] += static_cast<double>(gamma * x[e]);
      sg += static_cast<double>(gamma - 1.f);
    } else {
#pragma unroll
      for (int e = 0; e < E; ++e) sx[e] += static_cast<double>(x[e]);
    }
    sn += 1.0;
  }
  // CTA reduce (per lane slot) then one partial [D + 2] per CTA (center_partial_final_kernel adds them in a fixed order)
  double* acc = part + static_cast<int64_t>(blockIdx.x) * (D + 2);
  __shared__ double red[kRowWarps][32 * E + 2];
#pragma unroll
  for (int e = 0; e < E; ++e) red[warp][lane + 32 * e] = sx[e];
  if (lane == 0) { red[warp][32 * E] = sg; red[warp][32 * E + 1] = sn; }
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * E + 2; i += blockDim.x) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kRowWarps; ++w) s += red[w][i];
    if (i < 32 * E) { if (i < D) acc[i] = s; }
    else acc[D + (i - 32 * E)] = s;
  }
}
// second stage, fixed order: 32 x 8 threads, row y adds the partials y, y + 8, .. (batches of 8 loads in flight: one thread
// per element walked up to 296 partials one L2 round trip at a time, 12 us), the rows meet in shared memory
__global__ void center_partial_final_kernel(const double* __restrict__ part, int nblk, int D, double* acc) {
  __shared__ double sh[8][33];
  const int n = D + 2;
  for (int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + threadIdx.x;
    double s = 0.0;
    if (i < n) {
      int b = threadIdx.y;
      for (; b + 7 * 8 < nblk; b += 8 * 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = part[static_cast<int64_t>(b + u * 8) * n + i];
#pragma unroll
        for (int u = 0; u < 8; ++u) s += v[u];
      }
      for (; b < nblk; b += 8) s += part[static_cast<int64_t>(b) * n + i];
    }
    sh[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && i < n) {
      double t = 0.0;
#pragma unroll
      for (int y = 0; y < 8; ++y) t += sh[y][threadIdx.x];
      acc[i] += t;
    }
    __syncthreads();
  }
}

// single warp
template <int E>
__global__ void center_finalize_kernel(int flavour, const double* __restrict__ acc, int D, float eps, float* center) {
  const int lane = threadIdx.x & 31;
  const double cnt = acc[D + 1];
  float m[E];
  if (flavour == COSKAD_SCORE_POINCARE || flavour == COSKAD_SCORE_POINCARE_NOPROJ) {
    // two_mean = num / clamp_abs(den, 1e-10); mobius_scalar_mul(0.5, two_mean)
    const float den = static_cast<float>(acc[D]);
    const float dena = (den >= 0.f ? 1.f : -1.f) * (fabsf(den) + 1e-10f);
#pragma unroll
    for (int e = 0; e < E; ++e) { const int i = lane + 32 * e; m[e] = (i < D) ? static_cast<float>(acc[i]) / dena : 0.f; }
    const float n = fmaxf(sqrtf(vec_sumsq(m)), 1e-15f);
    const float th = clamped_tanh(0.5f * clamped_artanh(n, 1e-7f));
#pragma unroll
    for (int e = 0; e < E; ++e) m[e] = th * (m[e] / n);
  } else {
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int i = lane + 32 * e;
      float c = (i < D && cnt > 0.0) ? static_cast<float>(acc[i] / cnt) : 0.f;
      if (flavour == COSKAD_SCORE_EUCLID && eps > 0.f) {   // euclidean_encoder_staticCenter.py:121-122
        if (fabsf(c) < eps && c < 0.f) c = -eps;
        if (fabsf(c) < eps && c > 0.f) c = eps;
      }
      m[e] = c;
    }
  }
  store_row<E>(center, D, lane, m);
}

}  // namespace coskad
