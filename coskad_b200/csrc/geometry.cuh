// geometry.cuh -- latent-geometry device math (warp-shuffle reductions).
//
// A latent vector of D <= 32*E floats is spread over the 32 lanes of a warp, E elements per
// lane (element i lives in lane i%32, slot i/32; absent elements are 0).  Every function is
// called by all 32 lanes.
//
// POINCARE flavour = geoopt==0.5.0 geoopt/manifolds/stereographic/math.py with k = -1
// (third-party, pinned by the reference's environment.yml:247; call sites
// models/hyperbolic_encoder.py:110,122,147,157,179,181,266, utils/eval_utils.py:67,
// eval_COSKAD.py:195).  With k = -1: sabs(k)^0.5 == 1.0f exactly, so tan_k == tanh and
// artan_k == artanh.
// POINCARE_HM flavour = utils/hyper_math.py (c = +1): expmap0 :302-306, project :100-105,
// mobius_add :173-179, dist :207-210.
#pragma once
#include "common.cuh"

namespace coskad {

struct GeoConst {
  float min_norm;     // norm clamp
  float proj_eps;     // project: maxnorm = 1 - proj_eps
  float atanh_eps;    // artanh clamp +-(1 - atanh_eps)
  float den_add;      // hyper_math adds 1e-5 to the Mobius denominator ...
  float den_min;      // ... geoopt clamps it at 1e-15
};
__device__ __forceinline__ GeoConst geo_geoopt() { return GeoConst{1e-15f, 4e-3f, 1e-7f, 0.f, 1e-15f}; }
__device__ __forceinline__ GeoConst geo_hm() { return GeoConst{1e-5f, 1e-3f, 1e-5f, 1e-5f, -3.0e38f}; }

template <int E>
__device__ __forceinline__ float vec_sumsq(const float (&x)[E]) {
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < E; ++e) s = fmaf(x[e], x[e], s);
  return warp_sum(s);
}
template <int E>
__device__ __forceinline__ float vec_dot(const float (&x)[E], const float (&y)[E]) {
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < E; ++e) s = fmaf(x[e], y[e], s);
  return warp_sum(s);
}

__device__ __forceinline__ float clamped_tanh(float x) { return tanhf(fminf(fmaxf(x, -15.f), 15.f)); }
__device__ __forceinline__ float clamped_artanh(float x, float eps) {
  const float hi = 1.f - eps;
  x = fminf(fmaxf(x, -hi), hi);
  return 0.5f * (logf(1.f + x) - logf(1.f - x));
}

// expmap0(u): geoopt tan_k(|u|) * (u/|u|);  hyper_math tanh(|u|) * u / |u|
template <int E>
__device__ __forceinline__ void expmap0(float (&u)[E], const GeoConst& g, bool hm) {
  const float n = fmaxf(sqrtf(vec_sumsq(u)), g.min_norm);
  const float th = clamped_tanh(n);
#pragma unroll
  for (int e = 0; e < E; ++e) u[e] = hm ? (th * u[e]) / n : th * (u[e] / n);
}
// project(x): clip to the ball of radius 1 - eps
template <int E>
__device__ __forceinline__ void project(float (&x)[E], const GeoConst& g) {
  const float n = fmaxf(sqrtf(vec_sumsq(x)), g.min_norm);
  const float maxnorm = 1.f - g.proj_eps;
  if (n > maxnorm) {
#pragma unroll
    for (int e = 0; e < E; ++e) x[e] = x[e] / n * maxnorm;
  }
}
// dist(a, b) = 2 artanh(|(-a) (+) b|)
template <int E>
__device__ __forceinline__ float poincare_dist(const float (&a)[E], const float (&b)[E], const GeoConst& g) {
  float x[E];
#pragma unroll
  for (int e = 0; e < E; ++e) x[e] = -a[e];
  const float x2 = vec_sumsq(x), y2 = vec_sumsq(b), xy = vec_dot(x, b);
  const float ca = 1.f + 2.f * xy + y2;     // 1 - 2k xy - k y2, k = -1
  const float cb = 1.f - x2;                // 1 + k x2
  float den = 1.f + 2.f * xy + x2 * y2;     // 1 - 2k xy + k^2 x2 y2
  den = fmaxf(den + g.den_add, g.den_min);
  float r[E];
#pragma unroll
  for (int e = 0; e < E; ++e) r[e] = (ca * x[e] + cb * b[e]) / den;
  const float rn = sqrtf(vec_sumsq(r));
  return 2.f * clamped_artanh(rn, g.atanh_eps);
}
template <int E>
__device__ __forceinline__ float euclid_score(const float (&z)[E], const float (&c)[E], int D) {
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < E; ++e) { const float d = c[e] - z[e]; s = fmaf(d, d, s); }
  return warp_sum(s) / static_cast<float>(D);
}
// 1 - F.cosine_similarity(c, z) (torch >= 2.0: each vector divided by max(norm, 1e-8) first)
template <int E>
__device__ __forceinline__ float cosine_score(const float (&z)[E], const float (&c)[E]) {
  const float nz = fmaxf(sqrtf(vec_sumsq(z)), 1e-8f), nc = fmaxf(sqrtf(vec_sumsq(c)), 1e-8f);
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < E; ++e) s = fmaf(c[e] / nc, z[e] / nz, s);
  return 1.f - warp_sum(s);
}

// Full per-window score from the raw head output u (modified in place to the projected latent
// for the Poincare flavours).  flavour is warp-uniform.
template <int E>
__device__ __forceinline__ float score_from_latent(int flavour, float (&u)[E], const float (&c)[E], int D) {
  switch (flavour) {
    case 1: { const GeoConst g = geo_geoopt(); expmap0(u, g, false); project(u, g); return poincare_dist(u, c, g); }
    case 2: { const GeoConst g = geo_geoopt(); expmap0(u, g, false); return poincare_dist(u, c, g); }
    case 3: return euclid_score(u, c, D);
    case 4: return cosine_score(u, c);
    case 5: { const GeoConst g = geo_hm(); expmap0(u, g, true); project(u, g); return poincare_dist(u, c, g); }
    default: return 0.f;
  }
}

}  // namespace coskad
