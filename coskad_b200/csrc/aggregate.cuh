// aggregate.cuh -- frame-level score aggregation, bit-exact with the reference's numpy path.
//
// Reference (utils/eval_utils.py:69-74, eval_COSKAD.py:201-211), per person:
//     pose = np.zeros((w, n_frames))                       # float64
//     for n in range(w): pose[n, frames[n] - 1] = loss[n]  # frame id 0 wraps to the last column
//     pose = where(pose == 0.0, nan, pose); fig = nanmean(pose, 0); fig = where(isnan(fig), 0, fig)
//   clip_score = amax over the clip's persons.
// nanmean(axis=0) of a C-contiguous [w, n_frames] float64 array adds the rows one after another
// (SURVEY.md fact 9), so  fig[f] = (sum over windows n in dataset order with f in frames[n]-1 and
// loss[n] != 0 of (double)loss[n]) / count, computed here by one warp per person that walks its
// windows sequentially (lanes = the T frame slots of a window) -- deterministic, no atomics.
#pragma once
#include "common.cuh"

namespace coskad {

constexpr int kAggWarps = 4;

__global__ void person_curves_kernel(const float* __restrict__ score, const int64_t* __restrict__ frames, int T,
                                     const int64_t* __restrict__ win_idx, const int64_t* __restrict__ person_off,
                                     const int32_t* __restrict__ person_clip, int64_t n_persons,
                                     const int64_t* __restrict__ clip_off, double* sum, int32_t* cnt,
                                     const int64_t* __restrict__ person_out_off) {
  const int lane = threadIdx.x & 31;
  const int64_t person = static_cast<int64_t>(blockIdx.x) * kAggWarps + (threadIdx.x >> 5);
  if (person >= n_persons) return;
  const int32_t clip = person_clip[person];
  const int64_t F = clip_off[clip + 1] - clip_off[clip];
  double* s = sum + person_out_off[person];
  int32_t* c = cnt + person_out_off[person];
  for (int64_t f = lane; f < F; f += 32) { s[f] = 0.0; c[f] = 0; }
  __syncwarp();
  for (int64_t k = person_off[person]; k < person_off[person + 1]; ++k) {
    const int64_t n = win_idx[k];
    const float v = score[n];
    // T <= 32 frame slots, one per lane
    int64_t col = -1;
    if (lane < T) {
      col = frames[n * T + lane] - 1;
      if (col < 0) col += F;                 // numpy negative index: frame id 0 -> last frame
      if (col < 0 || col >= F) col = -1;     // out of range ids raise upstream; ignored here
    }
    // a frame id repeated inside one window is one assignment, not two
    bool dup = false;
    for (int j = 0; j < T; ++j) {
      const int64_t cj = __shfl_sync(0xffffffffu, col, j);
      if (j < lane && cj == col) dup = true;
    }
    if (col >= 0 && !dup && v != 0.0f) {     // exact zero == "absent" (eval_COSKAD.py:201)
      s[col] += static_cast<double>(v);
      c[col] += 1;
    }
    __syncwarp();
  }
  for (int64_t f = lane; f < F; f += 32) s[f] = (c[f] > 0) ? s[f] / static_cast<double>(c[f]) : 0.0;
}

// clip_score[f] = max over the clip's persons (np.amax, eval_COSKAD.py:211); persons of a clip are contiguous
__global__ void clip_max_kernel(const double* __restrict__ person_curve, const int64_t* __restrict__ person_out_off,
                                const int64_t* __restrict__ clip_person_off, const int64_t* __restrict__ clip_off,
                                int64_t n_clips, double* __restrict__ out) {
  const int64_t clip = blockIdx.y;
  if (clip >= n_clips) return;
  const int64_t F = clip_off[clip + 1] - clip_off[clip];
  const int64_t p0 = clip_person_off[clip], p1 = clip_person_off[clip + 1];
  for (int64_t f = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; f < F;
       f += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    double m = 0.0;
    for (int64_t p = p0; p < p1; ++p) {
      const double v = person_curve[person_out_off[p] + f];
      m = (p == p0) ? v : (v > m ? v : m);
    }
    out[clip_off[clip] + f] = m;
  }
}


// ---- score post-processing: shift by `shift` frames + Gaussian smoothing, per clip curve ----------------------------
// Reference: utils/eval_utils.py:200-207 score_process: scores_shifted[shift:] = score[:-shift] (shift = 11), then
// scipy.ndimage.gaussian_filter1d(scores_shifted, 30) = correlate1d with the normalised 241-tap kernel, mode 'reflect'
// (half-sample symmetric extension: d c b a | a b c d | d c b a).  The accumulation order is the one of scipy's
// NI_Correlate1D symmetric-kernel branch: tmp = x[l] w[0]; for jj = -r .. -1: tmp += (x[l+jj] + x[l-jj]) * w[jj] --
// separate multiply and add (no FMA contraction), float64.  weights = the kernel as numpy computes it (host), w[r] = centre.
__device__ __forceinline__ double shifted_reflect(const double* __restrict__ c, int64_t F, int64_t i, int shift) {
  const int64_t period = 2 * F;
  i %= period;
  if (i < 0) i += period;
  if (i >= F) i = period - 1 - i;
  return (i >= shift) ? c[i - shift] : 0.0;
}
__global__ void score_process_kernel(const double* __restrict__ curves, const int64_t* __restrict__ curve_off, int64_t n_curves,
                                     int shift, const double* __restrict__ weights, int radius, double* __restrict__ out) {
  const int64_t cv = blockIdx.x;
  if (cv >= n_curves) return;
  const int64_t o = curve_off[cv], F = curve_off[cv + 1] - o;
  const double* c = curves + o;
  for (int64_t l = threadIdx.x; l < F; l += blockDim.x) {
    double tmp = __dmul_rn(shifted_reflect(c, F, l, shift), weights[radius]);
    for (int jj = -radius; jj < 0; ++jj) {
      const double pair = __dadd_rn(shifted_reflect(c, F, l + jj, shift), shifted_reflect(c, F, l - jj, shift));
      tmp = __dadd_rn(tmp, __dmul_rn(pair, weights[radius + jj]));
    }
    out[o + l] = tmp;
  }
}

}  // namespace coskad
