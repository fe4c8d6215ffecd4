// fused_eval_tc.cuh -- fused eval hot path, tensor-core generation: the dense channel mixing (65 % of the MACs) runs on the
// 5th-gen tensor cores (tcgen05.mma kind::tf32, 3xTF32 split accumulate), the graph contractions, the linear head and
// the geometry stay on the CUDA cores (packed FFMA2).  Same arithmetic contract as fused_eval.cuh (see there for the
// reference citations); score parity rtol 1e-4 is kept by the hi/lo operand split (tc.cuh).  DESIGN.md section 4 describes
// the structure; in short, per CTA (one per SM, 384 threads = 12 warps = 3 groups x 4 TMEM lane quarters, 227 KB smem,
// all 512 TMEM columns), per tile of kNW = 3 windows:
//   * activations live in shared memory as channel planes [row = window*C + c][205]; H1 (layer-1 output) and H4
//     (layer-4 output) never exist in memory: H1 is split straight into the layer-2 TMEM operand, H4 is consumed by
//     the head from the TMEM accumulators;
//   * layers 2/3 (N = 32): each warp group owns one window, stages its two M-tiles (128 positions each) into private
//     TMEM A buffers, issues its own MMAs and runs its own epilogue (SmallPipe);
//   * layer 4 (N = 64): 384 D columns + two shared 64-column A buffers handed from tile to tile through per-tile mbarriers
//     (TcPipe), staged by all three warp groups; the residual
//     half is issued before the layer-4 graph contraction and executes behind it; the head starts on the j = 0 tiles
//     while the j = 1 tiles still execute;
//   * tile software pipelining: head reduction + score of tile i run in the first stage of tile i+1; the layer-1 graph
//     contraction of tile i+1 runs on the quarter-3 warps during the head of tile i;
//   * optional front end: windows gathered from trajectory rows + test-time affine transform in the input stage;
//   * kDec = true appends the auto-encoder's decoder (folded first layer, layers 1-2 mixing on the tensor cores).
#pragma once
#include <type_traits>
#include "common.cuh"
#include "fused_eval.cuh"
#include "geometry.cuh"
#include "tc.cuh"

namespace coskad {

constexpr int kTcThreads = 384;
constexpr int kTcWarps = kTcThreads / 32;        // 12 = 4 TMEM lane quarters x 3 groups
constexpr int kTcTiles = 2 * kNW;                // M-tiles per CTA tile: (half j, window n) -> ti = j*kNW + n
constexpr uint32_t kTcColD = 0;                  // D: 64 columns per M-tile
constexpr uint32_t kTcColA = 64 * kTcTiles;      // A: 2 buffers x (32 hi + 32 lo) columns
static_assert(kTcColA + 128 <= 512, "TMEM budget");

// blob = [B_hi image Kp*N][B_lo image Kp*N][bias N][slope,0,0,0]
__host__ __device__ constexpr int tc_blob_floats(int Kp, int N) { return 2 * Kp * N + N + 4; }
constexpr int kTcWsFloats = tc_blob_floats(32, 32);   // small buffer: L1 (Kp 8), L3
constexpr int kTcWbFloats = tc_blob_floats(32, 64);   // big buffer: L2, L4 residual half, L4 graph half

struct FusedTcParams {
  const float* eTw[4];
  const float* eAw[4];
  const float* mixL1;     // FP32 blob of layer 1 (fused_eval.cuh mix layout: [4][32] + bias + slope): K = 4 is too thin for an MMA
  const float* tcL2;      // mix-first: K 32 (H1), N 32 = [U 16 | Rsd 16]
  const float* tcL3;      // K 32 = [G3 16 | H2 16], N 32
  const float* tcL4X;     // K 32 (H3, residual conv), N 64
  const float* tcL4G;     // K 32 (G4, tcn conv), N 64, carries bias + slope
  const float* head_w;    // [16][kF]
  const float* head_w4;   // [64][4][204][4]: (c, dq, p, i) = head_w[4 dq + i][c*204 + p] (fold.cuh pack_head4_kernel)
  const float* head_b;    // [16]
  const float* x;
  const float* center;
  float* z;
  float* score;
  int64_t B;
  int head_rows, D, flavour;
  // optional front end (coskad_encode_score_traj_fwd): windows gathered from trajectory rows + test-time affine transform
  const float* traj;          // [traj_rows, 2*kV] (x0,y0,x1,y1,.. per frame), nullptr = windows come from x
  const int64_t* win_row;     // [B] first trajectory row of window i
  int64_t traj_rows;
  const int* trans;           // [B] index into mats, nullptr = no transform
  const float* mats;          // [n_mats][6]: x' = m0 x + m1 y + m2, y' = m3 x + m4 y + m5
  int n_mats;
  // decoder (fused_eval_tc_kernel<true>, auto-encoder): same blobs as FusedParams (fused_eval.cuh)
  const float* dM;            // [32*204][DL] folded rev_btlnk + first decoder layer
  const float* dm0;           // [32*204]
  float d_slope0;
  int DL;
  const float* dTw[3];
  const float* dAw[3];
  const float* dWm[3];
  const float* tcD2;          // tensor-core blobs of decoder layers 1 (32->16 mix-first) and 2 (16->32)
  const float* tcD3;
  float* xhat;                // [B,2,12,17] or null
  float* rec_score;           // [B] or null
};

constexpr int kTcGB = (kRSmall + 3) & ~3;                             // keep what follows 16-byte aligned (cp.async 16, UMMA descriptors)
static_assert((2 * kRBig + 2 * kRSmall + kTcGB) % 4 == 0, "weight staging buffers must be 16-byte aligned");
constexpr int kTcSmemFloats = 2 * kRBig + 2 * kRSmall + kTcGB        // R0, R1, XB[2], GB
                              + kTwFloats + kAwFloats + kTcWsFloats + kTcWbFloats
                              + 2 * kTcWarps * kNW * kDP + kNW * kDP + 32   // zpart[2], zfin, center
                              + 40;                                     // mbarriers (18 x 8 B) + tmem base + task counter
constexpr int kTcSmemBytes = kTcSmemFloats * 4;
static_assert(kTcSmemBytes <= 227 * 1024, "shared memory plan exceeds 227 KB");

struct TcPipe {
  uint64_t* tdone;    // [6] MMAs of tile ti complete (1 arrival: tcgen05.commit of the group that issued it)
  uint64_t* done;     // all MMAs of a phase complete (3 arrivals: one commit per warp group)
  uint64_t* half;     // optional: the first kNW tiles (position half j = 0) of the phase complete (3 arrivals); nullptr = not signalled
  uint32_t n_half;
  uint32_t tbase;
  uint32_t n_phase, n_done;   // completed layer-4 phases / done waits (identical on every thread)
};

// One layer-4 mixing phase over the 6 M-tiles: D[ti] (+)= [src1 | src2](positions of ti, K channels) * B^T  (N = 64: the
// 384 D columns leave room for two shared 64-column A buffers).  All 12 warps produce: warp group g (4 warps = the 4 TMEM lane
// quarters) stages tiles g (j = 0) and g + 3 (j = 1) of its window into A buffer ti & 1, syncs its 4 warps on a named barrier,
// and its first lane issues the tile's MMAs and commits them to the tile's own mbarrier tdone[ti].  A buffer is handed from
// tile ti - 2 to tile ti (across groups): the producers of tile ti wait on tdone[ti - 2] of this phase (tiles 0 and 1: on
// tdone[4] / tdone[5] of the previous phase).  One mbarrier per tile, used once per phase, keeps the parity waits unambiguous
// whatever the lag between the groups (a shared per-buffer barrier would let a group that skipped a generation pass early).
// The chains 4 <- 2 <- 0 and 5 <- 3 <- 1 are acyclic: every group issues its first tile before it waits for a later one.
// K1 + K2 <= 32; Kp = K rounded up to 8 (the extra columns are staged as zeros).
struct NoFiller { __device__ __forceinline__ void operator()() const {} };
// `filler` runs on every warp between its two tiles: useful work while the other groups' MMAs free the next A buffer
template <int K1, int K2, int N, class Filler = NoFiller>
__device__ __forceinline__ void tc_mix_phase(TcPipe& P, const float* src1, const float* src2, const float* Bhi, const float* Blo,
                                             bool accumulate, int warp, int lane, Filler filler = Filler{}) {
  constexpr int K = K1 + K2;
  constexpr int Kp = (K + 7) & ~7;
  static_assert(Kp <= 32 && N % 16 == 0 && N <= 64, "bad mixing phase shape");
  const int q = warp & 3, g = warp >> 2;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int ti = g + kNW * r;                        // = j * kNW + n with n = g, j = r
    const int b = ti & 1;
    const int n = g;
    const int p = r * 128 + q * 32 + lane;
    const bool valid = p < kP;
    const int pc = valid ? p : kP - 1;
    const uint32_t abuf = P.tbase + (static_cast<uint32_t>(q * 32) << 16) + kTcColA + 64u * b;
    if (ti >= 2) tc::mbar_wait(&P.tdone[ti - 2], P.n_phase & 1);                       // previous user, this phase
    else if (P.n_phase > 0) tc::mbar_wait(&P.tdone[ti + 4], (P.n_phase - 1) & 1);        // last user of the previous phase
    tc::fence_after_sync();
#pragma unroll
    for (int k0 = 0; k0 < Kp; k0 += 16) {
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int k = k0 + u;
        float a = 0.f;
        if (k < K1) a = src1[(n * K1 + k) * kCS + pc];
        else if (k < K) a = src2[(n * K2 + (k - K1)) * kCS + pc];
        if (!valid) a = 0.f;
        tc::split_tf32(a, hi[u], lo[u]);
      }
      tc::tmem_st16(abuf + k0, hi);
      tc::tmem_st16(abuf + 32 + k0, lo);
    }
    tc::wait_st();
    tc::fence_before_sync();
    asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory");
    if (q == 0 && lane == 0) {
      tc::fence_after_sync();
      const uint32_t idesc = tc::make_idesc_tf32(128, N);
      constexpr uint32_t lbo = (N / 8) * 128, sbo = 128;
      const uint32_t bh = tc::smem_u32(Bhi), bl = tc::smem_u32(Blo);
      const uint32_t d = P.tbase + kTcColD + 64u * ti;
      const uint32_t a = P.tbase + kTcColA + 64u * b;
#pragma unroll
      for (int kb = 0; kb < Kp / 8; ++kb) {
        const uint64_t dh = tc::make_smem_desc(bh + kb * 2 * lbo, lbo, sbo);
        const uint64_t dl = tc::make_smem_desc(bl + kb * 2 * lbo, lbo, sbo);
        tc::mma_tf32_ts(d, a + kb * 8, dh, idesc, (accumulate || kb > 0) ? 1u : 0u);
        tc::mma_tf32_ts(d, a + 32 + kb * 8, dh, idesc, 1u);
        tc::mma_tf32_ts(d, a + kb * 8, dl, idesc, 1u);
      }
      tc::mma_commit(&P.tdone[ti]);
      if (r == 0 && P.half != nullptr) tc::mma_commit(P.half);    // 3 arrivals: the j = 0 tiles of all windows are complete
      if (r == 1) tc::mma_commit(P.done);                          // 3 arrivals: all MMAs of the phase are complete
    }
    __syncwarp();
    if (r == 0) filler();
  }
  P.n_phase += 1;                                      // identical on every thread
}
__device__ __forceinline__ void tc_wait_done(TcPipe& P) {
  tc::mbar_wait(P.done, P.n_done & 1);
  P.n_done += 1;
  tc::fence_after_sync();
}

// ---- N = 32 mixing phases (layers 2 and 3): self-contained per warp group ---------------------------------------------
// Warp group g (4 warps = the 4 TMEM lane quarters) owns window n = g: it stages the two M-tiles (position halves j) of
// its window into its own A buffers, syncs on a 128-thread named barrier, its first lane issues the 12 MMAs of the tile
// and commits to a per-tile mbarrier, and the same warps run the epilogue of their tiles as soon as that barrier flips.
// No cross-group hand-off, all 12 warps produce.  TMEM columns of group g: A(j) = 160 g + 64 j, D(0) = 160 g + 128,
// D(1) = 160 g (aliases A(0): the MMAs of tile 1 are issued by the same thread after those of tile 0 and the tensor
// pipe executes them in issue order, so A(0) is dead when D(1) is first written).
constexpr uint32_t kSmGroupCols = 160;
static_assert(kSmGroupCols * kNW <= 512, "TMEM budget of the small phases");
__device__ __forceinline__ uint32_t sm_col_a(int g, int j) { return kSmGroupCols * g + 64u * j; }
__device__ __forceinline__ uint32_t sm_col_d(int g, int j) { return j == 0 ? kSmGroupCols * g + 128u : kSmGroupCols * g; }
__device__ __forceinline__ void group_bar(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }

struct SmallPipe {
  uint64_t* done;     // [6] = (g, j)
  uint32_t tbase;
  uint32_t uses;      // completed small phases (identical on every thread): wait parity = uses & 1
};

// 32 staged channels (hi at +0..31, lo at +32..63) of one thread's position -> this warp's lanes of A buffer `abuf`
__device__ __forceinline__ void sm_store_a(uint32_t abuf, const float (&a)[32]) {
#pragma unroll
  for (int k0 = 0; k0 < 32; k0 += 16) {
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) tc::split_tf32(a[k0 + u], hi[u], lo[u]);
    tc::tmem_st16(abuf + k0, hi);
    tc::tmem_st16(abuf + 32 + k0, lo);
  }
}
// tile staged (all 4 warps of the group): publish, then one lane issues D = A(32 ch) x B^T (3xTF32) and commits
template <int N>
__device__ __forceinline__ void sm_publish_and_issue(const SmallPipe& P, int g, int j, int q, int lane, const float* Bhi, const float* Blo) {
  tc::wait_st();
  tc::fence_before_sync();
  group_bar(g);
  if (q == 0 && lane == 0) {
    tc::fence_after_sync();
    const uint32_t idesc = tc::make_idesc_tf32(128, N);
    constexpr uint32_t lbo = (N / 8) * 128, sbo = 128;
    const uint32_t bh = tc::smem_u32(Bhi), bl = tc::smem_u32(Blo);
    const uint32_t d = P.tbase + sm_col_d(g, j), a = P.tbase + sm_col_a(g, j);
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) {
      const uint64_t dh = tc::make_smem_desc(bh + kb * 2 * lbo, lbo, sbo);
      const uint64_t dl = tc::make_smem_desc(bl + kb * 2 * lbo, lbo, sbo);
      tc::mma_tf32_ts(d, a + kb * 8, dh, idesc, kb > 0 ? 1u : 0u);
      tc::mma_tf32_ts(d, a + 32 + kb * 8, dh, idesc, 1u);
      tc::mma_tf32_ts(d, a + kb * 8, dl, idesc, 1u);
    }
    tc::mma_commit(&P.done[g * 2 + j]);
  }
  __syncwarp();
}
// epilogue of the small phases: D (32 columns) -> (+bias, PReLU) -> planes; SPLIT: mix-first layout (U | Rsd)
template <bool SPLIT>
__device__ __forceinline__ void sm_epilogue(const SmallPipe& P, int g, int q, int lane, float* dst, float* dstR, const float* bias,
                                            float slope) {
  constexpr int CO = SPLIT ? 16 : 32;
  const int n = g;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    if (j == 1 && q == 3) break;                      // positions >= 224 do not exist
    const int p = j * 128 + q * 32 + lane;
    tc::mbar_wait(&P.done[g * 2 + j], P.uses & 1);
    tc::fence_after_sync();
    const uint32_t d = P.tbase + (static_cast<uint32_t>(q * 32) << 16) + sm_col_d(g, j);
    uint32_t v0[16], v1[16];
    tc::tmem_ld16(d, v0);
    tc::tmem_ld16(d + 16, v1);
    tc::wait_ld();
    if (p < kP) {
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const float a = __uint_as_float(v0[u]) + bias[u];
        const float b = __uint_as_float(v1[u]) + bias[16 + u];
        if (SPLIT) {
          dst[(n * CO + u) * kCS + p] = a;              // U: bias slot is 0
          dstR[(n * CO + u) * kCS + p] = b;             // Rsd + folded bias
        } else {
          dst[(n * CO + u) * kCS + p] = prelu(a, slope);
          dst[(n * CO + 16 + u) * kCS + p] = prelu(b, slope);
        }
      }
    }
  }
  tc::fence_before_sync();
}

template <bool kDec, int NDQ>
__global__ void __launch_bounds__(kTcThreads, 1) fused_eval_tc_kernel(const __grid_constant__ FusedTcParams Pm) {
  extern __shared__ __align__(128) float smem_tc[];
  float* R0 = smem_tc;
  float* R1 = R0 + kRBig;
  float* XB = R1 + kRBig;
  float* GB = XB + 2 * kRSmall;
  float* TB = GB + kTcGB;
  float* AB = TB + kTwFloats;
  float* WMs = AB + kAwFloats;
  float* WMb = WMs + kTcWsFloats;
  float* zpart = WMb + kTcWbFloats;
  float* zfin = zpart + 2 * kTcWarps * kNW * kDP;        // decoder variant: latents of the tile
  float* cen = zfin + kNW * kDP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(cen + 32);     // full[2], empty[2], done, small-phase done[6]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);   // bars[11]: layer-4 graph half, j = 0 tiles complete; bars[12..17]: layer-4 tiles
  int* task_ctr = reinterpret_cast<int*>(tmem_slot + 1);          // dynamic task hand-out of the layer-4 temporal stage

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (Pm.B + kNW - 1) / kNW;
  // every CTA of the grid has work (grid <= ntiles), so the TMEM allocation below is unconditional

  auto load_x = [&](float* dst, int64_t tile) {
    const int64_t w0 = tile * kNW;
    if (Pm.traj == nullptr) {
      for (int i = tid; i < kNW * 2 * kP; i += kTcThreads) {
        const int r = i / kP, p = i - r * kP;
        int64_t w = w0 + (r >> 1);
        if (w >= Pm.B) w = Pm.B - 1;
        cp_async4(dst + r * kCS + p, Pm.x + w * (2 * kP) + (r & 1) * kP + p);
      }
    } else {
      // window = kT consecutive trajectory rows of 2*kV interleaved coordinates: 408 contiguous floats in (t, v, c) order,
      // de-interleaved into the two channel planes (the sliding windows are never materialised in HBM)
      for (int i = tid; i < kNW * 2 * kP; i += kTcThreads) {
        const int n = i / (2 * kP), e = i - n * (2 * kP);      // e = p*2 + c
        int64_t w = w0 + n;
        if (w >= Pm.B) w = Pm.B - 1;
        int64_t row = Pm.win_row[w];
        row = row < 0 ? 0 : (row > Pm.traj_rows - kT ? Pm.traj_rows - kT : row);
        cp_async4(dst + (n * 2 + (e & 1)) * kCS + (e >> 1), Pm.traj + row * (2 * kV) + e);
      }
    }
  };
  // test-time affine transform of a landed input tile, in place (utils/dataset_utils.py:270-284 apply_pose_transform); run by
  // 96 threads: 32 lanes per window, so the window's matrix is fetched once per thread
  auto transform_x = [&](float* buf, int64_t tile, int t96) {
    if (Pm.trans == nullptr || t96 >= kNW * 32) return;
    const int n = t96 >> 5, l = t96 & 31;
    int64_t w = tile * kNW + n;
    if (w >= Pm.B) w = Pm.B - 1;
    int ti = Pm.trans[w];
    ti = ti < 0 ? 0 : (ti >= Pm.n_mats ? Pm.n_mats - 1 : ti);
    const float* m = Pm.mats + ti * 6;
    const float m0 = __ldg(m), m1 = __ldg(m + 1), m2 = __ldg(m + 2), m3 = __ldg(m + 3), m4 = __ldg(m + 4), m5 = __ldg(m + 5);
    float* bx = buf + (n * 2) * kCS;
    float* by = bx + kCS;
    for (int p = l; p < kP; p += 32) {
      const float x = bx[p], y = by[p];
      bx[p] = fmaf(m0, x, fmaf(m1, y, m2));
      by[p] = fmaf(m3, x, fmaf(m4, y, m5));
    }
  };
  auto acopy = [&](float* dst, const float* src, int nfloats) {
    for (int i = tid * 4; i < nfloats; i += kTcThreads * 4) cp_async16(dst + i, src + i);
  };
  // stage boundary: async copies landed and visible to the tensor-core (async) proxy, TMEM accesses ordered
  auto boundary = [&]() {
    cp_async_wait_all();
    tc::fence_proxy_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
  };

  // ---- prologue ------------------------------------------------------------------------------------
  // cen[0..15]: center (zero padded), cen[16..31]: head bias
  if (tid < 16) cen[tid] = (Pm.center != nullptr && tid < Pm.D) ? Pm.center[tid] : 0.f;
  else if (tid < 32) cen[tid] = __ldg(Pm.head_b + (tid - 16));
  if (warp == 0) tc::tmem_alloc(tmem_slot, 512);
  if (tid == 32) {
    tc::mbar_init(&bars[4], kNW);                                 // bars[0..3]: unused
    tc::mbar_init(&bars[11], kNW);
#pragma unroll
    for (int i = 0; i < kTcTiles; ++i) tc::mbar_init(&bars[12 + i], 1);
#pragma unroll
    for (int i = 0; i < 6; ++i) tc::mbar_init(&bars[5 + i], 1);
    tc::fence_mbar_init();
  }
  load_x(XB, blockIdx.x);
  acopy(TB, Pm.eTw[0], kTwFloats);
  acopy(AB, Pm.eAw[0], kAwFloats);
  acopy(WMs, Pm.mixL1, mix_blob_floats(4, 32));
  acopy(WMb, Pm.tcL2, tc_blob_floats(32, 32));
  cp_async_commit();
  boundary();
  TcPipe pipe;
  pipe.tdone = &bars[12]; pipe.done = &bars[4];
  pipe.half = nullptr; pipe.n_half = 0;
  pipe.tbase = *tmem_slot;
  pipe.n_phase = 0; pipe.n_done = 0;
  SmallPipe spipe;
  spipe.done = &bars[5]; spipe.tbase = pipe.tbase; spipe.uses = 0;
  const int gq = warp & 3, gg = warp >> 2;          // TMEM lane quarter, warp group (= window in the small phases)

  // head reduction, geometry and score of a finished tile (its per-warp partial sums are in zpart[buf]): warp 9 + n takes
  // window n.  Deferred into stage S0 of the NEXT tile (whose opening barrier publishes zpart) so that it overlaps the
  // layer-1 contraction instead of costing two barriers of its own.
  // phase 0: latents only (-> zfin, what the decoder waits for); phase 1: outputs (z, score) from zfin; phase 2: both
  auto finalize = [&](int64_t ftile, int buf, int phase) {
    const int n = warp - (kTcWarps - kNW);
    if (n < 0) return;
    const int64_t w = ftile * kNW + n;
    float s = 0.f;
    if (phase != 1) {
      if (lane < kDP && w < Pm.B) {
        s = cen[16 + lane];
#pragma unroll
        for (int ww = 0; ww < kTcWarps; ++ww) s += zpart[(buf * kTcWarps + ww) * (kNW * kDP) + n * kDP + lane];
      }
      if (kDec && lane < kDP) zfin[n * kDP + lane] = s;   // ragged last tile: the decoder runs on a finite dummy latent (0)
      if (phase == 0) return;
    } else {
      s = lane < kDP ? zfin[n * kDP + lane] : 0.f;
    }
    if (w >= Pm.B) return;
    float u[1] = {s};
    if (Pm.z != nullptr && lane < Pm.head_rows) Pm.z[w * Pm.head_rows + lane] = u[0];
    if (Pm.score != nullptr) {
      if (lane >= Pm.D) u[0] = 0.f;
      const float c[1] = {lane < 16 ? cen[lane] : 0.f};
      const float sc = score_from_latent<1>(Pm.flavour, u, c, Pm.D);
      if (lane == 0) Pm.score[w] = sc;
    }
  };

  // N = 32 mixing phase from channel planes: A = [src1 (k1 channels) | src2 (32 - k1 channels)] of this group's window, staged
  // and issued per warp group (see SmallPipe); the caller runs sm_epilogue and advances spipe.uses
  auto small_mix = [&](const float* src1, auto k1c, const float* src2, const float* blob) {
    constexpr int k1 = decltype(k1c)::value;
    const int n = gg;
    const uint32_t lane_base = spipe.tbase + (static_cast<uint32_t>(gq * 32) << 16);
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      if (jj == 0 || gq < 3) {
        const int p = jj * 128 + gq * 32 + lane;
        const int pc = p < kP ? p : kP - 1;
        float a[32];
#pragma unroll
        for (int k = 0; k < 32; ++k)
          a[k] = (k < k1) ? src1[(n * k1 + k) * kCS + pc] : src2[(n * (32 - k1) + (k - k1)) * kCS + pc];
        sm_store_a(lane_base + sm_col_a(gg, jj), a);
      }
      sm_publish_and_issue<32>(spipe, gg, jj, gq, lane, blob, blob + 32 * 32);
    }
  };

  // ---- layer-1 graph contraction of the FIRST tile (all warps).  For every later tile it is done one tile ahead by the
  //      three warps of TMEM lane quarter 3 during the head stage, where they have half the work of the other quarters
  //      (positions 224..255 do not exist): see the end of stage S11.
  boundary();
  if (Pm.trans != nullptr) { transform_x(XB, blockIdx.x, tid); __syncthreads(); }
  temporal_stage_l1<kTcWarps>(XB, GB, TB, warp, lane);
  boundary();
  spatial_stage_l1<kTcWarps>(GB, AB, warp, lane);

  int cur = 0;
  int64_t last_tile = -1;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, cur ^= 1) {
    float* X0 = XB + cur * kRSmall;
    const int64_t next_tile = tile + gridDim.x;

    // ---- S2+S3: L1 mix (K = 4, FP32 pipe, registers) -> H1 split straight into the TMEM A operand (never stored) ->
    //             L2 (32->16) mix-first on tensor cores: U (R1 rows 0..47) | Rsd (R1 rows 48..95)
    boundary();
    acopy(AB, Pm.eAw[1], kAwFloats);
    acopy(TB, Pm.eTw[1], kTwFloats);
    if (next_tile < ntiles) load_x(XB + (cur ^ 1) * kRSmall, next_tile);
    cp_async_commit();
    if (!kDec && last_tile >= 0) finalize(last_tile, cur ^ 1, 2);
    last_tile = tile;
    float* U2 = R1;
    float* Rsd2 = R1 + kNW * kC2 * kCS;
    const float slope2 = WMb[2 * 32 * 32 + 32];
    {
      const int n = gg;
      const bool has1 = gq < 3;                        // tile j = 1 holds positions 128..203: none in quarter 3
      unsigned long long h2[2][kC1 / 2];             // packed channel pairs (FFMA2)
#pragma unroll
      for (int jj = 0; jj < 2; ++jj)
#pragma unroll
        for (int c = 0; c < kC1 / 2; ++c) h2[jj][c] = 0ull;
      float in[2][4];
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int p = jj * 128 + gq * 32 + lane;
        const int pc = p < kP ? p : kP - 1;
        in[jj][0] = GB[(n * 2 + 0) * kCS + pc]; in[jj][1] = GB[(n * 2 + 1) * kCS + pc];
        in[jj][2] = X0[(n * 2 + 0) * kCS + pc]; in[jj][3] = X0[(n * 2 + 1) * kCS + pc];
      }
      const ulonglong2* w4 = reinterpret_cast<const ulonglong2*>(WMs);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned long long x0 = dup2(in[0][k]), x1 = dup2(in[1][k]);
#pragma unroll
        for (int c4 = 0; c4 < kC1 / 4; ++c4) {
          const ulonglong2 w = w4[k * (kC1 / 4) + c4];
          ffma2(h2[0][2 * c4], x0, w.x); ffma2(h2[0][2 * c4 + 1], x0, w.y);
          ffma2(h2[1][2 * c4], x1, w.x); ffma2(h2[1][2 * c4 + 1], x1, w.y);
        }
      }
      const float slope1 = WMs[4 * kC1 + kC1];
      float h[2][kC1];
#pragma unroll
      for (int c = 0; c < kC1 / 2; ++c) {
        const float b0 = WMs[4 * kC1 + 2 * c], b1 = WMs[4 * kC1 + 2 * c + 1];
        h[0][2 * c] = prelu(lo2(h2[0][c]) + b0, slope1); h[0][2 * c + 1] = prelu(hi2(h2[0][c]) + b1, slope1);
        h[1][2 * c] = prelu(lo2(h2[1][c]) + b0, slope1); h[1][2 * c + 1] = prelu(hi2(h2[1][c]) + b1, slope1);
      }
      const uint32_t lane_base = spipe.tbase + (static_cast<uint32_t>(gq * 32) << 16);
      sm_store_a(lane_base + sm_col_a(gg, 0), h[0]);
      sm_publish_and_issue<32>(spipe, gg, 0, gq, lane, WMb, WMb + 32 * 32);
      if (has1) sm_store_a(lane_base + sm_col_a(gg, 1), h[1]);
      sm_publish_and_issue<32>(spipe, gg, 1, gq, lane, WMb, WMb + 32 * 32);
    }
    sm_epilogue<true>(spipe, gg, gq, lane, U2, Rsd2, WMb + 2 * 32 * 32, 0.f);
    spipe.uses += 1;
    // ---- S4: L2 temporal in place on U
    boundary();
    acopy(WMb, Pm.tcL4X, 2 * 32 * 64);
    acopy(WMs, Pm.tcL3, tc_blob_floats(32, 32));
    cp_async_commit();
    temporal_stage_c16<kTcWarps>(U2, U2, TB, warp, lane);
    // ---- S5: L2 spatial in place + residual + PReLU -> H2 (R1 rows 0..47)
    boundary();
    acopy(TB, Pm.eTw[2], kTwFloats);
    cp_async_commit();
    spatial_stage_c16<EpiAddResPrelu, kTcWarps>(U2, AB, EpiAddResPrelu{Rsd2, slope2}, warp, lane);
    // ---- S6: L3 temporal: H2 -> G3 (R1 rows 48..95)
    boundary();
    acopy(AB, Pm.eAw[2], kAwFloats);
    cp_async_commit();
    float* H2 = R1;
    float* G3 = R1 + kNW * kC2 * kCS;
    temporal_stage_c16<kTcWarps>(H2, G3, TB, warp, lane);
    // ---- S7: L3 spatial in place on G3
    boundary();
    acopy(TB, Pm.eTw[3], kTwFloats);
    cp_async_commit();
    spatial_stage_c16<EpiIdentity, kTcWarps>(G3, AB, EpiIdentity{}, warp, lane);
    // ---- S8: L3 mix on tensor cores: [G3 | H2] x W -> H3 (R0, 32 ch)
    boundary();
    acopy(AB, Pm.eAw[3], kAwFloats);
    cp_async_commit();
    if (tid == 0) *task_ctr = 0;
    small_mix(G3, std::integral_constant<int, kC2>{}, H2, WMs);
    sm_epilogue<false>(spipe, gg, gq, lane, R0, nullptr, WMs + 2 * 32 * 32, WMs[2 * 32 * 32 + 32]);
    spipe.uses += 1;
    // ---- S9: L4 residual half issued first (inputs H3 = R0): runs on the tensor cores while the CUDA cores do the
    //          layer-4 graph contraction below;  L4 temporal: H3 (R0) -> G4 (R1)
    boundary();
    if (!kDec) acopy(WMs, Pm.mixL1, mix_blob_floats(4, 32));      // (the decoder needs WMs; it reloads the blob at S21)
    cp_async_commit();
    // (the producer warps take a temporal task after each staged tile instead of blocking on the A-buffer hand-off)
    tc_mix_phase<kC3, 0, kC4>(pipe, R0, nullptr, WMb, WMb + 32 * 64, false, warp, lane, [&]() {
      const int task = next_task(task_ctr, lane);
      if (task < 2 * kV) temporal_task_c32(R0, R1, TB, task, lane);
    });
    temporal_stage_c32<kTcWarps>(R0, R1, TB, warp, lane, task_ctr);
    // ---- S10: L4 spatial in place on R1; the residual-half MMAs are long done: reload WMb with the graph half
    boundary();
    tc_wait_done(pipe);
    acopy(TB, Pm.eTw[0], kTwFloats);
    acopy(WMb, Pm.tcL4G, tc_blob_floats(32, 64));
    cp_async_commit();
    spatial_stage_c32<kTcWarps>(R1, AB, warp, lane);
    // ---- S11: L4 graph half accumulates onto D, then PReLU + linear head straight from TMEM (H4 never stored)
    boundary();
    pipe.half = &bars[11];
    tc_mix_phase<kC3, 0, kC4>(pipe, R1, nullptr, WMb, WMb + 32 * 64, true, warp, lane);
    pipe.half = nullptr;
    // the head starts on the j = 0 tiles as soon as their MMAs are complete, while the j = 1 tiles still execute
    tc::mbar_wait(&bars[11], pipe.n_half & 1);
    pipe.n_half += 1;
    tc::fence_after_sync();
    bool all_done = false;
    {
      const float* bias = WMb + 2 * 32 * 64;
      const float slope4 = bias[kC4];
      // the head weights are re-read by every tile of every CTA: keep them in L2 ahead of the window stream that flows through it
      uint64_t l2_keep;
      asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(l2_keep));
      const int q = warp & 3, sub = warp >> 2;
      unsigned long long z2[kNW][kDP / 2];           // packed (d, d+1) accumulators for fma.rn.f32x2
#pragma unroll
      for (int n = 0; n < kNW; ++n)
#pragma unroll
        for (int dp = 0; dp < kDP / 2; ++dp) z2[n][dp] = 0ull;
      // 8 units (half j, 16-channel chunk) per lane quarter, split 3/3/2 over the three warp groups
      // (quarter 3 holds no j = 1 positions: its 4 units go 2/2/0, and its warps run the next tile's layer 1 afterwards)
      const int u0 = q < 3 ? sub * 3 : sub * 2, u1 = q < 3 ? ((sub == 2) ? 8 : u0 + 3) : (sub == 2 ? u0 : u0 + 2);
      // the head only streams the rows it has: the kernel is instantiated per NDQ = ceil(head_rows / 4) groups of 4 rows
      // (16 rows: 4, the VAE's 8 + 1: 3, latent 8: 2); the weight stream and the FMAs of this stage scale with NDQ
      for (int unit = u0; unit < u1; ++unit) {
        const int j = unit >> 2, c0 = (unit & 3) * 16;
        const int p = j * 128 + q * 32 + lane;
        if (j == 1 && !all_done) { tc_wait_done(pipe); all_done = true; }
        uint32_t v[kNW][16];
#pragma unroll
        for (int n = 0; n < kNW; ++n)
          tc::tmem_ld16(pipe.tbase + (static_cast<uint32_t>(q * 32) << 16) + kTcColD + 64u * (j * kNW + n) + c0, v[n]);
        tc::wait_ld();
        if (p < kP) {
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const float b = bias[c0 + u];
            // NDQ x LDG.128 (streamed once per tile: no L1 allocation, L2 evict_last): 4 NDQ latent rows of feature (c0+u, p) as (d, d+1) pairs
            const ulonglong2* wp = reinterpret_cast<const ulonglong2*>(Pm.head_w4) + static_cast<size_t>((c0 + u) * (kDP / 4)) * kP + p;
            unsigned long long w2[kDP / 2];
#pragma unroll
            for (int dq = 0; dq < NDQ; ++dq) {
              asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u64 {%0, %1}, [%2], %3;"
                           : "=l"(w2[2 * dq]), "=l"(w2[2 * dq + 1]) : "l"(wp + dq * kP), "l"(l2_keep));
            }
#pragma unroll
            for (int n = 0; n < kNW; ++n) {
              const float h = prelu(__uint_as_float(v[n][u]) + b, slope4);
              const unsigned long long hh = dup2(h);
#pragma unroll
              for (int dp = 0; dp < 2 * NDQ; ++dp) ffma2(z2[n][dp], hh, w2[dp]);
            }
          }
        }
      }
      if (!all_done) tc_wait_done(pipe);
      // cross-lane reduction of the 48 partial sums by recursive halving (48 SHFL instead of 240): after the four
      // halving steps lane L holds the 3 sums with index 24 b4 + 12 b3 + 6 b2 + 3 b1 + {0,1,2} (b_i = bit i of L),
      // already combined over 16 lanes; the last step adds the lane pair
      float r[kNW * kDP];
#pragma unroll
      for (int n = 0; n < kNW; ++n)
#pragma unroll
        for (int dp = 0; dp < kDP / 2; ++dp) {
          r[n * kDP + 2 * dp] = __uint_as_float(static_cast<uint32_t>(z2[n][dp] & 0xffffffffull));
          r[n * kDP + 2 * dp + 1] = __uint_as_float(static_cast<uint32_t>(z2[n][dp] >> 32));
        }
      int idx = 0;
#pragma unroll
      for (int step = 0; step < 4; ++step) {
        const int off = 16 >> step;                  // lane offset 16, 8, 4, 2
        const int half = (kNW * kDP / 2) >> step;    // values kept: 24, 12, 6, 3
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
          const float keep = up ? r[i + half] : r[i];
          const float send = up ? r[i] : r[i + half];
          r[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
        idx += up ? half : 0;
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) r[i] += __shfl_xor_sync(0xffffffffu, r[i], 1);
      if ((lane & 1) == 0) {
#pragma unroll
        for (int i = 0; i < 3; ++i) zpart[(cur * kTcWarps + warp) * (kNW * kDP) + idx + i] = r[i];
      }
    }
    // ---- layer-1 graph contraction of the NEXT tile on the three quarter-3 warps (they had half the head work).
    //      All MMAs of this tile are complete (tc_wait_done above), so the weight images in WMb are dead (the head of
    //      the other warps only reads the bias block behind them): stage the layer-2 blob there; AB (layer-4 A) is dead
    //      since S10: stage the layer-1 A.  TB already holds the layer-1 T (loaded during S10).
    if (gq == 3 && next_tile < ntiles) {
      const int t96 = gg * 32 + lane;
      if (!kDec) for (int i = t96 * 4; i < tc_blob_floats(32, 32); i += 96 * 4) cp_async16(WMb + i, Pm.tcL2 + i);
      for (int i = t96 * 4; i < kAwFloats; i += 96 * 4) cp_async16(AB + i, Pm.eAw[0] + i);
      cp_async_commit();
      float* Xn = XB + (cur ^ 1) * kRSmall;
      if (Pm.trans != nullptr) { transform_x(Xn, next_tile, t96); asm volatile("bar.sync 4, 96;" ::: "memory"); }
      temporal_stage_l1<3>(Xn, GB, TB, gg, lane);
      cp_async_wait_all();
      asm volatile("bar.sync 4, 96;" ::: "memory");
      spatial_stage_l1<3>(GB, AB, gg, lane);
    }

    if constexpr (kDec) {
      // ================= decoder (models/sts/ae.py:210-230, models/common/components.py:168-179) on the FP32 pipe ==========
      // ---- S12: latents of this tile (the decoder consumes them), stage weights of the first decoder layers
      boundary();
      acopy(TB, Pm.dTw[0], kTwFloats);
      acopy(AB, Pm.dAw[0], kAwFloats);
      acopy(WMs, Pm.tcD2, tc_blob_floats(32, 32));
      acopy(WMb, Pm.tcD3, tc_blob_floats(32, 32));
      cp_async_commit();
      finalize(tile, cur, 0);
      __syncthreads();
      finalize(tile, cur, 1);                          // outputs (z, latent score) off the decoder's critical path
      // ---- S13: folded first decoder layer: H = PReLU(M z + m0) -> R0 (32 ch)
      //      (rev_btlnk models/sts/ae.py:222 + decoder layer 0 linear part, collapsed at set_decoder)
      {
        // M (32*204 rows x DL) streams from L2 once per tile: all loads of a batch of kFB rows are issued before the first
        // FMA (as a load -> FMA loop this stage exposed the L2 latency 17 times and took 16 % of the auto-encoder's time)
        constexpr int kFB = 6;
        constexpr int kFN = (kC1 * kP) / kTcThreads;            // 17 rows per thread
        static_assert(kFN * kTcThreads == kC1 * kP, "folded decoder layer: row split");
        const int DL = Pm.DL;
        float zr[kNW][kDP];                                     // the tile's latents, once per thread
#pragma unroll
        for (int n = 0; n < kNW; ++n)
#pragma unroll
          for (int d = 0; d < kDP; ++d) zr[n][d] = (d < 8 || DL > 8) ? zfin[n * kDP + d] : 0.f;
        for (int b0 = 0; b0 < kFN; b0 += kFB) {
          float m0[kFB];
          float4 ma[kFB], mb[kFB], mc[kFB], md[kFB];
#pragma unroll
          for (int j = 0; j < kFB; ++j) {
            const int i = tid + (b0 + j < kFN ? b0 + j : kFN - 1) * kTcThreads;
            m0[j] = __ldg(Pm.dm0 + i);
            const float4* m4 = reinterpret_cast<const float4*>(Pm.dM + static_cast<size_t>(i) * DL);
            ma[j] = __ldg(m4); mb[j] = __ldg(m4 + 1);
            if (DL > 8) { mc[j] = __ldg(m4 + 2); md[j] = __ldg(m4 + 3); }
          }
#pragma unroll
          for (int j = 0; j < kFB; ++j) {
            if (b0 + j < kFN) {
              const int i = tid + (b0 + j) * kTcThreads;
              const int co = i / kP, p = i - co * kP;
#pragma unroll
              for (int n = 0; n < kNW; ++n) {
                // two independent partial sums halve the dependent FMA chain
                float o = m0[j], o2 = 0.f;
                o = fmaf(ma[j].x, zr[n][0], o); o2 = fmaf(ma[j].y, zr[n][1], o2); o = fmaf(ma[j].z, zr[n][2], o); o2 = fmaf(ma[j].w, zr[n][3], o2);
                o = fmaf(mb[j].x, zr[n][4], o); o2 = fmaf(mb[j].y, zr[n][5], o2); o = fmaf(mb[j].z, zr[n][6], o); o2 = fmaf(mb[j].w, zr[n][7], o2);
                if (DL > 8) {
                  o = fmaf(mc[j].x, zr[n][8], o); o2 = fmaf(mc[j].y, zr[n][9], o2); o = fmaf(mc[j].z, zr[n][10], o); o2 = fmaf(mc[j].w, zr[n][11], o2);
                  o = fmaf(md[j].x, zr[n][12], o); o2 = fmaf(md[j].y, zr[n][13], o2); o = fmaf(md[j].z, zr[n][14], o); o2 = fmaf(md[j].w, zr[n][15], o2);
                }
                R0[(n * kC1 + co) * kCS + p] = prelu(o + o2, Pm.d_slope0);
              }
            }
          }
        }
      }
      // ---- S14: D2 (32->16) mix-first: R0 -> U (R1 rows 0..47), Rsd (rows 48..95)
      boundary();
      float* Ud = R1;
      float* Rsdd = R1 + kNW * kC2 * kCS;
      const float dslope1 = WMs[2 * 32 * 32 + 32];
      small_mix(R0, std::integral_constant<int, kC1>{}, nullptr, WMs);                                  // tensor cores, like encoder layer 2
      sm_epilogue<true>(spipe, gg, gq, lane, Ud, Rsdd, WMs + 2 * 32 * 32, 0.f);
      spipe.uses += 1;
      // ---- S15: D2 temporal in place
      boundary();
      acopy(WMs, Pm.dWm[2], mix_blob_floats(32, 4));
      cp_async_commit();
      temporal_stage_c16<kTcWarps>(Ud, Ud, TB, warp, lane);
      // ---- S16: D2 spatial + residual + PReLU -> R1 rows 0..47
      boundary();
      acopy(TB, Pm.dTw[1], kTwFloats);
      cp_async_commit();
      spatial_stage_c16<EpiAddResPrelu, kTcWarps>(Ud, AB, EpiAddResPrelu{Rsdd, dslope1}, warp, lane);
      // ---- S17: D3 (16->32) temporal: R1 rows 0..47 -> rows 48..95
      boundary();
      acopy(AB, Pm.dAw[1], kAwFloats);
      cp_async_commit();
      temporal_stage_c16<kTcWarps>(R1, R1 + kNW * kC2 * kCS, TB, warp, lane);
      // ---- S18: D3 spatial in place
      boundary();
      acopy(TB, Pm.dTw[2], kTwFloats);
      cp_async_commit();
      spatial_stage_c16<EpiIdentity, kTcWarps>(R1 + kNW * kC2 * kCS, AB, EpiIdentity{}, warp, lane);
      // ---- S19: D3 mix -> R0 (32 ch)
      boundary();
      acopy(AB, Pm.dAw[2], kAwFloats);
      cp_async_commit();
      small_mix(R1 + kNW * kC2 * kCS, std::integral_constant<int, kC2>{}, R1, WMb);                     // tensor cores, like encoder layer 3
      sm_epilogue<false>(spipe, gg, gq, lane, R0, nullptr, WMb + 2 * 32 * 32, WMb[2 * 32 * 32 + 32]);
      spipe.uses += 1;
      // ---- S20: D4 (32->2) mix-first: R0 -> U (R1 rows 0..5), Rsd (R1 rows 6..11); GB stays reserved for the next tile's
      //           layer-1 output
      boundary();
      acopy(WMb, Pm.tcL2, tc_blob_floats(32, 32));
      cp_async_commit();
      float* U4 = R1;
      float* Rsd4 = R1 + kNW * kC0 * kCS;
      const float dslope3 = WMs[kC3 * 2 * kC0 + 2 * kC0];
      {
        EpiSplit<kC0> epi{U4, Rsd4, WMs + kC3 * 2 * kC0};
        mix_stage<kC3, 0, 2 * kC0, 4, EpiSplit<kC0>, kTcWarps>(R0, nullptr, WMs, epi, warp, lane);
      }
      // ---- S21: D4 temporal in place
      boundary();
      acopy(WMs, Pm.mixL1, mix_blob_floats(4, 32));
      cp_async_commit();
      temporal_stage_l1<kTcWarps>(U4, U4, TB, warp, lane);
      // ---- S22: D4 spatial + residual + PReLU -> xhat in U4
      boundary();
      spatial_stage_l1<kTcWarps, EpiAddResPrelu>(U4, AB, warp, lane, EpiAddResPrelu{Rsd4, dslope3});
      // ---- S23: reconstruction score mean_{c,t,v}(x - xhat)^2, optional xhat store
      boundary();
      if (warp < kNW) {
        const int64_t w = tile * kNW + warp;
        if (w < Pm.B) {
          float s = 0.f;
          for (int i = lane; i < 2 * kP; i += 32) {
            const int c = i / kP, p = i - c * kP;
            const float xh = U4[(warp * 2 + c) * kCS + p];
            const float d = X0[(warp * 2 + c) * kCS + p] - xh;
            s = fmaf(d, d, s);
            if (Pm.xhat != nullptr) Pm.xhat[w * (2 * kP) + i] = xh;
          }
          s = warp_sum(s);
          if (lane == 0 && Pm.rec_score != nullptr) Pm.rec_score[w] = s / static_cast<float>(2 * kP);
        }
      }
    }
  }
  cp_async_wait_all();
  tc::fence_before_sync();
  __syncthreads();
  if (!kDec && last_tile >= 0) finalize(last_tile, cur ^ 1, 2);
  if (warp == 0) tc::tmem_dealloc(pipe.tbase, 512);
}

}  // namespace coskad
