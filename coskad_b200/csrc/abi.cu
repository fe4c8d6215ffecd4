// abi.cu -- the C ABI of libcoskad_b200.so (see include/coskad_b200.h for the contract).
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/coskad_b200.h"
#include "aggregate.cuh"
#include "common.cuh"
#include "fold.cuh"
#include "fused_eval.cuh"
#include "fused_eval_tc.cuh"
#include "latent_ops.cuh"
#include "train_ops.cuh"
#include "train_tc.cuh"
#include "tc_test.cuh"

using namespace coskad;

struct coskad_ctx {
  int device = 0;
  int sm_count = 0;
  std::string err;
  int64_t launches = 0;
  // packed weights (one allocation each)
  float* enc_pack = nullptr;
  float* dec_pack = nullptr;
  bool enc_set = false, dec_set = false;
  int head_rows = 0;
  FusedParams fp{};
  FusedTcParams tp{};
  int fused_impl = 1;      // 1 = tensor-core mixing (fused_eval_tc_kernel), 0 = FP32 CUDA-core kernel
  // scratch
  int32_t* cnt_scratch = nullptr;
  size_t cnt_scratch_elems = 0;
  bool smem_attr_set = false;
  // training: per-CTA partial sums of the cross-CTA reductions (second stage: partial_sum_kernel), stream-ordered reuse
  float* ws = nullptr;
  size_t ws_bytes = 0;
  int train_impl = 1;      // 1 = 1x1 convolutions / weight gradients on tcgen05 (3xTF32), 0 = the FP32 CUDA-core kernels (A/B)
  int contract_rows = 0;   // rows per block of the training graph-contraction kernels: 48 (two blocks per SM) or 96 (one); 0 = unset
};

// COSKAD_CONTRACT_ROWS=96 selects the one-block-per-SM contraction kernels (A/B measurement); default 48
static int contract_rows(coskad_ctx* ctx) {
  if (ctx->contract_rows == 0) {
    const char* e = getenv("COSKAD_CONTRACT_ROWS");
    ctx->contract_rows = (e != nullptr && atoi(e) == kCRows) ? kCRows : kCRowsSmall;
  }
  return ctx->contract_rows;
}

static thread_local std::string g_create_err;

static int fail(coskad_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf; else g_create_err = buf;
  return code;
}
#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(ctx, COSKAD_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define CK_LAUNCH()                                                                           \
  do {                                                                                        \
    ctx->launches++;                                                                          \
    cudaError_t e_ = cudaGetLastError();                                                      \
    if (e_ != cudaSuccess)                                                                    \
      return fail(ctx, COSKAD_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

static inline size_t align4(size_t n) { return (n + 3) & ~static_cast<size_t>(3); }

// the partial-sum workspace: one allocation, grown on demand (never inside a stream capture: the eager warm-up steps of a
// graph-captured training loop size it first)
static int ensure_ws(coskad_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->ws_bytes) return COSKAD_OK;
  const size_t want = bytes < (static_cast<size_t>(32) << 20) ? (static_cast<size_t>(32) << 20) : bytes;
  cudaFree(ctx->ws);
  ctx->ws = nullptr;
  ctx->ws_bytes = 0;
  CK(cudaMalloc(&ctx->ws, want));
  ctx->ws_bytes = want;
  return COSKAD_OK;
}

extern "C" int coskad_abi_version(void) { return COSKAD_ABI_VERSION; }

extern "C" const char* coskad_last_error(const coskad_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

extern "C" int64_t coskad_launch_count(const coskad_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int coskad_create(coskad_ctx** out, int device, int n_frames, int n_joints) {
  coskad_ctx* ctx = nullptr;
  if (!out) return fail(nullptr, COSKAD_ERR_ARG, "out is NULL");
  *out = nullptr;
  if (n_frames != kT || n_joints != kV)
    return fail(nullptr, COSKAD_ERR_ARG,
                "libcoskad_b200 is built for %d frames x %d joints (every reference config); got %d x %d",
                kT, kV, n_frames, n_joints);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
    return fail(nullptr, COSKAD_ERR_NO_DEVICE, "no CUDA device visible: libcoskad_b200 has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(nullptr, COSKAD_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess)
    return fail(nullptr, COSKAD_ERR_CUDA, "cudaGetDeviceProperties failed");
  if (prop.major != 10)
    return fail(nullptr, COSKAD_ERR_NO_DEVICE, "device %d is sm_%d%d; libcoskad_b200 carries sm_100a code only", device,
                prop.major, prop.minor);
  ctx = new coskad_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  *out = ctx;
  return COSKAD_OK;
}

extern "C" int coskad_destroy(coskad_ctx* ctx) {
  if (!ctx) return COSKAD_OK;
  cudaSetDevice(ctx->device);
  cudaFree(ctx->enc_pack);
  cudaFree(ctx->dec_pack);
  cudaFree(ctx->cnt_scratch);
  cudaFree(ctx->ws);
  delete ctx;
  return COSKAD_OK;
}

// expected channel plan of the fused kernel
static const int kEncC[5] = {kC0, kC1, kC2, kC3, kC4};
static const int kDecC[5] = {kC4, kC3, kC2, kC1, kC0};

static int check_layer(coskad_ctx* ctx, const coskad_layer_params& L, int ci, int co, int idx, const char* what) {
  if (L.c_in != ci || L.c_out != co)
    return fail(ctx, COSKAD_ERR_ARG, "%s layer %d is %d->%d; the fused sm_100a kernel is built for %d->%d", what, idx,
                L.c_in, L.c_out, ci, co);
  if (!L.A || !L.T || !L.w1 || !L.bn1_w || !L.bn1_b || !L.bn1_rm || !L.bn1_rv || !L.prelu)
    return fail(ctx, COSKAD_ERR_ARG, "%s layer %d has NULL parameters", what, idx);
  if (L.w2 && (!L.bn2_w || !L.bn2_b || !L.bn2_rm || !L.bn2_rv))
    return fail(ctx, COSKAD_ERR_ARG, "%s layer %d residual BatchNorm parameters are NULL", what, idx);
  return COSKAD_OK;
}

extern "C" int coskad_set_encoder(coskad_ctx* ctx, int n_layers, const coskad_layer_params* L, const float* head_w,
                                  const float* head_b, int head_rows, void* stream_) {
  if (!ctx) return COSKAD_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (n_layers != 4) return fail(ctx, COSKAD_ERR_ARG, "encoder must have 4 ST_GCNN layers, got %d", n_layers);
  if (head_rows < 1 || head_rows > kDP) return fail(ctx, COSKAD_ERR_ARG, "head_rows must be in [1,%d], got %d", kDP, head_rows);
  if (!L || !head_w) return fail(ctx, COSKAD_ERR_ARG, "NULL weights");
  for (int i = 0; i < 4; ++i) {
    int rc = check_layer(ctx, L[i], kEncC[i], kEncC[i + 1], i, "encoder");
    if (rc) return rc;
  }
  CK(cudaSetDevice(ctx->device));
  // layout: per layer Tw, Aw, Wm ; then head_w[16][F], head_b[16]
  size_t off = 0, oT[4], oA[4], oW[4];
  for (int i = 0; i < 4; ++i) {
    const bool mf = kEncC[i + 1] < kEncC[i];
    const int K = mf ? kEncC[i] : 2 * kEncC[i], CO = mf ? 2 * kEncC[i + 1] : kEncC[i + 1];
    oT[i] = off; off += kTwFloats;
    oA[i] = off; off += kAwFloats;
    oW[i] = off; off += align4(mix_blob_floats(K, CO));
  }
  const size_t oHW = off; off += static_cast<size_t>(kDP) * kF;
  const size_t oHB = off; off += kDP;
  const size_t oHW4 = off; off += static_cast<size_t>(kDP) * kF;
  // tensor-core blobs of the mixing stages (fused_eval_tc.cuh)
  const size_t oT1 = off; off += align4(tc_blob_floats(8, 32));
  const size_t oT2 = off; off += align4(tc_blob_floats(32, 32));
  const size_t oT3 = off; off += align4(tc_blob_floats(32, 32));
  const size_t oT4X = off; off += align4(tc_blob_floats(32, 64));
  const size_t oT4G = off; off += align4(tc_blob_floats(32, 64));
  if (!ctx->enc_pack) CK(cudaMalloc(&ctx->enc_pack, off * sizeof(float)));
  for (int i = 0; i < 4; ++i) {
    const bool mf = kEncC[i + 1] < kEncC[i];
    fold_layer_kernel<<<8, 256, 0, st>>>(L[i], mf ? 1 : 0, ctx->enc_pack + oT[i], ctx->enc_pack + oA[i], ctx->enc_pack + oW[i]);
    CK_LAUNCH();
    ctx->fp.eTw[i] = ctx->enc_pack + oT[i];
    ctx->fp.eAw[i] = ctx->enc_pack + oA[i];
    ctx->fp.eWm[i] = ctx->enc_pack + oW[i];
  }
  pack_head_kernel<<<64, 256, 0, st>>>(head_w, head_b, head_rows, ctx->enc_pack + oHW, ctx->enc_pack + oHB);
  CK_LAUNCH();
  ctx->fp.head_w = ctx->enc_pack + oHW;
  ctx->fp.head_b = ctx->enc_pack + oHB;
  pack_head4_kernel<<<64, 256, 0, st>>>(head_w, head_rows, ctx->enc_pack + oHW4);
  CK_LAUNCH();
  ctx->tp.head_w4 = ctx->enc_pack + oHW4;
  fold_layer_tc_kernel<<<8, 256, 0, st>>>(L[0], 0, 8, 32, ctx->enc_pack + oT1);
  CK_LAUNCH();
  fold_layer_tc_kernel<<<8, 256, 0, st>>>(L[1], 1, 32, 32, ctx->enc_pack + oT2);
  CK_LAUNCH();
  fold_layer_tc_kernel<<<8, 256, 0, st>>>(L[2], 0, 32, 32, ctx->enc_pack + oT3);
  CK_LAUNCH();
  fold_layer_tc_kernel<<<8, 256, 0, st>>>(L[3], 2, 32, 64, ctx->enc_pack + oT4X);
  CK_LAUNCH();
  fold_layer_tc_kernel<<<8, 256, 0, st>>>(L[3], 3, 32, 64, ctx->enc_pack + oT4G);
  CK_LAUNCH();
  for (int i = 0; i < 4; ++i) { ctx->tp.eTw[i] = ctx->fp.eTw[i]; ctx->tp.eAw[i] = ctx->fp.eAw[i]; }
  ctx->tp.mixL1 = ctx->fp.eWm[0]; ctx->tp.tcL2 = ctx->enc_pack + oT2; ctx->tp.tcL3 = ctx->enc_pack + oT3;
  ctx->tp.tcL4X = ctx->enc_pack + oT4X; ctx->tp.tcL4G = ctx->enc_pack + oT4G;
  ctx->tp.head_w = ctx->fp.head_w; ctx->tp.head_b = ctx->fp.head_b;
  ctx->head_rows = head_rows;
  ctx->enc_set = true;
  return COSKAD_OK;
}

extern "C" int coskad_set_decoder(coskad_ctx* ctx, const float* rev_w, const float* rev_b, int latent_dim, int n_layers,
                                  const coskad_layer_params* L, void* stream_) {
  if (!ctx) return COSKAD_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (n_layers != 4) return fail(ctx, COSKAD_ERR_ARG, "decoder must have 4 ST_GCNN layers, got %d", n_layers);
  if (latent_dim != 8 && latent_dim != 16) return fail(ctx, COSKAD_ERR_ARG, "decoder latent_dim must be 8 or 16, got %d", latent_dim);
  if (!L || !rev_w || !rev_b) return fail(ctx, COSKAD_ERR_ARG, "NULL weights");
  for (int i = 0; i < 4; ++i) {
    int rc = check_layer(ctx, L[i], kDecC[i], kDecC[i + 1], i, "decoder");
    if (rc) return rc;
  }
  CK(cudaSetDevice(ctx->device));
  const int DL = latent_dim;
  size_t off = 0, oT[3], oA[3], oW[3];
  const size_t oM = off; off += static_cast<size_t>(kC3) * kP * DL;
  const size_t om0 = off; off += static_cast<size_t>(kC3) * kP;
  for (int i = 0; i < 3; ++i) {       // decoder layers 1..3
    const int ci = kDecC[i + 1], co = kDecC[i + 2];
    const bool mf = co < ci;
    const int K = mf ? ci : 2 * ci, CO = mf ? 2 * co : co;
    oT[i] = off; off += kTwFloats;
    oA[i] = off; off += kAwFloats;
    oW[i] = off; off += align4(mix_blob_floats(K, CO));
  }
  // tensor-core blobs of decoder layers 1 (32->16, mix-first) and 2 (16->32): same shapes as encoder layers 2 and 3
  const size_t oTD2 = off; off += align4(tc_blob_floats(32, 32));
  const size_t oTD3 = off; off += align4(tc_blob_floats(32, 32));
  if (!ctx->dec_pack) CK(cudaMalloc(&ctx->dec_pack, off * sizeof(float)));
  // layer 0 collapse in float64
  const int rows = (DL + 1) * kC4;
  double *In = nullptr, *G1 = nullptr, *G = nullptr;
  CK(cudaMalloc(&In, sizeof(double) * rows * kP));
  CK(cudaMalloc(&G1, sizeof(double) * rows * kP));
  CK(cudaMalloc(&G, sizeof(double) * rows * kP));
  dec_basis_kernel<<<128, 256, 0, st>>>(rev_w, rev_b, DL, kC4, In);
  CK_LAUNCH();
  dec_temporal_kernel<<<128, 256, 0, st>>>(In, L[0].T, rows, G1);
  CK_LAUNCH();
  dec_spatial_kernel<<<128, 256, 0, st>>>(G1, L[0].A, rows, G);
  CK_LAUNCH();
  dec_mix_kernel<<<128, 256, 0, st>>>(L[0], In, G, DL, ctx->dec_pack + oM, ctx->dec_pack + om0);
  CK_LAUNCH();
  float slope0 = 0.f;
  CK(cudaMemcpyAsync(&slope0, L[0].prelu, sizeof(float), cudaMemcpyDeviceToHost, st));
  for (int i = 0; i < 3; ++i) {
    const int ci = kDecC[i + 1], co = kDecC[i + 2];
    fold_layer_kernel<<<8, 256, 0, st>>>(L[i + 1], co < ci ? 1 : 0, ctx->dec_pack + oT[i], ctx->dec_pack + oA[i], ctx->dec_pack + oW[i]);
    CK_LAUNCH();
    ctx->fp.dTw[i] = ctx->dec_pack + oT[i];
    ctx->fp.dAw[i] = ctx->dec_pack + oA[i];
    ctx->fp.dWm[i] = ctx->dec_pack + oW[i];
  }
  fold_layer_tc_kernel<<<8, 256, 0, st>>>(L[1], 1, 32, 32, ctx->dec_pack + oTD2);
  CK_LAUNCH();
  fold_layer_tc_kernel<<<8, 256, 0, st>>>(L[2], 0, 32, 32, ctx->dec_pack + oTD3);
  CK_LAUNCH();
  ctx->tp.tcD2 = ctx->dec_pack + oTD2;
  ctx->tp.tcD3 = ctx->dec_pack + oTD3;
  CK(cudaStreamSynchronize(st));
  cudaFree(In); cudaFree(G1); cudaFree(G);
  ctx->fp.dM = ctx->dec_pack + oM;
  ctx->fp.dm0 = ctx->dec_pack + om0;
  ctx->fp.d_slope0 = slope0;
  ctx->fp.DL = DL;
  ctx->dec_set = true;
  return COSKAD_OK;
}

static int ensure_smem_attr(coskad_ctx* ctx) {
  if (ctx->smem_attr_set) return COSKAD_OK;
  CK(cudaFuncSetAttribute(fused_eval_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  CK(cudaFuncSetAttribute(fused_eval_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  CK(cudaFuncSetAttribute(fused_eval_tc_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
  CK(cudaFuncSetAttribute(fused_eval_tc_kernel<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
  CK(cudaFuncSetAttribute(fused_eval_tc_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
  CK(cudaFuncSetAttribute(fused_eval_tc_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
  CK(cudaFuncSetAttribute(fused_eval_tc_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
  ctx->smem_attr_set = true;
  return COSKAD_OK;
}

static int fused_grid(const coskad_ctx* ctx, int64_t B) {
  const int64_t ntiles = (B + kNW - 1) / kNW;
  return static_cast<int>(ntiles < ctx->sm_count ? ntiles : ctx->sm_count);   // persistent: one CTA per SM
}

static int launch_encode_score(coskad_ctx* ctx, int flavour, const float* x, const float* traj, int64_t traj_rows,
                               const int64_t* win_row, const int32_t* trans, const float* mats, int n_mats,
                               const float* center, int64_t B, float* z, float* score, void* stream_) {
  if (!ctx->enc_set) return fail(ctx, COSKAD_ERR_STATE, "coskad_set_encoder has not been called");
  if (flavour < COSKAD_SCORE_NONE || flavour > COSKAD_SCORE_POINCARE_HM) return fail(ctx, COSKAD_ERR_ARG, "unknown flavour %d", flavour);
  if (B == 0) return COSKAD_OK;                      // empty batch: nothing to do (its output pointers may be NULL)
  if (flavour != COSKAD_SCORE_NONE && (!score || !center)) return fail(ctx, COSKAD_ERR_ARG, "score/center is NULL for flavour %d", flavour);
  CK(cudaSetDevice(ctx->device));
  int rc = ensure_smem_attr(ctx);
  if (rc) return rc;
  FusedParams p = ctx->fp;
  p.x = x; p.center = center; p.z = z; p.score = (flavour == COSKAD_SCORE_NONE) ? nullptr : score;
  p.xhat = nullptr; p.rec_score = nullptr;
  p.B = B; p.head_rows = ctx->head_rows; p.flavour = flavour;
  // the VAE head stacks fc_var under fc_mean: the geometry sees only the latent rows
  p.D = (ctx->head_rows == 9) ? 8 : ctx->head_rows;
  if (ctx->fused_impl == 1 || traj != nullptr) {
    FusedTcParams t = ctx->tp;
    t.x = x; t.center = center; t.z = z; t.score = p.score; t.B = B; t.head_rows = p.head_rows; t.D = p.D; t.flavour = flavour;
    t.traj = traj; t.win_row = win_row; t.traj_rows = traj_rows; t.trans = trans; t.mats = mats; t.n_mats = n_mats;
    // one instantiation per number of 4-row head groups: the head streams and multiplies only the rows it has
    const int ndq = (t.head_rows + 3) / 4;
    const dim3 grid(fused_grid(ctx, B));
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    if (ndq <= 2) fused_eval_tc_kernel<false, 2><<<grid, kTcThreads, kTcSmemBytes, st>>>(t);
    else if (ndq == 3) fused_eval_tc_kernel<false, 3><<<grid, kTcThreads, kTcSmemBytes, st>>>(t);
    else fused_eval_tc_kernel<false, 4><<<grid, kTcThreads, kTcSmemBytes, st>>>(t);
  } else {
    fused_eval_kernel<false><<<fused_grid(ctx, B), kThreads, kSmemBytes, static_cast<cudaStream_t>(stream_)>>>(p);
  }
  CK_LAUNCH();
  return COSKAD_OK;
}

extern "C" int coskad_encode_score_fwd(coskad_ctx* ctx, int flavour, const float* x, const float* center, int64_t B,
                                       float* z, float* score, void* stream_) {
  if (!ctx) return COSKAD_ERR_ARG;
  if (B < 0 || (B > 0 && !x)) return fail(ctx, COSKAD_ERR_ARG, "bad x/B");
  return launch_encode_score(ctx, flavour, x, nullptr, 0, nullptr, nullptr, nullptr, 0, center, B, z, score, stream_);
}

extern "C" int coskad_encode_score_traj_fwd(coskad_ctx* ctx, int flavour, const float* traj, int64_t traj_rows,
                                            const int64_t* win_row, const int32_t* trans, const float* mats, int n_mats,
                                            const float* center, int64_t N, float* z, float* score, void* stream_) {
  if (!ctx) return COSKAD_ERR_ARG;
  if (N < 0 || (N > 0 && (!traj || !win_row))) return fail(ctx, COSKAD_ERR_ARG, "bad traj/win_row/N");
  if (N > 0 && traj_rows < kT) return fail(ctx, COSKAD_ERR_ARG, "traj_rows (%lld) is shorter than one window (%d frames)",
                                           static_cast<long long>(traj_rows), kT);
  if (trans != nullptr && (!mats || n_mats < 1)) return fail(ctx, COSKAD_ERR_ARG, "trans given without transformation matrices");
  return launch_encode_score(ctx, flavour, nullptr, traj, traj_rows, win_row, trans, mats, n_mats, center, N, z, score, stream_);
}

extern "C" int coskad_set_fused_impl(coskad_ctx* ctx, int impl) {
  if (!ctx) return COSKAD_ERR_ARG;
  if (impl != 0 && impl != 1) return fail(ctx, COSKAD_ERR_ARG, "fused impl must be 0 (FP32 CUDA cores) or 1 (tcgen05 mixing)");
  ctx->fused_impl = impl;
  return COSKAD_OK;
}

extern "C" int coskad_autoencode_score_fwd(coskad_ctx* ctx, const float* x, const float* center, int64_t B, float* z,
                                           float* xhat, float* rec_score, float* lat_score, void* stream_) {
  if (!ctx) return COSKAD_ERR_ARG;
  if (!ctx->enc_set || !ctx->dec_set) return fail(ctx, COSKAD_ERR_STATE, "encoder/decoder weights not set");
  if (ctx->head_rows != ctx->fp.DL) return fail(ctx, COSKAD_ERR_STATE, "encoder head rows (%d) != decoder latent (%d)", ctx->head_rows, ctx->fp.DL);
  if (B < 0 || (B > 0 && !x)) return fail(ctx, COSKAD_ERR_ARG, "bad x/B");
  if (lat_score && !center) return fail(ctx, COSKAD_ERR_ARG, "center is NULL");
  if (B == 0) return COSKAD_OK;
  CK(cudaSetDevice(ctx->device));
  int rc = ensure_smem_attr(ctx);
  if (rc) return rc;
  FusedParams p = ctx->fp;
  p.x = x; p.center = center; p.z = z; p.score = lat_score; p.xhat = xhat; p.rec_score = rec_score;
  p.B = B; p.head_rows = ctx->head_rows; p.D = ctx->head_rows; p.flavour = COSKAD_SCORE_EUCLID;
  if (ctx->fused_impl == 1) {
    // tensor-core encoder + FP32 decoder stages in one kernel
    FusedTcParams t = ctx->tp;
    t.x = x; t.center = center; t.z = z; t.score = lat_score; t.B = B; t.head_rows = p.head_rows; t.D = p.D; t.flavour = p.flavour;
    t.dM = p.dM; t.dm0 = p.dm0; t.d_slope0 = p.d_slope0; t.DL = p.DL;
    for (int i = 0; i < 3; ++i) { t.dTw[i] = p.dTw[i]; t.dAw[i] = p.dAw[i]; t.dWm[i] = p.dWm[i]; }
    t.xhat = xhat; t.rec_score = rec_score;
    if (t.head_rows <= 8) fused_eval_tc_kernel<true, 2><<<fused_grid(ctx, B), kTcThreads, kTcSmemBytes, static_cast<cudaStream_t>(stream_)>>>(t);
    else fused_eval_tc_kernel<true, 4><<<fused_grid(ctx, B), kTcThreads, kTcSmemBytes, static_cast<cudaStream_t>(stream_)>>>(t);
  } else {
    fused_eval_kernel<true><<<fused_grid(ctx, B), kThreads, kSmemBytes, static_cast<cudaStream_t>(stream_)>>>(p);
  }
  CK_LAUNCH();
  return COSKAD_OK;
}

// Debug aid for the parity tests: run the first tile of the fused kernel up to `stage` and copy the
// CTA's activation buffers out: [R0 | R1 | GB | GB2 | zfin] = 2*96*205 + 2*6*205 + 48 floats.
extern "C" int coskad_debug_fused_stage(coskad_ctx* ctx, int with_decoder, const float* x, int64_t B, int stage,
                                        float* dbg_out, void* stream_) {
  if (!ctx) return COSKAD_ERR_ARG;
  if (!ctx->enc_set || (with_decoder && !ctx->dec_set)) return fail(ctx, COSKAD_ERR_STATE, "weights not set");
  if (B < 1 || !x || !dbg_out) return fail(ctx, COSKAD_ERR_ARG, "bad arguments");
  CK(cudaSetDevice(ctx->device));
  int rc = ensure_smem_attr(ctx);
  if (rc) return rc;
  FusedParams p = ctx->fp;
  p.x = x; p.center = nullptr; p.z = nullptr; p.score = nullptr; p.xhat = nullptr; p.rec_score = nullptr;
  p.B = B; p.head_rows = ctx->head_rows; p.D = ctx->head_rows; p.flavour = COSKAD_SCORE_NONE;
  p.dbg = dbg_out; p.dbg_stage = stage;
  if (with_decoder) fused_eval_kernel<true><<<1, kThreads, kSmemBytes, static_cast<cudaStream_t>(stream_)>>>(p);
  else fused_eval_kernel<false><<<1, kThreads, kSmemBytes, static_cast<cudaStream_t>(stream_)>>>(p);
  CK_LAUNCH();
  return COSKAD_OK;
}
// test aid: out[128,N] = A[128,K] W^T on the tcgen05 path (3xTF32), W given as K-major canonical images (hi, lo)
extern "C" int coskad_debug_tc_mix(coskad_ctx* ctx, const float* A, const float* Bhi, const float* Blo, int K, int N,
                                   int swap_strides, float* out, void* stream_) {
  if (!ctx) return COSKAD_ERR_ARG;
  if (K % 16 != 0 || K < 16 || K > 64 || N % 16 != 0 || N < 16 || N > 64) return fail(ctx, COSKAD_ERR_ARG, "tc test: K,N in {16..64} multiples of 16");
  CK(cudaSetDevice(ctx->device));
  tc_mix_test_kernel<<<1, 128, 2 * N * K * sizeof(float), static_cast<cudaStream_t>(stream_)>>>(A, Bhi, Blo, K, N, swap_strides, out);
  CK_LAUNCH();
  return COSKAD_OK;
}
extern "C" int coskad_debug_fused_floats(void) { return 2 * kRBig + 2 * kRSmall + kNW * kDP; }
extern "C" int coskad_fused_tile_windows(void) { return kNW; }

// ---- latent ops -----------------------------------------------------------------------------------
static inline int row_grid(const coskad_ctx* ctx, int64_t B) {
  int64_t g = (B + kRowWarps - 1) / kRowWarps;
  const int64_t cap = static_cast<int64_t>(ctx->sm_count) * 8;
  return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}
#define CHECK_BD()                                                                                          \
  if (!ctx) return COSKAD_ERR_ARG;                                                                          \
  if (B < 0 || D < 1 || D > 512) return fail(ctx, COSKAD_ERR_ARG, "bad B/D (D must be in [1,512]), got B=%lld D=%d", (long long)B, D); \
  if (B == 0) return COSKAD_OK;                                                                             \
  CK(cudaSetDevice(ctx->device));

extern "C" int coskad_geom_map(coskad_ctx* ctx, int op, const float* in, int64_t B, int D, float* out, void* stream_) {
  CHECK_BD();
  if (op < 0 || op > COSKAD_MAP_L2NORMALIZE) return fail(ctx, COSKAD_ERR_ARG, "unknown map op %d", op);
  if (!in || !out) return fail(ctx, COSKAD_ERR_ARG, "NULL pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (D <= 32) geom_map_kernel<1><<<row_grid(ctx, B), kRowWarps * 32, 0, st>>>(op, in, B, D, out);
  else if (D <= 128) geom_map_kernel<4><<<row_grid(ctx, B), kRowWarps * 32, 0, st>>>(op, in, B, D, out);
  else geom_map_kernel<16><<<row_grid(ctx, B), kRowWarps * 32, 0, st>>>(op, in, B, D, out);
  CK_LAUNCH();
  return COSKAD_OK;
}

extern "C" int coskad_dist(coskad_ctx* ctx, int flavour, const float* a, const float* b, int b_bcast, int64_t B, int D,
                           float* out, void* stream_) {
  CHECK_BD();
  if (flavour < COSKAD_SCORE_POINCARE || flavour > COSKAD_SCORE_POINCARE_HM) return fail(ctx, COSKAD_ERR_ARG, "unknown flavour %d", flavour);
  if (!a || !b || !out) return fail(ctx, COSKAD_ERR_ARG, "NULL pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (D <= 32) dist_kernel<1><<<row_grid(ctx, B), kRowWarps * 32, 0, st>>>(flavour, a, b, b_bcast, B, D, out);
  else if (D <= 128) dist_kernel<4><<<row_grid(ctx, B), kRowWarps * 32, 0, st>>>(flavour, a, b, b_bcast, B, D, out);
  else dist_kernel<16><<<row_grid(ctx, B), kRowWarps * 32, 0, st>>>(flavour, a, b, b_bcast, B, D, out);
  CK_LAUNCH();
  return COSKAD_OK;
}

extern "C" int coskad_geom_map_bwd(coskad_ctx* ctx, int op, const float* in, const float* gout, int64_t B, int D, float* gin,
                                   void* stream_) {
  CHECK_BD();
  if (op != COSKAD_MAP_EXPMAP0 && op != COSKAD_MAP_PROJECT && op != COSKAD_MAP_EXPMAP0_PROJECT && op != COSKAD_MAP_L2NORMALIZE)
    return fail(ctx, COSKAD_ERR_ARG, "geom_map_bwd: op %d has no backward (geoopt flavour and L2 normalise only)", op);
  if (D > 128) return fail(ctx, COSKAD_ERR_ARG, "geom_map_bwd supports D <= 128");
  if (!in || !gout || !gin) return fail(ctx, COSKAD_ERR_ARG, "NULL pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (D <= 32) geom_map_bwd_kernel<1><<<row_grid(ctx, B), kRowWarps * 32, 0, st>>>(op, in, gout, B, D, gin);
  else geom_map_bwd_kernel<4><<<row_grid(ctx, B), kRowWarps * 32, 0, st>>>(op, in, gout, B, D, gin);
  CK_LAUNCH();
  return COSKAD_OK;
}

extern "C" int coskad_dist_bwd(coskad_ctx* ctx, int flavour, const float* a, const float* b, int b_bcast, const float* gs,
                               int64_t B, int D, float* ga, float* gb, void* stream_) {
  CHECK_BD();
  if (flavour != COSKAD_SCORE_POINCARE && flavour != COSKAD_SCORE_POINCARE_NOPROJ && flavour != COSKAD_SCORE_EUCLID &&
      flavour != COSKAD_SCORE_COSINE)
    return fail(ctx, COSKAD_ERR_ARG, "dist_bwd: flavour %d has no backward", flavour);
  if (D > 128) return fail(ctx, COSKAD_ERR_ARG, "dist_bwd supports D <= 128");
  if (!a || !b || !gs || (!ga && !gb)) return fail(ctx, COSKAD_ERR_ARG, "NULL pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (D <= 32) dist_bwd_kernel<1><<<row_grid(ctx, B), kRowWarps * 32, 0, st>>>(flavour, a, b, b_bcast, gs, B, D, ga, gb);
  else dist_bwd_kernel<4><<<row_grid(ctx, B), kRowWarps * 32, 0, st>>>(flavour, a, b, b_bcast, gs, B, D, ga, gb);
  CK_LAUNCH();
  return COSKAD_OK;
}

extern "C" int coskad_dist0(coskad_ctx* ctx, const float* x, int64_t B, int D, float* out, void* stream_) {
  CHECK_BD();
  if (!x || !out) return fail(ctx, COSKAD_ERR_ARG, "NULL pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (D <= 32) dist0_kernel<1><<<row_grid(ctx, B), kRowWarps * 32, 0, st>>>(x, B, D, out);
  else if (D <= 128) dist0_kernel<4><<<row_grid(ctx, B), kRowWarps * 32, 0, st>>>(x, B, D, out);
  else dist0_kernel<16><<<row_grid(ctx, B), kRowWarps * 32, 0, st>>>(x, B, D, out);
  CK_LAUNCH();
  return COSKAD_OK;
}

extern "C" int coskad_ps_sample(coskad_ctx* ctx, const float* mu, const float* t, const float* v, int64_t B, int D,
                                float* z, void* stream_) {
  CHECK_BD();
  if (D < 2 || D > 32) return fail(ctx, COSKAD_ERR_ARG, "ps_sample supports 2 <= D <= 32");
  if (!mu || !t || !v || !z) return fail(ctx, COSKAD_ERR_ARG, "NULL pointer");
  ps_sample_kernel<<<row_grid(ctx, B), kRowWarps * 32, 0, static_cast<cudaStream_t>(stream_)>>>(mu, t, v, B, D, z);
  CK_LAUNCH();
  return COSKAD_OK;
}

extern "C" int coskad_poincare_score_bwd(coskad_ctx* ctx, const float* z, const float* center, const float* dscore,
                                         int64_t B, int D, int with_project, float* dz, void* stream_) {
  CHECK_BD();
  if (D > 32) return fail(ctx, COSKAD_ERR_ARG, "poincare_score_bwd supports D <= 32");
  if (!z || !center || !dscore || !dz) return fail(ctx, COSKAD_ERR_ARG, "NULL pointer");
  poincare_score_bwd_kernel<<<row_grid(ctx, B), kRowWarps * 32, 0, static_cast<cudaStream_t>(stream_)>>>(
      z, center, dscore, B, D, with_project, dz);
  CK_LAUNCH();
  return COSKAD_OK;
}

extern "C" int coskad_center_partial(coskad_ctx* ctx, int flavour, const float* zproj, int64_t B, int D, double* acc,
                                     void* stream_) {
  CHECK_BD();
  if (D > 128) return fail(ctx, COSKAD_ERR_ARG, "center ops support D <= 128");
  if (!zproj || !acc) return fail(ctx, COSKAD_ERR_ARG, "NULL pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  int g = row_grid(ctx, B);
  if (g > ctx->sm_count * 2) g = ctx->sm_count * 2;
  // per-CTA float64 partials, added to acc in a fixed order (no atomics: the center is bit-reproducible)
  { const int rc = ensure_ws(ctx, sizeof(double) * static_cast<size_t>(g) * (D + 2)); if (rc) return rc; }
  double* part = reinterpret_cast<double*>(ctx->ws);
  if (D <= 32) center_partial_kernel<1><<<g, kRowWarps * 32, 0, st>>>(flavour, zproj, B, D, part);
  else center_partial_kernel<4><<<g, kRowWarps * 32, 0, st>>>(flavour, zproj, B, D, part);
  CK_LAUNCH();
  center_partial_final_kernel<<<1, dim3(32, 8), 0, st>>>(part, g, D, acc);
  CK_LAUNCH();
  return COSKAD_OK;
}

extern "C" int coskad_mahalanobis(coskad_ctx* ctx, const float* z, const float* center, const float* VI, int64_t B, int D,
                                  float* out, void* stream_) {
  CHECK_BD();
  if (D > kMahD) return fail(ctx, COSKAD_ERR_ARG, "mahalanobis supports D <= %d, got %d", kMahD, D);
  if (!z || !center || !VI || !out) return fail(ctx, COSKAD_ERR_ARG, "NULL pointer");
  mahalanobis_kernel<<<row_grid(ctx, B), kRowWarps * 32, 0, static_cast<cudaStream_t>(stream_)>>>(z, center, VI, B, D, out);
  CK_LAUNCH();
  return COSKAD_OK;
}

extern "C" int coskad_mahalanobis_bwd(coskad_ctx* ctx, const float* z, const float* center, const float* VI, const float* gs,
                                      int64_t B, int D, float* gz, void* stream_) {
  CHECK_BD();
  if (D > kMahD) return fail(ctx, COSKAD_ERR_ARG, "mahalanobis supports D <= %d, got %d", kMahD, D);
  if (!z || !center || !VI || !gs || !gz) return fail(ctx, COSKAD_ERR_ARG, "NULL pointer");
  mahalanobis_bwd_kernel<<<row_grid(ctx, B), kRowWarps * 32, 0, static_cast<cudaStream_t>(stream_)>>>(z, center, VI, gs, B, D, gz);
  CK_LAUNCH();
  return COSKAD_OK;
}

extern "C" int coskad_cov_partial(coskad_ctx* ctx, const float* z, const float* mu, int64_t B, int D, double* acc, void* stream_) {
  CHECK_BD();
  if (D > kMahD) return fail(ctx, COSKAD_ERR_ARG, "cov_partial supports D <= %d, got %d", kMahD, D);
  if (!z || !mu || !acc) return fail(ctx, COSKAD_ERR_ARG, "NULL pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  int g = row_grid(ctx, B);
  if (g > ctx->sm_count * 2) g = ctx->sm_count * 2;
  const int n = D * D + 1;
  { const int rc = ensure_ws(ctx, sizeof(double) * static_cast<size_t>(g) * n); if (rc) return rc; }
  double* part = reinterpret_cast<double*>(ctx->ws);
  if (D <= 8) cov_partial_kernel<8><<<g, kRowWarps * 32, 0, st>>>(z, mu, B, D, part);
  else if (D <= 16) cov_partial_kernel<16><<<g, kRowWarps * 32, 0, st>>>(z, mu, B, D, part);
  else cov_partial_kernel<32><<<g, kRowWarps * 32, 0, st>>>(z, mu, B, D, part);
  CK_LAUNCH();
  cov_partial_final_kernel<<<(n + 127) / 128, 128, 0, st>>>(part, g, n, acc);
  CK_LAUNCH();
  return COSKAD_OK;
}

extern "C" int coskad_center_finalize(coskad_ctx* ctx, int flavour, const double* acc, int D, float eps, float* center,
                                      void* stream_) {
  if (!ctx) return COSKAD_ERR_ARG;
  if (D < 1 || D > 128 || !acc || !center) return fail(ctx, COSKAD_ERR_ARG, "bad arguments");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (D <= 32) center_finalize_kernel<1><<<1, 32, 0, st>>>(flavour, acc, D, eps, center);
  else center_finalize_kernel<4><<<1, 32, 0, st>>>(flavour, acc, D, eps, center);
  CK_LAUNCH();
  return COSKAD_OK;
}

// shift + Gaussian smoothing of n_curves float64 curves (curve_off CSR); weights [2*radius+1] device doubles
extern "C" int coskad_score_process(coskad_ctx* ctx, const double* curves, const int64_t* curve_off, int64_t n_curves,
                                    int shift, const double* weights, int radius, double* out, void* stream_) {
  if (!ctx) return COSKAD_ERR_ARG;
  if (n_curves < 0 || shift < 0 || radius < 0) return fail(ctx, COSKAD_ERR_ARG, "negative argument");
  if (n_curves == 0) return COSKAD_OK;
  if (!curves || !curve_off || !weights || !out) return fail(ctx, COSKAD_ERR_ARG, "NULL pointer");
  if (curves == out) return fail(ctx, COSKAD_ERR_ARG, "score_process cannot run in place");
  CK(cudaSetDevice(ctx->device));
  score_process_kernel<<<static_cast<unsigned>(n_curves), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      curves, curve_off, n_curves, shift, weights, radius, out);
  CK_LAUNCH();
  return COSKAD_OK;
}

extern "C" int coskad_frame_aggregate(coskad_ctx* ctx, const float* score, const int64_t* frames, int T,
                                      const int64_t* win_idx, const int64_t* person_off, const int32_t* person_clip,
                                      const int64_t* person_out_off, int64_t n_persons, const int64_t* clip_person_off,
                                      const int64_t* clip_off, int64_t n_clips, int64_t total_person_frames,
                                      int64_t max_clip_frames, double* person_out, double* out, void* stream_) {
  if (!ctx) return COSKAD_ERR_ARG;
  if (T < 1 || T > 32) return fail(ctx, COSKAD_ERR_ARG, "T must be in [1,32], got %d", T);
  if (n_persons < 0 || n_clips < 0) return fail(ctx, COSKAD_ERR_ARG, "negative counts");
  if (n_persons == 0 || n_clips == 0) return COSKAD_OK;
  if (!score || !frames || !win_idx || !person_off || !person_clip || !person_out_off || !clip_person_off || !clip_off ||
      !person_out || !out)
    return fail(ctx, COSKAD_ERR_ARG, "NULL pointer");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (static_cast<size_t>(total_person_frames) > ctx->cnt_scratch_elems) {
    CK(cudaStreamSynchronize(st));
    cudaFree(ctx->cnt_scratch);
    ctx->cnt_scratch = nullptr;
    CK(cudaMalloc(&ctx->cnt_scratch, sizeof(int32_t) * static_cast<size_t>(total_person_frames)));
    ctx->cnt_scratch_elems = static_cast<size_t>(total_person_frames);
  }
  const int g1 = static_cast<int>((n_persons + kAggWarps - 1) / kAggWarps);
  person_curves_kernel<<<g1, kAggWarps * 32, 0, st>>>(score, frames, T, win_idx, person_off, person_clip, n_persons,
                                                       clip_off, person_out, ctx->cnt_scratch, person_out_off);
  CK_LAUNCH();
  if (n_clips > 65535) return fail(ctx, COSKAD_ERR_ARG, "more than 65535 clips per call");
  int gx = static_cast<int>((max_clip_frames + 255) / 256);
  if (gx < 1) gx = 1;
  clip_max_kernel<<<dim3(gx, static_cast<unsigned>(n_clips)), 256, 0, st>>>(person_out, person_out_off, clip_person_off,
                                                                          clip_off, n_clips, out);
  CK_LAUNCH();
  return COSKAD_OK;
}

// ---- diagnostics -----------------------------------------------------------------------------------
__global__ void ffma_peak_kernel(float* out, int iters, float a, float b) {
  float r[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = a * static_cast<float>(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = fmaf(r[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += r[i];
  if (s == 12345.678f) out[0] = s;     // never true; keeps the loop alive
}

extern "C" int coskad_measure_fp32_peak(coskad_ctx* ctx, double* tflops, void* stream_) {
  if (!ctx || !tflops) return COSKAD_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  float* d = nullptr;
  CK(cudaMalloc(&d, 16));
  const int blocks = ctx->sm_count * 2, threads = 1024, iters = 1 << 15;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  ffma_peak_kernel<<<blocks, threads, 0, st>>>(d, iters, 0.999f, 0.001f);   // warm-up
  CK_LAUNCH();
  double best = 0.0;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0, st));
    ffma_peak_kernel<<<blocks, threads, 0, st>>>(d, iters, 0.999f, 0.001f);
    CK_LAUNCH();
    CK(cudaEventRecord(e1, st));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 16.0 * iters * static_cast<double>(blocks) * threads;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  *tflops = best;
  return COSKAD_OK;
}

extern "C" int coskad_measure_tf32_peak(coskad_ctx* ctx, double* tflops, void* stream_) {
  if (!ctx || !tflops) return COSKAD_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  float* d = nullptr;
  CK(cudaMalloc(&d, 16));
  const int blocks = ctx->sm_count, iters = 1 << 14;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  tf32_peak_kernel<<<blocks, 128, 0, st>>>(256, d);   // warm-up
  CK_LAUNCH();
  double best = 0.0;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0, st));
    tf32_peak_kernel<<<blocks, 128, 0, st>>>(iters, d);
    CK_LAUNCH();
    CK(cudaEventRecord(e1, st));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 128.0 * 256.0 * 8.0 * iters * static_cast<double>(blocks);
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
  *tflops = best;
  return COSKAD_OK;
}

// ---- training path --------------------------------------------------------------------------------
#define TRAIN_PRE()                                    \
  if (!ctx) return COSKAD_ERR_ARG;                     \
  CK(cudaSetDevice(ctx->device));                      \
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
static inline int ew_grid(const coskad_ctx* ctx, int64_t n, int threads = kTrainThreads) {
  int64_t g = (n + threads - 1) / threads;
  const int64_t cap = static_cast<int64_t>(ctx->sm_count) * 8;
  return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}
static bool chan_ok(int c) { return c == 2 || c == 16 || c == 32 || c == 64; }

extern "C" int coskad_train_contract_fwd(coskad_ctx* ctx, const float* X, const float* A, const float* T, int64_t R,
                                         float* G1, float* G, void* stream_) {
  TRAIN_PRE();
  if (R <= 0) return COSKAD_OK;
  if (contract_rows(ctx) == kCRows) {
    const int64_t nblk = (R + kCRows - 1) / kCRows;
    const int g = static_cast<int>(nblk < ctx->sm_count ? nblk : ctx->sm_count);
    CK(cudaFuncSetAttribute(train_contract_fwd_kernel<kCRows>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCSmemBytes));
    train_contract_fwd_kernel<kCRows><<<g, kCThreads, kCSmemBytes, st>>>(X, A, T, R, G1, G);
  } else {
    constexpr int smem = contract_smem_bytes(kCRowsSmall);
    const int64_t nblk = (R + kCRowsSmall - 1) / kCRowsSmall;
    const int g = static_cast<int>(nblk < 2 * ctx->sm_count ? nblk : 2 * ctx->sm_count);
    CK(cudaFuncSetAttribute(train_contract_fwd_kernel<kCRowsSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    train_contract_fwd_kernel<kCRowsSmall><<<g, kCThreads, smem, st>>>(X, A, T, R, G1, G);
  }
  CK_LAUNCH();
  return COSKAD_OK;
}

template <typename TOut>
static int launch_partial_sum(coskad_ctx* ctx, const float* part, int nparts, int64_t stride, int64_t n, TOut* out, cudaStream_t st) {
  if (nparts >= 32) partial_sum_wide_kernel<TOut><<<static_cast<unsigned>((n + 31) / 32), dim3(32, kPsRows), 0, st>>>(part, nparts, stride, n, out);
  else partial_sum_kernel<TOut><<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(part, nparts, stride, n, out);
  CK_LAUNCH();
  return COSKAD_OK;
}

extern "C" int coskad_set_train_impl(coskad_ctx* ctx, int impl) {
  if (!ctx) return COSKAD_ERR_ARG;
  if (impl != 0 && impl != 1) return fail(ctx, COSKAD_ERR_ARG, "train impl must be 0 (FP32 CUDA cores) or 1 (tcgen05), got %d", impl);
  ctx->train_impl = impl;
  return COSKAD_OK;
}

extern "C" int coskad_train_contract_bwd(coskad_ctx* ctx, const float* dG, const float* dXres, const float* X,
                                         const float* G1, const float* A, const float* T, int64_t R, float* dX,
                                         float* dA, float* dT, void* stream_) {
  TRAIN_PRE();
  if (R <= 0) return COSKAD_OK;
  const bool big = contract_rows(ctx) == kCRows;
  const int nr = big ? kCRows : kCRowsSmall, cap = big ? ctx->sm_count : 2 * ctx->sm_count;
  const int64_t nblk = (R + nr - 1) / nr;
  const int g = static_cast<int>(nblk < cap ? nblk : cap);
  { const int rc = ensure_ws(ctx, sizeof(float) * static_cast<size_t>(g) * kContractPart); if (rc) return rc; }
  if (big) {
    CK(cudaFuncSetAttribute(train_contract_bwd_kernel<kCRows>, cudaFuncAttributeMaxDynamicSharedMemorySize, kCSmemBytes));
    train_contract_bwd_kernel<kCRows><<<g, kCThreads, kCSmemBytes, st>>>(dG, dXres, X, G1, A, T, R, dX, ctx->ws);
  } else {
    constexpr int smem = contract_smem_bytes(kCRowsSmall);
    CK(cudaFuncSetAttribute(train_contract_bwd_kernel<kCRowsSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    train_contract_bwd_kernel<kCRowsSmall><<<g, kCThreads, smem, st>>>(dG, dXres, X, G1, A, T, R, dX, ctx->ws);
  }
  CK_LAUNCH();
  if (dT == dA + kT * kV * kV)      // the usual case (one zeroed gradient buffer per layer): one second-stage launch for both
    return launch_partial_sum<float>(ctx, ctx->ws, g, kContractPart, kContractPart, dA, st);
  { const int rc = launch_partial_sum<float>(ctx, ctx->ws, g, kContractPart, kT * kV * kV, dA, st); if (rc) return rc; }
  return launch_partial_sum<float>(ctx, ctx->ws + kT * kV * kV, g, kContractPart, kV * kT * kT, dT, st);
}

// chan_gemm_kernel launcher: K input channels -> M output channels over B*204 positions
static int launch_chan_gemm(coskad_ctx* ctx, bool stats_on, const float* in1, const float* in2, const float* Wa, const float* Wb,
                            int w_is_km, const float* b1, const float* b2, int64_t B, int K, int M, float* out1, float* out2,
                            double* stats, cudaStream_t st) {
  const int TM = M < 4 ? M : 4, n_ct = M / TM;
  const int CE = (n_ct >= kGWarps) ? 128 : 256;          // fewer than 8 output tiles: two 128-position slabs per chunk
  const size_t smem = sizeof(float) * (2 * static_cast<size_t>(K) * M + 2 * static_cast<size_t>(K) * (CE + 4) + 4 * static_cast<size_t>(M));
  const int64_t nchunk = (B * kP + CE - 1) / CE;
  const int64_t cap = static_cast<int64_t>(ctx->sm_count) * 2;
  const int g = static_cast<int>(nchunk < cap ? nchunk : cap);
  if (stats_on) {
    CK(cudaFuncSetAttribute(chan_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    chan_gemm_kernel<true><<<g, kGThreads, smem, st>>>(in1, in2, Wa, Wb, w_is_km, b1, b2, B, K, M, CE, out1, out2, stats);
  } else {
    CK(cudaFuncSetAttribute(chan_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    chan_gemm_kernel<false><<<g, kGThreads, smem, st>>>(in1, in2, Wa, Wb, w_is_km, b1, b2, B, K, M, CE, out1, out2, stats);
  }
  CK_LAUNCH();
  return COSKAD_OK;
}

// Dynamic shared memory requested only to make the block scheduler's CTAs-per-SM limit equal to what TMEM can hold: with
// the kernels' small footprint it places up to 5-9 CTAs on an SM, the ones past the TMEM capacity block inside tcgen05.alloc
// until a resident CTA exits, and other SMs sit idle meanwhile (measured: 2-3x the kernel time).
template <typename K>
static size_t tmem_pad_smem(K kernel, int per_sm) {
  cudaFuncAttributes attr{};
  if (cudaFuncGetAttributes(&attr, kernel) != cudaSuccess) return 0;
  const size_t target = (static_cast<size_t>(228) * 1024) / static_cast<size_t>(per_sm) - 1536;   // 1 KB per CTA is reserved
  return target > attr.sharedSizeBytes ? target - attr.sharedSizeBytes : 0;
}
struct BnFinalizeArgs {
  float eps, momentum;
  float *rm1, *rv1, *rm2, *rv2, *mi;
  int64_t *nbt1, *nbt2;
};
// tcgen05 forward convolution: y1, y2 and the BatchNorm statistics (per-warp partials -> fixed-order second stage)
template <int CIP, int CO, int COP>
static int launch_tc_mix_fwd(coskad_ctx* ctx, const float* G, const float* X, const float* W1, const float* b1, const float* W2,
                             const float* b2, int64_t B, int CI, float* y1, float* y2, double* stats, cudaStream_t st,
                             const BnFinalizeArgs* bn = nullptr) {
  const int64_t E = B * kP, ntiles = (E + kTcT - 1) / kTcT;
  int per_sm = 512 / tmem_alloc_cols(4 * CIP + 2 * COP);                // TMEM columns bound the resident CTAs
  if (per_sm > 4) per_sm = 4;
  const int64_t cap = static_cast<int64_t>(ctx->sm_count) * per_sm;
  const int g = static_cast<int>(ntiles < cap ? ntiles : cap);
  { const int rc = ensure_ws(ctx, sizeof(float) * static_cast<size_t>(g) * 4 * CO); if (rc) return rc; }
  // no shared-memory padding here (tmem_pad_smem): measured 15-35 % slower for this kernel -- the large carve-out shrinks L1,
  // and the 128-byte row segments of neighbouring warps share cache lines (rows start at multiples of 816 B)
  tc_mix_fwd_kernel<CIP, CO, COP><<<g, kTcT, 0, st>>>(G, X, W1, b1, W2, b2, E, CI, y1, y2, ctx->ws);
  CK_LAUNCH();
  if (bn != nullptr) {      // statistics second stage + BatchNorm finalize in one launch
    train_bn_stats_finalize_kernel<<<(CO + 7) / 8, dim3(32, kBsfRows), 0, st>>>(ctx->ws, g, static_cast<double>(E), CO, bn->eps,
                                                                                bn->momentum, bn->rm1, bn->rv1, bn->rm2, bn->rv2,
                                                                                bn->mi, bn->nbt1, bn->nbt2);
    CK_LAUNCH();
    return COSKAD_OK;
  }
  return launch_partial_sum<double>(ctx, ctx->ws, g, 4 * CO, 4 * CO, stats, st);
}

extern "C" int coskad_train_mix_fwd(coskad_ctx* ctx, const float* G, const float* X, const float* W1, const float* b1,
                                    const float* W2, const float* b2, int64_t B, int CI, int CO, float* y1, float* y2,
                                    double* stats, void* stream_) {
  TRAIN_PRE();
  if (!chan_ok(CI) || !chan_ok(CO)) return fail(ctx, COSKAD_ERR_ARG, "training kernels support channels {2,16,32,64}, got %d->%d", CI, CO);
  if (B <= 0) return COSKAD_OK;
  if (ctx->train_impl == 1) {
#define TC_FWD(ci, co, cip, cop) if (CI == ci && CO == co) return launch_tc_mix_fwd<cip, co, cop>(ctx, G, X, W1, b1, W2, b2, B, CI, y1, y2, stats, st)
    TC_FWD(2, 32, 8, 32); TC_FWD(32, 16, 32, 16); TC_FWD(16, 32, 16, 32); TC_FWD(32, 64, 32, 64);      // encoder
    TC_FWD(64, 32, 64, 32); TC_FWD(32, 2, 32, 16);                                                       // decoder
#undef TC_FWD
  }
  return launch_chan_gemm(ctx, true, G, X, W1, W2, 0, b1, b2, B, CI, CO, y1, y2, stats, st);
}

extern "C" int coskad_train_mix_fwd_bn(coskad_ctx* ctx, const float* G, const float* X, const float* W1, const float* b1,
                                       const float* W2, const float* b2, int64_t B, int CI, int CO, float* y1, float* y2,
                                       float eps, float momentum, float* rm1, float* rv1, float* rm2, float* rv2, float* mi,
                                       int64_t* nbt1, int64_t* nbt2, void* stream_) {
  TRAIN_PRE();
  if (!chan_ok(CI) || !chan_ok(CO)) return fail(ctx, COSKAD_ERR_ARG, "training kernels support channels {2,16,32,64}, got %d->%d", CI, CO);
  if (ctx->train_impl != 1) return fail(ctx, COSKAD_ERR_STATE, "coskad_train_mix_fwd_bn is the tensor-core path (coskad_set_train_impl(ctx, 1))");
  if (!mi) return fail(ctx, COSKAD_ERR_ARG, "coskad_train_mix_fwd_bn: mi is NULL");
  if (B <= 0) return COSKAD_OK;
  const BnFinalizeArgs bn{eps, momentum, rm1, rv1, rm2, rv2, mi, nbt1, nbt2};
#define TC_FWD(ci, co, cip, cop) if (CI == ci && CO == co) return launch_tc_mix_fwd<cip, co, cop>(ctx, G, X, W1, b1, W2, b2, B, CI, y1, y2, nullptr, st, &bn)
  TC_FWD(2, 32, 8, 32); TC_FWD(32, 16, 32, 16); TC_FWD(16, 32, 16, 32); TC_FWD(32, 64, 32, 64);      // encoder
  TC_FWD(64, 32, 64, 32); TC_FWD(32, 2, 32, 16);                                                       // decoder
#undef TC_FWD
  return fail(ctx, COSKAD_ERR_ARG, "coskad_train_mix_fwd_bn: no tensor-core kernel for %d -> %d channels", CI, CO);
}

extern "C" int coskad_train_bn_finalize(coskad_ctx* ctx, const double* stats, int64_t n_per_channel, int CO, float eps,
                                        float momentum, float* rm1, float* rv1, float* rm2, float* rv2, float* mi,
                                        int64_t* nbt1, int64_t* nbt2, void* stream_) {
  TRAIN_PRE();
  train_bn_finalize_kernel<<<(CO + 63) / 64, 64, 0, st>>>(stats, static_cast<double>(n_per_channel), CO, eps, momentum, rm1,
                                                         rv1, rm2, rv2, mi, nbt1, nbt2);
  CK_LAUNCH();
  return COSKAD_OK;
}

extern "C" int coskad_train_bn_prelu_bwd_grads(coskad_ctx* ctx, const float* dout, const float* y1, const float* y2,
                                               const float* mi, const float* g1, const float* be1, const float* g2,
                                               const float* be2, const float* slope, int64_t B, int CO, double* red,
                                               float* dg1, float* dbe1, float* dg2, float* dbe2, float* dslope, void* stream_) {
  TRAIN_PRE();
  if (B <= 0) return COSKAD_OK;
  if (!red) return fail(ctx, COSKAD_ERR_ARG, "coskad_train_bn_prelu_bwd_grads: red is NULL");
  if (CO > 64) return fail(ctx, COSKAD_ERR_ARG, "bn_prelu_bwd supports c_out <= 64, got %d", CO);
  int nb = static_cast<int>((B + 7) / 8);
  const int cap = (ctx->sm_count * 8 + CO - 1) / CO;
  if (nb > cap) nb = cap;
  { const int rc = ensure_ws(ctx, sizeof(float) * static_cast<size_t>(nb) * 4 * CO); if (rc) return rc; }
  train_bn_prelu_bwd_reduce_kernel<<<dim3(CO, nb), kTrainThreads, 0, st>>>(dout, y1, y2, mi, g1, be1, g2, be2, slope, B, CO, ctx->ws);
  CK_LAUNCH();
  train_bn_prelu_bwd_reduce_final_kernel<true><<<(3 * CO + 31) / 32 + 1, dim3(32, kPsRows), 0, st>>>(ctx->ws, nb, CO, red, dg1, dbe1,
                                                                                                 dg2, dbe2, dslope);
  CK_LAUNCH();
  return COSKAD_OK;
}

extern "C" int coskad_train_bn_param_grads(coskad_ctx* ctx, const double* red, int CO, float* dg1, float* dbe1, float* dg2,
                                           float* dbe2, float* dslope, void* stream_) {
  TRAIN_PRE();
  if (red == nullptr || CO <= 0) return fail(ctx, COSKAD_ERR_ARG, "coskad_train_bn_param_grads: red is NULL or CO = %d", CO);
  train_bn_param_grads_kernel<<<(CO + 63) / 64, 64, 0, st>>>(red, CO, dg1, dbe1, dg2, dbe2, dslope);
  CK_LAUNCH();
  return COSKAD_OK;
}

extern "C" int coskad_train_bn_prelu_fwd(coskad_ctx* ctx, const float* y1, const float* y2, const float* mi,
                                         const float* g1, const float* be1, const float* g2, const float* be2,
                                         const float* slope, int64_t B, int CO, float* out, void* stream_) {
  TRAIN_PRE();
  if (B <= 0) return COSKAD_OK;
  train_bn_prelu_fwd_kernel<<<ew_grid(ctx, B * CO * 32), kTrainThreads, 0, st>>>(y1, y2, mi, g1, be1, g2, be2, slope, B, CO, out);
  CK_LAUNCH();
  return COSKAD_OK;
}

extern "C" int coskad_train_bn_prelu_bwd(coskad_ctx* ctx, const float* dout, const float* y1, const float* y2,
                                         const float* mi, const float* g1, const float* be1, const float* g2,
                                         const float* be2, const float* slope, int64_t B, int CO, double* red,
                                         float* dy1, float* dy2, void* stream_) {
  TRAIN_PRE();
  if (B <= 0) return COSKAD_OK;
  if (3 * CO + 32 > kTrainThreads) return fail(ctx, COSKAD_ERR_ARG, "bn_prelu_bwd supports c_out <= 64, got %d", CO);
  int nb = static_cast<int>((B + 7) / 8);
  const int cap = (ctx->sm_count * 8 + CO - 1) / CO;
  if (nb > cap) nb = cap;
  { const int rc = ensure_ws(ctx, sizeof(float) * static_cast<size_t>(nb) * 4 * CO); if (rc) return rc; }
  train_bn_prelu_bwd_reduce_kernel<<<dim3(CO, nb), kTrainThreads, 0, st>>>(dout, y1, y2, mi, g1, be1, g2, be2, slope, B, CO, ctx->ws);
  CK_LAUNCH();
  train_bn_prelu_bwd_reduce_final_kernel<false><<<(3 * CO + 31) / 32 + 1, dim3(32, kPsRows), 0, st>>>(ctx->ws, nb, CO, red, nullptr,
                                                                                                  nullptr, nullptr, nullptr, nullptr);
  CK_LAUNCH();
  if (dy1 && dy2) {        // NULL: the tensor-core backward (coskad_train_mix_bwd_tc) applies the BatchNorm / PReLU backward itself
    train_bn_prelu_bwd_apply_kernel<<<ew_grid(ctx, B * CO * 32), kTrainThreads, 0, st>>>(dout, y1, y2, mi, g1, be1, g2, be2,
                                                                                      slope, red, B, CO, dy1, dy2);
    CK_LAUNCH();
  }
  return COSKAD_OK;
}

template <int CO, int COK, int NP>
static int launch_tc_bwd_data(coskad_ctx* ctx, const float* dout, const float* y1, const float* y2, const float* mi,
                              const float* g1, const float* be1, const float* g2, const float* be2, const float* slope,
                              const double* red, const float* W1, const float* W2, int64_t B, int CI, float* dy1, float* dy2,
                              float* dG, float* dXres, cudaStream_t st) {
  const int64_t E = B * kP, ntiles = (E + kTcT - 1) / kTcT;
  int per_sm = 512 / tmem_alloc_cols(4 * kBwdKC + 2 * NP);
  if (per_sm > 4) per_sm = 4;
  // COSKAD_BWD_ASYNC=0 selects the register-load kernel (A/B measurement); default: cp.async staging one pass ahead.  The staging
  // buffer (24 KB) may cost a resident CTA where the weight images are large (64 -> 64 channels: 3 instead of 4 per SM).
  static const bool async = [] { const char* e = getenv("COSKAD_BWD_ASYNC"); return e == nullptr || atoi(e) != 0; }();
  if (async) {
    cudaFuncAttributes attr{};
    CK(cudaFuncGetAttributes(&attr, tc_mix_bwd_data_kernel<CO, COK, NP, true>));
    const size_t stage = sizeof(float) * kBwdStageFloats;
    while (per_sm > 1 && (attr.sharedSizeBytes + stage + 1024) * per_sm > static_cast<size_t>(227) * 1024) --per_sm;
    const size_t target = (static_cast<size_t>(228) * 1024) / static_cast<size_t>(per_sm) - 1536;
    const size_t dyn = target > attr.sharedSizeBytes + stage ? target - attr.sharedSizeBytes : stage;
    const int64_t cap = static_cast<int64_t>(ctx->sm_count) * per_sm;
    const int g = static_cast<int>(ntiles < cap ? ntiles : cap);
    CK(cudaFuncSetAttribute(tc_mix_bwd_data_kernel<CO, COK, NP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(dyn)));
    tc_mix_bwd_data_kernel<CO, COK, NP, true><<<g, kTcT, dyn, st>>>(dout, y1, y2, mi, g1, be1, g2, be2, slope, red, W1, W2, E, CI, dy1,
                                                                    dy2, dG, dXres);
    CK_LAUNCH();
    return COSKAD_OK;
  }
  const int64_t cap = static_cast<int64_t>(ctx->sm_count) * per_sm;
  const int g = static_cast<int>(ntiles < cap ? ntiles : cap);
  const size_t pad = tmem_pad_smem(tc_mix_bwd_data_kernel<CO, COK, NP, false>, per_sm);
  CK(cudaFuncSetAttribute(tc_mix_bwd_data_kernel<CO, COK, NP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(pad)));
  tc_mix_bwd_data_kernel<CO, COK, NP, false><<<g, kTcT, pad, st>>>(dout, y1, y2, mi, g1, be1, g2, be2, slope, red, W1, W2, E, CI, dy1, dy2,
                                                               dG, dXres);
  CK_LAUNCH();
  return COSKAD_OK;
}
template <int CO, int CI8, int KT>
static int launch_tc_bwd_weight(coskad_ctx* ctx, const float* dy1, const float* dy2, const float* G, const float* X, int64_t B,
                                int CI, float* dW1, float* db1, float* dW2, float* db2, cudaStream_t st) {
  const int64_t E = B * kP, ntiles = (E + KT - 1) / KT;
  const size_t smem = sizeof(float) * wg_smem_floats(CO, CI8, KT);
  int per_sm = static_cast<int>((200 * 1024) / (smem + 1024));
  const int tm = 512 / tmem_alloc_cols(2 * wg_n(CI8));
  if (per_sm > tm) per_sm = tm;
  if (per_sm > 3) per_sm = 3;
  if (per_sm < 1) per_sm = 1;
  const int64_t cap = static_cast<int64_t>(ctx->sm_count) * per_sm;
  const int g = static_cast<int>(ntiles < cap ? ntiles : cap);
  const size_t per = static_cast<size_t>(2) * CO * (CI + 1);
  { const int rc = ensure_ws(ctx, sizeof(float) * per * g); if (rc) return rc; }
  CK(cudaFuncSetAttribute(tc_mix_bwd_weight_kernel<CO, CI8, KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  tc_mix_bwd_weight_kernel<CO, CI8, KT><<<g, kTcT, smem, st>>>(dy1, dy2, G, X, E, CI, ctx->ws);
  CK_LAUNCH();
  tc_wgrad_reduce_kernel<<<static_cast<unsigned>((per + 31) / 32), dim3(32, kPsRows), 0, st>>>(ctx->ws, g, CO, CI, dW1, db1, dW2, db2);
  CK_LAUNCH();
  return COSKAD_OK;
}

// Tensor-core backward of the two 1x1 convolutions of a layer with the BatchNorm-train + PReLU backward fused in:
// (dout, y1, y2, red) -> dy1, dy2 (scratch, written once) -> dG, dXres, dW1 += , db1 +=, dW2 +=, db2 +=
extern "C" int coskad_train_mix_bwd_tc(coskad_ctx* ctx, const float* dout, const float* y1, const float* y2, const float* mi,
                                       const float* g1, const float* be1, const float* g2, const float* be2,
                                       const float* slope, const double* red, const float* G, const float* X, const float* W1,
                                       const float* W2, int64_t B, int CI, int CO, float* dy1, float* dy2, float* dG,
                                       float* dXres, float* dW1, float* db1, float* dW2, float* db2, void* stream_) {
  TRAIN_PRE();
  if (!chan_ok(CI) || !chan_ok(CO)) return fail(ctx, COSKAD_ERR_ARG, "training kernels support channels {2,16,32,64}, got %d->%d", CI, CO);
  if (B <= 0) return COSKAD_OK;
  if (!dout || !y1 || !y2 || !mi || !red || !G || !X || !W1 || !W2 || !dy1 || !dy2 || !dG || !dXres || !dW1 || !dW2)
    return fail(ctx, COSKAD_ERR_ARG, "NULL pointer");
  int rc = COSKAD_ERR_ARG;
#define TC_BWD(ci, co, cok, np, ci8, kt)                                                                                       \
  if (CI == ci && CO == co) {                                                                                                \
    rc = launch_tc_bwd_data<co, cok, np>(ctx, dout, y1, y2, mi, g1, be1, g2, be2, slope, red, W1, W2, B, CI, dy1, dy2, dG, dXres, st); \
    if (rc) return rc;                                                                                                       \
    return launch_tc_bwd_weight<co, ci8, kt>(ctx, dy1, dy2, G, X, B, CI, dW1, db1, dW2, db2, st);                                \
  }
  TC_BWD(2, 32, 32, 16, 8, 64) TC_BWD(32, 16, 16, 32, 32, 64) TC_BWD(16, 32, 32, 16, 16, 64) TC_BWD(32, 64, 64, 32, 32, 32)   // encoder
  TC_BWD(64, 32, 32, 64, 64, 32) TC_BWD(32, 2, 16, 32, 32, 64)                                                          // decoder
#undef TC_BWD
  return fail(ctx, COSKAD_ERR_ARG, "train_mix_bwd_tc: unsupported channel pair %d -> %d", CI, CO);
}

extern "C" int coskad_train_mix_bwd(coskad_ctx* ctx, const float* dy1, const float* dy2, const float* G, const float* X,
                                    const float* W1, const float* W2, int64_t B, int CI, int CO, float* dG, float* dXres,
                                    float* dW1, float* db1, float* dW2, float* db2, void* stream_) {
  TRAIN_PRE();
  if (!chan_ok(CI) || !chan_ok(CO)) return fail(ctx, COSKAD_ERR_ARG, "training kernels support channels {2,16,32,64}, got %d->%d", CI, CO);
  if (B <= 0) return COSKAD_OK;
  {
    // dG[ci] = sum_co W1[co,ci] dy1[co], dXres[ci] = sum_co W2[co,ci] dy2[co]: K = CO, M = CI, weights already [k][m]
    const int rc = launch_chan_gemm(ctx, false, dy1, dy2, W1, W2, 1, nullptr, nullptr, B, CO, CI, dG, dXres, nullptr, st);
    if (rc) return rc;
  }
  const int tco = CO < 4 ? CO : 4, tci = CI < 4 ? CI : 4;
  const int ntile = (CO / tco) * (CI / tci);
  if ((CO >= 4 && CO % 4) || (CI >= 4 && CI % 4) || ntile > 128 || ntile < 8)
    return fail(ctx, COSKAD_ERR_ARG, "train_mix_bwd: unsupported channel pair %d -> %d", CI, CO);
  // two staging buffers of [2 (CO + CI)][WC + 4] floats; >= 32 KB for the final reduction over the k-slices
  const int wc = (CO + CI) > 48 ? 64 : 128;
  if ((ntile & (ntile - 1)) || (wc * ntile / kTrainThreads) % 4)
    return fail(ctx, COSKAD_ERR_ARG, "train_mix_bwd: unsupported tile split for %d -> %d", CI, CO);
  size_t smem2 = sizeof(float) * 2 * 2 * (CO + CI) * (wc + 4);
  if (smem2 < sizeof(float) * 32 * kTrainThreads) smem2 = sizeof(float) * 32 * kTrainThreads;
  int g2 = static_cast<int>((B * kP + wc - 1) / wc);
  if (g2 > ctx->sm_count * 2) g2 = ctx->sm_count * 2;
  if (wc == 64) {
    CK(cudaFuncSetAttribute(train_mix_bwd_weight_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem2)));
    train_mix_bwd_weight_kernel<64><<<g2, kTrainThreads, smem2, st>>>(dy1, dy2, G, X, B, CI, CO, dW1, db1, dW2, db2);
  } else {
    CK(cudaFuncSetAttribute(train_mix_bwd_weight_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem2)));
    train_mix_bwd_weight_kernel<128><<<g2, kTrainThreads, smem2, st>>>(dy1, dy2, G, X, B, CI, CO, dW1, db1, dW2, db2);
  }
  CK_LAUNCH();
  return COSKAD_OK;
}

// mode 0: out[b,d] = sum_f A[b,f] W(d,f) + bias[d]; mode 1: out[b,f] = sum_d a[b,d] W(d,f) + bias[f];
// mode 2: dW(d,f) += sum_b a[b,d] A[b,f].  w_is_fd: W stored [F,D] (rev_btlnk) instead of [D,F] (btlnk).
extern "C" int coskad_train_linear(coskad_ctx* ctx, int mode, const float* a_small, const float* A_wide, const float* W,
                                   int w_is_fd, const float* bias, int64_t B, int F, int D, float* out, void* stream_) {
  TRAIN_PRE();
  if (D < 1 || D > 16) return fail(ctx, COSKAD_ERR_ARG, "linear: D must be in [1,16], got %d", D);
  if (F % 4) return fail(ctx, COSKAD_ERR_ARG, "linear: F must be a multiple of 4 (16-byte loads), got %d", F);
  if (B <= 0) return COSKAD_OK;
  const int64_t sd = w_is_fd ? 1 : F, sf = w_is_fd ? D : 1;
  if (mode == 0) {
    // the feature slices write partial sums [kLinSlices][B][D]; the second stage adds them in a fixed order
    const size_t n = static_cast<size_t>(B) * D;
    { const int rc = ensure_ws(ctx, sizeof(float) * n * kLinSlices); if (rc) return rc; }
    CK(cudaMemsetAsync(out, 0, sizeof(float) * n, st));
    const int rows_per_block = (kTrainThreads / 32) * kLinRows;
    lin_reduce_f_kernel<16><<<dim3(static_cast<unsigned>((B + rows_per_block - 1) / rows_per_block), kLinSlices), kTrainThreads, 0, st>>>(
        A_wide, W, sd, sf, bias, B, F, D, ctx->ws);
    CK_LAUNCH();
    return launch_partial_sum<float>(ctx, ctx->ws, kLinSlices, static_cast<int64_t>(n), static_cast<int64_t>(n), out, st);
  }
  else if (mode == 1) lin_expand_f_kernel<16><<<dim3((F + kTrainThreads - 1) / kTrainThreads, static_cast<unsigned>((B + kExpRows - 1) / kExpRows)), kTrainThreads, 0, st>>>(a_small, W, sd, sf, bias, B, F, D, out);
  else if (mode == 2) {
    // row slices: >= 16 rows per warp, and as many blocks as fill whole waves of the 2 resident blocks per SM (102 feature blocks
    // x 4 slices = 408 blocks were 1.38 waves on 296 slots: the second wave ran at 38 % occupancy)
    int nb = B >= 512 ? 4 : 1;
    if (B >= 1024) {
      static const int env_nb = [] { const char* e = getenv("COSKAD_WGRAD_SLICES"); return e ? atoi(e) : 0; }();
      const int fb = (F + kWgF - 1) / kWgF, slots = 2 * ctx->sm_count;
      double best = 0.0;
      for (int c = 2; c <= 8 && B / c >= 128; ++c) {
        const int blocks = fb * c, waves = (blocks + slots - 1) / slots;
        const double eff = static_cast<double>(blocks) / (static_cast<double>(waves) * slots);
        if (eff > best + 1e-9) { best = eff; nb = c; }
      }
      if (env_nb > 0) nb = env_nb;
    }
    const size_t n = static_cast<size_t>(D) * F;
    { const int rc = ensure_ws(ctx, sizeof(float) * n * nb); if (rc) return rc; }
    lin_wgrad_kernel<16><<<dim3((F + kWgF - 1) / kWgF, nb), kTrainThreads, 0, st>>>(a_small, A_wide, sd, sf, B, F, D, ctx->ws);
    CK_LAUNCH();
    return launch_partial_sum<float>(ctx, ctx->ws, nb, static_cast<int64_t>(n), static_cast<int64_t>(n), out, st);
  } else return fail(ctx, COSKAD_ERR_ARG, "linear: unknown mode %d", mode);
  CK_LAUNCH();
  return COSKAD_OK;
}

extern "C" int coskad_adam_step(coskad_ctx* ctx, float* p, const float* g, float* m, float* v, int64_t n, const float* lr,
                                float beta1, float beta2, float eps, int64_t* step, float* scratch, void* stream_) {
  TRAIN_PRE();
  if (n <= 0) return COSKAD_OK;
  if (!p || !g || !m || !v || !lr || !step || !scratch) return fail(ctx, COSKAD_ERR_ARG, "coskad_adam_step: NULL pointer");
  if (n % 4) return fail(ctx, COSKAD_ERR_ARG, "coskad_adam_step: n must be a multiple of 4 (pad the flat buffers), got %lld", (long long)n);
  if ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15)
    return fail(ctx, COSKAD_ERR_ARG, "coskad_adam_step: buffers must be 16-byte aligned");
  adam_tick_kernel<<<1, 1, 0, st>>>(step, scratch, beta1, beta2);
  CK_LAUNCH();
  const int64_t n4 = n / 4;
  const int blocks = static_cast<int>((n4 + 255) / 256 < 2 * ctx->sm_count ? (n4 + 255) / 256 : 2 * ctx->sm_count);
  adam_flat_kernel<<<blocks, 256, 0, st>>>(p, g, m, v, n4, lr, scratch, beta1, beta2, eps);
  CK_LAUNCH();
  return COSKAD_OK;
}

extern "C" int coskad_train_col_sum(coskad_ctx* ctx, const float* a, int64_t B, int N, float* out, void* stream_) {
  TRAIN_PRE();
  if (B <= 0 || N <= 0) return COSKAD_OK;
  const int64_t slices = (B + 63) / 64;               // >= 8 rows per thread
  const int ns = static_cast<int>(slices < 32 ? slices : 32);
  { const int rc = ensure_ws(ctx, sizeof(float) * static_cast<size_t>(N) * ns); if (rc) return rc; }
  col_sum_kernel<<<dim3((N + 31) / 32, static_cast<unsigned>(ns)), kTrainThreads, 0, st>>>(a, B, N, ctx->ws);
  CK_LAUNCH();
  return launch_partial_sum<float>(ctx, ctx->ws, ns, N, N, out, st);
}
