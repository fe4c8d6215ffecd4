/* coskad_b200.h -- C ABI of libcoskad_b200.so (hand-written sm_100a CUDA kernels for the
 * COSKAD anomaly-scoring hot path).
 *
 * The reference (aleflabo/COSKAD) is pure Python: it has no FFI layer.  Each entry point below
 * therefore replaces a *Python call site* of the reference; the citation after "replaces:" is
 * the reference file:line whose arithmetic the entry point reproduces.  The Python binding a
 * maintainer would add is a ctypes stub -- see INTEGRATION.md and coskad_b200/_lib.py.
 *
 * Conventions
 *   - every function returns 0 (COSKAD_OK) or a negative coskad_status; it never throws/exits;
 *     coskad_last_error(ctx) gives a human readable message for the last failure on that ctx.
 *   - all data pointers are DEVICE pointers on the ctx's device unless the name ends in _host;
 *     the caller owns every buffer; the library owns only the ctx (packed weights, scratch).
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  Calls are
 *     asynchronous with respect to the host and ordered on that stream.
 *   - a ctx is not thread-safe: one ctx per (device, host thread).
 *   - layouts are the reference's: windows x[B, C, T, V] float32 contiguous (NCHW with
 *     C = coordinates, T = frames, V = joints), latents z[B, D] float32, scores [B] float32.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef COSKAD_B200_H_
#define COSKAD_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COSKAD_ABI_VERSION 2

typedef struct coskad_ctx coskad_ctx;

typedef enum {
  COSKAD_OK = 0,
  COSKAD_ERR_ARG = -1,        /* bad argument / unsupported shape */
  COSKAD_ERR_CUDA = -2,       /* CUDA runtime error (message in coskad_last_error) */
  COSKAD_ERR_STATE = -3,      /* weights not set, wrong call order */
  COSKAD_ERR_NO_DEVICE = -4   /* no usable sm_100 device */
} coskad_status;

/* Latent geometry / score flavour. */
typedef enum {
  COSKAD_SCORE_NONE = 0,        /* no score, latent only                                        */
  COSKAD_SCORE_POINCARE = 1,    /* project(expmap0(z)) then dist(x, c), geoopt 0.5.0 constants,
                                   k=-1.  replaces: eval_COSKAD.py:194-196 + utils/eval_utils.py:66-67 */
  COSKAD_SCORE_POINCARE_NOPROJ = 2, /* expmap0 only (validation path) replaces: models/hyperbolic_encoder.py:266 */
  COSKAD_SCORE_EUCLID = 3,      /* mean_d (c_d - z_d)^2        replaces: utils/eval_utils.py:61-64   */
  COSKAD_SCORE_COSINE = 4,      /* 1 - cos(c, z)               replaces: eval_COSKAD.py:81           */
  COSKAD_SCORE_POINCARE_HM = 5  /* utils/hyper_math.py constants (c=+1): expmap0 :302-306,
                                   project :100-105, dist :207-210 (dead code upstream; pinned oracle) */
} coskad_flavour;

/* One ST_GCNN_layer's parameters (device pointers, float32), reference names in comments.
 * replaces: models/graph_layers/stsgcn.py:47-91 (module state) */
typedef struct {
  int32_t c_in, c_out;
  const float* A;        /* gcn.A            [T, V, V]                 stsgcn.py:134 */
  const float* T;        /* gcn.T            [V, T, T]                 stsgcn.py:138 */
  const float* w1;       /* tcn.0.weight     [c_out, c_in] (1x1 conv)  stsgcn.py:57  */
  const float* b1;       /* tcn.0.bias       [c_out] or NULL                         */
  const float* bn1_w;    /* tcn.1.weight / bias / running_mean / running_var [c_out] */
  const float* bn1_b;
  const float* bn1_rm;
  const float* bn1_rv;
  const float* w2;       /* residual.0.weight [c_out, c_in]; NULL => nn.Identity (stsgcn.py:80) */
  const float* b2;       /* residual.0.bias or NULL                                  */
  const float* bn2_w;    /* residual.1.*                                             */
  const float* bn2_b;
  const float* bn2_rm;
  const float* bn2_rv;
  const float* prelu;    /* prelu.weight [1]                            stsgcn.py:82 */
} coskad_layer_params;

/* ---- lifetime ---------------------------------------------------------------------------- */
int coskad_abi_version(void);
/* n_frames / n_joints: the fused sm_100a kernels are built for the shapes every reference
 * config uses (12 frames, 17 joints); other shapes are rejected with COSKAD_ERR_ARG. */
int coskad_create(coskad_ctx** out, int device, int n_frames, int n_joints);
int coskad_destroy(coskad_ctx* ctx);
const char* coskad_last_error(const coskad_ctx* ctx);   /* ctx may be NULL: last create() error */

/* ---- weights ------------------------------------------------------------------------------
 * Encoder = n_layers ST_GCNN layers + a linear head of `head_rows` rows over the flattened
 * (c,t,v) features.  For STSE/STSAE the head is btlnk (head_rows = latent_dim); for STSVAE it is
 * fc_mean stacked on fc_var (head_rows = latent_dim + 1).  Eval-mode BatchNorm is folded into
 * the 1x1 convolutions on the device, in float64, at this call.
 * replaces: models/sts/ae.py:60-72,147-157 (module construction) + nn.Module.eval() semantics */
int coskad_set_encoder(coskad_ctx* ctx, int n_layers, const coskad_layer_params* layers_host,
                       const float* head_w /*[head_rows, F]*/, const float* head_b /*[head_rows] or NULL*/,
                       int head_rows, void* stream);
/* Decoder = rev_btlnk Linear(latent -> F) + n_layers ST_GCNN layers.
 * replaces: models/sts/ae.py:200-230 */
int coskad_set_decoder(coskad_ctx* ctx, const float* rev_w /*[F, latent]*/, const float* rev_b /*[F]*/,
                       int latent_dim, int n_layers, const coskad_layer_params* layers_host, void* stream);

/* ---- fused eval hot path ------------------------------------------------------------------
 * x[B,2,12,17] -> 4 fused ST_GCNN layers -> head -> geometry -> score, one persistent kernel;
 * activations never leave the SM.  z (nullable) receives the raw head output [B, head_rows];
 * score (nullable iff flavour NONE) receives [B].  center: [latent] device pointer.
 * replaces: STSE.forward models/sts/ae.py:108-121 (+ the per-person scoring in eval_COSKAD.py:186-199) */
int coskad_encode_score_fwd(coskad_ctx* ctx, int flavour, const float* x, const float* center,
                            int64_t B, float* z, float* score, void* stream);
/* Same path fed from TRAJECTORIES: window construction and the test-time affine transforms happen on the device, in the
 * kernel's input stage, so the 12x (stride-1 sliding windows) x n_transform redundant window tensor is never materialised.
 *   traj [traj_rows, 2*V] float32: the reference's per-person coordinate rows (x0,y0,x1,y1,.. per frame, after scaling;
 *        trajectory.coordinates), any number of persons concatenated;
 *   win_row [N] int64: window i = the 12 consecutive rows starting at win_row[i] (x[c][t][v] = traj[win_row[i]+t][2v+c]),
 *        clamped to [0, traj_rows-12];
 *   trans [N] int32 (nullable) indexes mats [n_mats][6] = rows 0,1 of the 3x3 affine matrices (get_aff_trans_mat):
 *        x' = m0 x + m1 y + m2,  y' = m3 x + m4 y + m5.
 * replaces: the window materialisation utils/preprocessing.py:58-89 (_aggregate_rnn_autoencoder_data) and
 * PoseDatasetRobust.__getitem__ utils/dataset.py:65-74 -> apply_pose_transform utils/dataset_utils.py:270-284,
 * feeding STSE.forward models/sts/ae.py:108-121 */
int coskad_encode_score_traj_fwd(coskad_ctx* ctx, int flavour, const float* traj, int64_t traj_rows,
                                 const int64_t* win_row, const int32_t* trans, const float* mats /*[n_mats,6]*/, int n_mats,
                                 const float* center, int64_t N, float* z, float* score, void* stream);
/* Kernel generation used by coskad_encode_score_fwd: 1 (default) = channel mixing on tcgen05 tensor cores (3xTF32),
 * 0 = the all-FP32 CUDA-core kernel (kept for A/B measurement and as the decoder path). */
int coskad_set_fused_impl(coskad_ctx* ctx, int impl);
/* Euclidean auto-encoder: also runs the decoder and the reconstruction score.
 * xhat (nullable) [B,2,12,17]; rec_score (nullable) [B] = mean_{c,t,v}(x - xhat)^2;
 * lat_score (nullable) [B] = mean_d (c_d - z_d)^2.
 * replaces: STSAE.forward models/sts/ae.py:233-250 + utils/eval_utils.py:77-90 */
int coskad_autoencode_score_fwd(coskad_ctx* ctx, const float* x, const float* center, int64_t B,
                                float* z, float* xhat, float* rec_score, float* lat_score, void* stream);

/* ---- geometry on latents (the gmath namespace) -------------------------------------------
 * op codes for coskad_geom_map: z[B,D] -> out[B,D] */
typedef enum {
  COSKAD_MAP_EXPMAP0 = 0,          /* gmath.expmap0(u, k=-1)                  */
  COSKAD_MAP_PROJECT = 1,          /* gmath.project(x, k=-1)                  */
  COSKAD_MAP_EXPMAP0_PROJECT = 2,  /* project(expmap0(u))                     */
  COSKAD_MAP_EXPMAP0_HM = 3,       /* utils/hyper_math.py expmap0 (c=1)       */
  COSKAD_MAP_PROJECT_HM = 4,       /* utils/hyper_math.py project (c=1)       */
  COSKAD_MAP_L2NORMALIZE = 5       /* z / ||z||   (models/sts/vae.py:81)      */
} coskad_map_op;
int coskad_geom_map(coskad_ctx* ctx, int op, const float* in, int64_t B, int D, float* out, void* stream);
/* Pairwise-broadcast distance/score between a[B,D] and b (b_is_broadcast: b is [D], else [B,D]).
 * flavour POINCARE: gmath.dist(a, b, k=-1); POINCARE_HM: hyper_math.dist; EUCLID: mean (b-a)^2;
 * COSINE: 1 - cos(b, a).  replaces: geoopt stereographic math dist / utils/eval_utils.py:61-67 */
int coskad_dist(coskad_ctx* ctx, int flavour, const float* a, const float* b, int b_is_broadcast,
                int64_t B, int D, float* out, void* stream);
/* Vector-Jacobian products of the two calls above, so that the reference's own training_step can differentiate
 * through gmath.expmap0 / project / dist one call at a time.  gin[B,D] = (d map(in)/d in)^T gout; ops EXPMAP0, PROJECT,
 * EXPMAP0_PROJECT, L2NORMALIZE.  ga / gb [B,D] (either nullable) = gs[B] * d f(a,b)/d a, d b per row; flavours POINCARE,
 * EUCLID, COSINE.   replaces: autograd through geoopt at models/hyperbolic_encoder.py:147,157 */
int coskad_geom_map_bwd(coskad_ctx* ctx, int op, const float* in, const float* gout, int64_t B, int D, float* gin, void* stream);
int coskad_dist_bwd(coskad_ctx* ctx, int flavour, const float* a, const float* b, int b_is_broadcast, const float* gs,
                    int64_t B, int D, float* ga, float* gb, void* stream);
/* dist0(x) = 2 artanh(||x||).   replaces: models/hyperbolic_encoder.py:181 */
int coskad_dist0(coskad_ctx* ctx, const float* x, int64_t B, int D, float* out, void* stream);
/* PowerSpherical reparameterised sample from explicit noise: z = Householder_{e1->mu}([t, sqrt(1-t^2) v]);
 * mu [B,D] unit vectors, t [B] = 2 Beta(alpha,beta) - 1, v [B,D-1] unit vectors.
 * replaces: PowerSpherical(loc, scale).rsample() models/sts/vae.py:110,129 (power_spherical, un-vendored) */
int coskad_ps_sample(coskad_ctx* ctx, const float* mu, const float* t, const float* v, int64_t B, int D, float* z, void* stream);
/* backward of score = dist(project?(expmap0(z)), c) w.r.t. z (training loss, hyperbolic_encoder.py:147-157),
 * arg order dist(c, x) as in training. dz[B,D] = dscore[B] * d score/d z. */
int coskad_poincare_score_bwd(coskad_ctx* ctx, const float* z, const float* center, const float* dscore,
                              int64_t B, int D, int with_project, float* dz, void* stream);

/* ---- center update -------------------------------------------------------------------------
 * acc is [D+2] float64 on the device and is ACCUMULATED into (zero it first):
 *   POINCARE: acc[0..D) += sum gamma_i x_i, acc[D] += sum (gamma_i - 1), acc[D+1] += count
 *   EUCLID / COSINE: acc[0..D) += sum z_i, acc[D+1] += count
 * Partial sums from several shards/ranks add (all-reduce them), then finalize once.
 * replaces: gmath.weighted_midpoint models/hyperbolic_encoder.py:122,179; the running sums of
 * models/euclidean_encoder_staticCenter.py:105-124; models/spherical_vae.py:110-116 */
int coskad_center_partial(coskad_ctx* ctx, int flavour, const float* zproj, int64_t B, int D, double* acc, void* stream);
/* eps: center_tolerance clamp of the Euclidean flavour (<=0 disables). */
int coskad_center_finalize(coskad_ctx* ctx, int flavour, const double* acc, int D, float eps, float* center, void* stream);

/* ---- Mahalanobis distance to the center (distance: 'mahalanobis'; D <= 32) -------------------
 * out[b] = sqrt((z_b - c)^T VI (z_b - c)), VI [D, D] row-major (the inverse covariance matrix).
 * replaces: utils/eval_utils.py:28-38 mahalanobis(u, v, VI, reduce='none') as called by
 * windows_based_loss_mahalanobis (:41-55) and models/euclidean_encoder_staticCenter.py:185 */
int coskad_mahalanobis(coskad_ctx* ctx, const float* z, const float* center, const float* VI, int64_t B, int D,
                       float* out, void* stream);
/* gz[b, :] = gs[b] (VI + VI^T)(z_b - c) / (2 out[b])  -- autograd of the above w.r.t. z (training loss,
 * models/euclidean_encoder_staticCenter.py:185); rows with a zero distance get a zero gradient */
int coskad_mahalanobis_bwd(coskad_ctx* ctx, const float* z, const float* center, const float* VI, const float* gs,
                           int64_t B, int D, float* gz, void* stream);
/* acc (double)[D*D + 1] += sum_b (z_b - mu)(z_b - mu)^T (row-major) and the row count in acc[D*D]; partial sums of
 * several batches / shards add (all-reduce them), then cov = acc / (count - 1) and VI = inverse(cov) on the host side.
 * replaces: batch_cov_mat_step + compute_inv_cov_mat, models/euclidean_encoder_staticCenter.py:40-46,133-142 */
int coskad_cov_partial(coskad_ctx* ctx, const float* z, const float* mu, int64_t B, int D, double* acc, void* stream);

/* ---- frame-level aggregation ---------------------------------------------------------------
 * Bit-exact replacement of the Python triple loop: for each person scatter score -> frames
 * (index frames-1 with the reference's wrap of frame id 0 -> last frame), treat exact zeros as
 * absent, float64 mean in window (dataset) order, then max over the clip's persons.
 *   score[N] f32; frames[N,T] i64 (1-based frame ids);
 *   win_idx[*] i64  window ids grouped by person, dataset order inside a person;
 *   person_off[n_persons+1] i64  CSR offsets into win_idx;
 *   person_clip[n_persons] i32   clip of each person; persons of one clip are contiguous;
 *   person_out_off[n_persons+1] i64  offsets of each person's curve in person_out
 *                                (person p owns n_frames(clip(p)) doubles);
 *   clip_person_off[n_clips+1] i64   CSR offsets: persons of clip c;
 *   clip_off[n_clips+1] i64      clip c owns out[clip_off[c] .. clip_off[c+1]), n_frames(c) long;
 *   total_person_frames = person_out_off[n_persons]; max_clip_frames = max_c n_frames(c);
 *   person_out f64 (per-person curves, always written), out f64 (per-clip per-frame scores).
 * A clip without persons yields zeros (the reference raises on np.stack of an empty list).
 * replaces: utils/eval_utils.py:69-74 + eval_COSKAD.py:201-211 */
int coskad_frame_aggregate(coskad_ctx* ctx, const float* score, const int64_t* frames, int T,
                           const int64_t* win_idx, const int64_t* person_off, const int32_t* person_clip,
                           const int64_t* person_out_off, int64_t n_persons,
                           const int64_t* clip_person_off, const int64_t* clip_off, int64_t n_clips,
                           int64_t total_person_frames, int64_t max_clip_frames,
                           double* person_out, double* out, void* stream);

/* Score post-processing of per-clip curves on the device: out[f] = gaussian_filter1d(shifted, sigma)[f] with
 * shifted[shift:] = curve[:-shift]; float64, scipy's 'reflect' boundary and accumulation order.
 *   curves f64 concatenated, curve_off[n_curves+1] i64 CSR offsets, weights[2*radius+1] f64 = the normalised Gaussian
 *   kernel exactly as scipy builds it (radius = int(4 sigma + 0.5); the caller computes it with numpy), out != curves.
 * replaces: utils/eval_utils.py:200-207 score_process (eval_COSKAD.py:217) */
int coskad_score_process(coskad_ctx* ctx, const double* curves, const int64_t* curve_off, int64_t n_curves, int shift,
                         const double* weights, int radius, double* out, void* stream);

/* ---- training path (per-layer kernels, train-mode BatchNorm with per-GPU batch statistics) ------
 * Activations are [B, C, 204] float32 in HBM between kernels (the batch statistics of BatchNorm sit
 * between the convolution and the activation).  Gradient buffers that are ACCUMULATED into (dA, dT,
 * dW*, db*, stats, red, linear mode 2 / col_sum outputs) must be zeroed by the caller.
 * replaces: ST_GCNN_layer.forward under autograd, models/graph_layers/stsgcn.py:94-156 */
/* G1 = einsum('nctv,vtq->ncqv', X, T); G = einsum('nctv,tvw->nctw', G1, A); R = B*C rows  (stsgcn.py:154-155) */
int coskad_train_contract_fwd(coskad_ctx* ctx, const float* X, const float* A, const float* T, int64_t R,
                              float* G1, float* G, void* stream);
/* dX = dXres (nullable) + T^T A^T dG;  dA += G1 (x) dG;  dT += X (x) dG1 */
int coskad_train_contract_bwd(coskad_ctx* ctx, const float* dG, const float* dXres, const float* X, const float* G1,
                              const float* A, const float* T, int64_t R, float* dX, float* dA, float* dT, void* stream);
/* y1 = conv1x1(G; W1,b1), y2 = conv1x1(X; W2,b2)  (stsgcn.py:57,71); stats (double)[4*CO] += sum y1, sum y1^2, sum y2, sum y2^2 */
int coskad_train_mix_fwd(coskad_ctx* ctx, const float* G, const float* X, const float* W1, const float* b1,
                         const float* W2, const float* b2, int64_t B, int CI, int CO, float* y1, float* y2,
                         double* stats, void* stream);
/* mi[4*CO] = mean1, invstd1, mean2, invstd2 (biased variance, eps); running stats updated like nn.BatchNorm2d
 * (momentum, unbiased variance); n_per_channel = B*204; nbt1 / nbt2 (nullable): the two num_batches_tracked counters
 * (int64 device scalars), incremented by one.  replaces: nn.BatchNorm2d(train) forward bookkeeping, stsgcn.py:62,79 */
int coskad_train_bn_finalize(coskad_ctx* ctx, const double* stats, int64_t n_per_channel, int CO, float eps,
                             float momentum, float* rm1, float* rv1, float* rm2, float* rv2, float* mi, int64_t* nbt1,
                             int64_t* nbt2, void* stream);
/* coskad_train_mix_fwd + coskad_train_bn_finalize for the training forward on the tensor-core path: the statistics' second stage and
 * the BatchNorm finalize are one launch and no intermediate `stats` buffer exists (the values are the same).
 * replaces: nn.Conv2d 1x1 + nn.BatchNorm2d(train) statistics of both branches, models/graph_layers/stsgcn.py:56-62,74-79 */
int coskad_train_mix_fwd_bn(coskad_ctx* ctx, const float* G, const float* X, const float* W1, const float* b1, const float* W2,
                            const float* b2, int64_t B, int CI, int CO, float* y1, float* y2, float eps, float momentum,
                            float* rm1, float* rv1, float* rm2, float* rv2, float* mi, int64_t* nbt1, int64_t* nbt2,
                            void* stream);
/* coskad_train_bn_prelu_bwd (dy1 = dy2 = NULL form) + coskad_train_bn_param_grads in two launches instead of three: `red`
 * (double)[3*CO+1] is ASSIGNED (no zero fill needed), the parameter gradients are accumulated (each nullable). */
int coskad_train_bn_prelu_bwd_grads(coskad_ctx* ctx, const float* dout, const float* y1, const float* y2, const float* mi,
                                    const float* g1, const float* be1, const float* g2, const float* be2, const float* slope,
                                    int64_t B, int CO, double* red, float* dg1, float* dbe1, float* dg2, float* dbe2,
                                    float* dslope, void* stream);
/* parameter gradients of the layer's two BatchNorms and its PReLU from `red` (coskad_train_bn_prelu_bwd), ACCUMULATED into
 * dg1 / dbe1 / dg2 / dbe2 [CO] and dslope [1] (each nullable): d beta1 = d beta2 = red[0..CO), d gamma1 = red[CO..2CO),
 * d gamma2 = red[2CO..3CO), d slope = red[3CO].  replaces: autograd of nn.BatchNorm2d / nn.PReLU parameters, stsgcn.py:62,79,82 */
int coskad_train_bn_param_grads(coskad_ctx* ctx, const double* red, int CO, float* dg1, float* dbe1, float* dg2,
                                float* dbe2, float* dslope, void* stream);
/* out = PReLU(BN1(y1) + BN2(y2))   (stsgcn.py:106-110) */
int coskad_train_bn_prelu_fwd(coskad_ctx* ctx, const float* y1, const float* y2, const float* mi, const float* g1,
                              const float* be1, const float* g2, const float* be2, const float* slope, int64_t B, int CO,
                              float* out, void* stream);
/* red (double)[3*CO+1] += sum ds, sum ds*yhat1, sum ds*yhat2 per channel (= d beta, d gamma1, d gamma2) and d slope;
 * dy1, dy2 = gradients w.r.t. the conv outputs (both NULL: only `red` is reduced) */
int coskad_train_bn_prelu_bwd(coskad_ctx* ctx, const float* dout, const float* y1, const float* y2, const float* mi,
                              const float* g1, const float* be1, const float* g2, const float* be2, const float* slope,
                              int64_t B, int CO, double* red, float* dy1, float* dy2, void* stream);
/* dG = W1^T dy1, dXres = W2^T dy2; dW1 += dy1 G^T, db1 += sum dy1, dW2 += dy2 X^T, db2 += sum dy2 */
int coskad_train_mix_bwd(coskad_ctx* ctx, const float* dy1, const float* dy2, const float* G, const float* X,
                         const float* W1, const float* W2, int64_t B, int CI, int CO, float* dG, float* dXres,
                         float* dW1, float* db1, float* dW2, float* db2, void* stream);
/* The same backward on the tensor cores (tcgen05, 3xTF32), with the BatchNorm-train + PReLU backward fused in: call
 * coskad_train_bn_prelu_bwd with dy1 = dy2 = NULL first (it then only reduces `red`), then this.  dy1 / dy2 are scratch
 * [B, CO, 204] (written once, read by the weight-gradient kernel); dW*, db* are accumulated into (zero them first).
 * replaces: autograd of nn.Conv2d 1x1 + nn.BatchNorm2d(train) + nn.PReLU, models/graph_layers/stsgcn.py:56-82,106-110 */
int coskad_train_mix_bwd_tc(coskad_ctx* ctx, const float* dout, const float* y1, const float* y2, const float* mi,
                            const float* g1, const float* be1, const float* g2, const float* be2, const float* slope,
                            const double* red, const float* G, const float* X, const float* W1, const float* W2, int64_t B,
                            int CI, int CO, float* dy1, float* dy2, float* dG, float* dXres, float* dW1, float* db1,
                            float* dW2, float* db2, void* stream);
/* One Adam step over flat buffers p, g, m, v [n] (n % 4 == 0, 16-byte aligned): step += 1 (int64 device scalar), then
 * m = lerp(m, g, 1 - beta1); v = beta2 v + (1 - beta2) g^2; p -= lr / (1 - beta1^t) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps); lr is a
 * device scalar (schedulers update it in place; graph-replay safe), scratch is 2 device floats.
 * replaces: torch.optim.Adam(self.parameters(), lr=opt_lr) . step(), models/hyperbolic_encoder.py:199 (defaults: no weight decay) */
int coskad_adam_step(coskad_ctx* ctx, float* p, const float* g, float* m, float* v, int64_t n, const float* lr, float beta1,
                     float beta2, float eps, int64_t* step, float* scratch, void* stream);
/* 1 (default): coskad_train_mix_fwd runs on the tensor cores; 0: the FP32 CUDA-core kernels (A/B measurement).  Every
 * cross-CTA reduction of the training path (BatchNorm statistics, weight / bias / graph-operator gradients, linear heads)
 * is a fixed-order two-stage sum through a ctx-owned workspace: two runs on the same inputs are bit-identical.  The
 * workspace is sized by the first (eager) call; do not make the FIRST training call of a ctx inside a stream capture. */
int coskad_set_train_impl(coskad_ctx* ctx, int impl);
/* linear layers over the F = C*204 flattened features (btlnk / fc_mean / fc_var / rev_btlnk, models/sts/ae.py:155,206):
 * mode 0: out[B,D] = A_wide[B,F] W^T + bias; mode 1: out[B,F] = a_small[B,D] W + bias[F]; mode 2: out(=dW) += a_small^T A_wide.
 * w_is_fd = 0: W is [D,F] (btlnk); 1: W is [F,D] (rev_btlnk).  D <= 16, F % 4 == 0 (rows are read with 16-byte loads:
 * every COSKAD layout has F = C*T*V with T*V a multiple of 4); mode 2 accumulates into a zeroed / running dW. */
int coskad_train_linear(coskad_ctx* ctx, int mode, const float* a_small, const float* A_wide, const float* W, int w_is_fd,
                        const float* bias, int64_t B, int F, int D, float* out, void* stream);
/* out[N] += column sums of a[B,N] (bias gradients) */
int coskad_train_col_sum(coskad_ctx* ctx, const float* a, int64_t B, int N, float* out, void* stream);

/* ---- diagnostics ---------------------------------------------------------------------------- */
/* Sustained FP32-FMA rate of the device (TFLOP/s) from a register-resident FFMA loop; used by
 * bench.py as the measured denominator of the FP32 roofline. */
int coskad_measure_fp32_peak(coskad_ctx* ctx, double* tflops, void* stream);
/* Sustained tcgen05 kind::tf32 rate (TFLOP/s, one product per MAC -- a 3xTF32 product costs three) from back-to-back
 * M128 N256 K8 MMAs with A in TMEM and B in shared memory on every SM; the measured denominator of roofline_tensor. */
int coskad_measure_tf32_peak(coskad_ctx* ctx, double* tflops, void* stream);
/* test aid: run tile 0 of the fused kernel up to `stage` (S0..S23 of fused_eval.cuh) and dump the
 * CTA's shared-memory activations; coskad_debug_fused_floats() floats are written to dbg_out. */
int coskad_debug_fused_stage(coskad_ctx* ctx, int with_decoder, const float* x, int64_t B, int stage,
                             float* dbg_out, void* stream);
int coskad_debug_fused_floats(void);
/* test aid for the tcgen05 plumbing: out[128,N] = A[128,K] W[N,K]^T with 3xTF32; Bhi/Blo are the canonical K-major
 * shared-memory images of trunc_tf32(W) and W - trunc_tf32(W) */
int coskad_debug_tc_mix(coskad_ctx* ctx, const float* A, const float* Bhi, const float* Blo, int K, int N,
                        int swap_strides, float* out, void* stream);
int coskad_fused_tile_windows(void);
/* number of kernel launches issued through this ctx since creation */
int64_t coskad_launch_count(const coskad_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* COSKAD_B200_H_ */
