/*
 * eda_b200 — C ABI of the B200-native (sm_100a) hot path of yanmin-wu/EDA.
 *
 * This header is the drop-in boundary.  Every entry point takes plain device pointers,
 * sizes and a CUDA stream handle (cudaStream_t passed as void*; NULL = legacy default
 * stream) and returns 0 on success or a negative EDA_ERR_* code.  Nothing here exits the
 * process (the reference's CUDA_CHECK_ERRORS() calls exit(-1),
 * pointnet2/_ext_src/include/cuda_utils.h:35-44), nothing allocates: the CALLER owns every
 * buffer including scratch, and every call is asynchronous with respect to the host and
 * ordered on `stream` — the same stream contract as the reference ops, which launch on
 * at::cuda::getCurrentCUDAStream() of the current device.
 *
 * All tensors are dense row-major ("contiguous"); float = IEEE fp32, index = int32,
 * exactly as the reference checks (pointnet2/_ext_src/include/utils.h:10-30).
 *
 * Each function cites the reference interface it replaces (path relative to the
 * reference repository root).  The reference-side binding a maintainer adds is shown in
 * INTEGRATION.md.
 */
#ifndef EDA_B200_H_
#define EDA_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define EDA_API __attribute__((visibility("default")))
#else
#define EDA_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define EDA_OK 0
#define EDA_ERR_INVALID_ARGUMENT (-1) /* null pointer, negative size, unsupported shape */
#define EDA_ERR_CUDA_LAUNCH (-2)      /* cudaGetLastError() != cudaSuccess after launch */
#define EDA_ERR_NO_DEVICE (-3)        /* no sm_100 device / driver */
#define EDA_ERR_UNSUPPORTED (-4)      /* configuration outside the compiled template set */

/* Library/ABI version (major*10000 + minor*100 + patch). */
EDA_API int eda_version(void);
/* Static human-readable string for an EDA_ERR_* code (never NULL). */
EDA_API const char *eda_error_string(int code);
/* Text of the last CUDA error seen by this thread inside the library ("" if none). */
EDA_API const char *eda_last_cuda_error(void);
/* Number of CUDA kernels this library has launched in this process so far (bookkeeping only). */
EDA_API unsigned long long eda_launch_count(void);

/* ---------------------------------------------------------------------------------------
 * Furthest point sampling.
 * Replaces: at::Tensor furthest_point_sampling(at::Tensor points, const int nsamples)
 *           pointnet2/_ext_src/src/sampling.cpp:70-91 (kernel sampling_gpu.cu:74-178).
 * xyz (B,N,3) f32 -> idxs (B,m) i32, bit-exact with the reference incl. its tie-break.
 * `scratch`: device buffer of eda_fps_scratch_bytes(B,N,m) bytes (may be NULL when that
 * is 0).  The reference allocates its (B,N) `tmp` per call (sampling.cpp:78-80); here the
 * running minima live in registers and scratch is only needed by the large-N fallback.
 */
EDA_API size_t eda_fps_scratch_bytes(int B, int N, int m);
/* FPS of an FPS-ordered set (the backbone's stages 2-4 sample from the previous stage's output):
 * eda_fps_identity_check verifies, in parallel and with the reference's exact distance arithmetic, whether the answer for
 * scene b is 0, 1, ..., m-1 — true iff at every step i < m point i is the STRICT maximiser of the running minimum, so
 * no tie-break is involved — and writes not_identity[b] = 0 (verified) or 1 (ties, duplicates, skipped points, NaNs:
 * run the real thing).  dsel is scratch (B*m floats).  eda_furthest_point_sampling_ex is eda_furthest_point_sampling
 * (+ optional progress milestones) that consults such a flag array on the device: verified scenes get the identity
 * immediately, the others the full algorithm — bit-exact either way, no host synchronisation. */
EDA_API int eda_fps_identity_check(const float *xyz, int B, int n, int m, float *dsel, int *not_identity, void *stream);
EDA_API int eda_furthest_point_sampling_ex(const float *xyz, int B, int N, int m, void *scratch, int *idxs,
                                           int *progress, int every, const int *not_identity, void *stream);
EDA_API int eda_furthest_point_sampling(const float *xyz, int B, int N, int m, void *scratch, int *idxs,
                                void *stream);

/* Same sampling, publishing progress: after every `every` samples (and after the last one) of a scene, that
 * scene's cluster makes its indices so far visible device-wide and adds 1 to progress[j], j = 0, 1, ... the index
 * of the milestone — `progress` is an array of ceil(m / every) words, ONE PER MILESTONE, so that progress[j]
 * reaching (previous total + B) means every scene of the batch has passed milestone j even when the scenes run in
 * several waves (never reset by the library: the caller tracks each word's running total).  Together with
 * eda_stream_wait_value32 this lets a second stream run ball query + the fused MLP on the first centres while
 * the strictly serial sampling of the remaining ones continues on otherwise idle SMs.  every >= 2. */
EDA_API int eda_furthest_point_sampling_progress(const float *xyz, int B, int N, int m, void *scratch, int *idxs,
                                                 int *progress, int every, void *stream);
/* Measurement aid (no reference counterpart): the sampler's per-iteration reduction + cluster-exchange chain alone
 * (warp arg-max, shared-memory stage, bar.sync, st.async/mbarrier exchange between the `cluster` CTAs of a scene,
 * cluster arg-max) for `iters` dependent iterations on B clusters of `threads` threads — the latency floor bench.py
 * quotes next to the sampler's cycles per iteration.  sink: B words (keeps the chain alive).  Also reports through
 * eda_fps_plan which decomposition eda_furthest_point_sampling uses for (N). */
EDA_API int eda_selftest_fps_exchange(int B, int cluster, int threads, int iters, unsigned int *sink, void *stream);
EDA_API int eda_fps_plan(int B, int N, int m, int *cluster, int *threads, int *points_per_thread);
/* Stream-ordered wait on a device word: work queued on `stream` after this call starts once
 * (int)(*addr - value) >= 0 (cuStreamWaitValue32; no SM is occupied while waiting). */
EDA_API int eda_stream_wait_value32(void *stream, const int *addr, int value);

/* Ball query.
 * Replaces: at::Tensor ball_query(at::Tensor new_xyz, at::Tensor xyz, const float radius,
 *           const int nsample)  pointnet2/_ext_src/src/ball_query.cpp:13-37
 *           (kernel ball_query_gpu.cu:14-49).
 * new_xyz (B,M,3), xyz (B,N,3) -> idx (B,M,nsample) i32; every slot is written (zeros for
 * an empty ball), so idx need not be pre-zeroed.  Bit-exact. */
EDA_API int eda_ball_query(const float *new_xyz, const float *xyz, int B, int N, int M, float radius,
                   int nsample, int *idx, void *stream);

/* Ball query for centres m0 .. m0+mc-1 of the Mtot centres of every scene, centre coordinates read through
 * the FPS indices (xyz[centre_idx[b*Mtot + j]], bit-identical to new_xyz); writes rows m0.. of idx (B,Mtot,nsample). */
EDA_API int eda_ball_query_range(const float *xyz, const int *centre_idx, int B, int N, int Mtot, int m0, int mc,
                                 float radius, int nsample, int *idx, void *stream);

/* Grouping.  Replaces group_points / group_points_grad,
 * pointnet2/_ext_src/src/group_points.cpp:17-65 (kernels group_points_gpu.cu:13-80).
 * points (B,C,N), idx (B,M,S) -> out (B,C,M,S).
 * grad: grad_out (B,C,M,S) -> grad_points (B,C,N); the callee zero-fills grad_points. */
EDA_API int eda_group_points(const float *points, const int *idx, int B, int C, int N, int M, int S,
                     float *out, void *stream);
EDA_API int eda_group_points_grad(const float *grad_out, const int *idx, int B, int C, int N, int M, int S,
                          float *grad_points, void *stream);

/* Gather.  Replaces gather_points / gather_points_grad,
 * pointnet2/_ext_src/src/sampling.cpp:20-69 (kernels sampling_gpu.cu:13-62).
 * points (B,C,N), idx (B,M) -> out (B,C,M);  grad_out (B,C,M) -> grad_points (B,C,N) (zero-filled here). */
EDA_API int eda_gather_points(const float *points, const int *idx, int B, int C, int N, int M, float *out,
                      void *stream);
EDA_API int eda_gather_points_grad(const float *grad_out, const int *idx, int B, int C, int N, int M,
                           float *grad_points, void *stream);

/* 3-NN + interpolation.  Replaces three_nn / three_interpolate / three_interpolate_grad,
 * pointnet2/_ext_src/src/interpolate.cpp:19-104 (kernels interpolate_gpu.cu:14-159).
 * unknown (B,n,3), known (B,m,3) -> dist2 (B,n,3) f32 (SQUARED distances, as the reference
 * op returns; the sqrt is applied by the Python wrapper, pointnet2_utils.py:142), idx (B,n,3).
 * points (B,C,m), idx (B,n,3), weight (B,n,3) -> out (B,C,n).
 * grad_out (B,C,n) -> grad_points (B,C,m) (zero-filled here). */
EDA_API int eda_three_nn(const float *unknown, const float *known, int B, int n, int m, float *dist2,
                 int *idx, void *stream);
EDA_API int eda_three_interpolate(const float *points, const int *idx, const float *weight, int B, int C,
                          int m, int n, float *out, void *stream);
EDA_API int eda_three_interpolate_grad(const float *grad_out, const int *idx, const float *weight, int B,
                               int C, int n, int m, float *grad_points, void *stream);

/* Feature propagation in the row-major layout its MLP consumes.  Replaces, as ONE pass, the weight arithmetic +
 * three_interpolate + torch.cat of PointnetFPModule.forward (pointnet2/pointnet2_modules.py:393-410;
 * pointnet2_utils.py:142 for the sqrt): with dist2 / idx (B,n,3) from eda_three_nn,
 *   weight_k = (1 / (sqrt(dist2_k) + 1e-8)) / sum_k(1 / (sqrt(dist2_k) + 1e-8))        (written to `weight` (B,n,3), may be NULL)
 *   x0[(b,j)] = [ sum_k weight_k * known[b, idx_k, :]  (C2) | skip[b, j, :]  (C1) ]       x0 is (B*n, C2 + C1)
 * known (B,m,C2; row stride ldk) and skip (B,n,C1; row stride lds; NULL with C1 = 0) are POINT-major.
 * eda_fp_scatter_rows is the backward of the interpolated half (three_interpolate_grad, interpolate_gpu.cu:121-148):
 *   dknown (B,m,C2) point-major (zero-filled here) += weight_k * dx[(b,j), 0:C2] at idx_k. */
EDA_API int eda_fp_gather_rows(const float *known, int ldk, const float *skip, int lds, const int *idx,
                               const float *dist2, int B, int n, int m, int C2, int C1, float *x0, float *weight,
                               void *stream);
EDA_API int eda_fp_scatter_rows(const float *dx, int ldx, const int *idx, const float *weight, int B, int n, int m, int C2,
                                float *dknown, void *stream);

/* ---------------------------------------------------------------------------------------
 * Fused set-abstraction grouped MLP: neighbour gather + centre/normalise + concat + 3 x [1x1 conv ->
 * BatchNorm -> ReLU] + max over nsample, one kernel, tcgen05 kind::tf32 with fp32 accumulation.
 * Replaces the composition QueryAndGroup.forward (minus the ball query itself)
 *   pointnet2/pointnet2_utils.py:344-359  ->  SharedMLP  pointnet2/pytorch_utils.py:11-36,67-120
 *   ->  F.max_pool2d  pointnet2/pointnet2_modules.py:251-267.
 *
 * Shapes: xyz (B,N,3), new_xyz (B,M,3), idx (B,M,S) from eda_ball_query, features POINT-major: row
 * (b,i) = feat + (b*N+i)*feat_stride, C floats (C may be 0, feat NULL); conv weights W1 (C1,3+C) with
 * the reference's column order [xyz | features], W2 (C2,C1), W3 (C3,C2), no bias (pytorch_utils.py:87).
 * Supported: C1,C2 <= 128, C3 <= 256, all multiples of 16; S a power of two >= 16.  Anything else
 * returns EDA_ERR_UNSUPPORTED and the host composes the unfused ops.
 *
 * eda_sa_mlp_pack: folds per-output-channel `scale_l` (NULL = 1) into W_l, rounds to tf32 and lays the
 *   first `nlayers` layers out in the streaming order of the kernel; `packed` holds
 *   eda_sa_mlp_packed_floats(C,C1,C2,C3) floats (0 = unsupported dims).
 * eda_sa_mlp_forward, stats_layer == 0: out (B,M,C3) POINT-major = max_s relu(conv3'(relu(conv2'(relu(
 *   conv1'(x) + shift1)) + shift2)) + shift3)   (zero-filled by the callee).
 *   stats_layer == l in 1..3: layers < l as above, layer l's conv output is only reduced to
 *   stats[0:C_l] = sum, stats[C_l:2C_l] = sum of squares over all B*M*S rows (zero-filled by the
 *   callee) — the batch statistics train-mode BatchNorm2d needs.  The sums are fp64 (atomicAdd on double): over
 *   ~1e6 rows, var = E[z^2] - E[z]^2 from fp32 sums loses up to 1e-2 relative when |mean| >> std.
 * eda_bn_finalize: (sum, sumsq, count) + gamma/beta/eps -> scale = gamma/sqrt(var+eps),
 *   shift = beta - mean*scale; optionally updates running_mean/var with `momentum` (unbiased variance),
 *   as nn.BatchNorm2d does in training.  count <= 0: use running_mean/var instead (eval mode).
 *   save_mean / save_invstd may be NULL.
 * eda_transpose_last2: (B,R,C) -> (B,C,R). */
EDA_API size_t eda_sa_mlp_packed_floats(int C, int C1, int C2, int C3);
EDA_API int eda_sa_mlp_pack(const float *W1, const float *W2, const float *W3, const float *scale1,
                            const float *scale2, const float *scale3, int C, int C1, int C2, int C3, int nlayers,
                            float *packed, void *stream);
EDA_API int eda_sa_mlp_forward(const float *xyz, const float *new_xyz, const float *feat, int feat_stride,
                               const int *idx, const float *packed, const float *shift1, const float *shift2,
                               const float *shift3, int B, int N, int M, int S, int C, int C1, int C2, int C3,
                               float radius, int normalize_xyz, int stats_layer, float *out, double *stats,
                               void *stream);
/* eda_sa_mlp_forward (stats_layer 0) for centres m0 .. m0+mc-1 only, centres given by FPS index; `out`
 * (B,Mtot,C3) must have been zero-filled by the caller (once, before the first range). */
EDA_API int eda_sa_mlp_forward_range(const float *xyz, const int *centre_idx, const float *feat, int feat_stride,
                                     const int *idx, const float *packed, const float *shift1, const float *shift2,
                                     const float *shift3, int B, int N, int Mtot, int m0, int mc, int S, int C, int C1,
                                     int C2, int C3, float radius, int normalize_xyz, float *out, void *stream);
EDA_API int eda_bn_finalize(const double *stats, double count, const float *gamma, const float *beta, float eps,
                            float momentum, float *running_mean, float *running_var, int update_running, int C,
                            float *scale, float *shift, float *save_mean, float *save_invstd, void *stream);
EDA_API int eda_transpose_last2(const float *in, int B, int R, int C, float *out, void *stream);
/* Same for a column slice of wider rows: in is (B, R, ld) and columns col0 .. col0+C-1 become out (B, C, R). */
EDA_API int eda_transpose_strided(const float *in, int col0, int ld, int B, int R, int C, float *out, void *stream);

/* ---------------------------------------------------------------------------------------
 * Cross-modal attention layers (models/encoder_decoder_layers.py).  All activations are row-major
 * "batch-first" matrices: row (b, s) of a (B, S, E) tensor is one contiguous E-float row.
 *
 * eda_linear_forward: Y[R x N] = act((X [+ P])[R x K] W[N x K]^T + bias), optionally followed by
 *   LayerNorm(residual + Y) * gamma + beta over the N columns (layer_norm != 0; residual may be NULL).
 *   Replaces F.linear inside nn.MultiheadAttention (in/out projections, torch/nn/functional.py
 *   math path as called from encoder_decoder_layers.py:87-117,149-183,366-401), the residual +
 *   nn.LayerNorm after each attention / FFN block (:94-96,106-107,118-122,371-405), the FFN linears
 *   (:53-59,322-328) and the 1x1 Conv1d's of PositionEmbeddingLearned (:24-28).
 *   Up to 3 problems with the same K, N and epilogue share one launch (q / k / v in-projections).
 *   `w_packed` comes from eda_linear_pack (W (N,K) row-major, optional per-row scale folded in, tf32
 *   rounding, kernel streaming order); eda_linear_packed_floats(N,K) floats, 0 = unsupported
 *   (N % 16 == 0, 16 <= N <= 320).  x, pos: R x K contiguous; y, residual: R x N contiguous,
 *   16-byte aligned.  tcgen05 kind::tf32, fp32 accumulate.
 *
 * eda_attention_forward: ctx (B,Nq,H*D) = softmax(q k^T * scale + mask) v per head, q (B,Nq,H*D),
 *   k (B,Nk,H*D) already projected, v CHANNEL-major vt (B,H*D,ldv) with ldv >= Nk, ldv % 4 == 0 (written
 *   that way by eda_linear_forward with y_batch_rows = Nk: four consecutive keys of one channel are one
 *   16-byte unit of the PV operand); k and vt are expected to hold tf32-representable values (eda_linear_forward
 *   with round_tf32; other values are truncated by the tensor core); key_padding_mask (B,Nk) bytes, nonzero = ignore (may be NULL).
 *   Replaces the bmm/softmax/bmm of the same math path.  Head dims compiled: D in {32, 36, 64}.  A fully masked row
 *   yields NaN as in the reference. */
typedef struct eda_linear_problem {
  const float *x;        /* (rows, K) */
  const float *pos;      /* (rows, K) added to x before the product, or NULL */
  const float *w_packed; /* eda_linear_pack output for W (N, K) */
  const float *bias;     /* (N) or NULL */
  const float *residual; /* (rows, N) or NULL; only read when layer_norm != 0 */
  float *y;              /* (rows, N); or channel-major when y_batch_rows > 0 (below) */
  int rows;
  int y_batch_rows;      /* 0: y row-major.  T > 0: rows are (batch, t < T) and y is (batch, N, y_ld) channel-major,
                            y[(row / T * N + col) * y_ld + row % T] — how the attention kernel wants V */
  int y_ld;              /* >= T; padding columns are not written */
  int round_tf32;        /* != 0: round outputs to tf32 (round-to-nearest): K and V projections, whose outputs
                            eda_attention_forward feeds to the tensor cores as they are */
  float *pre_ln;         /* optional (rows, N): receives the LayerNorm input (residual + product), which
                            eda_layernorm_backward needs; NULL = not stored */
  int y_row_stride;      /* 0: rows of y are N floats apart.  > N (multiple of 4): y is a column block of a wider
                            row-major matrix (row-major output only) */
  int reserved;
} eda_linear_problem;
EDA_API size_t eda_linear_packed_floats(int N, int K);
EDA_API int eda_linear_pack(const float *W, const float *scale, int N, int K, float *packed, void *stream);
/* Same packing for a strided view: element (n, k) of the (N, K) weight is W[n * stride_n + k * stride_k].  With
 * stride_n = 1, stride_k = ld it packs the TRANSPOSE of a row-major matrix, which turns eda_linear_forward into the
 * activation-gradient GEMM dX = dY W of a layer y = x W^T (autograd's mm backward in the reference). */
EDA_API int eda_linear_pack_strided(const float *W, long long stride_n, long long stride_k, int N, int K, float *packed,
                                    void *stream);
/* Batched packing: `descs_device` is an array of `count` descriptors IN DEVICE MEMORY; weight i (N_i, K_i; element (n, k) at
 * w[n * stride_n + k * stride_k]) is packed into dst_i (eda_linear_packed_floats(N_i, K_i) floats).  max_elements =
 * max_i N_i * roundup8(K_i).  One launch for all weights of a model. */
typedef struct eda_linear_pack_desc {
  const float *w;
  float *dst;
  long long stride_n, stride_k;
  int N, K, Kpad, reserved; /* Kpad = K rounded up to a multiple of 8 */
} eda_linear_pack_desc;
EDA_API int eda_linear_pack_batch(const void *descs_device, int count, int max_elements, void *stream);
EDA_API int eda_linear_forward(const eda_linear_problem *probs, int nprobs, int K, int N, int relu,
                               const float *ln_gamma, const float *ln_beta, float ln_eps, int layer_norm,
                               float dropout_p, unsigned int dropout_seed, const unsigned int *dropout_epoch, void *stream);
/* Development aid: clock64() phase stamps of CTA 0 of the most recent eda_linear_forward launch (synchronises).  A
 * negative n returns the first -n (<= 16) stamps of the most recent tcgen05 eda_wgrad launch instead. */
EDA_API int eda_debug_timestamps(long long *host_out, int n);
EDA_API int eda_debug_timestamps_attn(long long *host_out, int n); /* same, attention kernel, key block 1 */
EDA_API int eda_attention_forward(const float *q, const float *k, const float *vt, int ldv,
                                  const unsigned char *key_padding_mask, int B, int Nq, int Nk, int H, int D,
                                  float scale, float dropout_p, unsigned int dropout_seed, const unsigned int *dropout_epoch, float *ctx, void *stream);
/* Training variant: also writes lse (B, H, Nq), the log-sum-exp of every query's masked, scaled scores (NULL = skip). */
EDA_API int eda_attention_forward_lse(const float *q, const float *k, const float *vt, int ldv,
                                      const unsigned char *key_padding_mask, int B, int Nq, int Nk, int H, int D,
                                      float scale, float dropout_p, unsigned int dropout_seed, const unsigned int *dropout_epoch, float *ctx, float *lse,
                                      void *stream);

/* ---------------------------------------------------------------------------------------
 * Backward pass of the attention layers.  In the reference this is autograd through
 * nn.MultiheadAttention's math path, nn.Linear and nn.LayerNorm (torch/nn/functional.py:6607-6665 as used by
 * models/encoder_decoder_layers.py:37-124,127-186,288-407): per module ~45 library kernels and two (B*H, Nq, Nk)
 * tensors in HBM.  Here:
 *
 * eda_attention_backward: q (B,Nq,H*D), k (B,Nk,H*D), v (B,Nk,H*D; scenes v_batch_stride floats apart) are the
 *   PROJECTED inputs of eda_attention_forward_lse (v row-major here), ctx / lse its outputs, dctx the gradient of ctx.
 *   Recomputes the probabilities from lse tile by tile and writes dq, dk, dv (same layouts; fully overwritten, no
 *   atomics, deterministic).  Instead of v, the values may be given channel-major as the forward call took them:
 *   vt (B, H*D, ldv) with v = NULL — no transposed copy is needed then.  delta (B,H,Nq) is scratch (rowsum(dctx * ctx)).  The dropout mask of the forward call
 *   (dropout_p, dropout_seed) is regenerated from the same hash.  Head dims compiled: D in {32, 36}.
 * eda_wgrad: for every problem i, dw_i (N, K; row stride ldw) += dy_i (rows, N; ldy)^T x_i (rows, K; ldx) and, when db_i
 *   is not NULL, db_i (N) += column sums of dy_i.  ACCUMULATES (atomics / TMA bulk reductions): zero the outputs first.
 *   N, K, ldy, ldx multiples of 4, dy / x 16-byte aligned.  Up to 6 problems of one (N, K) per launch.  From 3000 rows in
 *   total (N >= 64, K >= 32) the tcgen05 kernel runs (operands by TMA as they lie in memory), else warp-level mma.sync.
 * eda_layernorm_backward: y = LayerNorm(u) * gamma + beta over the last dim N (<= 384, multiple of 4).  Writes
 *   du (rows, N); dgamma / dbeta (N) are ACCUMULATED (may be NULL).  With dropout_p > 0 also writes
 *   dproj = du * keep / (1 - p), the gradient of the GEMM output that eda_linear_forward(dropout_p, dropout_seed)
 *   dropped before the residual add (single-problem launch); with dropout_p == 0 dproj is not written (it equals du).
 * eda_relu_backward: out = dy * [y > 0] * scale, n elements (multiple of 4). */
typedef struct eda_wgrad_problem {
  const float *dy; /* (rows, N), row stride ldy */
  const float *x;  /* (rows, K), row stride ldx */
  float *dw;       /* (N, K), row stride ldw; accumulated */
  float *db;       /* (N) accumulated, or NULL */
  long long rows;
  int ldy, ldx, ldw;
  const float *x_scale; /* optional (K): x is consumed as relu(x * x_scale[k] + x_shift[k]) — the folded   */
  const float *x_shift; /* BatchNorm + ReLU of the layer that produced it (both NULL: x as it is)          */
} eda_wgrad_problem;
EDA_API int eda_attention_backward(const float *q, const float *k, const float *v, long long v_batch_stride,
                                   const float *vt, int ldv, const float *dctx, const float *ctx, const float *lse,
                                   const unsigned char *key_padding_mask, int B, int Nq, int Nk, int H, int D,
                                   float scale, float dropout_p, unsigned int dropout_seed, const unsigned int *dropout_epoch, float *delta, float *dq,
                                   float *dk, float *dv, void *stream);
/* Same result on the tcgen05 tensor cores (TMEM-resident score tiles; csrc/attn_bwd_tc.cu).  Additionally takes
 * channel-major copies kt (B, H*D, ldk) of k and qt, dctx_t (B, H*D, ldq) of q and dctx (ld >= N, multiple of 4; padding
 * columns are never read): the second products dq += dS k, dk += dS^T q, dv += P^T dctx want four consecutive column
 * indices of one channel per 16-byte operand unit. */
EDA_API int eda_attention_backward_tc(const float *q, const float *k, const float *v, long long v_batch_stride,
                                      const float *kt, int ldk, const float *qt, const float *dctx_t, int ldq,
                                      const float *dctx, const float *ctx, const float *lse,
                                      const unsigned char *key_padding_mask, int B, int Nq, int Nk, int H, int D,
                                      float scale, float dropout_p, unsigned int dropout_seed, const unsigned int *dropout_epoch, float *delta, float *dq,
                                      float *dk, float *dv, void *stream);
EDA_API int eda_wgrad(const eda_wgrad_problem *probs, int nprobs, int N, int K, void *stream);
/* Same accumulation for a tiny contraction width the tensor-core kernel does not take (K <= 8, any alignment): the 3- /
 * 6-channel first conv of PositionEmbeddingLearned (models/encoder_decoder_layers.py:24-28).  fp32 FMAs. */
EDA_API int eda_wgrad_small(const float *dy, int ldy, const float *x, int ldx, long long rows, int N, int K, float *dw,
                            int ldw, float *db, void *stream);
EDA_API int eda_layernorm_backward(const float *dy, const float *u, const float *gamma, float eps, long long rows, int N,
                                   float *du, float *dproj, float *dgamma, float *dbeta, float dropout_p,
                                   unsigned int dropout_seed, const unsigned int *dropout_epoch, void *stream);
EDA_API int eda_relu_backward(const float *dy, const float *y, float scale, long long n, float *out, void *stream);
/* Train-mode dropout (nn.Dropout after attention / FFN blocks, attention-probability dropout of
 * nn.MultiheadAttention(dropout=p), encoder_decoder_layers.py:47-59,94,106,118 ...): dropout_p > 0 makes
 * eda_linear_forward zero each output element (after bias / ReLU, before the residual) and eda_attention_forward each
 * softmax probability (after normalisation) with probability p and scale the rest by 1/(1-p).  The decision is a
 * counter-based hash of (dropout_seed, element position) — not torch's Philox stream, so masks differ from the
 * reference's for the same torch seed (train-mode parity is defined at p = 0).  eda_dropout_mask regenerates the
 * keep-mask (1.0 / 0.0) a forward call applied, out[a * cols + b] for a < rows, b < cols, hashing
 * (a * a_mul + a_add, b):  linear problem i of a launch: a_mul = 3, a_add = i, rows = R, cols = N;
 * attention: a_mul = 1, a_add = 0, rows = B*H*Nq (row (b*H + h)*Nq + q), cols = Nk. */
EDA_API int eda_dropout_mask(unsigned int seed, const unsigned int *dropout_epoch, float p, long long rows, int cols,
                             unsigned int a_mul, unsigned int a_add, float *out, void *stream);
/* nn.Dropout as a stand-alone pass with the same decision function (used after a BatchNorm + ReLU apply, where no GEMM
 * epilogue carries it: ThreeLayerMLP of the prediction heads, models/modules.py:88-106): out = keep ? x / (1 - p) : 0,
 * element (a, b) hashed as (a * a_mul + a_add, b) like eda_dropout_mask. */
EDA_API int eda_dropout_apply(const float *x, unsigned int seed, const unsigned int *dropout_epoch, float p, long long rows,
                              int cols, unsigned int a_mul, unsigned int a_add, float *out, void *stream);
/* Dropout epoch: every dropout-applying entry point (forward and backward) takes `dropout_epoch`, an optional device
 * word whose value the kernel adds to dropout_seed when it RUNS (NULL = off).  A CUDA graph of a training step freezes
 * the host-drawn seeds; with an epoch word that a captured device op increments once per step, every replay still draws
 * fresh masks, identical in that step's forward and backward.  It is a plain per-call argument: the library keeps no
 * dropout state (its only state is the per-thread last-error string, a launch counter and thread-safe per-device
 * caches of kernel attributes).  The word must stay allocated while kernels launched with it run. */

/* ---------------------------------------------------------------------------------------
 * Backward pass of the fused set-abstraction stage (eda_sa_mlp_forward).  In the reference: autograd through
 * QueryAndGroup / SharedMLP / max_pool2d (pointnet2/pointnet2_utils.py:209-257,317-376, pytorch_utils.py:11-36,
 * pointnet2_modules.py:251-267) = cuDNN convolution / BatchNorm backward on (B,C,npoint,nsample) tensors + the atomic
 * scatter of group_points_grad (group_points_gpu.cu:48-80).  Here the grouped rows are row-major (row = (b, centre,
 * sample), column = channel): every layer's GEMM runs on eda_linear_forward / eda_wgrad and these entry points are the
 * memory-bound stages in between (one pass over HBM each).  scale / shift are the folded BatchNorm terms
 * (y = z * scale + shift), mean / invstd the statistics they were built from, stats = [sum(dy) (C), sum(dy * zhat) (C)]
 * (ACCUMULATED: zero first; they are also the gradients of the BatchNorm bias / weight), count = rows the batch
 * statistics were taken over, batch_stats = 0 for eval-mode (running-statistics) BatchNorm.
 *   eda_sa_gather_rows          x0 (B*M*S, K0pad) = [features[idx] (C) | (xyz[idx] - new_xyz) [/ radius] (3) | 0]
 *   eda_bn_relu_apply           out = relu(z * scale[c] + shift[c]),  z (rows, C)
 *   eda_sa_pool_backward        per (centre, channel): first arg-max row of y3 over the S samples -> amax (centres, C)
 *                               (-1 where the max is <= 0: the ReLU kills it), stats of layer 3
 *   eda_sa_pool_backward_apply  z3 <- dz3 = scale (dy3 - mean(dy3) - zhat3 mean(dy3 zhat3)),  dy3 = grad_out at amax
 *   eda_bn_relu_backward_stats  stats of dy = da * [z * scale + shift > 0]
 *   eda_bn_relu_backward_apply  da <- dz = scale (dy - mean(dy) - zhat mean(dy zhat))
 *   eda_sa_scatter_rows         dfeat (B, N, C) point-major += dx0[:, :C] scattered by idx (red.global.add) */
/* eda_rows_gemm: y (rows, N; row stride ldy) = f(x) (rows, K; ldx) W'^T with W'[n][k] = w[n * w_stride_n + k * w_stride_k]
 * (so a row-major (N, K) weight is (K, 1), and its use as dX = dY W is (1, ld)) and f(x) = x, or
 * relu(x * in_scale[k] + in_shift[k]) when in_scale / in_shift are given.  K, N multiples of 8, K <= 288; tf32 operands,
 * fp32 accumulation.  Persistent row-streaming kernels for the 10^5 - 10^6-row GEMMs of the SA stage: tcgen05 with
 * TMA-fed operands and TMA tile stores from 16 384 rows (x, y 16-byte aligned, ldx, ldy multiples of 4), warp-level
 * mma.sync below that. */
EDA_API int eda_rows_gemm(const float *x, int ldx, const float *in_scale, const float *in_shift, const float *w,
                          long long w_stride_n, long long w_stride_k, long long rows, int K, int N, float *y, int ldy,
                          void *stream);
/* Same GEMM, additionally accumulating the column statistics of its OUTPUT into stats (2N doubles: [sum_r y[r][n],
 * sum_r y[r][n]^2], zero first): the BatchNorm batch statistics of the layer, taken in the epilogue while the tile is
 * still in registers instead of by a second pass over y (eda_col_stats).  stats NULL = plain eda_rows_gemm. */
EDA_API int eda_rows_gemm_stats(const float *x, int ldx, const float *in_scale, const float *in_shift, const float *w,
                                long long w_stride_n, long long w_stride_k, long long rows, int K, int N, float *y,
                                int ldy, double *stats, void *stream);
EDA_API int eda_sa_gather_rows(const float *xyz, const float *new_xyz, const float *feat, int feat_stride,
                               const int *idx, int B, int N, int M, int S, int C, int K0pad, float radius,
                               int normalize_xyz, float *x0, void *stream);
EDA_API int eda_bn_relu_apply(const float *z, const float *scale, const float *shift, long long rows, int C, float *out,
                              void *stream);
EDA_API int eda_sa_pool_backward(const float *z3, const float *scale, const float *shift, const float *mean,
                                 const float *invstd, const float *grad_out, long long centres, int S, int C, int *amax,
                                 float *stats, void *stream);
EDA_API int eda_sa_pool_backward_apply(float *z3, const int *amax, const float *grad_out, const float *scale,
                                       const float *mean, const float *invstd, const float *stats, double count,
                                       int batch_stats, long long centres, int S, int C, void *stream);
EDA_API int eda_bn_relu_backward_stats(const float *da, const float *z, const float *scale, const float *shift,
                                       const float *mean, const float *invstd, long long rows, int C, float *stats,
                                       void *stream);
EDA_API int eda_bn_relu_backward_apply(float *da, const float *z, const float *scale, const float *shift,
                                       const float *mean, const float *invstd, const float *stats, double count,
                                       int batch_stats, long long rows, int C, void *stream);
/* Row-major TRAINING forward of the same stage (activations are kept for the backward pass instead of being
 * recomputed): z_l from eda_rows_gemm, then
 *   eda_col_stats        stats[0:C] += sum_r z[r][c], stats[C:2C] += sum_r z[r][c]^2 (fp64; zero first) -> eda_bn_finalize
 *   eda_sa_pool_forward  out (centres, C) = max(0, max_s (z3 * scale + shift)) over the S rows of each centre; optional
 *                        amax (centres, C): the first row attaining it, -1 where nothing is positive
 *   eda_sa_pool_backward_stats  the reductions of eda_sa_pool_backward from such a saved amax (no scan over S) */
EDA_API int eda_col_stats(const float *z, long long rows, int C, double *stats, void *stream);
EDA_API int eda_sa_pool_forward(const float *z3, const float *scale, const float *shift, long long centres, int S, int C,
                                float *out, int *amax, void *stream);
EDA_API int eda_sa_pool_backward_stats(const float *z3, const int *amax, const float *mean, const float *invstd,
                                       const float *grad_out, long long centres, int S, int C, float *stats,
                                       void *stream);
EDA_API int eda_sa_scatter_rows(const float *dx0, const int *idx, int B, int N, int M, int S, int C, int K0pad,
                                float *dfeat, void *stream);

/* ---------------------------------------------------------------------------------------
 * Synchronised BatchNorm statistics over the GPUs of one node through NVLink peer memory (the reference converts every
 * BatchNorm to nn.SyncBatchNorm when more than one GPU is used, main_utils.py:335-338; torch's implementation issues an
 * NCCL all-gather per layer in the forward and an all-reduce per layer in the backward pass).
 *
 * peer_buffers_dev: DEVICE array of `world` pointers, entry r = rank r's exchange buffer as mapped into THIS process
 * (symmetric memory / CUDA IPC; the host side obtains them from torch.distributed._symmetric_memory); every buffer holds
 * eda_peer_buffer_bytes(world, max_elems) bytes and must be zero-filled once before first use.  All ranks must issue the
 * same sequence of peer calls (like any collective).  world <= 16.
 *   eda_peer_allreduce      data[0:n] <- sum over ranks, in place (fp32 when is_f64 == 0, else fp64; accumulation in
 *                           fp64, ranks added in rank order: bit-identical results on every rank).  One single-CTA launch:
 *                           remote stores of the local vector into every peer, release/acquire flags, ordered local sum.
 *   eda_bn_finalize_peer    eda_bn_finalize with that exchange fused in front: stats = THIS rank's [sum z, sum z^2],
 *                           count = rows of ALL ranks.
 * A peer that never arrives turns into a non-zero error word in the buffer (bytes 4..7) after ~2 s instead of a hang. */
EDA_API size_t eda_peer_buffer_bytes(int world, int max_elems);
EDA_API int eda_peer_allreduce(void *const *peer_buffers_dev, int world, int rank, int max_elems, void *data, int n,
                               int is_f64, void *stream);
EDA_API int eda_bn_finalize_peer(void *const *peer_buffers_dev, int world, int rank, int max_elems, const double *stats,
                                 double count, const float *gamma, const float *beta, float eps, float momentum,
                                 float *running_mean, float *running_var, int update_running, int C, float *scale,
                                 float *shift, float *save_mean, float *save_invstd, void *stream);

/* ---------------------------------------------------------------------------------------
 * Hardware self-test of the tcgen05/TMEM building blocks the fused kernels rely on (no reference
 * counterpart).  D[128,N] = A[128,K] * W[N,K]^T with kind::tf32, fp32 accumulate, one CTA.
 * mode 0: A from shared memory; mode 1: A from tensor memory; mode 2: A from shared memory, B staged
 * MN-major (the layout the attention kernel uses for V).  N % 16 == 0, 16 <= N <= 256;
 * K % 16 == 0, 16 <= K <= 128. */
EDA_API int eda_selftest_umma(const float *A, const float *W, int N, int K, int mode, float *D, void *stream);
/* Shared-memory operand layout probe (development aid): one kind::tf32 MMA (K = 8) whose B descriptor has
 * the given LBO / SBO / majorness / swizzle layout_type / start offset over a region holding float(i) at float index i; D (8, N) receives, for
 * (k, n), the float index the hardware read as B(n, k). */
EDA_API int eda_selftest_umma_probe(int N, int lbo_bytes, int sbo_bytes, int b_mn_major, int layout_type,
                                    int start_offset_bytes, float *D, void *stream);

/* Tensor-pipe rate probe (measurement aid, bench / DESIGN numbers): every CTA issues `iters` kind::tf32 MMAs of shape
 * 128 x N x 8 from one thread and writes [issue-loop cycles, cycles until completion] to cycles_device[2 * cta].
 * a_mode 0 / 1 / 2: A from shared memory (no swizzle) / tensor memory / shared memory SWIZZLE_128B; b_swizzled: B in
 * SWIZZLE_128B instead of the packed chunk-major layout; precomputed: descriptors hoisted out of the issue loop;
 * waiting_warps: extra warps of the CTA that sit in an mbarrier wait for the duration (as producer / staging warps do). */
EDA_API int eda_selftest_umma_rate(int N, int a_mode, int b_swizzled, int precomputed, int iters, int ctas,
                                   int waiting_warps, long long *cycles_device, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* EDA_B200_H_ */
