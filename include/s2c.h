/*
 * s2c.h -- C ABI of libs2c.so, the sm_100a (B200) implementation of the Scan2Cap
 * point-cloud -> caption hot path.
 *
 * This is the drop-in boundary: every entry point below replaces one function of the
 * reference's pybind module `pointnet2._ext` (lib/pointnet2/_ext_src/src/bindings.cpp:6-19)
 * or one Python-level step of the reference models, and takes only plain device pointers,
 * sizes and a CUDA stream -- no torch types.  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions (all entry points)
 *   - every pointer is a DEVICE pointer to a dense, contiguous array in the layout quoted;
 *     float = IEEE fp32, int = int32, unless stated otherwise;
 *   - outputs are CALLER-allocated.  Where the reference relies on zero-filled outputs
 *     (torch::zeros in its C++ wrappers) the kernel here writes every element itself, so the
 *     caller does not need to clear them, except for the *_grad functions whose `accumulate`
 *     semantics are stated per function;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); work is
 *     enqueued asynchronously on it, there is no hidden synchronisation and no global state;
 *   - return value: 0 = S2C_OK, otherwise an S2C_ERR_* code; s2c_last_error() returns a
 *     thread-local human-readable message for the last failing call of this thread.
 *     (The reference prints and calls exit(-1): include/cuda_utils.h:30-39.)
 */
#ifndef S2C_H_
#define S2C_H_

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define S2C_API __attribute__((visibility("default")))
#else
#define S2C_API
#endif

#define S2C_OK 0
#define S2C_ERR_INVALID_ARGUMENT 1
#define S2C_ERR_CUDA 2
#define S2C_ERR_UNSUPPORTED 3

/* Library / ABI version (major*10000 + minor*100 + patch). */
S2C_API int s2c_version(void);
/* Message of the last error raised on the calling thread ("" if none). */
S2C_API const char *s2c_last_error(void);

/* ------------------------------------------------------------------------------------------
 * furthest_point_sampling -- replaces _ext.furthest_point_sampling
 *   reference: src/sampling.cpp:66-87, src/sampling_gpu.cu:59-229
 *   xyz (B,N,3) f32  ->  idx (B,m) int32.   idx[:,0] = 0; bit-exact with the reference,
 *   including its tie-breaking (shared-memory tree order) and the |p|^2 <= 1e-3 skip rule.
 *   No scratch buffer is needed (the reference's (B,N) `temp` lives in registers here).
 *   new_xyz (B,m,3) f32 may be NULL; if given, the sampled coordinates xyz[b, idx[b,j]] are
 *   written too (fuses the gather_points call of pointnet2_modules.py:238-240).
 * ---------------------------------------------------------------------------------------- */
S2C_API int s2c_furthest_point_sampling(const float *xyz, int B, int N, int m, int *idx, float *new_xyz,
                                void *stream);

/* gather_points -- replaces _ext.gather_points (src/sampling.cpp:15-40, sampling_gpu.cu:8-30)
 *   points (B,C,N), idx (B,m) -> out (B,C,m) */
S2C_API int s2c_gather_points(const float *points, const int *idx, int B, int C, int N, int m, float *out,
                      void *stream);

/* gather_points_grad -- replaces _ext.gather_points_grad (sampling.cpp:41-65, sampling_gpu.cu:34-57)
 *   grad_out (B,C,m), idx (B,m) -> grad_points (B,C,N).  The output is zero-filled by this call
 *   (as torch::zeros does in the reference) and then accumulated with float atomics. */
S2C_API int s2c_gather_points_grad(const float *grad_out, const int *idx, int B, int C, int N, int m,
                           float *grad_points, void *stream);

/* ball_query -- replaces _ext.ball_query (src/ball_query.cpp:8-32, ball_query_gpu.cu:9-54)
 *   new_xyz (B,M,3), xyz (B,n,3) -> idx (B,M,nsample) int32: the first `nsample` point indices
 *   (in index order) with d2 < radius*radius (fp32); unfilled slots repeat the first hit; an
 *   empty ball gives zeros.  Bit-exact with the reference.  1 <= nsample <= 1024.
 *   cnt (B,M) int32 may be NULL; if given it receives min(#hits, nsample) per centre. */
S2C_API int s2c_ball_query(const float *new_xyz, const float *xyz, int B, int n, int M, float radius,
                   int nsample, int *idx, int *cnt, void *stream);

/* group_points -- replaces _ext.group_points (src/group_points.cpp:12-36, group_points_gpu.cu:8-28)
 *   points (B,C,N), idx (B,np,ns) -> out (B,C,np,ns) */
S2C_API int s2c_group_points(const float *points, const int *idx, int B, int C, int N, int npoints,
                     int nsample, float *out, void *stream);

/* group_points_grad -- replaces _ext.group_points_grad (group_points.cpp:38-62, group_points_gpu.cu:43-64)
 *   grad_out (B,C,np,ns), idx (B,np,ns) -> grad_points (B,C,N); zero-filled by this call, then
 *   accumulated with float atomics (order unspecified, as in the reference). */
S2C_API int s2c_group_points_grad(const float *grad_out, const int *idx, int B, int C, int N, int npoints,
                          int nsample, float *grad_points, void *stream);

/* three_nn -- replaces _ext.three_nn (src/interpolate.cpp:14-40, interpolate_gpu.cu:9-68)
 *   unknown (B,n,3), known (B,m,3) -> dist2 (B,n,3) f32 (SQUARED distances, as the extension
 *   returns them), idx (B,n,3) int32.  Bit-exact (strict <, first index wins, double bests). */
S2C_API int s2c_three_nn(const float *unknown, const float *known, int B, int n, int m, float *dist2,
                 int *idx, void *stream);

/* three_interpolate -- replaces _ext.three_interpolate (interpolate.cpp:42-68, interpolate_gpu.cu:72-112)
 *   points (B,C,m), idx (B,n,3), weight (B,n,3) -> out (B,C,n) */
S2C_API int s2c_three_interpolate(const float *points, const int *idx, const float *weight, int B, int C,
                          int m, int n, float *out, void *stream);

/* three_interpolate_grad -- replaces _ext.three_interpolate_grad (interpolate.cpp:70-99, interpolate_gpu.cu:116-154)
 *   grad_out (B,C,n), idx, weight (B,n,3) -> grad_points (B,C,m); zero-filled by this call. */
S2C_API int s2c_three_interpolate_grad(const float *grad_out, const int *idx, const float *weight, int B,
                               int C, int n, int m, float *grad_points, void *stream);

/* ------------------------------------------------------------------------------------------
 * query_and_group -- one fused call for QueryAndGroup.forward (lib/pointnet2/pointnet2_utils.py:317-376
 *   with use_xyz=True, ret_grouped_xyz=True, normalize_xyz as given), i.e. ball_query ->
 *   group_points(xyz^T) - new_xyz -> (/ radius) -> group_points(features) -> cat.
 *     xyz       (B,n,3)
 *     new_xyz   (B,M,3)
 *     features  per-point features, C channels (C may be 0, then features may be NULL):
 *                 feat_layout 0: (B,C,n)  channel-major, the reference's layout
 *                 feat_layout 1: (B,n,C)  point-major with row stride feat_stride floats
 *                                (lets the caller pass point_clouds[...,3:] without a transpose)
 *     idx       (B,M,nsample) int32 out (same values as s2c_ball_query)
 *     grouped   out, 3+C channels, xyz channels first:
 *                 out_layout 0: (B,3+C,M,nsample)   the reference's layout
 *                 out_layout 1: (B,M,nsample,3+C)   channels-last (what the grouped-MLP kernels eat)
 *                 out_layout 2: (B,M,nsample,Cp)    channels-last, Cp = 3+C rounded up to a multiple of 4, pad = 0
 *                                                   (16-byte aligned rows for the TMA-fed MLP kernel)
 *   normalize_xyz != 0 multiplies the relative coordinates by the fp32 reciprocal of radius
 *   (what torch's CUDA div-by-python-scalar does).
 * ---------------------------------------------------------------------------------------- */
S2C_API int s2c_query_and_group(const float *xyz, const float *new_xyz, const float *features, int B, int n,
                        int M, int C, int feat_layout, long long feat_stride, float radius,
                        int nsample, int normalize_xyz, int out_layout, int *idx, float *grouped,
                        void *stream);

/* ------------------------------------------------------------------------------------------
 * knn_adjacency -- replaces the 256-iteration Python loop GraphModule._create_adjacent_mat /
 *   _query_locals (models/graph_module.py:182-233) and its copy in the caption module
 *   (models/caption_module.py:322-381): for every (scene, target) pair, the indicator row of the
 *   `num_locals` proposals closest to the target box.
 *     corners    (B,K,8,3) FLOAT64 box corners (data_dict["bbox_corner"])
 *     mask       (B,K) int64, 0 = invalid proposal (data_dict["bbox_mask"])
 *     targets    (B,T) int64 target proposal ids, or NULL: T must be K and target t is box t
 *     corner_mode   1: query_mode "corner" (min over the target's 8 corners), 0: "center"
 *     include_self  1: D[target] = 0, 0: D[target] = 1e30   (applied after the other masks)
 *     iou_threshold CONF.TRAIN.OVERLAID_THRESHOLD (0.5)
 *     adjacent   (B,T,K) f32 out, exactly num_locals ones per row
 *     neighbours (B,T,num_locals) int32 out (may be NULL): the selected ids in ascending order
 *   Distances are float64 like the reference's; ties (only the 1e30 sentinels in practice) go to
 *   the smaller index (torch.topk leaves them implementation-defined).  K <= 1024.
 * ---------------------------------------------------------------------------------------- */
S2C_API int s2c_knn_adjacency(const double *corners, const long long *mask, const long long *targets, int B,
                              int K, int T, int num_locals, int corner_mode, int include_self,
                              double iou_threshold, float *adjacent, int *neighbours, void *stream);

/* ------------------------------------------------------------------------------------------
 * mlp_layer_fwd -- one layer of the per-group shared MLP (SharedMLP: 1x1 Conv2d without bias, the
 *   BatchNorm + ReLU of the PREVIOUS layer folded into the operand load), on the tcgen05 tensor cores
 *   (3xTF32, fp32 accumulation in TMEM).  Replaces, per layer, the reference's cuDNN conv + BatchNorm2d
 *   + ReLU kernel triple (lib/pointnet2/pytorch_utils.py:88-120 as used at pointnet2_modules.py:251).
 *     A          (R, lda) fp32 rows = every (scene, group, sample); K valid columns
 *     pro_scale / pro_shift  [K] or both NULL: a' = relu(a * scale[k] + shift[k]) applied on load
 *     W          (N, K) fp32 row-major (Conv2d weight (Cout,Cin,1,1)); N multiple of 16, <= 256
 *     C          (R, ldc) fp32 out = a' * W^T   (PRE-BatchNorm output of this layer)
 *     stat_sum / stat_sumsq  [N] float64 or both NULL: += column sums of C and of C^2 (the batch
 *                statistics of this layer's BatchNorm); the caller zeroes them
 * ---------------------------------------------------------------------------------------- */
S2C_API int s2c_mlp_layer_fwd(const float *A, long long lda, long long R, int K, const float *pro_scale,
                              const float *pro_shift, const float *W, int N, float *C, long long ldc,
                              double *stat_sum, double *stat_sumsq, void *stream);

/* pool_fwd -- last BatchNorm + ReLU folded into the max over nsample (F.max_pool2d, pointnet2_modules.py:255-257):
 *   Y (G*ns, ldy) pre-BatchNorm rows, N channels -> out (G, N) = max_s relu(Y*scale+shift); argmax (G, N) int32
 *   (may be NULL) = first maximising sample, the element the backward pass routes the gradient to.  ns = 1 turns
 *   this into a plain BatchNorm + ReLU (PointnetFPModule). */
S2C_API int s2c_pool_fwd(const float *Y, long long ldy, long long G, int ns, int N, const float *scale,
                         const float *shift, float *out, int *argmax, void *stream);

/* pool_bwd_stats -- per-channel sums the last layer's BatchNorm backward needs, straight from the pooled gradient:
 *   sum_g[c] += sum_G g, sum_gy[c] += sum_G g * y  with g = dpool[G,c] where relu(bn(y)) was active at the arg-max
 *   element y = Y[G*ns + argmax[G,c], c], else 0.  float64 accumulators, zeroed by the caller. */
S2C_API int s2c_pool_bwd_stats(const float *dpool, const int *argmax, const float *Y, long long ldy, long long G,
                               int ns, int N, const float *scale, const float *shift, double *sum_g,
                               double *sum_gy, void *stream);

/* mlp_layer_fwd_v2 -- same contract as s2c_mlp_layer_fwd, warp-specialised and fed by the TMA engine
 *   (cp.async.bulk + mbarrier pipeline: loader / transform / MMA / epilogue warps, double-buffered TMEM).
 *   The A operand reaches the tensor core through TENSOR memory (transform warps: tcgen05.st, MMA: tcgen05.mma [d],[a],b);
 *   the environment variable S2C_MLP_ATM=0 selects the shared-memory operand path instead (A/B measurements only).
 *   Extra requirements: N in {64,128,256}; K, lda, ldc multiples of 4; A, C 16-byte aligned;
 *   wprep = workspace of ceil(K/32)*N*256 bytes (16-byte aligned) for the split / swizzled weights. */
S2C_API int s2c_mlp_layer_fwd_v2(const float *A, long long lda, long long R, int K, const float *pro_scale,
                                 const float *pro_shift, const float *W, int N, float *C, long long ldc,
                                 double *stat_sum, double *stat_sumsq, void *wprep, void *stream);

/* mlp_probe -- profiling aid of the layer kernels (mlp_layer_fwd_v2 / bwd_data / bwd_input): CTA 0 of every following
 *   launch stamps clock64() at its pipeline hand-offs (loader / transform / MMA / epilogue warps) into buf (device memory,
 *   `capacity` 8-byte slots; slot layout at S2C_PROBE in csrc/mlp2.cu).  buf = NULL switches it off (the default).
 *   No reference counterpart; used by tools/mlp_pipe_probe.py. */
S2C_API int s2c_mlp_probe(unsigned long long *buf, int capacity);

/* mlp_layer_bwd_data -- backward "data" pass of one shared-MLP layer l on the tensor cores (same pipeline as
 *   mlp_layer_fwd_v2), fusing BatchNorm backward, the 1x1-conv data gradient, the previous layer's ReLU mask and the
 *   reductions of the previous layer's BatchNorm backward:
 *       dY_l   = a[k]*g_l + b[k]*Y_l + c[k]              (BatchNorm-backward of layer l is affine per channel)
 *       g_prev = (dY_l * W_l) masked by relu(bn_{l-1}(Y_prev)) > 0                     -> C (R, ldc)
 *       stat_sum[n] += sum_r g_prev,  stat_sumsq[n] += sum_r g_prev * Y_prev           (float64; caller zeroes)
 *   g_l: dense G (R, ldg), or for the LAST layer rebuilt from the pooled gradient: dpool / argmax (R/ns, K) with
 *   last_scale/last_shift [K] = that layer's folded BatchNorm (its ReLU mask).  W (K, N) row-major = Conv2d weight
 *   of layer l (K = C_l, N = C_{l-1}); N in {64,128,256}; K and all leading dimensions multiples of 4.
 *   dY_out (R, K) optional: dY_l written back.  wprep: workspace of ceil(K/32)*N*256 bytes.
 *   Replaces autograd through BatchNorm2d/ReLU/Conv2d of pytorch_utils.py:88-120 (cuDNN dgrad + ATen BN backward). */
S2C_API int s2c_mlp_layer_bwd_data(const float *G, long long ldg, const float *Y, long long ldy, long long R, int K,
                                   const float *a, const float *b, const float *c, const float *dpool,
                                   const int *argmax, int ns, const float *last_scale, const float *last_shift,
                                   const float *W, int N, const float *Yprev, long long ldyp,
                                   const float *prev_scale, const float *prev_shift, float *C, long long ldc,
                                   float *dY_out, double *stat_sum, double *stat_sumsq, void *wprep, void *stream);

/* mlp_layer_bwd_weight -- weight gradient of one shared-MLP layer on the tensor cores (MN-major tcgen05 operands,
 *   3xTF32, per-CTA accumulation in TMEM, fp32 atomics into dW):
 *       dW[C x P] += sum_r dY[r,:]^T * X'[r,:]
 *   dY (R, lddy): dense, or -- when a/b/c [C] are given -- formed on the fly as a*dY_in + b*Y + c (BatchNorm
 *   backward of this layer from its masked upstream gradient and its pre-BN output Y);
 *   X (R, ldx): previous layer's pre-BN output with xs/xh [P] (X' = relu(X*xs+xh)), or the raw layer input (xs = NULL).
 *   C <= 256, P <= 288, ceil(C/128)*ceil(P/32) <= 16; leading dimensions multiples of 4.  The caller zero-fills dW.
 *   Replaces the cuDNN/cuBLAS weight-gradient GEMM of Conv2d in SharedMLP (pytorch_utils.py:88-95). */
S2C_API int s2c_mlp_layer_bwd_weight(const float *dY, long long lddy, const float *Y, long long ldy, const float *a,
                                     const float *b, const float *c, const float *X, long long ldx, const float *xs,
                                     const float *xh, long long R, int C, int P, float *dW, long long lddw,
                                     void *stream);

/* query_and_group_grid -- same contract and bit-identical results as s2c_query_and_group / s2c_ball_query, with a
 *   uniform-grid pre-filter (cell edge >= r): per scene a counting sort of the points by cell, then each centre tests
 *   only the points of the 27 surrounding cells and recovers "the first nsample indices in index order" from a
 *   per-warp bitmap.  ~200 distance tests per centre instead of n.  grouped may be NULL (ball query only), idx may be
 *   NULL (grouped only).  workspace: s2c_ball_query_grid_workspace_bytes(B, n) bytes of device memory. */
/* detection_loss -- the VoteNet detection loss of lib/loss_helper.py (compute_vote_loss :24-69, compute_objectness_loss
 *   :71-111, compute_box_and_sem_cls_loss :113-187 on utils/nn_distance.py:32-59) forward AND backward in one launch:
 *   det_loss = vote + 0.5*objectness + box + 0.1*sem_cls, box = center + 0.1*heading_cls + heading_reg + 0.1*size_cls
 *   + size_reg (:409, :472-476; the caller applies the x10).  net (B,K,W) are the proposal head's outputs per proposal,
 *   W = 2 + 3 + 2*NH + 4*NS + NC in the reference's channel order (proposal_module.py:105-144); center = agg_xyz +
 *   net[...,2:5] is passed (and its gradient returned) separately.  stats[16] = {det_loss, vote, objectness, center,
 *   heading_cls, heading_reg, size_cls, size_reg, sem_cls, box, obj_acc, pos_ratio, neg_ratio, 0, 0, 0}.  Label outputs
 *   (objectness_label / object_assignment int64, objectness_mask f32; (B,K)) use the reference's fp32 operation order.
 *   scratch: (B*K + B*G) ints. */
S2C_API int s2c_detection_loss(int B, int S, int N, int K, int G, int NH, int NS, int NC, const float *vote_xyz,
                               const float *seed_xyz, const int *seed_inds, long long seed_ld, const float *vote_label,
                               const long long *vote_label_mask, const float *agg_xyz, const float *net,
                               const float *center, const float *center_label, const long long *heading_class_label,
                               const float *heading_residual_label, const long long *size_class_label,
                               const float *size_residual_label, const long long *sem_cls_label,
                               const float *box_label_mask, const float *mean_size, float *stats,
                               long long *objectness_label, float *objectness_mask, long long *object_assignment,
                               float *d_vote_xyz, float *d_net, float *d_center, int *scratch, void *stream);

/* Post-processing of benchmark/predict.py:176-190 (lib/ap_helper.py:40-178 parse_predictions) on the device.
 * points_in_boxes_count: count[b,k] = number of points of scene b inside the axis-aligned box (min xyz, max xyz; f64)
 *   -- replaces extract_pc_in_box3d's scipy Delaunay hull per box (data/scannet/model_util_scannet.py:13-22; ScanNet
 *   boxes have heading 0, :130-134).  xyz_ld = floats between consecutive points (3 for (B,N,3), 3+C for point_clouds).
 * nms3d: greedy 3-D NMS of utils/nms.py (nms_3d_faster :57-107; same_class_only = nms_3d_faster_samecls :110-150),
 *   float64, only boxes with valid != 0 take part; keep[b,k] = 1 for the picked boxes. */
S2C_API int s2c_points_in_boxes_count(const float *xyz, long long xyz_ld, int B, int N, const double *boxes, int K,
                                      int *count, void *stream);
S2C_API int s2c_nms3d(const double *boxes, const double *score, const long long *cls, const int *valid, int B, int K,
                      double iou_threshold, int old_type, int same_class_only, int *keep, void *stream);

/* Tuning knob of the TMA gather epilogue of s2c_query_and_group_grid (ring geometry per warp): 0 = default
 * (8-row tiles x 3 per mover, 10 mover + 10 query warps), 1 = 8x4x8, 2 = 8x4x9, 3 = 8x5x7, 4 = 8x3x12, 5 = 16x3x6; -1 = disable the TMA path (LDG/STG epilogue; used by the
 * tests to cross-check the two epilogues bit for bit).  Process-wide; not part of the reference surface. */
S2C_API int s2c_query_and_group_grid_tune(int variant);
S2C_API long long s2c_ball_query_grid_workspace_bytes(int B, int n);
S2C_API int s2c_query_and_group_grid(const float *xyz, const float *new_xyz, const float *features, int B, int n,
                                     int M, int C, int feat_layout, long long feat_stride, float radius,
                                     int nsample, int normalize_xyz, int out_layout, int *idx, float *grouped,
                                     void *workspace, long long workspace_bytes, void *stream);

/* ball_query_grid_build / query_and_group_grid_prebuilt -- the two halves of s2c_query_and_group_grid.  The uniform grid
 *   depends only on (xyz, radius), so a caller that has the next batch's coordinates can build it ahead of time (e.g. on
 *   a copy stream during the previous training step) and keep only the query / gather kernel on the critical path.
 *   `workspace` (s2c_ball_query_grid_workspace_bytes(B, n) bytes) carries the grid; it holds no absolute pointers and
 *   may be copied between equally aligned buffers.  Same arguments, same bit-identical results as the one-call form
 *   (ball_query_gpu.cu:9-44 + group_points_gpu.cu:8-64 + pointnet2_utils.py:317-376). */
S2C_API int s2c_ball_query_grid_build(const float *xyz, int B, int n, float radius, void *workspace,
                                      long long workspace_bytes, void *stream);
S2C_API int s2c_query_and_group_grid_prebuilt(const float *xyz, const float *new_xyz, const float *features, int B, int n,
                                              int M, int C, int feat_layout, long long feat_stride, float radius,
                                              int nsample, int normalize_xyz, int out_layout, int *idx, float *grouped,
                                              void *workspace, long long workspace_bytes, void *stream);

/* bn_finalize -- per-layer BatchNorm bookkeeping of the fused shared-MLP path in ONE launch (replaces nn.BatchNorm2d's
 *   statistics handling, lib/pointnet2/pytorch_utils.py:88-120 / torch batch_norm):
 *   use_batch_stats: mean = sum/R, var = max(sumsq/R - mean^2, 0) from the float64 column sums of the GEMM epilogue,
 *   else the running statistics; update_running: running = (1-momentum)*running + momentum*{mean, unbiased var},
 *   ++num_batches_tracked (may be NULL); momentum < 0 = nn.BatchNorm(momentum=None): cumulative moving average with
 *   factor 1/num_batches_tracked (value after the increment; the counter is then required).  Outputs [N]: mean, invstd (float64) and the folded fp32 affine
 *   scale = gamma*invstd, shift = beta - mean*scale that the next kernel applies in its operand prologue. */
S2C_API int s2c_bn_finalize(const double *sum, const double *sumsq, long long R, int N, const float *gamma,
                            const float *beta, double eps, double momentum, int use_batch_stats, int update_running,
                            float *running_mean, float *running_var, long long *num_batches_tracked, double *mean,
                            double *invstd, float *scale, float *shift, void *stream);

/* bn_backward_coeffs -- BatchNorm backward of one layer as the per-channel affine map dY = a*g + b*y + c
 *   (batch statistics; b = c = 0 in evaluation mode) plus grad_gamma = sum g*xhat and grad_beta = sum g, from the
 *   float64 sums (sum g, sum g*y) a backward GEMM / pooling epilogue produced.  All arrays [N]. */
S2C_API int s2c_bn_backward_coeffs(const double *sum_g, const double *sum_gy, const double *mean,
                                   const double *invstd, const float *gamma, long long R, int N, int batch_stats,
                                   float *grad_gamma, float *grad_beta, float *a, float *b, float *c, void *stream);

/* group_rows_grad -- gradient of grouped rows w.r.t. a point-major tensor (group_points_grad_kernel,
 *   group_points_gpu.cu:43-64, on the channels-last layout of the fused path):
 *   rows (B, T, ld) gradient of the grouped tensor, channels [c0, c0+C) wanted; idx (B, T) the neighbour list;
 *   out (B, n, C) = scale * sum over t with idx[b,t]==k of rows[b,t,c0:c0+C]  (zero-filled here, fp32 atomics --
 *   one red.global.add.v4.f32 per 4 channels when C is a multiple of 4 and out is 16-byte aligned). */
S2C_API int s2c_group_rows_grad(const float *rows, long long ld, int c0, int C, const int *idx, int B, long long T,
                                int n, float scale, float *out, void *stream);

/* caption_decode_fwd / caption_decode_bwd -- the teacher-forced top-down caption decoder of
 *   TopDownSceneCaptionModule.forward_sample_batch (models/caption_module.py:428-500, step :250-292) for all T
 *   words in ONE launch per direction (a thread-block cluster walks the recurrence; see csrc/caption.cu).
 *   Sizes: B scenes, T words, K proposals, E = emb_size, H = hidden_size, F = feat_size (E, F multiples of 4, H of 64).
 *   Inputs (fp32): pre_word (B,T,E) = W_td[:, :E] w_t ; pre_tgt (B,E) = W_td[:, E+H:] target + b_td ;
 *     mapped (B,K,H) = map_feat(obj) ; obj (B,K,F) ; valid (B,K) 0/1 local-context mask ;
 *     w_tdh = W_td[:, E:E+H] (E rows, stride ld_tdh) ; GRU cells (w_ih (3H,E), w_hh (3H,H), b_ih, b_hh (3H)) ;
 *     w_hidd (H,H) ; w_att (H) ; w_lang (E,F+H), b_lang (E).
 *   Per-step outputs, all (T,B,.): u, lang (E) ; h1, r1, z1, n1, hn1, q, r2, z2, n2, hn2, h2 (H) ; probs (K) =
 *     softmax over the valid proposals (exact zeros elsewhere; uniform 1/K for a scene with no valid proposal,
 *     as softmax of K equal -1e30 scores) ; att (F).  h2 is the decoder output, probs the attention map.
 *   Backward: d_h2 (T,B,H) and optional d_probs (T,B,K) in; transposed weights wt_* (row-major W^T) in;
 *     per-step gradients out, (T,B,.): dgi2, dgh2, dgi1, dgh1 (3H) ; dlang, du (E) ; datt (F) ; dq (H) ;
 *     d_mapped (B,K,H) and d_obj (B,K,F) are ACCUMULATED INTO (zero-fill them first) ;
 *     d_watt (ceil(B/8), H): per-cluster partial sums of d w_att.  The weight gradients are GEMMs over those
 *     (T*B)-row stacks and are left to the caller. */
typedef struct s2c_caption_params {
  int B, T, K, E, H, F;
  long long ld_tdh;
  const float *pre_word, *pre_tgt, *mapped, *obj, *valid;
  const float *w_tdh, *w_ih1, *w_hh1, *b_ih1, *b_hh1, *w_hidd, *w_att, *w_lang, *b_lang, *w_ih2, *w_hh2, *b_ih2, *b_hh2;
  float *u, *h1, *r1, *z1, *n1, *hn1, *q, *probs, *att, *lang, *r2, *z2, *n2, *hn2, *h2;
  float *scores; /* (T,B,K) scratch: raw attention scores exchanged between the CTAs of the cluster */
  /* backward only */
  const float *wt_tdh, *wt_ih1, *wt_hh1, *wt_hidd, *wt_lang, *wt_ih2, *wt_hh2;
  const float *d_h2, *d_probs;
  float *dgi2, *dgh2, *dlang, *datt, *dq, *dgi1, *dgh1, *du, *d_mapped, *d_obj, *d_watt;
  /* optional profiling aid: (T, 8) %globaltimer stamps (ns) taken by CTA 0 at the stage boundaries of every word */
  long long *dbg_ts;
  /* optional: one zero-initialised unsigned int of device memory.  When given (and B <= 8, H/4 <= #SMs, cooperative
   * launch available) the recurrence runs as a persistent cooperative grid with the weights resident in shared memory
   * (csrc/caption_grid.cu) and this word is its grid-barrier counter; otherwise the cluster kernels are used. */
  unsigned int *grid_bar;
} s2c_caption_params;
S2C_API int s2c_caption_decode_fwd(const s2c_caption_params *params, void *stream);
S2C_API int s2c_caption_decode_bwd(const s2c_caption_params *params, void *stream);

/* gemm_tn -- out (M,N) = A^T X with A (R, lda), X (R, ldx) fp32 row-major and a SHORT reduction length R (the T*B rows
 *   of the caption decoder): the weight gradients dW = dGates^T Inputs that autograd computes with cuBLAS after the
 *   recurrence (caption_module.py:428-500).  colsum (M), optional: column sums of A (the bias gradients). */
S2C_API int s2c_gemm_tn(const float *A, long long lda, const float *X, long long ldx, int R, int M, int N, float *out,
                        long long ldo, float *colsum, void *stream);

/* gemm -- C (M,N; ldc) = A (M,K) * B (K,N) [+ bias[n]] [ReLU], plain fp32 FMAs, operands as strided views
 *   A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn].  The nn.Linear layers of the caption module whose widths
 *   (300 / 812 / 3500) are not multiples of the tensor-core tiles -- map_feat, the word / target terms of map_topdown,
 *   classifier (models/caption_module.py:216-240) -- forward (B = W^T: sbk = 1, sbn = ldw) and input gradient
 *   (B = W: sbk = ldw, sbn = 1); their weight gradients are s2c_gemm_tn.  Replaces the framework's cuBLAS calls. */
S2C_API int s2c_gemm(const float *A, long long sam, long long sak, const float *B, long long sbk, long long sbn,
                     const float *bias, int relu, int M, int N, int K, float *C, long long ldc, void *stream);

/* mlp_layer_bwd_input -- input gradient of the FIRST layer of a shared MLP on the tensor cores:
 *       C (R, N; ldc) = (a*G + b*Y + c) * W[:, block of N columns]
 *   G (R, ldg) the layer's masked upstream gradient, Y (R, ldy) its pre-BatchNorm output, a/b/c [K] from
 *   s2c_bn_backward_coeffs, W the Conv2d weight (K rows, row stride ldw) offset by the caller to the wanted block of
 *   input columns; N in {64,128,256}.  dY_out (R, K) optional.  wprep: ceil(K/32)*N*256 bytes.
 *   Replaces the dgrad of SharedMLP's first Conv2d (pytorch_utils.py:88-95 under autograd). */
S2C_API int s2c_mlp_layer_bwd_input(const float *G, long long ldg, const float *Y, long long ldy, long long R, int K,
                                    const float *a, const float *b, const float *c, const float *W, long long ldw, int N,
                                    float *C, long long ldc, float *dY_out, void *wprep, void *stream);

/* col_sum -- out[c] = sum_r A[r, c] for a tall fp32 matrix A (R, lda), M columns: the bias gradient of a Linear /
 *   Conv1d layer (ATen's sum(0) under autograd).  out is zero-filled here (fp32 atomics over row chunks). */
S2C_API int s2c_col_sum(const float *A, long long lda, long long R, int M, float *out, void *stream);

/* ------------------------------------------------------------------------------------------
 * edgeconv_fwd / edgeconv_bwd -- one EdgeConv layer of the relational graph (models/graph_module.py:22-115:
 *   EdgeConv.message :102-109 = Linear(2*Cin, Cout) -> ReLU -> Linear(Cout, Cout) on [x_i, x_j - x_i];
 *   MessagePassing.propagate :44-100 with aggr = "add") over ONE batched graph.  Replaces, per layer, PyG's
 *   index_select x2 + cat + two cuBLAS GEMMs + scatter-add and their autograd backward.
 *     x          (Nn, Cin) fp32 node features; row / col (E) int64 = edge_index[0] / edge_index[1]
 *                (x_j = x[row[e]], x_i = x[col[e]], the message is aggregated at col[e])
 *     edge_mask  (E) bytes or NULL: 0 = the slot is not an edge (message 0, no gradient)
 *     W1 (Cout, 2*Cin), b1 (Cout), W2 (Cout, Cout), b2 (Cout)  = map_edge.0 / map_edge.2 of the reference
 *     z (E, 2*Cin), Y1 (E, Cout)   out, kept for backward: the gathered edge input and the pre-ReLU hidden layer
 *     msg (E, Cout)  out: the masked messages (what propagate returns as `message`)
 *     agg (Nn, Cout) out or NULL: sum of the messages at col[e] (zero-filled here)
 *   Cout in {64,128,256}; 2*Cin a multiple of 64, <= 512.  workspace: s2c_edgeconv_workspace_bytes(E, Cin, Cout,
 *   backward) bytes, 256-byte aligned.  GEMMs on the tcgen05 kernels (3xTF32, fp32 accumulate).
 *   bwd: dagg (Nn, Cout) / dmsg (E, Cout) = gradients of the two outputs (either may be NULL);
 *        dx (Nn, Cin) or NULL, dW1, db1, dW2, db2 are overwritten. */
S2C_API long long s2c_edgeconv_workspace_bytes(long long E, int Cin, int Cout, int backward);
S2C_API int s2c_edgeconv_fwd(const float *x, long long Nn, int Cin, const long long *row, const long long *col,
                             const unsigned char *edge_mask, long long E, const float *W1, const float *b1,
                             const float *W2, const float *b2, int Cout, float *z, float *Y1, float *msg, float *agg,
                             void *workspace, void *stream);
S2C_API int s2c_edgeconv_bwd(const float *dagg, const float *dmsg, long long Nn, int Cin, const long long *row,
                             const long long *col, const unsigned char *edge_mask, long long E, const float *W1,
                             const float *b1, const float *W2, int Cout, const float *z, const float *Y1, float *dx,
                             float *dW1, float *db1, float *dW2, float *db2, void *workspace, void *stream);

/* adam_step -- torch.optim.Adam (amsgrad=False, maximize=False) over ONE flat fp32 buffer each of parameters, gradients and
 *   the two moments: the optimizer.step() of the reference's training loop (lib/solver.py:293-300; Adam created at
 *   scripts/train.py:134) as one streaming pass instead of ~27 multi-tensor launches.
 *     n        elements (multiple of 4; the caller pads the flat buffers), all four buffers 16-byte aligned
 *     hyper    device [lr, beta1, beta2, eps, weight_decay] (read at run time: a captured graph follows lr changes)
 *     steps    device [nsteps] float step counters (torch keeps one per parameter tensor, all equal): incremented here
 *     coef     device [2] scratch
 *     grad_scale  multiplies the gradient first (1/world for a summed all-reduce, else 1). */
S2C_API int s2c_adam_step(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, long long n,
                          const float *hyper, float *steps, int nsteps, float *coef, float grad_scale, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* S2C_H_ */
