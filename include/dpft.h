/*
 * dpft.h -- C ABI of the B200-native trust-region inverse-compositional solver.
 *
 * This is the drop-in boundary for ONE path of smartroboticslab/deep_prob_feature_track:
 * the per-pyramid-level Gauss-Newton solve that LeastSquareTracking.forward runs.  The
 * reference has no FFI of its own (it is pure PyTorch); the interface it exposes for this
 * path is the nn.Module surface
 *
 *   TrustRegionInverseWUncertainty.forward / .forward_residuals   code/models/algorithms.py:611, :725
 *   TrustRegionBase.forward / .forward_residuals                  code/models/algorithms.py:45,  :123
 *   DirectSolverNet.forward                                       code/models/algorithms.py:1604
 *   LeastSquareTracking.forward (coarse-to-fine chain)            code/models/LeastSquareTracking.py:345-446
 *
 * and the Python modules in deep_prob_feature_track_b200/algorithms.py mirror it one to one
 * and call the entry points below through ctypes (INTEGRATION.md shows the binding).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous fp32 (or uint8 for masks) unless noted;
 *   - feature maps are NCHW: (B,C,H,W); depth / inverse-depth / masks are (B,1,H,W); K is (B,4)
 *     = [fx, fy, cx, cy] ALREADY scaled to the level (the reference passes K / 2^level);
 *   - poses are R (B,3,3) row-major and t (B,3); "pose rows" pack them as 12 floats R|t;
 *   - `stream` is a cudaStream_t passed as void*; nothing here allocates or synchronises;
 *   - every function returns 0 on success or a negative DPFT_E* / positive cudaError_t code.
 *     dpft_last_error() gives a message for the calling thread.
 *   - levels are listed in the order they are solved: COARSE FIRST.
 */
#ifndef DPFT_H_
#define DPFT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPFT_ABI_VERSION 1
#define DPFT_MAX_LEVELS 8

/* error codes */
#define DPFT_EINVAL   (-1)  /* bad argument (shape, NULL pointer, unsupported flag mix: DPFT_SHARED_KEYFRAME or
                               DPFT_PAIRWISE_EXTREMES with DPFT_COMBINE_ICP or without DPFT_FUSED_SOBEL;
                               DPFT_SIGMA_BROADCAST without DPFT_FUSED_SOBEL; occ_out with DPFT_PAIRWISE_EXTREMES) */
#define DPFT_ENOSPACE (-2)  /* workspace too small */

/* flags */
#define DPFT_REMOVE_TRU_SIGMA 0x01u /* alg:1974-1979  mask pixels whose sigma sits on the batch-global min/max */
#define DPFT_COMBINE_ICP      0x02u /* alg:668-689    add the point-to-plane term (needs depth0/depth1)        */
#define DPFT_NO_PDL           0x04u /* launch iterations without programmatic dependent launch (debug)        */
#define DPFT_LAUNCH_PER_ITERATION 0x10u /* one launch per Gauss-Newton iteration instead of the single
                                          cooperative launch (the ICP term, occ_out and the materialised
                                          gradients always use it)                                           */
#define DPFT_STAGED_FOOTPRINT 0x20u /* launch-per-iteration path: levels with C == 8, W % 4 == 0, W >= 60 and 16-byte
                                       aligned x1 / sigma1 / invd1 run the kernel that stages the lookup footprint
                                       in shared memory (cp.async rows); other levels are unaffected              */
#define DPFT_SIGMA_BROADCAST  0x100u /* dpft_uic_forward: sigma0 / sigma1 are (B,1,H,W) -- the ONE uncertainty map per
                                        frame that the reference's encoder emits and then repeats to C channels
                                        (alg:1425-1427, uncertainty_channel = 1); results are those of the repeated
                                        tensors, the map is read, warped and differentiated once per pixel.
                                        Fused launch-per-iteration kernels only.                               */
#define DPFT_SHARED_KEYFRAME  0x40u /* x0, sigma0, invd0 (and obj_mask0) have batch size 1: every pair of the
                                       batch tracks against the same keyframe (kf_vo.py keyframe mode); forward only */
#define DPFT_PAIRWISE_EXTREMES 0x80u /* with DPFT_REMOVE_TRU_SIGMA: sigma extremes per pair, i.e. the semantics of
                                        calling the reference once per pair with B = 1 (kf_vo.py:156-166); forward only */
#define DPFT_QUEUE            0x200u /* dpft_uic_forward: the finest level (all its iterations) as ONE launch whose warps
                                        take warp tiles from a work queue; dependencies are per frame pair, not per
                                        launch (csrc/uic_queue.cu).  It pays when the call holds several waves of tiles
                                        and the pairs are not all coupled: several groups (dpft_uic_options.group), or
                                        no DPFT_REMOVE_TRU_SIGMA, or pairwise extremes.  Taken when the problem
                                        qualifies (C == 8, DPFT_FUSED_SOBEL, no DPFT_COMBINE_ICP, no occ_out,
                                        iters >= 1), ignored otherwise.  Same results as the launch-per-iteration path
                                        up to summation order.                                                        */
#define DPFT_FUSED_SOBEL      0x08u /* recompute the unit Sobel gradients inside every iteration (sliding register
                                       window) instead of materialising them once per level                    */

/* status bits written to *status (device int32, OR-ed; zero it before the call) */
#define DPFT_ST_NONFINITE 0x01 /* a weighted residual / normal-equation entry was NaN or Inf (alg:886,1988) */
#define DPFT_ST_SINGULAR  0x02 /* a damped 6x6 system was not positive definite                              */

/* One pyramid level of one batch of frame pairs. */
typedef struct dpft_level {
  const float *x0, *x1;         /* (B,C,H,W) keyframe / live feature maps                         */
  const float *sigma0, *sigma1; /* (B,C,H,W) their uncertainty maps (U_IC only, else NULL)        */
  const float *invd0, *invd1;   /* (B,1,H,W) inverse depth                                        */
  const float *depth0, *depth1; /* (B,1,H,W) metric depth, only with DPFT_COMBINE_ICP, else NULL  */
  const float *K;               /* (B,4) intrinsics at this level                                 */
  const uint8_t *obj_mask0;     /* (B,1,H,W) optional object mask of the keyframe (1 = object)    */
  const uint8_t *obj_mask1;     /* (B,1,H,W) optional object mask of the live frame               */
  uint8_t *occ_out;             /* optional (iters,B,H,W): the validity mask of every iteration
                                   (1 = excluded), exactly what compute_inverse_residuals returns */
  int32_t H, W;
} dpft_level_t;

/*
 * Optional knobs of dpft_uic_forward_ex / dpft_uic_workspace_bytes_ex.  Zero-initialise, set struct_bytes =
 * sizeof(dpft_uic_options_t), then the fields you want; a zero field means "default".  The library keeps no copy.
 */
typedef struct dpft_uic_options {
  uint32_t struct_bytes;
  int32_t group;          /* DPFT_REMOVE_TRU_SIGMA: pairs per sigma-extreme group.  The B pairs of the call are
                             `B / group` independent batches of `group` consecutive pairs, each with the reference's
                             batch-global extremes (alg:1976-1979) -- the results of B / group separate calls, from
                             one call.  0 = B (one batch); 1 = DPFT_PAIRWISE_EXTREMES.  B % group == 0; groups other
                             than B need DPFT_FUSED_SOBEL and exclude DPFT_COMBINE_ICP and occ_out.
                             aux_hist then has (n_levels*iters, B / group, 4) entries.                            */
  int32_t tile_rows[DPFT_MAX_LEVELS]; /* DPFT_QUEUE: rows per warp tile at level l (coarse first); 0 = chosen   */
  int32_t queue_ctas;     /* DPFT_QUEUE: CTAs (4 worker warps each) to launch; 0 = what the device holds        */
  int32_t queue_levels;   /* DPFT_QUEUE: how many of the finest levels run as work-queue launches (one launch per
                             level); 0 = 1.  The coarser levels keep one launch per iteration.                   */
  int32_t cta_slots;      /* launch-per-iteration kernels: resident CTA slots the tile heights are planned for  */
  int32_t tiling;         /* staged kernel: 0 balanced dealt tiles, 1 rectangular, 2 balanced linear ranges     */
  int32_t generic_geometry; /* 1: never pick the instantiations specialised for 160x120 / 80x60 levels          */
  float *launch_ms;       /* HOST array (n_levels*iters) or NULL.  When set, the call measures the device time of every
                             Gauss-Newton iteration (CUDA events around every launch, or %globaltimer stamps of the
                             iteration completions inside a single launch), WAITS for the stream and fills the array:
                             a measurement aid, not a way to run the solver.                                      */
  const float *icp_weight[DPFT_MAX_LEVELS]; /* DPFT_COMBINE_ICP: per level (coarse first) a DEVICE map (B,1,H,W) that
                             scales the point-to-plane residual and Jacobian pixel by pixel -- the output of a learned
                             ScaleNet evaluated at the level's first iteration (alg:677-682) -- or NULL: the scalar w_icp. */
  float *queue_kernel_ms; /* HOST array (queue_levels entries, coarse level first) or NULL; only read together with
                             launch_ms.  Receives the device time of each work-queue KERNEL alone: CUDA events recorded on
                             `stream` right before and right after its launch (its helper launches -- queue init, sigma0
                             extremes -- stay outside).                                                          */
  int32_t small_levels;   /* launch-per-iteration kernels, levels whose live frame (x1, sigma1, invd1 of one pair) fits
                             the shared memory of a CTA twice per SM: 0 = stage it there and look the footprint up on
                             chip (the default), 1 = global-memory lookups as on any other level (a measurement knob;
                             same masks bit for bit, same sums up to the order of the fp32 partial sums).
                             2 = levels narrower than 44 columns that run as work-queue launches (queue_levels) keep the
                             plain tile routine instead of the staged routine's narrow form (a measurement knob too). */
  int32_t sigma_detect;   /* full (B,C,H,W) sigma tensors whose C channels are copies of channel 0 -- what the reference's
                             encoder hands over (alg:1425-1427) -- are found by a device-side check at the start of the call
                             and served by the one-map tile routines (same results; no host synchronisation: kernels for
                             both cases are enqueued and the ones that do not apply return at once).  0 = check (default;
                             costs one read of the sigma tensors), 1 = do not check.  Fused kernels without
                             DPFT_COMBINE_ICP and occ_out only.                                                   */
} dpft_uic_options_t;

/* ABI version of the loaded library (== DPFT_ABI_VERSION). */
int dpft_abi_version(void);

/* Message of the last error on this thread ("" if none). */
const char *dpft_last_error(void);

/* Bytes of device scratch the calls below need for this problem (same arguments). */
size_t dpft_uic_workspace_bytes(const dpft_level_t *levels, int n_levels, int B, int C, int iters,
                                uint32_t flags);

/*
 * U_IC coarse-to-fine solve: for every level (coarse first) run `iters` Gauss-Newton iterations of
 * TrustRegionInverseWUncertainty.forward (alg:611-723): SE(3) warp in inverse depth, bilinear lookup
 * of x1 / sigma1 / invd1, validity mask, uncertainty-normalised residual, C x 6 Jacobian, J^T J and
 * J^T r reduced over the pair, H = J^T J + 1e-6 tr I, xi = H^-1 b, inverse-compositional update.
 * n_levels = 1 is exactly one module call; n_levels = 4 is the chain LeastSquareTracking.forward runs.
 *
 *   pose_in    (B,12)                       starting pose rows
 *   pose_hist  (n_levels*iters + 1, B, 12)  OUT: pose before iteration k at row k; the result is the last row
 *   sys_hist   (n_levels*iters, B, 27)      OUT: per iteration the 21 upper-triangular entries of J^T W J
 *                                           (row-major i<=j) followed by the 6 entries of J^T W r
 *   aux_hist   (n_levels*iters, 4)          OUT, optional (NULL): with DPFT_REMOVE_TRU_SIGMA the batch-global
 *                                           [min, max] of the warped sigma of that iteration and [min, max] of
 *                                           sigma0 -- the values dpft_uic_backward needs to rebuild the masks
 *   w_icp      scalar weight of the ICP term (the reference's ScaleNet('None') constant, 0.01)
 *   status     device int32, OR-ed with DPFT_ST_* bits
 */
int dpft_uic_forward(const dpft_level_t *levels, int n_levels, int B, int C, int iters, uint32_t flags,
                     float w_icp, const float *pose_in, float *pose_hist, float *sys_hist, float *aux_hist,
                     int32_t *status, void *workspace, size_t workspace_bytes, void *stream);

/* dpft_uic_workspace_bytes / dpft_uic_forward with options (opt may be NULL: identical to the plain calls). */
size_t dpft_uic_workspace_bytes_ex(const dpft_level_t *levels, int n_levels, int B, int C, int iters,
                                   uint32_t flags, const dpft_uic_options_t *opt);
int dpft_uic_forward_ex(const dpft_level_t *levels, int n_levels, int B, int C, int iters, uint32_t flags,
                        float w_icp, const float *pose_in, float *pose_hist, float *sys_hist, float *aux_hist,
                        int32_t *status, void *workspace, size_t workspace_bytes, void *stream,
                        const dpft_uic_options_t *opt);

/*
 * Depth stage of LeastSquareTracking._preprocess (LeastSquareTracking.py:656-661, 668-674; ImagePyramids
 * alg:1201-1219): invd = clamp(1/depth, 0, 10) with the pixels on the batch-global minimum, then on the
 * batch-global maximum, zeroed; max-pooled pyramids (kernel = stride = 2^l, floor sizes) of invd and, when
 * depth_out is not NULL, of depth.  Outputs are listed FINE level first (level l has (H >> l) x (W >> l) pixels),
 * each (B,1,h,w).  workspace: 8 bytes.
 */
int dpft_preprocess_depth(const float *depth, int B, int H, int W, int n_levels, float *const *invd_out,
                          float *const *depth_out, void *workspace, size_t workspace_bytes, void *stream);

/*
 * IC tracker (TrustRegionBase alg:45-139 + DirectSolverNet alg:1604-1691), split around the learned networks
 * the reference calls inside the loop.  `level` needs x0, x1, invd0, invd1, K (+ optional object masks).
 *   dpft_ic_gradients      gx, gy (B,C,H,W) <- unit Sobel gradient of x0 (feature_gradient, alg:1844-1865)
 *   dpft_ic_residual       r_out (B,C,H,W) <- x1(warp) - x0, 1e-3 where masked; occ_out (B,H,W) the mask
 *                          (compute_warped_residual, alg:1919-1957)
 *   dpft_ic_normal_matrix  A21 (B,21) <- upper triangle of sum w J J^T, J_c = gx_c du/dxi + gy_c dv/dxi;
 *                          weights (B,C,H,W) or NULL for ones (alg:71-75)
 *   dpft_ic_rhs            rhs (S,B,6) <- sum w J r(pose_s) for poses (S,B,12) (alg:1623-1624, 1682-1683)
 *   dpft_ic_update         pose_out (S,B,12) <- inverse-compositional update of pose_in (B,12) with
 *                          xi = H^-1 rhs (rhs: (B,6)), H = A + D + 1e-6 tr(A) I and
 *                            mode 0: D = 0 (S = 1)                      lev_mar_H, alg:2094-2103
 *                            mode 1: D = lambdas[s] diag(A)              residual-volume trials, alg:1676-1680
 *                            mode 2: D = diag(damp[b]), damp (B,6), S=1  learned damping, alg:1688-1691
 *                          H_out (S,B,21) optional.
 */
int dpft_ic_gradients(const dpft_level_t *level, int B, int C, float *gx, float *gy, void *stream);
int dpft_ic_residual(const dpft_level_t *level, int B, int C, const float *pose, float *r_out, uint8_t *occ_out,
                     void *stream);
int dpft_ic_normal_matrix(const dpft_level_t *level, int B, int C, const float *gx, const float *gy,
                          const float *weights, float *A21, void *stream);
int dpft_ic_rhs(const dpft_level_t *level, int B, int C, const float *gx, const float *gy, const float *weights,
                const float *poses, int S, float *rhs, void *stream);
int dpft_ic_update(int B, int S, int mode, const float *A21, const float *rhs, const float *lambdas,
                   const float *damp, const float *pose_in, float *pose_out, float *H_out, int32_t *status,
                   void *stream);

/*
 * Adjoints of the five entry points above, for training the IC tracker (reference: torch.autograd over
 * alg:45-121, 1604-1691, 1919-1957).  g_* inputs are gradients w.r.t. the matching forward output; outputs
 * marked (+=) are accumulated with atomics and must be zeroed (or hold a running sum) before the call, the
 * others are overwritten.  Masks and the 1e-3 fill are stop-gradients.
 *   dpft_ic_gradients_backward      g_x0 (+=)  <- g_gx, g_gy
 *   dpft_ic_residual_backward       g_x0, g_x1 (+=), g_pose (B,12) (+=)  <- g_r; pose as in the forward call
 *   dpft_ic_normal_matrix_backward  g_gx, g_gy, g_weights (NULL if weights NULL)  <- g_A21 (B,21)
 *   dpft_ic_rhs_backward            g_gx, g_gy, g_weights, g_x0, g_x1 (+=), g_poses (S,B,12) (+=)  <- g_rhs (S,B,6)
 *   dpft_ic_update_backward         g_A21 (B,21), g_rhs (B,6), g_damp (B,6; mode 2, may be NULL),
 *                                   g_pose_in (B,12)  <- g_pose_out (S,B,12)
 */
int dpft_ic_gradients_backward(const dpft_level_t *level, int B, int C, const float *g_gx, const float *g_gy,
                               float *g_x0, void *stream);
int dpft_ic_residual_backward(const dpft_level_t *level, int B, int C, const float *pose, const float *g_r,
                              float *g_x0, float *g_x1, float *g_pose, void *stream);
int dpft_ic_normal_matrix_backward(const dpft_level_t *level, int B, int C, const float *gx, const float *gy,
                                   const float *weights, const float *g_A21, float *g_gx, float *g_gy,
                                   float *g_weights, void *stream);
int dpft_ic_rhs_backward(const dpft_level_t *level, int B, int C, const float *gx, const float *gy,
                         const float *weights, const float *poses, int S, const float *g_rhs, float *g_gx,
                         float *g_gy, float *g_weights, float *g_x0, float *g_x1, float *g_poses, void *stream);
int dpft_ic_update_backward(int B, int S, int mode, const float *A21, const float *rhs, const float *lambdas,
                            const float *damp, const float *pose_in, const float *g_pose_out, float *g_A21,
                            float *g_rhs, float *g_damp, float *g_pose_in, void *stream);

/*
 * Context tensors of the learned networks the reference calls inside the loop, written by the kernels
 * that evaluate the residual:
 *   dpft_ic_context       context (B,4,H,W) = [ |r|, x0, x1, bilinear_up(w_prior) ]: the input of
 *                         DeepRobustEstimator('MultiScale2w') (alg:1471-1474) for a ONE-channel IC level at `pose` (B,12);
 *                         w_prior (B,1,hp,wp) or NULL (ones); occ_out (B,H,W) optional.
 *   dpft_uic_icp_context  icp_r (B,1,H,W): point-to-plane residual / sigma_icp, 1e-6 where its mask is set, and
 *                         feat_norm (B,1,H,W) = sqrt(sum_c wres_c^2) of the masked uncertainty-weighted feature residual:
 *                         all ScaleNet (alg:1535-1567) reads of its two residual inputs.  Needs depth0 / depth1.
 *                         Workspace: dpft_uic_residual_workspace_bytes with DPFT_COMBINE_ICP.
 */
int dpft_ic_context(const dpft_level_t *level, int B, const float *pose, const float *w_prior, int hp, int wp,
                    float *context, uint8_t *occ_out, void *stream);
int dpft_uic_icp_context(const dpft_level_t *level, int B, int C, uint32_t flags, const float *pose, float *icp_r,
                         float *feat_norm, void *workspace, size_t workspace_bytes, void *stream);

/*
 * forward_residuals of the U_IC tracker (alg:725-786 with compute_avg_loss alg:2119-2137): per frame pair the
 * sum over valid pixels of the squared uncertainty-weighted residuals (plus the squared weighted point-to-plane
 * residual with DPFT_COMBINE_ICP) divided by the number of valid pixels, at the given pose.  loss: (B).
 */
size_t dpft_uic_residual_workspace_bytes(const dpft_level_t *level, int B, int C, uint32_t flags);
int dpft_uic_residual_loss(const dpft_level_t *level, int B, int C, uint32_t flags, float w_icp,
                           const float *pose, float *loss, void *workspace, size_t workspace_bytes, void *stream);

/* Gradient maps of one level, same shapes as the inputs; the backward ACCUMULATES into them (zero them first). */
typedef struct dpft_level_grad {
  float *g_x0, *g_x1;         /* (B,C,H,W) */
  float *g_sigma0, *g_sigma1; /* (B,C,H,W) */
} dpft_level_grad_t;

size_t dpft_uic_backward_workspace_bytes(const dpft_level_t *levels, int n_levels, int B, int C, int iters,
                                         uint32_t flags);

/*
 * Reverse-mode of dpft_uic_forward (what torch.autograd derives from alg:611-723):
 * gradients w.r.t. x0, x1, sigma0, sigma1 of every level and w.r.t. the starting pose.  Masks, the 1e-6
 * fill and the batch-global extremes are stop-gradients.  Recompute-based: takes the forward's inputs and
 * its pose_hist / sys_hist / aux_hist, nothing per-pixel is saved.
 *
 *   grad_pose_hist (n_levels*iters + 1, B, 12)  IN: dL/d(pose row k); zero for rows that are not outputs
 *                                               (the outputs are the rows (l+1)*iters, one per level)
 *   grad_A         (n_levels, B, 36) or NULL    IN: dL/d(J^T W J of the last iteration of level l) (uncer_prop)
 *   grad_pose_in   (B,12)                       OUT: dL/d(pose_in)
 */
int dpft_uic_backward(const dpft_level_t *levels, const dpft_level_grad_t *grads, int n_levels, int B, int C,
                      int iters, uint32_t flags, float w_icp, const float *pose_hist, const float *sys_hist,
                      const float *aux_hist, const float *grad_pose_hist, const float *grad_A,
                      float *grad_pose_in, void *workspace, size_t workspace_bytes, void *stream);

/*
 * Measurement aid for bench.py: same work as dpft_uic_forward, but every Gauss-Newton launch is bracketed
 * with CUDA events on `stream`, the stream is waited for, and the device time of launch k (milliseconds) is
 * written to the HOST array launch_ms[n_levels*iters].  Events between launches switch their overlap off, so
 * these are clean per-kernel durations, not a way to run the solver.
 */
int dpft_uic_forward_timed(const dpft_level_t *levels, int n_levels, int B, int C, int iters, uint32_t flags,
                           float w_icp, const float *pose_in, float *pose_hist, float *sys_hist,
                           float *aux_hist, int32_t *status, void *workspace, size_t workspace_bytes,
                           void *stream, float *launch_ms);

/*
 * 3D end-point-error loss on the pose pyramid (criterions.py:101-136 compute_RT_EPE_loss, EPE3D_loss :22-46):
 *   loss[b] = sum_n mean over valid pixels of || (R_est[b,n] p + t_est[b,n]) - (R_gt[b] p + t_gt[b]) ||,
 *   p = [(u - cx)/fx, (v - cy)/fy, 1] * depth[b,v,u];  a pixel is valid when the target point has no NaN and
 *   invalid[b,v,u] <= 0 (invalid may be NULL); a sample without valid pixels gives 0.
 * depth, invalid: (B,h,w); K: (B,4) already scaled to (h,w); R_est (B,N,3,3), t_est (B,N,3), N <= 8.
 * The backward writes d loss / d R_est and d t_est scaled by g_loss (B); the target is a constant, as in the reference.
 */
int dpft_pose_epe_loss(const float *depth, const float *invalid, const float *K, const float *R_gt, const float *t_gt,
                       const float *R_est, const float *t_est, int B, int N, int h, int w, float *loss, void *stream);
int dpft_pose_epe_loss_backward(const float *depth, const float *invalid, const float *K, const float *R_gt,
                                const float *t_gt, const float *R_est, const float *t_est, int B, int N, int h, int w,
                                const float *g_loss, float *g_R_est, float *g_t_est, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DPFT_H_ */
