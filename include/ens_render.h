/*
 * ens_render.h -- C ABI of the B200-native fused ray renderer for EvenNICER-SLAM.
 *
 * The reference has no FFI layer: its seam is the Python call signatures of
 * Renderer / NICE / get_samples.  Each entry point below names the
 * reference function whose body it replaces (paths relative to /root/reference).
 * The Python drop-ins in evennicer_slam_b200/ bind these with ctypes; INTEGRATION.md
 * shows the stub a maintainer adds on the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - the caller owns every buffer (inputs, outputs, gradient accumulators, workspace);
 *     the library never allocates, frees or retains a pointer after return;
 *   - all work is enqueued on `stream` (a cudaStream_t) and returns without host sync;
 *   - return value 0 = ENS_OK, negative = error (ens_strerror);  no aborts, no exceptions;
 *   - gradient buffers are ACCUMULATED INTO (caller zero-initialises);
 *   - no global mutable state: safe from several processes / threads on one GPU.
 *
 * Layouts
 *   - "reference" grid layout: float32 [32][Z][Y][X]  (torch [1,32,Z,Y,X], EvenNICER_SLAM.py:241-273)
 *   - "native"    grid layout: float32 [Z][Y][X][32]  (one 128-byte line per voxel) -- what the
 *     render kernels read and scatter into; ens_grid_to_native / ens_grid_from_native convert.
 *   - decoder weights: a packed blob per decoder built by ens_pack_decoder from the decoder's
 *     individual tensors in state_dict order; decoder gradients come back as ONE flat float32
 *     buffer per decoder holding the tensors back-to-back in state_dict order and reference
 *     shapes (so the host can hand out views).
 */
#ifndef ENS_RENDER_H_
#define ENS_RENDER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ENS_ABI_VERSION 6

typedef void *ens_stream_t; /* cudaStream_t */

enum {
  ENS_OK = 0,
  ENS_EINVAL = -1,       /* null pointer / bad enum / negative size */
  ENS_ESHAPE = -2,       /* inconsistent sizes */
  ENS_ECUDA = -3,        /* a CUDA call failed; ens_last_error() has the message */
  ENS_ENCCL = -4,        /* reserved (collectives live in torch.distributed) */
  ENS_EUNSUPPORTED = -5  /* iMAP modes: N_importance>0, occupancy=False, perturb>0, lindisp */
};

/* stage of NICE.forward (src/conv_onet/models/decoder.py:312-342) */
enum { ENS_STAGE_COARSE = 0, ENS_STAGE_MIDDLE = 1, ENS_STAGE_FINE = 2, ENS_STAGE_COLOR = 3 };
/* grid level / decoder index */
enum { ENS_LEVEL_COARSE = 0, ENS_LEVEL_MIDDLE = 1, ENS_LEVEL_FINE = 2, ENS_LEVEL_COLOR = 3 };

#define ENS_C_DIM 32          /* configs/nice_slam.yaml:111 (model.c_dim)  */
#define ENS_HIDDEN 32         /* decoder.py:295 (hidden_size)              */
#define ENS_EMBED 93          /* decoder.py:127 (Fourier mapping size)     */
#define ENS_MAX_SAMPLES 64    /* N_samples + N_surface upper bound per ray */

/* Scene = what `c` (dict of grids) + `decoders` carry in the reference. */
typedef struct EnsScene {
  const float *grid[4];        /* native layout, indexed by ENS_LEVEL_*; NULL if absent       */
  int32_t dims[4][3];          /* (Z,Y,X) per level                                            */
  double bound[3][2];          /* slam.bound: used by middle/fine/color decoders and the mask  */
  double coarse_bound[3][2];   /* bound * coarse_bound_enlarge (EvenNICER_SLAM.py:182)         */
  const float *weights[4];     /* packed decoder blobs from ens_pack_decoder; NULL if absent   */
} EnsScene;

/* rendering knobs (cfg['rendering'], Renderer.py:11-19) */
typedef struct EnsRenderCfg {
  int32_t n_samples;           /* N_samples   (32)                                             */
  int32_t n_surface;           /* N_surface   (16); forced to 0 without gt_depth / stage coarse */
  int32_t n_importance;        /* must be 0                                                     */
  int32_t lindisp;             /* must be 0                                                     */
  float perturb;               /* must be 0                                                     */
  int32_t occupancy;           /* must be 1                                                     */
  const float *t_vals;         /* device float32 [n_samples]: torch.linspace(0,1,N_samples)      */
  const double *t_vals_surface;/* device float64 [n_surface]: torch.linspace(0,1,N_surface).double() */
} EnsRenderCfg;

/* gradient sinks of ens_render_bwd; any member may be NULL (= not needed). */
typedef struct EnsGrads {
  float *grid[4];              /* native layout, accumulate-into                                */
  float *decoder[4];           /* flat per-decoder buffers (state_dict order), accumulate-into   */
  float *rays_o;               /* [R][3], written (not accumulated)                              */
  float *rays_d;               /* [R][3], written                                                */
} EnsGrads;

int ens_version(void);
const char *ens_strerror(int code);
/* message of the last CUDA error a call of this thread returned ENS_ECUDA for ("" if none); the error itself has been
 * cleared from the runtime (cudaGetLastError), so it does not leak into later calls. */
const char *ens_last_error(void);

/* number of floats in a packed decoder blob / in the flat gradient buffer of a decoder */
int64_t ens_packed_decoder_floats(int level);
int64_t ens_decoder_grad_floats(int level);
/* number of tensors in a decoder's state_dict (22 for coarse-less MLPs: 10 fc_c + _B + 10 pts + 2 out = 23; 12 coarse) */
int ens_decoder_num_tensors(int level);
/* bytes of scratch ens_render_bwd needs for R rays x S samples (the tcgen05 backward keeps the sample points, the
 * per-point output gradients, d L / d p and the folded weight-gradient sums there) */
int64_t ens_bwd_workspace_bytes(int64_t n_rays, int n_samples_total, int want_decoder_grads);
/* bytes of the saved-for-backward buffer ens_render_fwd can fill (relu masks, plus the hidden activations when
 * decoder gradients will be wanted) so that ens_render_bwd does not recompute the forward -- what torch autograd
 * keeps as saved tensors in the reference.  0 = not applicable for this stage / sample count (pass NULL). */
int64_t ens_fwd_saved_bytes(int64_t n_rays, int n_samples_total, int stage, int want_decoder_grads);
/* the same by saved kind (the saved_with_activations argument of ens_render_fwd / ens_render_bwd): 0, 1 as above; 2 = relu
 * masks written by the tcgen05 forward; 3 = the relu outputs r_0..r_4 of every decoder written by the tcgen05 forward
 * (640 B per point and decoder) -- what the tcgen05 backward needs for the decoder gradients. */
int64_t ens_fwd_saved_bytes_kind(int64_t n_rays, int n_samples_total, int stage, int kind);
/* bytes of optional scratch for a forward that will NOT be followed by a backward (render_img, visualisation): with it
 * ens_render_fwd runs sample placement, the tcgen05 decode and the compositing as separate kernels over per-point
 * arrays in this scratch (faster for large batches); without it the fused kernel is used.  0 = not applicable. */
int64_t ens_fwd_scratch_bytes(int64_t n_rays, int n_samples_total, int stage);

/* [32][Z][Y][X] -> [Z][Y][X][32]  and back.  n_vox = Z*Y*X.  (layout of `c`, EvenNICER_SLAM.py:217-275) */
int ens_grid_to_native(const float *ref_layout, float *native, int64_t n_vox, ens_stream_t stream);
int ens_grid_from_native(const float *native, float *ref_layout, int64_t n_vox, int accumulate,
                         ens_stream_t stream);

/* Gather a decoder's individual tensors (device pointers, state_dict order, listed in a HOST array)
 * into the packed blob the kernels stage into shared memory.  Replaces nothing in the reference; it
 * is the price of accepting nn.Parameter-per-tensor storage (decoder.py:108-164, 224-250). */
int ens_pack_decoder(int level, const float *const *tensors_host, int n_tensors, float *packed,
                     ens_stream_t stream);

/* max(gt_depth) -> out[0] = (double)(max*1.2f)  [Renderer.py:110],  out[1] = (double)max  [Renderer.py:145].
 * `out` is a device double[2]; pass it to ens_render_fwd/bwd as depth_max. */
int ens_depth_max(const float *gt_depth, int64_t n, double *out, ens_stream_t stream);

/* common.get_samples (src/common.py:92-187) after the torch.randint draw: for each flat index into
 * the crop [H0:H1, W0:W1] produce pixel coords, gathered depth (f32) / colour (f64 or f32) and the ray
 * (get_rays_from_uv, common.py:74-89) for c2w (device float32, row stride `c2w_stride` >= 4, 3 rows). */
int ens_sample_rays(const int64_t *indices, int64_t n, int H0, int H1, int W0, int W1, int H, int W,
                    float fx, float fy, float cx, float cy, const float *c2w, int c2w_stride,
                    const float *depth, const void *color, int color_is_f64,
                    float *pix_i, float *pix_j, float *rays_o, float *rays_d, float *out_depth,
                    void *out_color, ens_stream_t stream);

/* common.get_rays / get_rays_rescale (src/common.py:300-340): rays of an (nH x nW) pixel lattice whose
 * coordinates are lin_w[nW], lin_h[nH] (device float32: the reference's torch.linspace vectors).
 * nH == 0 selects "pairs" mode: nW rays with pixel (lin_w[k], lin_h[k]) = get_rays_from_uv (common.py:74-89). */
int ens_lattice_rays(const float *lin_w, int nW, const float *lin_h, int nH, float fx, float fy, float cx,
                     float cy, const float *c2w, int c2w_stride, float *rays_o, float *rays_d,
                     ens_stream_t stream);

/* backward of both ray generators into c2w: g_c2w (device float32 [3][4], accumulate-into).
 * pix_i/pix_j are the per-ray pixel coordinates (for a lattice: pass lin_w/lin_h with nW>0). */
int ens_rays_bwd(const float *pix_i, const float *pix_j, int64_t n, int nW, float fx, float fy, float cx,
                 float cy, const float *g_rays_o, const float *g_rays_d, float *g_c2w,
                 ens_stream_t stream);

/* common.get_camera_from_tensor / quad2rotation (src/common.py:189-228): n camera tensors
 * [qw,qx,qy,qz,tx,ty,tz] (un-normalised quaternion allowed) -> n x [3][4] float32 [R|t]; and the
 * backward g_c2w [n][3][4] -> g_cam [n][7] (written).  In eager PyTorch this is ~60 tiny kernels per pose. */
int ens_pose_fwd(const float *cam_tensors, int n, float *c2w, ens_stream_t stream);
int ens_pose_bwd(const float *cam_tensors, int n, const float *g_c2w, float *g_cam, ens_stream_t stream);

/* Renderer.eval_points (src/utils/Renderer.py:24-62; clone at src/utils/Mesher.py:281-319):
 * pts [N][3] float64 or float32 -> out [N][4] float32 (r,g,b,occ).  apply_bound_mask != 0 applies the
 * `ret[~mask,3] = 100` rule for points outside slam.bound (Renderer.py:43-58); 0 gives the bare
 * NICE.forward (src/conv_onet/models/decoder.py:312-342) that Mesher.eval_points calls. */
int ens_eval_points(const EnsScene *scene, int stage, const void *pts, int pts_is_f64, int64_t n,
                    int apply_bound_mask, float *out4, ens_stream_t stream);

/* Renderer.render_batch_ray forward (src/utils/Renderer.py:64-199) with raw2outputs_nerf_color
 * (src/common.py:256-297).  gt_depth may be NULL (and is ignored for stage coarse); depth_max is the
 * device double[2] from ens_depth_max (NULL iff gt_depth is NULL).
 * Outputs: depth f64 [R], var f64 [R], color f32 [R][3].  Optional (may be NULL): z_vals f64 [R][S],
 * weights f32 [R][S], raw f32 [R][S][4] (pre-sigmoid; `raw` is also what ens_render_bwd wants back).
 * saved / saved_bytes: optional buffer of ens_fwd_saved_bytes(n_rays, S, stage, saved_with_activations == 1) bytes,
 * filled for ens_render_bwd (16-byte aligned; NULL = the backward recomputes).
 * saved_with_activations: 0 = relu masks only (enough for a backward without decoder gradients), 1 = masks + hidden
 * activations (decoder gradients), 2 = masks only, written by the tcgen05 decode path (placement -> decode -> compositing;
 * needs `scratch`; ENS_EUNSUPPORTED if that path cannot run), 3 = relu outputs of every block, written by the tcgen05
 * decode path (needs `scratch`; the forward of the tcgen05 backward with decoder gradients) -- pass the same value to
 * ens_render_bwd.
 * scratch / scratch_bytes: optional, see ens_fwd_scratch_bytes (used when saved is NULL or saved_with_activations == 2). */
int ens_render_fwd(const EnsScene *scene, const EnsRenderCfg *cfg, int stage, const float *rays_o,
                   const float *rays_d, const float *gt_depth, const double *depth_max, int64_t n_rays,
                   double *depth, double *var, float *color, double *z_vals, float *weights, float *raw,
                   void *saved, int64_t saved_bytes, int saved_with_activations, void *scratch,
                   int64_t scratch_bytes, ens_stream_t stream);

/* Backward of ens_render_fwd into grid features, decoder weights and rays.
 * Recomputes sample placement and activations; `raw` is the [R][S][4] tensor saved by the forward.
 * g_depth f64 [R], g_var f64 [R], g_color f32 [R][3]; any may be NULL (= zero).
 * workspace: ens_bwd_workspace_bytes bytes (only touched when decoder grads are requested and nothing was saved).
 * saved / saved_bytes / saved_with_activations: the buffer ens_render_fwd filled for the same rays (or NULL). */
int ens_render_bwd(const EnsScene *scene, const EnsRenderCfg *cfg, int stage, const float *rays_o,
                   const float *rays_d, const float *gt_depth, const double *depth_max, int64_t n_rays,
                   const float *raw, const double *g_depth, const double *g_var, const float *g_color,
                   const EnsGrads *grads, void *workspace, int64_t workspace_bytes, const void *saved,
                   int64_t saved_bytes, int saved_with_activations, ens_stream_t stream);

/* ---- fused frustum-masked Adam step on the feature grids ---------------------------------
 * One launch replaces, for every grid level of a mapping iteration (src/Mapper.py): the boolean-mask gather of the
 * optimisable features (:343-361), the index_put of them into the grid before each render (:451-458),
 * torch.optim.Adam.step on the grid parameter groups (:396-423, :625; default betas/eps, no weight decay, no amsgrad)
 * and the write-back (:633-641).  All buffers are in the NATIVE layout [Z][Y][X][32]; `grad` is the dense gradient
 * ens_render_bwd accumulated.  The voxels listed in voxel_index (the nonzero positions of the frustum mask
 * `torch.from_numpy(mask).permute(2,1,0)` flattened over [Z][Y][X], ascending, built once per optimize_map call) are
 * updated, all 32 channels (Mapper.py:345-346); other voxels and their moments are left untouched.  voxel_index == NULL
 * selects every voxel (n_selected must equal n_voxels).  exp_avg / exp_avg_sq are dense, caller-owned, zero-initialised
 * when the optimiser is created (one per optimize_map call in the reference).
 * step: 1-based Adam step of this call.  dyn: NULL, or a DEVICE double[1 + n_levels] = {step, lr[0..n_levels)} read by the
 * kernel instead of `step` / `lr` -- lets a captured CUDA graph advance the step and change the stage's learning rates.
 * clear_grad: also zero `grad` of the voxels it updates. */
typedef struct EnsAdamLevel {
  float *grid;               /* native layout, updated in place                          */
  float *grad;               /* native layout; const unless clear_grad                   */
  float *exp_avg;            /* first moment, native layout                              */
  float *exp_avg_sq;         /* second moment, native layout                             */
  const int32_t *voxel_index;/* device int32 [n_selected], ascending, or NULL = every voxel */
  int64_t n_selected;        /* number of selected voxels                                */
  int64_t n_voxels;          /* Z*Y*X (< 2^31)                                           */
  double lr;                 /* cfg['mapping']['stage'][stage]['<level>_lr'] * lr_factor */
} EnsAdamLevel;
int ens_grid_adam_step(const EnsAdamLevel *levels_host, int n_levels, double beta1, double beta2, double eps,
                       int64_t step, const double *dyn, int clear_grad, ens_stream_t stream);

/* ---- sharded mapping step: exchange only the voxels some rank touched ---------------------------------------
 * grad: a native-layout gradient [n_vox][32] as ens_render_bwd accumulates it.
 * ens_grid_touched: flags[v] = 1 iff voxel v has a non-zero channel (device int32 [n_vox], overwritten).
 * ens_grid_compact: with flags MAX-reduced over the ranks and pos = their inclusive prefix sum (device int32 [n_vox], over all
 * levels that share `compact`), to_compact != 0 copies the flagged rows to compact[pos[v] - 1] ([capacity][32] floats);
 * to_compact == 0 copies them back (after the SUM all-reduce of `compact`).  Rows past `capacity` are left alone and counted
 * in *overflow (device int32, accumulated): a caller that sees it non-zero must fall back to the dense all-reduce. */
int ens_grid_touched(const float *grad, int64_t n_vox, int32_t *flags, ens_stream_t stream);
int ens_grid_compact(float *grad, const int32_t *flags, const int32_t *pos, int64_t n_vox, float *compact, int64_t capacity,
                     int to_compact, int32_t *overflow, ens_stream_t stream);

/* Adam over many small tensors in ONE launch: the other parameter groups of the mapper's optimizer (decoder weights,
 * camera tensors; src/Mapper.py:396-423, :625) and the tracker's camera tensor (src/Tracker.py:335-342).  Same
 * arithmetic as ens_grid_adam_step.  params_host / grads_host / sizes_host / groups_host: host arrays of n_tensors device
 * pointers, element counts and group indices; lrs_host[n_groups] (n_groups <= 8); exp_avg / exp_avg_sq: device float
 * buffers of sum(sizes) elements, tensor i at the sum of the sizes before it, zero-initialised by the caller.
 * dyn: NULL or device double[1 + n_groups] = {step, lr per group} read at run time (CUDA-graph replays). */
int ens_tensors_adam_step(float *const *params_host, const float *const *grads_host, const int64_t *sizes_host,
                          const int *groups_host, int n_tensors, const double *lrs_host, int n_groups, float *exp_avg,
                          float *exp_avg_sq, double beta1, double beta2, double eps, int64_t step, const double *dyn,
                          ens_stream_t stream);

/* ---- blurred-L2 event loss, value + gradient in one launch ------------------------------
 * Replaces src/Tracker.py:204-224 and src/Mapper.py:593-615:
 *     loss = ((gt - pred)**2).sum() + sum_k w_k * ((gaussian_blur(gt, ks_k) - gaussian_blur(pred, ks_k))**2).sum()
 *     loss *= balancer;  loss.backward()
 * gaussian_blur = torchvision.transforms.functional.gaussian_blur on a tensor: reflect padding by ks/2 and a depthwise
 * conv2d with the outer product of the normalised 1-d kernel (sigma = 0.15 ks + 0.35 by default).
 * gt, pred: device float32 [H][W][C] (the layout of gt_event / full_event in the reference).
 * kernel_sizes_host[n_kernels] (odd, <= 15), kernels1d_host: the 1-d kernels back to back (sum ks_k floats, as
 * torchvision's _get_gaussian_kernel1d computes them in float32), kernel_weights_host[n_kernels]; n_kernels <= 4.
 * loss_parts: device double[2 + n_kernels] = {balancer * total, sum (gt-pred)^2, blurred sum per kernel} (overwritten).
 * g_pred: device float32 [H][W][C] = d(loss_parts[0]) / d pred (overwritten), or NULL. */
int ens_event_loss(const float *gt, const float *pred, int H, int W, int C, const int *kernel_sizes_host,
                   const float *kernels1d_host, const float *kernel_weights_host, int n_kernels, float balancer,
                   double *loss_parts, float *g_pred, ens_stream_t stream);

/* ---- frustum feature selection and keyframe overlap of the mapper --------------------------------
 * ens_frustum_mask replaces Mapper.get_mask_from_c2w (src/Mapper.py:115-186, numpy + cv2.remap on the host) for the
 * middle / fine / colour grids: voxel centre (xs[ix], ys[iy], zs[iz]) is selected when it projects inside the image with
 * 0 <= depth-along-axis <= bilinear(depth image) + 0.5 (zero depths read as the maximum over all voxels), or lies within
 * 0.5 of the camera centre.  w2c_host = HOST float[12..16], rows 0..2 of inv(c2w) (the reference inverts on the host too);
 * cam_centre_host = HOST float[3] = c2w[:3,3]; cam6_host = HOST double[6] = {H, W, fx, fy, cx, cy}; xs/ys/zs = DEVICE float
 * coordinate axes (torch.linspace of the bound, :132-134); depth = DEVICE float32 [H][W]; mask = DEVICE uint8, [NZ][NY][NX]
 * when mask_zyx != 0 (grid layout) else [NX][NY][NZ] (the reference's return value).  workspace: ens_frustum_workspace_bytes.
 * ens_keyframe_overlap replaces the projection loop of keyframe_selection_overlap (:222-241): counts[k] = number of
 * vertices [n][3] (DEVICE) that project more than `edge` pixels inside keyframe k's image, w2cs = DEVICE float [K][16]. */
int64_t ens_frustum_workspace_bytes(int NX, int NY, int NZ);
int ens_frustum_mask(const float *w2c_host, const float *cam_centre_host, const double *cam6_host, const float *xs,
                     const float *ys, const float *zs, int NX, int NY, int NZ, const float *depth, int H, int W,
                     int mask_zyx, uint8_t *mask, void *workspace, int64_t workspace_bytes, ens_stream_t stream);
int ens_keyframe_overlap(const float *w2cs, int n_keyframes, const double *cam6_host, float edge, const float *vertices,
                         int n_vertices, int *counts, ens_stream_t stream);

/* UNet input assembly of the event branch (the head of src/event_net.py:67-99, inference_event): img1 (previous GT colour,
 * HWC float64 or float32, H1 x W1) and img2 (rendered colour, HWC float32, H2 x W2) -> out [6][h][w] float32 =
 * cat(permute(img1), permute(img2)), each nearest-resized to h x w first when its size differs (torchvision Resize(NEAREST) =
 * ATen upsample_nearest2d indexing).  ens_unet_input_bwd: g_out [6][h][w] -> g_img2 [H2][W2][3] (written). */
int ens_unet_input(const void *img1, int img1_is_f64, int H1, int W1, const float *img2, int H2, int W2, int h, int w,
                   float *out, ens_stream_t stream);
int ens_unet_input_bwd(const float *g_out, int H2, int W2, int h, int w, float *g_img2, ens_stream_t stream);

/* ---- RGB-D loss glue of the callers, value + gradients in one launch ----------------------
 * tracker = 0 (src/Mapper.py:553-562):  loss = sum_{gt_depth>0} |gt_depth - depth|  (+ w_color * sum |gt_color - color|
 *                                       when use_color: the mapper's colour stage; the colour term is NOT masked)
 * tracker = 1 (src/Tracker.py:180-196): tmp = |gt_depth - depth| / sqrt(uncertainty + 1e-10)   (uncertainty detached)
 *                                       mask = (gt_depth > 0) [& tmp < 10 * median(tmp) when handle_dynamic]
 *                                       loss = sum_mask tmp (+ w_color * sum_mask |gt_color - color| when use_color)
 * median = torch.median: the lower middle element, found exactly.  All arithmetic float64, as in the reference.
 * depth, uncertainty: device double[n]; color: float[n][3]; gt_depth: float[n]; gt_color: float or double [n][3].
 * loss: device double[1] (overwritten); g_depth: double[n] = d loss / d depth, g_color: float[n][3] = d loss / d color
 * (overwritten; either may be NULL); workspace: device double[n], tracker only. */
int ens_rgbd_loss(int tracker, const double *depth, const double *uncertainty, const float *color, const float *gt_depth,
                  const void *gt_color, int gt_color_is_f64, int64_t n, int use_color, double w_color,
                  int handle_dynamic, double *loss, double *g_depth, float *g_color, double *workspace,
                  ens_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ENS_RENDER_H_ */
