/*
 * d4gs.h -- C ABI of libd4gs.so: the B200 (sm_100a) implementation of the
 * Deblur4DGS per-frame render hot path.
 *
 * The reference has no native ABI of its own for this path: it calls the
 * torch-extension API of gsplat==1.1.1 from Python
 * (flow3d/scene_model.py:5, 360-373) and plain PyTorch ops for the
 * deformation (flow3d/params.py:142-180, flow3d/scene_model.py:76-120).  Each
 * entry point below names the reference interface it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 / int32 / int64 data laid out
 *     row-major and contiguous; the library never allocates, frees or keeps
 *     device memory; workspaces are passed in by the caller (PyTorch tensors);
 *   - `stream` is a cudaStream_t (CUstream) -- work is only enqueued, never
 *     synchronised; functions are re-entrant and keep no global mutable state;
 *   - return 0 on success, non-zero on error; d4_last_error() returns a
 *     thread-local message; nothing throws across the boundary;
 *   - "*_cam_stride" = elements between consecutive cameras of a per-camera
 *     array, or 0 when all cameras share one copy.  A sub-exposure batch of N
 *     renders is expressed as C = N cameras with per-camera means / quats
 *     (stride G*3 / G*4) and shared viewmat / K (stride 0);
 *   - gradient outputs ("v_*") are ACCUMULATED into (atomicAdd) or overwritten
 *     as documented per function; unless stated the caller zero-fills them.
 */
#ifndef D4GS_H_
#define D4GS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *d4_stream_t; /* cudaStream_t */

#define D4_ABI_VERSION 2

int d4_version(void);
const char *d4_last_error(void);

/* ---- a8: projection + 2-D covariance ----------------------------------------
 * replaces gsplat.fully_fused_projection (fwd) as reached from
 * gsplat.rendering.rasterization, scene_model.py:360-373; fused with the first
 * pass of gsplat.isect_tiles (tiles_per_gauss).
 * means [*,G,3], quats [*,G,4] wxyz (normalised inside), scales [G,3],
 * viewmats [*,4,4], Ks [*,3,3].
 * out: radii i32 [C,G] (0 = culled), means2d [C,G,2], depths [C,G],
 *      conics [C,G,3], tiles_per_gauss i32 [C,G] (may be NULL).
 * cam_row0 (i32 [C], may be NULL) / window_height: per-camera ROW WINDOW for the multi-GPU tile-row bands
 * (SURVEY 8e): camera c renders rows [cam_row0[c], cam_row0[c] + window_height) of the width x height image.
 * The projection is that of the full image; means2d.y is returned relative to the window, Gaussians that
 * cannot touch it get radii = 0, and tile_h must be the window's.  Binning and blend then run on
 * C images of width x window_height.                                                                    */
int d4_project_fwd(const float *means, int64_t means_cam_stride, const float *quats,
                   int64_t quats_cam_stride, const float *scales, const float *viewmats,
                   int64_t viewmat_cam_stride, const float *Ks, int64_t k_cam_stride, int C, int G,
                   int width, int height, float eps2d, float near_plane, float far_plane,
                   float radius_clip, int tile_size, int tile_w, int tile_h, int32_t *radii,
                   float *means2d, float *depths, float *conics, int32_t *tiles_per_gauss,
                   const int32_t *cam_row0, int window_height, d4_stream_t stream);

/* ---- a12: projection backward -------------------------------------------------
 * replaces gsplat.fully_fused_projection backward.  v_means [*,G,3] and
 * v_quats [*,G,4] use the input strides; v_scales [G,3]; v_viewmats [C,4,4]
 * may be NULL.  All four are accumulated into: caller zero-fills.            */
int d4_project_bwd(const float *means, int64_t means_cam_stride, const float *quats,
                   int64_t quats_cam_stride, const float *scales, const float *viewmats,
                   int64_t viewmat_cam_stride, const float *Ks, int64_t k_cam_stride, int C, int G,
                   int width, int height, float eps2d, const int32_t *radii, const float *conics,
                   const float *v_means2d, const float *v_depths, const float *v_conics,
                   float *v_means, float *v_quats, float *v_scales, float *v_viewmats,
                   d4_stream_t stream);

/* ---- a9: tile binning -----------------------------------------------------------
 * replaces gsplat.isect_tiles (cumsum + second pass), the CUB radix sort inside
 * it, and gsplat.isect_offset_encode.                                            */

/* exclusive prefix sum of i32 counts; total (i64 device scalar) = sum of all */
size_t d4_scan_workspace_bytes(int64_t n);
int d4_exclusive_scan_i32(const int32_t *in, int64_t n, int32_t *out_exclusive, int64_t *total,
                          void *workspace, size_t workspace_bytes, d4_stream_t stream);

/* emit isect_ids i64 = (cam << (32+tile_n_bits)) | (tile << 32) | bits(depth)
 * and flatten_ids i32 = cam*G + g, in (cam, g, tile row-major) order          */
int d4_tile_n_bits(int n_tiles);
int d4_isect_emit(const float *means2d, const int32_t *radii, const float *depths,
                  const int32_t *cum_tiles_exclusive, int C, int G, int tile_size, int tile_w,
                  int tile_h, int64_t *isect_ids, int32_t *flatten_ids, d4_stream_t stream);

/* stable LSD radix sort of (u64 key, u32 value) pairs on key bits
 * [begin_bit, end_bit); ping-pongs between the a and b buffers; *result_in_b
 * (host int) tells where the sorted data ended up.                              */
size_t d4_sort_workspace_bytes(int64_t n);
int d4_sort_pairs_u64(uint64_t *keys_a, uint32_t *vals_a, uint64_t *keys_b, uint32_t *vals_b,
                      int64_t n, int begin_bit, int end_bit, void *workspace,
                      size_t workspace_bytes, int *result_in_b, d4_stream_t stream);

/* Tile-bucketed alternative to d4_isect_emit + d4_sort_pairs_u64 + d4_tile_offsets (same results,
 * ~8x less HBM traffic): count per (camera, tile) -> exclusive scan (== isect_offsets) ->
 * emit into the tile's segment (cursors zero-filled by the caller) -> per-tile shared-memory sort by
 * (depth bits, flatten id).  Usable when no tile holds more than d4_tile_sort_capacity() entries.  */
int d4_tile_sort_capacity(void);
int d4_tile_count(const float *means2d, const int32_t *radii, int C, int G, int tile_size, int tile_w,
                  int tile_h, int32_t *tile_counts, d4_stream_t stream);
int d4_bucket_emit(const float *means2d, const int32_t *radii, const float *depths, int C, int G,
                   int tile_size, int tile_w, int tile_h, const int32_t *tile_offsets, int32_t *cursors,
                   uint64_t *bucket_keys, int64_t capacity /* entries of bucket_keys; writes beyond are dropped */,
                   int bucket_stride /* > 0: fixed-stride buckets, tile t at bucket_keys[t * bucket_stride]; tile_offsets
                                        may be NULL and `cursors` ends up holding the per-tile counts */,
                   d4_stream_t stream);
int d4_tile_sort(const uint64_t *bucket_keys, const int32_t *tile_offsets, int64_t n_isects, int C,
                 int tile_w, int tile_h, int max_count, int64_t *isect_ids, int32_t *flatten_ids,
                 d4_stream_t stream);

/* offsets i32 [C,tile_h,tile_w]: first sorted index of each (camera, tile) */
int d4_tile_offsets(const int64_t *isect_ids_sorted, int64_t n_isects, int C, int tile_w,
                    int tile_h, int32_t *offsets, d4_stream_t stream);

/* ---- a10: blend forward -----------------------------------------------------------
 * replaces gsplat.rasterize_to_pixels (fwd) plus the Python around it in
 * gsplat.rendering.rasterization: depth appended as an extra colour channel for
 * "RGB+ED"/"RGB+D" (depths != NULL, background of that channel = 0) and the
 * expected-depth normalisation depth / max(alpha, 1e-10) (normalize_depth != 0).
 * means2d [C,G,2], conics [C,G,3], opacities [G], colors [*,G,D0],
 * backgrounds [C,D0] or NULL.  D = D0 + (depths ? 1 : 0), 1 <= D <= 64.
 * out: render_colors [C,H,W,D], render_alphas [C,H,W], last_ids i32 [C,H,W],
 *      acc_depth [C,H,W] (un-normalised depth, only when normalize_depth),
 *      hit_masks u8 [n_isects] or NULL (caller zero-fills): bit w of entry i set iff some pixel of
 *      the w-th 8x4 pixel block of the tile passed the alpha test for intersection i; hand it to
 *      d4_blend_bwd, which then visits only those (block, Gaussian) pairs.          */
int d4_blend_fwd(const float *means2d, const float *conics, const float *opacities,
                 const float *colors, int64_t colors_cam_stride, const float *depths,
                 const float *backgrounds, int C, int G, int D0, int width, int height,
                 int tile_size, int tile_w, int tile_h, const int32_t *tile_offsets,
                 const int32_t *flatten_ids, int64_t n_isects, int normalize_depth,
                 float *render_colors, float *render_alphas, int32_t *last_ids, float *acc_depth,
                 uint8_t *hit_masks, d4_stream_t stream);

/* ---- a11: blend backward ------------------------------------------------------------
 * replaces gsplat.rasterize_to_pixels backward (and autograd of the depth
 * normalisation).  v_means2d [C,G,2], v_conics [C,G,3], v_colors [*,G,D0]
 * (colors_cam_stride as forward), v_opacities [G], v_depths [C,G] (iff depths):
 * accumulated with atomics, caller zero-fills.                                   */
int d4_blend_bwd(const float *means2d, const float *conics, const float *opacities,
                 const float *colors, int64_t colors_cam_stride, const float *depths,
                 const float *backgrounds, int C, int G, int D0, int width, int height,
                 int tile_size, int tile_w, int tile_h, const int32_t *tile_offsets,
                 const int32_t *flatten_ids, int64_t n_isects, int normalize_depth,
                 const float *render_alphas, const int32_t *last_ids, const float *acc_depth,
                 const float *v_render_colors, const float *v_render_alphas, float *v_means2d,
                 float *v_conics, float *v_colors, float *v_opacities, float *v_depths,
                 const uint8_t *hit_masks /* from d4_blend_fwd, or NULL */,
                 int bwd_mode /* 0 = grouped kernel (default), 1 = warp-shuffle kernel */, d4_stream_t stream);

/* ---- a9 -> a10/a11: packed per-tile record slabs ("slab" path, the default of rasterization()) ----------
 * The blend kernels' own input format, built once per forward from what gsplat.isect_tiles /
 * isect_offset_encode produce: per intersection one 32-byte record
 *     (x, y, log2(opacity), local id | reach mask << 24), (A', B', C', depth)
 * with the conic pre-scaled to base 2, in the tile's depth order; records whose Gaussian cannot reach
 * alpha >= 1/255 on any pixel of the tile are dropped, so tile t owns records
 * [tile_offsets[t], tile_offsets[t] + rec_counts[t]).  recs: 32 * n_isects bytes, 16-byte aligned;
 * rec_counts i32 [C*tile_h*tile_w].  G < 2^24.  depths may be NULL (depth field = 0).
 * d4_tile_sort_pack = d4_tile_sort + d4_isect_pack in one kernel (ids still in shared memory).          */
int d4_isect_pack(const float *means2d, const float *conics, const float *opacities, const float *depths,
                  int C, int G, int tile_size, int tile_w, int tile_h, const int32_t *tile_offsets,
                  const int32_t *flatten_ids, int64_t n_isects, void *recs, int32_t *rec_counts,
                  d4_stream_t stream);
int d4_tile_sort_pack(const uint64_t *bucket_keys, const int32_t *tile_offsets, int64_t n_isects, int C,
                      int tile_w, int tile_h, int max_count, int64_t *isect_ids, int32_t *flatten_ids,
                      const float *means2d, const float *conics, const float *opacities,
                      const float *depths, int G, int tile_size, void *recs, int32_t *rec_counts,
                      d4_stream_t stream);
/* Capacity mode: the binning without its device -> host read-back (gsplat.isect_tiles syncs to size its outputs;
 * this variant is graph-capturable).  d4_scan_counts = d4_exclusive_scan_i32 that also leaves the largest count:
 * stats[0] = total intersections, stats[1] = largest per-tile count (device int64).  d4_tile_sort_pack_cap reads
 * the count from bin_stats[0] on the device; all per-intersection buffers hold `capacity` entries, the per-tile
 * shared-memory sort `sort_capacity` keys (<= d4_tile_sort_capacity_max()).  A tile that does not fit is left
 * without records and *overflow (device int64, caller zero-fills) is set to 1: the caller reads it whenever it
 * next synchronises and re-renders with a larger capacity.  With fixed-stride buckets (bucket_stride > 0) the
 * count pass disappears: d4_bucket_emit counts while it emits, d4_scan_counts turns the counts into offsets, and
 * the sort reads tile t's keys from its bucket and writes the sorted lists compactly at tile_offsets[t].          */
int d4_tile_sort_capacity_max(void);
int d4_scan_counts(const int32_t *counts, int64_t n, int32_t *offsets, int64_t *stats, void *workspace,
                   size_t workspace_bytes, d4_stream_t stream);
int d4_tile_sort_pack_cap(const uint64_t *bucket_keys, const int32_t *tile_offsets, const int64_t *bin_stats,
                          int64_t capacity, int sort_capacity, int C, int tile_w, int tile_h,
                          int64_t *isect_ids, int32_t *flatten_ids, const float *means2d, const float *conics,
                          const float *opacities, const float *depths, int G, int tile_size, void *recs,
                          int32_t *rec_counts, int64_t *overflow,
                          int bucket_stride /* as d4_bucket_emit; 0 = compact buckets at tile_offsets */,
                          d4_stream_t stream);
/* u32 words of the hit_bits buffer below: ((n_isects >> 5) + n_segments + 1) * 8 */
size_t d4_slab_hit_words(int64_t n_isects, int64_t n_segments);

/* a10 / a11 over the record slabs: same results as d4_blend_fwd / d4_blend_bwd (same per-pair arithmetic).
 * One CTA per (camera, tile) = 8 consumer warps + 1 producer warp that streams the tile's records
 * (cp.async.bulk) and colour rows (cp.async) through an mbarrier-synchronised shared-memory ring.
 * colors [*,G,D0] with D0 in {4, 8, 16, 32}, 16-byte aligned; with_depth blends the records' depth field
 * as channel D0.  last_ids index the RECORD list.  hit_bits (u32 [d4_slab_hit_words], may be NULL for a
 * forward without backward): per (32-record chunk, 8x4 pixel block) which records passed the alpha test;
 * the backward visits exactly those.  Gradients are accumulated with atomics: caller zero-fills.          */
int d4_blend_fwd_slab(const void *recs, const int32_t *tile_offsets, const int32_t *rec_counts,
                      const float *colors, int64_t colors_cam_stride, const float *backgrounds, int C, int G,
                      int D0, int with_depth, int width, int height, int tile_size, int tile_w, int tile_h,
                      int normalize_depth, float *render_colors, float *render_alphas, int32_t *last_ids,
                      float *acc_depth, uint32_t *hit_bits, d4_stream_t stream);
/* d4_blend_fwd_slab with the formulation chosen explicitly (same decisions -- alphas, last_ids, hit words -- bit for
 * bit; colours within fp32 rounding):
 *   variant 0  every accumulation on the fp32 pipe (packed FFMA2), a record is composited when it is met;
 *   variant 1  records are queued 16 at a time per 8x4 pixel block and their colours accumulated as
 *              O[32 x 16] += W^T . C on mma.sync.m16n8k8 (3xTF32 split); serves D0 == 16, else runs variant 0.
 * d4_blend_fwd_slab runs the library's default variant (d4_blend_fwd_slab_default_variant()).                     */
int d4_blend_fwd_slab_default_variant(void);
int d4_blend_fwd_slab_variant(const void *recs, const int32_t *tile_offsets, const int32_t *rec_counts,
                              const float *colors, int64_t colors_cam_stride, const float *backgrounds, int C, int G,
                              int D0, int with_depth, int width, int height, int tile_size, int tile_w, int tile_h,
                              int normalize_depth, float *render_colors, float *render_alphas, int32_t *last_ids,
                              float *acc_depth, uint32_t *hit_bits, int variant, d4_stream_t stream);
int d4_blend_bwd_slab(const void *recs, const int32_t *tile_offsets, const int32_t *rec_counts,
                      const float *colors, int64_t colors_cam_stride, const float *backgrounds, int C, int G,
                      int D0, int with_depth, int width, int height, int tile_size, int tile_w, int tile_h,
                      int normalize_depth, const float *render_alphas, const int32_t *last_ids,
                      const float *acc_depth, const float *v_render_colors, const float *v_render_alphas,
                      const uint32_t *hit_bits, float *v_means2d, float *v_conics, float *v_colors,
                      float *v_opacities, float *v_depths, d4_stream_t stream);
/* d4_blend_bwd_slab with the formulation chosen explicitly (same results within fp32 rounding):
 *   variant 0  every sum on the fp32 pipe (packed FFMA2);
 *   variant 1  the per-Gaussian gradient sums over a block's 32 pixels on mma.sync.m16n8k8 (3xTF32 split, fp32-grade);
 *   variant 2  also the <colour, cotangent> products of the recurrence.
 * Variants 1 / 2 serve D0 == 16 with 8-byte aligned v_colors / v_means2d; any other call runs variant 0.
 * d4_blend_bwd_slab runs the library's default variant (d4_blend_bwd_slab_default_variant()).                    */
int d4_blend_bwd_slab_default_variant(void);
int d4_blend_bwd_slab_variant(const void *recs, const int32_t *tile_offsets, const int32_t *rec_counts,
                              const float *colors, int64_t colors_cam_stride, const float *backgrounds, int C, int G,
                              int D0, int with_depth, int width, int height, int tile_size, int tile_w, int tile_h,
                              int normalize_depth, const float *render_alphas, const int32_t *last_ids,
                              const float *acc_depth, const float *v_render_colors, const float *v_render_alphas,
                              const uint32_t *hit_bits, float *v_means2d, float *v_conics, float *v_colors,
                              float *v_opacities, float *v_depths, int variant, d4_stream_t stream);

/* ---- a1-a6: motion-basis deformation at N sub-exposure timestamps ----------------------
 * replaces, fused: GaussianParams activations normalize(quats) / softmax(coefs)
 * (params.py:39-43), MotionBases.compute_transforms (params.py:142-180),
 * cont_6d_to_rmat (transforms.py:41-53), SceneModel.compute_poses_fg/all
 * (scene_model.py:76-120) and the camera sub-exposure transform
 * (scene_model.py:352-353) for all N iterations of the loop at
 * scene_model.py:323.
 * fg_means [Gf,3], fg_quats [Gf,4] raw wxyz, motion_coefs [Gf,K] raw logits,
 * bg_means [Gb,3], bg_quats [Gb,4] raw, rots [K,T,6], transls [K,T,3],
 * times [N], RTs [N,3,4] (NULL = identity).  K <= 64.
 * out: means [N,G,3], quats [N,G,4] (wxyz, unit), G = Gf + Gb, fg first.       */
int d4_deform_fwd(const float *fg_means, const float *fg_quats, const float *motion_coefs,
                  const float *bg_means, const float *bg_quats, const float *rots,
                  const float *transls, const float *times, const float *RTs, int Gf, int Gb, int K,
                  int T, int N, float *out_means, float *out_quats, d4_stream_t stream);

/* backward of d4_deform_fwd.  v_fg_means / v_fg_quats / v_motion_coefs /
 * v_bg_means / v_bg_quats are overwritten; v_rots [K,T,6], v_transls [K,T,3],
 * v_times [N], v_RTs [N,3,4] (may be NULL) are accumulated: caller zero-fills. */
int d4_deform_bwd(const float *fg_means, const float *fg_quats, const float *motion_coefs,
                  const float *bg_means, const float *bg_quats, const float *rots,
                  const float *transls, const float *times, const float *RTs, int Gf, int Gb, int K,
                  int T, int N, const float *v_out_means, const float *v_out_quats,
                  float *v_fg_means, float *v_fg_quats, float *v_motion_coefs, float *v_bg_means,
                  float *v_bg_quats, float *v_rots, float *v_transls, float *v_times, float *v_RTs,
                  d4_stream_t stream);

/* ---- a2/a3 stand-alone: MotionBases.compute_transforms (params.py:142-180) ----------------------
 * for the callers outside the render loop (trainer.py:478,485,701; init_utils.py:322).
 * coefs [G,K] are ALREADY softmaxed (as the reference passes them), ts [B] float frame
 * coordinates; out [G,B,3,4] = [R | t].  Backward: v_coefs [G,K] overwritten; v_rots [K,T,6],
 * v_transls [K,T,3], v_ts [B] accumulated (caller zero-fills).                                    */
int d4_compute_transforms_fwd(const float *coefs, const float *rots, const float *transls, const float *ts,
                              int G, int K, int T, int B, float *out, d4_stream_t stream);
int d4_compute_transforms_bwd(const float *coefs, const float *rots, const float *transls, const float *ts,
                              int G, int K, int T, int B, const float *v_out, float *v_coefs, float *v_rots,
                              float *v_transls, float *v_ts, d4_stream_t stream);

/* ---- a7: camera sub-exposure pose interpolation ---------------------------------------------
 * replaces MoveModel.forward_start_end_mid's pose part (move_model.py:143-147: pp.se3().Exp(),
 * _interpolate -> spline_utils.linear_interpolation :371-408, .Log(), se3_to_SE3 :204-215).
 * start6 / end6: the two se(3) 6-vectors [rho, phi] emitted by the MoveModel heads (device
 * pointers); u_i = linspace(0,1,N)[i].  RTs [N,3,4].  Backward accumulates into v_start6 /
 * v_end6 [6] (caller zero-fills).                                                              */
int d4_camera_interp_fwd(const float *start6, const float *end6, int N, float *RTs, d4_stream_t stream);
int d4_camera_interp_bwd(const float *start6, const float *end6, int N, const float *v_RTs, float *v_start6,
                         float *v_end6, d4_stream_t stream);

/* ---- a13: N-way combine of the sub-exposure renders -------------------------------------
 * replaces the stack/mean/max/min at scene_model.py:386-397: out = mean over N
 * of every channel, except channel max_ch (if >= 0) = max over N and channel
 * min_ch (if >= 0) = min over N; out_alpha = mean of alphas.
 * imgs [N,P,D], alphas [N,P] -> out_img [P,D], out_alpha [P].
 * ref_quirk != 0 reproduces the reference literally: there the last render's
 * tensor is overwritten in place by the average BEFORE the max/min are taken
 * (scene_model.py:391-393), so the extrema run over {r_0..r_{N-2}, mean}
 * instead of {r_0..r_{N-1}}.
 * arg_max / arg_min [P] (uint8, may be NULL when the channel is absent): the forward records per pixel which
 * sub-exposure won the max / min channel (255 = the mean itself, ref_quirk only); the backward routes from
 * them instead of re-reading the N images, so the stack need not be kept for backward.
 * Backward: v_imgs [N,P,D], v_alphas [N,P] overwritten (max/min route the
 * gradient to the first arg-extremum, as torch.max/min(dim) do).  N < 255.      */
int d4_combine_fwd(const float *imgs, const float *alphas, int N, int64_t P, int D, int max_ch,
                   int min_ch, int ref_quirk, float *out_img, float *out_alpha, uint8_t *arg_max,
                   uint8_t *arg_min, d4_stream_t stream);
int d4_combine_bwd(const uint8_t *arg_max, const uint8_t *arg_min, int N, int64_t P, int D, int max_ch,
                   int min_ch, const float *v_out_img, const float *v_out_alpha, float *v_imgs,
                   float *v_alphas, d4_stream_t stream);

/* ---- a13 for the 2-D multi-GPU partition: the same combine over (sub-exposure, tile-row band) units
 * spread across ranks.  A rank holds U units imgs [U,bh,W,D] / alphas [U,bh,W] with host-side coordinates subs /
 * bands (HOST int arrays, U <= 32).  d4_band_partial -> part [n_bands*bh, W, D+1] (sum of img / n_sub, alpha plane
 * as channel D) and ext [n_bands*bh, W, 2] (max of the max_ch value and of minus the min_ch value over the units with
 * sub < n_ext; -inf where the rank has none); the caller all-reduces part (SUM) and ext (MAX).  d4_band_winner ->
 * winner i32 [.., 2]: lowest sub-exposure among the rank's units attaining the global extremum (2^30 if none); the
 * caller all-reduces it (MIN).  d4_band_finalize -> out_img [H,W,D], out_alpha [H,W]; ref_quirk as d4_combine_fwd
 * (winner := -1 where the mean itself wins).  d4_band_bwd routes v_out / v_alpha to v_imgs / v_alphas (overwritten). */
int d4_band_partial(const float *imgs, const float *alphas, const int32_t *subs, const int32_t *bands, int U,
                    int n_sub, int n_ext, int n_bands, int bh, int W, int D, int max_ch, int min_ch,
                    float *part, float *ext, d4_stream_t stream);
int d4_band_winner(const float *imgs, const int32_t *subs, const int32_t *bands, int U, int n_ext, int n_bands,
                   int bh, int W, int D, int max_ch, int min_ch, const float *ext, int32_t *winner,
                   d4_stream_t stream);
int d4_band_finalize(const float *part, const float *ext, int32_t *winner, int H, int W, int D, int max_ch,
                     int min_ch, int ref_quirk, int n_ext, float *out_img, float *out_alpha, d4_stream_t stream);
int d4_band_bwd(const int32_t *winner, const int32_t *subs, const int32_t *bands, int U, int n_sub, int n_bands,
                int bh, int W, int D, int H, int max_ch, int min_ch, const float *v_out, const float *v_alpha,
                float *v_imgs, float *v_alphas, d4_stream_t stream);

/* ---- f1 ("next" row): activations + fg|bg concatenation + feature-vector assembly ------------------
 * replaces GaussianParams' activations (params.py:39-43, 70-84: exp / sigmoid / sigmoid), the fg|bg
 * torch.cat of SceneModel.get_*_all (scene_model.py:122-143) and the colors_override assembly
 * [rgb | fg mask | extra track channels] (scene_model.py:205-289) with one pass per render call.
 * raw fg_* [Gf,.], bg_* [Gb,.], extra [G,E] or NULL (E = 0); out: scales [G,3], opac [G],
 * colors [G, 3 + (with_mask ? 1 : 0) + E].  Backward overwrites every v_* output (v_extra may be NULL). */
int d4_assemble_fwd(const float *fg_scales, const float *bg_scales, const float *fg_opac, const float *bg_opac,
                    const float *fg_colors, const float *bg_colors, const float *extra, int Gf, int Gb, int E,
                    int with_mask, float *scales, float *opac, float *colors, d4_stream_t stream);
int d4_assemble_bwd(const float *scales, const float *opac, const float *colors, const float *v_scales,
                    const float *v_opac, const float *v_colors, int Gf, int Gb, int E, int with_mask,
                    float *v_fg_scales, float *v_bg_scales, float *v_fg_opac, float *v_bg_opac,
                    float *v_fg_colors, float *v_bg_colors, float *v_extra, d4_stream_t stream);

/* ---- f3 ("next" row): densification statistics --------------------------------------------------
 * replaces Trainer._prepare_control_step's per-render loop (flow3d/trainer.py:967-989) for the N
 * sub-exposure renders of one frame: for every Gaussian visible in render n (radii > 0)
 *   grad_norm_acc[g] += |(v_means2d.x * sx, v_means2d.y * sy)|,  vis_count[g] += 1,
 *   max_radii[g] = max(max_radii[g], radii / max(W,H))   (max_radii may be NULL: the reference's
 *   non-in-place index_put at trainer.py:989 discards that update).
 * v_means2d [N,G,2] (the .grad of meta["means2d"]), radii i32 [N,G]; sx = W/2 * batch * N,
 * sy = H/2 * batch * N (trainer.py:976-977).                                                       */
int d4_densify_stats(const float *v_means2d, const int32_t *radii, int N, int G, float sx, float sy,
                     float inv_max_wh, float *grad_norm_acc, int64_t *vis_count, float *max_radii,
                     d4_stream_t stream);

/* ---- f4 ("next" row): PWC-Net cost volume of the AlignedLoss front-end -----------------------------------
 * replaces _FunctionCorrelation.forward (flow3d/models/external/pwcnet/correlation/correlation.py:281-331) and its
 * CuPy kernels kernel_Correlation_rearrange (:8-33, twice) + kernel_Correlation_updateOutput (:35-103), reached from
 * flow3d/models/pwcnet.py:179,187.  first / second [B,C,H,W] (NCHW, contiguous) -> out [B,81,H,W]:
 *   out[b, (dy+4)*9 + (dx+4), y, x] = (1/C) sum_c first[b,c,y,x] * second[b,c,y+dy,x+dx],  dx, dy in [-4,4],
 * zero outside the image.                                                                                          */
int d4_correlation_fwd(const float *first, const float *second, int B, int C, int H, int W, float *out,
                       d4_stream_t stream);
/* replaces _FunctionCorrelation.backward (correlation.py:336-385) and its kernels kernel_Correlation_updateGradFirst
 * (:105-167) / kernel_Correlation_updateGradSecond (:169-233), one launch for the whole batch instead of two per
 * sample.  grad_out [B,81,H,W] -> grad_first / grad_second [B,C,H,W] (either may be NULL); same summation order as the
 * reference.  The reference's own caller never needs it (PWC-Net runs under no_grad, loss_utils.py:171-172).          */
int d4_correlation_bwd(const float *first, const float *second, const float *grad_out, int B, int C, int H, int W,
                       float *grad_first, float *grad_second, d4_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* D4GS_H_ */
