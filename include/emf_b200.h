/*
 * emf_b200.h -- C ABI of the B200-native multi-volume TSDF engine.
 *
 * This is the drop-in boundary for EM-Fusion's dense hot path (SURVEY.md
 * section 8b).  Every entry point takes plain pointers, sizes and a CUDA
 * stream; nothing allocates, frees or synchronises; all work is asynchronous
 * on the given stream (as the reference's level-1 operators are,
 * /root/reference/src/core/cuda/TSDF.cu:422).  Return value: EMF_OK or a
 * negative EMF_ERR_* code; nothing throws across the boundary.
 *
 * Conventions shared with the reference:
 *  - volumes are continuous float arrays, element (z*Ry + y, x), x fastest
 *    (reference src/core/TSDF.cpp:35-42); all pointers 16-byte aligned;
 *  - images are 2-D device arrays with a byte pitch (cv::cuda::GpuMat::step);
 *  - poses are RELATIVE poses already composed by the caller exactly as the
 *    reference host code does (src/core/TSDF.cpp:112,141,162): row-major 3x3
 *    rotation + translation, packed floats like cv::Matx33f::val / cv::Vec3f::val;
 *  - intrinsics are the row-major 3x3 camera matrix (cv::Matx33f).
 *
 * Paths below are relative to /root/reference.
 */
#ifndef EMF_B200_H
#define EMF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EMF_OK 0
#define EMF_ERR_INVALID (-1)     /* bad argument (null pointer, size, alignment) */
#define EMF_ERR_CUDA (-2)        /* CUDA launch error (cudaPeekAtLastError) */
#define EMF_ERR_UNSUPPORTED (-3) /* valid but unsupported (e.g. too many volumes in one batch) */

#if defined(__GNUC__)
#define EMF_API __attribute__((visibility("default")))
#else
#define EMF_API
#endif

#define EMF_MAX_VOLUMES 96 /* volumes per batched call (background + objects) */

typedef struct CUstream_st* emf_stream_t; /* == cudaStream_t; NULL = default stream */

/* 2-D device image (cv::cuda::GpuMat header): ptr, byte pitch, width, height. */
typedef struct emf_image {
    void* ptr;
    size_t pitch;
    int width;
    int height;
} emf_image;

/* Relative rigid pose: x' = R x + t. */
typedef struct emf_pose {
    float R[9];
    float t[3];
} emf_pose;

/* TSDFParams subset consumed by the path (include/EMFusion/core/data.h:32-71). */
typedef struct emf_tsdf_params {
    float max_tsdf_weight; /* 64  */
    float assoc_sigma;     /* 0.02 */
    float alpha;           /* 0.8 */
    float uni_prior;       /* 1.0 */
} emf_tsdf_params;

/* One TSDF volume (emf::TSDF / emf::ObjTSDF device state,
 * include/EMFusion/core/TSDF.h:285-300, ObjTSDF.h:197-215). */
typedef struct emf_volume {
    float* tsdf;           /* tsdfVol      Rx*Ry*Rz floats */
    float* weights;        /* tsdfWeights  Rx*Ry*Rz floats */
    const float* grads;    /* tsdfGrads (float3 per voxel) or NULL: gradients are then taken on the fly */
    const float* fg_probs; /* fgProbs (objects) or NULL (background) */
    /* Optional (objects): device array of 6 ints written by emf_compute_fg_probs_box -- inclusive voxel bounds
     * (x0, y0, z0, x1, y1, z1) of the voxels with fgProb > 0.5 (x0 > x1: none).  A ray of ObjTSDF::raycast can only
     * hit where the masked weight is positive, so rays that miss this box are not marched.  NULL = off. */
    const int32_t* fg_box;
    /* Optional acceleration state (no reference counterpart; results are unchanged).  NULL = off.
     * const_bits: three consecutive bitmaps -- a 4-voxel x-segment is all +1 / all 0 / all -1 -- one bit per
     *   segment, rows padded to whole 32-bit words: words per map = emf_bitmap_words_per_row(Rx) * Ry * Rz.
     *   Maintained by emf_integrate_volumes*; whoever zeroes the volume calls emf_reset_bitmaps; anyone else
     *   writing tsdf must clear all three maps.  Requires Rx % 4 == 0 and 16-byte aligned arrays.
     * brick_map: emf_brick_map_bytes(res) bytes, derived from const_bits by emf_update_brick_maps: one byte per
     *   8^3 brick, (m << 4) | (P << 3) | D -- the whole brick holds constant m (1: +1, 2: 0, 3: -1; 0 = mixed), so
     *   does every brick within D - 1 bricks of it, and (P) so does the 2x2x2 block of bricks starting at it.
     *   The raycast skips the march samples these certify (raycast.cu). */
    uint32_t* const_bits;
    uint8_t* brick_map;
    int res[3];            /* volumeRes (x, y, z) */
    float voxel_size;
    float truncdist;
    int id;                /* 0 = background, >0 = ObjTSDF::id */
} emf_volume;

/* ---------------------------------------------------------------------------
 * Level 1: operator API -- one symbol per reference free function.
 * ------------------------------------------------------------------------- */

/* emf::cuda::EMFusion::computePoints, include/EMFusion/core/cuda/EMFusion.cuh:39-40
 * (src/core/cuda/EMFusion.cu:29-61).  depth: W x H f32; points: W x H float3. */
EMF_API int emf_compute_points(const emf_image* depth, const emf_image* points, const float K[9], emf_stream_t stream);

/* emf::EMFusion::preprocessDepth, src/core/EMFusion.cpp:294-305 (cv::cuda::bilateralFilter + NaN patch + zero patch; five
 * launches) fused with computePoints when points_out != NULL: one launch.  depth_raw, depth_out: W x H f32 (distinct);
 * kernel_size / sigma_depth / sigma_spatial = Params::bilateral_kernel_size / _sigma_depth / _sigma_spatial (7, 0.04 m,
 * 4.5 px, include/EMFusion/core/data.h:92-94).  PARITY UNPINNED: the filter arithmetic is OpenCV-CUDA's (un-vendored,
 * absent here); its published kernel is restated (csrc/preprocess.cu). */
EMF_API int emf_preprocess_depth(const emf_image* depth_raw, const emf_image* depth_out, const emf_image* points_out,
                         const float K[9], int kernel_size, float sigma_depth, float sigma_spatial, emf_stream_t stream);

/* emf::cuda::TSDF::updateTSDF, include/EMFusion/core/cuda/TSDF.cuh:115-123
 * (src/core/cuda/TSDF.cu:327-427).  T_oc = cam_pose^-1 * pose. */
EMF_API int emf_update_tsdf(const emf_image* depth, const emf_image* assoc_weights, float* tsdf, float* weights,
                    const emf_pose* T_oc, const float K[9], const int res[3], float voxel_size,
                    float truncdist, float max_weight, emf_stream_t stream);

/* emf::TSDF::updateGradients = setTo(0) + emf::cuda::TSDF::computeTSDFGrads,
 * src/core/TSDF.cpp:120-123, include/EMFusion/core/cuda/TSDF.cuh:133-135.
 * One pass; the zero planes are written by the same kernel. grads: float3 per voxel. */
EMF_API int emf_compute_tsdf_grads(const float* tsdf, float* grads, const int res[3], emf_stream_t stream);

/* emf::cuda::TSDF::raycastTSDF, include/EMFusion/core/cuda/TSDF.cuh:160-169
 * (src/core/cuda/TSDF.cu:466-601).  T_co = pose^-1 * cam_pose.
 * raylengths is in/out (non-zero = far clip); vertices/normals (float3) and mask (u8)
 * are written at hit pixels only -- callers pre-clear, as in the reference.
 * grads may be NULL (forward differences on the fly, bit-identical);
 * fg_probs non-NULL applies ObjTSDF::raycast's weight masking
 * (src/core/ObjTSDF.cpp:209-210: weights where fgProb > 0.5, else 0) inline.
 * hit_voxel (optional, W x H x 3 int32, continuous): trunc(v*) at hit pixels. */
EMF_API int emf_raycast_tsdf(const float* tsdf, const float* grads, const float* weights, const float* fg_probs,
                     const emf_image* raylengths, const emf_image* vertices, const emf_image* normals,
                     const emf_image* mask, const emf_pose* T_co, const float K[9], const int res[3],
                     float voxel_size, float truncdist, int32_t* hit_voxel, emf_stream_t stream);

/* emf::cuda::TSDF::getVolumeVals (1-channel), include/EMFusion/core/cuda/TSDF.cuh:204-210
 * (src/core/cuda/TSDF.cu:662-726).  vals is fully overwritten (0 where not gathered). */
EMF_API int emf_get_volume_vals(const float* vol, const emf_image* points, const emf_pose* T_co, const int res[3],
                        float voxel_size, const emf_image* vals, emf_stream_t stream);

/* emf::cuda::ObjTSDF::updateFgBgProbs, include/EMFusion/core/cuda/ObjTSDF.cuh:49-56
 * (src/core/cuda/ObjTSDF.cu:29-107).  mask/occluded: W x H u8; fgbg: float2 per voxel. */
EMF_API int emf_update_fgbg_probs(const emf_image* mask, const emf_image* occluded, const float* tsdf,
                          const float* weights, float* fgbg, const emf_pose* T_oc, const float K[9],
                          const int res[3], float voxel_size, emf_stream_t stream);

/* emf::ObjTSDF::computeFgProbs, src/core/ObjTSDF.cpp:218-226 (6 OpenCV launches -> 1).
 * fg_vol_mask (u8, 255 where fgProb > 0.5) may be NULL. */
EMF_API int emf_compute_fg_probs(const float* fgbg, int64_t n_voxels, float* fg_probs, uint8_t* fg_vol_mask,
                         emf_stream_t stream);

/* emf_compute_fg_probs plus the bounding box of the foreground voxels (emf_volume::fg_box), one launch.
 * fg_box: 6 ints on the device, fully rewritten. */
EMF_API int emf_compute_fg_probs_box(const float* fgbg, const int res[3], float* fg_probs, uint8_t* fg_vol_mask,
                             int32_t* fg_box, emf_stream_t stream);

/* emf::cuda::TSDF::copyValues, include/EMFusion/core/cuda/TSDF.cuh:228-230 (src/core/cuda/TSDF.cu:768-819):
 * dst(x - offset) = src(x) for every source voxel whose target lies inside dst; the rest of dst is untouched.
 * channels = floats per voxel (1: tsdf / weights, 2: fg/bg counts, 3: gradients). */
EMF_API int emf_copy_values(const float* src, float* dst, int channels, const int offset[3], const int src_res[3],
                    const int dst_res[3], emf_stream_t stream);

/* ---------------------------------------------------------------------------
 * Level 2: class-surface operations (one call = one reference method).
 * ------------------------------------------------------------------------- */

/* emf::TSDF::computeAssociation (src/core/TSDF.cpp:125-156) when vol->fg_probs == NULL,
 * emf::ObjTSDF::computeAssociation (src/core/ObjTSDF.cpp:181-201) otherwise.
 * assoc_out: W x H f32, fully overwritten, NOT normalised.
 * assoc_mask_out (optional u8): associationMask (255 where the TSDF gather returned 0). */
EMF_API int emf_compute_association(const emf_volume* vol, const emf_image* points, const emf_pose* T_co,
                            const emf_tsdf_params* params, const emf_image* assoc_out,
                            const emf_image* assoc_mask_out, emf_stream_t stream);

/* Device part of emf::ObjTSDF::resize (src/core/ObjTSDF.cpp:116-147): the new tsdf / weights / fg-bg arrays are
 * written completely in one launch -- copied where the old grid covers the voxel (dst(x) = src(x + offset), the
 * reference's pixOffset), zero elsewhere -- replacing four setTo(0) and four copyValues passes.  src_fgbg / dst_fgbg
 * are both NULL for a volume without fg/bg counts.  The gradient volume is not moved: rebuild it with
 * emf_compute_tsdf_grads if a consumer needs it materialised. */
EMF_API int emf_resize_volume(const float* src_tsdf, const float* src_weights, const float* src_fgbg, const int src_res[3],
                      float* dst_tsdf, float* dst_weights, float* dst_fgbg, const int dst_res[3], const int offset[3],
                      emf_stream_t stream);

/* emf::TSDF::getMesh / emf::ObjTSDF::getMesh = emf::cuda::TSDF::marchingCubes (include/EMFusion/core/cuda/TSDF.cuh:255-265,
 * src/core/cuda/TSDF.cu:855-1152) with the mask passes of src/core/TSDF.cpp:356-373 / src/core/ObjTSDF.cpp:247-268 folded in:
 * a cube is meshed iff all eight corners have weight > 0 (and fgProb > 0.5 when vol->fg_probs is given).
 * Two calls, because the sizes are data dependent: emf_mesh_count leaves {number of vertices, number of triangle ints}
 * in the first two int32 of the workspace (device memory, emf_mesh_workspace_bytes(res) bytes, 16-byte aligned; -1 = more
 * than 2^31 - 1); the caller reads them, allocates, and calls emf_mesh_extract with the SAME workspace: vertices and normals
 * 3 floats each (metres, volume frame; the "normals" are the interpolated forward-difference gradients, not normalised --
 * as in the reference, whose `float3 /= float` does nothing, common.cuh:170-173), triangles as VTK polygons (3, i0, i1, i2)
 * like cv::viz::Mesh::polygons.  One vertex per crossed edge and cube, cubes in (z, y, x) order, edges ascending: the
 * reference's order.  The reference's three scratch volumes (9 B / voxel) are not needed: 8 B per row of cubes. */
EMF_API size_t emf_mesh_workspace_bytes(const int res[3]);
EMF_API int emf_mesh_count(const emf_volume* vol, void* workspace, size_t workspace_bytes, emf_stream_t stream);
EMF_API int emf_mesh_extract(const emf_volume* vol, void* workspace, size_t workspace_bytes, float* vertices, float* normals,
                     int32_t* triangles, emf_stream_t stream);

/* ---------------------------------------------------------------------------
 * Level 3: frame-level batched operations (one call = one EMFusion method).
 * vols[0] is the background when has_background != 0; objects follow in the
 * reference's iteration order (ascending id for the normaliser, list order for
 * the composite -- identical in the reference because ids only grow).
 * ------------------------------------------------------------------------- */

/* emf::EMFusion::computeAssociationWeights, src/core/EMFusion.cpp:635-670.
 * T_co[i] = pose_i^-1 * cam_pose.  assoc_out[i]: W x H f32 per volume.
 * mode 0: compute + normalise across vols (single-GPU path; sum order = array order).
 * mode 1: compute un-normalised weights and write their per-pixel sum to
 *         norm_partial (multi-GPU: all-reduce norm_partial, then emf_assoc_normalise).
 * mode 3: as mode 1, but vols[0] is left out of the sum (a replica of the background whose weight the owning rank adds). */
EMF_API int emf_assoc_weights(int n_vol, const emf_volume* vols, const emf_pose* T_co, const emf_image* points,
                      const emf_tsdf_params* params, const emf_image* assoc_out, int mode,
                      const emf_image* norm_partial, emf_stream_t stream);

/* Divide every image by norm with x/0 -> 0 (src/core/EMFusion.cpp:659-665). */
EMF_API int emf_assoc_normalise(int n_img, const emf_image* assoc_io, const emf_image* norm, emf_stream_t stream);

/* Per-volume raycast of emf::EMFusion::raycast (src/core/EMFusion.cpp:745-754) for a whole
 * batch in one launch.  Per volume i: ray_out[i] (f32, fully written inside rect, 0 = no hit),
 * mask_out[i] (u8), vert_out[i]/norm_out[i] (float3, hit pixels only).  rects (n_vol x 4 ints:
 * x0, y0, x1, y1, exclusive upper) bound the pixels each volume can cover; pixels outside a
 * volume's rect are not touched and must be treated as "no hit" by the consumer.
 * stats (optional, 8 x uint64 on the device, accumulated): TSDF samples taken, march samples skipped by jumps
 * (constant bricks / certified slabs), jumps, weight samples -- the raycast's roofline numerator --, then warp iterations
 * of the march loop and the lanes active in them, for certified ([4], [5]) and plain ([6], [7]) rays. */
EMF_API int emf_raycast_volumes(int n_vol, const emf_volume* vols, const emf_pose* T_co, const float K[9],
                        const int* rects, const emf_image* ray_out, const emf_image* vert_out,
                        const emf_image* norm_out, const emf_image* mask_out, uint64_t* stats,
                        emf_stream_t stream);

/* emf_raycast_volumes with a workspace of emf_raycast_workspace_bytes(W, H) bytes (device memory, 16-byte aligned,
 * zeroed once before its first use).  With it, a pre-pass (k_ray_certify) certifies, per 8 x 4 pixel tile and per slab of voxels along
 * the volume axis the camera looks along, that every voxel a march sample of the tile can touch there holds exactly +1;
 * rays of the first volume without fg_probs (the background) then skip those samples -- no-ops of the reference's march
 * loop, src/core/cuda/TSDF.cu:523-572 -- in closed form.  Same results, bit for bit (stats[1], [2] count what was skipped);
 * the environment variable EMF_RAY_CERT=0 turns the pre-pass off for A/B measurements. */
EMF_API int emf_raycast_volumes_ws(int n_vol, const emf_volume* vols, const emf_pose* T_co, const float K[9],
                           const int* rects, const emf_image* ray_out, const emf_image* vert_out,
                           const emf_image* norm_out, const emf_image* mask_out, uint64_t* stats,
                           void* workspace, size_t workspace_bytes, emf_stream_t stream);
EMF_API size_t emf_raycast_workspace_bytes(int width, int height);
/* The same with the workspace's two uses selected separately (emf_raycast_volumes_ws = both).  The workspace must be ZEROED once
 * before its first use and then kept between frames.
 * EMF_RAY_CERTIFICATE: the pre-pass above.
 * EMF_RAY_SCHEDULE: the tiles (16 x 8 pixels) of the FIRST volume are marched longest first, in the order of the cost (march
 *   iterations of the tile's longest ray) the previous call recorded in the workspace; the blocks of the other volumes follow
 *   (EMF_RAY_FRONT=<percent> places them after that share of the first volume's tiles instead).  The launch otherwise ends with a few SMs marching the longest rays (rim of the frustum, shadow seams:
 *   3-4 x the median) that happen to start last.  Scheduling only: no result depends on the order. */
#define EMF_RAY_CERTIFICATE 1u
#define EMF_RAY_SCHEDULE 2u
#define EMF_RAY_WIDE 4u          /* the first volume without fg_probs is marched with four lanes per ray (eight samples of a ray in
                                  * flight; needs no workspace): shortens the longest dependent chain where a GPU traces a band of
                                  * the frame and is otherwise idle (multi-GPU), costs instructions where it is not */
#define EMF_RAY_SCHEDULE_DEFER 8u /* with EMF_RAY_SCHEDULE: record the costs and use the order, but leave the sort of the next order to
                                  * emf_raycast_schedule_update (e.g. on a side stream, off the frame's critical path) */
EMF_API int emf_raycast_volumes_opt(int n_vol, const emf_volume* vols, const emf_pose* T_co, const float K[9],
                            const int* rects, const emf_image* ray_out, const emf_image* vert_out,
                            const emf_image* norm_out, const emf_image* mask_out, uint64_t* stats,
                            void* workspace, size_t workspace_bytes, unsigned options, emf_stream_t stream);
/* The sort that EMF_RAY_SCHEDULE_DEFER left out: turns the costs the last emf_raycast_volumes_opt call recorded into the order
 * of the next one.  rect0 = the first volume's rectangle of that call (NULL = the whole frame).  Must run after that call and
 * before the next one (the caller orders the streams). */
EMF_API int emf_raycast_schedule_update(int width, int height, const int rect0[4], void* workspace, size_t workspace_bytes,
                                emf_stream_t stream);

/* Compositing of emf::EMFusion::raycast, src/core/EMFusion.cpp:760-794, in one launch.
 * Objects i = 0..n_obj-1 in list order with ids[i]; background images bg_*.
 * Outputs: ray/vert/norm/seg (seg u8) and vis_count[n_obj] (int32, device; zeroed by the call). */
EMF_API int emf_raycast_composite(int n_obj, const int* ids, const int* rects, const emf_image* obj_ray,
                          const emf_image* obj_vert, const emf_image* obj_norm, const emf_image* obj_mask,
                          const emf_image* bg_ray, const emf_image* bg_vert, const emf_image* bg_norm,
                          const emf_image* bg_mask, int boundary, const emf_image* ray, const emf_image* vert,
                          const emf_image* norm, const emf_image* seg, int32_t* vis_count, emf_stream_t stream);

/* Multi-GPU composite: every rank runs emf_raycast_composite on its own objects against an EMPTY background (mask == 0)
 * and the n_parts results (ray / vert / norm / seg images, e.g. gathered with NCCL) are merged here, on the rank that
 * owns the background, in (raylength, list order) order -- the outcome of the reference's sequential loop over all
 * objects (src/core/EMFusion.cpp:760-771) -- followed by the background rule, the fill and the visibility counts
 * (:773-794).  ids: all n_obj objects in global list order; vis_count[n_obj] (device, zeroed by the call).
 * band_rows > 0 (replicated background): rank p traced rows [p * band_rows, (p + 1) * band_rows) of the background;
 * band_ray/vert/norm/mask[p] point at its band (continuous rows of W floats / 3 W floats / W bytes) and the bg_* images
 * are assembled from the bands by this call (they are then outputs).  band_rows == 0: bg_* are complete inputs. */
EMF_API int emf_composite_merge(int n_parts, const emf_image* part_ray, const emf_image* part_vert, const emf_image* part_norm,
                        const emf_image* part_seg, int n_obj, const int* ids, const emf_image* bg_ray,
                        const emf_image* bg_vert, const emf_image* bg_norm, const emf_image* bg_mask, int boundary,
                        const emf_image* ray, const emf_image* vert, const emf_image* norm, const emf_image* seg,
                        int32_t* vis_count, int band_rows, const void* const* band_ray, const void* const* band_vert,
                        const void* const* band_norm, const void* const* band_mask, emf_stream_t stream);

/* emf::EMFusion::integrateDepth, src/core/EMFusion.cpp:865-889, one launch for all volumes.
 * T_oc[i] = cam_pose^-1 * pose_i; assoc[i] = that volume's association image. */
EMF_API int emf_integrate_volumes(int n_vol, const emf_volume* vols, const emf_pose* T_oc, const float K[9],
                          const emf_image* depth, const emf_image* assoc, float max_weight,
                          emf_stream_t stream);

/* emf_integrate_volumes with device-side visibility gating and optional counters.
 * gates[i] < 0: volume i is always integrated (background); otherwise it is integrated iff
 * gate_counts[gates[i]] > gate_thresh, read ON THE DEVICE -- gate_counts is the vis_count array that
 * emf_raycast_composite wrote earlier on the same stream, so the visibility filter of
 * emf::EMFusion::integrateDepth (src/core/EMFusion.cpp:869-872) needs no device->host round trip.
 * stats (optional, 8 x uint64 on the device, accumulated): voxels updated, marked occluded-unseen (-1),
 * occluded but already seen, check-only (behind camera / invalid depth), projected outside the image inside
 * the culling interval -- the exact-bytes roofline numerator -- and, from the segment-level kernel only: voxels that took
 * the exact per-voxel path, segments decided free, segments decided occluded. */
EMF_API int emf_integrate_volumes_gated(int n_vol, const emf_volume* vols, const emf_pose* T_oc, const float K[9],
                                const emf_image* depth, const emf_image* assoc, float max_weight,
                                const int32_t* gate_counts, const int* gates, int gate_thresh, uint64_t* stats,
                                emf_stream_t stream);

/* emf_integrate_volumes_gated with a workspace of emf_integrate_workspace_bytes(W, H) bytes (device memory, 16-byte
 * aligned, contents irrelevant): the call builds a (min, max) pyramid of the depth image in it and classifies whole
 * 4-voxel segments against it before any per-voxel work (integrate.cu).  Same results, bit for bit; without a
 * workspace (NULL) or with a camera matrix that is not a plain pinhole the per-voxel kernel runs. */
EMF_API int emf_integrate_volumes_ws(int n_vol, const emf_volume* vols, const emf_pose* T_oc, const float K[9],
                             const emf_image* depth, const emf_image* assoc, float max_weight,
                             const int32_t* gate_counts, const int* gates, int gate_thresh, uint64_t* stats,
                             void* workspace, size_t workspace_bytes, emf_stream_t stream);
EMF_API size_t emf_integrate_workspace_bytes(int width, int height);

/* emf_integrate_volumes_ws in two parts, so that the part that depends on the depth image and the poses alone can run on
 * another stream next to the raycast: phase 1 = depth pyramid + classification of whole 32 x 4 x 4 voxel bricks against it
 * (every volume: the visibility gate is not known yet), phase 2 = the integration itself (gated); phase 0 = both
 * (= emf_integrate_volumes_ws).  Same arguments in both calls; the workspace carries the state between them. */
EMF_API int emf_integrate_volumes_phase(int n_vol, const emf_volume* vols, const emf_pose* T_oc, const float K[9],
                                const emf_image* depth, const emf_image* assoc, float max_weight,
                                const int32_t* gate_counts, const int* gates, int gate_thresh, uint64_t* stats,
                                void* workspace, size_t workspace_bytes, int phase, emf_stream_t stream);

/* Derives brick_map from const_bits for every volume that carries both (others are skipped); two launches.
 * Call after integrating and before raycasting. */
EMF_API int emf_update_brick_maps(int n_vol, const emf_volume* vols, emf_stream_t stream);

/* Bytes of emf_volume::brick_map for a volume of this resolution (0 if the resolution is invalid). */
EMF_API size_t emf_brick_map_bytes(const int res[3]);

/* Acceleration state of a freshly zeroed volume: every segment / brick is "all 0". */
EMF_API int emf_reset_bitmaps(const emf_volume* vol, emf_stream_t stream);

/* 32-bit words per volume row in a segment bitmap. */
static inline int emf_bitmap_words_per_row(int rx) { return (rx / 4 + 31) / 32; }

/* Screen-space rectangle (x0,y0,x1,y1) that bounds every pixel whose ray can enter the
 * volume's raycast box; full frame if a corner is behind the camera.  Host-side helper. */
EMF_API int emf_volume_screen_rect(const int res[3], float voxel_size, const emf_pose* T_co, const float K[9],
                           int width, int height, int rect_out[4]);

/* W x H f32 image <- value (the reference's setTo on association images, src/core/EMFusion.cpp:55,916). */
EMF_API int emf_fill_image_f32(const emf_image* img, float value, emf_stream_t stream);

/* ---------------------------------------------------------------------------
 * Tracker: the device part of one Levenberg-Marquardt iteration of
 * emf::TSDF's SDF tracker for a batch of volumes in one launch.
 *
 * Replaces, per volume and iteration, TSDF::computeGradients / computeTSDFVals / computeTSDFWeights /
 * computeHuberWeights / normalizeTSDFWeights / combineWeights / computeHessians / reduceHessians' device part /
 * computeError (src/core/TSDF.cpp:194-265, 375-394): emf::cuda::TSDF::computePoseGradients, getVolumeVals x 2,
 * computeAb, multSingletonCol x 2 (include/EMFusion/core/cuda/TSDF.cuh:181-253) and ~12 OpenCV-CUDA launches.
 *
 * modes[i]: 0 = leave volume i alone (converged), 1 = linearise at T_co[i] (evaluateGradient), 2 = error only
 * (the trial pose of TSDF::computePoseUpdate, src/core/TSDF.cpp:312-315, or an iteration with evaluateGradient false).
 * Per volume i:
 *   assoc[i]        in  (mode 1)  association weights of the volume (W x H f32)
 *   int_weights[i]  mode 1: out, mode 2: in -- Huber x min(w, max) x assoc per pixel, BEFORE the NORM_INF scale and only
 *                   on the tile rectangle recorded in records[i][45..46] (everything outside is 0 by construction and is
 *                   not written unless one of the optional images below is requested).  The reference's intWeights =
 *                   this x records[i][44] inside the rectangle, 0 outside: emf_track_normalised_weights forms it.
 *   tsdf_vals[i]    out, optional (array or entry.ptr NULL): tsdfVals
 *   track_weights[i] out, optional, mode 1: trackWeights (Huber)
 *   pose_grads[i]   out, optional, mode 1: `grads`, W*H x 6 floats continuous (bit-identical to the reference's)
 *   records + i * EMF_TRACK_RECORD (device floats): [0..35] A row-major 6x6, [36..41] b, [42] error
 *                   sum f^2 w, [43] max of the clamped integration weights, [44] the NORM_INF scale, [45], [46] (bit
 *                   patterns of two uint32) the rectangle of 32 x 8 pixel tiles int_weights was written on:
 *                   tx0 | tx1 << 16 and ty0 | ty1 << 16.
 *                   Mode 1 writes all of it; mode 2 only [42], using the stored scale and int_weights.
 * K (optional): the camera matrix the points were un-projected with (emf_compute_points).  With it, image tiles whose
 * pixels cannot see a volume's box are not gathered at all (their outputs are the zeros the reference computes there).
 * workspace: emf_track_workspace_bytes(n_vol) bytes of device memory (16-byte aligned) that was zeroed ONCE with
 * emf_track_workspace_init; the call leaves it ready for the next one.  Sums are folded in a fixed order:
 * the same inputs give the same bits. */
#define EMF_TRACK_RECORD 48
EMF_API size_t emf_track_workspace_bytes(int n_vol);
EMF_API int emf_track_workspace_init(void* workspace, size_t workspace_bytes, emf_stream_t stream);
EMF_API int emf_track_linearise(int n_vol, const emf_volume* vols, const emf_pose* T_co, const int* modes,
                        const emf_image* points, const float K[9], const emf_image* assoc, float huber_thresh,
                        float max_tsdf_weight,
                        const emf_image* int_weights, const emf_image* tsdf_vals, const emf_image* track_weights,
                        float* const* pose_grads, float* records, void* workspace, size_t workspace_bytes,
                        emf_stream_t stream);
/* The whole Levenberg-Marquardt loop on the device.  Per volume a state record in DEVICE memory, initialised by the host as
 * emf::TSDF::prepareTracking does (src/core/TSDF.cpp:170-192): R, t = the orthonormalised rel_pose_CO; converged = 0,
 * first_iteration = 1, evaluate_gradient = 1, trial_pending = 0, nu = nu_init, counters 0.  One call enqueues n_iterations
 * iterations of every volume: per iteration two launches of the kernel behind emf_track_linearise, whose last CTA per volume
 * also runs the host part of the iteration -- reduceHessians' convergence test, the damped 6 x 6 solve, SE(3) exp, the
 * step-size test (launch 1: linearise or error-only at the current pose, then the trial pose) and the gain ratio with
 * accept / reject and the damping update (launch 2: error at the trial pose), src/core/TSDF.cpp:267-338 statement by
 * statement.  Nothing is read back in between; converged volumes cost nothing; the caller polls `converged` (e.g. with an
 * asynchronous copy of the states every few iterations) and reads R, t when done.  T_co_hint: the starting poses on the
 * host (only used to size each volume's grid). */
typedef struct emf_track_state {
    double R[9], t[3];          /* rel_pose_CO, in/out */
    double R_old[9], t_old[3];  /* the pose a rejected trial step returns to */
    double mu, nu, rho;
    float A[36], b[6], x[6];
    float err, err_new;
    int converged, first_iteration, evaluate_gradient, trial_pending;
    int iterations, linearisations;
} emf_track_state;
typedef struct emf_track_lm_params {  /* TSDFParams, include/EMFusion/core/data.h:32-71 */
    float tau, eps1, eps2, nu_init, huber_thresh, max_tsdf_weight;
} emf_track_lm_params;
EMF_API int emf_track_iterate(int n_vol, const emf_volume* vols, emf_track_state* states, const emf_pose* T_co_hint,
                      const emf_image* points, const float K[9], const emf_image* assoc, const emf_track_lm_params* lm,
                      const emf_image* int_weights, float* records, void* workspace, size_t workspace_bytes,
                      int n_iterations, emf_stream_t stream);

/* out <- int_weights x record[44]: intWeights as emf::TSDF::getTrackingWeights sees it (src/core/TSDF.cpp:340-343). */
EMF_API int emf_track_normalised_weights(const emf_image* int_weights, const float* record, const emf_image* out,
                                 emf_stream_t stream);

/* ---------------------------------------------------------------------------
 * Level 4: the native frame engine -- one host call per frame.
 *
 * emf::EMFusion::processFrame's hot part (src/core/EMFusion.cpp:76-103 minus tracking and Mask R-CNN):
 * computePoints, computeAssociationWeights, raycast (+ composite), integrateDepth, as five launches on one stream
 * with no device->host read on the path: the visibility filter of integrateDepth (:869-872) is evaluated on the
 * device from the composite's counters.  The engine owns only image scratch; volumes stay caller-owned.
 * ------------------------------------------------------------------------- */
typedef struct emf_engine emf_engine;

typedef struct emf_engine_config {
    int width, height;          /* Params::frameSize */
    float K[9];                 /* Params::intr */
    emf_tsdf_params params;     /* Params::tsdfParams */
    int boundary;               /* Params::boundary (20) */
    int visibility_thresh;      /* Params::visibilityThresh (1600) */
} emf_engine_config;

/* phases of emf_engine_frame (any combination, executed in this order) */
#define EMF_FRAME_POINTS 0x1u          /* points <- depth */
#define EMF_FRAME_ASSOC 0x2u           /* association of every volume, normalised (single GPU) */
#define EMF_FRAME_ASSOC_PARTIAL 0x4u   /* un-normalised weights + their per-pixel sum in EMF_IMG_NORM (multi-GPU, before the all-reduce) */
#define EMF_FRAME_NORMALISE 0x8u       /* divide by EMF_IMG_NORM (multi-GPU, after the all-reduce) */
#define EMF_FRAME_RAYCAST 0x10u        /* raycast of every volume */
#define EMF_FRAME_COMPOSITE 0x20u      /* composite + visibility counters (against an empty background if there is none) */
#define EMF_FRAME_INTEGRATE 0x40u      /* integrate the background and the visible objects */
#define EMF_FRAME_INTEGRATE_ALL 0x80u  /* ... every volume regardless of visibility (first frame) */
#define EMF_FRAME_COMPOSITE_NOBG 0x100u /* ... composite against an empty background even if there is one (multi-GPU pre-composite) */
#define EMF_FRAME_ASSOC_PARTIAL_NOBG 0x400u /* ... as ASSOC_PARTIAL, the background (a replica) left out of the partial sum */
#define EMF_FRAME_TIMED 0x200u         /* record stage events for emf_engine_stage_ms */
#define EMF_FRAME_INTEGRATE_BG 0x800u   /* integrate only the background (no visibility counter gates it: multi-GPU, before the merge) */
#define EMF_FRAME_INTEGRATE_OBJ 0x1000u /* integrate only the (visible) objects */
#define EMF_FRAME_ALL (EMF_FRAME_POINTS | EMF_FRAME_ASSOC | EMF_FRAME_RAYCAST | EMF_FRAME_COMPOSITE | EMF_FRAME_INTEGRATE)

/* engine-owned images (emf_engine_image) */
enum {
    EMF_IMG_POINTS = 0, EMF_IMG_NORM, EMF_IMG_RAY, EMF_IMG_VERT, EMF_IMG_NORMALS, EMF_IMG_SEG,   /* frame level */
    EMF_IMG_VOL_ASSOC = 16, EMF_IMG_VOL_RAY, EMF_IMG_VOL_VERT, EMF_IMG_VOL_NORMALS, EMF_IMG_VOL_MASK /* per volume (index) */
};

EMF_API emf_engine* emf_engine_create(const emf_engine_config* cfg);   /* NULL on failure */
EMF_API void emf_engine_destroy(emf_engine* e);

/* The volumes of the following frames: vols[0] is the background if has_background, objects follow in list order.
 * Descriptors are copied.  When the count changes the image scratch is reallocated: images of volumes that were
 * already there (same id) keep their content, a new volume's association image starts at 1 (src/core/EMFusion.cpp:55,916),
 * everything else at 0. */
EMF_API int emf_engine_set_volumes(emf_engine* e, int n_vol, const emf_volume* vols, int has_background, emf_stream_t stream);

/* One frame (or some of its phases).  depth: W x H f32 on the device.  T_co[i] = pose_i^-1 * cam_pose,
 * T_oc[i] = cam_pose^-1 * pose_i for the n_vol volumes (either may be NULL if no phase needs it). */
EMF_API int emf_engine_frame(emf_engine* e, const emf_image* depth, const emf_pose* T_co, const emf_pose* T_oc, unsigned flags,
                     emf_stream_t stream);

/* One frame with HOST buffers at both ends: emf::EMFusion::processFrame between `depth_raw.upload(frame.getDepth())`
 * (src/core/EMFusion.cpp:72) and the host-side consumers of the model view (getLastMasks / the renderers, :131-200).
 * depth_host: W x H f32, continuous, page-locked.  The call queues: upload (own stream) -> emf_engine_frame(flags) on
 * `stream` -> download of the composite's segmentation (u8) and ray lengths (f32) into engine-owned page-locked memory
 * (own stream; download = 0: skipped, e.g. on a rank that does not hold the composite) and returns at once with a ticket.
 * emf_engine_result_host blocks until that frame's result has arrived and returns pointers to it (valid until two further
 * frames have been submitted: the engine double-buffers).  The upload of frame n + 1 and the download of frame n thus run
 * under the kernels of their neighbours, where the reference uploads, processes and downloads strictly in sequence.
 * emf_engine_depth_slot: the device copy of a submitted frame's depth (for callers that drive the phases themselves). */
EMF_API int emf_engine_submit_host(emf_engine* e, const float* depth_host, const emf_pose* T_co, const emf_pose* T_oc,
                           unsigned flags, int download, emf_stream_t stream, long long* ticket_out);
EMF_API int emf_engine_result_host(emf_engine* e, long long ticket, const uint8_t** seg_host, const float** ray_host);
EMF_API int emf_engine_depth_slot(emf_engine* e, long long ticket, emf_image* depth_out);

/* Device times of {association, raycast + composite, integrate} of the last EMF_FRAME_TIMED frame (waits for it). */
EMF_API int emf_engine_stage_ms(emf_engine* e, float ms[3]);

/* View of an engine-owned image (valid until the next emf_engine_set_volumes that changes the volume count). */
EMF_API int emf_engine_image(emf_engine* e, int what, int index, emf_image* out);

/* Visibility counters of the last composite, one per object in list order: on the device (for gating and for the
 * multi-GPU path to overwrite), and copied to the host (waits for the asynchronous copy only). */
EMF_API int32_t* emf_engine_vis_counts_device(emf_engine* e);
EMF_API int emf_engine_vis_counts(emf_engine* e, int32_t* counts_out, int n);

/* Restrict the background's raycast to image rows [y0, y1) (multi-GPU with a replicated background: every rank traces
 * a band of rows, the bands are gathered on the compositing rank).  y0 = 0, y1 = height restores the full frame. */
EMF_API int emf_engine_set_background_rows(emf_engine* e, int y0, int y1);

/* Integrate volume vol_index at the next EMF_FRAME_INTEGRATE whatever its visibility counter says (an object created
 * after the last composite: emf::EMFusion::createObj adds it to vis_objs, src/core/EMFusion.cpp:918). */
EMF_API int emf_engine_force_integrate(emf_engine* e, int vol_index);

/* ---------------------------------------------------------------------------
 * Multi-GPU exchange over NVLink peer memory (csrc/xchg.cu): one process per GPU, every rank owns an exchange buffer
 * that its peers open through CUDA IPC.  Replaces the all-reduce of the association normaliser, the gather of the
 * pre-composites and the broadcast of the visibility counters by flags + direct peer loads / stores
 * on the frame's stream.  Flags carry the frame number (they only grow), so nothing is ever reset.
 * ------------------------------------------------------------------------- */
EMF_API int emf_xchg_alloc(size_t bytes, void** ptr_out, unsigned char handle_out[64]);   /* cudaMalloc + zero + IPC handle */
EMF_API int emf_xchg_open(const unsigned char handle[64], void** ptr_out);               /* a peer's buffer as a device pointer */
EMF_API int emf_xchg_close(void* peer_ptr);
EMF_API int emf_xchg_free(void* ptr);
/* *flags[i] = value for i < n (n <= 16; local or peer pointers), after everything queued before it on the stream. */
EMF_API int emf_xchg_signal(int n, uint32_t* const* flags, uint32_t value, emf_stream_t stream);
/* The stream waits until flags[i] >= value for every i < n (n <= 32; LOCAL memory).  ">=" is the sign of the 32-bit
 * difference (int32_t)(flags[i] - value) >= 0: flags are frame counters that only grow and may wrap; a waiter must not be
 * more than 2^31 - 1 signals behind its producer.  After timeout_s seconds it gives up and stores 1 + i in *err (device
 * memory, caller-zeroed) instead of hanging; the host mirror (PeerExchange.poll_errors) turns that into an exception in
 * the next frame. */
EMF_API int emf_xchg_wait(const uint32_t* flags, int n, uint32_t value, uint32_t* err, double timeout_s, emf_stream_t stream);
/* out = parts[0] + parts[1] + ... in this order (W x H f32, continuous parts, 16-byte aligned; local or peer pointers). */
EMF_API int emf_xchg_sum_images(int n_parts, const float* const* parts, const emf_image* out, emf_stream_t stream);
/* dst[r][i] = src[i] for i < count, r < n_dst (the visibility counters into every rank's buffer). */
EMF_API int emf_xchg_scatter_u32(const uint32_t* src, int count, int n_dst, uint32_t* const* dst, emf_stream_t stream);

/* The all-reduce of the association normaliser fused into its consumer: the kernel waits for flags[i] >= value, i < n_parts
 * (local memory, polled inside the kernel by every CTA; time-out as emf_xchg_wait), sums parts[0..n_parts) (W x H f32,
 * continuous; local or peer pointers) in this order into norm_out and divides the n_img images by it with x/0 -> 0
 * (src/core/EMFusion.cpp:653-665). */
EMF_API int emf_assoc_normalise_parts(int n_img, const emf_image* assoc_io, int n_parts, const float* const* parts,
                              const emf_image* norm_out, const uint32_t* flags, uint32_t value, uint32_t* err,
                              double timeout_s, emf_stream_t stream);
/* Engine side of it: EMF_FRAME_ASSOC_PARTIAL* writes the partial normaliser to `target` (a slot of this rank's exchange
 * buffer; NULL = the engine's own image again); emf_engine_normalise_from_parts = emf_assoc_normalise_parts over the
 * engine's association images and EMF_IMG_NORM. */
EMF_API int emf_engine_set_partial_norm_target(emf_engine* e, const emf_image* target);
/* Multi-GPU: where the composite (EMF_FRAME_COMPOSITE*) writes {ray f32, vert float3, normals float3, seg u8} and where the
 * background's raycast writes {ray, vert, normals, mask} -- four W x H images each, e.g. inside this rank's exchange buffer,
 * so that the merging rank reads them in place; NULL = the engine's own images again. */
EMF_API int emf_engine_set_composite_target(emf_engine* e, const emf_image* target4);
EMF_API int emf_engine_set_background_target(emf_engine* e, const emf_image* target4);
/* Multi-GPU: gate the integrate of volume i by counts[index[i]] > visibility_thresh (device counters of the MERGED frame,
 * e.g. in this rank's exchange buffer; index[i] = the volume's position in the global object list, ignored for the
 * background) instead of the engine's own counters; counts == NULL restores them.  n = the engine's volume count. */
EMF_API int emf_engine_set_gate_source(emf_engine* e, const int32_t* counts, const int* index, int n);
/* Engine options.  EMF_OPT_RAY_CERTIFICATE: 1 = the background's raycast uses the ray-space certificate and the
 * four-lanes-per-ray march (emf_raycast_volumes_ws), 0 = the plain march, -1 = the environment variable EMF_RAY_CERT decides. */
#define EMF_OPT_RAY_CERTIFICATE 1
#define EMF_OPT_RAY_WIDE 2          /* 1 = four lanes per background ray (EMF_RAY_WIDE above), 0 = off, -1 = environment variable EMF_RAY_WIDE */
EMF_API int emf_engine_set_option(emf_engine* e, int option, int value);
EMF_API int emf_engine_normalise_from_parts(emf_engine* e, int n_parts, const float* const* parts, const uint32_t* flags,
                                    uint32_t value, uint32_t* err, double timeout_s, emf_stream_t stream);

/* Library identification: returns a static string "emf_b200 <version> sm_100a". */
EMF_API const char* emf_version(void);

#ifdef __cplusplus
}
#endif
#endif /* EMF_B200_H */
