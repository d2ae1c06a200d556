/*
 * kb_b200.h -- C ABI of the B200-native (sm_100a) post-network hot path of keypoint_bench.
 *
 * The reference (linyicheng1/keypoint_bench) is pure Python: its "plugin interface" for this
 * path is the set of module-level functions in utils/extracter.py, utils/matcher.py,
 * utils/projection.py and tasks/repeatability.py.  Each entry point below replaces the body of
 * one of them (file:line cited per function); `keypoint_bench_b200/utils/*.py` are the
 * same-named Python shims a maintainer drops in (see INTEGRATION.md for the ctypes binding).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - all tensors are dense, row-major, float32 / int32 unless stated;
 *   - the library never allocates or frees caller-visible memory: outputs and a scratch
 *     workspace (size from the matching *_workspace_bytes) are passed in;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*) unless its
 *     comment says it synchronises; no hidden cudaDeviceSynchronize;
 *   - return value: 0 = OK, < 0 = argument / workspace error (KB_ERR_*), > 0 = cudaError_t.
 */
#ifndef KB_B200_H
#define KB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KB_OK 0
#define KB_ERR_BAD_ARG (-1)
#define KB_ERR_WORKSPACE (-2)
#define KB_ERR_UNSUPPORTED (-3)

#if defined(__GNUC__)
#define KB_API __attribute__((visibility("default")))
#else
#define KB_API
#endif

typedef void* kb_stream_t; /* cudaStream_t */

/* Library version (major*10000 + minor*100 + patch). */
KB_API int kb_version(void);
/* Human-readable text for a return code of this library. */
KB_API const char* kb_error_string(int code);
/* Process-wide experiment knobs (tests and A/B measurements only; product callers never touch them).  Results never
 * depend on KB_KNOB_TC_CLUSTER or KB_KNOB_REP_NO_SORT; KB_KNOB_TC_DEBUG is a bit mask of TIMING experiments of the
 * tensor-core search kernel and makes results WRONG with bits 1, 2 or 8 set (4 = in-kernel pipeline wait counters).
 * Returns the previous value, or KB_ERR_BAD_ARG for an unknown knob. */
#define KB_KNOB_TC_CLUSTER 1  /* 2 = CTA-pair (cta_group::2) tensor-core search kernel; default 1 */
#define KB_KNOB_TC_DEBUG 2    /* default 0 */
#define KB_KNOB_REP_NO_SORT 3 /* 1 = kb_repeat_counts uses the tile-walking kernels instead of the sorted sweeps */
#define KB_KNOB_TC_ONE_PASS 4 /* 1 = tensor-core cross-check from column-group maxima of ONE Gram pass (redux.sync in the
                               * epilogue) instead of a second Gram with rows and columns swapped; same pairs, measured
                               * slower, kept for A/B measurements.  Set it before sizing the workspace. */
#define KB_KNOB_TC_BF16X3 6   /* operand split of the tensor-core Gram: 0 = automatic (fp16 halves and two products,
                               * (hi_x + lo_x).hi_y, for D > 64; bf16 halves and three products hi.hi + lo.hi + hi.lo for
                               * D <= 64), 1 = bf16 x 3 always, 2 = fp16 x 2 always; same pairs */
#define KB_KNOB_SAMPLE_4CH 7  /* 1 = kb_sample_desc stages low-resolution maps four channels per CTA (the older kernel; same bits) */
#define KB_KNOB_SPARSE_PROF 5 /* 1 = the per-map resolve kernel of kb_detect records clock64 at its phase boundaries */
KB_API int kb_debug_knob(int knob, int value);
/* Diagnostics: the 16 values map 0's CTA of the last kb_detect recorded under KB_KNOB_SPARSE_PROF (synchronises):
 * [0..7] clock64 at start / after the cut / candidates loaded / cell grid built / decisions done / kept compacted /
 * sorted / rows written, [8] candidates on chip, [9] kept interior, [10] [11] list lengths, [12] attempt. */
KB_API int kb_debug_sparse_prof(long long* host_out);

/* ---------------------------------------------------------------------------------------------
 * Stage 1a -- fast_nms(image_probs, nms_dist, max_iter, min_value)   utils/extracter.py:6-100
 *
 * Round-faithful NMS on B independent [H,W] maps: per round, a pixel is a local maximum iff it is
 * the FIRST maximal entry of its zero-padded (2r+1)^2 window (extracter.py:54-70); the loop stops
 * when the total number of maxima (summed over the batch, extracter.py:73) repeats; otherwise every
 * pixel with another maximum within Chebyshev distance r becomes `min_value` (extracter.py:81-96).
 * Valid for any sign of the scores, any min_value, any max_iter (-1 = until the count repeats).
 * One cooperative launch; termination is decided on the device (no host round trip per round).
 * `out` may not alias `score`.  `rounds` (device int32[1], may be NULL) receives the number of
 * suppression rounds executed.  nms_dist == 0 copies the input (the reference returns its input).
 * ------------------------------------------------------------------------------------------- */
KB_API size_t kb_fast_nms_workspace_bytes(int B, int H, int W);
KB_API int kb_fast_nms(const float* score, float* out, int B, int H, int W, int nms_dist, int max_iter,
                float min_value, int* rounds, void* ws, size_t ws_bytes, kb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Stage 1b -- remove_border_points + prob_map_to_positions_with_prob + the top-k / min_score tail
 * of detection()                                   utils/extracter.py:164-190, 129-161, 217-220
 *
 * For each of the B maps: pixels inside the border (border_dist rows/cols removed on every side)
 * with value > threshold are listed in raster order as (x=(col+.5)/W, y=(row+.5)/H, p); if more
 * than top_k qualify (top_k > 0) the top_k by (score desc, raster index asc) are kept, sorted in
 * that order, else raster order is kept; rows with p <= min_score are then dropped when
 * min_score > 0.  Outputs are padded to `cap` rows per map: xyp [B,cap,3], raster [B,cap] (flat
 * row*W+col), count [B] (rows written), total [B] (may be NULL; number of qualifying pixels before
 * truncation -- if it exceeds cap while top_k <= 0 the list is truncated to the first cap).
 * ------------------------------------------------------------------------------------------- */
KB_API size_t kb_select_workspace_bytes(int B, int H, int W, int top_k);
KB_API int kb_select(const float* nms_map, int B, int H, int W, int border_dist, float threshold,
              float min_score, int top_k, int cap, float* xyp, int* raster, int* count, int* total,
              void* ws, size_t ws_bytes, kb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * detection(score_map, params)                                       utils/extracter.py:193-221
 *
 * NMS + border + threshold + top-k + min_score for B independent maps in one asynchronous call
 * (the batched sibling of the per-image drop-in).  Output layout as kb_select with cap = top_k.
 * Non-negative maps with threshold >= 0 take the sparse exact path (greedy NMS in priority order
 * over the pixels that can reach the top_k, equal to the fixed point of the reference's rounds);
 * anything else -- or any image the sparse path could not certify -- runs the round-faithful
 * kernel of kb_fast_nms on the device, so results are exact in all cases.
 * `path` (device int32[B], may be NULL) reports per map which path produced the result
 * (1 = sparse, 2 = round-faithful).
 * ------------------------------------------------------------------------------------------- */
KB_API size_t kb_detect_workspace_bytes(int B, int H, int W, int nms_dist, int top_k, float threshold);
KB_API int kb_detect(const float* score, int B, int H, int W, int nms_dist, int border_dist, float threshold,
              float min_score, int top_k, float* xyp, int* raster, int* count, int* path, void* ws,
              size_t ws_bytes, kb_stream_t stream);
/* Measurement hook: kb_detect restricted to the kernels selected by `phases` (bit 0 threshold estimate,
 * bit 1 streaming round-1 kernel, bit 2 per-map resolve + fallback) on a workspace that a full call
 * (phases = 7) has filled before.  kb_detect == phases 7.  Round 1 has three kernels with identical output: the packed
 * streaming kernel (pairs of maps as half2; the automatic choice for batches large enough to fill the GPU with long
 * bands), the tiled kernel (everything else) and a full-width fp32 streaming kernel: bit 3 forces the tiled one, bit 4
 * the fp32 streaming one, bit 5 the packed one (tests and A/B measurements; at most one of the three). */
KB_API int kb_detect_phases(const float* score, int B, int H, int W, int nms_dist, int border_dist, float threshold,
                     float min_score, int top_k, float* xyp, int* raster, int* count, int* path, void* ws,
                     size_t ws_bytes, int phases, kb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * simple_nms(scores, nms_radius)                                   models/lightglue.py:904-920
 *
 * The max-pool NMS of the LightGlue-style extractor on B independent [H,W]
 * maps: M = (s == pool(s)); twice { supp = pool(M) > 0; ss = supp ? 0 : s;
 * M |= (ss == pool(ss)) & ~supp }; out = M ? s : 0, pool = (2r+1)^2 maximum with -inf outside the
 * image.  Ties are all kept (the comparison is ==).  0 <= nms_radius <= 16.  `out` may not alias
 * `score`.  The rest of the reference's extract() (border = -1, `> threshold`, torch.topk,
 * sample_descriptors, models/lightglue.py:929-979) is kb_select + kb_sample_desc(normalize=1,
 * coord_mode=1).
 * ------------------------------------------------------------------------------------------- */
KB_API size_t kb_simple_nms_workspace_bytes(int B, int H, int W);
KB_API int kb_simple_nms(const float* score, float* out, int B, int H, int W, int nms_radius, void* ws,
                  size_t ws_bytes, kb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Tensor Lucas-Kanade matcher: OpticalFlow.__call__ / optical_flow_tensor
 *                                                        utils/matcher.py:7-142, 188-203
 *
 * Tracks n keypoints of image 0 into image 1.  img0/img1 [B,C,H,W]; pts0_px and
 * init_px [B,n_max,2] in PIXELS of the full-resolution image (the reference's pts*(W-1,H-1) and its
 * randomly displaced, clamped start, matcher.py:52-61 -- drawing the random start is the caller's job);
 * count [B] (NULL = n_max).  Pyramid level j >= 1 is avg_pool2d(img, kernel 2j, stride 2j) of the
 * original image with coordinates scaled by 2^j; per level `iterations` Gauss-Newton steps on
 * win_size^2*C-sample patches (bilinear, zero padding) with the per-channel Sobel gradients of image 1;
 * a step is skipped where det(G) <= 1e-6, and uses the reference's update
 * delta_b = b_b * sum_i inv(G)[b,i] (matcher.py:139).  out_px [B,n_max,2] in pixels (what
 * optical_flow_tensor returns).  win_size odd; levels <= 8; C*win_size^2 <= 51200.
 * ------------------------------------------------------------------------------------------- */
KB_API size_t kb_lk_workspace_bytes(int B, int C, int H, int W, int levels);
KB_API int kb_lk_track(const float* img0, const float* img1, int B, int C, int H, int W, const float* pts0_px,
                const float* init_px, const int* count, int n_max, int win_size, int levels, int iterations,
                float* out_px, void* ws, size_t ws_bytes, kb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Stage 2 -- descriptor sampling      utils/matcher.py:221-226 and models/lightglue.py:24-41
 *
 * desc [B,C,h,w] (NCHW as the backbones emit it); pts [B,n_max,pts_stride] with (x,y) in the first
 * two columns; count [B] valid rows per map (NULL = n_max everywhere).  Output out [B,n_max,C].
 * coord_mode 0: brute_force_matcher's mapping g=(p-0.5)*2, tap at ((g+1)/2)*(size-1);
 * coord_mode 1: lightglue.sample_descriptors' mapping of PIXEL keypoints with cell size `s`.
 * Bilinear, align_corners=True, zero padding.  normalize != 0 divides by max(||.||_2, 1e-12).
 * ------------------------------------------------------------------------------------------- */
KB_API int kb_sample_desc(const float* desc, int B, int C, int h, int w, const float* pts, int pts_stride,
                   const int* count, int n_max, int normalize, int coord_mode, int s, float* out,
                   kb_stream_t stream);
/* The same sampling (normalize = 0) for the 2 * pairs maps of a batch of image pairs (maps [0, pairs) = image 0, maps
 * [pairs, 2 * pairs) = image 1; desc [2 * pairs, C, h, w], pts / count / out as above) FUSED with the operand
 * preparation of the kb_match_mnn(algo = 1) call that follows on (out[0 : pairs], out[pairs : 2 * pairs], n_max = m_max):
 * besides the float32 rows it writes the 16-bit operand rows and partial row norms into that call's workspace
 * `match_ws` (sized by kb_match_workspace_bytes(pairs, n_max, n_max, C, 1)), and the match call passes phases 6 | 8
 * (bit 3 = operands already in place) to kb_match_mnn_phases.  Same pairs as the two separate calls; one pass over the
 * float32 rows less.  kb_sample_desc_operands_supported: low-resolution maps sampled densely (the plane-staged sampler's
 * case), C a multiple of 64 up to 256, n_max <= 1024; otherwise KB_ERR_UNSUPPORTED -- callers then make the two calls. */
KB_API int kb_sample_desc_operands_supported(int C, int h, int w, int n_max);
KB_API int kb_sample_desc_operands(const float* desc, int pairs, int C, int h, int w, const float* pts, int pts_stride,
                            const int* count, int n_max, int coord_mode, int s, float* out, void* match_ws,
                            size_t match_ws_bytes, kb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Stage 3 -- mutual-nearest-neighbour matching
 *            utils/matcher.py:227-234 -> skimage.feature.match_descriptors (float64 cdist)
 *
 * d0 [B,n_max,D], d1 [B,m_max,D] float32; n0/n1 [B] valid rows (NULL = max).  For every row i of
 * d0 the first-of-ties Euclidean nearest column j (distances evaluated in float64 from the float32
 * inputs); with cross_check the pair survives only if i is also the first-of-ties nearest row of j;
 * pairs with distance >= max_distance are dropped (strict <; pass INFINITY to disable).  Output
 * sorted by i ascending: pairs [B,n_max,2] int32, dist [B,n_max] float64 (may be NULL: the reference's
 * brute_force_matcher keeps only the index pairs, and the tensor-core path then decides the max_distance gate
 * from its certified score, evaluating float64 distances only inside the error band), count [B] (assigned for
 * every b).  `algo` 0 = float64 SIMT evaluation of every distance; 1 = tcgen05 tensor-core
 * candidate search (split-bf16 Gram in TMEM) with float64 certification of the winners (D <= 256);
 * -1 = automatic (1 when supported, else 0).  Both produce the same pairs.
 * ------------------------------------------------------------------------------------------- */
KB_API size_t kb_match_workspace_bytes(int B, int n_max, int m_max, int D, int algo);
KB_API int kb_match_mnn(const float* d0, const float* d1, const int* n0, const int* n1, int B, int n_max,
                 int m_max, int D, double max_distance, int cross_check, int algo, int* pairs,
                 double* dist, int* count, void* ws, size_t ws_bytes, kb_stream_t stream);
/* kb_match_mnn restricted to the parts selected by `phases` (bit 0 operand preparation, bit 1 tensor-core search
 * kernel, bit 2 certification / rescans / gate / pair compaction) on a workspace that a full call (phases = 7) has
 * filled before (a measurement hook); tensor-core path only.  kb_match_mnn == phases 7.  Bit 3 instead of bit 0: the
 * operand rows were written by kb_sample_desc_operands (phases 6 | 8 is the product call after it). */
KB_API int kb_match_mnn_phases(const float* d0, const float* d1, const int* n0, const int* n1, int B, int n_max,
                        int m_max, int D, double max_distance, int cross_check, int algo, int* pairs, double* dist,
                        int* count, void* ws, size_t ws_bytes, int phases, kb_stream_t stream);
/* Diagnostics (tests only): byte offsets inside an algo=1 workspace after kb_match_mnn returned:
 * off[0] records of direction 0: per row FOUR records (one per column slice of the epilogue) of 32 B
 * (float best, second, third, pad; int argbest, argsecond, pad, pad), off[1] same for direction 1, off[2] int32[2] = rows that needed the exact float64 rescan,
 * rows settled by the two-candidate exact check,
 * off[3]/off[4] float32 squared row norms of d0/d1, off[5] int64[512][8] pipeline wait counters of the search
 * kernel (filled with kb_debug_knob(KB_KNOB_TC_DEBUG, 4), scripts/tc_pipeline_profile.py).  `off` has SIX entries. */
KB_API int kb_match_tc_debug_offsets(int B, int n_max, int m_max, int D, size_t* off);

/* ---------------------------------------------------------------------------------------------
 * Stage 4a -- warp_homography                                      utils/projection.py:137-167
 *
 * pts [B,n_max,pts_stride] normalised (x,y); count [B] (NULL = n_max); H33 [B,9] row-major;
 * wh [B,2] = (width, height) of the target image as float.  p = pts*(w-1,h-1); q = H[p,1];
 * uv = q.xy/q.z; valid iff 0<=u<=w-1 and 0<=v<=h-1.  Outputs, compacted in input order:
 * kp_valid [B,n_max,2], kp_warp [B,n_max,2] (both divided back by (w-1,h-1)), ids [B,n_max],
 * ids_out [B,n_max], n_valid [B].
 * ------------------------------------------------------------------------------------------- */
KB_API int kb_warp_homography(const float* pts, int pts_stride, const int* count, int B, int n_max,
                       const float* H33, const float* wh, float* kp_valid, float* kp_warp, int* ids,
                       int* ids_out, int* n_valid, kb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Stage 4a' -- warp_se3 + interpolate_depth                      utils/projection.py:194-267, 270-372
 *
 * Depth-based covisibility (warp(mode='se3'), the MegaDepth-style warp01_params of datasets/megadepth.py:333-352).
 * pts [B,n_max,pts_stride] normalised (x,y); depth0 [B,h0,w0], depth1 [B,h1,w1]; kinv0 [B,9] = inverse of
 * intrinsics0 (the reference inverts it on the fly, projection.py:46), k1 [B,9] = intrinsics1, pose01 [B,16]
 * row-major 4x4, bbox0 / bbox1 [B,2] = (row, col) crop offsets.  A keypoint is scaled by (w0,h0), needs an
 * interpolated depth in view 0 (four corners inside a 10-pixel border, all > 0), is unprojected (+bbox0 +0.5),
 * moved, projected (-bbox1 -0.5) and compared with view 1's interpolated depth (|dz| < 0.05).  Outputs, in input
 * order: kp_valid / kp_warp [B,n_max,2] (divided by (w0,h0) / (w1,h1)), ids [B,n_max], n_valid [B];
 * ids_out [B,n_max] = points projected outside view 1's valid-corner area followed by the occluded ones,
 * n_out [B].  Points without depth in either view appear in neither list (as in the reference).
 * ------------------------------------------------------------------------------------------- */
KB_API int kb_warp_se3(const float* pts, int pts_stride, const int* count, int B, int n_max, const float* depth0,
                int h0, int w0, const float* depth1, int h1, int w1, const float* kinv0, const float* k1,
                const float* pose01, const float* bbox0, const float* bbox1, float* kp_valid, float* kp_warp,
                int* ids, int* ids_out, int* n_valid, int* n_out, kb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Stage 4b -- val_key_points core                              tasks/repeatability.py:39-51, 9-36, 69-85
 *
 * k0c/k01c [B,a_max,2], k1c/k10c [B,b_max,2] (outputs of kb_warp_homography both ways), na/nb [B].
 * dist_mutual = (|k0c_i - k10c_j| + |k1c_j - k01c_i|)/2 with the index diagonal i==j<min(A,B) set
 * to 99999; mutual pairs = entries equal to both their row max and column max of
 * v = (-dist_mutual) - min(-dist_mutual) (float32, as the reference rounds it); scaled by
 * `scale01` (warp01['resize'] or width) and compared with th.  Outputs: stats [B,4] float64 =
 * (gt_num, sum of scaled distances <= th, number of mutual pairs, 0); errors [B,a_max] =
 * scale10 * row minimum (may be NULL); pairs [B,pair_cap,2] unordered mutual pairs (may be NULL).
 * ------------------------------------------------------------------------------------------- */
KB_API size_t kb_repeat_workspace_bytes(int B, int a_max, int b_max);
KB_API int kb_repeat_counts(const float* k0c, const float* k01c, const int* na, const float* k1c,
                     const float* k10c, const int* nb, int B, int a_max, int b_max, float scale01,
                     float scale10, float th, double* stats, float* errors, int* pairs, int pair_cap,
                     void* ws, size_t ws_bytes, kb_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Stage 4c -- MHA corner error                                           tasks/MHA.py:51-72
 *
 * h_est / h_real [B,9] float64 row-major; the four x/y-swapped corners of the reference are
 * projected with both, rescaled by (resize_h/h, resize_w/w), mean L2 -> mean_dist [B] float64,
 * flags [B,n_th] float64 = (mean_dist <= th[t]).  valid [B] int32 (may be NULL): 0 -> all flags 0
 * (the reference returns zeros when RANSAC fails or no covisible keypoints exist).
 * ------------------------------------------------------------------------------------------- */
KB_API int kb_corner_error(const double* h_est, const double* h_real, const int* valid, int B, int w, int h,
                    int resize_h, int resize_w, const double* th, int n_th, double* mean_dist,
                    double* flags, kb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* KB_B200_H */
