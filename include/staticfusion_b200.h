/*
 * staticfusion_b200.h — C ABI of the B200-native StaticFusion solver.
 *
 * This library replaces, for ONE path only, the CPU front-end of
 * raluca-scona/staticfusion: the per-frame coarse-to-fine joint odometry +
 * static/dynamic segmentation solve.  The reference has no FFI / plugin layer;
 * its boundary is the C++ object `class StaticFusion` (StaticFusion.h:66-189)
 * whose callers (StaticFusion-datasets.cpp:171-190) write four input images,
 * call three methods and read three outputs.  Every entry point below cites
 * the reference interface it stands in for.  All citations are relative to
 * the upstream tree.
 *
 * Conventions
 *   - plain pointers and sizes only; caller owns every buffer it passes;
 *   - images are float32, rows x cols; depth in metres with 0 = invalid,
 *     intensity in [0,1] (StaticFusion.h:88-89).  `col_major != 0` means the
 *     buffer is laid out like the reference's Eigen::MatrixXf (column-major);
 *     otherwise row-major (what cv::eigen2cv hands to Reconstruction::fuseFrame,
 *     StaticFusion-datasets.cpp:188-190);
 *   - 4x4 poses are float32[16] COLUMN-major, exactly an Eigen::Matrix4f
 *     (StaticFusion.h:110) so `fuseFrame(..., &T_odometry, ...)` can consume them;
 *   - every function returns 0 on success or a negative SF_E_* code; nothing is
 *     thrown across the ABI.  The solver itself never fails on data: degenerate
 *     inputs set bits in the per-pair status word instead (SF_STATUS_*);
 *   - a context is bound to one CUDA device and is not thread-safe (the
 *     reference object is single-threaded too); use one context per host
 *     thread / GPU.
 *   - there is NO CPU fallback: without a CUDA device sf_create returns
 *     SF_E_CUDA.
 */
#ifndef STATICFUSION_B200_H
#define STATICFUSION_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SF_NUM_CLUSTERS 24 /* NUM_CLUSTERS, StaticFusion.h:61 */

/* error codes */
#define SF_OK 0
#define SF_E_INVALID (-1) /* bad argument / parameter */
#define SF_E_CUDA (-2)    /* CUDA runtime error or no device; see sf_last_error() */
#define SF_E_NOMEM (-3)
#define SF_E_STATE (-4) /* call order violated (e.g. sf_run_solver before inputs were set) */

/* per-pair status bits (0 = nominal) */
#define SF_STATUS_NO_VALID_PIXELS 1 /* validPixels empty at some step: reference divides by zero (FrontEnd.cpp:505-509) */
#define SF_STATUS_ZERO_RESIDUAL 2   /* mean |B| == 0 (identical images): reference produces NaN (FrontEnd.cpp:615) */
#define SF_STATUS_SINGULAR 4        /* a zero pivot was met in the 6x6 normal equations */
#define SF_STATUS_INTERNAL 16       /* the IRLS loop kernel gave up waiting for work that never came (never expected): results of the batch are invalid */

/* memory space of a pointer argument */
#define SF_MEM_HOST 0
#define SF_MEM_DEVICE 1

/* trace record layout (debug / parity tests), identical to the oracle's */
#define SF_TRACE_MAX_IRLS 12
#define SF_TRACE_HDR 96
#define SF_TRACE_IRLS 34
#define SF_TRACE_STEP (SF_TRACE_HDR + SF_TRACE_MAX_IRLS * SF_TRACE_IRLS)

/*
 * Tunables = the public fields the reference's drivers assign after construction
 * (StaticFusion.h:115-146,171-172; values: StaticFusion-datasets.cpp:79-94,121,156-165).
 */
typedef struct sf_params {
    int rows, cols;          /* StaticFusion.h:116 (240 x 320 for res_factor 2) */
    int ctf_levels;          /* StaticFusion.h:120; reference default log2(cols/40)+2 */
    int max_iter_per_level;  /* StaticFusion.h:143 */
    int max_iter_irls;       /* StaticFusion.h:142 */
    int use_motion_filter;   /* StaticFusion.h:139 */
    int enable_segmentation; /* 0 = "everything static" variant (FrontEnd.cpp:606-607): b == 1, no k-means */
    float fovh;              /* StaticFusion.h:115, radians; used for BOTH axes (FrontEnd.cpp:378-386) */
    float k_photometric_res; /* StaticFusion.h:144 */
    float irls_delta_threshold; /* StaticFusion.h:145 */
    float kc_cauchy, kb, kz;    /* StaticFusion.h:146,172 */
    float lambda_reg, lambda_prior; /* StaticFusion.h:171 */
    float previous_speed_const_weight, previous_speed_eig_weight; /* StaticFusion.h:140-141 */
    float outer_exit_threshold; /* 0.04, hard-coded at FrontEnd.cpp:1130; <= 0 disables the exit */
} sf_params;

typedef struct sf_ctx sf_ctx; /* opaque: owns the device arena, stream and per-pair state */

/* Fills *p with the values all three reference drivers use (StaticFusion-datasets.cpp:79-94)
 * for a rows x cols solver; ctf_levels = log2(cols/40)+2 (FrontEnd.cpp:61). */
void sf_default_params(sf_params* p, int rows, int cols);

/* Replaces StaticFusion::StaticFusion(res_factor) (FrontEnd.cpp:52-181) minus GUI / GL backend.
 * max_batch = number of frame pairs solved per sf_solve_batch call.  flags: bit 0 = record trace. */
int sf_create(sf_ctx** out, const sf_params* p, int device, int max_batch, int flags);
void sf_destroy(sf_ctx* ctx);

/* Re-assign the tunables (the drivers rewrite `kb` per frame, StaticFusion-datasets.cpp:156-165).
 * rows, cols, ctf_levels, max_iter_per_level must equal the values given to sf_create; everything else, fovh included
 * (the focal lengths of every level are recomputed), may change.  A call with unchanged values is free: the captured
 * launch schedule is only rebuilt when a value differs. */
int sf_set_params(sf_ctx* ctx, const sf_params* p);

/* ---- drop-in trio: one pair at a time, host buffers, the reference's call order ------------------ */

/* Stand in for writing StaticFusion::depthCurrent / intensityCurrent (StaticFusion.h:88). */
int sf_set_current(sf_ctx* ctx, const float* depth, const float* intensity, int col_major);
/* Stand in for writing StaticFusion::depthPrediction / intensityPrediction (StaticFusion.h:89). */
int sf_set_prediction(sf_ctx* ctx, const float* depth, const float* intensity, int col_major);
/* StaticFusion::twist_odometry_old (StaticFusion.h:111): motion-filter state carried between frames. */
int sf_set_twist_old(sf_ctx* ctx, const float twist_old[6]);
/* StaticFusion::createImagePyramid(bool old_im) (StaticFusion.h:126, FrontEnd.cpp:256). */
int sf_create_image_pyramid(sf_ctx* ctx, int old_im);
/* StaticFusion::runSolver(bool create_image_pyr) (StaticFusion.h:135, FrontEnd.cpp:1071). */
int sf_run_solver(sf_ctx* ctx, int create_image_pyr);
/* StaticFusion::buildSegmImage() (StaticFusion.h:177, SegmentationBackground.cpp:176). */
int sf_build_segm_image(sf_ctx* ctx);
/* ---- depth pre-filter: the step right before the path ------------------------------------------- */
/* Reconstruction::getFilteredDepth(cv::Mat depth, Eigen::MatrixXf& depthMat) (Reconstruction.cpp:722-732), i.e.
 * Shaders/depth_bilateral.frag:30-76 followed by Shaders/depth_metric.frag:28-40: n_images row-major rows x cols
 * uint16 millimetre images -> float metres (0 = invalid: outside [0.3 m, max_depth_m] before or after filtering).
 * depth_mm / depth_out may be host or device memory (SF_MEM_*); col_major_out = 1 writes the Eigen layout (host only).
 * max_depth_m is the reference's depthCutoff (gui default depth_max = 4.5, FrontEnd.cpp:167,174). */
int sf_filter_depth(sf_ctx* ctx, int n_images, const uint16_t* depth_mm, int in_space, float max_depth_m, float* depth_out,
                    int out_space, int col_major_out);

/* ---- image-sequence loader: where depthCurrent / intensityCurrent come from on recorded sequences ---- */
/* StaticFusion::loadImageFromSequenceAssoc(depthFile, rgbFile, res_factor) (StaticFusion.h:122, FrontEnd.cpp:216-254),
 * the part after cv::imread: n_images decoded images at full resolution (rows*res_factor x cols*res_factor, row-major;
 * bgr = 3 bytes per pixel in cv::imread's BGR order, depth_raw = uint16 millimetres) -> rows x cols outputs, vertically
 * flipped and decimated as the reference does (:231,249): intensity in [0,1] (intensityCurrent), depth in metres
 * (depthCurrent), depth_mm (StaticFusion::depth_mm) and the 3-byte colour image (StaticFusion::color_full).  Every output
 * and either input may be NULL.  Buffers are host or device memory (SF_MEM_*); col_major_out = 1 writes the two float
 * images in the Eigen layout (host only).  PNG decoding stays on the host (the reference uses OpenCV for it). */
int sf_convert_frames(sf_ctx* ctx, int n_images, const uint8_t* bgr, const uint16_t* depth_raw, int res_factor, int in_space,
                      float* intensity, float* depth, uint16_t* depth_mm, uint8_t* color_full, int out_space, int col_major_out);

/* ---- 5-frame history: completes the segmentation image exactly as the drivers produce it ---------- */
/* Stand in for the drivers' writes to depthBuffer / intensityBuffer / odomBuffer[slot % 5] (StaticFusion.h:94-96):
 * sf_buffer_set with explicit images (bootstrap, StaticFusion-datasets.cpp:114-116; T = NULL means identity, else
 * column-major 4x4; depth = intensity = NULL writes odomBuffer[slot % 5] only), sf_buffer_push(im_count) with the current frame and T_odometry of the last sf_run_solver
 * (:130-132, :182-184; device-to-device). */
int sf_buffer_set(sf_ctx* ctx, int slot, const float* depth, const float* intensity, const float T[16], int col_major);
int sf_buffer_push(sf_ctx* ctx, int index);
/* StaticFusion::computeResidualsAgainstPreviousImage(int index) (StaticFusion.h:132, FrontEnd.cpp:896): call between
 * sf_run_solver and sf_build_segm_image once im_count >= 5 (StaticFusion-datasets.cpp:175-177).  The result,
 * perClusterAverageResidual (StaticFusion.h:93), persists in the context like the reference's member. */
int sf_compute_residuals_against_previous_image(sf_ctx* ctx, int index);
/* perClusterAverageResidual of every pair of the last solve: n_pairs*24 floats, NaN = no history / empty cluster. */
int sf_get_per_cluster_average_residual(sf_ctx* ctx, float* out);
/* Batched form: when on, sf_solve_sequence / sf_launch after sf_upload_sequence run the history stage for every pair
 * k >= 4 of the sequence (frame k+1 against frame k-4 through the increments of pairs k-4..k) before the segmentation
 * image is built; pairs 0..3 have no history (NaN), so callers that split a sequence overlap the pieces by 4 pairs. */
int sf_set_history(sf_ctx* ctx, int on);

/* Outputs the drivers read afterwards.  Any pointer may be NULL.
 *   T_odometry      StaticFusion.h:110 (column-major 4x4)
 *   twist_old_out   StaticFusion.h:111 after FrontEnd.cpp:1143-1144
 *   b_segm          StaticFusion.h:166
 *   b_perpixel      StaticFusion.h:167 rows x cols (layout per col_major)
 *   labels          StaticFusion.h:155 clusterAllocation[0], int32 rows x cols, 24 = no depth
 */
int sf_get_outputs(sf_ctx* ctx, float T_odometry[16], float twist_old_out[6], float b_segm[SF_NUM_CLUSTERS],
                   float* b_perpixel, int32_t* labels, int col_major, int* irls_iterations, int* status);

/* ---- batched path: n_pairs independent (prediction, current) pairs per call ---------------------- */

/*
 * Solve n_pairs <= max_batch independent pairs = createImagePyramid(true) + runSolver(true) +
 * buildSegmImage() on each (StaticFusion-datasets.cpp:171-180).  Images are row-major, pair-major:
 * image k of a stack starts at offset k*rows*cols.
 *   depth_cur/inten_cur/depth_pred/inten_pred : n_pairs images each (in_space = SF_MEM_HOST | SF_MEM_DEVICE)
 *   twist_old_in  : n_pairs*6 floats or NULL (= zeros), always host
 * Outputs (out_space applies to b_perpixel and labels_u8 only; the small arrays are always host), any may be NULL:
 *   T_odometry    : n_pairs*16, column-major 4x4 each
 *   twist_old_out : n_pairs*6
 *   b_segm        : n_pairs*24
 *   b_perpixel    : n_pairs*rows*cols float, row-major
 *   labels_u8     : n_pairs*rows*cols uint8 cluster labels (24 = no depth)
 *   irls_iters    : n_pairs, total IRLS iterations executed over all levels
 *   status        : n_pairs, SF_STATUS_* bits
 * The call is synchronous: outputs are valid on return.
 */
int sf_solve_batch(sf_ctx* ctx, int n_pairs, const float* depth_cur, const float* inten_cur,
                   const float* depth_pred, const float* inten_pred, int in_space, const float* twist_old_in,
                   float* T_odometry, float* twist_old_out, float* b_segm, float* b_perpixel, uint8_t* labels_u8,
                   int out_space, int* irls_iters, int* status);

/*
 * Frame-to-frame odometry over a sequence (the bootstrap form, StaticFusion-datasets.cpp:109-144:
 * prediction := previous raw frame).  n_frames images -> n_frames-1 pairs (pair k = frames k, k+1);
 * each frame's pyramid is built once and shared by the two pairs that use it.  n_frames-1 <= max_batch.
 * Output layout as in sf_solve_batch with n_pairs = n_frames-1.
 */
int sf_solve_sequence(sf_ctx* ctx, int n_frames, const float* depth, const float* inten, int in_space,
                      const float* twist_old_in, float* T_odometry, float* twist_old_out, float* b_segm,
                      float* b_perpixel, uint8_t* labels_u8, int out_space, int* irls_iters, int* status);

/* ---- split-phase variants used by the benchmark to time the device-resident region -------------- */
/* Upload only (frames land in the level-0 slots of the pyramids); in_space as above. */
int sf_upload_pairs(sf_ctx* ctx, int n_pairs, const float* depth_cur, const float* inten_cur,
                    const float* depth_pred, const float* inten_pred, int in_space, const float* twist_old_in);
int sf_upload_sequence(sf_ctx* ctx, int n_frames, const float* depth, const float* inten, int in_space,
                       const float* twist_old_in);
/* Enqueue the whole solve for the uploaded batch on the context's stream; no host synchronisation. */
/* sf_upload_sequence for frames still in their decoded file form (see sf_convert_frames): the conversion runs on the
 * device straight into the solver's frame slots, 5 * res_factor^2 bytes per pixel cross PCIe instead of 8. */
int sf_upload_sequence_raw(sf_ctx* ctx, int n_frames, const uint8_t* bgr, const uint16_t* depth_raw, int res_factor, int in_space,
                           const float* twist_old_in);
int sf_launch(sf_ctx* ctx);
/* Wait for the stream. */
int sf_sync(sf_ctx* ctx);
/* Copy results of the last sf_launch back (synchronises). */
int sf_download(sf_ctx* ctx, float* T_odometry, float* twist_old_out, float* b_segm, float* b_perpixel,
                uint8_t* labels_u8, int out_space, int* irls_iters, int* status);
/* Same for pairs [first_pair, first_pair + n) of the last solve (callers that overlap pieces of a sequence drop the
 * overlap this way); per_cluster_residual: n*24 floats or NULL (always host). */
int sf_download_range(sf_ctx* ctx, int first_pair, int n, float* T_odometry, float* twist_old_out, float* b_segm,
                      float* b_perpixel, uint8_t* labels_u8, int out_space, int* irls_iters, int* status,
                      float* per_cluster_residual);
/* Split-phase form: _begin enqueues the device->host copies behind the solve and returns at once (give it page-locked
 * buffers), _end waits for them and fills the small per-pair arrays passed to _begin.  One download in flight per context;
 * the next sf_upload_* / sf_launch on the context may be issued only after _end. */
int sf_download_range_begin(sf_ctx* ctx, int first_pair, int n, float* T_odometry, float* twist_old_out, float* b_segm,
                            float* b_perpixel, uint8_t* labels_u8, int out_space, int* irls_iters, int* status,
                            float* per_cluster_residual);
int sf_download_range_end(sf_ctx* ctx);
/* cudaStream_t of the context as an integer handle (for CUDA-event timing by the caller). */
uint64_t sf_stream(sf_ctx* ctx);
/* Number of kernel launches enqueued by the last sf_launch. */
int sf_last_launch_count(sf_ctx* ctx);
/* Number of concurrent pair ranges (streams) the schedule of the current batch is cut into: 1 for small batches, up to 3
 * for large ones (environment variable SF_LANES overrides).  The result does not depend on it. */
int sf_last_lane_count(sf_ctx* ctx);

/* ---- measurement hooks (bench.py) ---------------------------------------------------------------- */
#define SF_PROF_CLASSES 9 /* 0 init, 1 pyramid, 2 clustering, 3 warp, 4 linearise, 5 irls (the loop / fused kernels: whole IRLS loops; pass 1 with
                             SF_IRLS_LOOP=0), 6 irls_pass2 (SF_IRLS_LOOP=0 only), 7 pose_update, 8 finish.  SF_NVTX=1 adds NVTX ranges per stage. */
#define SF_PROF_LEVELS 8
/* Record a CUDA-event pair around every kernel group of the following sf_launch calls (on the context's stream). */
int sf_profile_enable(sf_ctx* ctx, int on);
/* After sf_sync: summed device time [ms] and launch counts per (class, pyramid level) of the last sf_launch;
 * both arrays hold SF_PROF_CLASSES*SF_PROF_LEVELS entries, index class*SF_PROF_LEVELS + image_level. */
int sf_profile_read(sf_ctx* ctx, float* ms, int* launches);
/* The same event pairs one by one, in launch order: up to `capacity` records of (class, image level, milliseconds);
 * *n_records receives the number of records of the last sf_launch (may exceed capacity). */
int sf_profile_read_records(sf_ctx* ctx, int capacity, int* cls, int* level, float* ms, int* n_records);
/* Per pair and step (ctf_levels*max_iter_per_level steps, index level*max_iter_per_level + k) of the last solve:
 * valid pixels N (0 = step not executed) and IRLS iterations run.  Arrays hold n_pairs*steps ints. */
int sf_get_step_stats(sf_ctx* ctx, int* n_valid, int* irls_iters);
/* Lloyd iterations kMeans3DCoord ran for every pair of the last solve (1..9, KMeans.cpp:167-228; 0 without segmentation). */
int sf_get_kmeans_iterations(sf_ctx* ctx, int* iterations);
/* on != 0: the host->device copies of sf_upload_sequence_raw (host input) and the device->host copies of
 * sf_download_range_begin run on two extra streams of the context, ordered against the solve stream by events: a
 * context's next upload overlaps its current solve, its download overlaps the solves of other contexts, and no call
 * waits on the host (sf_download_range_end does).  Results are unchanged.  Replaces the blocking glDownload / upload
 * round trips around the solver in the drivers (StaticFusion-datasets.cpp:166-190). */
int sf_set_copy_streams(sf_ctx* ctx, int on);
/* The result rows of the last solve where they lie in DEVICE memory: n_rows rows of row_floats 4-byte words each --
 * T_odometry (16, row-major), twist_odometry_old (6), b_segm (24), then two int32 (IRLS iterations, status bits).  Valid
 * until the next upload / launch on this context; ordered after the solve on sf_stream(ctx).  For collectives that gather
 * the per-frame poses across GPUs without a host round trip (StaticFusion-datasets.cpp:186-196 reads them per frame). */
int sf_result_rows_device(sf_ctx* ctx, const float** rows, int* n_rows, int* row_floats);

/* ---- introspection for parity tests ------------------------------------------------------------ */
/* Halt the next solves right after the linearisation of step (level*max_iter_per_level + k); -1 = run to the end. */
int sf_debug_set_stop_step(sf_ctx* ctx, int stop_step);
/* Copy a named float plane of pair `pair` at pyramid level `image_level` to host (row-major).
 * names: depth, intensity, depth_pred, intensity_pred, depth_warped, intensity_warped,
 *        depth_inter, xx_inter, yy_inter, dcu, dcv, dct, ddu, ddv, ddt, weights_c, weights_d, null;
 *        depth_warped_ref, intensity_warped_ref (level 0, after the history stage) */
int sf_debug_get_plane(sf_ctx* ctx, const char* name, int pair, int image_level, float* out);
/* Cluster labels of a pyramid level as int32 (24 = no depth), row-major. */
int sf_debug_get_labels(sf_ctx* ctx, int pair, int image_level, int32_t* out);
/* k-means centres [3][24] (z,x,y rows) and 24x24 connectivity of a pair. */
int sf_debug_get_kmeans(sf_ctx* ctx, int pair, float centres[3 * SF_NUM_CLUSTERS],
                        uint8_t connectivity[SF_NUM_CLUSTERS * SF_NUM_CLUSTERS]);
/* Trace of a pair: ctf_levels*max_iter_per_level records of SF_TRACE_STEP floats (needs flags bit 0). */
int sf_debug_get_trace(sf_ctx* ctx, int pair, float* out, int n_floats);

/* Last CUDA / argument error message of the calling thread's most recent failing call. */
const char* sf_last_error(void);
/* ABI version of this header. */
int sf_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* STATICFUSION_B200_H */
