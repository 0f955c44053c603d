/*
 * sf_oracle.h — C ABI of the CPU oracle (TEST INFRASTRUCTURE, not product).
 *
 * The oracle is a dependency-free CPU restatement of StaticFusion's joint
 * odometry + static/dynamic segmentation solver, following (file:line are
 * relative to the upstream reference tree):
 *     FrontEnd.cpp:256-892, 1071-1146      pyramid, warp, linearisation, IRLS, pose update
 *     SegmentationBackground.cpp:53-197    seg prior, 24x24 seg solve, per-pixel image
 *     KMeans.cpp:52-391                    24-means clustering, connectivity, label pyramid
 *
 * PARITY PINNED AGAINST THE REFERENCE'S OWN CODE: the reference ships no tests or
 * golden vectors, but its solver sources (KMeans.cpp, SegmentationBackground.cpp,
 * FrontEnd.cpp:256-1146, StaticFusion.h) are compiled unmodified from where they
 * lie against a header shim (oracle/ref_shim, `make ref` -> oracle/_ref/) and the
 * oracle's reference-literal policy (ORC_ACCUM_F32) reproduces them BIT FOR BIT,
 * stage by stage and end to end (tests/test_oracle_vs_reference.py; fixtures
 * generated from that build in tests/golden/reference_golden_*.npz).
 * What is not pinned: Eigen's / MRPT's internal arithmetic (not in the reference
 * tree, not installed): both sides use documented stand-ins for it.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product
 * (staticfusion_b200/) never links, imports or calls it.
 */
#ifndef SF_ORACLE_H
#define SF_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NUM_CLUSTERS 24
#define ORC_NUM_STAGES 8 /* pyramid, k-means, warp, linearise, IRLS, seg-solve, pose update, per-pixel image */
#define ORC_TRACE_MAX_IRLS 12
#define ORC_TRACE_HDR 96
#define ORC_TRACE_IRLS 34
#define ORC_TRACE_STEP (ORC_TRACE_HDR + ORC_TRACE_MAX_IRLS * ORC_TRACE_IRLS)

/* accumulation policy for cross-pixel sums and the small dense algebra */
#define ORC_ACCUM_F32 0   /* reference-literal: sequential float sums in the reference's traversal order, float algebra */
#define ORC_ACCUM_EXACT 1 /* every cross-pixel sum is an order-independent integer sum, double algebra: the CUDA path's contract */
#define ORC_ACCUM_F64 2   /* like EXACT but plain double sums for the normal equations / |res|^2: measures EXACT's quantisation */

typedef struct {
    int rows, cols;          /* finest level (StaticFusion.h:116) */
    int ctf_levels;          /* StaticFusion.h:120 */
    int max_iter_per_level;  /* StaticFusion.h:143 */
    int max_iter_irls;       /* StaticFusion.h:142 */
    int use_motion_filter;   /* StaticFusion.h:139 */
    int enable_segmentation; /* 0: b_segm == 1 everywhere ("Everything static", FrontEnd.cpp:606-607), k-means skipped */
    float fovh;              /* StaticFusion.h:115 (radians) */
    float k_photometric_res; /* StaticFusion.h:144 */
    float irls_delta_threshold;
    float kc_cauchy, kb, kz;
    float lambda_reg, lambda_prior;
    float previous_speed_const_weight, previous_speed_eig_weight;
    float outer_exit_threshold; /* 0.04 hard-coded at FrontEnd.cpp:1130; <=0 disables the exit */
} orc_params;

typedef struct orc_ctx orc_ctx;

orc_ctx* orc_create(const orc_params* p, int accum_mode);
void orc_destroy(orc_ctx* c);
void orc_set_params(orc_ctx* c, const orc_params* p);

/* inputs: row-major rows x cols float; depth in metres (0 = invalid), intensity in [0,1] */
void orc_set_current(orc_ctx* c, const float* depth, const float* intensity);
void orc_set_prediction(orc_ctx* c, const float* depth, const float* intensity);
void orc_set_twist_old(orc_ctx* c, const float twist_old[6]);

/* the reference's three entry points (StaticFusion.h:126,135,177) */
void orc_create_image_pyramid(orc_ctx* c, int old_im);
/* stop_step >= 0: return right after computeSegPrior of step (level*max_iter_per_level + k) for stage dumps */
void orc_run_solver(orc_ctx* c, int create_image_pyr, int stop_step);
void orc_build_segm_image(orc_ctx* c);
/* wall time accumulated per stage since creation / the last reset (std::chrono::steady_clock; CPU baseline) */
void orc_get_stage_seconds(orc_ctx* c, double out[ORC_NUM_STAGES], int reset);

/* 5-frame history (StaticFusion.h:92-99; FrontEnd.cpp:896-1069).  The drivers copy the current frame and T_odometry
 * into slot im_count % 5 after every frame (StaticFusion-datasets.cpp:114-116, 130-132, 182-184) and call
 * computeResidualsAgainstPreviousImage(im_count) between runSolver and buildSegmImage once im_count >= 5 (:175-177). */
void orc_buffer_set(orc_ctx* c, int slot, const float* depth, const float* intensity, const float T_rowmajor[16]);
void orc_buffer_push(orc_ctx* c, int index);
void orc_compute_residuals_against_previous_image(orc_ctx* c, int index);
void orc_get_per_cluster_average_residual(const orc_ctx* c, float out[ORC_NUM_CLUSTERS]);
void orc_set_per_cluster_average_residual(orc_ctx* c, const float in[ORC_NUM_CLUSTERS]);
void orc_set_T(orc_ctx* c, const float T_rowmajor[16]);
/* names: depth_warped_ref, intensity_warped_ref, cumulative (full resolution, row-major) */
int  orc_get_residual_image(const orc_ctx* c, const char* name, float* out_rowmajor);

/* Depth pre-filter (Shaders/depth_bilateral.frag + depth_metric.frag via Reconstruction::getFilteredDepth,
 * Reconstruction.cpp:722-732): row-major u16 millimetres -> row-major float metres.  exact: 0 = libm expf, 1 = the
 * reproducible exp of the CUDA contract.  Parity unpinned (GLSL on the reference side). */
void orc_filter_depth(const uint16_t* depth_mm, int rows, int cols, float max_depth_m, int exact, float* out);
float orc_det_expf(float a);

/* Image-sequence loader, conversion half (StaticFusion::loadImageFromSequenceAssoc, FrontEnd.cpp:216-254; the PNG decode
 * itself is OpenCV/libpng).  bgr: full_rows x full_cols x 3 bytes as cv::imread(..., CV_LOAD_IMAGE_COLOR) returns them;
 * depth_raw: full_rows x full_cols uint16 (millimetres).  Outputs are (full_rows/res_factor) x (full_cols/res_factor),
 * row-major, vertically flipped and decimated exactly as the reference does: intensity [0,1], depth metres,
 * depth_mm (StaticFusion::depth_mm) and colour (StaticFusion::color_full, 3 bytes per pixel).  Pinned bit-for-bit
 * against the reference's own function (oracle/_ref, tests/test_tum_io.py). */
void orc_convert_frame(const uint8_t* bgr, const uint16_t* depth_raw, int full_rows, int full_cols, int res_factor,
                       float* intensity, float* depth, uint16_t* depth_mm, uint8_t* color);
/* Trajectory bookkeeping of the callers: currPose = currPose * T_odometry in float (Reconstruction.cpp:256,265), the
 * TUM-format quaternion (Eigen::Quaternionf(Matrix3f), Eigen/src/Geometry/Quaternion.h: trace / largest-diagonal branches)
 * of Datasets::writeTrajectoryFile (Datasets.cpp:252-266) and Reconstruction::savePly (Reconstruction.cpp:460-484).
 * Matrices row-major; q = (x, y, z, w). */
void orc_pose_compose(const float A_rowmajor[16], const float B_rowmajor[16], float out_rowmajor[16]);
void orc_quat_from_rotation(const float T_rowmajor[16], float q_xyzw[4]);

/* stand-alone stages for unit tests */
void orc_kmeans(orc_ctx* c); /* kMeans3DCoord + createClustersPyramidUsingKMeans on the current pyramid */
void orc_warp_level(orc_ctx* c, int image_level, const float T_odometry_rowmajor[16]);

/* outputs */
void orc_get_T(const orc_ctx* c, float out_rowmajor[16]);
void orc_get_twists(const orc_ctx* c, float twist_odometry[6], float twist_old[6], float twist_level[6]);
void orc_get_b_segm(const orc_ctx* c, float out[ORC_NUM_CLUSTERS]);
void orc_get_b_perpixel(const orc_ctx* c, float* out_rowmajor);
int  orc_get_labels(const orc_ctx* c, int image_level, int32_t* out_rowmajor);
void orc_get_kmeans(const orc_ctx* c, float out[3 * ORC_NUM_CLUSTERS]);      /* [3][24]: z,x,y rows */
void orc_get_connectivity(const orc_ctx* c, uint8_t out[ORC_NUM_CLUSTERS * ORC_NUM_CLUSTERS]);
int  orc_get_status(const orc_ctx* c);
int  orc_get_total_irls(const orc_ctx* c);
/* named float image of one pyramid level, row-major; returns 0 on success.
 * names: depth, intensity, xx, yy, depth_pred, intensity_pred, depth_warped, intensity_warped,
 *        depth_inter, intensity_inter, xx_inter, yy_inter, dcu, dcv, dct, ddu, ddv, ddt,
 *        weights_c, weights_d, null */
int  orc_get_image(const orc_ctx* c, const char* name, int image_level, float* out_rowmajor);
/* trace: ctf_levels*max_iter_per_level records of ORC_TRACE_STEP floats */
int  orc_trace_size(const orc_ctx* c);
void orc_get_trace(const orc_ctx* c, float* out);

/* small dense algebra exposed for the numpy-twin tests (double versions) */
void orc_se3_exp(const double xi[6], double T_rowmajor[16]);
void orc_se3_log(const double T_rowmajor[16], double xi[6]);
int  orc_ldlt_solve(int n, const double* A_rowmajor, const double* b, double* x);
void orc_jacobi_eig6(const double A_rowmajor[36], double evals[6], double evecs_rowmajor[36]);

#ifdef __cplusplus
}
#endif
#endif
