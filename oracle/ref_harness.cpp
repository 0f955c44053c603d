/*
 * ref_harness.cpp — drives the REFERENCE's own solver sources (compiled unmodified against oracle/ref_shim)
 * through a small C ABI so that tests can compare them with the oracle stage by stage.  TEST INFRASTRUCTURE.
 *
 * The only reference code restated here is the buffer-allocation part of StaticFusion::StaticFusion
 * (FrontEnd.cpp:52-170): the original constructor also opens a Pangolin window and builds the GL
 * Reconstruction (FrontEnd.cpp:171-180), which cannot exist in this environment.
 */
#include <StaticFusion.h>

using namespace Eigen;

/* FrontEnd.cpp:52-170 without the GUI / GL back-end */
StaticFusion::StaticFusion(unsigned int res_factor)
{
    rows = 480/res_factor;  cols = 640/res_factor;                      /* :55-56 */
    fovh = M_PI*62.5/180.0; fovv = M_PI*48.5/180.0;                     /* :57-58 */
    width = 640/res_factor; height = 480/res_factor;                    /* :59-60 */
    ctf_levels = log2(cols/40) + 2;                                     /* :61 */
    k_photometric_res = 0.15f; irls_delta_threshold = 1e-6f; max_iter_irls = 10; max_iter_per_level = 2;   /* :66-69 */
    previous_speed_const_weight = 0.05f; previous_speed_eig_weight = 0.5f; kc_Cauchy = 0.5f; kb = 1.25f; kz = 1.5f;  /* :70-74 */
    use_motion_filter = false;                                          /* :76 */
    cam_pose.setFromValues(0,0,0,0,0,0); cam_oldpose = cam_pose; twist_odometry_old.fill(0.f);  /* :79-81 */
    depthCurrent.setSize(height,width); depthPrediction.setSize(height,width);                   /* :84-87 */
    intensityCurrent.setSize(height,width); intensityPrediction.setSize(height,width);
    dct.resize(rows,cols); ddt.resize(rows,cols); dcu.resize(rows,cols); ddu.resize(rows,cols);  /* :89-94 */
    dcv.resize(rows,cols); ddv.resize(rows,cols); Null.resize(rows,cols);
    weights_c.setSize(rows,cols); weights_d.setSize(rows,cols);
    intensityBuffer.resize(bufferLength); depthBuffer.resize(bufferLength); odomBuffer.resize(bufferLength);  /* :96-103 */
    for (int i=0; i<bufferLength; i++) { intensityBuffer[i].resize(rows, cols); depthBuffer[i].resize(rows, cols); }
    perClusterAverageResidual.fill(std::numeric_limits<float>::quiet_NaN());                     /* :105 */
    const unsigned int pyr_levels = round(log2(width/cols)) + ctf_levels;                        /* :108-143, sized for 8 levels so ctf_levels may be re-assigned */
    const unsigned int alloc_levels = std::max(pyr_levels, 8u);
    intensityPyr.resize(alloc_levels); intensityPredPyr.resize(alloc_levels); intensityInterPyr.resize(alloc_levels);
    depthPyr.resize(alloc_levels); depthPredPyr.resize(alloc_levels); depthInterPyr.resize(alloc_levels);
    xxPyr.resize(alloc_levels); xxInterPyr.resize(alloc_levels); xxPredPyr.resize(alloc_levels);
    yyPyr.resize(alloc_levels); yyInterPyr.resize(alloc_levels); yyPredPyr.resize(alloc_levels);
    intensityWarpedPyr.resize(alloc_levels); depthWarpedPyr.resize(alloc_levels);
    xxWarpedPyr.resize(alloc_levels); yyWarpedPyr.resize(alloc_levels); clusterAllocation.resize(alloc_levels);
    xxBuffer.setSize(height, width); yyBuffer.setSize(height, width); xxBuffer.assign(0.f); yyBuffer.assign(0.f);
    for (unsigned int i = 0; i<alloc_levels; i++)
    {
        const unsigned int s = pow(2.f,int(i));
        cols_i = width/s; rows_i = height/s;
        if (rows_i < 1 || cols_i < 1) break;
        intensityPyr[i].resize(rows_i, cols_i); intensityPredPyr[i].resize(rows_i, cols_i); intensityInterPyr[i].resize(rows_i, cols_i);
        depthPyr[i].resize(rows_i, cols_i); depthInterPyr[i].resize(rows_i, cols_i); depthPredPyr[i].resize(rows_i, cols_i);
        depthPyr[i].assign(0.f); depthPredPyr[i].assign(0.f);
        xxPyr[i].resize(rows_i, cols_i); xxInterPyr[i].resize(rows_i, cols_i); xxPredPyr[i].resize(rows_i, cols_i);
        xxPyr[i].assign(0.f); xxPredPyr[i].assign(0.f);
        yyPyr[i].resize(rows_i, cols_i); yyInterPyr[i].resize(rows_i, cols_i); yyPredPyr[i].resize(rows_i, cols_i);
        yyPyr[i].assign(0.f); yyPredPyr[i].assign(0.f);
        intensityWarpedPyr[i].resize(rows_i,cols_i); depthWarpedPyr[i].resize(rows_i,cols_i);
        xxWarpedPyr[i].resize(rows_i,cols_i); yyWarpedPyr[i].resize(rows_i,cols_i);
        clusterAllocation[i].resize(rows_i, cols_i); clusterAllocation[i].assign(0);
    }
    const Vector4f v_mask(1.f, 2.f, 2.f, 1.f);                          /* :146-149 */
    for (unsigned int i=0; i<4; i++)
        for (unsigned int j=0; j<4; j++)
            convMask(i,j) = v_mask(i)*v_mask(j)/36.f;
    b_segm_perpixel.setSize(rows,cols); b_segm_perpixel.fill(0.5f); b_segm.fill(0.5f);           /* :154-156 */
    b_prior.fill(0.f); lambda_t_w.fill(0.f);
    depthWarpedRefference = MatrixXf::Zero(rows, cols); intensityWarpedRefference = MatrixXf::Zero(rows, cols);  /* :158-159 */
    T_odometry.setIdentity(); twist_odometry.fill(0.f); twist_level_odometry.fill(0.f); est_cov.fill(0.f);
    for (int i = 0; i < NUM_CLUSTERS; i++) for (int j = 0; j < NUM_CLUSTERS; j++) connectivity[i][j] = (i == j);
    depth_mm = cv::Mat(height, width, CV_16U, 0.0);                                             /* :161-162 */
    color_full = cv::Mat(height, width, CV_8UC3, cv::Scalar(0,0,0));
    confidence = 0.25f; depth_max = 4.5f; reconstruction = nullptr; gui = nullptr;               /* :167-168 */
}

/* ---- C ABI -------------------------------------------------------------------------------------------- */
static void in_rowmajor(MatrixXf& m, const float* src) {
    for (int v = 0; v < m.rows(); v++) for (int u = 0; u < m.cols(); u++) m(v, u) = src[(size_t)v * m.cols() + u];
}
static void out_rowmajor(const Dense<float>& m, int rows, int cols, float* dst) {
    for (int v = 0; v < rows; v++) for (int u = 0; u < cols; u++) dst[(size_t)v * cols + u] = m(v, u);
}

extern "C" {

typedef struct {
    int ctf_levels, max_iter_per_level, max_iter_irls, use_motion_filter;
    float k_photometric_res, irls_delta_threshold, kc_cauchy, kb, kz, lambda_reg, lambda_prior;
    float previous_speed_const_weight, previous_speed_eig_weight;
} ref_params;

#ifdef SF_B200_BINDING  /* the library built from ref_binding.cpp: the class's solver methods forward to libstaticfusion_b200.so */
void ref_binding_release(void* h);
int ref_is_b200_binding(void) { return 1; }
#else
static void ref_binding_release(void*) {}
int ref_is_b200_binding(void) { return 0; }
#endif
void* ref_create(int res_factor) { return new StaticFusion((unsigned)res_factor); }
void ref_destroy(void* h) { ref_binding_release(h); delete static_cast<StaticFusion*>(h); }
int ref_rows(void* h) { return (int)static_cast<StaticFusion*>(h)->rows; }
int ref_cols(void* h) { return (int)static_cast<StaticFusion*>(h)->cols; }
int ref_default_levels(void* h) { return (int)static_cast<StaticFusion*>(h)->ctf_levels; }

/* the drivers assign these public fields after construction (StaticFusion-datasets.cpp:79-94) */
void ref_set_params(void* h, const ref_params* p) {
    StaticFusion& s = *static_cast<StaticFusion*>(h);
    s.ctf_levels = p->ctf_levels; s.max_iter_per_level = p->max_iter_per_level; s.max_iter_irls = p->max_iter_irls;
    s.use_motion_filter = p->use_motion_filter != 0; s.k_photometric_res = p->k_photometric_res;
    s.irls_delta_threshold = p->irls_delta_threshold; s.kc_Cauchy = p->kc_cauchy; s.kb = p->kb; s.kz = p->kz;
    s.lambda_reg = p->lambda_reg; s.lambda_prior = p->lambda_prior;
    s.previous_speed_const_weight = p->previous_speed_const_weight; s.previous_speed_eig_weight = p->previous_speed_eig_weight;
}
void ref_set_current(void* h, const float* d, const float* i) { StaticFusion& s = *static_cast<StaticFusion*>(h); in_rowmajor(s.depthCurrent, d); in_rowmajor(s.intensityCurrent, i); }
void ref_set_prediction(void* h, const float* d, const float* i) { StaticFusion& s = *static_cast<StaticFusion*>(h); in_rowmajor(s.depthPrediction, d); in_rowmajor(s.intensityPrediction, i); }
void ref_set_twist_old(void* h, const float* t) { StaticFusion& s = *static_cast<StaticFusion*>(h); for (int k = 0; k < 6; k++) s.twist_odometry_old(k) = t[k]; }
void ref_set_T(void* h, const float* T_rowmajor) { StaticFusion& s = *static_cast<StaticFusion*>(h); for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) s.T_odometry(r, c) = T_rowmajor[r * 4 + c]; }

/* the reference's own methods */
void ref_create_image_pyramid(void* h, int old_im) { static_cast<StaticFusion*>(h)->createImagePyramid(old_im != 0); }
void ref_run_solver(void* h, int create_image_pyr) { static_cast<StaticFusion*>(h)->runSolver(create_image_pyr != 0); }
void ref_build_segm_image(void* h) { static_cast<StaticFusion*>(h)->buildSegmImage(); }
/* the drivers' ring-buffer writes (StaticFusion-datasets.cpp:114-116, 182-184), then the reference's own method */
void ref_buffer_set(void* h, int slot, const float* d, const float* i, const float* T_rowmajor) {
    StaticFusion& s = *static_cast<StaticFusion*>(h);
    const int b = ((slot % s.bufferLength) + s.bufferLength) % s.bufferLength;
    in_rowmajor(s.depthBuffer[b], d); in_rowmajor(s.intensityBuffer[b], i);
    Eigen::Matrix4f T; for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) T(r, c) = T_rowmajor[r * 4 + c];
    s.odomBuffer[b] = T;
}
void ref_buffer_push(void* h, int index) {
    StaticFusion& s = *static_cast<StaticFusion*>(h);
    s.depthBuffer[index % s.bufferLength] = s.depthCurrent.replicate(1,1);
    s.intensityBuffer[index % s.bufferLength] = s.intensityCurrent.replicate(1,1);
    s.odomBuffer[index % s.bufferLength] = s.T_odometry;
}
void ref_compute_residuals_against_previous_image(void* h, int index) { static_cast<StaticFusion*>(h)->computeResidualsAgainstPreviousImage(index); }
void ref_get_per_cluster_average_residual(void* h, float* out) { StaticFusion& s = *static_cast<StaticFusion*>(h); for (int l = 0; l < NUM_CLUSTERS; l++) out[l] = s.perClusterAverageResidual(l); }
int ref_get_residual_image(void* h, const char* name, float* out) {
    StaticFusion& s = *static_cast<StaticFusion*>(h);
    const std::string n(name);
    if (n == "depth_warped_ref") out_rowmajor(s.depthWarpedRefference, s.rows, s.cols, out);
    else if (n == "intensity_warped_ref") out_rowmajor(s.intensityWarpedRefference, s.rows, s.cols, out);
    else if (n == "cumulative") out_rowmajor(s.cumulativeResiduals, s.rows, s.cols, out);
    else return -2;
    return 0;
}
/* ---- image-sequence loader (FrontEnd.cpp:183-254): the reference's own loadAssoc / loadImageFromSequenceAssoc ---- */
/* register a decoded image under a file name for the shim's cv::imread: type 16 = 8-bit BGR, 2 = 16-bit single channel */
void ref_register_image(const char* name, const void* data, int rows, int cols, int type) {
    cv::Mat m(rows, cols, type, 0.0);
    std::memcpy(m.data, data, (size_t)rows * cols * cv::Mat::elem(type));
    cv::shim_image_registry()[name] = m;
}
void ref_clear_images(void) { cv::shim_image_registry().clear(); }
int ref_load_image_from_sequence_assoc(void* h, const char* depthFile, const char* rgbFile, int res_factor) {
    return static_cast<StaticFusion*>(h)->loadImageFromSequenceAssoc(depthFile, rgbFile, (unsigned)res_factor) ? 1 : 0;
}
void ref_get_current(void* h, float* depth, float* intensity, unsigned short* depth_mm, unsigned char* color_full) {
    StaticFusion& s = *static_cast<StaticFusion*>(h);
    out_rowmajor(s.depthCurrent, s.rows, s.cols, depth); out_rowmajor(s.intensityCurrent, s.rows, s.cols, intensity);
    std::memcpy(depth_mm, s.depth_mm.data, sizeof(unsigned short) * s.rows * s.cols);
    std::memcpy(color_full, s.color_full.data, 3u * s.rows * s.cols);
}
/* loadAssoc: returns the number of entries (or -1 when the file cannot be opened); entry k is read back with ref_assoc_entry */
static std::vector<double> g_assoc_ts; static std::vector<std::string> g_assoc_depth, g_assoc_color;
int ref_load_assoc(void* h, const char* dir, const char* assocFile) {
    g_assoc_ts.clear(); g_assoc_depth.clear(); g_assoc_color.clear();
    if (!static_cast<StaticFusion*>(h)->loadAssoc(dir, assocFile, g_assoc_ts, g_assoc_depth, g_assoc_color)) return -1;
    return (int)g_assoc_ts.size();
}
double ref_assoc_entry(int k, char* depth_path, char* color_path, int cap) {
    std::snprintf(depth_path, cap, "%s", g_assoc_depth[k].c_str()); std::snprintf(color_path, cap, "%s", g_assoc_color[k].c_str());
    return g_assoc_ts[k];
}

#ifndef SF_B200_BINDING  /* stage-by-stage entry points into the reference's own solver code (absent from the bound library) */
void ref_kmeans(void* h) { StaticFusion& s = *static_cast<StaticFusion*>(h); s.kMeans3DCoord(); s.createClustersPyramidUsingKMeans(); }
/* one warp of pyramid level `image_level` with the current T_odometry (FrontEnd.cpp:775) */
void ref_warp_level(void* h, int image_level) {
    StaticFusion& s = *static_cast<StaticFusion*>(h);
    s.image_level = image_level; s.rows_i = s.rows >> image_level; s.cols_i = s.cols >> image_level;
    s.warpImagesAccurateInverse();
}
/* the linearisation stages of one step, in runSolver's order (FrontEnd.cpp:1097-1124); first != 0 copies Pred -> Warped */
void ref_linearise_level(void* h, int level_i, int first) {
    StaticFusion& s = *static_cast<StaticFusion*>(h);
    s.level = level_i;
    const unsigned int sc = pow(2.f, int(s.ctf_levels - (level_i + 1)));
    s.cols_i = s.cols / sc; s.rows_i = s.rows / sc;
    s.image_level = s.ctf_levels - level_i - 1;
    if (first) {
        s.depthWarpedPyr[s.image_level] = s.depthPredPyr[s.image_level];
        s.intensityWarpedPyr[s.image_level] = s.intensityPredPyr[s.image_level];
        s.xxWarpedPyr[s.image_level] = s.xxPredPyr[s.image_level];
        s.yyWarpedPyr[s.image_level] = s.yyPredPyr[s.image_level];
    } else
        s.warpImagesAccurateInverse();
    s.calculateCoord(); s.calculateDerivatives(); s.computeWeights(); s.computeSegPrior();
}
void ref_solve_level(void* h) { static_cast<StaticFusion*>(h)->solveOdometryAndSegmJoint(); }
#endif

/* state readers (row-major out) */
int ref_get_image(void* h, const char* name, int L, float* out) {
    StaticFusion& s = *static_cast<StaticFusion*>(h);
    const std::string n(name);
    const int r = s.rows >> L, c = s.cols >> L;
    const MatrixXf* m = nullptr;
    if (n == "depth") m = &s.depthPyr[L]; else if (n == "intensity") m = &s.intensityPyr[L];
    else if (n == "xx") m = &s.xxPyr[L]; else if (n == "yy") m = &s.yyPyr[L];
    else if (n == "depth_pred") m = &s.depthPredPyr[L]; else if (n == "intensity_pred") m = &s.intensityPredPyr[L];
    else if (n == "depth_warped") m = &s.depthWarpedPyr[L]; else if (n == "intensity_warped") m = &s.intensityWarpedPyr[L];
    else if (n == "depth_inter") m = &s.depthInterPyr[L]; else if (n == "intensity_inter") m = &s.intensityInterPyr[L];
    else if (n == "xx_inter") m = &s.xxInterPyr[L]; else if (n == "yy_inter") m = &s.yyInterPyr[L];
    else if (n == "dcu") m = &s.dcu; else if (n == "dcv") m = &s.dcv; else if (n == "dct") m = &s.dct;
    else if (n == "ddu") m = &s.ddu; else if (n == "ddv") m = &s.ddv; else if (n == "ddt") m = &s.ddt;
    else if (n == "weights_c") m = &s.weights_c; else if (n == "weights_d") m = &s.weights_d;
    else if (n == "b_segm_perpixel") m = &s.b_segm_perpixel;
    if (m) { out_rowmajor(*m, r, c, out); return 0; }
    if (n == "null") { for (int v = 0; v < r; v++) for (int u = 0; u < c; u++) out[(size_t)v * c + u] = s.Null(v, u) ? 1.f : 0.f; return 0; }
    return -1;
}
void ref_get_labels(void* h, int L, int* out) {
    StaticFusion& s = *static_cast<StaticFusion*>(h);
    const int r = s.rows >> L, c = s.cols >> L;
    for (int v = 0; v < r; v++) for (int u = 0; u < c; u++) out[(size_t)v * c + u] = s.clusterAllocation[L](v, u);
}
void ref_get_kmeans(void* h, float* out /*[3][24]*/, unsigned char* conn /*[24][24]*/) {
    StaticFusion& s = *static_cast<StaticFusion*>(h);
    for (int r = 0; r < 3; r++) for (int l = 0; l < NUM_CLUSTERS; l++) out[r * NUM_CLUSTERS + l] = s.kmeans(r, l);
    for (int i = 0; i < NUM_CLUSTERS; i++) for (int j = 0; j < NUM_CLUSTERS; j++) conn[i * NUM_CLUSTERS + j] = s.connectivity[i][j] ? 1 : 0;
}
void ref_get_pose(void* h, float* T_rowmajor, float* twist_odometry, float* twist_old, float* twist_level, float* est_cov) {
    StaticFusion& s = *static_cast<StaticFusion*>(h);
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) T_rowmajor[r * 4 + c] = s.T_odometry(r, c);
    for (int k = 0; k < 6; k++) { twist_odometry[k] = s.twist_odometry(k); twist_old[k] = s.twist_odometry_old(k); twist_level[k] = s.twist_level_odometry(k); }
    if (est_cov) for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) est_cov[r * 6 + c] = s.est_cov(r, c);
}
void ref_get_seg(void* h, float* b_segm, float* b_prior, float* lambda_t_w) {
    StaticFusion& s = *static_cast<StaticFusion*>(h);
    for (int l = 0; l < NUM_CLUSTERS; l++) { b_segm[l] = s.b_segm[l]; b_prior[l] = s.b_prior[l]; lambda_t_w[l] = s.lambda_t_w[l]; }
}
int ref_num_valid(void* h) { return (int)static_cast<StaticFusion*>(h)->validPixels.size(); }

}  /* extern "C" */
