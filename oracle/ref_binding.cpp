/*
 * ref_binding.cpp — THE BINDING OF INTEGRATION.md, COMPILED: the reference's own `class StaticFusion` (StaticFusion.h,
 * unmodified, from /root/reference) with its solver methods forwarded to libstaticfusion_b200.so through the C ABI.
 * TEST INFRASTRUCTURE (oracle/Makefile target `ref_b200` -> oracle/_ref/libsf_ref_b200.so): it proves on hardware that the
 * 30-line forwarding shown in INTEGRATION.md drops into the class and that a driver loop (StaticFusion-datasets.cpp:109-184,
 * replayed by oracle/ref_harness.cpp's C ABI) gets the CUDA path's results through the reference's own public fields.
 *
 * Differences from the text in INTEGRATION.md: the context pointer lives in a side table instead of a new member, because
 * the reference header is compiled as it lies in /root/reference and cannot gain a field here.
 */
#include <StaticFusion.h>

#include <map>
#include <stdexcept>
#include <string>

#include "../include/staticfusion_b200.h"

using namespace Eigen;

static std::map<const StaticFusion*, sf_ctx*>& table() { static std::map<const StaticFusion*, sf_ctx*> t; return t; }
static std::map<const StaticFusion*, sf_params>& last_params() { static std::map<const StaticFusion*, sf_params> t; return t; }
static void ok(int rc) { if (rc != SF_OK) throw std::runtime_error(std::string("staticfusion_b200: ") + sf_last_error()); }

static sf_params toParams(const StaticFusion& s) {
    sf_params p; sf_default_params(&p, (int)s.rows, (int)s.cols);
    p.ctf_levels = (int)s.ctf_levels; p.max_iter_per_level = (int)s.max_iter_per_level; p.max_iter_irls = (int)s.max_iter_irls;
    p.use_motion_filter = s.use_motion_filter ? 1 : 0; p.fovh = s.fovh; p.k_photometric_res = s.k_photometric_res;
    p.irls_delta_threshold = s.irls_delta_threshold; p.kc_cauchy = s.kc_Cauchy; p.kb = s.kb; p.kz = s.kz;
    p.lambda_reg = s.lambda_reg; p.lambda_prior = s.lambda_prior;
    p.previous_speed_const_weight = s.previous_speed_const_weight; p.previous_speed_eig_weight = s.previous_speed_eig_weight;
    return p;
}
/* create once (the drivers assign the tunables after construction), then push the possibly changed tunables (kb per frame) */
static sf_ctx* sync(StaticFusion& s) {
    const sf_params p = toParams(s);
    sf_ctx*& c = table()[&s];
    sf_params& cur = last_params()[&s];
    if (c && (p.ctf_levels != cur.ctf_levels || p.max_iter_per_level != cur.max_iter_per_level)) { sf_destroy(c); c = nullptr; }
    if (!c) ok(sf_create(&c, &p, /*device*/ 0, /*max_batch*/ 1, 0));
    else ok(sf_set_params(c, &p));
    cur = p;
    return c;
}
extern "C" void ref_binding_release(void* h) {
    auto it = table().find(static_cast<const StaticFusion*>(h));
    if (it != table().end()) { sf_destroy(it->second); table().erase(it); }
}

/* ---- the three solver entry points (StaticFusion.h:126,135,177) ------------------------------------------------ */
void StaticFusion::createImagePyramid(bool old_im) {
    sf_ctx* b200 = sync(*this);
    if (old_im) ok(sf_set_prediction(b200, depthPrediction.data(), intensityPrediction.data(), /*col_major*/ 1));
    else        ok(sf_set_current(b200, depthCurrent.data(), intensityCurrent.data(), 1));
    ok(sf_create_image_pyramid(b200, old_im ? 1 : 0));
}
void StaticFusion::runSolver(bool create_image_pyr) {
    sf_ctx* b200 = sync(*this);
    if (create_image_pyr) ok(sf_set_current(b200, depthCurrent.data(), intensityCurrent.data(), 1));
    ok(sf_set_twist_old(b200, twist_odometry_old.data()));
    ok(sf_run_solver(b200, create_image_pyr ? 1 : 0));
    ok(sf_get_outputs(b200, T_odometry.data(), twist_odometry_old.data(), b_segm.data(), nullptr, nullptr, 1, nullptr, nullptr));
    cam_oldpose = cam_pose;                                    /* FrontEnd.cpp:1134-1137 stays on the host (MRPT, double) */
    mrpt::math::CMatrixDouble44 aux_acu = T_odometry;
    cam_pose = cam_pose + mrpt::poses::CPose3D(aux_acu);
}
void StaticFusion::buildSegmImage() {
    sf_ctx* b200 = sync(*this);
    ok(sf_build_segm_image(b200));
    ok(sf_get_outputs(b200, nullptr, nullptr, nullptr, b_segm_perpixel.data(), clusterAllocation[0].data(), /*col_major*/ 1, nullptr, nullptr));
}
/* ---- 5-frame residual check (StaticFusion.h:132; the ring buffers stay public members the drivers write) ---------- */
void StaticFusion::computeResidualsAgainstPreviousImage(int index) {
    sf_ctx* b200 = sync(*this);
    const int idx_to_warp = (index - bufferLength) % bufferLength;                      /* FrontEnd.cpp:898 */
    ok(sf_buffer_set(b200, idx_to_warp, depthBuffer[idx_to_warp].data(), intensityBuffer[idx_to_warp].data(), odomBuffer[idx_to_warp].data(), 1));
    for (int i = index - bufferLength + 1; i < index; i++)                               /* the increments in between, :901-909 */
        ok(sf_buffer_set(b200, i % bufferLength, nullptr, nullptr, odomBuffer[i % bufferLength].data(), 1));
    ok(sf_compute_residuals_against_previous_image(b200, index));                        /* before buildSegmImage */
    ok(sf_get_per_cluster_average_residual(b200, perClusterAverageResidual.data()));
}
/* ---- recorded sequences: decode as before (the shim's cv::imread), convert on the device -------------------------- */
bool StaticFusion::loadImageFromSequenceAssoc(const std::string& depthFile, const std::string& rgbFile, unsigned int res_factor) {
    sf_ctx* b200 = sync(*this);
    cv::Mat color = cv::imread(rgbFile.c_str(), CV_LOAD_IMAGE_COLOR);
    if (color.data == NULL) { printf("End of sequence (or color image not found...)\n"); return true; }
    cv::Mat depth = cv::imread(depthFile.c_str(), -1);
    ok(sf_convert_frames(b200, 1, color.data, (const uint16_t*)depth.data, (int)res_factor, SF_MEM_HOST, intensityCurrent.data(),
                         depthCurrent.data(), (uint16_t*)depth_mm.data, color_full.data, SF_MEM_HOST, /*col_major*/ 1));
    return false;
}
