// StaticFusionB200.hpp — C++ host-side mirror of `class StaticFusion` (reference StaticFusion.h:66-189) for the ONE
// path this project replaces.  Same public field and method names, same argument meaning, same call order as the
// reference drivers use (StaticFusion-datasets.cpp:109-200); everything else of the reference class (GUI, GL
// Reconstruction, loaders) is out of scope and absent.  All compute is forwarded to the C ABI of
// libstaticfusion_b200.so (include/staticfusion_b200.h); there is no CPU fallback.
//
// The reference stores images as Eigen::MatrixXf (column-major).  Eigen is not a dependency here: `MatrixXf` below
// is a minimal column-major float matrix with the same (row, col) indexing and data() layout, so an Eigen user can
// pass `m.data()` straight through.
#pragma once
#include <cstring>
#include <cstdint>
#include <cstdio>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/staticfusion_b200.h"
#include "TumIO.hpp"

namespace sfb200 {

struct MatrixXf {  // column-major, like Eigen::MatrixXf
    int rows_ = 0, cols_ = 0;
    std::vector<float> a;
    MatrixXf() = default;
    MatrixXf(int r, int c) : rows_(r), cols_(c), a((size_t)r * c, 0.f) {}
    void resize(int r, int c) { rows_ = r; cols_ = c; a.assign((size_t)r * c, 0.f); }
    void fill(float v) { for (auto& x : a) x = v; }
    int rows() const { return rows_; }
    int cols() const { return cols_; }
    float& operator()(int v, int u) { return a[(size_t)u * rows_ + v]; }
    float operator()(int v, int u) const { return a[(size_t)u * rows_ + v]; }
    float* data() { return a.data(); }
    const float* data() const { return a.data(); }
    void swap(MatrixXf& o) { std::swap(rows_, o.rows_); std::swap(cols_, o.cols_); a.swap(o.a); }
    MatrixXf replicate(int, int) const { return *this; }  // the drivers copy with replicate(1,1)
};
struct MatrixXi {
    int rows_ = 0, cols_ = 0;
    std::vector<int32_t> a;
    void resize(int r, int c) { rows_ = r; cols_ = c; a.assign((size_t)r * c, 0); }
    int32_t& operator()(int v, int u) { return a[(size_t)u * rows_ + v]; }
    int32_t operator()(int v, int u) const { return a[(size_t)u * rows_ + v]; }
    int32_t* data() { return a.data(); }
};
struct Matrix4f {  // column-major 4x4, like Eigen::Matrix4f
    float m[16];
    Matrix4f() { setIdentity(); }
    void setIdentity() { for (int i = 0; i < 16; i++) m[i] = (i % 5 == 0) ? 1.f : 0.f; }
    float& operator()(int r, int c) { return m[c * 4 + r]; }
    float operator()(int r, int c) const { return m[c * 4 + r]; }
    const float* data() const { return m; }
};

class StaticFusion {
public:
    // ---- fields the drivers write (StaticFusion.h:88-89, 115-146, 171-172) ----
    MatrixXf depthCurrent, intensityCurrent;        // new frame
    MatrixXf depthPrediction, intensityPrediction;  // model prediction / previous frame
    unsigned int rows, cols, width, height, ctf_levels;
    float fovh;
    bool use_motion_filter;
    float previous_speed_const_weight, previous_speed_eig_weight;
    unsigned int max_iter_irls, max_iter_per_level;
    float k_photometric_res, irls_delta_threshold, kc_Cauchy, kb;
    float lambda_reg, lambda_prior, kz;
    // ---- fields the drivers read (StaticFusion.h:110-111, 155, 166-167) ----
    Matrix4f T_odometry;
    float twist_odometry_old[6];
    std::vector<MatrixXi> clusterAllocation;  // only level 0 is exported (what the backend consumes, Reconstruction.h:173)
    float b_segm[SF_NUM_CLUSTERS];
    MatrixXf b_segm_perpixel;
    // ---- 5-frame history the drivers maintain (StaticFusion.h:92-96; StaticFusion-datasets.cpp:114-116, 182-184) ----
    std::vector<MatrixXf> depthBuffer, intensityBuffer;
    std::vector<Matrix4f> odomBuffer;
    int bufferLength = 5;
    float perClusterAverageResidual[SF_NUM_CLUSTERS];  // NaN until computeResidualsAgainstPreviousImage ran (FrontEnd.cpp:105)
    // ---- what the image-sequence loader fills besides depthCurrent / intensityCurrent (StaticFusion.h:71; cv::Mat there) ----
    std::vector<uint16_t> depth_mm;   // rows x cols, row-major millimetres  (handed to fuseFrame / getFilteredDepth)
    std::vector<uint8_t> color_full;  // rows x cols x 3, row-major            (handed to fuseFrame)
    int irls_iterations = 0, status = 0;  // extras: SF_STATUS_* bits replace the reference's undefined behaviour

    // StaticFusion::StaticFusion(res_factor), FrontEnd.cpp:52-181 (solver part only)
    explicit StaticFusion(unsigned int res_factor, int device = 0) {
        sf_params p;
        sf_default_params(&p, 480 / (int)res_factor, 640 / (int)res_factor);
        rows = height = p.rows; cols = width = p.cols; ctf_levels = p.ctf_levels; fovh = p.fovh;
        // constructor defaults of the reference (FrontEnd.cpp:66-76); the drivers overwrite them
        k_photometric_res = 0.15f; irls_delta_threshold = 1e-6f; max_iter_irls = 10; max_iter_per_level = 2;
        previous_speed_const_weight = 0.05f; previous_speed_eig_weight = 0.5f; kc_Cauchy = 0.5f; kb = 1.25f; kz = 1.5f;
        use_motion_filter = false;
        lambda_reg = 0.35f; lambda_prior = 0.5f;  // uninitialised in the reference; driver values used
        depthCurrent.resize(rows, cols); intensityCurrent.resize(rows, cols);
        depthPrediction.resize(rows, cols); intensityPrediction.resize(rows, cols);
        b_segm_perpixel.resize(rows, cols); b_segm_perpixel.fill(0.5f);
        clusterAllocation.resize(1); clusterAllocation[0].resize(rows, cols);
        for (int i = 0; i < 6; i++) twist_odometry_old[i] = 0.f;
        for (int l = 0; l < SF_NUM_CLUSTERS; l++) b_segm[l] = 0.5f;
        depthBuffer.resize(bufferLength); intensityBuffer.resize(bufferLength); odomBuffer.resize(bufferLength);  // FrontEnd.cpp:96-103
        for (int i = 0; i < bufferLength; i++) { depthBuffer[i].resize(rows, cols); intensityBuffer[i].resize(rows, cols); }
        for (int l = 0; l < SF_NUM_CLUSTERS; l++) perClusterAverageResidual[l] = std::numeric_limits<float>::quiet_NaN();
        device_ = device;
    }
    ~StaticFusion() { sf_destroy(ctx_); }
    StaticFusion(const StaticFusion&) = delete;
    StaticFusion& operator=(const StaticFusion&) = delete;

    // StaticFusion::loadAssoc, FrontEnd.cpp:183
    bool loadAssoc(const std::string& dir, const std::string& assocFile, std::vector<double>& timestamps,
                   std::vector<std::string>& filesDepth, std::vector<std::string>& filesColor) const {
        return sfb200::loadAssoc(dir, assocFile, timestamps, filesDepth, filesColor);
    }
    // StaticFusion::loadImageFromSequenceAssoc, FrontEnd.cpp:216: returns true at the end of the sequence (colour image
    // missing).  PNG decode on the host, flip / decimation / conversions on the device (sf_convert_frames).
    bool loadImageFromSequenceAssoc(const std::string& depthFile, const std::string& rgbFile, unsigned int res_factor) {
        std::vector<uint8_t> bgr;
        std::vector<uint16_t> raw;
        int r = 0, c = 0, rd = 0, cd = 0;
        if (!imread_color_bgr(rgbFile, bgr, r, c)) {
            std::printf("End of sequence (or color image not found...)\n");
            return true;
        }
        if (!imread_depth_u16(depthFile, raw, rd, cd)) throw std::runtime_error("staticfusion_b200: depth image not found: " + depthFile);
        if (r != (int)(height * res_factor) || c != (int)(width * res_factor) || rd != r || cd != c)
            throw std::runtime_error("staticfusion_b200: images must be (height x width) * res_factor");
        ensure();
        depth_mm.resize((size_t)rows * cols); color_full.resize((size_t)rows * cols * 3);
        check(sf_convert_frames(ctx_, 1, bgr.data(), raw.data(), (int)res_factor, SF_MEM_HOST, intensityCurrent.data(), depthCurrent.data(),
                                depth_mm.data(), color_full.data(), SF_MEM_HOST, 1));
        return false;
    }
    // StaticFusion::createImagePyramid(bool old_im), FrontEnd.cpp:256
    void createImagePyramid(bool old_im) {
        ensure();
        if (old_im) check(sf_set_prediction(ctx_, depthPrediction.data(), intensityPrediction.data(), 1));
        else check(sf_set_current(ctx_, depthCurrent.data(), intensityCurrent.data(), 1));
        check(sf_create_image_pyramid(ctx_, old_im ? 1 : 0));
    }
    // StaticFusion::runSolver(bool create_image_pyr), FrontEnd.cpp:1071
    void runSolver(bool create_image_pyr) {
        ensure();
        if (create_image_pyr) check(sf_set_current(ctx_, depthCurrent.data(), intensityCurrent.data(), 1));
        check(sf_set_twist_old(ctx_, twist_odometry_old));
        check(sf_run_solver(ctx_, create_image_pyr ? 1 : 0));
        check(sf_get_outputs(ctx_, T_odometry.m, twist_odometry_old, b_segm, nullptr, nullptr, 1, &irls_iterations, &status));
    }
    // Reconstruction::getFilteredDepth(cv::Mat depth, Eigen::MatrixXf& depthMat), Reconstruction.cpp:722-732: row-major u16
    // millimetres in, column-major metres out (what cv::cv2eigen produced); max_depth = the GUI's depthCutoff
    void getFilteredDepth(const uint16_t* depth_mm, MatrixXf& depthMat, float max_depth = 4.5f) {
        ensure();
        depthMat.resize(rows, cols);
        check(sf_filter_depth(ctx_, 1, depth_mm, SF_MEM_HOST, max_depth, depthMat.data(), SF_MEM_HOST, 1));
    }
    // StaticFusion::computeResidualsAgainstPreviousImage(int index), FrontEnd.cpp:896: between runSolver and buildSegmImage
    // once im_count >= bufferLength (StaticFusion-datasets.cpp:175-177).  The ring buffers are public members the drivers
    // assign directly, so the slots this call reads are pushed to the device here: the image of five frames ago and the
    // four increments in between.
    void computeResidualsAgainstPreviousImage(int index) {
        ensure();
        const int idx_to_warp = (index - bufferLength) % bufferLength;
        check(sf_buffer_set(ctx_, idx_to_warp, depthBuffer[idx_to_warp].data(), intensityBuffer[idx_to_warp].data(), odomBuffer[idx_to_warp].m, 1));
        for (int i = index - bufferLength + 1; i < index; i++)
            check(sf_buffer_set(ctx_, i % bufferLength, nullptr, nullptr, odomBuffer[i % bufferLength].m, 1));
        check(sf_compute_residuals_against_previous_image(ctx_, index));
        check(sf_get_per_cluster_average_residual(ctx_, perClusterAverageResidual));
    }
    // StaticFusion::buildSegmImage(), SegmentationBackground.cpp:176
    void buildSegmImage() {
        ensure();
        check(sf_build_segm_image(ctx_));
        check(sf_get_outputs(ctx_, nullptr, nullptr, nullptr, b_segm_perpixel.data(), clusterAllocation[0].data(), 1, nullptr, nullptr));
    }

private:
    sf_ctx* ctx_ = nullptr;
    int device_ = 0;
    sf_params cur_{};
    static void check(int rc) {
        if (rc != SF_OK) throw std::runtime_error(std::string("staticfusion_b200: ") + sf_last_error());
    }
    sf_params params() const {
        sf_params p;
        sf_default_params(&p, (int)rows, (int)cols);
        p.ctf_levels = (int)ctf_levels; p.max_iter_per_level = (int)max_iter_per_level; p.max_iter_irls = (int)max_iter_irls;
        p.use_motion_filter = use_motion_filter ? 1 : 0; p.fovh = fovh; p.k_photometric_res = k_photometric_res;
        p.irls_delta_threshold = irls_delta_threshold; p.kc_cauchy = kc_Cauchy; p.kb = kb; p.kz = kz;
        p.lambda_reg = lambda_reg; p.lambda_prior = lambda_prior;
        p.previous_speed_const_weight = previous_speed_const_weight; p.previous_speed_eig_weight = previous_speed_eig_weight;
        return p;
    }
    // the drivers assign the public fields after construction: (re)create / update the context lazily
    void ensure() {
        const sf_params p = params();
        if (!ctx_ || p.ctf_levels != cur_.ctf_levels || p.max_iter_per_level != cur_.max_iter_per_level) {
            sf_destroy(ctx_); ctx_ = nullptr;
            check(sf_create(&ctx_, &p, device_, 1, 0));
        } else if (std::memcmp(&p, &cur_, sizeof(sf_params)) != 0) {
            check(sf_set_params(ctx_, &p));  // only when a field changed (kb per frame, StaticFusion-datasets.cpp:156-165)
        }
        cur_ = p;
    }
};

}  // namespace sfb200
