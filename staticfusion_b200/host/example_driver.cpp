// example_driver.cpp — the reference driver's call sequence (StaticFusion-datasets.cpp:75-144, 148-190) against the
// B200 solver through the C++ mirror.  Inputs: raw float32 column-major files or a built-in synthetic ramp scene.
//   example_driver [depth_prev.f32 inten_prev.f32 depth_cur.f32 inten_cur.f32]
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "StaticFusionB200.hpp"

using namespace sfb200;

static bool load(const char* path, MatrixXf& m) {
    FILE* f = std::fopen(path, "rb");
    if (!f) return false;
    const size_t n = std::fread(m.data(), sizeof(float), m.a.size(), f);
    std::fclose(f);
    return n == m.a.size();
}

static void synthetic(MatrixXf& d, MatrixXf& c, float shift) {
    for (int u = 0; u < d.cols(); u++)
        for (int v = 0; v < d.rows(); v++) {
            const float x = (u + shift) * 0.02f, y = v * 0.02f;
            d(v, u) = std::round((1.5f + 0.4f * std::sin(0.7f * x) + 0.3f * std::cos(0.9f * y)) * 1000.f) / 1000.f;
            c(v, u) = 0.5f + 0.25f * std::sin(3.1f * x + 0.5f * y) + 0.2f * std::sin(5.3f * y - x);
        }
}

int main(int argc, char** argv) {
    const unsigned int res_factor = 2;
    StaticFusion staticFusion(res_factor);
    // parameter block of the drivers, StaticFusion-datasets.cpp:79-94
    staticFusion.use_motion_filter = true;
    staticFusion.ctf_levels = (unsigned)std::log2(staticFusion.cols / 40) + 2;
    staticFusion.max_iter_per_level = 3;
    staticFusion.previous_speed_const_weight = 0.1f;
    staticFusion.previous_speed_eig_weight = 2.f;
    staticFusion.k_photometric_res = 0.15f;
    staticFusion.irls_delta_threshold = 0.0015f;
    staticFusion.max_iter_irls = 6;
    staticFusion.lambda_reg = 0.35f;
    staticFusion.lambda_prior = 0.5f;
    staticFusion.kc_Cauchy = 0.5f;
    staticFusion.kb = 1.5f;
    staticFusion.kz = 1.5f;

    if (argc == 5) {
        if (!load(argv[1], staticFusion.depthPrediction) || !load(argv[2], staticFusion.intensityPrediction) ||
            !load(argv[3], staticFusion.depthCurrent) || !load(argv[4], staticFusion.intensityCurrent)) {
            std::fprintf(stderr, "could not read the four %ux%u float32 column-major images\n", staticFusion.rows, staticFusion.cols);
            return 2;
        }
    } else {
        synthetic(staticFusion.depthPrediction, staticFusion.intensityPrediction, 0.f);
        synthetic(staticFusion.depthCurrent, staticFusion.intensityCurrent, 1.5f);
    }
    int im_count = 0;
    try {
        // bootstrap pair, StaticFusion-datasets.cpp:109-144
        staticFusion.depthBuffer[im_count % staticFusion.bufferLength] = staticFusion.depthPrediction.replicate(1,1);
        staticFusion.intensityBuffer[im_count % staticFusion.bufferLength] = staticFusion.intensityPrediction.replicate(1,1);
        staticFusion.odomBuffer[im_count % staticFusion.bufferLength] = Matrix4f();
        staticFusion.createImagePyramid(true);
        staticFusion.kb = 1.05f;
        im_count++;
        staticFusion.runSolver(true);
        staticFusion.buildSegmImage();
        staticFusion.depthBuffer[im_count % staticFusion.bufferLength] = staticFusion.depthCurrent.replicate(1,1);
        staticFusion.intensityBuffer[im_count % staticFusion.bufferLength] = staticFusion.intensityCurrent.replicate(1,1);
        staticFusion.odomBuffer[im_count % staticFusion.bufferLength] = staticFusion.T_odometry;
        // steady state, StaticFusion-datasets.cpp:148-184 (frame-to-frame: prediction := previous frame), synthetic input only
        for (int k = 0; argc != 5 && k < 6; k++) {
            staticFusion.depthPrediction.swap(staticFusion.depthCurrent);
            staticFusion.intensityPrediction.swap(staticFusion.intensityCurrent);
            synthetic(staticFusion.depthCurrent, staticFusion.intensityCurrent, 1.5f * (k + 2));
            im_count++;
            staticFusion.kb = 1.5f;
            staticFusion.createImagePyramid(true);
            staticFusion.runSolver(true);
            if (im_count - staticFusion.bufferLength >= 0) staticFusion.computeResidualsAgainstPreviousImage(im_count);
            staticFusion.buildSegmImage();
            staticFusion.depthBuffer[im_count % staticFusion.bufferLength] = staticFusion.depthCurrent.replicate(1,1);
            staticFusion.intensityBuffer[im_count % staticFusion.bufferLength] = staticFusion.intensityCurrent.replicate(1,1);
            staticFusion.odomBuffer[im_count % staticFusion.bufferLength] = staticFusion.T_odometry;
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    std::printf("T_odometry (column-major Eigen::Matrix4f layout):\n");
    for (int r = 0; r < 4; r++)
        std::printf("  % .6f % .6f % .6f % .6f\n", staticFusion.T_odometry(r, 0), staticFusion.T_odometry(r, 1),
                    staticFusion.T_odometry(r, 2), staticFusion.T_odometry(r, 3));
    double mean_b = 0;
    for (float x : staticFusion.b_segm_perpixel.a) mean_b += x;
    std::printf("frames %d, perClusterAverageResidual[0] %.5f\n", im_count, staticFusion.perClusterAverageResidual[0]);
    std::printf("irls iterations %d, status %d, mean static weight %.4f\n", staticFusion.irls_iterations, staticFusion.status,
                mean_b / staticFusion.b_segm_perpixel.a.size());
    return 0;
}
