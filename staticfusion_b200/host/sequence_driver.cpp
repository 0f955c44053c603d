// sequence_driver.cpp — StaticFusion-imagesequenceassoc.cpp (the reference's recorded-sequence driver) without its GL
// map: frame-to-frame odometry + segmentation over a TUM-format folder, through the C++ mirror of the reference class.
//   sequence_driver <dir> [max_frames] [out_prefix]
// <dir> holds rgbd_assoc.txt ("ts_rgb rgb_path ts_depth depth_path", FrontEnd.cpp:204) and 640x480 PNGs (8-bit colour,
// 16-bit depth in millimetres).  Call order and parameters follow StaticFusion-imagesequenceassoc.cpp:57-205; where the
// reference asks its map for the model prediction (getPredictedImages, :174) the previous frame is used instead
// (prediction := previous frame, as the reference itself does for the bootstrap pair, :110-111), and getFilteredDepth (:175)
// runs on the device.  Writes <out_prefix>.freiburg (pose graph, Reconstruction.cpp:460-484) and prints one line per frame.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "StaticFusionB200.hpp"

using namespace sfb200;

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "missing sequence directory\n"); return 2; }
    const std::string dir = argv[1];
    const size_t max_frames = argc > 2 ? (size_t)std::atol(argv[2]) : (size_t)-1;
    const std::string out_prefix = argc > 3 ? argv[3] : "trajectory";
    const unsigned int res_factor = 2;
    try {
        StaticFusion staticFusion(res_factor);
        staticFusion.use_motion_filter = true;                                     // :61
        staticFusion.ctf_levels = (unsigned)std::log2(staticFusion.cols / 40) + 2;  // :64
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

        int im_count = 1;                                                          // :83
        const unsigned int decimation = 1;
        std::vector<double> timestamps;
        std::vector<std::string> filesDepth, filesColor;
        const std::string assocFile = "/rgbd_assoc.txt";
        staticFusion.loadAssoc(dir, assocFile, timestamps, filesDepth, filesColor);
        if (filesDepth.empty() || filesColor.empty()) throw std::runtime_error("no image files");
        if (filesDepth.size() < 3) throw std::runtime_error("need at least three frames (the driver starts at index 1)");
        Trajectory trajectory;

        auto push_buffers = [&](const MatrixXf& d, const MatrixXf& i, const Matrix4f& T) {
            staticFusion.depthBuffer[im_count % staticFusion.bufferLength] = d.replicate(1, 1);
            staticFusion.intensityBuffer[im_count % staticFusion.bufferLength] = i.replicate(1, 1);
            staticFusion.odomBuffer[im_count % staticFusion.bufferLength] = T;
        };
        auto report = [&]() {
            double mean_b = 0;
            for (float x : staticFusion.b_segm_perpixel.a) mean_b += x;
            const Pose4f& P = trajectory.fuse(staticFusion.T_odometry.data(), (unsigned long long)im_count);  // fuseFrame(..., im_count, &T_odometry, ...) :132
            std::printf("frame %d ts %.4f  pose %.5f %.5f %.5f  irls %d status %d  mean static weight %.4f\n", im_count, timestamps[im_count],
                        P(0, 3), P(1, 3), P(2, 3), staticFusion.irls_iterations, staticFusion.status, mean_b / staticFusion.b_segm_perpixel.a.size());
        };

        // bootstrap pair, :103-135
        staticFusion.loadImageFromSequenceAssoc(filesDepth[im_count], filesColor[im_count], res_factor);
        staticFusion.depthPrediction.swap(staticFusion.depthCurrent);
        staticFusion.intensityPrediction.swap(staticFusion.intensityCurrent);
        push_buffers(staticFusion.depthPrediction, staticFusion.intensityPrediction, Matrix4f());
        im_count += decimation;
        staticFusion.loadImageFromSequenceAssoc(filesDepth[im_count], filesColor[im_count], res_factor);
        staticFusion.createImagePyramid(true);
        staticFusion.kb = 1.05f;
        staticFusion.runSolver(true);
        staticFusion.buildSegmImage();
        push_buffers(staticFusion.depthCurrent, staticFusion.intensityCurrent, staticFusion.T_odometry);
        report();

        // steady state, :141-200
        while (!((im_count + decimation) >= filesDepth.size()) && (size_t)im_count + 1 < max_frames) {
            im_count += decimation;
            staticFusion.depthPrediction.swap(staticFusion.depthCurrent);          // stands in for getPredictedImages (:174)
            staticFusion.intensityPrediction.swap(staticFusion.intensityCurrent);
            if (staticFusion.loadImageFromSequenceAssoc(filesDepth[im_count], filesColor[im_count], res_factor)) break;
            staticFusion.kb = 1.5f;                                                 // :160-170 once the model is initialised
            staticFusion.getFilteredDepth(staticFusion.depth_mm.data(), staticFusion.depthCurrent);  // :175
            staticFusion.createImagePyramid(true);
            staticFusion.runSolver(true);
            if (im_count - staticFusion.bufferLength >= 0) staticFusion.computeResidualsAgainstPreviousImage(im_count);
            staticFusion.buildSegmImage();
            push_buffers(staticFusion.depthCurrent, staticFusion.intensityCurrent, staticFusion.T_odometry);
            report();
        }
        if (!trajectory.saveFreiburg(out_prefix)) throw std::runtime_error("could not write " + out_prefix + ".freiburg");
        std::printf("wrote %s.freiburg (%zu poses)\n", out_prefix.c_str(), trajectory.poseGraph.size());
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
