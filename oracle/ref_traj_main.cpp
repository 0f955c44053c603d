/*
 * ref_traj_main.cpp — the REFERENCE's own trajectory writers, compiled from where they lie (TEST INFRASTRUCTURE):
 *   Datasets::writeTrajectoryFile           Utils/Datasets.cpp:252-266      (piped into the compiler after part A)
 *   the pose-graph block of Reconstruction::savePly   Reconstruction.cpp:460-485   (piped in after part B)
 * against oracle/ref_shim (Eigen stand-in: Matrix4f product, corner blocks, the Quaternionf(Matrix3f) constructor); the text
 * formatting itself (sprintf, std::ostream << float, setprecision / fixed) is the real C++ library, which is what this pins.
 * The translation unit is assembled by oracle/Makefile (target ref_traj): this file is cut at the two marker lines.
 *
 * stdin:  n, then n records of (double timestamp_obs, unsigned long long pose-log time, float currPose[16] row-major)
 * stdout: the Datasets text, a line "--", the .freiburg text
 */
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

using namespace std;

/* the members of Utils/Datasets.h:68-74 that writeTrajectoryFile touches */
class Datasets {
public:
    std::ofstream f_res;
    Eigen::Matrix4f rotateByZ;
    double timestamp_obs;
    void writeTrajectoryFile(Eigen::Matrix4f currPose, Eigen::MatrixXf& ddt);
};
//@@PART_A_END: Utils/Datasets.cpp:252-266 follows
//@@PART_B_BEGIN
/* the members of Reconstruction.h:185,209,212 that the pose-graph block of savePly reads */
static void save_pose_graph(const std::string& saveFilename, const std::vector<std::pair<unsigned long long int, Eigen::Matrix4f> >& poseGraph,
                            const std::vector<unsigned long long int>& poseLogTimes) {
//@@PART_B_END: Reconstruction.cpp:460-485 follows (its closing brace ends this function)
//@@PART_C_BEGIN
int main(int argc, char** argv) {
    if (argc < 2) return 2;
    const std::string prefix = argv[1];
    int n = 0;
    if (!(std::cin >> n)) return 2;
    Datasets ds;
    ds.f_res.open((prefix + ".txt").c_str());
    ds.rotateByZ = Eigen::Matrix4f::Identity();  /* Datasets.cpp:58-60: AngleAxisf(M_PI, UnitZ).toRotationMatrix() in float */
    const float s = std::sin((float)M_PI), c = std::cos((float)M_PI);
    ds.rotateByZ(0, 0) = c; ds.rotateByZ(0, 1) = -s; ds.rotateByZ(1, 0) = s; ds.rotateByZ(1, 1) = c;
    std::vector<std::pair<unsigned long long int, Eigen::Matrix4f> > poseGraph;
    std::vector<unsigned long long int> poseLogTimes;
    Eigen::MatrixXf ddt(1, 1);
    for (int k = 0; k < n; k++) {
        double ts; unsigned long long lt; float ddt_sum;
        Eigen::Matrix4f P;
        std::cin >> ts >> lt >> ddt_sum;
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { float v; std::cin >> v; P(i, j) = v; }
        ddt(0, 0) = ddt_sum;
        ds.timestamp_obs = ts;
        ds.writeTrajectoryFile(P, ddt);
        poseGraph.push_back(std::make_pair(lt, P));
        poseLogTimes.push_back(lt);
    }
    ds.f_res.close();
    save_pose_graph(prefix, poseGraph, poseLogTimes);
    return 0;
}
