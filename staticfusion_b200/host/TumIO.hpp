// TumIO.hpp — the wire formats either side of the solver, C++ host side (twin of staticfusion_b200/tum_io.py):
//   * loadAssoc                  StaticFusion::loadAssoc (FrontEnd.cpp:183-214): `rgbd_assoc.txt`
//   * read_png / imread_*        the two cv::imread calls of loadImageFromSequenceAssoc (FrontEnd.cpp:220,240).  OpenCV is not
//                                a dependency here: a small PNG reader (zlib inflate + the five scanline filters) covers what
//                                the TUM-format folders hold -- 8-bit RGB / RGBA / grey colour images and 16-bit grey depth,
//                                non-interlaced -- and yields cv::imread's layout (BGR bytes; native-endian uint16)
//   * Trajectory                 currPose = currPose * T_odometry (Reconstruction.cpp:256,265), Datasets::writeTrajectoryFile
//                                (Datasets.cpp:252-266) and the `.freiburg` pose graph of Reconstruction::savePly
//                                (Reconstruction.cpp:460-484): TUM format `timestamp tx ty tz qx qy qz qw`
// The flip / decimation / intensity conversion of the loader is device work: sf_convert_frames (include/staticfusion_b200.h).
#pragma once
#include <zlib.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace sfb200 {

// StaticFusion::loadAssoc (FrontEnd.cpp:183-214): same argument order, same return value, same parsing rules
inline bool loadAssoc(const std::string& dir, const std::string& assocFile, std::vector<double>& timestamps,
                      std::vector<std::string>& filesDepth, std::vector<std::string>& filesColor) {
    const std::string assocPath = dir + assocFile;
    if (assocPath.empty()) return false;
    std::ifstream assocIn(assocPath.c_str());
    if (!assocIn.is_open()) return false;
    std::string line;
    while (std::getline(assocIn, line)) {
        if (line.empty() || line.compare(0, 1, "#") == 0) continue;
        std::istringstream iss(line);
        double timestampDepth, timestampColor;
        std::string fileDepth, fileColor;
        if (!(iss >> timestampColor >> fileColor >> timestampDepth >> fileDepth)) break;
        timestamps.push_back(timestampDepth);
        filesDepth.push_back(dir + fileDepth);
        filesColor.push_back(dir + fileColor);
    }
    return true;
}

// ---- PNG ------------------------------------------------------------------------------------------------------
struct PngImage {
    int rows = 0, cols = 0, channels = 0, bit_depth = 0;  // as stored in the file
    std::vector<uint8_t> px;                               // rows * cols * channels samples, 16-bit samples native-endian
    bool empty() const { return px.empty(); }
};

namespace png_detail {
inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
}  // namespace png_detail

// Returns an empty image when the file does not exist (cv::imread's behaviour); throws on files it cannot represent.
inline PngImage read_png(const std::string& path) {
    using namespace png_detail;
    PngImage im;
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return im;
    std::vector<uint8_t> buf;
    uint8_t tmp[65536];
    size_t n;
    while ((n = std::fread(tmp, 1, sizeof(tmp), f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
    std::fclose(f);
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (buf.size() < 8 || std::memcmp(buf.data(), sig, 8) != 0) throw std::runtime_error(path + ": not a PNG file");
    std::vector<uint8_t> idat;
    int color_type = -1, interlace = 0;
    std::vector<uint8_t> palette;
    for (size_t pos = 8; pos + 12 <= buf.size();) {
        const uint32_t len = be32(&buf[pos]);
        const char* type = reinterpret_cast<const char*>(&buf[pos + 4]);
        const uint8_t* data = &buf[pos + 8];
        if (pos + 12 + len > buf.size()) throw std::runtime_error(path + ": truncated PNG chunk");
        if (!std::memcmp(type, "IHDR", 4)) {
            im.cols = (int)be32(data); im.rows = (int)be32(data + 4); im.bit_depth = data[8]; color_type = data[9]; interlace = data[12];
        } else if (!std::memcmp(type, "PLTE", 4)) palette.assign(data, data + len);
        else if (!std::memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
        else if (!std::memcmp(type, "IEND", 4)) break;
        pos += 12 + len;
    }
    if (interlace) throw std::runtime_error(path + ": interlaced PNG is not supported");
    if (im.bit_depth != 8 && im.bit_depth != 16) throw std::runtime_error(path + ": only 8- and 16-bit PNG samples are supported");
    const int file_channels = color_type == 0 ? 1 : color_type == 2 ? 3 : color_type == 3 ? 1 : color_type == 4 ? 2 : color_type == 6 ? 4 : 0;
    if (!file_channels || (color_type == 3 && im.bit_depth != 8)) throw std::runtime_error(path + ": unsupported PNG colour type");
    const size_t bps = (size_t)im.bit_depth / 8, bpp = bps * file_channels, stride = (size_t)im.cols * bpp;
    std::vector<uint8_t> raw((stride + 1) * im.rows);
    uLongf out_len = (uLongf)raw.size();
    if (uncompress(raw.data(), &out_len, idat.data(), (uLong)idat.size()) != Z_OK || out_len != raw.size())
        throw std::runtime_error(path + ": zlib inflate failed");
    // undo the scanline filters in place (PNG spec 9.2)
    std::vector<uint8_t> img(stride * im.rows);
    for (int y = 0; y < im.rows; y++) {
        const uint8_t ft = raw[(stride + 1) * y];
        const uint8_t* src = &raw[(stride + 1) * y + 1];
        uint8_t* cur = &img[stride * y];
        const uint8_t* up = y ? &img[stride * (y - 1)] : nullptr;
        for (size_t x = 0; x < stride; x++) {
            const int a = x >= bpp ? cur[x - bpp] : 0, b = up ? up[x] : 0, c = (up && x >= bpp) ? up[x - bpp] : 0;
            int v = src[x];
            switch (ft) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: v += paeth(a, b, c); break;
                default: throw std::runtime_error(path + ": bad PNG filter type");
            }
            cur[x] = (uint8_t)v;
        }
    }
    if (color_type == 3) {  // palette -> RGB
        im.channels = 3;
        im.px.resize((size_t)im.rows * im.cols * 3);
        for (size_t i = 0; i < (size_t)im.rows * im.cols; i++) {
            const size_t k = 3 * (size_t)img[i];
            if (k + 2 >= palette.size()) throw std::runtime_error(path + ": palette index out of range");
            im.px[3 * i] = palette[k]; im.px[3 * i + 1] = palette[k + 1]; im.px[3 * i + 2] = palette[k + 2];
        }
        return im;
    }
    im.channels = file_channels;
    if (bps == 2) {  // big-endian samples -> native uint16
        im.px.resize(img.size());
        uint16_t* o = reinterpret_cast<uint16_t*>(im.px.data());
        for (size_t i = 0; i < img.size() / 2; i++) o[i] = (uint16_t)((img[2 * i] << 8) | img[2 * i + 1]);
    } else
        im.px.swap(img);
    return im;
}

// cv::imread(path, CV_LOAD_IMAGE_COLOR): always 3 x 8-bit, BGR order (grey replicated, alpha dropped).  Empty when missing.
inline bool imread_color_bgr(const std::string& path, std::vector<uint8_t>& bgr, int& rows, int& cols) {
    const PngImage im = read_png(path);
    if (im.empty()) return false;
    if (im.bit_depth != 8) throw std::runtime_error(path + ": colour image must have 8-bit samples");
    rows = im.rows; cols = im.cols;
    bgr.resize((size_t)rows * cols * 3);
    for (size_t i = 0; i < (size_t)rows * cols; i++) {
        const uint8_t* s = &im.px[i * im.channels];
        if (im.channels >= 3) { bgr[3 * i] = s[2]; bgr[3 * i + 1] = s[1]; bgr[3 * i + 2] = s[0]; }
        else { bgr[3 * i] = bgr[3 * i + 1] = bgr[3 * i + 2] = s[0]; }
    }
    return true;
}
// cv::imread(path, -1) of a 16-bit single-channel depth image
inline bool imread_depth_u16(const std::string& path, std::vector<uint16_t>& depth, int& rows, int& cols) {
    const PngImage im = read_png(path);
    if (im.empty()) return false;
    if (im.bit_depth != 16 || im.channels != 1) throw std::runtime_error(path + ": depth image must be 16-bit single-channel");
    rows = im.rows; cols = im.cols;
    depth.resize((size_t)rows * cols);
    std::memcpy(depth.data(), im.px.data(), depth.size() * 2);
    return true;
}

// ---- trajectory -------------------------------------------------------------------------------------------------
struct Pose4f {  // row-major 4x4 float
    float m[16];
    Pose4f() { for (int i = 0; i < 16; i++) m[i] = (i % 5 == 0) ? 1.f : 0.f; }
    float& operator()(int r, int c) { return m[r * 4 + c]; }
    float operator()(int r, int c) const { return m[r * 4 + c]; }
    // Eigen Matrix4f product, float, sums in k order
    Pose4f operator*(const Pose4f& b) const {
        Pose4f o;
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) {
                float s = m[i * 4] * b.m[j];
                for (int k = 1; k < 4; k++) s += m[i * 4 + k] * b.m[k * 4 + j];
                o.m[i * 4 + j] = s;
            }
        return o;
    }
    static Pose4f fromColumnMajor(const float* cm) {  // Eigen::Matrix4f::data() / sf_get_outputs' T_odometry
        Pose4f o;
        for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) o(r, c) = cm[c * 4 + r];
        return o;
    }
};
// Eigen::Quaternionf(const Matrix3f&) (Eigen/src/Geometry/Quaternion.h): q = (x, y, z, w)
inline void quatFromRotation(const Pose4f& T, float q[4]) {
    float t = T(0, 0) + T(1, 1) + T(2, 2);
    if (t > 0.f) {
        t = std::sqrt(t + 1.f);
        q[3] = 0.5f * t;
        t = 0.5f / t;
        q[0] = (T(2, 1) - T(1, 2)) * t; q[1] = (T(0, 2) - T(2, 0)) * t; q[2] = (T(1, 0) - T(0, 1)) * t;
    } else {
        int i = 0;
        if (T(1, 1) > T(0, 0)) i = 1;
        if (T(2, 2) > T(i, i)) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(T(i, i) - T(j, j) - T(k, k) + 1.f);
        q[i] = 0.5f * t;
        t = 0.5f / t;
        q[3] = (T(k, j) - T(j, k)) * t; q[j] = (T(j, i) + T(i, j)) * t; q[k] = (T(k, i) + T(i, k)) * t;
    }
}

class Trajectory {
public:
    Pose4f currPose;                                   // Reconstruction.cpp:36
    std::vector<Pose4f> poseGraph;                     // Reconstruction.cpp:315
    std::vector<unsigned long long> poseLogTimes;      // Reconstruction.cpp:321
    Pose4f rotateByZ;                                  // Datasets.cpp:58-60: AngleAxisf(M_PI, UnitZ)
    Trajectory() {
        const float s = std::sin((float)M_PI), c = std::cos((float)M_PI);
        rotateByZ(0, 0) = c; rotateByZ(0, 1) = -s; rotateByZ(1, 0) = s; rotateByZ(1, 1) = c;
    }
    // the pose part of Reconstruction::fuseFrame (Reconstruction.cpp:255-265, 315-321); T_odometry column-major as the solver exports it
    const Pose4f& fuse(const float* T_odometry_colmajor, unsigned long long timestamp) {
        currPose = currPose * Pose4f::fromColumnMajor(T_odometry_colmajor);
        poseGraph.push_back(currPose);
        poseLogTimes.push_back(timestamp);
        return currPose;
    }
    // Datasets::writeTrajectoryFile (Datasets.cpp:252-266); ddt_sum = ddt.sumAll() (lines of equal consecutive depth images are skipped)
    bool writeTrajectoryLine(std::ostream& f_res, double timestamp_obs, float ddt_sum = 1.f) const {
        if (!(std::fabs(ddt_sum) > 0)) return false;
        const Pose4f convertedPose = currPose * rotateByZ;
        float q[4];
        quatFromRotation(convertedPose, q);
        char aux[24];
        std::snprintf(aux, sizeof(aux), "%.04f", timestamp_obs);
        f_res << aux << " " << convertedPose(0, 3) << " " << convertedPose(1, 3) << " " << convertedPose(2, 3) << " ";
        f_res << q[0] << " " << q[1] << " " << q[2] << " " << q[3] << std::endl;
        return true;
    }
    // the pose-graph part of Reconstruction::savePly (Reconstruction.cpp:460-484)
    void writeFreiburg(std::ostream& f) const {
        for (size_t i = 0; i < poseGraph.size(); i++) {
            std::stringstream strs;
            strs << std::setprecision(6) << std::fixed << (double)poseLogTimes.at(i) / 1000000.0 << " ";
            const Pose4f& P = poseGraph.at(i);
            f << strs.str() << P(0, 3) << " " << P(1, 3) << " " << P(2, 3) << " ";
            float q[4];
            quatFromRotation(P, q);
            f << q[0] << " " << q[1] << " " << q[2] << " " << q[3] << "\n";
        }
    }
    bool saveFreiburg(const std::string& saveFilename) const {
        std::ofstream f((saveFilename + ".freiburg").c_str(), std::fstream::out);
        if (!f.is_open()) return false;
        writeFreiburg(f);
        return true;
    }
};

}  // namespace sfb200
