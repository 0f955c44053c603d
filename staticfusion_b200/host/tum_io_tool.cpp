// tum_io_tool.cpp — CPU-only exerciser of TumIO.hpp for the test-suite (no GPU, no solver):
//   tum_io_tool assoc <dir> <assocFile>          -> one line per entry: "<timestamp %.17g> <depth path> <colour path>"
//   tum_io_tool color <png> <out.bin>            -> rows cols on stdout, BGR bytes in out.bin     (cv::imread COLOR layout)
//   tum_io_tool depth <png> <out.bin>            -> rows cols on stdout, uint16 samples in out.bin (cv::imread -1 layout)
//   tum_io_tool traj  <in.txt>                   -> in: lines "timestamp_us ts_obs ddt_sum T(16 floats, column-major)";
//                                                   out: the Datasets.cpp trajectory lines, a line "--", then the .freiburg text
#include <cstdio>
#include <iostream>

#include "TumIO.hpp"

using namespace sfb200;

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    const std::string cmd = argv[1];
    try {
        if (cmd == "assoc" && argc == 4) {
            std::vector<double> ts;
            std::vector<std::string> fd, fc;
            if (!loadAssoc(argv[2], argv[3], ts, fd, fc)) { std::printf("FAILED\n"); return 0; }
            for (size_t k = 0; k < ts.size(); k++) std::printf("%.17g %s %s\n", ts[k], fd[k].c_str(), fc[k].c_str());
            return 0;
        }
        if ((cmd == "color" || cmd == "depth") && argc == 4) {
            int rows = 0, cols = 0;
            std::vector<uint8_t> bgr;
            std::vector<uint16_t> d;
            const bool ok = cmd == "color" ? imread_color_bgr(argv[2], bgr, rows, cols) : imread_depth_u16(argv[2], d, rows, cols);
            if (!ok) { std::printf("MISSING\n"); return 0; }
            FILE* f = std::fopen(argv[3], "wb");
            if (!f) return 3;
            if (cmd == "color") std::fwrite(bgr.data(), 1, bgr.size(), f); else std::fwrite(d.data(), 2, d.size(), f);
            std::fclose(f);
            std::printf("%d %d\n", rows, cols);
            return 0;
        }
        if (cmd == "traj" && argc == 3) {
            std::ifstream in(argv[2]);
            Trajectory tr;
            unsigned long long t_us;
            double ts_obs;
            float ddt, T[16];
            while (in >> t_us >> ts_obs >> ddt) {
                for (int i = 0; i < 16; i++) in >> T[i];
                tr.fuse(T, t_us);
                tr.writeTrajectoryLine(std::cout, ts_obs, ddt);
            }
            std::cout << "--\n";
            tr.writeFreiburg(std::cout);
            return 0;
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 2;
}
