// Host-side check of the row-tile layout contract (staticfusion_b200/csrc/sf_device.cuh): compiled with nvcc, runs on the CPU.
// The IRLS kernels read tiles by POSITION and only form order-independent sums; linearise_kernel writes by PIXEL through
// tile_row_off / tile_label_off.  What both sides rely on:
//   1. tile_pos is a permutation of 0..63 that keeps every 16-pixel group in place (a consumer warp's 32 positions are 32
//      consecutive pixels);
//   2. the pixel pairs (4l, 4l+1) and (4l+2, 4l+3) a thread stores as one float2 / uchar2 are adjacent and 8-byte aligned;
//   3. the four threads of a 16-pixel group fill one whole 32-byte sector of a plane with each of their two stores;
//   4. rows and labels of all pixels of a level tile the record without overlap.
#include <cstdio>
#include <cstdint>
#include <set>
#include "sf_device.cuh"

using namespace sf;

int main() {
    int fails = 0;
    auto check = [&](bool ok, const char* what) { if (!ok) { std::printf("FAIL: %s\n", what); fails++; } };
    std::set<int> seen;
    for (int p = 0; p < ROW_TILE; p++) {
        const int q = tile_pos(p);
        check(q >= 0 && q < ROW_TILE, "position in range");
        check(q / 16 == p / 16, "16-pixel group keeps its place");
        seen.insert(q);
    }
    check((int)seen.size() == ROW_TILE, "permutation");
    for (int tile = 0; tile < 3; tile++)
        for (int l = 0; l < ROW_TILE / 4; l++) {
            const int p0 = tile * ROW_TILE + 4 * l;
            for (int h = 0; h < 4; h += 2) {
                check(tile_row_off(0, p0 + h + 1) == tile_row_off(0, p0 + h) + 4, "float2 pair adjacent");
                check(tile_row_off(0, p0 + h) % 8 == 0, "float2 pair aligned");
                check(tile_label_off(p0 + h + 1) == tile_label_off(p0 + h) + 1, "label pair adjacent");
                check(tile_label_off(p0 + h) % 2 == 0, "label pair aligned");
            }
        }
    for (int k = 0; k < NROWPL; k++)
        for (int grp = 0; grp < ROW_TILE / 16; grp++)
            for (int h = 0; h < 4; h += 2) {  // one store instruction of the four threads of the group
                std::set<size_t> sectors, bytes;
                for (int l = 0; l < 4; l++) {
                    const size_t o = tile_row_off(k, grp * 16 + 4 * l + h);
                    sectors.insert(o / 32);
                    for (int b = 0; b < 8; b++) bytes.insert(o + b);
                }
                check(sectors.size() == 1 && bytes.size() == 32, "one whole sector per store of a 4-thread group");
            }
    const int P = 300;  // a level with a partial last tile
    std::set<size_t> used;
    for (int p = 0; p < (int)tiles_per_pair(P) * ROW_TILE; p++) {
        for (int k = 0; k < NROWPL; k++)
            for (int b = 0; b < 4; b++) check(used.insert(tile_row_off(k, p) + b).second, "row bytes unique");
        check(used.insert(tile_label_off(p)).second, "label byte unique");
    }
    check(used.size() == tiles_per_pair(P) * TILE_BYTES && *used.rbegin() == tiles_per_pair(P) * TILE_BYTES - 1, "records tiled exactly");
    check(TILE_BYTES % 32 == 0, "tile records start on sector boundaries");
    std::printf(fails ? "tile layout: %d failures\n" : "tile layout ok\n", fails);
    return fails ? 1 : 0;
}
