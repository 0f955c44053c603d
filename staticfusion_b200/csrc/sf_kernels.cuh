// sf_kernels.cuh — launch interface between the C ABI (sf_api.cu) and the kernels (sf_k_*.cu).
#pragma once
#include "sf_device.cuh"

namespace sf {

// compact per-pair result block copied back to the host
struct PairOut {
    float T[16];        // row-major
    float twist_old[6]; // after FrontEnd.cpp:1143-1144
    float b_segm[NC];
    int irls_iters;
    int status;
};

// device arena: base pointers + strides, passed to kernels by value
struct Arena {
    // image pyramids, one per frame: [n_frames][pyr_stride]
    float* pyr_d;
    float* pyr_i;
    size_t pyr_stride;  // pixels per pyramid (sum over levels)
    // pair -> frame indices
    const int* cur_idx;
    const int* pred_idx;
    // per pair
    uint8_t* labels;            // [F][pyr_stride] cluster labels, 24 = no depth
    const uint8_t* seed_map;    // [P1] nearest k-means seed of every level-1 pixel (KMeans.cpp:87-101): a function of the image size only
    long long* acc_d;           // [F][P0] warp depth accumulator (fixed point 2^32)
    unsigned long long* acc_iw; // [F][P0] packed: weight sum (high 22 bits) | intensity sum (low 42 bits, 2^22)
    float* warp_d;              // [F][P0]
    float* warp_i;              // [F][P0]
    uint8_t* tiles;             // [F][tiles_per_pair(P0)][TILE_BYTES] raw Jacobian rows + valid-pixel labels
    float* dbg;                 // [F][NPLANES][P0] linearisation planes, only with the trace flag (else nullptr)
    int* gcount;                // [GCOUNT_CELLS]: [0] pairs active in the current step, [1] pairs still inside the IRLS loop, [2] / [3] lengths of
                                // iter_list0 / iter_list1, [4] generation of the IRLS item queue (one per step), [5] / [6] its head / tail, [7] blocks alive in the loop kernel
    unsigned long long* irls_q; // [q_cap] item queue of irls_loop_kernel: (generation << 32) | (pair << 12) | (pass << 11) | chunk
    int q_cap;
    int* iter_list0;            // [F] pairs that run IRLS iteration it (odd it); pass 2 of iteration it appends the pairs that go on
    int* iter_list1;            // [F] ... to the other list (even it), so every pass launch is sized by the pairs still iterating
    int* active_list;           // [F] indices of the pairs active in the current step (first gcount[0] entries)
    int* work_ctr;              // [MAX_WORK_CTRS] dynamic item counters, one per pass launch of a solve (zeroed by init_pairs)
    PairCtl* ctl;               // [F]
    PairOut* out;               // [F]
    float* b_perpixel;          // [F][P0]
    float* pcar;                // [F][NC] perClusterAverageResidual (NaN until the 5-frame history ran, FrontEnd.cpp:105)
    float* ring_d;              // [5][P0] depthBuffer     (StaticFusion.h:94), drop-in path only
    float* ring_i;              // [5][P0] intensityBuffer
    float* ring_T;              // [5][16] odomBuffer, row-major
    float* trace;               // [F][steps][SF_TRACE_STEP] or nullptr
    int* stepstat;              // [F][steps][2]: valid pixels, IRLS iterations
    size_t P0;                  // pixels of level 0
    int max_blocks;
    int num_sms;
    int trace_steps;
};

constexpr int MAX_WORK_CTRS = 4096;
constexpr int GCOUNT_CELLS = 8;

struct LaunchCfg {
    cudaStream_t stream;
    int n_pairs;
    int n_frames;
    int* next_ctr;  // host-side running index into Arena::work_ctr for this solve
};

void prepare_kernels();     // one-time kernel attribute setup (dynamic shared memory sizes); call outside stream capture
int irls_chunk_iters(int P);  // deterministic function of the level size only

// every launcher returns the number of kernels it enqueued
int launch_init_pairs(const Arena& a, const DevParams& p, const float* twist_old_in_dev, const LaunchCfg& c);
int launch_pyramids(const Arena& a, const LevelGeom* geom, int levels, const LaunchCfg& c);
int launch_kmeans(const Arena& a, const DevParams& p, const LevelGeom* geom, int levels, const LaunchCfg& c);
int launch_step_begin(const Arena& a, int level_i, int k, const LaunchCfg& c);
int launch_warp(const Arena& a, const LevelGeom& g, const LaunchCfg& c);
int launch_linearise(const Arena& a, const DevParams& p, const LevelGeom& g, int first, const LaunchCfg& c);
int launch_step_prep(const Arena& a, const DevParams& p, int level_i, int k, const LaunchCfg& c);
int launch_irls_pass1(const Arena& a, const DevParams& p, const LevelGeom& g, int level_i, int k, int it, const LaunchCfg& c);
int launch_irls_pass2(const Arena& a, const DevParams& p, const LevelGeom& g, int level_i, int k, int it, const LaunchCfg& c);
int launch_irls_fused(const Arena& a, const DevParams& p, const LevelGeom& g, int level_i, int k, const LaunchCfg& c);
int launch_irls_loop(const Arena& a, const DevParams& p, const LevelGeom& g, int level_i, int k, const LaunchCfg& c);
size_t irls_queue_capacity(int max_pairs, int max_iter_irls, size_t P0);  // items one IRLS loop of a lane can push
int launch_pose_update(const Arena& a, const DevParams& p, int level_i, int k, const LaunchCfg& c);
int launch_finish(const Arena& a, const DevParams& p, const LevelGeom& g0, const LaunchCfg& c);
// computeResidualsAgainstPreviousImage (FrontEnd.cpp:896-1069).  mode 0: pairs of a sequence, pair p >= 4 warps frame
// cur_idx[p]-5 with the increments of pairs p-4..p; mode 1: pair 0 against the ring buffers with the driver's im_count = index.
int launch_history(const Arena& a, const DevParams& p, const LevelGeom& g0, int mode, int index, const LaunchCfg& c);
// Reconstruction::getFilteredDepth: n images of rows x cols u16 millimetres -> float metres (strides in elements)
int launch_filter_depth(const uint16_t* in, float* out, int rows, int cols, int n, size_t in_stride, size_t out_stride, float max_depth_m,
                        cudaStream_t stream);
// StaticFusion::loadImageFromSequenceAssoc, conversion half: n decoded full-resolution images -> rows x cols outputs
// (f32_stride = elements between consecutive intensity / depth images, so they can land in the pyramids' level-0 slots)
int launch_convert_frames(const uint8_t* bgr, const uint16_t* depth_raw, int rows, int cols, int res_factor, int n, float* intensity, float* depth,
                          size_t f32_stride, uint16_t* depth_mm, uint8_t* color, cudaStream_t stream);
int launch_segm_image(const Arena& a, const LevelGeom& g0, const LaunchCfg& c);  // buildSegmImage

}  // namespace sf
