// sf_k_pyramid.cu - per-pair state initialisation and the image pyramid (createImagePyramid)
// Part of the sm_100a kernels of the StaticFusion joint odometry + segmentation solver (launch interface: sf_kernels.cuh).
// One launch of each kernel serves the whole batch of frame pairs; data-dependent exits (IRLS convergence FrontEnd.cpp:679,
// outer-loop exit :1130, k-means :227) are per-pair flags in PairCtl that later launches test, so the host enqueues a static
// schedule with no synchronisation.  Compiled with -fmad=false: float expressions keep the reference's operation order and
// rounding; fused multiply-adds appear only where written explicitly.  Reference citations are relative to the upstream tree.
#include <cstdlib>

#include "sf_common.cuh"

namespace sf {

// ------------------------------------------------------------------------------------------
// init: reset the per-pair control blocks (runSolver prologue, FrontEnd.cpp:1091)
// ------------------------------------------------------------------------------------------
__global__ void init_pairs_kernel(Arena a, const float* twist_old_in, int n_pairs) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < MAX_WORK_CTRS; i += gridDim.x * blockDim.x) a.work_ctr[i] = 0;
    const int pair = blockIdx.x * blockDim.x + threadIdx.x;
    if (pair >= n_pairs) return;
    PairCtl& c = a.ctl[pair];
    for (int i = 0; i < 16; i++) c.T[i] = (i % 5 == 0) ? 1.f : 0.f;
    for (int i = 0; i < 12; i++) c.Tinv[i] = (i % 5 == 0) ? 1.f : 0.f;
    for (int i = 0; i < 6; i++) {
        c.twist_old[i] = twist_old_in ? twist_old_in[pair * 6 + i] : 0.f;
        c.twist_odom[i] = 0.f; c.twist_level[i] = 0.f; c.var[i] = 0.f; c.prev_sol[i] = 0.f;
    }
    for (int i = 0; i < 36; i++) c.AtA[i] = 0.0;
    c.res_sq = 0.0;
    for (int l = 0; l < NC; l++) {
        c.b_segm[l] = 0.5f; c.b_prior[l] = 0.f; c.lambda_t_w[l] = 0.f;  // FrontEnd.cpp:156
        c.conn[l] = 1u << l;                                            // KMeans.cpp:311
        c.prior_fix[l] = 0; c.csize[l] = 0; c.cnonnull[l] = 0; c.lab_fix[l] = 0; c.lab_cnt[l] = 0;
        for (int k = 0; k < 3; k++) c.kmeans[k * NC + l] = 0.f;
    }
    c.max_wc_bits = 0; c.max_wd_bits = 0; c.fixBc = 0; c.fixBd = 0; c.n_valid = 0;
    for (int q = 0; q < 7; q++) { c.colmax_c[q] = 0; c.colmax_d[q] = 0; c.colbound[q] = 0.f; c.sexp[q] = 0; }
    for (int q = 0; q < 27; q++) c.acc_ne[q] = 0;
    c.acc_rs = 0; c.rexp = 0;
    c.inv_max_c = 0.f; c.inv_max_d = 0.f; c.aver_res = 0.f; c.aver_res_old = 0.f;
    c.active = 0; c.irls_done = 1; c.break_level = -1; c.it_done = 0; c.status = 0; c.total_irls = 0;
    c.ticket1 = 0; c.ticket2 = 0;
    for (int i = 0; i < 2 * a.trace_steps; i++) a.stepstat[(size_t)pair * 2 * a.trace_steps + i] = 0;
    if (a.trace) {
        float* t = a.trace + (size_t)pair * a.trace_steps * SF_TRACE_STEP;
        for (int i = 0; i < a.trace_steps * SF_TRACE_STEP; i++) t[i] = 0.f;
    }
}

// ------------------------------------------------------------------------------------------
// K1: pyramid level from its parent (createImagePyramid, FrontEnd.cpp:294-375)
// ------------------------------------------------------------------------------------------
// One thread makes the two horizontally adjacent outputs u = 2t - 1 and u = 2t of a row: their 4x4 source blocks lie inside
// the eight source columns 4t - 4 .. 4t + 3, i.e. two aligned float4 loads per source row and image (16 wide loads per two
// outputs instead of 64 scalar ones).  The per-output arithmetic and its order are the reference's.
__device__ __forceinline__ void pyr_interior(const float (&D)[4][8], const float (&I)[4][8], int o, float& out_d, float& out_i) {
    const float max_depth_dif = 0.1f;
    float db[16], ib[16];  // column-major 4x4 block at (v2-1,u2-1), :308-309
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) { db[r + 4 * c] = D[r][o + c]; ib[r + 4 * c] = I[r][o + c]; }
    float d0 = db[5], d1 = db[6], d2 = db[9], d3 = db[10];  // :311
    if (d1 < d0) { const float t = d1; d1 = d0; d0 = t; }
    if (d3 < d2) { const float t = d3; d3 = d2; d2 = t; }
    const float dcenter = (d3 < d1) ? fmaxf(d3, d0) : fmaxf(d1, d2);
    const float vm[4] = {1.f, 2.f, 2.f, 1.f};
    if (dcenter != 0.f) {
        float sum_d = 0.f, sum_c = 0.f, weight = 0.f;
#pragma unroll
        for (int k = 0; k < 16; k++) {  // :323-333
            const float cm = vm[k & 3] * vm[k >> 2] / 36.f;
            const float abs_dif = fabsf(db[k] - dcenter);
            if (abs_dif < max_depth_dif) {
                const float aux_w = cm * (max_depth_dif - abs_dif);
                weight += aux_w;
                sum_d += aux_w * db[k];
                sum_c += aux_w * ib[k];
            }
        }
        out_d = sum_d / weight;
        out_i = sum_c / weight;
    } else {  // :339-343
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 16; k++) s += (vm[k & 3] * vm[k >> 2] / 36.f) * ib[k];
        out_i = s;
        out_d = 0.f;
    }
}
__device__ __forceinline__ void pyr_border(const float (&D)[4][8], const float (&I)[4][8], int o, float& out_d, float& out_i) {
    // boundary, :347-373, 2x2 block at (v2,u2) in column-major order; source rows v2, v2+1 are D[1], D[2]
    const float d4[4] = {D[1][o], D[2][o], D[1][o + 1], D[2][o + 1]};
    const float i4[4] = {I[1][o], I[2][o], I[1][o + 1], I[2][o + 1]};
    out_i = 0.25f * (((i4[0] + i4[1]) + i4[2]) + i4[3]);
    float new_d = 0.f;
    unsigned cont = 0;
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (d4[k] != 0.f) { new_d += d4[k]; cont++; }
    out_d = cont ? new_d / float(cont) : 0.f;
}
__global__ void __launch_bounds__(256, 2) pyr_down_kernel(Arena a, LevelGeom gs, LevelGeom gd, int threads_per_row) {
    const int frame = blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = q / threads_per_row, t = q - v * threads_per_row;  // threads_per_row = cols/2 + 1
    if (v >= gd.rows) return;
    const float* ds = a.pyr_d + (size_t)frame * a.pyr_stride + gs.off;
    const float* is = a.pyr_i + (size_t)frame * a.pyr_stride + gs.off;
    float* dd = a.pyr_d + (size_t)frame * a.pyr_stride + gd.off + (size_t)v * gd.cols;
    float* id = a.pyr_i + (size_t)frame * a.pyr_stride + gd.off + (size_t)v * gd.cols;
    const bool row_in = (v > 0) && (v < gd.rows - 1);
    const bool lo_ok = t >= 1, hi_ok = 2 * t < gd.cols;  // the two float4 column groups that exist
    float D[4][8], I[4][8];
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const bool row_ok = row_in || r == 1 || r == 2;  // boundary rows only need source rows 2v, 2v+1
        const size_t o = (size_t)(2 * v - 1 + r) * gs.cols + (size_t)(4 * t) - 4;
        const float4 d_lo = (row_ok && lo_ok) ? ldg4(ds + o) : z4, d_hi = (row_ok && hi_ok) ? ldg4(ds + o + 4) : z4;
        const float4 i_lo = (row_ok && lo_ok) ? ldg4(is + o) : z4, i_hi = (row_ok && hi_ok) ? ldg4(is + o + 4) : z4;
        D[r][0] = d_lo.x; D[r][1] = d_lo.y; D[r][2] = d_lo.z; D[r][3] = d_lo.w; D[r][4] = d_hi.x; D[r][5] = d_hi.y; D[r][6] = d_hi.z; D[r][7] = d_hi.w;
        I[r][0] = i_lo.x; I[r][1] = i_lo.y; I[r][2] = i_lo.z; I[r][3] = i_lo.w; I[r][4] = i_hi.x; I[r][5] = i_hi.y; I[r][6] = i_hi.z; I[r][7] = i_hi.w;
    }
    if (lo_ok) {  // output u = 2t - 1: block columns 4t-3 .. 4t (array 1..4), 2x2 block columns 4t-2, 4t-1 (array 2, 3)
        const int u = 2 * t - 1;
        float od, oi;
        if (row_in && (u > 0) && (u < gd.cols - 1)) pyr_interior(D, I, 1, od, oi);
        else pyr_border(D, I, 2, od, oi);
        dd[u] = od; id[u] = oi;
    }
    if (hi_ok) {  // output u = 2t: block columns 4t-1 .. 4t+2 (array 3..6), 2x2 block columns 4t, 4t+1 (array 4, 5)
        const int u = 2 * t;
        float od, oi;
        if (row_in && (u > 0) && (u < gd.cols - 1)) pyr_interior(D, I, 3, od, oi);
        else pyr_border(D, I, 4, od, oi);
        dd[u] = od; id[u] = oi;
    }
}

int launch_init_pairs(const Arena& a, const DevParams&, const float* twist_old_in_dev, const LaunchCfg& c) {
    init_pairs_kernel<<<cdiv(c.n_pairs, 64), 64, 0, c.stream>>>(a, twist_old_in_dev, c.n_pairs);
    return 1;
}

int launch_pyramids(const Arena& a, const LevelGeom* geom, int levels, const LaunchCfg& c) {
    int n = 0;
    for (int l = 1; l < levels; l++) {
        const int tpr = geom[l].cols / 2 + 1;
        pyr_down_kernel<<<dim3(cdiv((size_t)tpr * geom[l].rows, 256), c.n_frames), 256, 0, c.stream>>>(a, geom[l - 1], geom[l], tpr);
        n++;
    }
    return n;
}
}  // namespace sf
