// sf_k_warp.cu - per-step control, forward splat of the prediction (warpImagesAccurateInverse), 5-frame history residuals
// Part of the sm_100a kernels of the StaticFusion joint odometry + segmentation solver (launch interface: sf_kernels.cuh).
// One launch of each kernel serves the whole batch of frame pairs; data-dependent exits (IRLS convergence FrontEnd.cpp:679,
// outer-loop exit :1130, k-means :227) are per-pair flags in PairCtl that later launches test, so the host enqueues a static
// schedule with no synchronisation.  Compiled with -fmad=false: float expressions keep the reference's operation order and
// rounding; fused multiply-adds appear only where written explicitly.  Reference citations are relative to the upstream tree.
#include <cstdlib>

#include "sf_common.cuh"

namespace sf {

// ------------------------------------------------------------------------------------------
// per-step control
// ------------------------------------------------------------------------------------------
// One block: resets the two global work counters ([0] pairs active in the step, [1] pairs inside the IRLS loop), decides
// which pairs take part in the step and compacts their indices into Arena::active_list, so that the per-pixel kernels
// of the step are sized by the ACTIVE pairs (steps that no pair needs any more cost one near-empty launch each).
constexpr int SB_THREADS = 1024;
__global__ void __launch_bounds__(SB_THREADS) step_begin_kernel(Arena a, int level_i, int n_pairs) {
    __shared__ int s_count;
    if (threadIdx.x == 0) {
        s_count = 0; a.gcount[1] = 0; a.gcount[2] = 0; a.gcount[3] = 0;
        a.gcount[4] += 1; a.gcount[5] = 0; a.gcount[6] = 0; a.gcount[7] = 0;  // IRLS item queue of this step: new generation, empty, no block alive
    }
    __syncthreads();
    for (int pair = threadIdx.x; pair < n_pairs; pair += SB_THREADS) {
        PairCtl& c = a.ctl[pair];
        const int active = (c.break_level != level_i) ? 1 : 0;  // FrontEnd.cpp:1130 leaves the k-loop of this level only
        c.active = active;
        c.irls_done = 1;
        if (!active) continue;
        c.max_wc_bits = 0; c.max_wd_bits = 0; c.fixBc = 0; c.fixBd = 0; c.n_valid = 0;
        for (int l = 0; l < NC; l++) { c.prior_fix[l] = 0; c.csize[l] = 0; c.cnonnull[l] = 0; }
        for (int q = 0; q < 7; q++) { c.colmax_c[q] = 0; c.colmax_d[q] = 0; }
        a.active_list[atomicAdd(&s_count, 1)] = pair;  // any order: every cross-pixel sum is an integer sum
    }
    __syncthreads();
    if (threadIdx.x == 0) a.gcount[0] = s_count;
}

// ------------------------------------------------------------------------------------------
// K4: forward splat of the prediction into the current view (warpImagesAccurateInverse,
// FrontEnd.cpp:775-871).  Integer weights; depth and intensity sums are fixed-point integers
// so the atomics commute and the result is deterministic.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void splat(long long* acc_d, unsigned long long* acc_iw, int idx, int w, long long qd, long long qi) {
    atomic_add_ll(acc_d + idx, (long long)w * qd);
    atomicAdd(acc_iw + idx, ((unsigned long long)w << 42) + (unsigned long long)((long long)w * qi));
}

// transform one source point and splat it into the 1-4 surrounding pixels (FrontEnd.cpp:814-867 and :963-1014)
__device__ __forceinline__ void splat_point(long long* acc_d, unsigned long long* acc_iw, const LevelGeom& g, const float* T,
                                            float xr, float yr, float z, float intensity_w) {
    const float x_w = T[0] * xr + T[1] * yr + T[2] * z + T[3];  // :814-816
    const float y_w = T[4] * xr + T[5] * yr + T[6] * z + T[7];
    const float depth_w = T[8] * xr + T[9] * yr + T[10] * z + T[11];
    const float fu = 100.f * (div_rn_guarded(g.f * x_w, depth_w) + g.disp_u);  // :819-820
    const float fv = 100.f * (div_rn_guarded(g.f * y_w, depth_w) + g.disp_v);
    if (!(fabsf(fu) < 1.0e9f) || !(fabsf(fv) < 1.0e9f)) return;  // non-finite / out of int range = out of bounds
    const int uwarp = (int)fu, vwarp = (int)fv;
    const int cols_lim = 100 * (g.cols - 1), rows_lim = 100 * (g.rows - 1);
    if ((uwarp >= 0) && (uwarp < cols_lim) && (vwarp >= 0) && (vwarp < rows_lim)) {
        const int uwarp_l = uwarp - uwarp % 100;
        const int uwarp_r = uwarp_l + 100;
        const int vwarp_d = vwarp - vwarp % 100;
        const int vwarp_u = vwarp_d + 100;
        const int delta_r = uwarp_r - uwarp;
        const int delta_l = 100 - delta_r;
        const int delta_u = vwarp_u - vwarp;
        const int delta_d = 100 - delta_u;
        const long long qd = fixq(depth_w, FIX_WARP_D), qi = fixq(intensity_w, FIX_WARP_I);
        if (min(delta_r, delta_l) + min(delta_u, delta_d) < 5) {  // :835-843
            const int ind_u = delta_r > delta_l ? uwarp_l / 100 : uwarp_r / 100;
            const int ind_v = delta_u > delta_d ? vwarp_d / 100 : vwarp_u / 100;
            splat(acc_d, acc_iw, ind_v * g.cols + ind_u, 200, qd, qi);
        } else {  // :846-867
            const int v_d = vwarp_d / 100, u_l = uwarp_l / 100;
            const int v_u = v_d + 1, u_r = u_l + 1;
            splat(acc_d, acc_iw, v_u * g.cols + u_r, delta_l + delta_d, qd, qi);
            splat(acc_d, acc_iw, v_u * g.cols + u_l, delta_r + delta_d, qd, qi);
            splat(acc_d, acc_iw, v_d * g.cols + u_r, delta_l + delta_u, qd, qi);
            splat(acc_d, acc_iw, v_d * g.cols + u_l, delta_r + delta_u, qd, qi);
        }
    }
}

#ifndef SF_WARP_BPS
#define SF_WARP_BPS 6
#endif
__global__ void __launch_bounds__(256, SF_WARP_BPS) warp_kernel(Arena a, LevelGeom g, int chunks_per_pair) {
    int cur_slot = -1, pair = 0;
    const float* src_d = nullptr;
    const float* src_i = nullptr;
    float T[12];  // the pair's inverse pose stays in registers while the block walks through the pair's pixels
    // depth and intensity of an item are loaded together and ONE ITEM AHEAD (while the block stays inside a pair): the two
    // dependent DRAM round trips per item (depth, then intensity only where the depth is valid) were 42 % of the stall samples
    float z_n = 0.f, i_n = 0.f;
    bool have_n = false;
    for (ItemWalk it(a.gcount[0] * chunks_per_pair, chunks_per_pair); it.more(); it.next()) {
        if (it.slot != cur_slot) {
            cur_slot = it.slot;
            pair = a.active_list[cur_slot];
            const size_t fo = (size_t)a.pred_idx[pair] * a.pyr_stride + g.off;
            src_d = a.pyr_d + fo; src_i = a.pyr_i + fo;
#pragma unroll
            for (int q = 0; q < 12; q++) T[q] = a.ctl[pair].Tinv[q];
            have_n = false;
        }
        const int p = it.rem * 256 + threadIdx.x;
        float z = z_n, intensity_w = i_n;
        if (!have_n) {
            z = 0.f; intensity_w = 0.f;
            if (p < g.P) { z = __ldg(src_d + p); intensity_w = __ldg(src_i + p); }
        }
        have_n = (it.rem + 1 < chunks_per_pair) && (it.item + 1 < it.end);  // block-uniform: the next item is in the same pair
        if (have_n) {
            const int pn = p + 256;
            z_n = 0.f; i_n = 0.f;
            if (pn < g.P) { z_n = __ldg(src_d + pn); i_n = __ldg(src_i + pn); }
            if ((threadIdx.x & 7) == 0 && pn + 256 < g.P) { prefetch_l2(src_d + pn + 256); prefetch_l2(src_i + pn + 256); }  // the sectors of the item after
        }
        if (p >= g.P || z == 0.f) continue;
        int i, j;
        split_rc(p, g, i, j);
        const float xr = (g.inv_f * (float(j) - g.disp_u)) * z;  // xxPredPyr, FrontEnd.cpp:386
        const float yr = (g.inv_f * (float(i) - g.disp_v)) * z;
        splat_point(a.acc_d + (size_t)pair * a.P0, a.acc_iw + (size_t)pair * a.P0, g, T, xr, yr, z, intensity_w);
    }
}

// K4b: divide by the accumulated weight (FrontEnd.cpp:875-891) and clear the accumulators for the next splat
// Two adjacent pixels per thread: both accumulators come in with independent 16-byte loads (P is even on every level)
__global__ void __launch_bounds__(256) warp_normalise_kernel(Arena a, LevelGeom g, int chunks_per_pair) {
    for (ItemWalk it(a.gcount[0] * chunks_per_pair, chunks_per_pair); it.more(); it.next()) {
        const int pair = a.active_list[it.slot];
        const int p = (it.rem * 256 + threadIdx.x) * 2;
        if (p >= g.P) continue;
        const size_t o = (size_t)pair * a.P0 + p;
        const ulonglong2 iw2 = *reinterpret_cast<const ulonglong2*>(a.acc_iw + o);
        const longlong2 dq2 = *reinterpret_cast<const longlong2*>(a.acc_d + o);
        const unsigned long long iw[2] = {iw2.x, iw2.y};
        const long long dq[2] = {dq2.x, dq2.y};
        float dw[2] = {0.f, 0.f}, iwv[2] = {0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const unsigned w = (unsigned)(iw[j] >> 42);
            const long long iq = (long long)(iw[j] & ((1ull << 42) - 1ull));
            if (w != 0u) {
                iwv[j] = (float)((double)iq / ((double)w * 4194304.0));
                dw[j] = (float)((double)dq[j] / ((double)w * 4294967296.0));
            }
        }
        if ((iw[0] | iw[1]) != 0ull) {
            *reinterpret_cast<ulonglong2*>(a.acc_iw + o) = make_ulonglong2(0ull, 0ull);
            *reinterpret_cast<longlong2*>(a.acc_d + o) = make_longlong2(0ll, 0ll);
        }
        *reinterpret_cast<float2*>(a.warp_d + o) = make_float2(dw[0], dw[1]);
        *reinterpret_cast<float2*>(a.warp_i + o) = make_float2(iwv[0], iwv[1]);
    }
}

// ------------------------------------------------------------------------------------------
// K8: 5-frame history residuals (computeResidualsAgainstPreviousImage, FrontEnd.cpp:896-1069)
// ------------------------------------------------------------------------------------------
// T = (prod of the four previous increments * T_odometry)^-1 (:901-909): products in double of the float increments,
// rounded once, rigid inverse in double.
__global__ void hist_pose_kernel(Arena a, int mode, int index, int n_pairs) {
    const int pair = blockIdx.x * blockDim.x + threadIdx.x;
    if (pair >= n_pairs) return;
    PairCtl& c = a.ctl[pair];
    const int cur = a.cur_idx[pair];
    const bool on = (mode == 1) ? (pair == 0) : (pair >= 4 && cur >= 5);
    c.hist_on = on ? 1 : 0;
    for (int l = 0; l < NC; l++) { c.hist_sum[l] = 0; c.hist_cnt[l] = 0; }
    if (!on) {
        if (mode == 0) for (int l = 0; l < NC; l++) a.pcar[pair * NC + l] = __int_as_float(0x7fc00000);
        return;
    }
    c.hist_ref = (mode == 1) ? (index % 5) : (cur - 5);
    double M[16], tmp[16];
    for (int i = 0; i < 16; i++) M[i] = (i % 5 == 0) ? 1.0 : 0.0;
    for (int j = 0; j <= 4; j++) {
        const float* B;
        if (j == 4) B = a.out[pair].T;
        else B = (mode == 1) ? (a.ring_T + 16 * ((index - 4 + j) % 5)) : a.out[pair - 4 + j].T;
        for (int r = 0; r < 4; r++)
            for (int q = 0; q < 4; q++) {
                double acc = 0.0;
                for (int k = 0; k < 4; k++) acc += M[r * 4 + k] * (double)B[k * 4 + q];
                tmp[r * 4 + q] = acc;
            }
        for (int i = 0; i < 16; i++) M[i] = tmp[i];
    }
    float Mf[16];
    for (int i = 0; i < 16; i++) Mf[i] = (float)M[i];
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) c.Thist[i * 4 + j] = Mf[j * 4 + i];
        double sd = 0.0;
        for (int j = 0; j < 3; j++) sd += (double)Mf[j * 4 + i] * (double)Mf[j * 4 + 3];
        c.Thist[i * 4 + 3] = (float)(-sd);
    }
}

// forward splat of the frame of five frames ago into the current view (:946-1021)
__global__ void __launch_bounds__(256) hist_warp_kernel(Arena a, LevelGeom g, const float* ref_d, const float* ref_i, size_t ref_stride) {
    const int pair = blockIdx.y;
    const PairCtl& c = a.ctl[pair];
    if (!c.hist_on) return;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= g.P) return;
    const size_t ro = (size_t)c.hist_ref * ref_stride + p;
    const float z = __ldg(ref_d + ro);
    const float dcur = __ldg(a.pyr_d + (size_t)a.cur_idx[pair] * a.pyr_stride + p);
    if (z == 0.f || dcur == 0.f) return;  // :951 tests the CURRENT depth at the source pixel
    const float intensity_w = __ldg(ref_i + ro);
    const int i = p / g.cols, j = p - i * g.cols;
    const float xr = (g.inv_f * (float(j) - g.disp_u)) * z;  // :925-929
    const float yr = (g.inv_f * (float(i) - g.disp_v)) * z;
    splat_point(a.acc_d + (size_t)pair * a.P0, a.acc_iw + (size_t)pair * a.P0, g, c.Thist, xr, yr, z, intensity_w);
}

// normalise (:1025-1035), residuals and per-cluster sums (:1039-1066); clears the accumulators
__global__ void __launch_bounds__(256) hist_reduce_kernel(Arena a, DevParams prm, LevelGeom g, const float* ref_d, size_t ref_stride) {
    const int pair = blockIdx.y;
    PairCtl& c = a.ctl[pair];
    if (!c.hist_on) return;
    __shared__ long long bins[8][NC];
    __shared__ int cnts[8][NC];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 8 * NC; i += 256) { (&bins[0][0])[i] = 0; (&cnts[0][0])[i] = 0; }
    __syncthreads();
    const int p = blockIdx.x * blockDim.x + tid;
    int lab = -1;
    long long q = 0;
    if (p < g.P) {
        const size_t o = (size_t)pair * a.P0 + p;
        const unsigned long long iw = a.acc_iw[o];
        float dw = 0.f, iwv = 0.f;
        if (iw != 0ull) {
            const long long dq = a.acc_d[o];
            const unsigned w = (unsigned)(iw >> 42);
            const long long iq = (long long)(iw & ((1ull << 42) - 1ull));
            if (w != 0u) {
                iwv = (float)((double)iq / ((double)w * 4194304.0));
                dw = (float)((double)dq / ((double)w * 4294967296.0));
            }
            a.acc_iw[o] = 0ull;
            a.acc_d[o] = 0ll;
        }
        a.warp_d[o] = dw;  // depthWarpedRefference / intensityWarpedRefference (StaticFusion.h:98-99)
        a.warp_i[o] = iwv;
        const size_t co = (size_t)a.cur_idx[pair] * a.pyr_stride + p;
        const float dcur = __ldg(a.pyr_d + co);
        const float zref = __ldg(ref_d + (size_t)c.hist_ref * ref_stride + p);
        const float idiff = (zref != 0.f && dcur != 0.f) ? __ldg(a.pyr_i + co) : 0.f;  // :939, :1018-1020
        const float dr = dcur - dw;
        const float ir = idiff - iwv;
        const float cr = fabsf(dr) + prm.k_photometric_res * fabsf(ir);  // :1041
        if (dw != 0.f && dcur != 0.f) {  // :1049
            lab = a.labels[(size_t)pair * a.pyr_stride + p];
            q = fixq(cr, 32);
        }
    }
    warp_group_add(lab, q, bins[warp], cnts[warp], lane);
    __syncthreads();
    if (tid < NC) {
        long long sacc = 0;
        int n = 0;
        for (int w = 0; w < 8; w++) { sacc += bins[w][tid]; n += cnts[w][tid]; }
        if (n) { atomic_add_ll(&c.hist_sum[tid], sacc); atomicAdd(&c.hist_cnt[tid], n); }
    }
}

// :1045-1046, :1068: counts start at 1, mean over 2*count; clusters without a pixel stay NaN
__global__ void hist_final_kernel(Arena a, int n_pairs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pairs * NC) return;
    const int pair = i / NC, l = i - pair * NC;
    const PairCtl& c = a.ctl[pair];
    if (!c.hist_on) return;
    const int n = c.hist_cnt[l];
    const float sum = n > 0 ? (float)fixval(c.hist_sum[l], 32) : __int_as_float(0x7fc00000);
    a.pcar[i] = sum / float(2 * (n + 1));
}

int launch_step_begin(const Arena& a, int level_i, int, const LaunchCfg& c) {
    step_begin_kernel<<<1, SB_THREADS, 0, c.stream>>>(a, level_i, c.n_pairs);
    return 1;
}

int launch_warp(const Arena& a, const LevelGeom& g, const LaunchCfg& c) {
    const int cpp = (int)cdiv(g.P, 256);
    const size_t items = (size_t)cpp * c.n_pairs, cap = (size_t)a.num_sms * 16;  // 8 resident blocks per SM, two rounds
    const unsigned grid = (unsigned)(items < cap ? items : cap);
    warp_kernel<<<grid, 256, 0, c.stream>>>(a, g, cpp);
    const int cpp2 = (int)cdiv(g.P, 512);  // the normalisation takes two pixels per thread
    const size_t items2 = (size_t)cpp2 * c.n_pairs;
    warp_normalise_kernel<<<(unsigned)(items2 < cap ? items2 : cap), 256, 0, c.stream>>>(a, g, cpp2);
    return 2;
}

int launch_history(const Arena& a, const DevParams& p, const LevelGeom& g0, int mode, int index, const LaunchCfg& c) {
    const float* ref_d = mode == 1 ? a.ring_d : a.pyr_d;
    const float* ref_i = mode == 1 ? a.ring_i : a.pyr_i;
    const size_t stride = mode == 1 ? a.P0 : a.pyr_stride;
    const dim3 grid(cdiv(g0.P, 256), c.n_pairs);
    hist_pose_kernel<<<cdiv(c.n_pairs, 64), 64, 0, c.stream>>>(a, mode, index, c.n_pairs);
    hist_warp_kernel<<<grid, 256, 0, c.stream>>>(a, g0, ref_d, ref_i, stride);
    hist_reduce_kernel<<<grid, 256, 0, c.stream>>>(a, p, g0, ref_d, stride);
    hist_final_kernel<<<cdiv((size_t)c.n_pairs * NC, 128), 128, 0, c.stream>>>(a, c.n_pairs);
    return 4;
}
}  // namespace sf
